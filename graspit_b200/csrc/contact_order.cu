// K6b: visit keys of contact candidates.  Compiled with -fmad=false (see obb_order.h).
//
// removeContactDuplicates depends on the order in which the reference's dual-tree recursion appends the candidates
// (collisionInterface.cpp:254-287, collisionAlgorithms.cpp:42-130).  That order is a property of each candidate's own path
// through the two reference-style OBB trees: at every node pair on the path the rule picks the node to split and visits the
// child with the smaller bboxDistanceSq first, so "my child is visited second" is one bit per level, and the candidates of a
// (pose, pair) group, sorted by these bit strings (then by feature code inside a triangle pair), stand in visit order.  One
// thread per candidate walks its path (two box distances per level, fp64, the reference's operation order) and writes the
// bits as a 128-bit key; the host sorts each group by it.  The host replay (contacts_host.cpp, obb_visit_order) evaluates
// the same function only where a group's paths branch; it remains the fallback for paths longer than 128 levels.
#include <cub/cub.cuh>
#include "b2c_internal.h"
// the box distance stays out of line in the visit-key kernel: inlined twice per level it took 166 registers and the warps
// waited on instruction fetch (ncu r02i); out of line 128 registers and 6 % off the whole batched call
#define B2C_ORDER_DIST_ATTR __noinline__
#include "obb_order.h"

namespace b2c {

// Candidates of one triangle pair (up to 15 feature pairs) share their path: only the first of them in the group (the
// "leader") walks it.  Leaders are found by one pass over the group and appended to a dense list, so that the walking
// warps are full; the others copy their leader's key afterwards.
__global__ void __launch_bounds__(256) contact_leader_kernel(OrderParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const int q = p.recs[i].query;
  const int lo = p.offsets[q] - p.base;
  const int ta = p.recs[i].tri_a, tb = p.recs[i].tri_b;
  int leader = i;
  for (int j = lo; j < i; j++)
    if (p.recs[j].tri_a == ta && p.recs[j].tri_b == tb) { leader = j; break; }
  p.leader[i] = leader;
  if (leader == i) p.leaders[atomicAdd(p.nleaders, 1)] = i;
}

__global__ void __launch_bounds__(256) contact_follow_kernel(OrderParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  const int l = p.leader[i];
  if (l == i) return;
  p.keys[2 * (size_t)i] = p.keys[2 * (size_t)l];
  p.keys[2 * (size_t)i + 1] = p.keys[2 * (size_t)l + 1];
  p.depth[i] = p.depth[l];
}

__global__ void __launch_bounds__(128) contact_order_kernel(OrderParams p) {
  const int li = blockIdx.x * blockDim.x + threadIdx.x;
  if (li >= (p.nleaders ? *p.nleaders : p.n)) return;
  const int i = p.nleaders ? p.leaders[li] : li;
  const ContactRec &r = p.recs[i];
  const int q = r.query;
  const int lpose = q / p.np, pair = q - lpose * p.np;
  const int2 ab = p.pairs[pair];
  const double *X1 = ab.x < p.nrb ? reinterpret_cast<const double *>(p.fk + (size_t)lpose * p.nrb + ab.x)
                                  : reinterpret_cast<const double *>(p.static_xf + (ab.x - p.nrb));
  const double *X2 = ab.y < p.nrb ? reinterpret_cast<const double *>(p.fk + (size_t)lpose * p.nrb + ab.y)
                                  : reinterpret_cast<const double *>(p.static_xf + (ab.y - p.nrb));
  OrderXf x;
  order_rel_xf(X1, X2, x);
  const OrderNode *A = p.nodes[ab.x], *B = p.nodes[ab.y];
  const int *tpa = p.tri_pos[ab.x], *tpb = p.tri_pos[ab.y];
  const int pa = tpa[r.tri_a], pb = tpb[r.tri_b];
  // leaf positions spanned by the candidates of this (pose, pair) group: levels of the walk where the whole group sits
  // under one child decide nothing and cost nothing (most of the upper levels: a group is one contact patch)
  int range[4] = {pa, pa, pb, pb};
  if (p.offsets) {
    const int glo = p.offsets[q] - p.base, ghi = (q + 1 < p.nqueries ? p.offsets[q + 1] : p.total) - p.base;
    for (int j = glo; j < ghi; j++) {
      const int ja = tpa[p.recs[j].tri_a], jb = tpb[p.recs[j].tri_b];
      range[0] = min(range[0], ja); range[1] = max(range[1], ja);
      range[2] = min(range[2], jb); range[3] = max(range[3], jb);
    }
  }
  unsigned long long hi = 0ull, lo = 0ull;
  int depth = 0, n1 = 0, n2 = 0;
  // Every lane first runs down the levels that decide nothing for its group, then the lanes of the warp evaluate their
  // deciding level's box distances together (one lane per walk, and the walks reach their deciding levels at different
  // depths: taken level by level the distance code ran with 7 of 32 lanes, ncu r02h).
  while (true) {
    OrderPick k;
    int st;
    while ((st = order_pick(A, B, n1, n2, pa, pb, range, k)) == 1 && depth <= 4096) {   // (a malformed tree cannot hang the kernel)
      order_advance(n1, n2, k);
      depth++;
    }
    if (st != 2 || depth > 4096) break;
    if (k.mine != order_first(A, B, x, n1, n2, k)) {
      if (depth < 64) hi |= 1ull << (63 - depth);
      else if (depth < 128) lo |= 1ull << (127 - depth);
    }
    order_advance(n1, n2, k);
    depth++;
  }
  p.keys[2 * (size_t)i] = hi;
  p.keys[2 * (size_t)i + 1] = lo;
  p.depth[i] = depth;
}

// Rank of every candidate inside its (pose, pair) group by (visit key, feature code, arrival index): thread / candidate,
// one pass over the keys of its group (a group holds ~40 candidates, a few hold thousands).  perm[first of the group + rank]
// = the candidate's index inside the group, i.e. the group in the reference's visit order without a sort on the host.  A
// candidate whose path was longer than the key writes nothing: the slot left at -1 tells the host to replay that group.
__global__ void __launch_bounds__(256) contact_rank_kernel(OrderParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
  if (p.depth[i] > 128) return;
  const int q = p.recs[i].query;
  const int lo = p.offsets[q] - p.base, hi = (q + 1 < p.nqueries ? p.offsets[q + 1] : p.total) - p.base;
  const unsigned long long k0 = p.keys[2 * (size_t)i], k1 = p.keys[2 * (size_t)i + 1];
  const int f = p.recs[i].feature;
  int rank = 0;
  for (int j = lo; j < hi; j++) {
    const unsigned long long a0 = p.keys[2 * (size_t)j], a1 = p.keys[2 * (size_t)j + 1];
    const int fj = p.recs[j].feature;
    const bool less = a0 != k0 ? a0 < k0 : (a1 != k1 ? a1 < k1 : (fj != f ? fj < f : j < i));
    rank += less ? 1 : 0;
  }
  p.perm[lo + rank] = i - lo;
}

cudaError_t launch_contact_order(const OrderParams &p, cudaStream_t s) {
  if (p.n <= 0) return cudaSuccess;
  cudaError_t e;
  if (p.nleaders) {
    e = cudaMemsetAsync(p.nleaders, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    contact_leader_kernel<<<(p.n + 255) / 256, 256, 0, s>>>(p);
    // (the grid covers the worst case -- every candidate its own leader; threads past the list return at once)
    contact_order_kernel<<<(p.n + 127) / 128, 128, 0, s>>>(p);
    contact_follow_kernel<<<(p.n + 255) / 256, 256, 0, s>>>(p);
  } else {
    contact_order_kernel<<<(p.n + 127) / 128, 128, 0, s>>>(p);
  }
  e = cudaGetLastError();
  if (e != cudaSuccess || !p.perm) return e;
  e = cudaMemsetAsync(p.perm, 0xff, sizeof(int) * (size_t)p.n, s);
  if (e != cudaSuccess) return e;
  contact_rank_kernel<<<(p.n + 255) / 256, 256, 0, s>>>(p);
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------------------------------------------------
// K6d: CollisionInterface::removeContactDuplicates (collisionInterface.cpp:254-287) per (pose, pair) group on the device.
// The rule is sequential in the candidates' visit order: candidate i, if still alive, is compared with every later alive
// candidate j; the two are "the same contact" when both their body-1 and body-2 points lie within the threshold; then the
// one with the larger distSq goes -- j, and the scan goes on, or i (also on equal distSq), and the scan of i ends.  One warp
// per group: i runs sequentially, the j's are taken 32 at a time; inside a batch of 32 everything before the first j that
// stops i is applied, nothing after it -- exactly the sequential outcome.  Same fp64 expressions, in the same order, as the
// host's scan (contacts_api.inc), and this file is compiled without fused multiply-adds.  Alive flags: one bit per
// candidate in shared memory (groups of more than DEDUPE_MAX candidates are left to the host).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int DEDUPE_MAX = 4096;
constexpr int DEDUPE_WARPS = 4;

__global__ void __launch_bounds__(32 * DEDUPE_WARPS) contact_dedupe_kernel(DedupeParams p) {
  __shared__ unsigned alive_s[DEDUPE_WARPS][DEDUPE_MAX / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = p.q0 + blockIdx.x * DEDUPE_WARPS + warp;
  if (q >= p.q1) return;
  const int lo = p.offsets[q], hi = q + 1 < p.nqueries ? p.offsets[q + 1] : p.total;
  const int k = hi - lo;
  if (k <= 0) { if (lane == 0) p.nsurv[q] = 0; return; }
  int *perm = p.perm + lo;
  const ContactRec *recs = p.recs + lo;
  bool bad = k > DEDUPE_MAX;
  for (int i = lane; i < k && !bad; i += 32) bad = perm[i] < 0;
  if (__any_sync(0xffffffffu, bad)) { if (lane == 0) p.nsurv[q] = -1; return; }
  unsigned *alive = alive_s[warp];
  const int nwords = (k + 31) >> 5;
  for (int w = lane; w < nwords; w += 32) alive[w] = (w == nwords - 1 && (k & 31)) ? ((1u << (k & 31)) - 1u) : 0xffffffffu;
  __syncwarp();
  const double t2 = p.thresh_sq;
  for (int i = 0; i < k; i++) {
    if (!((alive[i >> 5] >> (i & 31)) & 1u)) continue;
    const double *ci = recs[perm[i]].c;
    const double ci0 = ci[0], ci1 = ci[1], ci2 = ci[2], ci3 = ci[3], ci4 = ci[4], ci5 = ci[5], di = ci[12];
    for (int jb = i + 1; jb < k; jb += 32) {
      const int j = jb + lane;
      bool kill = false, stop = false;
      if (j < k && ((alive[j >> 5] >> (j & 31)) & 1u)) {
        const double *cj = recs[perm[j]].c;
        const double d1 = (cj[0] - ci0) * (cj[0] - ci0) + (cj[1] - ci1) * (cj[1] - ci1) + (cj[2] - ci2) * (cj[2] - ci2);
        if (!(d1 > t2)) {
          const double d2 = (cj[3] - ci3) * (cj[3] - ci3) + (cj[4] - ci4) * (cj[4] - ci4) + (cj[5] - ci5) * (cj[5] - ci5);
          if (!(d2 > t2)) {
            if (di < cj[12]) kill = true; else stop = true;
          }
        }
      }
      const unsigned ms = __ballot_sync(0xffffffffu, stop);
      const int first_stop = ms ? __ffs(ms) - 1 : 32;
      const unsigned km = __ballot_sync(0xffffffffu, kill && lane < first_stop);
      __syncwarp();
      if (lane == 0) {
        const int w0 = jb >> 5, sh = jb & 31;
        if (km) {
          alive[w0] &= ~(km << sh);
          if (sh && w0 + 1 < nwords) alive[w0 + 1] &= ~(km >> (32 - sh));
        }
        if (ms) alive[i >> 5] &= ~(1u << (i & 31));
      }
      __syncwarp();
      if (ms) break;
    }
  }
  // survivors to the front of the group's segment, visit order kept
  int cnt = 0;
  for (int ib = 0; ib < k; ib += 32) {
    const int i = ib + lane;
    const bool a = i < k && ((alive[i >> 5] >> (i & 31)) & 1u);
    const int v = a ? perm[i] : 0;
    const unsigned m = __ballot_sync(0xffffffffu, a);
    __syncwarp();
    if (a) perm[cnt + __popc(m & ((1u << lane) - 1u))] = v;
    cnt += __popc(m);
    __syncwarp();
  }
  if (lane == 0) p.nsurv[q] = cnt;
}

__global__ void contact_dedupe_counts_kernel(DedupeParams p) {       // max(nsurv, 0) for the scan
  const int q = p.q0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (q < p.q1) p.soff[q] = p.nsurv[q] > 0 ? p.nsurv[q] : 0;
}

__global__ void contact_gather_kernel(DedupeParams p, int r0, int r1) {
  const int i = r0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= r1) return;
  const int q = p.recs[i].query;
  const int lo = p.offsets[q];
  const int r = i - lo;
  if (r == 0 && q == p.q1 - 1) {}      // (nothing special: the total is written below by the last query's first thread)
  const int ns = p.nsurv[q];
  if (r < ns) p.out[p.soff[q] + r] = p.recs[lo + p.perm[lo + r]];
}

__global__ void contact_dedupe_total_kernel(DedupeParams p) {
  // survivors of [q0, q1): exclusive offset of the last query + its own count
  const int q = p.q1 - 1;
  *p.out_total = p.soff[q] + (p.nsurv[q] > 0 ? p.nsurv[q] : 0);
}

size_t contact_dedupe_temp_bytes(int nqueries) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int *)nullptr, (int *)nullptr, nqueries);
  return bytes;
}

cudaError_t launch_contact_dedupe(const DedupeParams &p, cudaStream_t s) {
  const int nq = p.q1 - p.q0;
  if (nq <= 0) return cudaSuccess;
  contact_dedupe_kernel<<<(nq + DEDUPE_WARPS - 1) / DEDUPE_WARPS, 32 * DEDUPE_WARPS, 0, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  contact_dedupe_counts_kernel<<<(nq + 255) / 256, 256, 0, s>>>(p);
  size_t tb = p.temp_bytes;
  e = cub::DeviceScan::ExclusiveSum(p.temp, tb, p.soff + p.q0, p.soff + p.q0, nq, s);
  if (e != cudaSuccess) return e;
  // the scan ran in place: soff[q] now = exclusive offset; the counts are still in nsurv
  contact_dedupe_total_kernel<<<1, 1, 0, s>>>(p);
  const int r0 = 0;      // record range of [q0, q1) is read from the offsets on the host side: see the caller
  (void)r0;
  return cudaGetLastError();
}

cudaError_t launch_contact_gather(const DedupeParams &p, int r0, int r1, cudaStream_t s) {
  if (r1 <= r0) return cudaSuccess;
  contact_gather_kernel<<<(r1 - r0 + 255) / 256, 256, 0, s>>>(p, r0, r1);
  return cudaGetLastError();
}

} // namespace b2c
