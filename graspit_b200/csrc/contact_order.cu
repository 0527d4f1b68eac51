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
#include "b2c_internal.h"
#include "obb_order.h"

namespace b2c {

__global__ void __launch_bounds__(128) contact_order_kernel(OrderParams p) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.n) return;
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
  int depth = 0, n1 = 0, n2 = 0, second = 0;
  while (order_step(A, B, x, n1, n2, pa, pb, &second, range)) {
    if (second) {
      if (depth < 64) hi |= 1ull << (63 - depth);
      else if (depth < 128) lo |= 1ull << (127 - depth);
    }
    depth++;
    if (depth > 4096) break;              // (a malformed tree cannot hang the kernel)
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
  contact_order_kernel<<<(p.n + 127) / 128, 128, 0, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess || !p.perm) return e;
  e = cudaMemsetAsync(p.perm, 0xff, sizeof(int) * (size_t)p.n, s);
  if (e != cudaSuccess) return e;
  contact_rank_kernel<<<(p.n + 255) / 256, 256, 0, s>>>(p);
  return cudaGetLastError();
}

} // namespace b2c
