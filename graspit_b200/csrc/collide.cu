// K3: collision legality of a pose batch in two phases (replaces GraspitCollision::allCollisions,
// graspitCollision.cpp:332-363, with getActivePairs :55-100 and the CollisionCallback recursion,
// collisionAlgorithms.cpp:42-147, collisionAlgorithms_inl.h:50-81).
//
//   broad_kernel    one THREAD per (pose, active pair): builds the pair's fp32 relative transform from the two
//                   fp64 body transforms and tests the two root boxes.  Survivors become work items
//                   (pose, pair) in a global list (warp-aggregated append).  Every lane is busy: most pairs of
//                   most poses end here.
//   narrow_kernel   one WARP per 32 work items, persistent.  The 32 body-pair traversals share ONE work stack
//                   in shared memory: an entry is (item slot, nodeA, nodeB); each lane pops one entry, tests both
//                   children of the larger node against the other node, survivors are compacted back with
//                   ballot/popc; leaf pairs go to a queue drained 32 triangle pairs at a time.  An item that found
//                   an intersection is marked done and its remaining entries are dropped when popped (the reference
//                   stops a pair's recursion the same way, CollisionCallback::distanceTest).  With 32 independent
//                   traversals feeding one stack the lanes stay full where a warp-per-pose traversal idles on
//                   narrow frontiers.
//   compact_kernel  poses still legal -> compact list for the energy kernel.
//
// Triangle pairs whose fp32 SAT is too close to call go to the ambiguity list and are re-decided in fp64 by
// recheck_kernel (kernels.cu), as before.
#include "b2c_internal.h"
#include "traverse.cuh"

namespace b2c {

__global__ void __launch_bounds__(256) broad_kernel(EvalParams p) {
  const long long total = (long long)p.npose * p.npairs;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool ov = false;
  int pose = 0, pi = 0;
  if (i < total) {
    pose = (int)(i / p.npairs);
    pi = (int)(i - (long long)pose * p.npairs);
    if (pi == 0) {
      if (p.legal) p.legal[pose] = 1;
      if (p.energy) p.energy[pose] = 0.f;
    }
    // optional per-pose pair selection (autograsp: only the pairs on the interest list of a pose that still moves)
    const bool on = !p.pair_enable || ((__ldg(&p.pair_enable[(size_t)pose * p.mask_words + (pi >> 6)]) >> (pi & 63)) & 1ull);
    if (on) {
      int2 ab = __ldg(&p.pairs[pi]);
      const MeshDev &ma = p.meshes[__ldg(&p.body_mesh[ab.x])];
      const MeshDev &mb = p.meshes[__ldg(&p.body_mesh[ab.y])];
      const double *XA = body_xf_ptr(p.fk, p.static_xf, p.nrb, pose, ab.x);
      const double *XB = body_xf_ptr(p.fk, p.static_xf, p.nrb, pose, ab.y);
      PairXf x;
      make_pair_xf(XA, XB, ma.extent, mb.extent, x);
      Node ra = load_node(ma, 0), rb = load_node(mb, 0);
      float ca[3], ha[3], cb[3], hb[3];
      node_ch(ra, ca, ha); node_ch(rb, cb, hb);
      ov = box_overlap(ca, ha, cb, hb, x);
    }
  }
  // warp-aggregated append
  unsigned m = __ballot_sync(FULL, ov);
  if (m) {
    int base = 0;
    if (lane == __ffs(m) - 1) base = atomicAdd(p.item_count, __popc(m));
    base = __shfl_sync(FULL, base, __ffs(m) - 1);
    if (ov) {
      int at = base + __popc(m & ((1u << lane) - 1u));
      if (at < p.item_cap) p.items[at] = make_int2(pose, pi);
      else if (p.overflow) atomicOr(p.overflow, OVF_ITEMS);      // cannot happen while item_cap >= npose * npairs; reported if it does
    }
  }
  if (p.stats) {
    unsigned n = __popc(__ballot_sync(FULL, i < total));
    if (lane == 0 && n) atomicAdd(&p.stats[STAT_BOX], (unsigned long long)n);
  }
}

// per-warp shared memory of the narrow phase
struct NarrowSmem {
  double *xr;                    // [32][12]  B <- A, fp64 (leaf re-centring)
  PairXf *pxf;                   // [32]      A <- B, fp32 (box tests)
  int4 *meta;                    // [32]      pose, pair, mesh A, mesh B
  unsigned long long *stack;     // [NSTACK_PHYS]
  unsigned long long *leafq;     // [NLEAF_CAP]
};

__device__ __forceinline__ unsigned long long pack_item_entry(int slot, int a, int b) {
  return ((unsigned long long)(unsigned)slot << 48) | ((unsigned long long)(unsigned)a << 24) | (unsigned long long)(unsigned)b;
}

#ifndef NARROW_MIN_BLOCKS
#define NARROW_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(256, NARROW_MIN_BLOCKS) narrow_kernel(EvalParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned char *base = smem_raw + (size_t)warp * NARROW_WARP_SMEM;
  NarrowSmem S;
  S.xr = reinterpret_cast<double *>(base);
  S.pxf = reinterpret_cast<PairXf *>(S.xr + 32 * 12);
  S.meta = reinterpret_cast<int4 *>(S.pxf + 32);
  S.stack = reinterpret_cast<unsigned long long *>(S.meta + 32);
  S.leafq = S.stack + NSTACK_PHYS;
  unsigned n_box = 0, n_tri = 0, n_amb = 0;
  int nitems = *p.item_count;
  if (nitems > p.item_cap) nitems = p.item_cap;

  int *cnt = reinterpret_cast<int *>(S.leafq + NLEAF_CAP);     // [32] open entries (stack + leaf queue) per item slot
  cnt[lane] = 0;
  __syncwarp();
  int sp = 0, lq = 0;
  unsigned done = 0xffffffffu;       // bit s set: item slot s needs no more work (found a hit, or is free)
  bool more = nitems > 0;            // unclaimed work items remain

  while (true) {
    // ---- refill: free item slots take new work items, so the shared stack keeps 32 lanes busy while single
    // traversals run dry (a slot is free when none of its entries is left in the stack or the leaf queue)
    if (more && sp < 32) {
      const bool freeslot = cnt[lane] == 0;
      const unsigned fm = __ballot_sync(FULL, freeslot);
      const int f = __popc(fm);
      if (f >= 8 || (f > 0 && sp + lq == 0)) {
        int first = 0;
        if (lane == 0) first = atomicAdd(p.pose_counter, f);
        first = __shfl_sync(FULL, first, 0);
        if (first + f >= nitems) more = false;
        const int it = first + __popc(fm & lt_mask);
        bool valid = freeslot && it < nitems;
        int pose = 0, pi = 0;
        if (valid) {
          int2 e = p.items[it];
          pose = e.x; pi = e.y;
          // FAST_COLLISION: a pose already known to collide needs no further pairs
          if (!p.mode_all && p.legal && *reinterpret_cast<volatile uint8_t *>(&p.legal[pose]) == 0) valid = false;
        }
        if (valid) {
          int2 ab = __ldg(&p.pairs[pi]);
          int mA = __ldg(&p.body_mesh[ab.x]), mB = __ldg(&p.body_mesh[ab.y]);
          const double *XA = body_xf_ptr(p.fk, p.static_xf, p.nrb, pose, ab.x);
          const double *XB = body_xf_ptr(p.fk, p.static_xf, p.nrb, pose, ab.y);
          PairXf x;
          make_pair_xf(XA, XB, p.meshes[mA].extent, p.meshes[mB].extent, x);
          S.pxf[lane] = x;
#pragma unroll
          for (int k = 0; k < 12; k++) S.xr[12 * lane + k] = rel_xf_element(XA, XB, k);
          S.meta[lane] = make_int4(pose, pi, mA, mB);
          cnt[lane] = 1;
        }
        const unsigned vm = __ballot_sync(FULL, valid);
        if (valid) S.stack[sp + __popc(vm & lt_mask)] = pack_item_entry(lane, 0, 0);
        sp += __popc(vm);
        done &= ~vm;
        __syncwarp();
      }
    }
    if (sp == 0 && lq == 0) {
      if (!more) break;
      continue;
    }

    if (lq >= 32 || (sp == 0 && lq > 0)) {
      // ---- 32 triangle pairs
      int n = lq < 32 ? lq : 32;
      int res = TT_SEP;
      int slot = 0, ta = 0, tb = 0;
      if (lane < n) {
        unsigned long long e = S.leafq[lq - n + lane];
        slot = (int)(e >> 48); ta = (int)((e >> 24) & 0xffffffu); tb = (int)(e & 0xffffffu);
        if (!((done >> slot) & 1u)) {
          int4 mt = S.meta[slot];
          const MeshDev &ma = p.meshes[mt.z];
          const MeshDev &mb = p.meshes[mt.w];
          const double *XR = S.xr + 12 * slot;
          d3 a1 = tri_vertex(ma.tris, ta, 0), a2 = tri_vertex(ma.tris, ta, 1), a3 = tri_vertex(ma.tris, ta, 2);
          d3 b1 = tri_vertex(mb.tris, tb, 0), b2 = tri_vertex(mb.tris, tb, 1), b3 = tri_vertex(mb.tris, tb, 2);
          // triangle A into B's frame in fp64 (reference: t1.applyTransform(mTran1To2), collisionAlgorithms_inl.h:58-60),
          // everything re-centred on its first vertex before rounding
          d3 P1 = xf_apply(XR, a1);
          d3 E1 = xf_rot(XR, a2 - a1);
          d3 E2 = xf_rot(XR, a3 - a2);
          f3 e1 = tof(E1), e2 = tof(E2);
          f3 p2 = e1, p3 = tof(E1 + E2);
          f3 q1 = tof(b1 - P1), q2 = tof(b2 - P1), q3 = tof(b3 - P1);
          res = tri_tri_filter(p2, p3, e1, e2, q1, q2, q3, tof(b2 - b1), tof(b3 - b2));
          n_tri++;
        }
        atomicSub(&cnt[slot], 1);
      }
      lq -= n;
      if (res == TT_AMBIG) {
        n_amb++;
        int4 mt = S.meta[slot];
        int at = atomicAdd(p.ambig_count, 1);
        if (at < p.ambig_cap) { Ambig a; a.pose = mt.x; a.pair = mt.y; a.tri_a = ta; a.tri_b = tb; p.ambig[at] = a; }
        else {
          // the fp64 recheck queue is full: fail safe (an undecided pair counts as a collision) and say so
          res = TT_HIT;
          if (p.overflow) atomicOr(p.overflow, OVF_COLLIDE_QUEUE);
        }
      }
      if (res == TT_HIT) {
        int4 mt = S.meta[slot];
        if (p.legal) p.legal[mt.x] = 0;
        if (p.pair_mask) atomicOr(&p.pair_mask[(size_t)mt.x * p.mask_words + (mt.y >> 6)], 1ull << (mt.y & 63));
      }
      // every slot that hit is finished (all lanes learn it); with FAST_COLLISION so is every other pair of the same
      // pose held by this warp (the pairs of a pose are neighbours in the item list, so that is most of them)
      {
        unsigned hm = __ballot_sync(FULL, res == TT_HIT);
        const int mypose = S.meta[lane].x;          // pose of item slot `lane`
        while (hm) {
          int src = __ffs(hm) - 1;
          hm &= hm - 1;
          int hs = __shfl_sync(FULL, slot, src);
          done |= 1u << hs;
          if (!p.mode_all) {
            int hp = __shfl_sync(FULL, mypose, hs);
            done |= __ballot_sync(FULL, mypose == hp && cnt[lane] > 0);
          }
        }
      }
      __syncwarp();
      continue;
    }

    int room = NSTACK_CAP - sp;
    if (room <= -60) {
      // the shared stack is about to outgrow its physical size (never seen: it would take a tree far deeper than 2^23
      // triangles allow): give up on the open items LOUDLY -- they count as collisions and the sticky word reports it
      if (p.overflow && lane == 0) atomicOr(p.overflow, OVF_STACK);
      const int mypose = S.meta[lane].x;
      if (cnt[lane] > 0 && p.legal) p.legal[mypose] = 0;
      cnt[lane] = 0; sp = 0; lq = 0; done = 0xffffffffu;
      __syncwarp();
      continue;
    }
    int k = sp < 32 ? sp : 32;
    if (k > room) k = room > 1 ? room : 1;
    bool have = lane < k;
    unsigned long long e = have ? S.stack[sp - 1 - lane] : 0ull;
    sp -= k;
    __syncwarp();
    bool push0 = false, push1 = false, leaf0 = false, leaf1 = false;
    unsigned long long c0 = 0ull, c1 = 0ull;
    if (have) {
      int slot = (int)(e >> 48), na = (int)((e >> 24) & 0xffffffu), nb = (int)(e & 0xffffffu);
      if (!((done >> slot) & 1u)) {
        int4 mt = S.meta[slot];
        const MeshDev &ma = p.meshes[mt.z];
        const MeshDev &mb = p.meshes[mt.w];
        Node a = load_node(ma, na), b = load_node(mb, nb);
        bool la = a.child < 0, lb = b.child < 0;
        if (la && lb) {
          push0 = leaf0 = true;
          c0 = pack_item_entry(slot, a.tri, b.tri);
        } else {
          const PairXf &x = S.pxf[slot];
          // split the larger box (reference: getBoxVolume comparison, collisionAlgorithms.cpp:63-70)
          const bool splitA = lb || (!la && (a.hx * a.hy * a.hz > b.hx * b.hy * b.hz));
          // both children of the split node against the other node, on ONE code path for every lane (operands are
          // selected, not branched on: lanes that split A and lanes that split B run the same box test together)
          const MeshDev &cm = splitA ? ma : mb;
          const int cbase = splitA ? a.child : b.child;
          const Node other = splitA ? b : a;
          const bool oleaf = splitA ? lb : la;
          const int oidx = splitA ? nb : na;
          Node ch0, ch1;
          load_pair(cm, cbase, ch0, ch1);
          float co[3], ho[3], cc[3], hc[3];
          node_ch(other, co, ho);
          node_ch(ch0, cc, hc);
          push0 = box_overlap(splitA ? cc : co, splitA ? hc : ho, splitA ? co : cc, splitA ? ho : hc, x);
          node_ch(ch1, cc, hc);
          push1 = box_overlap(splitA ? cc : co, splitA ? hc : ho, splitA ? co : cc, splitA ? ho : hc, x);
          leaf0 = push0 && oleaf && ch0.child < 0;
          leaf1 = push1 && oleaf && ch1.child < 0;
          {
            int m0 = leaf0 ? ch0.tri : cbase, t0 = leaf0 ? other.tri : oidx;
            int m1 = leaf1 ? ch1.tri : cbase + 1, t1 = leaf1 ? other.tri : oidx;
            c0 = pack_item_entry(slot, splitA ? m0 : t0, splitA ? t0 : m0);
            c1 = pack_item_entry(slot, splitA ? m1 : t1, splitA ? t1 : m1);
          }
          n_box += 2;
        }
      }
      int delta = (push0 ? 1 : 0) + (push1 ? 1 : 0) - 1;
      if (delta != 0) atomicAdd(&cnt[slot], delta);
    }
    unsigned s0 = __ballot_sync(FULL, push0 && !leaf0), s1 = __ballot_sync(FULL, push1 && !leaf1);
    unsigned l0 = __ballot_sync(FULL, push0 && leaf0), l1 = __ballot_sync(FULL, push1 && leaf1);
    if (push0 && !leaf0) S.stack[sp + __popc(s0 & lt_mask)] = c0;
    if (push1 && !leaf1) S.stack[sp + __popc(s0) + __popc(s1 & lt_mask)] = c1;
    sp += __popc(s0) + __popc(s1);
    if (push0 && leaf0) S.leafq[lq + __popc(l0 & lt_mask)] = c0;
    if (push1 && leaf1) S.leafq[lq + __popc(l0) + __popc(l1 & lt_mask)] = c1;
    lq += __popc(l0) + __popc(l1);
    __syncwarp();
  }

  if (p.stats) {
    unsigned long long v;
    v = n_box; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o); if (lane == 0 && v) atomicAdd(&p.stats[STAT_BOX], v);
    v = n_tri; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o); if (lane == 0 && v) atomicAdd(&p.stats[STAT_TRI], v);
    v = n_amb; for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o); if (lane == 0 && v) atomicAdd(&p.stats[STAT_AMBIG], v);
  }
}

// poses without a collision -> compact list (order: pose index within a warp, warps in completion order)
__global__ void __launch_bounds__(256) compact_kernel(EvalParams p) {
  const int pose = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool ok = pose < p.npose && p.legal[pose] != 0;
  unsigned m = __ballot_sync(FULL, ok);
  if (!m) return;
  int base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(p.legal_count, __popc(m));
  base = __shfl_sync(FULL, base, __ffs(m) - 1);
  if (ok) p.legal_list[base + __popc(m & ((1u << lane) - 1u))] = pose;
}

cudaError_t launch_collide(const EvalParams &p, int narrow_blocks, cudaStream_t s) {
  if (p.npose <= 0 || p.npairs <= 0) {
    // nothing can collide: every pose is legal
    if (p.npose > 0 && p.legal) { cudaError_t e = cudaMemsetAsync(p.legal, 1, p.npose, s); if (e != cudaSuccess) return e; }
    if (p.npose > 0 && p.energy) { cudaError_t e = cudaMemsetAsync(p.energy, 0, sizeof(float) * (size_t)p.npose, s); if (e != cudaSuccess) return e; }
    return cudaSuccess;
  }
  const size_t smem = (size_t)8 * NARROW_WARP_SMEM;
  static size_t granted[64] = {};
  { cudaError_t e = ensure_dyn_smem((const void *)narrow_kernel, smem, granted); if (e != cudaSuccess) return e; }
  const long long total = (long long)p.npose * p.npairs;
  broad_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  narrow_kernel<<<narrow_blocks, 256, smem, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_compact(const EvalParams &p, cudaStream_t s) {
  if (p.npose <= 0 || !p.legal_list) return cudaSuccess;
  compact_kernel<<<(p.npose + 255) / 256, 256, 0, s>>>(p);
  return cudaGetLastError();
}

} // namespace b2c
