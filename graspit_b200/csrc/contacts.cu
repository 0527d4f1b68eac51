// K6 contact-candidate extraction and K7 region extraction.
//
// K6 (contact_kernel): one warp per (pose, body pair).  Same cooperative work stack as the distance
// kernel, but with the fixed contact threshold as the bound: a node pair is kept while the lower bound of
// its box distance is <= threshold (reference ContactCallback::distanceTest, dSq <= thr^2,
// collisionAlgorithms.h:168-172).  Surviving triangle pairs are enumerated in fp64 exactly as
// triangleTriangleContact does (include/graspit/triangle_inl.h:284-404): nothing if the triangles
// intersect; else every vertex-face (6) and edge-edge (9) closest pair closer than the threshold (strict),
// skipping vertex-face pairs of triangle 1 whose foot is a vertex of triangle 2 and edge-edge pairs whose
// foot is an endpoint.  The reference detects those duplicates with exact == on coordinates; here they are
// detected from the feature (Voronoi region / clamped parameter), see SURVEY.md Appendix A.16.
// Each emitted contact follows ContactCallback::leafTest (collisionAlgorithms_inl.h:96-133): p1 moved to
// body 1's frame, n1 = normalize(p1 - p2) rotated to frame 1, n2 = -n1 in frame 2, distSq = the reference's
// mixed-frame |p1_frame1 - p2_frame2|^2.
//
// K7 (region_kernel): one thread per triangle of the body: triangles whose normal does not oppose the
// contact normal and whose closest point to the reference point lies within the radius are flagged
// (RegionCallback::leafTest, collisionAlgorithms_inl.h:252-275); the host collects their vertices.
#include <cub/cub.cuh>
#include <algorithm>
#include "b2c_internal.h"
#include "traverse.cuh"

namespace b2c {

__device__ __forceinline__ void c_node_ch(const Node &n, float *c, float *h) {
  c[0] = n.cx; c[1] = n.cy; c[2] = n.cz; h[0] = n.hx; h[1] = n.hy; h[2] = n.hz;
}
__device__ __forceinline__ d3 c_tri_vertex(const Tri *tris, int t, int k) {
  float4 v = __ldg(&tris[t].v[k]);
  return mk<double>((double)v.x, (double)v.y, (double)v.z);
}
__device__ __forceinline__ const double *c_body_xf(const Xf *fk, const Xf *static_xf, int nrb, int pose, int body) {
  return body < nrb ? reinterpret_cast<const double *>(fk + (size_t)pose * nrb + body)
                    : reinterpret_cast<const double *>(static_xf + (body - nrb));
}

__device__ __noinline__ void emit_contact(const ContactParams &p, int w, int ta, int tb, int feature, d3 p1, d3 p2,
                             const double *XA, const double *XB) {
  int slot = atomicAdd(p.out_count, 1);
  atomicAdd(&p.per_query[w], 1);
  if (slot >= p.out_cap) return;
  ContactRec r;
  r.query = w; r.tri_a = ta; r.tri_b = tb; r.feature = feature;
  d3 n = p1 - p2;
  double nn = sqrt(dot(n, n));
  d3 n1 = nn > 0.0 ? n * (1.0 / nn) : n;
  // back to body 1's frame: mTran2To1 = T1^-1 % T2
  d3 p1a = xf_apply_inv(XA, xf_apply(XB, p1));
  d3 n1a = xf_rot_t(XA, xf_rot(XB, n1));
  r.c[0] = p1a.x; r.c[1] = p1a.y; r.c[2] = p1a.z;
  r.c[3] = p2.x; r.c[4] = p2.y; r.c[5] = p2.z;
  r.c[6] = n1a.x; r.c[7] = n1a.y; r.c[8] = n1a.z;
  r.c[9] = -n1.x; r.c[10] = -n1.y; r.c[11] = -n1.z;
  d3 mixed = p1a - p2;
  r.c[12] = dot(mixed, mixed);
  p.out[slot] = r;
}

__device__ __noinline__ d3 closest_pt_triangle_dd(d3 a, d3 b, d3 c, d3 q, int *region) {
  return closest_pt_triangle<double>(a, b, c, q, region);
}
__device__ __noinline__ double seg_seg_dist2_dd(d3 p1, d3 q1, d3 p2, d3 q2, d3 &c1, d3 &c2, double *s, double *t) {
  return seg_seg_dist2<double>(p1, q1, p2, q2, c1, c2, s, t);
}

// all contacts of one triangle pair (a* already in B's frame)
__device__ __noinline__ void tri_pair_contacts(const ContactParams &p, int w, int ta, int tb, d3 a1, d3 a2, d3 a3,
                                               d3 b1, d3 b2, d3 b3, double thr2, const double *XA, const double *XB) {
  if (tri_tri_intersect_exact(a1, a2, a3, b1, b2, b3)) return;      // "Collision found when looking for contacts"
  d3 av[3] = {a1, a2, a3}, bv[3] = {b1, b2, b3};
  // The loops stay rolled and the helpers out of line: unrolled, this routine was several tens of KB of fp64 code and the
  // kernel spent its time waiting for instructions (ncu r01g: 4.5 warps stalled on "no instruction" per issue, 188 registers).
  // vertices of triangle 2 against face 1
#pragma unroll 1
  for (int k = 0; k < 3; k++) {
    d3 q = closest_pt_triangle_dd(a1, a2, a3, bv[k], nullptr);
    d3 d = bv[k] - q;
    if (dot(d, d) < thr2) emit_contact(p, w, ta, tb, k, q, bv[k], XA, XB);
  }
  // vertices of triangle 1 against face 2, unless the foot is a vertex of triangle 2 (already reported)
#pragma unroll 1
  for (int k = 0; k < 3; k++) {
    int region = 6;
    d3 q = closest_pt_triangle_dd(b1, b2, b3, av[k], &region);
    d3 d = q - av[k];
    if (dot(d, d) < thr2 && region > 2) emit_contact(p, w, ta, tb, 3 + k, av[k], q, XA, XB);
  }
  // nine edge-edge pairs, unless a foot is an endpoint
#pragma unroll 1
  for (int ij = 0; ij < 9; ij++) {
    const int i = ij / 3, j = ij - 3 * i;
    d3 c1, c2;
    double s, t;
    double dd = seg_seg_dist2_dd(av[i], av[(i + 1) % 3], bv[j], bv[(j + 1) % 3], c1, c2, &s, &t);
    if (dd < thr2 && s != 0.0 && s != 1.0 && t != 0.0 && t != 1.0) emit_contact(p, w, ta, tb, 6 + ij, c1, c2, XA, XB);
  }
}

constexpr int CSTACK_CAP = 448, CSTACK_PHYS = 512, CLEAF_CAP = 96, CEXACT_CAP = 64;
constexpr int CONTACT_WARP_SMEM = CSTACK_PHYS * 8 + CLEAF_CAP * 8 + CEXACT_CAP * 8 + 2 * 96 + 96 + 32;

// fp32 gate in front of the fp64 enumeration: the gap between the two triangles' projections on any axis is a lower
// bound of their distance, so a pair with such a gap above threshold + rounding head-room can neither touch nor
// intersect and the enumeration would emit nothing for it.  Axes: both face normals and the line of centroids.  Leaf
// boxes are axis aligned in the body frames and loose around slanted triangles; with a 0.2 mm threshold most queued
// pairs are of this kind (ncu r02g: 1 717 pairs enumerated per pose for 285 candidates).
__device__ __forceinline__ bool c_axis_far(f3 n, const f3 *a, const f3 *b, float lim2) {
  float a0 = dot(n, a[0]), a1 = dot(n, a[1]), a2 = dot(n, a[2]);
  float b0 = dot(n, b[0]), b1 = dot(n, b[1]), b2 = dot(n, b[2]);
  float gap = fmaxf(fminf(b0, fminf(b1, b2)) - fmaxf(a0, fmaxf(a1, a2)), fminf(a0, fminf(a1, a2)) - fmaxf(b0, fmaxf(b1, b2)));
  return gap > 0.f && gap * gap > lim2 * dot(n, n);
}
__device__ __noinline__ bool tri_pair_far(const Tri *tris_a, const Tri *tris_b, int ta, int tb, const PairXf *xp, float lim2) {
  f3 a[3], b[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    float4 va = __ldg(&tris_a[ta].v[k]), vb = __ldg(&tris_b[tb].v[k]);
    a[k] = mk<float>(va.x, va.y, va.z);
    b[k] = mk<float>(xp->r[0] * vb.x + xp->r[1] * vb.y + xp->r[2] * vb.z + xp->t[0],
                     xp->r[3] * vb.x + xp->r[4] * vb.y + xp->r[5] * vb.z + xp->t[1],
                     xp->r[6] * vb.x + xp->r[7] * vb.y + xp->r[8] * vb.z + xp->t[2]);
  }
  if (c_axis_far(cross(a[1] - a[0], a[2] - a[0]), a, b, lim2)) return true;
  if (c_axis_far(cross(b[1] - b[0], b[2] - b[0]), a, b, lim2)) return true;
  return c_axis_far((b[0] + b[1] + b[2]) - (a[0] + a[1] + a[2]), a, b, lim2);
}

// same PairXf construction as the evaluation kernel (kept local to this translation unit)
__device__ __forceinline__ void c_make_pair_xf(const double *XA, const double *XB, float extA, float extB, PairXf &o) {
  double r[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) r[3 * i + j] = XA[0 + i] * XB[0 + j] + XA[3 + i] * XB[3 + j] + XA[6 + i] * XB[6 + j];
  d3 dt = mk<double>(XB[9] - XA[9], XB[10] - XA[10], XB[11] - XA[11]);
  d3 t = xf_rot_t(XA, dt);
#pragma unroll
  for (int i = 0; i < 9; i++) { o.r[i] = (float)r[i]; o.ar[i] = fabsf((float)r[i]) + 1.0e-6f; }
  o.t[0] = (float)t.x; o.t[1] = (float)t.y; o.t[2] = (float)t.z;
  o.eps = 1.0e-6f * (fabsf(o.t[0]) + fabsf(o.t[1]) + fabsf(o.t[2]) + extA + extB) + 1.0e-5f;
  o.pad[0] = o.pad[1] = 0.f;
}

#ifndef CONTACT_MIN_BLOCKS
#define CONTACT_MIN_BLOCKS 2
#endif
__global__ void __launch_bounds__(256, CONTACT_MIN_BLOCKS) contact_kernel(ContactParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt_mask = (1u << lane) - 1u;
  unsigned char *base = smem_raw + (size_t)warp * CONTACT_WARP_SMEM;
  unsigned long long *stack = reinterpret_cast<unsigned long long *>(base);
  unsigned long long *leafq = stack + CSTACK_PHYS;
  double *XA = reinterpret_cast<double *>(leafq + CLEAF_CAP + CEXACT_CAP);
  double *XB = XA + 12;
  PairXf *xp = reinterpret_cast<PairXf *>(XB + 12);
  unsigned long long *exactq = leafq + CLEAF_CAP;
  const int total = p.npose * p.nq;
  const float thr = (float)p.threshold;
  // fp32 box bound with head-room so that the fp64 leaf test, not the culling, decides
  const float bound2 = (thr * 1.0001f + 1.0e-5f) * (thr * 1.0001f + 1.0e-5f);
  const double thr2 = p.threshold * p.threshold;

  const int nwork = p.work_list ? *p.work_count : total;
  while (true) {
    int w = 0;
    if (lane == 0) { w = atomicAdd(p.work_counter, 1); w = w < nwork ? (p.work_list ? p.work_list[w] : w) : -1; }
    w = __shfl_sync(FULL, w, 0);
    if (w < 0) break;
    int pose = w / p.nq, qi = w - pose * p.nq;
    int2 ab = __ldg(&p.queries[qi]);
    const MeshDev &ma = p.meshes[__ldg(&p.body_mesh[ab.x])];
    const MeshDev &mb = p.meshes[__ldg(&p.body_mesh[ab.y])];
    {
      const double *sa = c_body_xf(p.fk, p.static_xf, p.nrb, pose, ab.x);
      const double *sb = c_body_xf(p.fk, p.static_xf, p.nrb, pose, ab.y);
      if (lane < 12) XA[lane] = sa[lane];
      else if (lane < 24) XB[lane - 12] = sb[lane - 12];
    }
    __syncwarp();
    if (lane == 0) { PairXf x; c_make_pair_xf(XA, XB, ma.extent, mb.extent, x); *xp = x; }
    __syncwarp();
    const PairXf &x = *xp;
    int sp = 1, lq = 0, xq = 0;
    // head-room: the gate works on fp32 vertices through the fp32 pair transform
    const float far_lim = thr + 16.f * x.eps + 1.0e-3f, far_lim2 = far_lim * far_lim;
    // A stack entry names the node pair by what its two lanes will load: [63] which body's node is split, [24..47] the first
    // child of that node, [0..23] the other body's node -- decided by the lane that pushes it, which holds both boxes, so
    // the popping lanes issue their two loads at once instead of parent, then child (the wait for the parents' boxes was
    // 13 % of the kernel's samples, ncu r02i).
    int mode = 0;                        // 0: traverse, 1: nothing to do, 2: both roots are leaves (one triangle pair)
    if (lane == 0) {
      const Node ra = load_node(ma, 0), rb = load_node(mb, 0);
      const bool la = ra.child < 0, lb = rb.child < 0;
      if (la && lb) {
        mode = ra.tri >= 0 && rb.tri >= 0 ? 2 : 1;
        if (mode == 2) leafq[0] = ((unsigned long long)(unsigned)ra.tri << 24) | (unsigned)rb.tri;
      } else {
        const bool splitA = lb || (!la && (ra.hx * ra.hy * ra.hz > rb.hx * rb.hy * rb.hz));
        stack[0] = ((unsigned long long)(splitA ? 1u : 0u) << 63) | ((unsigned long long)(unsigned)(splitA ? ra.child : rb.child) << 24);
      }
    }
    mode = __shfl_sync(FULL, mode, 0);
    __syncwarp();
    if (mode == 1) continue;
    if (mode == 2) { sp = 0; lq = 1; }
    while (true) {
      if (lq >= 32 || (sp == 0 && lq > 0)) {
        // gate a full warp of queued triangle pairs; the survivors are packed for the fp64 enumeration
        int n = lq < 32 ? lq : 32;
        unsigned long long e = 0ull;
        bool keep = false;
        if (lane < n) {
          e = leafq[lq - n + lane];
          keep = p.exhaustive || !tri_pair_far(ma.tris, mb.tris, (int)((e >> 24) & 0xffffffu), (int)(e & 0xffffffu), xp, far_lim2);
        }
        unsigned km = __ballot_sync(FULL, keep);
        if (keep) exactq[xq + __popc(km & lt_mask)] = e;
        xq += __popc(km);
        lq -= n;
        __syncwarp();
        if (xq < 32 && (sp > 0 || lq > 0)) continue;
      }
      if (xq >= 32 || (sp == 0 && lq == 0 && xq > 0)) {
        int n = xq < 32 ? xq : 32;
        if (lane < n) {
          unsigned long long e = exactq[xq - n + lane];
          int ta = (int)((e >> 24) & 0xffffffu), tb = (int)(e & 0xffffffu);
          d3 a1 = xf_apply_inv(XB, xf_apply(XA, c_tri_vertex(ma.tris, ta, 0)));
          d3 a2 = xf_apply_inv(XB, xf_apply(XA, c_tri_vertex(ma.tris, ta, 1)));
          d3 a3 = xf_apply_inv(XB, xf_apply(XA, c_tri_vertex(ma.tris, ta, 2)));
          d3 b1 = c_tri_vertex(mb.tris, tb, 0), b2 = c_tri_vertex(mb.tris, tb, 1), b3 = c_tri_vertex(mb.tris, tb, 2);
          tri_pair_contacts(p, w, ta, tb, a1, a2, a3, b1, b2, b3, thr2, XA, XB);
        }
        xq -= n;
        __syncwarp();
        continue;
      }
      if (sp == 0) break;
      int room = CSTACK_CAP - sp;
      if (room <= -60) {
        // about to outgrow the physical stack: drop the rest of this pair and report it
        if (p.overflow && lane == 0) atomicOr(p.overflow, OVF_STACK);
        break;
      }
      // Two lanes per popped node pair, one per child of the node that is split: with the 0.2 mm band the front of live
      // node pairs is narrow (ncu r02g: 6 of 32 lanes in the box tests when one lane took a node pair and both children).
      int k = sp < 16 ? sp : 16;
      if (k > room) k = room > 1 ? room : 1;
      const int ent = lane >> 1, side = lane & 1;
      bool have = ent < k;
      unsigned long long e = have ? stack[sp - 1 - ent] : 0ull;
      sp -= k;
      __syncwarp();
      bool pp = false, ll = false;
      unsigned long long c = 0ull;
      if (have) {
        const bool sa = (e >> 63) != 0ull;
        const int first = (int)((e >> 24) & 0xffffffu) + side, other = (int)(e & 0xffffffu);
        const int na = sa ? first : other, nb = sa ? other : first;
        const Node a = load_node(ma, na), b = load_node(mb, nb);
        float ca[3], ha[3], cb[3], hb[3];
        c_node_ch(a, ca, ha);
        c_node_ch(b, cb, hb);
        pp = box_dist_lb(ca, ha, cb, hb, x) <= bound2;
        const bool la = a.child < 0, lb = b.child < 0;
        ll = pp && la && lb;
        if (ll) c = ((unsigned long long)(unsigned)a.tri << 24) | (unsigned)b.tri;
        else {
          const bool splitA = lb || (!la && (a.hx * a.hy * a.hz > b.hx * b.hy * b.hz));
          c = ((unsigned long long)(splitA ? 1u : 0u) << 63) | ((unsigned long long)(unsigned)(splitA ? a.child : b.child) << 24) |
              (unsigned)(splitA ? nb : na);
        }
      }
      unsigned s0 = __ballot_sync(FULL, pp && !ll), q0 = __ballot_sync(FULL, pp && ll);
      if (pp && !ll) stack[sp + __popc(s0 & lt_mask)] = c;
      sp += __popc(s0);
      if (pp && ll) leafq[lq + __popc(q0 & lt_mask)] = c;
      lq += __popc(q0);
      __syncwarp();
    }
    __syncwarp();
  }
}

__global__ void region_kernel(RegionParams p) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= p.mesh.ntri) return;
  d3 a = c_tri_vertex(p.mesh.tris, t, 0), b = c_tri_vertex(p.mesh.tris, t, 1), c = c_tri_vertex(p.mesh.tris, t, 2);
  d3 n = cross(b - a, c - a);
  double nn = sqrt(dot(n, n));
  if (nn > 0.0) n = n * (1.0 / nn);
  d3 cn = mk<double>(p.normal[0], p.normal[1], p.normal[2]);
  d3 ref = mk<double>(p.point[0], p.point[1], p.point[2]);
  unsigned char flag = 0;
  if (!(dot(n, cn) > 0.0)) {        // contact normals point inwards (collisionAlgorithms_inl.h:264-266)
    d3 q = closest_pt_triangle<double>(a, b, c, ref);
    d3 d = q - ref;
    if (!(dot(d, d) > p.radius * p.radius)) flag = 1;
  }
  p.flags[t] = flag;
}

// K7 batched: one warp per (body, point, normal, radius) query, lanes stride over the body's triangles with the same test
// as region_kernel; flagged triangles are appended to one list with warp-aggregated atomics.  Brute force over the
// triangles is deliberate: the bodies that get neighbourhoods are elastic fingertips and small objects (a few thousand
// triangles), and the test is ~100 fp64 flops.
// fp32 gate (round 2): |ref - a| - max(|b - a|, |c - a|) (the second term stored with the vertex) is a lower bound of the
// triangle's distance to ref; with
// head-room for the fp32 rounding, a triangle beyond the radius by this bound cannot be flagged.  22 of a fingertip's ~3 000
// triangles are, so the fp64 test runs on a packed queue of the few that pass.
__device__ __forceinline__ bool region_flag_exact(const RegionBatchParams &p, const MeshDev &m, int t, d3 cn, d3 ref, double r2) {
  d3 a = c_tri_vertex(m.tris, t, 0), b = c_tri_vertex(m.tris, t, 1), c = c_tri_vertex(m.tris, t, 2);
  d3 n = cross(b - a, c - a);
  double nn = sqrt(dot(n, n));
  if (nn > 0.0) n = n * (1.0 / nn);
  if (dot(n, cn) > 0.0) return false;
  d3 d = closest_pt_triangle<double>(a, b, c, ref) - ref;
  return !(dot(d, d) > r2);
}
__global__ void __launch_bounds__(256) region_batch_kernel(RegionBatchParams p) {
  __shared__ int queue_s[8][64];
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  int *queue = queue_s[threadIdx.x >> 5];
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  for (int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < p.nquery; q += nwarps) {
    const MeshDev &m = p.meshes[p.mesh_of[q]];
    const d3 cn = mk<double>(p.normal[3 * q], p.normal[3 * q + 1], p.normal[3 * q + 2]);
    const d3 ref = mk<double>(p.point[3 * q], p.point[3 * q + 1], p.point[3 * q + 2]);
    const double r = p.radius[q], r2 = r * r;
    const f3 reff = tof(ref);
    const float lim = (float)r * 1.00001f + 1.0e-3f + 1.0e-6f * (fabsf(reff.x) + fabsf(reff.y) + fabsf(reff.z));
    int nq = 0;
    for (int t0 = 0; t0 < m.ntri || nq > 0; t0 += 32) {
      const int t = t0 + lane;
      bool near = false;
      if (t < m.ntri) {
        const float4 va = __ldg(&m.tris[t].v[0]);              // .w: the triangle's radius about this vertex (mesh_create)
        const f3 ra = reff - mk<float>(va.x, va.y, va.z);
        near = !(sqrtf(dot(ra, ra)) - va.w > lim);
      }
      const unsigned nm = __ballot_sync(0xffffffffu, near);
      if (near) queue[nq + __popc(nm & lt)] = t;
      nq += __popc(nm);
      __syncwarp();
      if (nq < 32 && t0 + 32 < m.ntri) continue;
      // the exact test for a warp of queued triangles (the last, partial batch once the mesh is through)
      const int nb = nq < 32 ? nq : 32;
      int tq = -1;
      bool flag = false;
      if (lane < nb) { tq = queue[nq - nb + lane]; flag = region_flag_exact(p, m, tq, cn, ref, r2); }
      nq -= nb;
      __syncwarp();
      const unsigned mk_ = __ballot_sync(0xffffffffu, flag);
      if (mk_) {
        int base = 0;
        if (lane == __ffs(mk_) - 1) base = atomicAdd(p.out_count, __popc(mk_));
        base = __shfl_sync(0xffffffffu, base, __ffs(mk_) - 1);
        if (flag) {
          const int at = base + __popc(mk_ & lt);
          if (at < p.out_cap) p.out[at] = make_int2(tq, q);     // (triangle, query): as one 64-bit word the query is the high half
        }
      }
    }
  }
}

cudaError_t launch_region_batch(const RegionBatchParams &p, cudaStream_t s) {
  if (p.nquery <= 0) return cudaSuccess;
  const int blocks = std::max(1, std::min((p.nquery + 7) / 8, 148 * 8));
  region_batch_kernel<<<blocks, 256, 0, s>>>(p);
  return cudaGetLastError();
}

// flagged (triangle, query) pairs -> sorted by query, triangles ascending inside a query (the order b2c_region visits them):
// one radix sort of the pairs as 64-bit words
size_t region_sort_temp_bytes(int n, int query_bits) {
  size_t bytes = 0;
  cub::DoubleBuffer<unsigned long long> k(nullptr, nullptr);
  cub::DeviceRadixSort::SortKeys(nullptr, bytes, k, n, 0, 32 + query_bits);
  return bytes;
}
cudaError_t launch_region_sort(unsigned long long *pairs, unsigned long long *alt, int n, int query_bits, void *temp, size_t temp_bytes,
                               unsigned long long **sorted, cudaStream_t s) {
  cub::DoubleBuffer<unsigned long long> k(pairs, alt);
  cudaError_t e = cub::DeviceRadixSort::SortKeys(temp, temp_bytes, k, n, 0, 32 + query_bits, s);
  *sorted = k.Current();
  return e;
}

// ---- vertex lists of a region batch on the device -----------------------------------------------------------------
// RegionCallback::insertPoint (collisionAlgorithms.h:218-232) keeps the first occurrence of every vertex, compared with
// exact equality, in the order the flagged triangles are visited; bodyRegion then appends the query point itself
// (graspitCollision.cpp:495).  With the (triangle, query) pairs sorted by query and triangle: first[q] = the query's first
// pair; a vertex is kept when no earlier vertex of the same query equals it (equality is transitive, so this is the
// sequential rule); an exclusive scan of the keep flags gives every kept vertex its place, query q's list being shifted by
// the q query points appended before it.
__global__ void region_first_kernel(const int2 *hits, int nhits, int n, int *first) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k > nhits) return;
  const int qprev = k == 0 ? -1 : hits[k - 1].y, qcur = k == nhits ? n : hits[k].y;
  for (int q = qprev + 1; q <= qcur; q++) first[q] = k;
}
constexpr int REGION_STAGE = 384;          // vertices of one query staged in shared memory (a query holds ~70)
__global__ void __launch_bounds__(256) region_keep_kernel(const int2 *hits, const int *first, int n, const int *mesh_of,
                                                          const MeshDev *meshes, int *keep) {
  __shared__ float stage[8][REGION_STAGE][3];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int q = blockIdx.x * 8 + warp;
  if (q >= n) return;
  const int lo = first[q], k = 3 * (first[q + 1] - lo);
  const Tri *tris = meshes[mesh_of[q]].tris;
  float (*sv)[3] = stage[warp];
  for (int j = lane; j < k && j < REGION_STAGE; j += 32) {
    const float4 v = __ldg(&tris[hits[lo + j / 3].x].v[j % 3]);
    sv[j][0] = v.x; sv[j][1] = v.y; sv[j][2] = v.z;
  }
  __syncwarp();
  for (int j = lane; j < k; j += 32) {
    float vx, vy, vz;
    if (j < REGION_STAGE) { vx = sv[j][0]; vy = sv[j][1]; vz = sv[j][2]; }
    else { const float4 v = __ldg(&tris[hits[lo + j / 3].x].v[j % 3]); vx = v.x; vy = v.y; vz = v.z; }
    bool dup = false;
    const int ns = j < REGION_STAGE ? j : REGION_STAGE;
    for (int i = 0; i < ns && !dup; i++) dup = sv[i][0] == vx && sv[i][1] == vy && sv[i][2] == vz;
    for (int i = REGION_STAGE; i < j && !dup; i++) {
      const float4 u = __ldg(&tris[hits[lo + i / 3].x].v[i % 3]);
      dup = u.x == vx && u.y == vy && u.z == vz;
    }
    keep[3 * (size_t)lo + j] = dup ? 0 : 1;
  }
}
// thread j < 3 nhits: vertex j, written if kept; thread 3 nhits + q: query q's own point and its list's start
__global__ void region_write_kernel(const int2 *hits, int nhits, const int *first, int n, const int *mesh_of, const MeshDev *meshes,
                                    const int *keep, const int *pos, const double *points, double *out, long long cap, int *offsets) {
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long nv = 3ll * nhits;
  if (j < nv) {
    if (!keep[j]) return;
    const int2 h = hits[j / 3];
    const long long at = (long long)pos[j] + h.y;
    if (at >= cap) return;
    const float4 v = __ldg(&meshes[mesh_of[h.y]].tris[h.x].v[j % 3]);
    out[3 * at] = (double)v.x; out[3 * at + 1] = (double)v.y; out[3 * at + 2] = (double)v.z;
  } else if (j < nv + n) {
    const int q = (int)(j - nv);
    offsets[q] = pos[3 * (size_t)first[q]] + q;
    const long long at = (long long)pos[3 * (size_t)first[q + 1]] + q;
    if (q == n - 1) offsets[n] = (int)(at + 1);
    if (at >= cap) return;
    out[3 * at] = points[3 * (size_t)q]; out[3 * at + 1] = points[3 * (size_t)q + 1]; out[3 * at + 2] = points[3 * (size_t)q + 2];
  }
}
size_t region_lists_temp_bytes(int nhits) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int *)nullptr, (int *)nullptr, 3 * nhits + 1);
  return bytes;
}
// keep / pos: [3 nhits + 1] ints each; first: [n + 1].  After this pos[3 nhits] = the number of kept vertices.
cudaError_t launch_region_keep(const int2 *hits, int nhits, int n, const int *mesh_of, const MeshDev *meshes, int *first, int *keep,
                               int *pos, void *temp, size_t temp_bytes, cudaStream_t s) {
  region_first_kernel<<<(nhits + 1 + 255) / 256, 256, 0, s>>>(hits, nhits, n, first);
  cudaError_t e = cudaMemsetAsync(keep + 3 * (size_t)nhits, 0, sizeof(int), s);
  if (e != cudaSuccess) return e;
  region_keep_kernel<<<(n + 7) / 8, 256, 0, s>>>(hits, first, n, mesh_of, meshes, keep);
  return cub::DeviceScan::ExclusiveSum(temp, temp_bytes, keep, pos, 3 * nhits + 1, s);
}
// offsets: [n + 1] (device); out: [cap][3] doubles (device)
cudaError_t launch_region_write(const int2 *hits, int nhits, int n, const int *mesh_of, const MeshDev *meshes, const double *points,
                                const int *first, const int *keep, const int *pos, double *out, long long cap, int *offsets, cudaStream_t s) {
  const long long nt = 3ll * nhits + n;
  region_write_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, s>>>(hits, nhits, first, n, mesh_of, meshes, keep, pos, points, out, cap, offsets);
  return cudaGetLastError();
}

// Cull pass of a batched call, one thread per (pose, pair) query: only queries whose two root boxes come within the
// threshold go on the warps' work list.  (Of the 116 pairs of a HumanHand pose some 6 do; handed to a warp each, the
// other 110 cost a pair-transform set-up and one traversal step at 2 of 32 lanes -- a quarter of the kernel, ncu r02h.)
// A root-box lower bound above the threshold means no triangle pair within it, so the list changes no output.
__global__ void __launch_bounds__(256) contact_cull_kernel(ContactParams p) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
  const float thr = (float)p.threshold;
  const float bound2 = (thr * 1.0001f + 1.0e-5f) * (thr * 1.0001f + 1.0e-5f);
  bool keep = false;
  if (w < p.npose * p.nq) {
    const int pose = w / p.nq, qi = w - pose * p.nq;
    const int2 ab = __ldg(&p.queries[qi]);
    const MeshDev &ma = p.meshes[__ldg(&p.body_mesh[ab.x])];
    const MeshDev &mb = p.meshes[__ldg(&p.body_mesh[ab.y])];
    PairXf x;
    c_make_pair_xf(c_body_xf(p.fk, p.static_xf, p.nrb, pose, ab.x), c_body_xf(p.fk, p.static_xf, p.nrb, pose, ab.y), ma.extent, mb.extent, x);
    const Node a = load_node(ma, 0), b = load_node(mb, 0);
    float ca[3], ha[3], cb[3], hb[3];
    c_node_ch(a, ca, ha);
    c_node_ch(b, cb, hb);
    keep = box_dist_lb(ca, ha, cb, hb, x) <= bound2;
  }
  const unsigned m = __ballot_sync(FULL, keep);
  int base = 0;
  if (lane == 0 && m) base = atomicAdd(p.work_count, __popc(m));
  base = __shfl_sync(FULL, base, 0);
  if (keep) p.work_list[base + __popc(m & ((1u << lane) - 1u))] = w;
}

cudaError_t launch_contacts(const ContactParams &p, int blocks, cudaStream_t s) {
  if (p.work_list) {
    cudaError_t e = cudaMemsetAsync(p.work_count, 0, sizeof(int), s);
    if (e != cudaSuccess) return e;
    const long long n = (long long)p.npose * p.nq;
    contact_cull_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
  }
  size_t smem = (size_t)8 * CONTACT_WARP_SMEM;
  static size_t granted[64] = {};
  { cudaError_t e = ensure_dyn_smem((const void *)contact_kernel, smem, granted); if (e != cudaSuccess) return e; }
  contact_kernel<<<blocks, 256, smem, s>>>(p);
  return cudaGetLastError();
}

// ---- grouping of the candidate records by query (batched contacts) ------------------------------------------------
// contact_kernel emits records in arbitrary order and counts them per query; an exclusive scan of the counts gives each
// query its range, and one pass moves every record there.  The host then finds each (pose, pair) group contiguous.
__global__ void contact_scatter_kernel(const ContactRec *in, int n, const int *offsets, int *cursor, ContactRec *out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int q = in[i].query;
    const int slot = offsets[q] + atomicAdd(&cursor[q], 1);
    out[slot] = in[i];
  }
}

size_t contact_group_temp_bytes(int nqueries) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (const int *)nullptr, (int *)nullptr, nqueries);
  return bytes;
}

cudaError_t launch_contact_group(const ContactRec *in, int n, const int *per_query, int nqueries, int *offsets, int *cursor,
                                 ContactRec *out, void *temp, size_t temp_bytes, cudaStream_t s) {
  cudaError_t e = cub::DeviceScan::ExclusiveSum(temp, temp_bytes, per_query, offsets, nqueries, s);
  if (e != cudaSuccess) return e;
  e = cudaMemsetAsync(cursor, 0, sizeof(int) * (size_t)nqueries, s);
  if (e != cudaSuccess || n <= 0) return e;
  contact_scatter_kernel<<<std::min((n + 255) / 256, 148 * 8), 256, 0, s>>>(in, n, offsets, cursor, out);
  return cudaGetLastError();
}

// What the host needs of the grouping before it can enqueue the chunks: the first record of every chunk of poses and
// which pairs of the plan hold candidates at all (their reference-style trees are built on first use).  The full offset
// table follows on the copy stream.   out = [nslots chunk starts][np pair flags]
__global__ void contact_digest_kernel(const int *per_query, const int *offsets, int nqueries, int np, int stride, int *out, int nslots) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nqueries) return;
  if (per_query[q] > 0) out[nslots + q % np] = 1;
  if (q % stride == 0 && q / stride < nslots) out[q / stride] = offsets[q];
}
cudaError_t launch_contact_digest(const int *per_query, const int *offsets, int nqueries, int np, int stride, int *out, int nslots, cudaStream_t s) {
  cudaError_t e = cudaMemsetAsync(out, 0, sizeof(int) * (size_t)(nslots + np), s);
  if (e != cudaSuccess || nqueries <= 0) return e;
  contact_digest_kernel<<<(nqueries + 255) / 256, 256, 0, s>>>(per_query, offsets, nqueries, np, stride, out, nslots);
  return cudaGetLastError();
}

cudaError_t launch_region(const RegionParams &p, cudaStream_t s) {
  if (p.mesh.ntri <= 0) return cudaSuccess;
  region_kernel<<<(p.mesh.ntri + 255) / 256, 256, 0, s>>>(p);
  return cudaGetLastError();
}

} // namespace b2c
