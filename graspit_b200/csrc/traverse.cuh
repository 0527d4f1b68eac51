// Device helpers shared by the traversal kernels (kernels.cu, collide.cu).
#pragma once
#include "b2c_internal.h"

namespace b2c {

#define FULL 0xffffffffu

// stored node (16 B, bvh.h) -> decoded box.  q0 + q * qs with one fused rounding, like the host's decode_node.
__device__ __forceinline__ Node decode_nodeq(unsigned wx, unsigned wy, unsigned wz, unsigned ww, const float *q0, const float *qs) {
  float l0 = __fmaf_rn((float)(wx & 0xffffu), qs[0], q0[0]), l1 = __fmaf_rn((float)(wx >> 16), qs[1], q0[1]);
  float l2 = __fmaf_rn((float)(wy & 0xffffu), qs[2], q0[2]), h0 = __fmaf_rn((float)(wy >> 16), qs[0], q0[0]);
  float h1 = __fmaf_rn((float)(wz & 0xffffu), qs[1], q0[1]), h2 = __fmaf_rn((float)(wz >> 16), qs[2], q0[2]);
  Node n;
  n.cx = 0.5f * (l0 + h0); n.cy = 0.5f * (l1 + h1); n.cz = 0.5f * (l2 + h2);
  n.hx = 0.5f * (h0 - l0); n.hy = 0.5f * (h1 - l1); n.hz = 0.5f * (h2 - l2);
  const int link = (int)ww;
  n.child = link >= 0 ? link : -1;
  n.tri = link >= 0 ? -1 : ~link;
  return n;
}
__device__ __forceinline__ Node load_node(const MeshDev &m, int i) {
  uint4 w = __ldg(reinterpret_cast<const uint4 *>(m.nodes + i));
  return decode_nodeq(w.x, w.y, w.z, w.w, m.q0, m.qs);
}
// both children of a node (i = the parent's link, always even): one 256-bit load
__device__ __forceinline__ void load_pair(const MeshDev &m, int i, Node &c0, Node &c1) {
  unsigned r0, r1, r2, r3, r4, r5, r6, r7;
  asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
               : "l"(m.nodes + i));
  c0 = decode_nodeq(r0, r1, r2, r3, m.q0, m.qs);
  c1 = decode_nodeq(r4, r5, r6, r7, m.q0, m.qs);
}
// the same from a copy of the top of the tree in shared memory (i < ntop), else from global memory
__device__ __forceinline__ void fetch_pair(const MeshDev &m, const NodeQ *top, int ntop, int i, Node &c0, Node &c1) {
  if (i + 1 < ntop) {
    const uint4 *p = reinterpret_cast<const uint4 *>(top + i);
    uint4 a = p[0], b = p[1];
    c0 = decode_nodeq(a.x, a.y, a.z, a.w, m.q0, m.qs);
    c1 = decode_nodeq(b.x, b.y, b.z, b.w, m.q0, m.qs);
  } else {
    load_pair(m, i, c0, c1);
  }
}
// sibling pair as raw words (for kernels that test boxes in their stored lo/hi form)
__device__ __forceinline__ void fetch_pair_raw(const MeshDev &m, const NodeQ *top, int ntop, int i, uint4 &a, uint4 &b) {
  if (i + 1 < ntop) {
    const uint4 *p = reinterpret_cast<const uint4 *>(top + i);
    a = p[0]; b = p[1];
  } else {
    asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
                 : "l"(m.nodes + i));
  }
}
// squared distance from p to a stored box, straight from its quantised corners (no centre / half-extent decode)
__device__ __forceinline__ float point_nodeq_dist2(uint4 w, const float *q0, const float *qs, float px, float py, float pz) {
  float l0 = __fmaf_rn((float)(w.x & 0xffffu), qs[0], q0[0]), l1 = __fmaf_rn((float)(w.x >> 16), qs[1], q0[1]);
  float l2 = __fmaf_rn((float)(w.y & 0xffffu), qs[2], q0[2]), h0 = __fmaf_rn((float)(w.y >> 16), qs[0], q0[0]);
  float h1 = __fmaf_rn((float)(w.z & 0xffffu), qs[1], q0[1]), h2 = __fmaf_rn((float)(w.z >> 16), qs[2], q0[2]);
  float dx = fmaxf(fmaxf(l0 - px, px - h0), 0.f), dy = fmaxf(fmaxf(l1 - py, py - h1), 0.f), dz = fmaxf(fmaxf(l2 - pz, pz - h2), 0.f);
  return dx * dx + dy * dy + dz * dz;
}
__device__ __forceinline__ Node fetch_node(const MeshDev &m, const NodeQ *top, int ntop, int i) {
  if (i < ntop) {
    uint4 a = *reinterpret_cast<const uint4 *>(top + i);
    return decode_nodeq(a.x, a.y, a.z, a.w, m.q0, m.qs);
  }
  return load_node(m, i);
}
__device__ __forceinline__ void node_ch(const Node &n, float *c, float *h) {
  c[0] = n.cx; c[1] = n.cy; c[2] = n.cz; h[0] = n.hx; h[1] = n.hy; h[2] = n.hz;
}
__device__ __forceinline__ d3 tri_vertex(const Tri *tris, int t, int k) {
  float4 v = __ldg(&tris[t].v[k]);
  return mk<double>((double)v.x, (double)v.y, (double)v.z);
}

__device__ __forceinline__ unsigned long long pack_entry(int pair, int a, int b) {
  return ((unsigned long long)(unsigned)pair << 48) | ((unsigned long long)(unsigned)a << 24) | (unsigned long long)(unsigned)b;
}


__device__ __forceinline__ void make_pair_xf(const double *XA, const double *XB, float extA, float extB, PairXf &o) {
  // A <- B : R = Ra^T Rb, t = Ra^T (tb - ta), fp64 then rounded
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


__device__ __forceinline__ const double *body_xf_ptr(const Xf *fk, const Xf *static_xf, int nrb, int pose, int body) {
  return body < nrb ? reinterpret_cast<const double *>(fk + (size_t)pose * nrb + body)
                    : reinterpret_cast<const double *>(static_xf + (body - nrb));
}


// fp64 relative transform B <- A of a body pair: R = Rb^T Ra (row-major), t = Rb^T (ta - tb); element i of 12
__device__ __forceinline__ double rel_xf_element(const double *XA, const double *XB, int i) {
  if (i < 9) {
    int r = i / 3, c = i - 3 * r;
    return XB[r] * XA[c] + XB[3 + r] * XA[3 + c] + XB[6 + r] * XA[6 + c];
  }
  int r = i - 9;
  return XB[r] * (XA[9] - XB[9]) + XB[3 + r] * (XA[10] - XB[10]) + XB[6 + r] * (XA[11] - XB[11]);
}

} // namespace b2c
