// Device helpers shared by the traversal kernels (kernels.cu, collide.cu).
#pragma once
#include "b2c_internal.h"

namespace b2c {

#define FULL 0xffffffffu

__device__ __forceinline__ Node load_node(const Node *nodes, int i) {
  const float4 *p = reinterpret_cast<const float4 *>(nodes + i);
  float4 a = __ldg(p), b = __ldg(p + 1);
  Node n;
  n.cx = a.x; n.cy = a.y; n.cz = a.z; n.hx = a.w;
  n.hy = b.x; n.hz = b.y; n.child = __float_as_int(b.z); n.tri = __float_as_int(b.w);
  return n;
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
