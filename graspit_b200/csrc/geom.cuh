// Device geometry primitives of the B200 collision engine (sm_100a).
//
// What each routine must agree with (reference file:line) is noted on the routine; the
// formulation is ours.  Precision scheme (DESIGN.md "Precision"):
//   * pose algebra (FK, relative transforms) and leaf re-centring: fp64
//   * every geometric test: fp32 on coordinates re-centred next to the query, so the fp32
//     relative error applies to the *local* scale, not to world coordinates
//   * triangle-triangle SAT returns SEP / HIT / AMBIGUOUS with an explicit rounding bound;
//     AMBIGUOUS pairs are re-decided in fp64 (tri_tri_intersect_exact) by the recheck kernel
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace b2c {

// ------------------------------------------------------------------------------------------
// small vector helpers (templated on float/double)
// ------------------------------------------------------------------------------------------
template <typename T> struct V3 { T x, y, z; };
typedef V3<float> f3;
typedef V3<double> d3;

template <typename T> __host__ __device__ __forceinline__ V3<T> mk(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator+(V3<T> a, V3<T> b) { return mk<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator-(V3<T> a, V3<T> b) { return mk<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator*(V3<T> a, T s) { return mk<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> __host__ __device__ __forceinline__ T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> __host__ __device__ __forceinline__ V3<T> cross(V3<T> a, V3<T> b) {
  return mk<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <typename T> __host__ __device__ __forceinline__ T min3(T a, T b, T c) { return fmin(a, fmin(b, c)); }
template <typename T> __host__ __device__ __forceinline__ T max3(T a, T b, T c) { return fmax(a, fmax(b, c)); }
// Eigen's normalized(): divides by the norm, leaves a zero vector alone
__host__ __device__ __forceinline__ d3 d_normalized(d3 a) {
  const double n2 = a.x * a.x + a.y * a.y + a.z * a.z;
  if (n2 > 0.0) { const double n = sqrt(n2); return mk<double>(a.x / n, a.y / n, a.z / n); }
  return a;
}
__host__ __device__ __forceinline__ f3 tof(d3 a) { return mk<float>((float)a.x, (float)a.y, (float)a.z); }
__host__ __device__ __forceinline__ d3 tod(f3 a) { return mk<double>((double)a.x, (double)a.y, (double)a.z); }

// Rigid transform as rotation matrix (row-major) + translation, fp64.  12 doubles = 96 B.
struct Xf {
  double r[9];
  double t[3];
};
__host__ __device__ __forceinline__ d3 xf_rot(const double *r, d3 v) {
  return mk<double>(r[0] * v.x + r[1] * v.y + r[2] * v.z,
                    r[3] * v.x + r[4] * v.y + r[5] * v.z,
                    r[6] * v.x + r[7] * v.y + r[8] * v.z);
}
__host__ __device__ __forceinline__ d3 xf_rot_t(const double *r, d3 v) {   // R^T v
  return mk<double>(r[0] * v.x + r[3] * v.y + r[6] * v.z,
                    r[1] * v.x + r[4] * v.y + r[7] * v.z,
                    r[2] * v.x + r[5] * v.y + r[8] * v.z);
}
__host__ __device__ __forceinline__ d3 xf_apply(const double *rt, d3 v) {       // R v + t
  d3 o = xf_rot(rt, v);
  return mk<double>(o.x + rt[9], o.y + rt[10], o.z + rt[11]);
}
__host__ __device__ __forceinline__ d3 xf_apply_inv(const double *rt, d3 v) {   // R^T (v - t)
  return xf_rot_t(rt, mk<double>(v.x - rt[9], v.y - rt[10], v.z - rt[11]));
}

// ------------------------------------------------------------------------------------------
// Quaternion algebra mirroring GraspIt!'s transf (include/graspit/transform.h:47-49,
// src/transform.cpp:68-72,154-159): unit quaternion wxyz re-normalised after every product,
// translation composed through the quaternion rotation v + w*2(u x v) + u x 2(u x v).
// ------------------------------------------------------------------------------------------
struct Qt {
  double w, x, y, z;
  double tx, ty, tz;
};
__host__ __device__ __forceinline__ void q_normalize(Qt &a) {
  double n2 = a.w * a.w + a.x * a.x + a.y * a.y + a.z * a.z;
  if (n2 > 0.0) { double n = sqrt(n2); a.w /= n; a.x /= n; a.y /= n; a.z /= n; }
}
__host__ __device__ __forceinline__ d3 q_rot(const Qt &q, d3 v) {
  d3 u = mk<double>(q.x, q.y, q.z);
  d3 uv = cross(u, v);
  uv = uv + uv;
  d3 c = cross(u, uv);
  return mk<double>(v.x + q.w * uv.x + c.x, v.y + q.w * uv.y + c.y, v.z + q.w * uv.z + c.z);
}
__host__ __device__ __forceinline__ Qt q_compose(const Qt &a, const Qt &b) {   // a % b
  Qt r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
  r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x;
  q_normalize(r);
  d3 t = q_rot(a, mk<double>(b.tx, b.ty, b.tz));
  r.tx = a.tx + t.x; r.ty = a.ty + t.y; r.tz = a.tz + t.z;
  return r;
}
__host__ __device__ __forceinline__ Qt q_inverse(const Qt &a) {
  Qt r;
  double n2 = a.w * a.w + a.x * a.x + a.y * a.y + a.z * a.z;
  r.w = a.w / n2; r.x = -a.x / n2; r.y = -a.y / n2; r.z = -a.z / n2;
  r.tx = r.ty = r.tz = 0.0;
  d3 t = q_rot(r, mk<double>(a.tx, a.ty, a.tz));
  q_normalize(r);
  r.tx = -t.x; r.ty = -t.y; r.tz = -t.z;
  return r;
}
// rotation matrix (row-major) -> unit quaternion: Eigen's trace-branch algorithm, what transf::set(mat3, vec3) stores
// (include/graspit/transform.h:52-57); translation left at zero
__host__ __device__ __forceinline__ Qt q_from_matrix(const double *m) {
  double c[4];   // x y z w
  double t = m[0] + m[4] + m[8];
  if (t > 0.0) {
    t = sqrt(t + 1.0);
    c[3] = 0.5 * t;
    t = 0.5 / t;
    c[0] = (m[7] - m[5]) * t; c[1] = (m[2] - m[6]) * t; c[2] = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    c[i] = 0.5 * t;
    t = 0.5 / t;
    c[3] = (m[3 * k + j] - m[3 * j + k]) * t;
    c[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    c[k] = (m[3 * k + i] + m[3 * i + k]) * t;
  }
  Qt r;
  r.w = c[3]; r.x = c[0]; r.y = c[1]; r.z = c[2];
  r.tx = r.ty = r.tz = 0.0;
  q_normalize(r);
  return r;
}
__host__ __device__ __forceinline__ Qt q_identity() { Qt r; r.w = 1; r.x = r.y = r.z = 0; r.tx = r.ty = r.tz = 0; return r; }
// transf::AXIS_ANGLE_ROTATION (src/transform.cpp:23-34) incl. the IDENTITY snap for angle^2 < 1e-6
__host__ __device__ __forceinline__ Qt q_axis_angle(double angle, double ax, double ay, double az) {
  Qt r = q_identity();
  if (angle * angle < 0.000001) return r;
  double n = sqrt(ax * ax + ay * ay + az * az);
  if (n > 0.0) { ax /= n; ay /= n; az /= n; }
  double s, c;
  sincos(0.5 * angle, &s, &c);
  r.w = c; r.x = s * ax; r.y = s * ay; r.z = s * az;
  q_normalize(r);
  return r;
}
// Eigen's toRotationMatrix, row-major
__host__ __device__ __forceinline__ void q_to_xf(const Qt &q, double *rt) {
  const double tx = 2.0 * q.x, ty = 2.0 * q.y, tz = 2.0 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w;
  const double txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
  const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
  rt[0] = 1.0 - (tyy + tzz); rt[1] = txy - twz;         rt[2] = txz + twy;
  rt[3] = txy + twz;         rt[4] = 1.0 - (txx + tzz); rt[5] = tyz - twx;
  rt[6] = txz - twy;         rt[7] = tyz + twx;         rt[8] = 1.0 - (txx + tyy);
  rt[9] = q.tx; rt[10] = q.ty; rt[11] = q.tz;
}

// ------------------------------------------------------------------------------------------
// Boxes.  A BVH node is an axis-aligned box in its body's frame (centre c, half extents h).
// For a body pair the rotation R (A <- B), |R|+eps and translation t are shared by every node
// test of the pair.  Conservative replacement for bboxOverlap (include/graspit/bbox_inl.h:88-255):
// results of the exact leaf tests do not depend on which conservative culling is used
// (SURVEY.md Appendix A.10).
// ------------------------------------------------------------------------------------------
struct PairXf {          // fp32, 24 floats = 96 B
  float r[9];            // A <- B rotation, row-major
  float ar[9];           // |r| + 1e-6
  float t[3];            // A <- B translation
  float eps;             // absolute slack covering fp32 rounding of r, t at this pair's scale
  float pad[2];
};

__device__ __forceinline__ bool box_overlap(const float *ca, const float *ha, const float *cb, const float *hb,
                                            const PairXf &x) {
  const float *R = x.r, *AR = x.ar;
  // centre of B in A's frame, relative to A's centre
  float T0 = R[0] * cb[0] + R[1] * cb[1] + R[2] * cb[2] + x.t[0] - ca[0];
  float T1 = R[3] * cb[0] + R[4] * cb[1] + R[5] * cb[2] + x.t[1] - ca[1];
  float T2 = R[6] * cb[0] + R[7] * cb[1] + R[8] * cb[2] + x.t[2] - ca[2];
  const float e = x.eps;
  // the six face axes without branches (most separated pairs are separated on one of them), then ONE branch, then
  // the nine edge-edge axes: lanes of a warp diverge at most once inside a test
  bool sep = fabsf(T0) > ha[0] + hb[0] * AR[0] + hb[1] * AR[1] + hb[2] * AR[2] + e;
  sep |= fabsf(T1) > ha[1] + hb[0] * AR[3] + hb[1] * AR[4] + hb[2] * AR[5] + e;
  sep |= fabsf(T2) > ha[2] + hb[0] * AR[6] + hb[1] * AR[7] + hb[2] * AR[8] + e;
  sep |= fabsf(T0 * R[0] + T1 * R[3] + T2 * R[6]) > hb[0] + ha[0] * AR[0] + ha[1] * AR[3] + ha[2] * AR[6] + e;
  sep |= fabsf(T0 * R[1] + T1 * R[4] + T2 * R[7]) > hb[1] + ha[0] * AR[1] + ha[1] * AR[4] + ha[2] * AR[7] + e;
  sep |= fabsf(T0 * R[2] + T1 * R[5] + T2 * R[8]) > hb[2] + ha[0] * AR[2] + ha[1] * AR[5] + ha[2] * AR[8] + e;
  if (sep) return false;
  // 9 edge-edge axes A_i x B_j
  sep |= fabsf(T2 * R[3] - T1 * R[6]) > ha[1] * AR[6] + ha[2] * AR[3] + hb[1] * AR[2] + hb[2] * AR[1] + e;
  sep |= fabsf(T2 * R[4] - T1 * R[7]) > ha[1] * AR[7] + ha[2] * AR[4] + hb[0] * AR[2] + hb[2] * AR[0] + e;
  sep |= fabsf(T2 * R[5] - T1 * R[8]) > ha[1] * AR[8] + ha[2] * AR[5] + hb[0] * AR[1] + hb[1] * AR[0] + e;
  sep |= fabsf(T0 * R[6] - T2 * R[0]) > ha[0] * AR[6] + ha[2] * AR[0] + hb[1] * AR[5] + hb[2] * AR[4] + e;
  sep |= fabsf(T0 * R[7] - T2 * R[1]) > ha[0] * AR[7] + ha[2] * AR[1] + hb[0] * AR[5] + hb[2] * AR[3] + e;
  sep |= fabsf(T0 * R[8] - T2 * R[2]) > ha[0] * AR[8] + ha[2] * AR[2] + hb[0] * AR[4] + hb[1] * AR[3] + e;
  sep |= fabsf(T1 * R[0] - T0 * R[3]) > ha[0] * AR[3] + ha[1] * AR[0] + hb[1] * AR[8] + hb[2] * AR[7] + e;
  sep |= fabsf(T1 * R[1] - T0 * R[4]) > ha[0] * AR[4] + ha[1] * AR[1] + hb[0] * AR[8] + hb[2] * AR[6] + e;
  sep |= fabsf(T1 * R[2] - T0 * R[5]) > ha[0] * AR[5] + ha[1] * AR[2] + hb[0] * AR[7] + hb[1] * AR[6] + e;
  return !sep;
}

// Lower bound on the squared distance between the two boxes: squared positive gaps along the three
// face axes of A, and of B, the larger of the two sums (each sum is a valid bound because the three
// axes of one box are orthonormal).  Plays the role of bboxDistanceSq (bbox_inl.h:269-413); any
// valid lower bound yields the same minimum distance.  Slack eps keeps it conservative in fp32.
__device__ __forceinline__ float box_dist_lb(const float *ca, const float *ha, const float *cb, const float *hb,
                                             const PairXf &x) {
  const float *R = x.r, *AR = x.ar;
  float T0 = R[0] * cb[0] + R[1] * cb[1] + R[2] * cb[2] + x.t[0] - ca[0];
  float T1 = R[3] * cb[0] + R[4] * cb[1] + R[5] * cb[2] + x.t[1] - ca[1];
  float T2 = R[6] * cb[0] + R[7] * cb[1] + R[8] * cb[2] + x.t[2] - ca[2];
  const float e = x.eps;
  float g0 = fmaxf(0.f, fabsf(T0) - (ha[0] + hb[0] * AR[0] + hb[1] * AR[1] + hb[2] * AR[2] + e));
  float g1 = fmaxf(0.f, fabsf(T1) - (ha[1] + hb[0] * AR[3] + hb[1] * AR[4] + hb[2] * AR[5] + e));
  float g2 = fmaxf(0.f, fabsf(T2) - (ha[2] + hb[0] * AR[6] + hb[1] * AR[7] + hb[2] * AR[8] + e));
  float sa = g0 * g0 + g1 * g1 + g2 * g2;
  float h0 = fmaxf(0.f, fabsf(T0 * R[0] + T1 * R[3] + T2 * R[6]) - (hb[0] + ha[0] * AR[0] + ha[1] * AR[3] + ha[2] * AR[6] + e));
  float h1 = fmaxf(0.f, fabsf(T0 * R[1] + T1 * R[4] + T2 * R[7]) - (hb[1] + ha[0] * AR[1] + ha[1] * AR[4] + ha[2] * AR[7] + e));
  float h2 = fmaxf(0.f, fabsf(T0 * R[2] + T1 * R[5] + T2 * R[8]) - (hb[2] + ha[0] * AR[2] + ha[1] * AR[5] + ha[2] * AR[8] + e));
  float sb = h0 * h0 + h1 * h1 + h2 * h2;
  return fmaxf(sa, sb);
}

// squared distance from point p to the axis-aligned box (c,h); role of pointBoxDistanceSq
// (bbox_inl.h:32-54)
__device__ __forceinline__ float point_box_dist2(const float *c, const float *h, f3 p) {
  float dx = fmaxf(0.f, fabsf(p.x - c[0]) - h[0]);
  float dy = fmaxf(0.f, fabsf(p.y - c[1]) - h[1]);
  float dz = fmaxf(0.f, fabsf(p.z - c[2]) - h[2]);
  return dx * dx + dy * dy + dz * dz;
}

// ------------------------------------------------------------------------------------------
// Closest point on triangle (a,b,c) to p.  Must agree with closestPtTriangle
// (include/graspit/triangle_inl.h:149-197): same Voronoi-region classification (Ericson, RTCD 5.1.5),
// vertices returned exactly.  `region` (optional) reports the feature: 0-2 vertex, 3-5 edge
// (ab, ac, bc), 6 face -- the feature code replaces the reference's exact == comparisons when
// contacts are enumerated.
// ------------------------------------------------------------------------------------------
template <typename T>
__host__ __device__ __forceinline__ V3<T> closest_pt_triangle(V3<T> a, V3<T> b, V3<T> c, V3<T> p, int *region = nullptr) {
  V3<T> ab = b - a, ac = c - a, ap = p - a;
  T d1 = dot(ab, ap), d2 = dot(ac, ap);
  if (d1 <= T(0) && d2 <= T(0)) { if (region) *region = 0; return a; }
  V3<T> bp = p - b;
  T d3 = dot(ab, bp), d4 = dot(ac, bp);
  if (d3 >= T(0) && d4 <= d3) { if (region) *region = 1; return b; }
  T vc = d1 * d4 - d3 * d2;
  if (vc <= T(0) && d1 >= T(0) && d3 <= T(0)) {
    T v = d1 / (d1 - d3);
    if (region) *region = 3;
    return a + ab * v;
  }
  V3<T> cp = p - c;
  T d5 = dot(ab, cp), d6 = dot(ac, cp);
  if (d6 >= T(0) && d5 <= d6) { if (region) *region = 2; return c; }
  T vb = d5 * d2 - d1 * d6;
  if (vb <= T(0) && d2 >= T(0) && d6 <= T(0)) {
    T w = d2 / (d2 - d6);
    if (region) *region = 4;
    return a + ac * w;
  }
  T va = d3 * d6 - d5 * d4;
  if (va <= T(0) && (d4 - d3) >= T(0) && (d5 - d6) >= T(0)) {
    T w = (d4 - d3) / ((d4 - d3) + (d5 - d6));
    if (region) *region = 5;
    return b + (c - b) * w;
  }
  T denom = T(1) / (va + vb + vc);
  T v = vb * denom, w = vc * denom;
  if (region) *region = 6;
  return a + ab * v + ac * w;
}

template <typename T> __host__ __device__ __forceinline__ T clamp01(T x) { return x < T(0) ? T(0) : (x > T(1) ? T(1) : x); }

// Closest points of segments [p1,q1], [p2,q2]; must agree with segmSegmDistanceSq
// (triangle_inl.h:209-272), including EPSILON = 1e-3 on the *squared* segment length and the exact
// `denom != 0` parallel test.  s,t (optional) return the clamped parameters.
template <typename T>
__host__ __device__ __forceinline__ T seg_seg_dist2(V3<T> p1, V3<T> q1, V3<T> p2, V3<T> q2, V3<T> &c1, V3<T> &c2,
                                                    T *so = nullptr, T *to = nullptr) {
  const T EPS = T(1.0e-3);
  V3<T> d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
  T a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r);
  T s, t;
  if (a <= EPS && e <= EPS) {
    s = t = T(0);
    c1 = p1; c2 = p2;
    if (so) *so = s; if (to) *to = t;
    V3<T> d = c1 - c2;
    return dot(d, d);
  }
  if (a <= EPS) {
    s = T(0);
    t = clamp01(f / e);
  } else {
    T c = dot(d1, r);
    if (e <= EPS) {
      t = T(0);
      s = clamp01(-c / a);
    } else {
      T b = dot(d1, d2);
      T denom = a * e - b * b;
      if (denom != T(0)) s = clamp01((b * f - c * e) / denom);
      else s = T(0);
      t = (b * s + f) / e;
      if (t < T(0)) { t = T(0); s = clamp01(-c / a); }
      else if (t > T(1)) { t = T(1); s = clamp01((b - c) / a); }
    }
  }
  c1 = p1 + d1 * s;
  c2 = p2 + d2 * t;
  if (so) *so = s; if (to) *to = t;
  V3<T> d = c1 - c2;
  return dot(d, d);
}

// ------------------------------------------------------------------------------------------
// Triangle-triangle intersection.
// ------------------------------------------------------------------------------------------
enum { TT_SEP = 0, TT_HIT = 1, TT_AMBIG = 2 };

// classify one SAT axis in fp32.  p1 is the origin, so its projection is 0.
// gap > 0 <=> the reference's project6 would report separation (mn1 > mx2 || mn2 > mx1,
// triangle_inl.h:48-68).  `scale` = largest |coordinate| of the re-centred vertices.
__device__ __forceinline__ int sat_axis(f3 ax, f3 p2, f3 p3, f3 q1, f3 q2, f3 q3, float scale, float axmin) {
  float l1 = fabsf(ax.x) + fabsf(ax.y) + fabsf(ax.z);
  if (l1 <= axmin) return 1;                 // numerically null axis (parallel edges): cannot separate
  float P2 = dot(ax, p2), P3 = dot(ax, p3);
  float Q1 = dot(ax, q1), Q2 = dot(ax, q2), Q3 = dot(ax, q3);
  float mx1 = fmaxf(0.f, fmaxf(P2, P3)), mn1 = fminf(0.f, fminf(P2, P3));
  float mx2 = max3(Q1, Q2, Q3), mn2 = min3(Q1, Q2, Q3);
  float gap = fmaxf(mn1 - mx2, mn2 - mx1);
  float tol = 2.0e-6f * l1 * scale;
  if (gap > tol) return 0;                   // certainly separated
  if (gap < -tol) return 1;                  // certainly overlapping on this axis
  return 2;                                  // too close to call in fp32
}

// fp32 filter for triangleIntersection (triangle_inl.h:70-144).
//   p2,p3      : triangle 1 vertices relative to its first vertex (p1 = 0), in triangle 2's frame
//   e1,e2      : triangle 1 edges p2-p1, p3-p2 (e3 = -(e1+e2)), each rounded from fp64 with
//                *relative* accuracy (rotated exact differences)
//   q1..q3     : triangle 2 vertices relative to p1;  f1,f2: its edges q2-q1, q3-q2
__device__ __forceinline__ int tri_tri_filter(f3 p2, f3 p3, f3 e1, f3 e2, f3 q1, f3 q2, f3 q3, f3 f1, f3 f2) {
  f3 e3 = mk<float>(-(e1.x + e2.x), -(e1.y + e2.y), -(e1.z + e2.z));
  f3 f3_ = mk<float>(-(f1.x + f2.x), -(f1.y + f2.y), -(f1.z + f2.z));
  float scale = fmaxf(fmaxf(fmaxf(fabsf(p2.x), fabsf(p2.y)), fmaxf(fabsf(p2.z), fabsf(p3.x))),
                      fmaxf(fabsf(p3.y), fabsf(p3.z)));
  scale = fmaxf(scale, fmaxf(fmaxf(fabsf(q1.x), fabsf(q1.y)), fabsf(q1.z)));
  scale = fmaxf(scale, fmaxf(fmaxf(fabsf(q2.x), fabsf(q2.y)), fabsf(q2.z)));
  scale = fmaxf(scale, fmaxf(fmaxf(fabsf(q3.x), fabsf(q3.y)), fabsf(q3.z)));
  float le = fmaxf(fabsf(e1.x) + fabsf(e1.y) + fabsf(e1.z), fabsf(e2.x) + fabsf(e2.y) + fabsf(e2.z));
  float lf = fmaxf(fabsf(f1.x) + fabsf(f1.y) + fabsf(f1.z), fabsf(f2.x) + fabsf(f2.y) + fabsf(f2.z));
  const float axmin = 1.0e-5f * le * lf;     // edge-cross axes smaller than this are rounding noise
  int amb = 0, r;
  f3 n1 = cross(e1, e2), m1 = cross(f1, f2);
#define B2C_AXIS(AX, MIN)                                        \
  r = sat_axis((AX), p2, p3, q1, q2, q3, scale, (MIN));          \
  if (r == 0) return TT_SEP;                                     \
  amb |= (r == 2);
  B2C_AXIS(n1, 0.f)
  B2C_AXIS(m1, 0.f)
  B2C_AXIS(cross(e1, f1), axmin) B2C_AXIS(cross(e1, f2), axmin) B2C_AXIS(cross(e1, f3_), axmin)
  B2C_AXIS(cross(e2, f1), axmin) B2C_AXIS(cross(e2, f2), axmin) B2C_AXIS(cross(e2, f3_), axmin)
  B2C_AXIS(cross(e3, f1), axmin) B2C_AXIS(cross(e3, f2), axmin) B2C_AXIS(cross(e3, f3_), axmin)
  B2C_AXIS(cross(e1, n1), 0.f) B2C_AXIS(cross(e2, n1), 0.f) B2C_AXIS(cross(e3, n1), 0.f)
  B2C_AXIS(cross(f1, m1), 0.f) B2C_AXIS(cross(f2, m1), 0.f) B2C_AXIS(cross(f3_, m1), 0.f)
#undef B2C_AXIS
  return amb ? TT_AMBIG : TT_HIT;
}

// Exact-order fp64 restatement used only for AMBIGUOUS pairs and by the contact kernels.  Decides
// exactly like triangleIntersection (triangle_inl.h:70-144): 17 axes (n1, m1, 9 e x f, 3 e x n1,
// 3 f x m1), separated only if strictly mn1 > mx2 or mn2 > mx1, i.e. touching counts as intersecting.
// t1* are triangle 1's vertices already expressed in triangle 2's frame.
__host__ __device__ inline bool axis_separates(d3 ax, d3 p1, d3 p2, d3 p3, d3 q1, d3 q2, d3 q3) {
  double P1 = dot(ax, p1), P2 = dot(ax, p2), P3 = dot(ax, p3);
  double Q1 = dot(ax, q1), Q2 = dot(ax, q2), Q3 = dot(ax, q3);
  double mx1 = P1, mn1 = P1, mx2 = Q1, mn2 = Q1;
  if (P2 > mx1) mx1 = P2; if (P3 > mx1) mx1 = P3;
  if (P2 < mn1) mn1 = P2; if (P3 < mn1) mn1 = P3;
  if (Q2 > mx2) mx2 = Q2; if (Q3 > mx2) mx2 = Q3;
  if (Q2 < mn2) mn2 = Q2; if (Q3 < mn2) mn2 = Q3;
  return (mn1 > mx2) || (mn2 > mx1);
}
__host__ __device__ inline bool tri_tri_intersect_exact(d3 a1, d3 a2, d3 a3, d3 b1, d3 b2, d3 b3) {
  d3 p1 = a1 - a1, p2 = a2 - a1, p3 = a3 - a1;
  d3 q1 = b1 - a1, q2 = b2 - a1, q3 = b3 - a1;
  d3 e1 = p2 - p1, e2 = p3 - p2, e3 = p1 - p3;
  d3 f1 = q2 - q1, f2 = q3 - q2, f3_ = q1 - q3;
  d3 n1 = cross(e1, e2), m1 = cross(f1, f2);
  if (axis_separates(n1, p1, p2, p3, q1, q2, q3)) return false;
  if (axis_separates(m1, p1, p2, p3, q1, q2, q3)) return false;
  if (axis_separates(cross(e1, f1), p1, p2, p3, q1, q2, q3)) return false;
  if (axis_separates(cross(e1, f2), p1, p2, p3, q1, q2, q3)) return false;
  if (axis_separates(cross(e1, f3_), p1, p2, p3, q1, q2, q3)) return false;
  if (axis_separates(cross(e2, f1), p1, p2, p3, q1, q2, q3)) return false;
  if (axis_separates(cross(e2, f2), p1, p2, p3, q1, q2, q3)) return false;
  if (axis_separates(cross(e2, f3_), p1, p2, p3, q1, q2, q3)) return false;
  if (axis_separates(cross(e3, f1), p1, p2, p3, q1, q2, q3)) return false;
  if (axis_separates(cross(e3, f2), p1, p2, p3, q1, q2, q3)) return false;
  if (axis_separates(cross(e3, f3_), p1, p2, p3, q1, q2, q3)) return false;
  if (axis_separates(cross(e1, n1), p1, p2, p3, q1, q2, q3)) return false;
  if (axis_separates(cross(e2, n1), p1, p2, p3, q1, q2, q3)) return false;
  if (axis_separates(cross(e3, n1), p1, p2, p3, q1, q2, q3)) return false;
  if (axis_separates(cross(f1, m1), p1, p2, p3, q1, q2, q3)) return false;
  if (axis_separates(cross(f2, m1), p1, p2, p3, q1, q2, q3)) return false;
  if (axis_separates(cross(f3_, m1), p1, p2, p3, q1, q2, q3)) return false;
  return true;
}

// ------------------------------------------------------------------------------------------
// Triangle-triangle minimum distance for two triangles known not to intersect.  Must agree with
// the non-intersecting branch of triangleTriangleDistanceSq (triangle_inl.h:406-518): minimum over
// 6 vertex-face and 9 edge-edge closest pairs, first minimum wins on ties (strict <).
// pa lies on triangle 1 (a*), pb on triangle 2 (b*).
// ------------------------------------------------------------------------------------------
template <typename T>
__host__ __device__ __forceinline__ T tri_tri_dist2(V3<T> a1, V3<T> a2, V3<T> a3, V3<T> b1, V3<T> b2, V3<T> b3,
                                                    V3<T> &pa, V3<T> &pb) {
  T dmin, d;
  V3<T> x, y;
  pa = closest_pt_triangle(a1, a2, a3, b1); pb = b1; { V3<T> v = pb - pa; dmin = dot(v, v); }
  x = closest_pt_triangle(a1, a2, a3, b2); { V3<T> v = b2 - x; d = dot(v, v); } if (d < dmin) { dmin = d; pa = x; pb = b2; }
  x = closest_pt_triangle(a1, a2, a3, b3); { V3<T> v = b3 - x; d = dot(v, v); } if (d < dmin) { dmin = d; pa = x; pb = b3; }
  y = closest_pt_triangle(b1, b2, b3, a1); { V3<T> v = y - a1; d = dot(v, v); } if (d < dmin) { dmin = d; pa = a1; pb = y; }
  y = closest_pt_triangle(b1, b2, b3, a2); { V3<T> v = y - a2; d = dot(v, v); } if (d < dmin) { dmin = d; pa = a2; pb = y; }
  y = closest_pt_triangle(b1, b2, b3, a3); { V3<T> v = y - a3; d = dot(v, v); } if (d < dmin) { dmin = d; pa = a3; pb = y; }
#define B2C_EE(P, Q, R, S) d = seg_seg_dist2(P, Q, R, S, x, y); if (d < dmin) { dmin = d; pa = x; pb = y; }
  B2C_EE(a1, a2, b1, b2) B2C_EE(a1, a2, b2, b3) B2C_EE(a1, a2, b3, b1)
  B2C_EE(a2, a3, b1, b2) B2C_EE(a2, a3, b2, b3) B2C_EE(a2, a3, b3, b1)
  B2C_EE(a3, a1, b1, b2) B2C_EE(a3, a1, b2, b3) B2C_EE(a3, a1, b3, b1)
#undef B2C_EE
  return dmin;
}

// ------------------------------------------------------------------------------------------
// Compact form of the same minimum for the traversal hot path.  The minimum distance of two disjoint
// triangles is attained either between two edges or between a vertex and the INTERIOR of the other
// face; a vertex whose foot falls outside the face is covered by the edges that end in it.  So the
// six vertex-face candidates reduce to a plane projection + an inside test, and all fifteen
// candidates run as rolled loops over rotated vertex registers (small code: the unrolled form above
// is ~60 KB of SASS once inlined in a traversal kernel and stalls on instruction fetch).
// Same candidate order as triangleTriangleDistanceSq (triangle_inl.h:406-518), strict < on ties.
// The reference treats an edge whose SQUARED length is <= 1e-3 as a point (triangle_inl.h:213,223), and
// then relies on its full vertex-face tests to cover what that edge would have found; when such an edge
// is met `degenerate` is set and the caller must use tri_tri_dist2 (the plain form) for that pair.
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ float fast_div(float a, float b) {
#ifdef __CUDA_ARCH__
  return __fdividef(a, b);
#else
  return a / b;
#endif
}
__host__ __device__ __forceinline__ double fast_div(double a, double b) { return a / b; }

template <typename T>
__host__ __device__ __forceinline__ T seg_seg_dist2_fast(V3<T> p1, V3<T> q1, V3<T> p2, V3<T> q2, V3<T> &c1, V3<T> &c2,
                                                         bool &degenerate) {
  const T EPS = T(1.0e-3);
  V3<T> d1 = q1 - p1, d2 = q2 - p2, r = p1 - p2;
  T a = dot(d1, d1), e = dot(d2, d2), f = dot(d2, r), c = dot(d1, r), b = dot(d1, d2);
  T s = T(0), t = T(0);
  const bool da = a <= EPS, de = e <= EPS;      // degenerate segments (EPSILON on the squared length)
  degenerate = degenerate || da || de;
  if (!(da && de)) {
    if (da) {
      t = clamp01(fast_div(f, e));
    } else if (de) {
      s = clamp01(fast_div(-c, a));
    } else {
      T denom = a * e - b * b;
      if (denom != T(0)) s = clamp01(fast_div(b * f - c * e, denom));
      T tn = b * s + f;
      if (tn < T(0)) { t = T(0); s = clamp01(fast_div(-c, a)); }
      else if (tn > e) { t = T(1); s = clamp01(fast_div(b - c, a)); }
      else t = fast_div(tn, e);
    }
  }
  c1 = p1 + d1 * s;
  c2 = p2 + d2 * t;
  V3<T> d = c1 - c2;
  return dot(d, d);
}

template <typename T>
__host__ __device__ __forceinline__ T tri_tri_dist2_compact(V3<T> a0, V3<T> a1, V3<T> a2, V3<T> b0, V3<T> b1, V3<T> b2,
                                                            V3<T> &pa, V3<T> &pb, bool &degenerate,
                                                            T bound = T(3.0e38)) {
  // `bound`: squared distance the caller already holds; when a triangle lies wholly on one side of the other's
  // plane, farther than that, no candidate of this pair can matter and the routine returns early (3e38)
  T dmin = T(3.0e38);
  degenerate = false;
  pa = a0; pb = b0;
  // vertex-face: side 0 = vertices of b against face a, side 1 = vertices of a against face b
#pragma unroll 1
  for (int side = 0; side < 2; side++) {
    V3<T> ab = a1 - a0, ac = a2 - a0;
    V3<T> n = cross(ab, ac);
    T nn = dot(n, n);
    if (nn > T(0)) {
      T inv = fast_div(T(1), nn);
      {
        T s0 = dot(b0 - a0, n), s1 = dot(b1 - a0, n), s2 = dot(b2 - a0, n);
        bool pos = s0 > T(0) && s1 > T(0) && s2 > T(0), neg = s0 < T(0) && s1 < T(0) && s2 < T(0);
        if ((pos || neg) && fmin(s0 * s0, fmin(s1 * s1, s2 * s2)) * inv > bound) return T(3.0e38);
      }
#pragma unroll 1
      for (int k = 0; k < 3; k++) {
        V3<T> ap = b0 - a0;
        T w2 = dot(cross(ab, ap), n);       // nn * barycentric weight of a2
        T w1 = dot(cross(ap, ac), n);       // nn * barycentric weight of a1
        if (w1 >= T(0) && w2 >= T(0) && w1 + w2 <= nn) {
          T dn = dot(ap, n);
          T d = dn * dn * inv;
          if (d < dmin) {
            dmin = d;
            V3<T> foot = b0 - n * (dn * inv);
            if (side == 0) { pa = foot; pb = b0; } else { pa = b0; pb = foot; }
          }
        }
        V3<T> tmp = b0; b0 = b1; b1 = b2; b2 = tmp;
      }
    }
    V3<T> t0 = a0, t1 = a1, t2 = a2;
    a0 = b0; a1 = b1; a2 = b2; b0 = t0; b1 = t1; b2 = t2;
  }
  // edge-edge: (a0a1, a1a2, a2a0) x (b0b1, b1b2, b2b0)
#pragma unroll 1
  for (int i = 0; i < 3; i++) {
#pragma unroll 1
    for (int j = 0; j < 3; j++) {
      V3<T> x, y;
      T d = seg_seg_dist2_fast(a0, a1, b0, b1, x, y, degenerate);
      if (d < dmin) { dmin = d; pa = x; pb = y; }
      V3<T> tmp = b0; b0 = b1; b1 = b2; b2 = tmp;
    }
    V3<T> tmp = a0; a0 = a1; a1 = a2; a2 = tmp;
  }
  return dmin;
}


} // namespace b2c
