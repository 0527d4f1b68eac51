// Host-side, non-throughput parts of the CollisionInterface surface (see contacts_host.h).
#include "contacts_host.h"
#include "obb_order.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <list>

namespace b2c {

namespace {

struct P3 {
  double x, y, z;
};
inline P3 operator+(P3 a, P3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline P3 operator-(P3 a, P3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline P3 operator*(double s, P3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dotp(P3 a, P3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline P3 crossp(P3 a, P3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double normp(P3 a) { return std::sqrt(dotp(a, a)); }
inline P3 unitp(P3 a) {
  double n2 = dotp(a, a);
  if (n2 > 0.0) { double n = std::sqrt(n2); return {a.x / n, a.y / n, a.z / n}; }
  return a;
}
inline double comp(const P3 &a, int i) { return i == 0 ? a.x : (i == 1 ? a.y : a.z); }

struct M3 {
  double m[3][3];
  P3 col(int c) const { return {m[0][c], m[1][c], m[2][c]}; }
  void setcol(int c, P3 v) { m[0][c] = v.x; m[1][c] = v.y; m[2][c] = v.z; }
  P3 mul(P3 v) const {
    return {m[0][0] * v.x + m[0][1] * v.y + m[0][2] * v.z, m[1][0] * v.x + m[1][1] * v.y + m[1][2] * v.z,
            m[2][0] * v.x + m[2][1] * v.y + m[2][2] * v.z};
  }
};
inline M3 matmul(const M3 &a, const M3 &b) {
  M3 r;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      double s = 0.0;
      for (int k = 0; k < 3; k++) s += a.m[i][k] * b.m[k][j];
      r.m[i][j] = s;
    }
  return r;
}

// rotation matrix of the unit quaternion for an axis/angle rotation (transf::AXIS_ANGLE_ROTATION)
M3 axis_angle_matrix(double angle, P3 axis) {
  axis = unitp(axis);
  double s = std::sin(0.5 * angle), w = std::cos(0.5 * angle);
  double x = s * axis.x, y = s * axis.y, z = s * axis.z;
  double n = std::sqrt(w * w + x * x + y * y + z * z);
  w /= n; x /= n; y /= n; z /= n;
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y,
               tzz = tz * z;
  M3 r;
  r.m[0][0] = 1 - (tyy + tzz); r.m[0][1] = txy - twz; r.m[0][2] = txz + twy;
  r.m[1][0] = txy + twz; r.m[1][1] = 1 - (txx + tzz); r.m[1][2] = tyz - twx;
  r.m[2][0] = txz - twy; r.m[2][1] = tyz + twx; r.m[2][2] = 1 - (txx + tyy);
  return r;
}

// rotation matrix -> unit quaternion (wxyz), trace-branch algorithm (what transf::set(mat3) gets from Eigen)
void matrix_to_quat(const M3 &R, double q[4]) {
  double t = R.m[0][0] + R.m[1][1] + R.m[2][2];
  double c[4];   // x y z w
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    c[3] = 0.5 * t;
    t = 0.5 / t;
    c[0] = (R.m[2][1] - R.m[1][2]) * t;
    c[1] = (R.m[0][2] - R.m[2][0]) * t;
    c[2] = (R.m[1][0] - R.m[0][1]) * t;
  } else {
    int i = 0;
    if (R.m[1][1] > R.m[0][0]) i = 1;
    if (R.m[2][2] > R.m[i][i]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(R.m[i][i] - R.m[j][j] - R.m[k][k] + 1.0);
    c[i] = 0.5 * t;
    t = 0.5 / t;
    c[3] = (R.m[k][j] - R.m[j][k]) * t;
    c[j] = (R.m[j][i] + R.m[i][j]) * t;
    c[k] = (R.m[k][i] + R.m[i][k]) * t;
  }
  double n = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2] + c[3] * c[3]);
  q[0] = c[3] / n; q[1] = c[0] / n; q[2] = c[1] / n; q[3] = c[2] / n;
}

// Classic cyclic-by-largest-element Jacobi eigen-solver for a symmetric 3x3 (Golub & Van Loan 8.4),
// with the reference's stopping rule: at most 50 rotations, stop once the off-diagonal norm no longer
// decreases after the third rotation (collisionModel.cpp:612-679); Schur rotation skipped when the
// pivot is below 1e-4 (:585-602).
void jacobi3(double a[3][3], double v[3][3]) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) v[i][j] = (i == j) ? 1.0 : 0.0;
  double prevoff = std::numeric_limits<double>::max();
  for (int n = 0; n < 50; n++) {
    int p = 0, q = 1;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        if (i == j) continue;
        if (std::fabs(a[i][j]) > std::fabs(a[p][q])) { p = i; q = j; }
      }
    double c = 1.0, s = 0.0;
    if (std::fabs(a[p][q]) > 0.0001f) {
      double r = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
      double t = (r >= 0.0) ? 1.0 / (r + std::sqrt(1.0 + r * r)) : -1.0 / (-r + std::sqrt(1.0 + r * r));
      c = 1.0 / std::sqrt(1.0 + t * t);
      s = t * c;
    }
    double J[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
    J[p][p] = c; J[p][q] = s; J[q][p] = -s; J[q][q] = c;
    double t1[3][3], t2[3][3];
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        t1[i][j] = 0.0;
        for (int k = 0; k < 3; k++) t1[i][j] += v[i][k] * J[k][j];
      }
    std::memcpy(v, t1, sizeof t1);
    // a = (J^T a) J
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        t2[i][j] = 0.0;
        for (int k = 0; k < 3; k++) t2[i][j] += J[k][i] * a[k][j];
      }
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        t1[i][j] = 0.0;
        for (int k = 0; k < 3; k++) t1[i][j] += t2[i][k] * J[k][j];
      }
    std::memcpy(a, t1, sizeof t1);
    double off = 0.0;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        if (i != j) off += a[i][j] * a[i][j];
    if (n > 2 && off >= prevoff) return;
    prevoff = off;
  }
}

struct TriD {
  P3 v[3];
};

const double kPad = 1.0e-2;   // Leaf::TOLERANCE

struct ObbNode {
  M3 R;
  P3 center;
  P3 half;
  int child[2] = {-1, -1};
  int tri = -1;               // leaf: its triangle
  int leaf_lo = 0, leaf_hi = 0;   // leaves below this node, as a range of depth-first leaf positions
  M3 Rq;                      // R after the round trip through the unit quaternion transf stores (BoundingBox::setTran)
};

} // namespace

struct ObbTree {
  std::vector<ObbNode> nodes;
  std::vector<int> tri_pos;   // triangle -> depth-first position of its leaf
  std::vector<OrderNode> order;   // what the visit-order rule reads of every node (obb_order.h), same indices as nodes
};

namespace {

void fit_obb(const std::vector<TriD> &tris, const std::vector<int> &idx, ObbNode &out) {
  // area-weighted covariance (collisionModel.cpp:134-166)
  double cov[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
  double total = 0.0;
  P3 wm = {0, 0, 0};
  for (int t : idx) {
    const TriD &T = tris[t];
    P3 m = (1.0 / 3.0) * (T.v[0] + T.v[1] + T.v[2]);
    double area = 0.5 * normp(crossp(T.v[1] - T.v[0], T.v[2] - T.v[0]));
    total += area;
    wm = wm + area * m;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        cov[i][j] += (area / 12.0) * (9 * comp(m, i) * comp(m, j) + comp(T.v[0], i) * comp(T.v[0], j) +
                                      comp(T.v[1], i) * comp(T.v[1], j) + comp(T.v[2], i) * comp(T.v[2], j));
  }
  wm = (1.0 / total) * wm;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) cov[i][j] = (1.0 / total) * cov[i][j] - comp(wm, i) * comp(wm, j);
  double v[3][3];
  jacobi3(cov, v);
  // order the axes by eigenvalue with the reference's swap sequence (collisionModel.cpp:179-188)
  int first = 0, second = 1, third = 2;
  if (cov[1][1] > cov[0][0]) std::swap(first, second);
  if (cov[2][2] > cov[first][first]) std::swap(first, third);
  if (cov[third][third] > cov[second][second]) std::swap(second, third);
  P3 xA = {v[0][first], v[1][first], v[2][first]};
  P3 yA = {v[0][second], v[1][second], v[2][second]};
  P3 zA = crossp(unitp(xA), unitp(yA));
  yA = crossp(zA, unitp(xA));
  xA = crossp(yA, zA);
  M3 R;
  R.setcol(0, xA); R.setcol(1, yA); R.setcol(2, zA);
  // extents along the axes, padded (collisionModel.cpp:82-131)
  double mx[3] = {-1.0e10, -1.0e10, -1.0e10}, mn[3] = {1.0e10, 1.0e10, 1.0e10};
  for (int t : idx)
    for (int k = 0; k < 3; k++) {
      const P3 &p = tris[t].v[k];
      double d[3] = {dotp(p, xA), dotp(p, yA), dotp(p, zA)};
      for (int a = 0; a < 3; a++) {
        if (d[a] + kPad > mx[a]) mx[a] = d[a] + kPad;
        if (d[a] - kPad < mn[a]) mn[a] = d[a] - kPad;
      }
    }
  double h[3];
  for (int a = 0; a < 3; a++) h[a] = 0.5 * (mx[a] - mn[a]);
  P3 c = {mn[0] + h[0], mn[1] + h[1], mn[2] + h[2]};
  c = R.mul(c);
  for (int a = 0; a < 3; a++)
    if (h[a] < kPad) h[a] = kPad;
  // x axis := direction of largest extent (collisionModel.cpp:211-229)
  int big = 0;
  if (h[1] > h[0]) big = 1;
  if (h[2] > h[big]) big = 2;
  P3 hv = {h[0], h[1], h[2]};
  if (big != 0) {
    M3 rot = axis_angle_matrix(M_PI / 2.0, big == 1 ? P3{0, 0, 1} : P3{0, 1, 0});
    hv = rot.mul(hv);
    hv = {std::fabs(hv.x), std::fabs(hv.y), std::fabs(hv.z)};
    R = matmul(R, rot);
  }
  out.R = R;
  out.center = c;
  out.half = hv;
}

void split_node(ObbTree &T, const std::vector<TriD> &tris, int node, std::vector<int> idx) {
  if ((int)idx.size() <= 1) { T.nodes[node].tri = idx.empty() ? -1 : idx[0]; return; }
  ObbNode nd = T.nodes[node];
  P3 xA = nd.R.col(0), yA = nd.R.col(1), zA = nd.R.col(2);
  if (nd.half.y > nd.half.x + 1.0e-5) { std::swap(xA, yA); yA = -1.0 * yA; }
  if (nd.half.z > nd.half.y && nd.half.z > nd.half.x + 1.0e-5) { std::swap(xA, zA); zA = -1.0 * zA; }
  // split plane at the mean vertex (collisionModel.cpp:46-60,284)
  P3 mean = {0, 0, 0};
  for (int t : idx) { mean = mean + tris[t].v[0]; mean = mean + tris[t].v[1]; mean = mean + tris[t].v[2]; }
  mean = (1.0 / (3.0 * (int)idx.size())) * mean;
  double sep = dotp(mean, xA);
  std::vector<int> c1, c2;
  auto balanced = [&](double s) {
    c1.clear(); c2.clear();
    for (int t : idx) {
      P3 med = {0, 0, 0};
      med = med + tris[t].v[0]; med = med + tris[t].v[1]; med = med + tris[t].v[2];
      med = (1.0 / 3.0) * med;
      if (dotp(med, xA) < s) c1.push_back(t); else c2.push_back(t);
    }
  };
  balanced(sep);
  if (c1.empty() || c2.empty()) {
    // median of the centroid projections (collisionModel.cpp:62-80)
    std::vector<double> pr;
    for (int t : idx) {
      P3 m = tris[t].v[0] + tris[t].v[1] + tris[t].v[2];
      m = (1.0 / 3.0) * m;
      pr.push_back(dotp(m, xA));
    }
    double med = pr.front();
    if (pr.size() > 1) {
      std::nth_element(pr.begin(), pr.begin() + pr.size() / 2, pr.end());
      med = pr[pr.size() / 2];
    }
    balanced(med);
  }
  if (c1.empty() || c2.empty()) {
    c1.clear(); c2.clear();
    bool one = true;
    for (int t : idx) { (one ? c1 : c2).push_back(t); one = !one; }
  }
  int a = (int)T.nodes.size();
  T.nodes.emplace_back();
  T.nodes.emplace_back();
  fit_obb(tris, c1, T.nodes[a]);
  fit_obb(tris, c2, T.nodes[a + 1]);
  T.nodes[node].child[0] = a;
  T.nodes[node].child[1] = a + 1;
  split_node(T, tris, a, std::move(c1));
  split_node(T, tris, a + 1, std::move(c2));
}

void level_rec(const ObbTree &T, int node, int cur, int want, std::vector<ObbBox> &out) {
  const ObbNode &n = T.nodes[node];
  bool leaf = n.child[0] < 0;
  if (cur == want || leaf) {
    ObbBox b;
    b.t[0] = n.center.x; b.t[1] = n.center.y; b.t[2] = n.center.z;
    b.half[0] = n.half.x; b.half[1] = n.half.y; b.half[2] = n.half.z;
    matrix_to_quat(n.R, b.q);
    out.push_back(b);
  }
  if (!leaf && cur < want) {
    level_rec(T, n.child[0], cur + 1, want, out);
    level_rec(T, n.child[1], cur + 1, want, out);
  }
}

} // namespace

ObbTree *obb_tree_build(const float *tri9, int ntri) {
  ObbTree *T = new ObbTree();
  std::vector<TriD> tris(ntri);
  for (int i = 0; i < ntri; i++)
    for (int k = 0; k < 3; k++)
      tris[i].v[k] = {(double)tri9[9 * (size_t)i + 3 * k], (double)tri9[9 * (size_t)i + 3 * k + 1], (double)tri9[9 * (size_t)i + 3 * k + 2]};
  std::vector<int> idx(ntri);
  for (int i = 0; i < ntri; i++) idx[i] = i;
  T->nodes.reserve(2 * (size_t)std::max(ntri, 1));
  T->nodes.emplace_back();
  if (ntri > 0) {
    fit_obb(tris, idx, T->nodes[0]);
    split_node(*T, tris, 0, idx);
  } else {
    ObbNode &n = T->nodes[0];
    n.R = M3{{{1, 0, 0}, {0, 1, 0}, {0, 0, 1}}};
    n.center = {0, 0, 0};
    n.half = {kPad, kPad, kPad};
  }
  // depth-first leaf positions (explicit stack: degenerate meshes give deep trees) and the rotation each box really
  // carries: BoundingBox keeps a transf, i.e. the unit quaternion of R (bBox.h:38-59)
  T->tri_pos.assign(std::max(ntri, 1), 0);
  {
    int next = 0;
    std::vector<std::pair<int, int>> st;      // (node, state)
    st.push_back({0, 0});
    while (!st.empty()) {
      auto &top = st.back();
      ObbNode &n = T->nodes[top.first];
      if (n.child[0] < 0) {
        n.leaf_lo = next; n.leaf_hi = next + 1;
        if (n.tri >= 0) T->tri_pos[n.tri] = next;
        next++;
        st.pop_back();
      } else if (top.second == 0) {
        n.leaf_lo = next; top.second = 1; st.push_back({n.child[0], 0});
      } else if (top.second == 1) {
        top.second = 2; st.push_back({n.child[1], 0});
      } else {
        n.leaf_hi = next; st.pop_back();
      }
    }
  }
  for (ObbNode &n : T->nodes) {
    double q[4];
    matrix_to_quat(n.R, q);
    const double w = q[0], x = q[1], y = q[2], z = q[3];
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    n.Rq.m[0][0] = 1 - (tyy + tzz); n.Rq.m[0][1] = txy - twz; n.Rq.m[0][2] = txz + twy;
    n.Rq.m[1][0] = txy + twz; n.Rq.m[1][1] = 1 - (txx + tzz); n.Rq.m[1][2] = tyz - twx;
    n.Rq.m[2][0] = txz - twy; n.Rq.m[2][1] = tyz + twx; n.Rq.m[2][2] = 1 - (txx + tyy);
  }
  T->order.resize(T->nodes.size());
  for (size_t i = 0; i < T->nodes.size(); i++) {
    const ObbNode &n = T->nodes[i];
    OrderNode &o = T->order[i];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) o.Rq[r][c] = n.Rq.m[r][c];
    o.center[0] = n.center.x; o.center[1] = n.center.y; o.center[2] = n.center.z;
    o.half[0] = n.half.x; o.half[1] = n.half.y; o.half[2] = n.half.z;
    o.child0 = n.child[0]; o.child1 = n.child[1]; o.leaf_lo = n.leaf_lo; o.leaf_hi = n.leaf_hi;
  }
  return T;
}

// -------------------------------------------------------------------------------------------------
// Reference visit order of contact candidates.
//
// ContactCallback appends contacts in the order its dual-tree recursion reaches the leaf pairs
// (collisionAlgorithms.cpp:42-130), and removeContactDuplicates / compactContactSet depend on that order
// (collisionInterface.cpp:74-287: the survivor of a 3 mm cluster, first-group-wins).  The engine finds the
// candidate triangle pairs on the device in no particular order; this replays the reference's recursion over
// the reference-style OBB trees, restricted to node pairs that are ancestors of a candidate:
//   * the node that is split: the one that is not a leaf; of two branches the one with the larger box
//     volume (product of half sizes), ties -> the second model's (collisionAlgorithms.cpp:60-79);
//   * the child visited first: the one with the smaller bboxDistanceSq to the other node, ties -> child1
//     (collisionAlgorithms.cpp:90-101; bboxDistanceSq, bbox_inl.h:269-413, returns -1 for overlapping boxes).
// The box distance is only evaluated where BOTH children hold candidates (elsewhere the order is forced).
// -------------------------------------------------------------------------------------------------
namespace {

struct VisitCtx {
  const ObbTree *A, *B;
  OrderXf x;
  const int *ta, *tb;
  std::vector<int> *out;
  std::vector<int> scratch;
};

// idx[lo, hi): candidates below the node pair (n1, n2)
void visit_pair(VisitCtx &c, int n1, int n2, int *idx, int lo, int hi) {
  if (hi <= lo) return;
  // leaf-position ranges of the candidates in both trees: "every candidate sits under the same child" is then two
  // comparisons, so the long single-file descents between real branchings cost O(1) per level instead of a pass over
  // the candidates (a group holds ~40 candidates under ~30 levels of which a handful branch)
  int amin = 0x7fffffff, amax = -1, bmin = 0x7fffffff, bmax = -1;
  for (int i = lo; i < hi; i++) {
    const int pa = c.A->tri_pos[c.ta[idx[i]]], pb = c.B->tri_pos[c.tb[idx[i]]];
    amin = std::min(amin, pa); amax = std::max(amax, pa);
    bmin = std::min(bmin, pb); bmax = std::max(bmax, pb);
  }
  const OrderNode *NA = c.A->order.data(), *NB = c.B->order.data();
  while (true) {
    const OrderNode &N1 = NA[n1], &N2 = NB[n2];
    const bool l1 = N1.child0 < 0, l2 = N2.child0 < 0;
    if (l1 && l2) {
      for (int i = lo; i < hi; i++) c.out->push_back(idx[i]);
      return;
    }
    bool split1;             // true: recurse on model 1's node (nodeSwap == false)
    if (l1) split1 = false;
    else if (l2) split1 = true;
    else split1 = order_box_volume(N1) > order_box_volume(N2);
    const OrderNode &R = split1 ? N1 : N2;
    const OrderNode *T = split1 ? NA : NB;
    const std::vector<int> &tpos = split1 ? c.A->tri_pos : c.B->tri_pos;
    const OrderNode &C0 = T[R.child0];
    const int pmin = split1 ? amin : bmin, pmax = split1 ? amax : bmax;
    if (pmin >= C0.leaf_lo && pmax < C0.leaf_hi) {           // everything sits under child 1: no order to decide
      if (split1) n1 = R.child0; else n2 = R.child0;
      continue;
    }
    if (pmin >= C0.leaf_hi || pmax < C0.leaf_lo) {           // everything sits under child 2
      if (split1) n1 = R.child1; else n2 = R.child1;
      continue;
    }
    // a real branching: stable split of the candidates between the two children
    const int *tri = split1 ? c.ta : c.tb;
    if ((int)c.scratch.size() < hi - lo) c.scratch.resize(hi - lo);
    int k0 = lo, k1 = 0;
    for (int i = lo; i < hi; i++) {
      const int pos = tpos[tri[idx[i]]];
      if (pos >= C0.leaf_lo && pos < C0.leaf_hi) idx[k0++] = idx[i]; else c.scratch[k1++] = idx[i];
    }
    for (int i = 0; i < k1; i++) idx[k0 + i] = c.scratch[i];
    const int mid = k0;
    double d0, d1;
    if (split1) { d0 = order_distance_sq(C0, N2, c.x); d1 = order_distance_sq(T[R.child1], N2, c.x); }
    else { d0 = order_distance_sq(N1, C0, c.x); d1 = order_distance_sq(N1, T[R.child1], c.x); }
    const int first = d1 < d0 ? 1 : 0;
    const int f_lo = first == 0 ? lo : mid, f_hi = first == 0 ? mid : hi;
    const int s_lo = first == 0 ? mid : lo, s_hi = first == 0 ? hi : mid;
    const int fc = first == 0 ? R.child0 : R.child1, sc = first == 0 ? R.child1 : R.child0;
    if (split1) { visit_pair(c, fc, n2, idx, f_lo, f_hi); visit_pair(c, sc, n2, idx, s_lo, s_hi); }
    else { visit_pair(c, n1, fc, idx, f_lo, f_hi); visit_pair(c, n1, sc, idx, s_lo, s_hi); }
    return;
  }
}

} // namespace

void obb_visit_order(const ObbTree *A, const ObbTree *B, const double R21[9], const double t21[3], const int *tri_a,
                     const int *tri_b, int n, std::vector<int> &order) {
  order.clear();
  if (n <= 0) return;
  order.reserve(n);
  VisitCtx c;
  c.A = A; c.B = B; c.ta = tri_a; c.tb = tri_b; c.out = &order;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) c.x.R[i][j] = R21[3 * i + j];
  for (int i = 0; i < 3; i++) c.x.t[i] = t21[i];
  std::vector<int> idx(n);
  for (int i = 0; i < n; i++) idx[i] = i;
  visit_pair(c, 0, 0, idx.data(), 0, n);
}

const OrderNode *obb_order_nodes(const ObbTree *t, int *count) { if (count) *count = (int)t->order.size(); return t->order.data(); }
const int *obb_tri_positions(const ObbTree *t, int *count) { if (count) *count = (int)t->tri_pos.size(); return t->tri_pos.data(); }

void obb_tree_free(ObbTree *t) { delete t; }

void obb_tree_level(const ObbTree *t, int depth, std::vector<ObbBox> &out) {
  out.clear();
  if (!t || t->nodes.empty()) return;
  level_rec(*t, 0, 0, depth, out);
}

// -------------------------------------------------------------------------------------------------
// contact post-processing
// -------------------------------------------------------------------------------------------------
static inline double d2(const double a[3], const double b[3]) {
  double s = 0.0;
  for (int k = 0; k < 3; k++) s += (a[k] - b[k]) * (a[k] - b[k]);
  return s;
}

void remove_contact_duplicates(std::vector<b2c_contact> &c, double thr) {
  // The reference erases from its list while it scans (collisionInterface.cpp:254-287); the same scan over "alive" flags
  // visits the same elements in the same order and compacts once at the end (erasing 104-byte records from the middle of a
  // vector was the bulk of the batched post-processing time).
  const double t2 = thr * thr;
  const size_t n = c.size();
  if (n < 2) return;
  std::vector<unsigned char> alive(n, 1);
  for (size_t i = 0; i < n; i++) {
    if (!alive[i]) continue;
    for (size_t j = i + 1; j < n; j++) {
      if (!alive[j]) continue;
      if (d2(c[j].b1_pos, c[i].b1_pos) > t2 || d2(c[j].b2_pos, c[i].b2_pos) > t2) continue;
      if (c[i].distSq < c[j].distSq) {
        alive[j] = 0;
      } else {
        alive[i] = 0;
        break;
      }
    }
  }
  size_t k = 0;
  for (size_t i = 0; i < n; i++)
    if (alive[i]) { if (k != i) c[k] = c[i]; k++; }
  c.resize(k);
}

namespace {

// 2-D convex hull vertices, indices into pts.  The reference asks qhull for the hull of the projected
// contacts (collisionInterface.cpp:165-219, options "qhull n Pp"); qhull's default facet merging drops
// points that sit on a hull edge to within its round-off bound (qh_distround, a few hundred ulps of the
// largest coordinate), so a point is a vertex here only if it sticks out of the line through its hull
// neighbours by more than that bound.
std::vector<int> hull2d(const std::vector<std::pair<double, double>> &pts) {
  int n = (int)pts.size();
  std::vector<int> order(n);
  for (int i = 0; i < n; i++) order[i] = i;
  std::sort(order.begin(), order.end(), [&](int a, int b) { return pts[a] < pts[b]; });
  double maxabs = 0.0;
  for (auto &p : pts) maxabs = std::max(maxabs, std::max(std::fabs(p.first), std::fabs(p.second)));
  const double tol = 256.0 * std::numeric_limits<double>::epsilon() * std::max(maxabs, 1.0);
  // true if `a` does not stick out to the right of the directed line o -> b by more than tol
  auto drop = [&](int o, int a, int b) {
    double bx = pts[b].first - pts[o].first, by = pts[b].second - pts[o].second;
    double cr = (pts[a].first - pts[o].first) * by - (pts[a].second - pts[o].second) * bx;   // > 0: a right of o->b
    double len = std::sqrt(bx * bx + by * by);
    return cr <= tol * len;
  };
  std::vector<int> h(2 * n);
  int k = 0;
  for (int i = 0; i < n; i++) {
    while (k >= 2 && drop(h[k - 2], h[k - 1], order[i])) k--;
    h[k++] = order[i];
  }
  for (int i = n - 2, t = k + 1; i >= 0; i--) {
    while (k >= t && drop(h[k - 2], h[k - 1], order[i])) k--;
    h[k++] = order[i];
  }
  h.resize(k > 1 ? k - 1 : k);
  return h;
}

void perimeter(std::vector<b2c_contact> &set) {
  if (set.size() < 2) return;
  const double resabs = 1.0e-1;
  auto P = [&](size_t i) { return P3{set[i].b1_pos[0], set[i].b1_pos[1], set[i].b1_pos[2]}; };
  size_t e1 = 0, e2 = 1, cp = 0;
  P3 curr = {0, 0, 0}, test = {0, 0, 0};
  for (cp = 0; cp < set.size(); cp++) {
    curr = P(e2) - P(e1);
    test = P(cp) - P(e1);
    if (normp(crossp(test, curr)) > resabs) break;
    double dt = dotp(test, curr);
    if (dt < 0) e1 = cp;
    if (dt > dotp(curr, curr)) e2 = cp;
  }
  if (cp == set.size()) {
    std::vector<b2c_contact> tmp = {set[e1], set[e2]};
    set = tmp;
    return;
  }
  P3 normal = unitp(crossp(test, curr));
  P3 axis1 = unitp(test);
  P3 axis2 = crossp(normal, axis1);
  std::vector<std::pair<double, double>> pts;
  for (size_t i = 0; i < set.size(); i++) pts.push_back({dotp(P(i), axis1), dotp(P(i), axis2)});
  std::vector<int> h = hull2d(pts);
  std::vector<b2c_contact> tmp;
  for (int i : h) tmp.push_back(set[i]);
  set = tmp;
}

} // namespace

void compact_contact_set(std::vector<b2c_contact> &contacts) {
  if (contacts.size() < 2) return;
  const double NORMAL_TOL = 1e-3, DIST_TOL = 1.0e-1;
  std::list<std::vector<b2c_contact>> groups;
  for (auto &c : contacts) {
    auto sp = groups.begin();
    for (; sp != groups.end(); ++sp) {
      const b2c_contact &f = sp->front();
      double d = f.b1_normal[0] * c.b1_normal[0] + f.b1_normal[1] * c.b1_normal[1] + f.b1_normal[2] * c.b1_normal[2];
      if (d > 1.0 - NORMAL_TOL) break;
    }
    if (sp == groups.end()) {
      groups.push_back({c});
    } else {
      bool dup = false;
      for (auto &o : *sp)
        if (d2(o.b1_pos, c.b1_pos) < DIST_TOL) { dup = true; break; }
      if (!dup) sp->push_back(c);
    }
  }
  contacts.clear();
  for (auto &g : groups) {
    if (g.size() > 1) perimeter(g);
    for (auto &c : g) contacts.push_back(c);
  }
}

// ---------------------------------------------------------------------------------------------
// Contact frames and point-contact friction cones (World::findAllContacts -> addContacts,
// reference src/contactSetting.cpp:54-81,193-221; Contact::Contact, src/contact/contact.cpp:67-99;
// PointContact::setUpFrictionEdges, src/contact/pointContact.cpp:44-56; Contact::setUpFrictionEllipsoid,
// contact.cpp:180-216; Contact::wrenchFromFrictionEdge / computeWrenches, contact.cpp:124-165).
// Closed-form and tiny per contact: host work.
// ---------------------------------------------------------------------------------------------
namespace {
struct V3d { double x, y, z; };
inline V3d vsub(V3d a, V3d b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3d vadd(V3d a, V3d b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3d vmul(V3d a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double vdot(V3d a, V3d b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3d vcross(V3d a, V3d b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline V3d vnormalized(V3d a) {           // Eigen's normalized(): unchanged when the squared norm is 0
  double n2 = vdot(a, a);
  return n2 > 0.0 ? vmul(a, 1.0 / std::sqrt(n2)) : a;
}
inline V3d qrot(const double q[4], V3d v) {   // Eigen quaternion * vector: v + w*2(u x v) + u x 2(u x v)
  V3d u = {q[1], q[2], q[3]};
  V3d uv = vcross(u, v); uv = vadd(uv, uv);
  return vadd(vadd(v, vmul(uv, q[0])), vcross(u, uv));
}
} // namespace

int contact_frames(const b2c_contact *c, int n, int side, double *frame7, double *tangents6) {
  for (int i = 0; i < n; i++) {
    const double *pos = side == 1 ? c[i].b1_pos : c[i].b2_pos;
    const double *nrm = side == 1 ? c[i].b1_normal : c[i].b2_normal;
    V3d normal = vnormalized({nrm[0], nrm[1], nrm[2]});
    V3d tx, ty;
    // MACHINE_ZERO = 1e-9 (matvec3D.h:53)
    if (std::fabs(vdot(normal, {1, 0, 0})) > 1.0 - 1.0e-9) tx = vnormalized(vcross(normal, {0, 0, 1}));
    else tx = vnormalized(vcross(normal, {1, 0, 0}));
    ty = vnormalized(vcross(normal, tx));
    // transf::COORDINATE(loc, tangentX, tangentY) (transform.cpp:40-50)
    V3d nx = vnormalized(tx);
    V3d nz = vnormalized(vcross(nx, ty));
    V3d ny = vnormalized(vcross(nz, nx));
    M3 R;
    R.m[0][0] = nx.x; R.m[1][0] = nx.y; R.m[2][0] = nx.z;
    R.m[0][1] = ny.x; R.m[1][1] = ny.y; R.m[2][1] = ny.z;
    R.m[0][2] = nz.x; R.m[1][2] = nz.y; R.m[2][2] = nz.z;
    double q[4];
    matrix_to_quat(R, q);
    double *f = frame7 + 7 * (size_t)i;
    f[0] = q[0]; f[1] = q[1]; f[2] = q[2]; f[3] = q[3]; f[4] = pos[0]; f[5] = pos[1]; f[6] = pos[2];
    if (tangents6) {
      // what wrenchFromFrictionEdge reads back: frame.affine().col(0), col(1) -- transf::set(mat3) keeps R = the matrix
      double *t = tangents6 + 6 * (size_t)i;
      t[0] = nx.x; t[1] = nx.y; t[2] = nx.z; t[3] = ny.x; t[4] = ny.y; t[5] = ny.z;
    }
  }
  return n;
}

void friction_ellipsoid_edges(int nlat, const int *ndirs, const double *phi, const double *eccen, std::vector<double> &edges) {
  edges.clear();
  for (int i = 0; i < nlat; i++) {
    double cosphi = std::cos(phi[i]), sinphi = std::sin(phi[i]);
    for (int j = 0; j < ndirs[i]; j++) {
      double theta = j * 2 * M_PI / ndirs[i];
      double num = std::cos(theta) * cosphi;
      double denom = num * num / (eccen[0] * eccen[0]);
      num = std::sin(theta) * cosphi;
      denom += num * num / (eccen[1] * eccen[1]);
      num = sinphi;
      denom += num * num / (eccen[2] * eccen[2]);
      denom = std::sqrt(denom);
      double e[6] = {std::cos(theta) * cosphi / denom, std::sin(theta) * cosphi / denom, 0, 0, 0, sinphi / denom};
      edges.insert(edges.end(), e, e + 6);
    }
  }
}

int point_contact_wrenches(const b2c_contact *c, int n, int side, double cof, const double cog[3], double max_radius, double *wrench48) {
  // PCWF: one latitude (phi = 0), 8 directions, unit eccentricities (pointContact.cpp:44-56)
  const int ndirs[1] = {8};
  const double phi[1] = {0.0}, eccen[3] = {1, 1, 1};
  std::vector<double> edges;
  friction_ellipsoid_edges(1, ndirs, phi, eccen, edges);
  std::vector<double> frames(7 * (size_t)n), tang(6 * (size_t)n);
  contact_frames(c, n, side, frames.data(), tang.data());
  for (int i = 0; i < n; i++) {
    const double *pos = side == 1 ? c[i].b1_pos : c[i].b2_pos;
    const double *nrm = side == 1 ? c[i].b1_normal : c[i].b2_normal;
    V3d normal = vnormalized({nrm[0], nrm[1], nrm[2]});
    V3d tx = {tang[6 * i], tang[6 * i + 1], tang[6 * i + 2]}, ty = {tang[6 * i + 3], tang[6 * i + 4], tang[6 * i + 5]};
    V3d radius = vsub({pos[0], pos[1], pos[2]}, {cog[0], cog[1], cog[2]});
    for (int k = 0; k < 8; k++) {
      const double *e = &edges[6 * k];
      V3d force = vadd(vadd(normal, vmul(tx, cof * e[0])), vmul(ty, cof * e[1]));
      V3d torque = vmul(vadd(vcross(radius, force), vmul(normal, cof * e[5])), 1.0 / max_radius);
      double *w = wrench48 + 48 * (size_t)i + 6 * k;
      w[0] = force.x; w[1] = force.y; w[2] = force.z; w[3] = torque.x; w[4] = torque.y; w[5] = torque.z;
    }
  }
  return n;
}

int check_contact_normals(b2c_contact *c, int n, const double q1[4], const double t1[3], const double q2[4], const double t2[3]) {
  int flipped = 0;
  for (int i = 0; i < n; i++) {
    V3d n1 = qrot(q1, {c[i].b1_normal[0], c[i].b1_normal[1], c[i].b1_normal[2]});
    V3d n2 = qrot(q2, {c[i].b2_normal[0], c[i].b2_normal[1], c[i].b2_normal[2]});
    V3d p1 = vadd(qrot(q1, {c[i].b1_pos[0], c[i].b1_pos[1], c[i].b1_pos[2]}), {t1[0], t1[1], t1[2]});
    V3d p2 = vadd(qrot(q2, {c[i].b2_pos[0], c[i].b2_pos[1], c[i].b2_pos[2]}), {t2[0], t2[1], t2[2]});
    V3d d = vsub(p2, p1);
    bool ok = true;
    if (vdot(d, n1) > 0) { ok = false; for (int k = 0; k < 3; k++) c[i].b1_normal[k] = -c[i].b1_normal[k]; }
    if (vdot(d, n2) < 0) { ok = false; for (int k = 0; k < 3; k++) c[i].b2_normal[k] = -c[i].b2_normal[k]; }
    if (!ok) flipped++;
  }
  return flipped;
}


} // namespace b2c
