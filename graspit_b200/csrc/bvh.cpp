// Host-side construction of the flat device BVH (K2).
//
// The reference builds an oriented-box tree with one triangle per leaf (CollisionModel::build,
// src/Collision/Graspit/collisionModel.cpp:517-539).  Query results do not depend on the shape of a
// conservative hierarchy, so the device tree is chosen for the GPU instead:
// axis-aligned boxes in the body frame (24 B of geometry per node, one shared rotation per body
// pair), object-median splits (balanced: depth = ceil(log2 N), which bounds every traversal stack),
// one triangle per leaf, nodes stored breadth-first with siblings adjacent so that a prefix of the
// array is the top of the tree.
#include "bvh.h"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <queue>

namespace b2c {

namespace {
struct Build {
  float lo[3], hi[3];
  int left, right;   // indices into the temporary array, -1 for leaves
  int tri;
  int count;
};

struct Builder {
  const float *tri9;
  std::vector<float> cen;         // 3 per triangle
  std::vector<int> order;
  std::vector<Build> tmp;

  int build(int lo, int hi) {
    Build b;
    for (int k = 0; k < 3; k++) { b.lo[k] = 1e30f; b.hi[k] = -1e30f; }
    float clo[3] = {1e30f, 1e30f, 1e30f}, chi[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = lo; i < hi; i++) {
      const float *t = tri9 + 9 * (size_t)order[i];
      for (int v = 0; v < 3; v++)
        for (int k = 0; k < 3; k++) {
          b.lo[k] = std::min(b.lo[k], t[3 * v + k]);
          b.hi[k] = std::max(b.hi[k], t[3 * v + k]);
        }
      for (int k = 0; k < 3; k++) {
        clo[k] = std::min(clo[k], cen[3 * (size_t)order[i] + k]);
        chi[k] = std::max(chi[k], cen[3 * (size_t)order[i] + k]);
      }
    }
    b.left = b.right = -1;
    b.tri = -1;
    b.count = hi - lo;
    int me = (int)tmp.size();
    tmp.push_back(b);
    if (hi - lo == 1) {
      tmp[me].tri = order[lo];
      return me;
    }
    int axis = 0;
    float ext = chi[0] - clo[0];
    for (int k = 1; k < 3; k++)
      if (chi[k] - clo[k] > ext) { ext = chi[k] - clo[k]; axis = k; }
    int mid = lo + (hi - lo) / 2;
    std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi, [&](int a, int c) {
      float ca = cen[3 * (size_t)a + axis], cc = cen[3 * (size_t)c + axis];
      return ca < cc || (ca == cc && a < c);
    });
    int l = build(lo, mid);
    int r = build(mid, hi);
    tmp[me].left = l;
    tmp[me].right = r;
    return me;
  }
};
} // namespace

void build_bvh(const float *tri9, int ntri, std::vector<Node> &nodes, int *depth_out) {
  nodes.clear();
  if (ntri <= 0) {
    // empty mesh: a single far-away degenerate leaf that nothing can reach
    Node n;
    n.cx = n.cy = n.cz = 1.0e18f;
    n.hx = n.hy = n.hz = 0.f;
    n.child = -1;
    n.tri = -1;
    nodes.push_back(n);
    if (depth_out) *depth_out = 0;
    return;
  }
  Builder B;
  B.tri9 = tri9;
  B.cen.resize(3 * (size_t)ntri);
  B.order.resize(ntri);
  std::iota(B.order.begin(), B.order.end(), 0);
  for (int i = 0; i < ntri; i++)
    for (int k = 0; k < 3; k++)
      B.cen[3 * (size_t)i + k] = (tri9[9 * (size_t)i + k] + tri9[9 * (size_t)i + 3 + k] + tri9[9 * (size_t)i + 6 + k]) / 3.0f;
  B.tmp.reserve(2 * (size_t)ntri);
  B.build(0, ntri);

  // breadth-first renumbering, siblings adjacent
  nodes.resize(B.tmp.size());
  std::vector<int> newidx(B.tmp.size(), -1);
  std::queue<std::pair<int, int>> q;   // (tmp index, depth)
  int next = 1, depth = 0;
  newidx[0] = 0;
  q.push({0, 0});
  while (!q.empty()) {
    auto [ti, d] = q.front();
    q.pop();
    depth = std::max(depth, d);
    const Build &b = B.tmp[ti];
    Node n;
    float maxabs = 0.f;
    for (int k = 0; k < 3; k++) maxabs = std::max(maxabs, std::max(std::fabs(b.lo[k]), std::fabs(b.hi[k])));
    // padding: keeps fp32 culling conservative (the reference pads by Leaf::TOLERANCE = 0.01 mm,
    // collisionModel.cpp:43,122-128)
    const float pad = 1.0e-3f + 2.0e-6f * maxabs;
    n.cx = 0.5f * (b.lo[0] + b.hi[0]);
    n.cy = 0.5f * (b.lo[1] + b.hi[1]);
    n.cz = 0.5f * (b.lo[2] + b.hi[2]);
    n.hx = 0.5f * (b.hi[0] - b.lo[0]) + pad;
    n.hy = 0.5f * (b.hi[1] - b.lo[1]) + pad;
    n.hz = 0.5f * (b.hi[2] - b.lo[2]) + pad;
    if (b.left < 0) {
      n.child = -1;
      n.tri = b.tri;
    } else {
      newidx[b.left] = next;
      newidx[b.right] = next + 1;
      n.child = next;
      n.tri = b.count;
      next += 2;
      q.push({b.left, d + 1});
      q.push({b.right, d + 1});
    }
    nodes[newidx[ti]] = n;
  }
  if (depth_out) *depth_out = depth;
}

} // namespace b2c
