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
#include <cstdlib>
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
  bool use_sah = true;
  int nbins = 16, balance = 8;

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
    // split: binned surface-area heuristic over the three axes (16 bins on the centroid bounds); object median along
    // the longest centroid axis when no SAH split separates the set.  (B2C_BVH_MEDIAN=1 forces the median split.)
    int axis = 0;
    float ext = chi[0] - clo[0];
    for (int k = 1; k < 3; k++)
      if (chi[k] - clo[k] > ext) { ext = chi[k] - clo[k]; axis = k; }
    int mid = lo + (hi - lo) / 2;
    bool done = false;
    if (use_sah && hi - lo > 2) {
      const int NB = 64; const int nb = nbins;
      float best_cost = 3.0e38f; int best_axis = -1, best_bin = -1;
      for (int ax = 0; ax < 3; ax++) {
        float w = chi[ax] - clo[ax];
        if (!(w > 1.0e-9f)) continue;
        float blo[NB][3], bhi[NB][3]; int cnt[NB];
        for (int b = 0; b < nb; b++) { cnt[b] = 0; for (int k = 0; k < 3; k++) { blo[b][k] = 1e30f; bhi[b][k] = -1e30f; } }
        for (int i = lo; i < hi; i++) {
          int t = order[i];
          int b = (int)((cen[3 * (size_t)t + ax] - clo[ax]) / w * nb);
          b = b < 0 ? 0 : (b >= nb ? nb - 1 : b);
          cnt[b]++;
          const float *tv = tri9 + 9 * (size_t)t;
          for (int v = 0; v < 3; v++)
            for (int k = 0; k < 3; k++) { blo[b][k] = std::min(blo[b][k], tv[3 * v + k]); bhi[b][k] = std::max(bhi[b][k], tv[3 * v + k]); }
        }
        // sweep: area of the left / right unions
        float la[NB], ra[NB]; int lc[NB], rc[NB];
        float l0[3] = {1e30f, 1e30f, 1e30f}, l1[3] = {-1e30f, -1e30f, -1e30f}; int c = 0;
        auto area = [](const float *a, const float *b) { float x = b[0] - a[0], y = b[1] - a[1], z = b[2] - a[2]; return x * y + y * z + z * x; };
        for (int b = 0; b < nb; b++) {
          if (cnt[b]) for (int k = 0; k < 3; k++) { l0[k] = std::min(l0[k], blo[b][k]); l1[k] = std::max(l1[k], bhi[b][k]); }
          c += cnt[b]; lc[b] = c; la[b] = c ? area(l0, l1) : 0.f;
        }
        float r0[3] = {1e30f, 1e30f, 1e30f}, r1[3] = {-1e30f, -1e30f, -1e30f}; c = 0;
        for (int b = nb - 1; b >= 0; b--) {
          if (cnt[b]) for (int k = 0; k < 3; k++) { r0[k] = std::min(r0[k], blo[b][k]); r1[k] = std::max(r1[k], bhi[b][k]); }
          c += cnt[b]; rc[b] = c; ra[b] = c ? area(r0, r1) : 0.f;
        }
        for (int b = 0; b + 1 < nb; b++) {
          if (lc[b] == 0 || rc[b + 1] == 0) continue;
          float cost = la[b] * lc[b] + ra[b + 1] * rc[b + 1];
          if (cost < best_cost) { best_cost = cost; best_axis = ax; best_bin = b; }
        }
      }
      if (best_axis >= 0) {
        float w = chi[best_axis] - clo[best_axis];
        auto part = std::stable_partition(order.begin() + lo, order.begin() + hi, [&](int t) {
          int b = (int)((cen[3 * (size_t)t + best_axis] - clo[best_axis]) / w * nbins);
          b = b < 0 ? 0 : (b >= nbins ? nbins - 1 : b);
          return b <= best_bin;
        });
        int m = (int)(part - order.begin());
        // keep the tree from degenerating: a split more lopsided than 1:7 falls back to the median
        if (m > lo && m < hi && std::min(m - lo, hi - m) * balance >= (hi - lo)) { mid = m; done = true; }
      }
    }
    if (!done) {
      mid = lo + (hi - lo) / 2;
      std::nth_element(order.begin() + lo, order.begin() + mid, order.begin() + hi, [&](int a, int c) {
        float ca = cen[3 * (size_t)a + axis], cc = cen[3 * (size_t)c + axis];
        return ca < cc || (ca == cc && a < c);
      });
    }
    int l = build(lo, mid);
    int r = build(mid, hi);
    tmp[me].left = l;
    tmp[me].right = r;
    return me;
  }
};
} // namespace

void build_bvh(const float *tri9, int ntri, std::vector<NodeQ> &nodes, BvhQuant *quant, int *depth_out, Node *root_out) {
  nodes.clear();
  BvhQuant g;
  if (ntri <= 0) {
    // empty mesh: one far-away leaf holding triangle 0 (the caller stores a degenerate triangle there)
    for (int k = 0; k < 3; k++) { g.q0[k] = 1.0e18f; g.qs[k] = 0.f; }
    NodeQ n;
    for (int k = 0; k < 3; k++) { n.lo[k] = 0; n.hi[k] = 0; }
    n.link = ~0;
    nodes.push_back(n); nodes.push_back(n);
    if (quant) *quant = g;
    if (depth_out) *depth_out = 0;
    if (root_out) *root_out = decode_node(n, g);
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
  if (const char *e = getenv("B2C_BVH_MEDIAN")) B.use_sah = atoi(e) == 0;
  if (const char *e = getenv("B2C_BVH_BINS")) B.nbins = std::max(2, std::min(64, atoi(e)));
  if (const char *e = getenv("B2C_BVH_BALANCE")) B.balance = std::max(2, atoi(e));
  B.build(0, ntri);
  if (B.use_sah) {
    // per-lane traversal stacks (point_kernel) hold one pending sibling per level: keep the depth well inside them
    std::vector<int> dep(B.tmp.size(), 0);
    int maxd = 0;
    for (size_t i = 0; i < B.tmp.size(); i++) {      // parents precede children in tmp
      if (B.tmp[i].left >= 0) { dep[B.tmp[i].left] = dep[i] + 1; dep[B.tmp[i].right] = dep[i] + 1; maxd = std::max(maxd, dep[i] + 1); }
    }
    if (maxd > 48) {
      B.use_sah = false;
      B.tmp.clear();
      std::iota(B.order.begin(), B.order.end(), 0);
      B.build(0, ntri);
    }
  }

  // padding: keeps fp32 culling conservative (the reference pads by Leaf::TOLERANCE = 0.01 mm,
  // collisionModel.cpp:43,122-128)
  auto pad_of = [](const Build &b) {
    float maxabs = 0.f;
    for (int k = 0; k < 3; k++) maxabs = std::max(maxabs, std::max(std::fabs(b.lo[k]), std::fabs(b.hi[k])));
    return 1.0e-3f + 2.0e-6f * maxabs;
  };
  // quantisation grid: the padded root box split into 65535 steps per axis (a little slack so that every padded
  // child box, which the root contains up to its own padding, lands inside the grid)
  {
    const Build &r = B.tmp[0];
    float pad = 4.0f * pad_of(r);
    for (int k = 0; k < 3; k++) {
      g.q0[k] = r.lo[k] - pad;
      float span = (r.hi[k] + pad) - g.q0[k];
      g.qs[k] = std::max(span / 65535.0f, 1.0e-12f);
    }
  }
  auto quant_lo = [&](float v, int k) {
    double q = std::floor(((double)v - (double)g.q0[k]) / (double)g.qs[k]);
    // the device decodes q0 + q * qs in fp32: step down until the decoded value is not above v
    long long qi = (long long)std::max(0.0, std::min(65535.0, q));
    while (qi > 0 && std::fmaf((float)qi, g.qs[k], g.q0[k]) > v) qi--;
    return (unsigned short)qi;
  };
  auto quant_hi = [&](float v, int k) {
    double q = std::ceil(((double)v - (double)g.q0[k]) / (double)g.qs[k]);
    long long qi = (long long)std::max(0.0, std::min(65535.0, q));
    while (qi < 65535 && std::fmaf((float)qi, g.qs[k], g.q0[k]) < v) qi++;
    return (unsigned short)qi;
  };

  // breadth-first renumbering: [0] root, [1] unused, sibling pairs from index 2 on
  nodes.resize(B.tmp.size() + 1);
  std::vector<int> newidx(B.tmp.size(), -1);
  std::queue<std::pair<int, int>> q;   // (tmp index, depth)
  int next = 2, depth = 0;
  newidx[0] = 0;
  q.push({0, 0});
  while (!q.empty()) {
    auto [ti, d] = q.front();
    q.pop();
    depth = std::max(depth, d);
    const Build &b = B.tmp[ti];
    const float pad = pad_of(b);
    NodeQ n;
    for (int k = 0; k < 3; k++) { n.lo[k] = quant_lo(b.lo[k] - pad, k); n.hi[k] = quant_hi(b.hi[k] + pad, k); }
    if (b.left < 0) {
      n.link = ~b.tri;
    } else {
      newidx[b.left] = next;
      newidx[b.right] = next + 1;
      n.link = next;
      next += 2;
      q.push({b.left, d + 1});
      q.push({b.right, d + 1});
    }
    nodes[newidx[ti]] = n;
  }
  nodes[1] = nodes[0];
  if (quant) *quant = g;
  if (depth_out) *depth_out = depth;
  if (root_out) *root_out = decode_node(nodes[0], g);
}

} // namespace b2c
