// Flat BVH node formats + host builder interface (see bvh.cpp).
#pragma once
#include <cmath>
#include <vector>

namespace b2c {

// Decoded node (registers / host): axis-aligned box in the body frame (centre c, half extents h).
// child < 0 marks a leaf holding triangle `tri`.
struct Node {
  float cx, cy, cz, hx;
  float hy, hz;
  int child;
  int tri;
};

// Stored node: 16 B.  Box corners quantised to 16 bits on the mesh's own grid (origin q0, step qs per axis; lo rounded
// down, hi rounded up, so the stored box contains the padded fp32 box), link >= 0: index of the first child (children
// are adjacent: link, link + 1), link < 0: leaf holding triangle ~link.
// Layout of the array: [0] root, [1] unused, then sibling pairs at even indices -- a pair is 32 B, 32-B aligned, so
// one 256-bit load (LDG.E.256) fetches both children and never straddles a sector; the array is breadth-first, so a
// prefix is the top of the tree and can be staged into shared memory with one bulk copy.
// Traversal kernels are bound by L1 wavefronts under fully divergent addresses (one cache line per lane per load
// instruction): halving the bytes per child pair halves the load instructions per box test.
struct NodeQ {
  unsigned short lo[3], hi[3];
  int link;
};
static_assert(sizeof(NodeQ) == 16, "NodeQ must be 16 bytes");

struct BvhQuant {
  float q0[3];     // grid origin
  float qs[3];     // grid step
};

// tri9: ntri x 9 floats.  Output: nodes in device layout, their quantisation grid, depth of the tree, and the root box
// (decoded) for the mesh extent.
void build_bvh(const float *tri9, int ntri, std::vector<NodeQ> &nodes, BvhQuant *quant, int *depth_out, Node *root_out);

// host decode (same arithmetic as the device)
inline Node decode_node(const NodeQ &q, const BvhQuant &g) {
  Node n;
  float lo[3], hi[3];
  for (int k = 0; k < 3; k++) { lo[k] = std::fmaf((float)q.lo[k], g.qs[k], g.q0[k]); hi[k] = std::fmaf((float)q.hi[k], g.qs[k], g.q0[k]); }
  n.cx = 0.5f * (lo[0] + hi[0]); n.cy = 0.5f * (lo[1] + hi[1]); n.cz = 0.5f * (lo[2] + hi[2]);
  n.hx = 0.5f * (hi[0] - lo[0]); n.hy = 0.5f * (hi[1] - lo[1]); n.hz = 0.5f * (hi[2] - lo[2]);
  n.child = q.link >= 0 ? q.link : -1;
  n.tri = q.link >= 0 ? -1 : ~q.link;
  return n;
}

} // namespace b2c
