// Flat BVH node + host builder interface (see bvh.cpp).
#pragma once
#include <vector>

namespace b2c {

// 32 B, two 16-B vector loads.  Axis-aligned box in the body frame (centre c, half extents h).
// Nodes are stored breadth-first with the two children of a node adjacent (child, child+1), so the
// first K nodes are the top levels of the tree and can be staged into shared memory with one bulk
// copy.  child < 0 marks a leaf holding triangle `tri`; for inner nodes `tri` is the triangle count.
struct Node {
  float cx, cy, cz, hx;
  float hy, hz;
  int child;
  int tri;
};
static_assert(sizeof(Node) == 32, "Node must be 32 bytes");

// tri9: ntri x 9 floats.  Output: nodes in device layout; depth of the tree.
void build_bvh(const float *tri9, int ntri, std::vector<Node> &nodes, int *depth_out);

} // namespace b2c
