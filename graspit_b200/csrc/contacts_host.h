// Host-side pieces of the CollisionInterface surface that are not throughput-critical:
//   * the reference-style oriented-box hierarchy served by getBoundingVolumes
//   * contact-set post-processing (duplicate removal, perimeter compaction) shared by every backend
//     in the reference (src/Collision/collisionInterface.cpp:74-287)
#pragma once
#include <vector>
#include "../../include/b2c.h"

namespace b2c {

constexpr int FK_MAX_DOF_HOST = 32;

struct ObbBox {
  double t[3];
  double q[4];      // wxyz
  double half[3];
};

struct ObbTree;

// Oriented-box tree with the reference's construction rules (CollisionModel::build,
// src/Collision/Graspit/collisionModel.cpp:97-344,517-539,585-679): area-weighted covariance,
// Jacobi eigenvectors, box padded by 0.01 mm, x axis = longest extent, split at the mean vertex
// along x (median and alternating fallbacks), one triangle per leaf.
ObbTree *obb_tree_build(const float *tri9, int ntri);
void obb_tree_free(ObbTree *t);
// Node::getBVRecurse / Branch::getBVRecurse (collisionModel.cpp:437-456): boxes at `depth`, plus
// leaves that end above it, in depth-first order.
void obb_tree_level(const ObbTree *t, int depth, std::vector<ObbBox> &out);

// CollisionInterface::removeContactDuplicates (collisionInterface.cpp:254-287)
void remove_contact_duplicates(std::vector<b2c_contact> &contacts, double duplicate_threshold);
// CollisionInterface::compactContactSet / replaceContactSetWithPerimeter (collisionInterface.cpp:74-252)
void compact_contact_set(std::vector<b2c_contact> &contacts);

} // namespace b2c
