// Host-side pieces of the CollisionInterface surface that are not throughput-critical:
//   * the reference-style oriented-box hierarchy served by getBoundingVolumes
//   * contact-set post-processing (duplicate removal, perimeter compaction) shared by every backend
//     in the reference (src/Collision/collisionInterface.cpp:74-287)
#pragma once
#include <vector>
#include "../../include/b2c.h"

namespace b2c {

constexpr int FK_MAX_DOF_HOST = 32;
constexpr int FK_MAX_ROBOTS = 4;     // robots in one kinematic tree (fk_kernel keeps their parents' last-link poses in registers)

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
struct OrderNode;
// the tree as the visit-order rule reads it (obb_order.h) and the triangle -> depth-first leaf position table: for the device
const OrderNode *obb_order_nodes(const ObbTree *t, int *count);
const int *obb_tri_positions(const ObbTree *t, int *count);
// Node::getBVRecurse / Branch::getBVRecurse (collisionModel.cpp:437-456): boxes at `depth`, plus
// leaves that end above it, in depth-first order.
void obb_tree_level(const ObbTree *t, int depth, std::vector<ObbBox> &out);

// Reference visit order of candidate triangle pairs (tri_a[i] of tree A, tri_b[i] of tree B): the order in which
// ContactCallback's recursion reaches their leaf pairs (collisionAlgorithms.cpp:42-130).  R21 (row-major), t21:
// transform from model 2's frame to model 1's (mTran2To1).  `order` = indices into the candidate arrays; candidates
// of one leaf pair keep their input order.
void obb_visit_order(const ObbTree *A, const ObbTree *B, const double R21[9], const double t21[3], const int *tri_a,
                     const int *tri_b, int n, std::vector<int> &order);

// CollisionInterface::removeContactDuplicates (collisionInterface.cpp:254-287)
void remove_contact_duplicates(std::vector<b2c_contact> &contacts, double duplicate_threshold);
// CollisionInterface::compactContactSet / replaceContactSetWithPerimeter (collisionInterface.cpp:74-252)
void compact_contact_set(std::vector<b2c_contact> &contacts);

// Contact::Contact frame (contact.cpp:67-99): frame7 = [n][q wxyz, t]; tangents6 (optional) = [n][tangentX, tangentY]
int contact_frames(const b2c_contact *c, int n, int side, double *frame7, double *tangents6);
// Contact::setUpFrictionEllipsoid (contact.cpp:180-216): edges = 6 doubles per sample
void friction_ellipsoid_edges(int nlat, const int *ndirs, const double *phi, const double *eccen, std::vector<double> &edges);
// PointContact friction cone -> boundary wrenches (contact.cpp:124-165): wrench48 = [n][8][force(3), torque(3)]
int point_contact_wrenches(const b2c_contact *c, int n, int side, double cof, const double cog[3], double max_radius, double *wrench48);
// checkContactNormals (contactSetting.cpp:54-81): flips ill-oriented normals in place, returns how many contacts changed
int check_contact_normals(b2c_contact *c, int n, const double q1[4], const double t1[3], const double q2[4], const double t2[3]);

// soft-finger contacts (soft_contact_host.cpp; contactSetting.cpp:83-191, softContact.cpp, FitParabola.h)
double soft_neighborhood_radius(double youngs1, double youngs2);
int soft_contact_merge(std::vector<b2c_contact> &contacts, std::vector<std::vector<double>> &n1,
                       std::vector<std::vector<double>> &n2, const double q1[4], const double t1[3], const double q2[4],
                       const double t2[3]);
int soft_contact_fit(const b2c_contact *c, int n, int side, const double *nghbd_xyz, const int *offsets, b2c_soft_fit *fit);
int soft_contact_wrenches(const b2c_contact *c, int n, int side, const b2c_soft_fit *fit_this, const b2c_soft_fit *fit_mate,
                          const double q_this[4], const double q_mate[4], double youngs, double cof, const double cog[3],
                          double max_radius, double *wrench120, double *ellipse6);

} // namespace b2c
