// TEST INFRASTRUCTURE ONLY -- the oracle.  Never imported, linked or executed by the product path.
//
// Reference-pinned forward kinematics and state -> pose mapping.
//
// The reference's own arithmetic is compiled here, not restated:
//   * class DHTransform and transf come from the reference headers (include/graspit/joint.h, transform.h) verbatim;
//   * DHTransform's out-of-line members, the joints' getTran, KinematicChain::fwdKinematics and the four
//     PositionState*::getCoreTran bodies are the reference's OWN SOURCE LINES, cut out of the files where they lie under
//     /root/reference by oracle/Makefile (sed line ranges, checked by grep) into oracle/_ref/obj/extract/*.inc at build
//     time and #included below.  Those translation units as a whole need Qt / Coin / Robot / World and cannot be
//     compiled here; the extracted functions only need transf.  Nothing is copied into the repo.
//       joint.cpp:61-87             DHTransform::DHTransform, DHTransform::computeTran
//       joint.h:362-364, 396-398    PrismaticJoint::getTran, RevoluteJoint::getTran
//       kinematicChain.cpp:543-559  KinematicChain::fwdKinematics
//       searchStateImpl.cpp:125-135, 156-168, 218-252, 266-274   PositionState{Complete,AA,Ellipsoid,Approach}::getCoreTran
// What IS written here: the stand-in classes those member functions are attached to (data members with the reference's
// names, nothing else), HandObjectState::execute's one-line composition `mRefTran % mPosition->getCoreTran()`
// (searchState.cpp:520) and the attached-robot rule `lastLinkTran % childOffsetTran` (kinematicChain.cpp:524-526).
#include <vector>
#include <map>
#include <string>
#include <cstring>
#include <cmath>
#include <cassert>

#include "graspit/matvec3D.h"
#include "graspit/joint.h"             // reference header: DHTransform (+ Joint declarations, unused)

#include "extract/dh_members.inc"      // DHTransform::DHTransform(double,double,double,double), DHTransform::computeTran()

namespace {

// ---- joints: the two getTran bodies are the reference's lines -------------------------------------------------------
struct JointStub {
  DHTransform *DH;
  double c;              // Joint::c, the constant offset of the DOF expression (joint.cpp:254-261, 355-363)
  JointStub() : DH(NULL), c(0.0) {}
  virtual ~JointStub() { delete DH; }
  virtual transf getTran(double jointVal) const = 0;
  double getVal() const { return 0.0; }        // only reached with jointVals == NULL, which this driver never passes
};
struct RevoluteStub : public JointStub {
#include "extract/revolute_getTran.inc"
};
struct PrismaticStub : public JointStub {
#include "extract/prismatic_getTran.inc"
};

struct OwnerStub {
  transf tran;
  const transf &getTran() const { return tran; }
};

} // namespace

// the reference's member function below is attached to this stand-in (same member names as kinematicChain.h)
class KinematicChain {
  public:
    OwnerStub *owner;
    transf tran;
    int numJoints, numLinks, firstJointNum;
    std::vector<JointStub *> jointVec;
    int *lastJoint;
    void fwdKinematics(const double *jointVals, std::vector<transf> &newLinkTranVec) const;
};
#include "extract/fwdKinematics.inc"

namespace {
struct HandStub {
  transf approach;
  const transf &getApproachTran() const { return approach; }
};
struct VarState {
  HandStub *mHand;
  std::map<std::string, double> vars, params;
  double readVariable(const char *n) const { return vars.at(n); }
  double getParameter(const char *n) const { return params.at(n); }
};
} // namespace
struct PositionStateComplete : public VarState { transf getCoreTran() const; };
struct PositionStateAA : public VarState { transf getCoreTran() const; };
struct PositionStateEllipsoid : public VarState { transf getCoreTran() const; };
struct PositionStateApproach : public VarState { transf getCoreTran() const; };
#include "extract/getCoreTran_complete.inc"
#include "extract/getCoreTran_aa.inc"
#include "extract/getCoreTran_ellipsoid.inc"
#include "extract/getCoreTran_approach.inc"

namespace {
transf mk7(const double *q7) { return transf(Quaternion(q7[0], q7[1], q7[2], q7[3]), vec3(q7[4], q7[5], q7[6])); }
void put7(const transf &t, double *o) {
  o[0] = t.rotation().w(); o[1] = t.rotation().x(); o[2] = t.rotation().y(); o[3] = t.rotation().z();
  o[4] = t.translation().x(); o[5] = t.translation().y(); o[6] = t.translation().z();
}
} // namespace

extern "C" {

// One kinematic chain through the reference's fwdKinematics.
//   owner7, chain7: owner (robot base) transform and the chain's own transform, (q wxyz, t)
//   joints: njoints x 6 = revolute flag, DH theta, d, a, alpha (radians / mm, values at joint value 0 WITHOUT c), c
//           -- RevoluteJoint builds DHTransform(theta + c, d, a, alpha), PrismaticJoint DHTransform(theta, d + c, a, alpha)
//   jointVals: njoints joint values; last_joint: nlinks chain-local joint indices; out7: nlinks x 7
void ref_fk_chain(const double *owner7, const double *chain7, int njoints, const double *joints, const double *jointVals,
                  int nlinks, const int *last_joint, double *out7) {
  OwnerStub owner;
  owner.tran = mk7(owner7);
  KinematicChain kc;
  kc.owner = &owner; kc.tran = mk7(chain7); kc.numJoints = njoints; kc.numLinks = nlinks; kc.firstJointNum = 0;
  std::vector<int> lj(last_joint, last_joint + nlinks);
  kc.lastJoint = lj.data();
  for (int j = 0; j < njoints; j++) {
    const double *p = joints + 6 * j;
    JointStub *J;
    if (p[0] != 0.0) { J = new RevoluteStub(); J->DH = new DHTransform(p[1] + p[5], p[2], p[3], p[4]); }
    else { J = new PrismaticStub(); J->DH = new DHTransform(p[1], p[2] + p[5], p[3], p[4]); }
    J->c = p[5];
    kc.jointVec.push_back(J);
  }
  std::vector<transf> links(nlinks, transf::IDENTITY);
  kc.fwdKinematics(jointVals, links);
  for (int l = 0; l < nlinks; l++) put7(links[l], out7 + 7 * l);
  for (JointStub *J : kc.jointVec) delete J;
}

// child robot base = last link of the parent chain % childOffsetTran (KinematicChain::updateLinkPoses, kinematicChain.cpp:524-526)
void ref_fk_child_base(const double *last_link7, const double *offset7, double *out7) { put7(mk7(last_link7) % mk7(offset7), out7); }

// HandObjectState::execute (searchState.cpp:515-527): hand transform = mRefTran % mPosition->getCoreTran().
//   kind 0 SPACE_COMPLETE (Tx Ty Tz Qw Qx Qy Qz), 1 SPACE_AXIS_ANGLE (Tx Ty Tz theta phi alpha),
//   2 SPACE_ELLIPSOID (beta gamma tau dist; params a b c), 3 SPACE_APPROACH (dist, wrist 1, wrist 2)
int ref_state_to_base(int kind, const double *vars, const double *params, const double *ref7, const double *approach7, double *out7) {
  HandStub hand;
  hand.approach = mk7(approach7);
  transf core;
  if (kind == 0) {
    PositionStateComplete s; s.mHand = &hand;
    const char *n[7] = {"Tx", "Ty", "Tz", "Qw", "Qx", "Qy", "Qz"};
    for (int i = 0; i < 7; i++) s.vars[n[i]] = vars[i];
    core = s.getCoreTran();
  } else if (kind == 1) {
    PositionStateAA s; s.mHand = &hand;
    const char *n[6] = {"Tx", "Ty", "Tz", "theta", "phi", "alpha"};
    for (int i = 0; i < 6; i++) s.vars[n[i]] = vars[i];
    core = s.getCoreTran();
  } else if (kind == 2) {
    PositionStateEllipsoid s; s.mHand = &hand;
    const char *n[4] = {"beta", "gamma", "tau", "dist"};
    for (int i = 0; i < 4; i++) s.vars[n[i]] = vars[i];
    s.params["a"] = params[0]; s.params["b"] = params[1]; s.params["c"] = params[2];
    core = s.getCoreTran();
  } else if (kind == 3) {
    PositionStateApproach s; s.mHand = &hand;
    const char *n[3] = {"dist", "wrist 1", "wrist 2"};
    for (int i = 0; i < 3; i++) s.vars[n[i]] = vars[i];
    core = s.getCoreTran();
  } else {
    return -1;
  }
  put7(mk7(ref7) % core, out7);
  return 0;
}

} // extern "C"
