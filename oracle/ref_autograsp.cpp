// TEST INFRASTRUCTURE ONLY -- the oracle.  Never imported, linked or executed by the product path.
//
// Reference-pinned autograsp: Hand::autoGrasp -> Robot::moveDOFToContacts -> jumpDOFToContact ->
// getJointValuesFromDOF / interpolateJoints / stopJointsFromLink, the DOF classes' accumulateMove / updateVal /
// getJointValues, and the pair filter of World::findContacts, on the reference's own collision backend.
//
// The DOF classes are the reference's own header (include/graspit/dof.h).  The member functions are the reference's OWN
// SOURCE LINES, cut out of robot.cpp / robot.h / dof.cpp / world.cpp by oracle/Makefile (sed line ranges, checked by grep)
// into oracle/_ref/obj/extract/*.inc at build time and #included below; those translation units as a whole need Qt,
// Coin, the dynamics engine and the XML loaders.  Nothing is copied into the repo.
//   robot.cpp:1393-1399      Robot::setJointValuesAndUpdate
//   robot.cpp:1418-1467      Robot::interpolateJoints
//   robot.cpp:1475-1489      Robot::stopJointsFromLink
//   robot.cpp:1507-1564      Robot::getJointValuesFromDOF
//   robot.cpp:1593-1677      Robot::jumpDOFToContact
//   robot.cpp:1708-1810      Robot::moveDOFToContacts
//   robot.cpp:2427-2462      Hand::autoGrasp
//   robot.h:563-578, 610-612 Robot::updateDofVals, getJointValues, getDOFVals
//   dof.cpp:40-45, 82-119    DOF::DOF, initDOF, updateMinMax
//   dof.cpp:483-534          RigidDOF::getStaticRatio, getJointValues, accumulateMove, updateFromJointValues
//   dof.cpp:536-678          BreakAwayDOF: destructor, initDOF, getJointValues, accumulateMove, updateVal,
//                            updateFromJointValues, reset
//   dof.cpp:766-856          CompliantDOF::initDOF, getStaticRatio, getJointValues, updateFromJointValues, accumulateMove
//   contact.cpp:58, robot.cpp:68, worldElement.cpp:47   Contact::THRESHOLD, Robot::AUTO_GRASP_TIME_STEP, WorldElement::ONE_STEP
//   world.cpp:1515-1526      World::getCollisionReport
//   world.cpp:1681-1707      World::findContacts, up to the end of the pair filter (what follows registers Contact objects)
// What IS written here: the stand-in Joint / Link / KinematicChain / World / Robot classes those lines are attached to
// (data members with the reference's names), link poses through the reference's fwdKinematics lines (ref_fk_chain,
// oracle/ref_fk.cpp), Link::contactsPreventMotion == false (no contact is registered on a link that has just been moved:
// body.cpp:928-931), and the C entry point.
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <string>
#include <vector>

#include <QObject>
#include <QString>
#include <QTextStream>
#include "graspit/matvec3D.h"
#include "graspit/mytools.h"
#include "graspit/body.h"                                   // the collision oracle's stand-in Body
#include "graspit/Collision/collisionStructures.h"
#include "graspit/Collision/collisionInterface.h"
#include "graspit/debug.h"
#include <sstream>

// the reference reports an interpolation failure and the 500-cycle failsafe only through DBGA: keep the messages
static std::vector<std::string> g_dbga;
#undef DBGA
#define DBGA(STMT) { std::ostringstream dbga_s_; dbga_s_ << STMT; g_dbga.push_back(dbga_s_.str()); }

#define PROF_TIMER_FUNC(STR) ;

extern "C" void ref_fk_chain(const double *owner7, const double *chain7, int njoints, const double *joints, const double *jointVals,
                             int nlinks, const int *last_joint, double *out7);

// ---- stand-ins ------------------------------------------------------------------------------------------------------------
class WorldElement {
  public:
    static const double ONE_STEP;
    virtual ~WorldElement() {}
};
#include "extract/worldElement_ONE_STEP.inc"                // worldElement.cpp:47

class Joint {
  public:
    int num; double val, mn, mx, ratio, stiffness;
    double dh[6];                                           // revolute flag, theta, d, a, alpha, c  (ref_fk_chain's row)
    int getNum() const { return num; }
    double getVal() const { return val; }
    void setVal(double v) { val = v; }
    double getMin() const { return mn; }
    double getMax() const { return mx; }
    double getCouplingRatio() const { return ratio; }
    double getSpringStiffness() const { return stiffness; }
};

class Robot;
class Link : public Body {
  public:
    Robot *owner; int chainNum, linkNum;
    Link(Robot *o, int c, int l) : Body("link"), owner(o), chainNum(c), linkNum(l) {}
    virtual WorldElement *getOwner() { return (WorldElement *)owner; }
    int getChainNum() const { return chainNum; }
    int getLinkNum() const { return linkNum; }
    bool contactsPreventMotion(const transf &) const { return false; }
};

class World;
class KinematicChain {
  public:
    Robot *owner; World *world;
    transf tran;
    int firstJointNum;
    std::vector<Joint *> jointVec;
    std::vector<Link *> linkVec;
    std::vector<int> lastJoint;
    int getNumLinks() const { return (int)linkVec.size(); }
    int getNumJoints() const { return (int)jointVec.size(); }
    Link *getLink(int l) const { return linkVec[l]; }
    Joint *getJoint(int j) const { return jointVec[j]; }
    int getLastJoint(int l) const { return lastJoint[l]; }
    void getJointValues(double *jointVals) const { for (size_t j = 0; j < jointVec.size(); j++) jointVals[firstJointNum + j] = jointVec[j]->getVal(); }   // kinematicChain.cpp:487-492
    void setJointValues(const double *jointVals) { for (size_t j = 0; j < jointVec.size(); j++) jointVec[j]->setVal(jointVals[firstJointNum + j]); }      // :500-505
    void updateLinkPoses();
    void infinitesimalMotion(const double *, std::vector<transf> &) const {}        // only consumed by contactsPreventMotion
};

class World {
  public:
    CollisionInterface *mCollisionInterface;
    bool allCollisionsOFF;
    World() : mCollisionInterface(NULL), allCollisionsOFF(false) {}
    bool dynamicsAreOn() const { return false; }
    int getCollisionReport(CollisionReport *colReport, const std::vector<Body *> *interestList = NULL);
    double getDist(Body *b1, Body *b2) { position p1, p2; return mCollisionInterface->bodyToBodyDistance(b1, b2, p1, p2); }   // world.cpp:1596-1597
    void findContacts(CollisionReport &colReport);
};

struct ViewerStub { void render() {} };
struct IVmgrStub { ViewerStub v; ViewerStub *getViewer() { return &v; } };
struct GraspitCoreStub { World *getWorld() { return NULL; } IVmgrStub *getIVmgr() { return NULL; } };
static GraspitCoreStub *graspitCore = NULL;

namespace ContactNs { }
struct Contact { static const double THRESHOLD; };
#include "extract/contact_THRESHOLD.inc"                    // contact.cpp:58

#include "graspit/dof.h"

class Robot : public WorldElement {
  public:
    int numDOF, numChains, numJoints;
    std::vector<DOF *> dofVec;
    std::vector<KinematicChain *> chainVec;
    World *myWorld;
    transf baseTran;
    int iterations;                                         // moveDOFToContacts cycles, counted by moveDOFStepTaken's caller below
    static const double AUTO_GRASP_TIME_STEP;
    Robot() : numDOF(0), numChains(0), numJoints(0), myWorld(NULL), iterations(0) {}
    int getNumJoints() const { return numJoints; }
    int getNumDOF() const { return numDOF; }
    int getNumChains() const { return numChains; }
    KinematicChain *getChain(int c) const { return chainVec[c]; }
    QString getName() const { return QString("robot"); }
    const transf &getTran() const { return baseTran; }
    void moveDOFStepTaken(int, bool &) { iterations++; }
    inline void updateDofVals(double *dofVals);
    inline void getDOFVals(double *dofVals) const;
    inline void getJointValues(double *jointVals) const;
    void setJointValuesAndUpdate(const double *jointVals);
    int interpolateJoints(double *initialVals, double *finalVals, CollisionReport *colReport, double *interpolationTime);
    void stopJointsFromLink(Link *link, double *desiredJointVals, int *stoppedJoints);
    virtual bool getJointValuesFromDOF(const double *desiredDofVals, double *actualDofVals, double *jointVals, int *stoppedJoints);
    bool jumpDOFToContact(double *desiredVals, int *stoppedJoints, int *numCols = NULL);
    bool moveDOFToContacts(double *desiredVals, double *desiredSteps, bool stopAtContact, bool renderIt);
    void setDesiredDOFVals(double *) {}
};
#include "extract/robot_AUTO_GRASP_TIME_STEP.inc"           // robot.cpp:68
class Hand : public Robot {
  public:
    bool autoGrasp(bool renderIt, double speedFactor = 1.0, bool stopAtContact = false);
};

void KinematicChain::updateLinkPoses() {
  // kinematicChain.cpp:507-530: fwdKinematics(NULL, ...) on the current joint values, then every link is given its pose
  double owner7[7] = {owner->getTran().rotation().w(), owner->getTran().rotation().x(), owner->getTran().rotation().y(), owner->getTran().rotation().z(),
                      owner->getTran().translation().x(), owner->getTran().translation().y(), owner->getTran().translation().z()};
  double chain7[7] = {tran.rotation().w(), tran.rotation().x(), tran.rotation().y(), tran.rotation().z(),
                      tran.translation().x(), tran.translation().y(), tran.translation().z()};
  std::vector<double> j6(6 * jointVec.size()), jv(jointVec.size()), out(7 * linkVec.size());
  for (size_t j = 0; j < jointVec.size(); j++) { memcpy(&j6[6 * j], jointVec[j]->dh, 6 * sizeof(double)); jv[j] = jointVec[j]->getVal(); }
  ref_fk_chain(owner7, chain7, (int)jointVec.size(), j6.data(), jv.data(), (int)linkVec.size(), lastJoint.data(), out.data());
  for (size_t l = 0; l < linkVec.size(); l++) {
    const double *o = &out[7 * l];
    const transf t(Quaternion(o[0], o[1], o[2], o[3]), vec3(o[4], o[5], o[6]));
    linkVec[l]->setTranRaw(t);
    world->mCollisionInterface->setBodyTransform(linkVec[l], t);
  }
}

// ---- the reference's lines ------------------------------------------------------------------------------------------------
#include "extract/dof_ctor.inc"
#include "extract/dof_init.inc"
#include "extract/dof_rigid.inc"
#include "extract/dof_breakaway.inc"
#include "extract/dof_compliant.inc"
#include "extract/robot_inline_vals.inc"
#include "extract/robot_getDOFVals.inc"
#include "extract/robot_setJointValuesAndUpdate.inc"
#include "extract/robot_interpolateJoints.inc"
#include "extract/robot_stopJointsFromLink.inc"
#include "extract/robot_getJointValuesFromDOF.inc"
#include "extract/robot_jumpDOFToContact.inc"
#include "extract/robot_moveDOFToContacts.inc"
#include "extract/hand_autoGrasp.inc"
#include "extract/world_getCollisionReport.inc"
#include "extract/world_findContacts_filter.inc"
}   // closes World::findContacts after its pair filter (the cut ends there)

// ---- members of the DOF classes that are not on this path ---------------------------------------------------------------
DOF::DOF(const DOF *) { abort(); }
void DOF::updateSetPoint() {}
void DOF::setTrajectory(double *, int) {}
void DOF::addToTrajectory(double *, int) {}
void DOF::callController(double) {}
double DOF::PDPositionController(double) { return 0; }
bool DOF::dynamicsProgress() { return false; }
bool DOF::readParametersFromXml(const TiXmlElement *) { return true; }
bool DOF::writeToStream(QTextStream &) { return true; }
bool DOF::readFromStream(QTextStream &) { return true; }
void RigidDOF::setForce(double) {}
int RigidDOF::getNumLimitConstraints() { return 0; }
void RigidDOF::buildDynamicLimitConstraints(std::map<Body *, int> &, int, double *, double *, int &) {}
void RigidDOF::buildDynamicCouplingConstraints(std::map<Body *, int> &, int, double *, double *, int &) {}
bool BreakAwayDOF::writeToStream(QTextStream &) { return true; }
double BreakAwayDOF::getSaveVal() const { return q; }
bool BreakAwayDOF::readFromStream(QTextStream &) { return true; }
bool BreakAwayDOF::readParametersFromXml(const TiXmlElement *) { return true; }
bool BreakAwayDOF::computeStaticJointTorques(double *, double) { return true; }
bool CompliantDOF::computeStaticJointTorques(double *, double) { return true; }
int CompliantDOF::getNumLimitConstraints() { return 0; }
void CompliantDOF::buildDynamicLimitConstraints(std::map<Body *, int> &, int, double *, double *, int &) {}
void CompliantDOF::setForce(double) {}
double CompliantDOF::PDPositionController(double) { return 0; }
bool CompliantDOF::dynamicsProgress() { return false; }
double CompliantDOF::smoothProfileController(double) { return 0; }

// ---- the world the C entry points build -------------------------------------------------------------------------------------
#include "graspit/Collision/Graspit/graspitCollision.h"

namespace {
// stoppedJoints is a local of moveDOFToContacts: the DOFs see it on every accumulateMove, and the last calls (the cycle
// that finds no movement left) see its final state
std::vector<int> g_lastStopped;
int g_numJoints = 0;
template <class Base> struct Spy : public Base {
  virtual bool accumulateMove(double q1, double *jointVals, int *stoppedJoints) {
    if (stoppedJoints) g_lastStopped.assign(stoppedJoints, stoppedJoints + g_numJoints);
    return Base::accumulateMove(q1, jointVals, stoppedJoints);
  }
};
struct DofAccess : public DOF { static void setVel(DOF *d, double v) { static_cast<DofAccess *>(d)->defaultVelocity = v; } };

struct AgWorld {
  World world;
  GraspitCollision ci;
  std::vector<Body *> bodies;
  Hand hand;
  int base, mount;
  AgWorld() : base(-1), mount(-1) { world.mCollisionInterface = &ci; hand.myWorld = &world; }
  ~AgWorld() {
    for (DOF *d : hand.dofVec) delete d;
    for (KinematicChain *c : hand.chainVec) { for (Joint *j : c->jointVec) delete j; delete c; }
    for (Body *b : bodies) { ci.removeBody(b); delete b; }
  }
};
transf mk7(const double *q7) { return transf(Quaternion(q7[0], q7[1], q7[2], q7[3]), vec3(q7[4], q7[5], q7[6])); }
} // namespace

extern "C" {

void *ref_ag_create() { return new AgWorld(); }
void ref_ag_destroy(void *h) { delete static_cast<AgWorld *>(h); }

// chain >= 0: link `link` of chain `chain` of the hand; chain == -1: the hand's base (palm) or mount piece; -2: not part of the hand
int ref_ag_add_body(void *h, const double *tri9, int ntri, int chain, int link) {
  AgWorld *W = static_cast<AgWorld *>(h);
  Body *b = chain >= -1 ? new Link(&W->hand, chain, link) : new Body("body");
  std::vector<Triangle> tris;
  for (int i = 0; i < ntri; i++) {
    const double *t = tri9 + 9 * i;
    tris.push_back(Triangle(position(t[0], t[1], t[2]), position(t[3], t[4], t[5]), position(t[6], t[7], t[8])));
  }
  b->setTriangles(tris);
  W->ci.addBody(b);
  W->bodies.push_back(b);
  return (int)W->bodies.size() - 1;
}
void ref_ag_set_transform(void *h, int id, const double *t7) {
  AgWorld *W = static_cast<AgWorld *>(h);
  W->bodies[id]->setTranRaw(mk7(t7));
  W->ci.setBodyTransform(W->bodies[id], mk7(t7));
}
void ref_ag_activate_pair(void *h, int a, int b, int active) {
  AgWorld *W = static_cast<AgWorld *>(h);
  W->ci.activatePair(W->bodies[a], W->bodies[b], active != 0);
}

// joints: nj x 11 = dof, coupling ratio, revolute flag, DH theta, d, a, alpha (at joint value 0, WITHOUT the offset c), c,
//                   min, max, spring stiffness;  chains: per chain tran7, njoints, nlinks; link_last_joint / link_body per link
//                   (chain-major); dof_type: 0 Rigid, 1 BreakAway, 2 Compliant; dof_velocity: <defaultVelocity>
int ref_ag_define_hand(void *h, int nchains, const double *chain_tran7, const int *chain_njoints, const int *chain_nlinks,
                       const double *joints, const int *link_last_joint, const int *link_body, int ndof, const int *dof_type,
                       const double *dof_velocity, int base_body, int mount_body, double *dof_min_out, double *dof_max_out) {
  AgWorld *W = static_cast<AgWorld *>(h);
  Hand &R = W->hand;
  W->base = base_body; W->mount = mount_body;
  int jn = 0, ln = 0;
  std::vector<std::vector<Joint *> > byDof(ndof);
  for (int c = 0; c < nchains; c++) {
    KinematicChain *kc = new KinematicChain();
    kc->owner = &R; kc->world = &W->world; kc->tran = mk7(chain_tran7 + 7 * c); kc->firstJointNum = jn;
    for (int j = 0; j < chain_njoints[c]; j++, jn++) {
      const double *p = joints + 11 * jn;
      Joint *J = new Joint();
      J->num = jn; J->val = 0.0; J->ratio = p[1]; J->mn = p[8]; J->mx = p[9]; J->stiffness = p[10];
      const double row[6] = {p[2], p[3], p[4], p[5], p[6], p[7]};
      memcpy(J->dh, row, sizeof row);
      kc->jointVec.push_back(J);
      byDof[(int)p[0]].push_back(J);
    }
    for (int l = 0; l < chain_nlinks[c]; l++, ln++) {
      kc->lastJoint.push_back(link_last_joint[ln]);
      kc->linkVec.push_back(static_cast<Link *>(W->bodies[link_body[ln]]));
    }
    R.chainVec.push_back(kc);
  }
  R.numChains = nchains; R.numJoints = jn; R.numDOF = ndof;
  g_numJoints = jn;
  for (int d = 0; d < ndof; d++) {
    DOF *D = dof_type[d] == 1 ? (DOF *)new Spy<BreakAwayDOF>() : dof_type[d] == 2 ? (DOF *)new Spy<CompliantDOF>() : (DOF *)new Spy<RigidDOF>();
    D->initDOF(&R, byDof[d]);                               // the reference's lines: joint list, min / max from the joint limits
    DofAccess::setVel(D, dof_velocity[d]);
    R.dofVec.push_back(D);
    if (dof_min_out) dof_min_out[d] = D->getMin();
    if (dof_max_out) dof_max_out[d] = D->getMax();
  }
  return jn;
}

// Hand::autoGrasp(false, speedFactor, stopAtContact) from the posture (base7, dof0).
//   info[0] = autoGrasp's return value, [1] = successful jumpDOFToContact calls, [2] = "Interpolation failed!" seen,
//   [3] = "MoveDOF failsafe hit" seen;  stopped_out = stoppedJoints as last shown to the DOFs
int ref_ag_autograsp(void *h, const double *base7, const double *dof0, double speedFactor, int stopAtContact,
                     double *dof_out, double *joints_out, int *stopped_out, int *info) {
  AgWorld *W = static_cast<AgWorld *>(h);
  Hand &R = W->hand;
  g_dbga.clear();
  g_lastStopped.assign(R.numJoints, 0);
  R.iterations = 0;
  R.baseTran = mk7(base7);
  if (W->base >= 0) { W->bodies[W->base]->setTranRaw(R.baseTran); W->ci.setBodyTransform(W->bodies[W->base], R.baseTran); }
  if (W->mount >= 0) { W->bodies[W->mount]->setTranRaw(R.baseTran); W->ci.setBodyTransform(W->bodies[W->mount], R.baseTran); }
  // the posture HandObjectState::execute leaves: DOFs forced to their values (Robot::forceDOFVals, robot.cpp:1830-1846:
  // reset, updateVal, joint values from the DOFs, link poses)
  std::vector<double> jv(R.numJoints, 0.0);
  for (int d = 0; d < R.numDOF; d++) { R.dofVec[d]->reset(); R.dofVec[d]->updateVal(dof0[d]); R.dofVec[d]->getJointValues(jv.data()); }
  R.setJointValuesAndUpdate(jv.data());
  const bool moved = R.autoGrasp(false, speedFactor, stopAtContact != 0);
  R.getDOFVals(dof_out);
  R.getJointValues(joints_out);
  for (int j = 0; j < R.numJoints; j++) stopped_out[j] = g_lastStopped[j];
  info[0] = moved ? 1 : 0; info[1] = R.iterations; info[2] = 0; info[3] = 0;
  for (const std::string &m : g_dbga) {
    if (m.find("Interpolation failed") != std::string::npos) info[2] = 1;
    if (m.find("failsafe") != std::string::npos) info[3] = 1;
  }
  return 0;
}

} // extern "C"
