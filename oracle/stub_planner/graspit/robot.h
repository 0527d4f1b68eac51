// TEST INFRASTRUCTURE ONLY (oracle stub) -- never linked into the product.
// Shadows include/graspit/robot.h (Qt / Coin / dynamics-bound) with the members the reference's search-state code
// (src/EGPlanner/searchState.cpp, searchStateImpl.cpp) and the product's batched SearchEnergy / planner bindings use.
// A Hand here is plain data: the posture it was last given, limits, eigengrasps, approach transform.
#ifndef B2C_ORACLE_STUB_ROBOT_H
#define B2C_ORACLE_STUB_ROBOT_H
#include <vector>
#include <string>
#include <QString>
#include "graspit/mytools.h"
#include "graspit/matvec3D.h"
#include "graspit/body.h"
#include "graspit/eigenGrasp.h"

class World;
#include "graspit/gloveInterface.h"

class DOF {
    double mMin, mMax;
  public:
    DOF(double mn = 0, double mx = 0) : mMin(mn), mMax(mx) {}
    double getMin() const { return mMin; }
    double getMax() const { return mMax; }
    double getVal() const { return 0.0; }
};

class Robot {
  protected:
    std::vector<DOF> mDofs;
    std::vector<double> mDofVals, mDesired;
    transf mTran, mApproach, mSavedTran;
    std::vector<double> mSavedDofs;
    EigenGraspInterface *mEigenGrasps;
    World *mWorld;
    std::string mName;
  public:
    Robot(World *w = NULL, const char *name = "hand") : mEigenGrasps(NULL), mWorld(w), mName(name) {}
    // cloning / rendering hooks of EGPlanner::createAndUseClone (egPlanner.cpp:205-233): the harness never clones
    void cloneFrom(const Robot *o) { *this = *o; }
    void setRenderGeometry(bool) {}
    bool getRenderGeometry() const { return false; }
    // Robot::saveState / restoreState (robot.h:405-408): the posture SearchEnergy::analyzeState puts back
    void saveState() { mSavedTran = mTran; mSavedDofs = mDofVals; }
    void restoreState() { mTran = mSavedTran; mDofVals = mSavedDofs; }
    void showVirtualContacts(bool) {}
    void setTransparency(float) {}
    GloveInterface *getGloveInterface() { static GloveInterface g; return &g; }
    void setName(const QString &n) { mName = n.toStdString(); }
    virtual ~Robot() {}
    int getNumDOF() const { return (int)mDofs.size(); }
    const DOF *getDOF(int i) const { return &mDofs[i]; }
    EigenGraspInterface *getEigenGrasps() const { return mEigenGrasps; }
    const transf &getTran() const { return mTran; }
    int setTran(const transf &t) { mTran = t; return SUCCESS; }
    void storeDOFVals(double *d) const { for (size_t i = 0; i < mDofVals.size(); i++) d[i] = mDofVals[i]; }
    void getDOFVals(double *d) const { storeDOFVals(d); }
    void forceDOFVals(double *d) { for (size_t i = 0; i < mDofVals.size(); i++) mDofVals[i] = d[i]; }
    void setDesiredDOFVals(double *d) { mDesired.assign(d, d + mDofs.size()); }
    const transf &getApproachTran() const { return mApproach; }
    World *getWorld() const { return mWorld; }
    bool isA(const char *) const { return false; }
    QString getName() const { return QString(mName); }
    // Robot::checkSetDOFVals (include/graspit/robot.h:698-709): clamp to the limits
    bool checkSetDOFVals(double *dofVals) const {
      bool atLeastOneValid = false;
      for (int d = 0; d < getNumDOF(); d++) {
        if (dofVals[d] > mDofs[d].getMax()) { dofVals[d] = mDofs[d].getMax(); }
        else if (dofVals[d] < mDofs[d].getMin()) { dofVals[d] = mDofs[d].getMin(); }
        else { atLeastOneValid = true; }
      }
      return atLeastOneValid;
    }
    // harness side
    void configure(const std::vector<double> &mn, const std::vector<double> &mx, const transf &approach, EigenGraspInterface *eg) {
      mDofs.clear();
      for (size_t i = 0; i < mn.size(); i++) mDofs.push_back(DOF(mn[i], mx[i]));
      mDofVals.assign(mn.size(), 0.0); mApproach = approach; mEigenGrasps = eg;
    }
    const std::vector<double> &dofVals() const { return mDofVals; }
};

class Hand : public Robot {
  public:
    Hand(World *w = NULL, const char *name = "hand") : Robot(w, name) {}
};
class Barrett : public Hand { public: Barrett(World *w, const char *n) : Hand(w, n) {} };
class Pr2Gripper : public Hand { public: Pr2Gripper(World *w, const char *n) : Hand(w, n) {} };
class Pr2Gripper2010 : public Hand { public: Pr2Gripper2010(World *w, const char *n) : Hand(w, n) {} };
class RobotIQ : public Hand { public: RobotIQ(World *w, const char *n) : Hand(w, n) {} };
#endif
