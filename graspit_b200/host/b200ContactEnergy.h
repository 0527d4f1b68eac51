// CONTACT_ENERGY on the B200 engine behind GraspIt!'s own SearchEnergy interface
// (reference include/graspit/EGPlanner/energy/searchEnergy.h:44-121, src/EGPlanner/energy/searchEnergy.cpp:104-188,
// src/EGPlanner/energy/contactEnergy.cpp:9-55).  Compiles against the reference's unmodified headers.
//
//   analyzeState(isLegal, energy, state, noChange)   the reference's entry point, one state: EGPlanner / SimAnn /
//                                                    SimAnnPlanner::mainLoop call it unchanged
//   analyzeStates(states, legal, energy)             the batched entry point: every state in ONE b2c_eval_batch call
//
// A state reaches the engine as its search variables (position variables, then posture variables) in the layout the
// engine maps on the device -- SPACE_AXIS_ANGLE / SPACE_ELLIPSOID / SPACE_APPROACH + POSE_EIGEN, SPACE_COMPLETE +
// POSE_DOF -- together with mRefTran; any other combination goes through the reference's own
// HandObjectState::getTotalTran / PostureState::getHandDOF and reaches the engine as (hand transform, DOF values).
// Registered with the factory as "B200_CONTACT_ENERGY" (REGISTER_SEARCH_ENERGY_CREATOR, searchEnergyFactory.cpp:39-47).
#ifndef B200_CONTACT_ENERGY_H
#define B200_CONTACT_ENERGY_H

#include <vector>

#include "graspit/EGPlanner/energy/searchEnergy.h"
#include "b2c.h"

class GraspPlanningState;

class B200ContactEnergy : public SearchEnergy
{
    b2c_ctx *mCtx;
    int mRobotId, mObjectId;
    mutable bool mLastLegal;
    mutable double mLastEnergy;
    mutable std::string mError;
    //! the hand as it stands (Hand::getTran, Hand::getDOFVals) through the engine; caches legality and energy
    bool evalCurrent() const;
  protected:
    virtual double energy() const;
    virtual bool legal() const;
  public:
    B200ContactEnergy();
    //! the engine context that holds the world, the hand's robot id and the object's body id in it
    void setEngine(b2c_ctx *ctx, int robotId, int objectBodyId);
    const std::string &error() const {return mError;}

    virtual void analyzeState(bool &isLegal, double &stateEnergy, const GraspPlanningState *state, bool noChange = true);
    virtual void analyzeCurrentPosture(Hand *h, Body *o, bool &isLegal, double &stateEnergy, bool noChange = true);
    //! all states in one call; legal[i] / energy[i] are what analyzeState(.., states[i], true) returns.  False on an
    //! engine error (see error())
    bool analyzeStates(const std::vector<const GraspPlanningState *> &states, std::vector<char> *legal,
                       std::vector<double> *energy) const;
    //! B2C_INPUT_* layout of a state's variables, or -1 when the engine has no device-side mapping for it
    static int inputModeOf(const GraspPlanningState *state);
};

#endif
