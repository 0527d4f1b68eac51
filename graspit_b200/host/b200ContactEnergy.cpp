// See b200ContactEnergy.h.
#include "b200ContactEnergy.h"

#include <cmath>
#include <cstring>

#include "graspit/EGPlanner/searchState.h"
#include "graspit/robot.h"
#include "graspit/body.h"

B200ContactEnergy::B200ContactEnergy() : mCtx(NULL), mRobotId(-1), mObjectId(-1), mLastLegal(false), mLastEnergy(0.0)
{
  mContactType = CONTACT_PRESET;
}

void B200ContactEnergy::setEngine(b2c_ctx *ctx, int robotId, int objectBodyId)
{
  mCtx = ctx; mRobotId = robotId; mObjectId = objectBodyId;
}

int B200ContactEnergy::inputModeOf(const GraspPlanningState *state)
{
  const StateType pos = state->readPosition()->getType(), post = state->readPosture()->getType();
  if (post == POSE_EIGEN) {
    if (pos == SPACE_AXIS_ANGLE) {return B2C_INPUT_AA_EIGEN;}
    if (pos == SPACE_ELLIPSOID) {return B2C_INPUT_ELLIPSOID_EIGEN;}
    if (pos == SPACE_APPROACH) {return B2C_INPUT_APPROACH_EIGEN;}
  }
  if (post == POSE_DOF && pos == SPACE_COMPLETE) {return B2C_INPUT_COMPLETE_DOF;}
  return -1;
}

static void putTransf(const transf &t, double q[4], double tr[3])
{
  q[0] = t.rotation().w(); q[1] = t.rotation().x(); q[2] = t.rotation().y(); q[3] = t.rotation().z();
  tr[0] = t.translation().x(); tr[1] = t.translation().y(); tr[2] = t.translation().z();
}

bool B200ContactEnergy::analyzeStates(const std::vector<const GraspPlanningState *> &states, std::vector<char> *legal,
                                      std::vector<double> *energy) const
{
  const int n = (int)states.size();
  legal->assign(n, 0); energy->assign(n, 0.0);
  if (n == 0) {return true;}
  if (!mCtx || mRobotId < 0) {mError = "B200ContactEnergy: setEngine() was not called"; return false;}
  // one layout and one reference transform per call: what a planner's states share (they are copies of its model state)
  const int mode = inputModeOf(states[0]);
  const int npos = states[0]->readPosition()->getNumVariables(), npost = states[0]->readPosture()->getNumVariables();
  // mRefTran is not exposed; it is what getTotalTran() = mRefTran % getCoreTran() leaves (searchState.h:309)
  const transf ref = states[0]->getTotalTran() % states[0]->readPosition()->getCoreTran().inverse();
  bool uniform = mode >= 0;
  for (int i = 1; i < n && uniform; i++) {
    uniform = inputModeOf(states[i]) == mode && states[i]->readPosition()->getNumVariables() == npos &&
              states[i]->readPosture()->getNumVariables() == npost;
    if (uniform) {
      const transf r = states[i]->getTotalTran() % states[i]->readPosition()->getCoreTran().inverse();
      const double qd = r.rotation().w() * ref.rotation().w() + r.rotation().x() * ref.rotation().x() +
                        r.rotation().y() * ref.rotation().y() + r.rotation().z() * ref.rotation().z();
      uniform = (r.translation() - ref.translation()).norm() < 1.0e-9 && std::fabs(std::fabs(qd) - 1.0) < 1.0e-12;
    }
  }
  b2c_eval_args a;
  memset(&a, 0, sizeof a);
  a.robot = mRobotId; a.object = mObjectId; a.npose = n;
  a.flags = mContactType == CONTACT_LIVE ? B2C_EVAL_LIVE : B2C_EVAL_PRESET;
  std::vector<float> x;
  if (uniform) {
    a.input_mode = mode; a.stride = npos + npost;
    x.resize((size_t)n * a.stride);
    for (int i = 0; i < n; i++) {
      float *row = &x[(size_t)i * a.stride];
      for (int v = 0; v < npos; v++) {row[v] = (float)states[i]->readPosition()->readVariable(v);}
      for (int v = 0; v < npost; v++) {row[npos + v] = (float)states[i]->readPosture()->readVariable(v);}
    }
    putTransf(ref, a.ref_q, a.ref_t);
    if (mode == B2C_INPUT_ELLIPSOID_EIGEN) {
      a.position_params[0] = states[0]->readPosition()->getParameter("a");
      a.position_params[1] = states[0]->readPosition()->getParameter("b");
      a.position_params[2] = states[0]->readPosition()->getParameter("c");
    }
  } else {
    // the reference's own mapping on the host: hand transform and DOF values per state (HandObjectState::execute)
    const int nd = states[0]->getHand()->getNumDOF();
    a.input_mode = B2C_INPUT_RAW; a.stride = 7 + nd;
    x.resize((size_t)n * a.stride);
    std::vector<double> dof(nd);
    for (int i = 0; i < n; i++) {
      float *row = &x[(size_t)i * a.stride];
      double q[4], t[3];
      putTransf(states[i]->getTotalTran(), q, t);
      for (int k = 0; k < 4; k++) {row[k] = (float)q[k];}
      for (int k = 0; k < 3; k++) {row[4 + k] = (float)t[k];}
      states[i]->readPosture()->getHandDOF(dof.data());
      for (int d = 0; d < nd; d++) {row[7 + d] = (float)dof[d];}
    }
    a.ref_q[0] = 1.0;
  }
  a.states = x.data();
  std::vector<unsigned char> lg(n);
  std::vector<float> en(n);
  a.legal = lg.data(); a.energy = en.data();
  if (b2c_eval_batch(mCtx, &a) != B2C_OK) {mError = b2c_last_error(mCtx); return false;}
  for (int i = 0; i < n; i++) {(*legal)[i] = (char)lg[i]; (*energy)[i] = en[i];}
  return true;
}

void B200ContactEnergy::analyzeState(bool &isLegal, double &stateEnergy, const GraspPlanningState *state, bool noChange)
{
  // SearchEnergy::analyzeState (searchEnergy.cpp:148-188): a state on the avoid list is illegal outright
  if (mAvoidList) {
    std::list<GraspPlanningState *>::const_iterator it;
    int i = 0;
    for (it = mAvoidList->begin(); it != mAvoidList->end(); it++) {
      if ((*it)->distance(state) < mThreshold) {isLegal = false; stateEnergy = 0.0; return;}
      i++;
    }
  }
  std::vector<const GraspPlanningState *> one(1, state);
  std::vector<char> lg; std::vector<double> en;
  if (!analyzeStates(one, &lg, &en)) {isLegal = false; stateEnergy = 0.0; return;}
  isLegal = lg[0] != 0;
  stateEnergy = isLegal ? en[0] : 0.0;
  // the reference leaves the world in the analysed state when asked to (noChange == false and the state is legal)
  if (!noChange && isLegal) {state->execute();}
}

bool B200ContactEnergy::evalCurrent() const
{
  mLastLegal = false; mLastEnergy = 0.0;
  if (!mCtx || mRobotId < 0 || !mHand) {mError = "B200ContactEnergy: no engine / no hand"; return false;}
  const int nd = mHand->getNumDOF();
  std::vector<float> row(7 + nd);
  std::vector<double> dof(nd);
  double q[4], t[3];
  putTransf(mHand->getTran(), q, t);
  for (int k = 0; k < 4; k++) {row[k] = (float)q[k];}
  for (int k = 0; k < 3; k++) {row[4 + k] = (float)t[k];}
  mHand->getDOFVals(dof.data());
  for (int d = 0; d < nd; d++) {row[7 + d] = (float)dof[d];}
  b2c_eval_args a;
  memset(&a, 0, sizeof a);
  a.robot = mRobotId; a.object = mObjectId; a.npose = 1; a.input_mode = B2C_INPUT_RAW; a.states = row.data(); a.stride = 7 + nd;
  a.ref_q[0] = 1.0; a.flags = mContactType == CONTACT_LIVE ? B2C_EVAL_LIVE : B2C_EVAL_PRESET;
  unsigned char lg = 0; float en = 0.f;
  a.legal = &lg; a.energy = &en;
  if (b2c_eval_batch(mCtx, &a) != B2C_OK) {mError = b2c_last_error(mCtx); return false;}
  mLastLegal = lg != 0; mLastEnergy = en;
  return true;
}

bool B200ContactEnergy::legal() const
{
  evalCurrent();
  return mLastLegal;
}

double B200ContactEnergy::energy() const
{
  return mLastEnergy;          // legal() ran first (SearchEnergy::analyzeCurrentPosture, searchEnergy.cpp:127-146)
}

void B200ContactEnergy::analyzeCurrentPosture(Hand *h, Body *o, bool &isLegal, double &stateEnergy, bool)
{
  mHand = h; mObject = o;
  isLegal = legal();
  stateEnergy = isLegal ? energy() : 0.0;
}
