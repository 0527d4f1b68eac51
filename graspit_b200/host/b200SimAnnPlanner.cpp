// See b200SimAnnPlanner.h.
#include "b200SimAnnPlanner.h"

#include <cstring>

#include "graspit/EGPlanner/searchState.h"
#include "graspit/EGPlanner/energy/searchEnergy.h"
#include "graspit/EGPlanner/simAnn.h"
#include "graspit/robot.h"
#include "b200ContactEnergy.h"

//#define GRASPITDBG
#include "graspit/debug.h"

static const int B200_BEST_LIST_SIZE = 20;      // BEST_LIST_SIZE, simAnnPlanner.cpp:36

B200SimAnnPlanner::B200SimAnnPlanner(Hand *h, b2c_ctx *ctx, int robotId, int objectBodyId, int numChains, unsigned long long seed,
                                     int rank, int world)
  : SimAnnPlanner(h), mCtx(ctx), mRobotId(robotId), mObjectId(objectBodyId), mNumChains(numChains), mAnnealer(-1), mRank(rank),
    mWorld(world), mSeed(seed), mParams(SimAnnParams::ANNEAL_DEFAULT())
{
  // the energy calculator the base class asked the factory for evaluates one state per call on the CPU; single states
  // (EGPlanner::getGrasp, showGrasp ...) go through the engine as well
  delete mEnergyCalculator;
  B200ContactEnergy *e = new B200ContactEnergy();
  e->setEngine(ctx, robotId, objectBodyId);
  mEnergyCalculator = e;
}

B200SimAnnPlanner::~B200SimAnnPlanner()
{
  if (mAnnealer >= 0) {b2c_anneal_destroy(mCtx, mAnnealer);}
  if (mCtx) {b2c_exchange_destroy(mCtx);}
}

void B200SimAnnPlanner::setAnnealingParameters(SimAnnParams params)
{
  SimAnnPlanner::setAnnealingParameters(params);
  mParams = params;
}

bool B200SimAnnPlanner::initialized()
{
  return SimAnnPlanner::initialized() && mCtx != NULL;
}

bool B200SimAnnPlanner::createAnnealer()
{
  if (mAnnealer >= 0) {b2c_anneal_destroy(mCtx, mAnnealer); mAnnealer = -1;}
  const GraspPlanningState *s = mCurrentState;
  if (B200ContactEnergy::inputModeOf(s) != B2C_INPUT_AA_EIGEN) {
    mError = "the device annealer searches SPACE_AXIS_ANGLE + POSE_EIGEN states (the planner dialog's default)";
    return false;
  }
  // the search variables, position first (the engine's layout), with the limits and jumps the reference gave them
  const int npos = s->readPosition()->getNumVariables(), npost = s->readPosture()->getNumVariables(), nv = npos + npost;
  std::vector<double> vmin(nv), vmax(nv), vjump(nv);
  std::vector<unsigned char> circ(nv);
  std::vector<float> init((size_t)mNumChains * nv);
  for (int i = 0; i < nv; i++) {
    const SearchVariable *v = i < npos ? s->readPosition()->getVariable(i) : s->readPosture()->getVariable(i - npos);
    vmin[i] = v->mMinVal; vmax[i] = v->mMaxVal; vjump[i] = v->mMaxJump; circ[i] = v->isCircular() ? 1 : 0;
    // every chain starts from the planner's current state, like SimAnnPlanner's single chain (resetParameters)
    for (int c = 0; c < mNumChains; c++) {init[(size_t)c * nv + i] = (float)v->getValue();}
  }
  const transf ref = s->getTotalTran() % s->readPosition()->getCoreTran().inverse();
  b2c_anneal_desc d;
  memset(&d, 0, sizeof d);
  d.robot = mRobotId; d.object = mObjectId; d.nchain = mNumChains; d.nvar = nv;
  d.vmin = vmin.data(); d.vmax = vmax.data(); d.vjump = vjump.data(); d.circular = circ.data();
  d.ref_q[0] = ref.rotation().w(); d.ref_q[1] = ref.rotation().x(); d.ref_q[2] = ref.rotation().y(); d.ref_q[3] = ref.rotation().z();
  d.ref_t[0] = ref.translation().x(); d.ref_t[1] = ref.translation().y(); d.ref_t[2] = ref.translation().z();
  d.YC = mParams.YC; d.HC = mParams.HC; d.YDIMS = (int)mParams.YDIMS; d.HDIMS = (int)mParams.HDIMS;
  d.NBR_ADJ = mParams.NBR_ADJ; d.ERR_ADJ = mParams.ERR_ADJ; d.T0 = mParams.DEF_T0; d.K0 = (int)mParams.DEF_K0;
  d.seed = mSeed; d.chain_offset = mRank * mNumChains;
  d.flags = mEnergyCalculator->getContactType() == CONTACT_LIVE ? B2C_EVAL_LIVE : B2C_EVAL_PRESET;
  if (b2c_anneal_create(mCtx, &d, init.data(), &mAnnealer) != B2C_OK) {mError = b2c_last_error(mCtx); mAnnealer = -1; return false;}
  if (mWorld == 1) {
    // a lone GPU still reads its best rows through the exchange (one slot, no peers)
    if (b2c_exchange_create(mCtx, 0, 1, 4, B200_BEST_LIST_SIZE, nv + 1, NULL, NULL) != B2C_OK) {mError = b2c_last_error(mCtx); return false;}
  }
  mRows.assign((size_t)mWorld * B200_BEST_LIST_SIZE * (nv + 1), 0.f);
  mStepsSeen.assign(mWorld, 0);
  return true;
}

void B200SimAnnPlanner::resetParameters()
{
  SimAnnPlanner::resetParameters();
  createAnnealer();
}

void B200SimAnnPlanner::mergeRows()
{
  const int npos = mCurrentState->readPosition()->getNumVariables(), npost = mCurrentState->readPosture()->getNumVariables();
  const int nv = npos + npost, n = mWorld * B200_BEST_LIST_SIZE;
  for (int k = 0; k < n; k++) {
    const float *r = &mRows[(size_t)k * (nv + 1)];
    const double e = r[nv];
    if (!(e < 1.0e7)) {continue;}                                  // empty slot
    // the reference's JUMP branch (simAnnPlanner.cpp:127-143), state by state
    double worstEnergy;
    if ((int)mBestList.size() < B200_BEST_LIST_SIZE) {worstEnergy = 1.0e5;}
    else {worstEnergy = mBestList.back()->getEnergy();}
    if (!(e < worstEnergy)) {continue;}
    GraspPlanningState *insertState = new GraspPlanningState(mCurrentState);
    for (int i = 0; i < npos; i++) {insertState->getPosition()->getVariable(i)->setValue(r[i]);}
    for (int i = 0; i < npost; i++) {insertState->getPosture()->getVariable(i)->setValue(r[npos + i]);}
    insertState->setEnergy(e);
    insertState->setLegal(true);
    insertState->setItNumber(mCurrentStep);
    if (!addToListOfUniqueSolutions(insertState, &mBestList, 0.2)) {
      delete insertState;
    } else {
      mBestList.sort(GraspPlanningState::compareStates);
      while ((int)mBestList.size() > B200_BEST_LIST_SIZE) {
        delete(mBestList.back());
        mBestList.pop_back();
      }
    }
  }
}

void B200SimAnnPlanner::mainLoop()
{
  if (mAnnealer < 0 && !createAnnealer()) {DBGA("B200SimAnnPlanner: " << mError); return;}
  if (b2c_anneal_step(mCtx, mAnnealer, 1) != B2C_OK || b2c_anneal_publish(mCtx, mAnnealer, NULL) != B2C_OK ||
      b2c_exchange_collect(mCtx, mRows.data(), mStepsSeen.data(), 0) != B2C_OK) {
    mError = b2c_last_error(mCtx);
    DBGA("B200SimAnnPlanner: " << mError);
    return;
  }
  mCurrentStep++;
  mergeRows();
  render();
  if (mCurrentStep % 100 == 0 && !mMultiThread) {Q_EMIT update();}
}

bool B200SimAnnPlanner::step()
{
  const int before = mCurrentStep;
  mainLoop();
  return mCurrentStep == before + 1;
}
