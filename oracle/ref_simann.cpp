// TEST INFRASTRUCTURE ONLY -- the oracle.  Never imported, linked or executed by the product path.
//
// Reference-pinned simulated annealing: the reference's OWN SimAnn (src/EGPlanner/simAnn.cpp),
// GraspPlanningState / VariableSet / SearchVariable (searchState.cpp, searchStateImpl.cpp) and SimAnnParams
// (simAnnParams.cpp), compiled verbatim by oracle/Makefile against the stand-in Hand / World / Coin / Qt headers of
// oracle/stub_planner and oracle/shim, driven here with a closed-form SearchEnergy so that a trajectory depends on
// nothing but SimAnn itself and libc's rand() stream.  oracle/port/simann.py, fed the same rand() stream, has to
// reproduce it (tests/test_simann_port.py); the CUDA annealer is held to the port (tests/test_gpu_anneal.py).
//
// What IS written here: the closed-form energy, SearchEnergy's base members (constructor = the reference's lines), and
// the C entry point.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "graspit/EGPlanner/searchState.h"
#include "graspit/EGPlanner/searchStateImpl.h"
#include "graspit/EGPlanner/energy/searchEnergy.h"
#include "graspit/EGPlanner/simAnn.h"
#include "graspit/EGPlanner/simAnnParams.h"
#include "graspit/robot.h"
#include "graspit/body.h"
#include "graspit/world.h"
#include "graspit/graspitCore.h"
#include "graspit/debug.h"

GraspitCore *graspitCore = NULL;                                        // src/main.cpp
SearchEnergyFactory *SearchEnergyFactory::searchEnergyFactory = NULL;   // searchEnergyFactory.cpp:35
#include "extract/searchEnergy_ctor.inc"
SearchEnergy::~SearchEnergy() {}
void SearchEnergy::createQualityMeasures() {}
void SearchEnergy::setHandAndObject(Hand *h, Body *o) { mHand = h; mObject = o; }
bool SearchEnergy::legal() const { return true; }
double SearchEnergy::getEpsQual() { return 0.0; }
double SearchEnergy::getVolQual() { return 0.0; }
void SearchEnergy::analyzeCurrentPosture(Hand *, Body *, bool &isLegal, double &stateEnergy, bool) { isLegal = true; stateEnergy = 0; }
void SearchEnergy::analyzeState(bool &isLegal, double &stateEnergy, const GraspPlanningState *, bool) { isLegal = true; stateEnergy = 0; }

namespace {
// variables in the order SimAnn::stateNeighbor draws them and the engine lays them out: position first, then posture
// (HandObjectState::getVariable(i) counts the posture first, searchState.cpp:541-548)
const SearchVariable *var(const GraspPlanningState *s, int i) {
  const int np = s->readPosition()->getNumVariables();
  return i < np ? s->readPosition()->getVariable(i) : s->readPosture()->getVariable(i - np);
}
SearchVariable *var(GraspPlanningState *s, int i) {
  const int np = s->getPosition()->getNumVariables();
  return i < np ? s->getPosition()->getVariable(i) : s->getPosture()->getVariable(i - np);
}
// energy = sum_i w_i (x_i - c_i)^2 over all variables (position first, then posture); a state is illegal when
// sin(3 x_0) * cos(2 x_last) > cut -- cheap, deterministic, and it exercises the retry loop of SimAnn::iterate.
class ClosedFormEnergy : public SearchEnergy
{
  public:
    std::vector<double> w, c;
    double cut;
    virtual double energy() const { return 0.0; }
    virtual void analyzeState(bool &isLegal, double &stateEnergy, const GraspPlanningState *state, bool = true) {
      const int n = state->getNumVariables();
      double e = 0.0;
      for (int i = 0; i < n; i++) { const double d = var(state, i)->getValue() - c[i]; e += w[i] * d * d; }
      const double g = sin(3.0 * var(state, 0)->getValue()) * cos(2.0 * var(state, n - 1)->getValue());
      isLegal = !(g > cut);
      stateEnergy = e;
    }
};
} // namespace

extern "C" {

// Runs `iters` calls of SimAnn::iterate on one state in the SPACE_AXIS_ANGLE + POSE_EIGEN layout (6 + neg variables)
// after srand(seed).  preset: 0 DEFAULT, 1 LOOP, 2 MODIFIED, 3 STRICT, 4 ONLINE.
//   start[nvar], w[nvar], c[nvar]; out_result[iters] = SimAnn::Result (0 FAIL, 1 JUMP, 2 KEEP as the enum orders them);
//   out_state[iters][nvar], out_energy[iters] = the current state after each call; out_step[iters] = getCurrentStep()
//   table[nvar][4] = mMinVal, mMaxVal, mMaxJump, circular of the reference's own SearchVariables (for the port's table)
int ref_simann_run(int neg, unsigned seed, int preset, int iters, const double *start, const double *w, const double *c, double cut,
                   double start_energy, int *out_result, double *out_state, double *out_energy, int *out_step, double *table) {
  World world; GraspitCore core; core.setWorld(&world); graspitCore = &core;
  EigenGraspInterface eg;
  std::vector<double> egv((size_t)neg * 4, 0.0), origin(4, 0.0), norm(4, 1.0);
  eg.configure(neg, 4, egv.data(), origin.data(), norm.data());
  Hand hand(&world, "hand");
  hand.configure(std::vector<double>(4, 0.0), std::vector<double>(4, 1.0), transf::IDENTITY, &eg);
  GraspableBody object("object");
  GraspPlanningState cur(&hand);
  cur.setObject(&object);
  cur.setPositionType(SPACE_AXIS_ANGLE);
  cur.setPostureType(POSE_EIGEN);
  cur.setRefTran(transf::IDENTITY);
  const int nvar = cur.getNumVariables();
  if (nvar != 6 + neg) return -1;
  for (int i = 0; i < nvar; i++) {
    SearchVariable *v = var(&cur, i);
    v->setValue(start[i]);
    if (table) { table[4 * i] = v->mMinVal; table[4 * i + 1] = v->mMaxVal; table[4 * i + 2] = v->mMaxJump; table[4 * i + 3] = v->isCircular() ? 1.0 : 0.0; }
  }
  cur.setEnergy(start_energy);
  ClosedFormEnergy E;
  E.w.assign(w, w + nvar); E.c.assign(c, c + nvar); E.cut = cut;
  SimAnn sa;
  SimAnnParams p = preset == 1 ? SimAnnParams::ANNEAL_LOOP() : preset == 2 ? SimAnnParams::ANNEAL_MODIFIED() :
                   preset == 3 ? SimAnnParams::ANNEAL_STRICT() : preset == 4 ? SimAnnParams::ANNEAL_ONLINE() : SimAnnParams::ANNEAL_DEFAULT();
  sa.setParameters(p);
  sa.reset();
  srand(seed);                                       // SimAnn's constructor seeded with time(NULL) (simAnn.cpp:85)
  for (int it = 0; it < iters; it++) {
    const SimAnn::Result r = sa.iterate(&cur, &E, NULL);
    out_result[it] = (int)r;
    for (int i = 0; i < nvar; i++) out_state[(size_t)it * nvar + i] = var(&cur, i)->getValue();
    out_energy[it] = cur.getEnergy();
    out_step[it] = sa.getCurrentStep();
  }
  graspitCore = NULL;
  return nvar;
}

} // extern "C"
