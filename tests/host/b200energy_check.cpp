// TEST -- the batched SearchEnergy / SimAnnPlanner bindings (graspit_b200/host/b200ContactEnergy.{h,cpp},
// b200SimAnnPlanner.{h,cpp}) A/B against the reference's own planner stack.
//
// What runs here is the REFERENCE's code wherever it could be compiled: GraspPlanningState / VariableSet / SearchVariable
// (src/EGPlanner/searchState.cpp, searchStateImpl.cpp), EGPlanner, SimAnnPlanner, SimAnn, SimAnnParams
// (src/EGPlanner/egPlanner.cpp, simAnnPlanner.cpp, simAnn.cpp, simAnnParams.cpp; src/profiling.cpp) -- all verbatim,
// against the stand-in Hand / World / Coin / Qt headers of oracle/stub_planner and oracle/shim -- plus
// SearchEnergy::SearchEnergy / analyzeCurrentPosture / analyzeState, the reference's own lines cut out of
// src/EGPlanner/energy/searchEnergy.cpp (64-76, 126-146, 148-188) by oracle/Makefile.  The two SearchEnergy subclasses
// under test differ only in legal() / energy():
//   RefContactEnergy   (this file)  reference GraspitCollision + the reference's fwdKinematics lines, through the oracle
//                                   library (ref_fk_chain, ref_eval_batch: oracle/ref_fk.cpp, oracle/ref_driver.cpp)
//   B200ContactEnergy  (product)    the CUDA engine through the C ABI
//
//   b200energy_check scene.txt
//     1. analyzeStates on N reference GraspPlanningStates (one engine call) == N x RefContactEnergy::analyzeState
//     2. the other state layouts (ELLIPSOID / APPROACH + EIGEN, COMPLETE + DOF) the same way
//     3. the reference's SimAnnPlanner::mainLoop, unchanged, driven once with each energy from the same rand() seed:
//        same legal / jump sequence, energies to tolerance, same best list
//     4. B200SimAnnPlanner (K chains on the device): EGPlanner's own best list (addToListOfUniqueSolutions) filled from the
//        device rows; every listed state re-evaluated by RefContactEnergy gives its energy back
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include "graspit/EGPlanner/searchState.h"
#include "graspit/EGPlanner/searchStateImpl.h"
#include "graspit/EGPlanner/energy/searchEnergy.h"
#include "graspit/EGPlanner/simAnnPlanner.h"
#include "graspit/EGPlanner/simAnn.h"
#include "graspit/EGPlanner/simAnnParams.h"
#include "graspit/robot.h"
#include "graspit/body.h"
#include "graspit/world.h"
#include "graspit/graspitCore.h"
#include "graspit/debug.h"

#include "b200ContactEnergy.h"
#include "b200SimAnnPlanner.h"
#include "scene_dump.h"

// ---- the oracle library (reference GraspitCollision + reference FK lines) ---------------------------------------------
extern "C" {
void *ref_create();
void ref_destroy(void *w);
int ref_add_body(void *w, const double *tri9, int ntri);
void ref_set_transform(void *w, int id, const double *q_wxyz, const double *t);
void ref_activate_pair(void *w, int a, int b, int active);
void ref_eval_batch(void **worlds, int nthreads, int npose, int nhand, const int *hand_ids, int object_id, const double *tf7,
                    int nvc, const int *vc_link, const double *vc_loc, const double *vc_normal, int mode, int want_dist,
                    unsigned char *legal, double *energy, double *dist);
void ref_fk_chain(const double *owner7, const double *chain7, int njoints, const double *joints, const double *jointVals,
                  int nlinks, const int *last_joint, double *out7);
}

// ---- pieces of the reference the harness has to supply ------------------------------------------------------------------
GraspitCore *graspitCore = NULL;                                   // src/main.cpp
SearchEnergyFactory *SearchEnergyFactory::searchEnergyFactory = NULL;   // src/EGPlanner/energy/searchEnergyFactory.cpp:35
void EGPlanner::update() {}                                        // Qt signals: moc would generate the emitters
void EGPlanner::complete() {}
// SearchEnergy's constructor, analyzeCurrentPosture and analyzeState: the reference's own lines
#include "extract/searchEnergy_ctor.inc"
#include "extract/searchEnergy_analyzeCurrentPosture.inc"
#include "extract/searchEnergy_analyzeState.inc"
// the rest of the base class touches quality measures and World::noCollision: not on this path
SearchEnergy::~SearchEnergy() {}
void SearchEnergy::createQualityMeasures() {}
void SearchEnergy::setHandAndObject(Hand *h, Body *o) { mHand = h; mObject = o; }     // searchEnergy.cpp:88-96 minus the quality measures
bool SearchEnergy::legal() const { return true; }
double SearchEnergy::getEpsQual() { return 0.0; }
double SearchEnergy::getVolQual() { return 0.0; }

static int g_fail = 0, g_checks = 0;
#define CHECK(cond, ...) do { g_checks++; if (!(cond)) { g_fail++; printf("FAIL %s:%d  %s  ", __FILE__, __LINE__, #cond); printf(__VA_ARGS__); printf("\n"); } } while (0)

// ---- the reference-side SearchEnergy -------------------------------------------------------------------------------------
struct RefWorldSide {
  void *w;
  std::vector<int> ids;            // scene body -> oracle id
  const SceneDump *S;
  std::vector<int> handIds, vcLink;
  std::vector<double> vcLoc, vcNormal;
};

class RefContactEnergy : public SearchEnergy
{
    const RefWorldSide *mR;
    mutable bool mLegal;
    mutable double mEnergy;
    void evalCurrent() const {
      const SceneDump &S = *mR->S;
      double owner7[7];
      const transf &t = mHand->getTran();
      owner7[0] = t.rotation().w(); owner7[1] = t.rotation().x(); owner7[2] = t.rotation().y(); owner7[3] = t.rotation().z();
      owner7[4] = t.translation().x(); owner7[5] = t.translation().y(); owner7[6] = t.translation().z();
      std::vector<double> dof(S.ndof);
      mHand->getDOFVals(dof.data());
      // links chain-major, then the palm: World::findVirtualContacts order = the order of hand_ids
      std::vector<double> tf7(7 * (S.linkBody.size() + 1));
      for (size_t c = 0; c < S.chains.size(); c++) {
        const b2c_chain_desc &cd = S.chains[c];
        double chain7[7] = {cd.tran_q[0], cd.tran_q[1], cd.tran_q[2], cd.tran_q[3], cd.tran_t[0], cd.tran_t[1], cd.tran_t[2]};
        std::vector<double> j6(6 * cd.njoints), jv(cd.njoints);
        for (int j = 0; j < cd.njoints; j++) {
          const b2c_joint_desc &jd = S.joints[cd.first_joint + j];
          // RevoluteJoint: DHTransform(0 + c, d, a, alpha), c = the folded theta; PrismaticJoint: DHTransform(theta, 0 + c, ...)
          double row[6] = {(double)jd.revolute, jd.revolute ? 0.0 : jd.theta, jd.revolute ? jd.d : 0.0, jd.a, jd.alpha, jd.revolute ? jd.theta : jd.d};
          memcpy(&j6[6 * j], row, sizeof row);
          jv[j] = dof[jd.dof] * jd.ratio;                        // RigidDOF::accumulateMove, dof.cpp:500-514
        }
        ref_fk_chain(owner7, chain7, cd.njoints, j6.data(), jv.data(), cd.nlinks, &S.linkLastJoint[cd.first_link], &tf7[7 * cd.first_link]);
      }
      memcpy(&tf7[7 * S.linkBody.size()], owner7, sizeof owner7);
      unsigned char lg = 0; double en = 0.0;
      void *worlds[1] = {mR->w};
      ref_eval_batch(worlds, 1, 1, (int)mR->handIds.size(), mR->handIds.data(), mR->ids[S.object], tf7.data(), (int)mR->vcLink.size(),
                     mR->vcLink.data(), mR->vcLoc.data(), mR->vcNormal.data(), mContactType == CONTACT_LIVE ? 1 : 0, 0, &lg, &en, NULL);
      mLegal = lg != 0; mEnergy = en;
    }
  protected:
    virtual bool legal() const { evalCurrent(); return mLegal; }
    virtual double energy() const { return mEnergy; }
  public:
    RefContactEnergy(const RefWorldSide *r) : mR(r), mLegal(false), mEnergy(0.0) { mContactType = CONTACT_PRESET; }
};

// SearchEnergy::getSearchEnergy (searchEnergy.cpp:203-217) asks the factory; here the two energies under test are registered
static const RefWorldSide *g_ref = NULL;
static b2c_ctx *g_ctx = NULL; static int g_robot = -1, g_object = -1;
class RegisteredRefEnergy : public RefContactEnergy { public: RegisteredRefEnergy() : RefContactEnergy(g_ref) {} };
class RegisteredB200Energy : public B200ContactEnergy { public: RegisteredB200Energy() { setEngine(g_ctx, g_robot, g_object); } };
SearchEnergy *SearchEnergy::getSearchEnergy(std::string t)
{
  SearchEnergy *se = SearchEnergyFactory::getInstance()->createEnergy(t);
  if (!se) { se = SearchEnergyFactory::getInstance()->createEnergy("CONTACT_ENERGY"); }
  return se;
}

// the reference planner with its protected loop exposed
class SteppedSimAnnPlanner : public SimAnnPlanner
{
  public:
    SteppedSimAnnPlanner(Hand *h) : SimAnnPlanner(h) {}
    void useEnergy(SearchEnergy *e) { delete mEnergyCalculator; mEnergyCalculator = e; }
    void reset() { resetParameters(); }
    void step() { mainLoop(); }
    const GraspPlanningState *current() const { return mCurrentState; }
    const std::list<GraspPlanningState *> &bestList() const { return mBestList; }
};

static void configureHand(Hand *h, EigenGraspInterface *eg, const SceneDump &S)
{
  eg->configure(S.neg, S.ndof, S.eg.data(), S.egOrigin.data(), S.egNorm.data());
  transf approach(Quaternion(S.approachQ[0], S.approachQ[1], S.approachQ[2], S.approachQ[3]), vec3(S.approachT[0], S.approachT[1], S.approachT[2]));
  h->configure(S.dofMin, S.dofMax, approach, eg);
}

int main(int argc, char **argv)
{
  if (argc < 2) { fprintf(stderr, "usage: %s scene.txt\n", argv[0]); return 2; }
  SceneDump S;
  if (!readSceneDump(argv[1], &S)) { printf("FAILED cannot read %s\n", argv[1]); return 1; }
  std::vector<int> ids;
  if (!loadDumpIntoEngine(S, 0, &g_ctx, &ids, &g_robot)) { printf("FAILED engine: %s\n", g_ctx ? b2c_last_error(g_ctx) : "no device"); return 1; }
  g_object = ids[S.object];

  // the same world inside the reference backend
  RefWorldSide R;
  R.w = ref_create(); R.S = &S;
  for (size_t i = 0; i < S.bodies.size(); i++) {
    std::vector<double> t(S.bodies[i].tri9.begin(), S.bodies[i].tri9.end());
    R.ids.push_back(ref_add_body(R.w, t.data(), (int)(t.size() / 9)));
    ref_set_transform(R.w, R.ids.back(), S.bodies[i].q, S.bodies[i].t);
  }
  for (auto &p : S.disabled) ref_activate_pair(R.w, R.ids[p.first], R.ids[p.second], 0);
  for (int b : S.linkBody) R.handIds.push_back(R.ids[b]);
  R.handIds.push_back(R.ids[S.base]);
  for (auto &v : S.vcs) {
    int slot = v.body == S.base ? (int)S.linkBody.size() : (int)(std::find(S.linkBody.begin(), S.linkBody.end(), v.body) - S.linkBody.begin());
    R.vcLink.push_back(slot);
    for (int k = 0; k < 3; k++) { R.vcLoc.push_back(v.loc[k]); R.vcNormal.push_back(v.normal[k]); }
  }
  g_ref = &R;
  REGISTER_SEARCH_ENERGY_CREATOR("CONTACT_ENERGY", RegisteredRefEnergy);
  REGISTER_SEARCH_ENERGY_CREATOR("B200_CONTACT_ENERGY", RegisteredB200Energy);

  World world; GraspitCore core; core.setWorld(&world); graspitCore = &core;
  EigenGraspInterface eg;
  Hand hand(&world, "Barrett");
  configureHand(&hand, &eg, S);
  GraspableBody object("mug");
  object.setTranRaw(transf(Quaternion(S.refQ[0], S.refQ[1], S.refQ[2], S.refQ[3]), vec3(S.refT[0], S.refT[1], S.refT[2])));

  // ---- 1 + 2: analyzeStates vs N x reference analyzeState, every layout the engine maps on the device ----------------
  struct Layout { StateType pos, post; const char *name; };
  const Layout layouts[4] = {{SPACE_AXIS_ANGLE, POSE_EIGEN, "AXIS_ANGLE + EIGEN"}, {SPACE_ELLIPSOID, POSE_EIGEN, "ELLIPSOID + EIGEN"},
                             {SPACE_APPROACH, POSE_EIGEN, "APPROACH + EIGEN"}, {SPACE_COMPLETE, POSE_DOF, "COMPLETE + DOF"}};
  std::mt19937_64 rng(17);
  for (const Layout &L : layouts) {
    GraspPlanningState model(&hand);
    model.setObject(&object);
    model.setPositionType(L.pos);
    model.setPostureType(L.post);
    model.setRefTran(object.getTran());
    model.reset();
    const int N = L.pos == SPACE_AXIS_ANGLE ? 2000 : 500;
    std::vector<GraspPlanningState *> states;
    for (int i = 0; i < N; i++) {
      GraspPlanningState *s = new GraspPlanningState(&model);
      for (int v = 0; v < s->getNumVariables(); v++) {
        SearchVariable *sv = s->getVariable(v);
        double lo = sv->mMinVal, hi = sv->mMaxVal;
        if (L.pos == SPACE_COMPLETE && v >= s->readPosture()->getNumVariables() + 3) { lo = -1; hi = 1; }   // quaternion components
        // the engine's state rows are fp32: give both sides the same (fp32-representable) values
        sv->setValue((double)(float)std::uniform_real_distribution<double>(lo, hi)(rng));
      }
      states.push_back(s);
    }
    B200ContactEnergy b200;
    b200.setEngine(g_ctx, g_robot, g_object);
    RefContactEnergy refE(&R);
    std::vector<const GraspPlanningState *> cs(states.begin(), states.end());
    std::vector<char> lg; std::vector<double> en;
    CHECK(b200.analyzeStates(cs, &lg, &en), "%s: %s", L.name, b200.error().c_str());
    CHECK(B200ContactEnergy::inputModeOf(cs[0]) >= 0, "%s has a device-side mapping", L.name);
    int nlegal = 0, legalMismatch = 0; double worst = 0.0;
    for (int i = 0; i < N; i++) {
      bool rl; double re;
      refE.analyzeState(rl, re, states[i], true);                 // the reference's own analyzeState (extracted lines)
      if (rl != (lg[i] != 0)) { legalMismatch++; continue; }
      if (rl) { nlegal++; worst = std::max(worst, std::fabs(en[i] - re) / std::max(std::fabs(re), 1e-9)); }
    }
    CHECK(legalMismatch == 0, "%s: %d legality mismatches", L.name, legalMismatch);
    CHECK(nlegal > N / 20 && worst < 1.0e-4, "%s: %d legal, worst energy error %.3g", L.name, nlegal, worst);
    // single-state entry point of the product == its batched one, and it leaves the hand in the state when asked to
    for (int i = 0; i < 20; i++) {
      bool l1; double e1;
      b200.analyzeState(l1, e1, states[i], false);
      CHECK(l1 == (lg[i] != 0) && (!l1 || e1 == en[i]), "%s state %d: single %d %.9g vs batch %d %.9g", L.name, i, (int)l1, e1, (int)lg[i], en[i]);
      if (l1) {
        transf want = states[i]->getTotalTran();
        CHECK((hand.getTran().translation() - want.translation()).norm() < 1e-12, "%s state %d: hand left at the analysed state", L.name, i);
      }
    }
    printf("%-20s %d states: %d legal, 0 legality mismatches, worst energy rel err %.2e\n", L.name, N, nlegal, worst);
    for (GraspPlanningState *s : states) delete s;
  }

  // ---- 3: the reference's SimAnnPlanner, unchanged, on either energy ------------------------------------------------------
  {
    GraspPlanningState model(&hand);
    model.setObject(&object);
    model.setPositionType(SPACE_AXIS_ANGLE);
    model.setPostureType(POSE_EIGEN);
    model.setRefTran(object.getTran());
    model.reset();
    const int STEPS = 400;
    std::vector<double> trace[2]; std::vector<int> legalTrace[2];
    std::vector<double> best[2];
    for (int side = 0; side < 2; side++) {
      Hand h2(&world, "Barrett");
      EigenGraspInterface eg2;
      configureHand(&h2, &eg2, S);
      SteppedSimAnnPlanner planner(&h2);
      SearchEnergy *e = SearchEnergy::getSearchEnergy(side == 0 ? "CONTACT_ENERGY" : "B200_CONTACT_ENERGY");
      e->setContactType(CONTACT_PRESET);
      planner.useEnergy(e);
      planner.setModelState(&model);
      planner.reset();
      srand(12345);                                   // SimAnn seeds rand() with time(NULL) (simAnn.cpp:85): same stream for both
      for (int s = 0; s < STEPS; s++) {
        planner.step();
        trace[side].push_back(planner.current()->getEnergy());
        legalTrace[side].push_back(planner.getCurrentStep());
      }
      for (const GraspPlanningState *b : planner.bestList()) best[side].push_back(b->getEnergy());
    }
    int same = 0;
    double worst = 0.0;
    for (int s = 0; s < STEPS; s++) {
      if (legalTrace[0][s] != legalTrace[1][s]) break;
      same++;
      worst = std::max(worst, std::fabs(trace[0][s] - trace[1][s]) / std::max(std::fabs(trace[0][s]), 1e-9));
    }
    CHECK(same == STEPS, "SimAnnPlanner on the two energies follows the same trajectory for %d of %d iterations", same, STEPS);
    CHECK(worst < 1.0e-4, "current-state energies along the trajectory agree to %.3g", worst);
    CHECK(best[0].size() == best[1].size() && !best[0].empty(), "best lists: %zu vs %zu states", best[0].size(), best[1].size());
    for (size_t i = 0; i < std::min(best[0].size(), best[1].size()); i++)
      CHECK(std::fabs(best[0][i] - best[1][i]) <= 1.0e-4 * std::fabs(best[0][i]), "best list entry %zu: %.9g vs %.9g", i, best[0][i], best[1][i]);
    printf("SimAnnPlanner (reference mainLoop), %d iterations: identical step counters, energies within %.2e, %zu solutions\n", STEPS, worst, best[0].size());
  }

  // ---- 4: B200SimAnnPlanner: K chains on the device, EGPlanner's best list on the host --------------------------------------
  {
    GraspPlanningState model(&hand);
    model.setObject(&object);
    model.setPositionType(SPACE_AXIS_ANGLE);
    model.setPostureType(POSE_EIGEN);
    model.setRefTran(object.getTran());
    model.reset();
    B200SimAnnPlanner planner(&hand, g_ctx, g_robot, g_object, 8192, 1234);
    planner.setAnnealingParameters(SimAnnParams::ANNEAL_DEFAULT());
    planner.setModelState(&model);
    CHECK(planner.resetPlanner(), "resetPlanner");
    CHECK(planner.initialized() && planner.getType() == PLANNER_SIM_ANN, "planner type / initialised");
    const int K0 = planner.getCurrentStep();
    int okSteps = 0;
    for (int s = 0; s < 60; s++) okSteps += planner.step() ? 1 : 0;
    CHECK(okSteps == 60 && planner.getCurrentStep() == K0 + 60, "%d steps, counter %d -> %d (%s)", okSteps, K0, planner.getCurrentStep(), planner.error().c_str());
    const int n = planner.getListSize();
    CHECK(n >= 5 && n <= 20, "%d solutions in EGPlanner's best list", n);
    RefContactEnergy refE(&R);
    double prev = -1.0e300, worst = 0.0;
    for (int i = 0; i < n; i++) {
      const GraspPlanningState *g = planner.getGrasp(i);
      CHECK(g->getEnergy() >= prev, "best list sorted");
      prev = g->getEnergy();
      bool rl; double re;
      refE.analyzeState(rl, re, g, true);
      CHECK(rl, "solution %d is legal for the reference", i);
      worst = std::max(worst, std::fabs(re - g->getEnergy()) / std::max(std::fabs(re), 1e-9));
      for (int j = i + 1; j < n; j++) CHECK(g->distance(planner.getGrasp(j)) >= 0.2, "solutions %d and %d distinct (HandObjectState::distance)", i, j);
    }
    CHECK(worst < 1.0e-4, "listed energies vs the reference: %.3g", worst);
    printf("B200SimAnnPlanner: 8192 chains x 60 steps, %d solutions, best %.4f, energies vs reference within %.2e\n", n, n ? planner.getGrasp(0)->getEnergy() : 0.0, worst);
  }

  ref_destroy(R.w);
  b2c_destroy(g_ctx);
  printf("%s: %d checks, %d failed\n", g_fail ? "FAILED" : "PASSED", g_checks, g_fail);
  return g_fail ? 1 : 0;
}
