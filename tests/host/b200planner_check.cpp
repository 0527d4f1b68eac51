// TEST -- C++ host planner (graspit_b200/host/b200Planner.{h,cpp}) over the C ABI.
//   b200planner_check cpu                 host-only logic: presets, variable tables, best list, state files
//   b200planner_check gpu scene.txt       + batched annealing and ListPlanner on cuda:0 with the dumped scene
// scene.txt is written by tests/test_host_planner.py (plain numbers; triangle files as in b200collision_check).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <random>
#include <string>
#include <vector>

#include <execinfo.h>
#include <signal.h>
#include <unistd.h>

#include "b200Planner.h"

extern "C" int cudaGetDeviceCount(int *count);      // libcudart (no CUDA headers needed on the host side)

using namespace b200;
static int g_fail = 0, g_checks = 0;
#define CHECK(cond, ...) do { g_checks++; if (!(cond)) { g_fail++; printf("FAIL %s:%d  %s  ", __FILE__, __LINE__, #cond); printf(__VA_ARGS__); printf("\n"); } } while (0)

static PlannerState st(float tx, float ty, float tz, float al, double e)
{
  PlannerState s;
  s.vars = {tx, ty, tz, 0.5f, 0.2f, al, 0.f, 0.f}; s.energy = e; s.itNumber = 0;
  return s;
}

static void cpuChecks()
{
  SimAnnParams p = SimAnnParams::ANNEAL_DEFAULT();
  CHECK(p.YC == 7.0 && p.HC == 7.0 && p.YDIMS == 8 && p.HDIMS == 8 && p.NBR_ADJ == 1.0 && p.ERR_ADJ == 1.0e-6 && p.DEF_T0 == 1.0e6 && p.DEF_K0 == 30000, "ANNEAL_DEFAULT");
  CHECK(SimAnnParams::ANNEAL_LOOP().DEF_K0 == 35000 && SimAnnParams::ANNEAL_STRICT().YDIMS == 2 && SimAnnParams::ANNEAL_ONLINE().HDIMS == 3, "presets");
  std::vector<SearchVariable> v = axisAngleEigenVariables(2);
  CHECK(v.size() == 8 && v[2].mName == "Tz" && v[2].mMaxVal == 350 && v[2].mDefaultValue == 350 && v[4].mCircular && !v[3].mCircular, "variables");
  CHECK(std::fabs(v[3].mMaxJump - M_PI / 5) < 1e-15 && v[6].mMinVal == -4 && v[6].mMaxJump == 2 && v[6].mName == "EG 0", "variables");
  transf7 ref = {{1, 0, 0, 0}, {0, 0, 0}};
  // distance = max(|dt| / maxRadius, 0.5 |angle| / pi)
  transf7 a = {{1, 0, 0, 0}, {0, 0, 0}}, b = {{0, 1, 0, 0}, {30, 0, 0}}, c = {{1, 0, 0, 0}, {30, 40, 0}};
  CHECK(std::fabs(stateDistance(a, b, 100.0) - 0.5) < 1e-12, "%g", stateDistance(a, b, 100.0));
  CHECK(std::fabs(stateDistance(a, c, 100.0) - 0.5) < 1e-12, "%g", stateDistance(a, c, 100.0));
  // pose of a state: translation composed with the reference frame (90 deg about z, shifted)
  double cz = std::sqrt(0.5);
  transf7 ref2 = {{cz, 0, 0, cz}, {10, 0, 0}};
  transf7 ps = statePose(st(1, 2, 3, 0.f, 0), ref2);
  CHECK(std::fabs(ps.t[0] - 8) < 1e-9 && std::fabs(ps.t[1] - 1) < 1e-9 && std::fabs(ps.t[2] - 3) < 1e-9, "%g %g %g", ps.t[0], ps.t[1], ps.t[2]);
  CHECK(std::fabs(ps.q[0] - cz) < 1e-12 && std::fabs(ps.q[3] - cz) < 1e-12, "identity snap for tiny alpha");
  // best list rules (same sequence as tests/test_simann_port.py::test_best_list_rules)
  BestList bl(ref, 100.0, 3);
  CHECK(bl.offer(st(0, 0, 0, 1, 10.0)), "first");
  CHECK(!bl.offer(st(5, 0, 0, 1, 11.0)), "within 0.2 of a better state");
  CHECK(bl.offer(st(5, 0, 0, 1, 9.0)) && bl.list().size() == 1 && bl.list()[0].energy == 9.0, "replaces the worse neighbour");
  CHECK(bl.offer(st(50, 0, 0, 1, 12.0)) && bl.offer(st(100, 0, 0, 1, 8.0)), "room");
  CHECK(bl.list()[0].energy == 8.0 && bl.list()[1].energy == 9.0 && bl.list()[2].energy == 12.0, "sorted");
  CHECK(!bl.offer(st(-80, 0, 0, 1, 13.0)), "full list: must beat the worst");
  CHECK(bl.offer(st(-80, 0, 0, 1, 10.0)) && bl.list().size() == 3 && bl.list()[2].energy == 10.0, "trim");
  CHECK(bl.worstEnergy() == 10.0 && BestList(ref).worstEnergy() == 1.0e5, "worstEnergy");
  // state files: the reference's "%d " / "%f " format
  std::vector<PlannerState> states;
  std::mt19937 rng(3);
  std::uniform_real_distribution<float> u(-3.f, 3.f);
  for (int i = 0; i < 40; i++) { PlannerState s; s.energy = 0; s.itNumber = 0; for (int k = 0; k < 8; k++) s.vars.push_back(u(rng)); states.push_back(s); }
  FILE *fp = tmpfile();
  CHECK(writeStatesToFile(fp, states, 2), "write");
  rewind(fp);
  char first[256];
  CHECK(fgets(first, sizeof first, fp) != NULL, "read back");
  char expect[256];
  snprintf(expect, sizeof expect, "4 %f %f \n", (double)states[0].vars[6], (double)states[0].vars[7]);
  CHECK(strcmp(first, expect) == 0, "'%s' vs '%s'", first, expect);
  rewind(fp);
  std::vector<PlannerState> back;
  CHECK(readStatesFromFile(fp, 2, &back) && back.size() == 40, "%zu states", back.size());
  for (size_t i = 0; i < back.size(); i++) for (int k = 0; k < 8; k++) CHECK(std::fabs(back[i].vars[k] - states[i].vars[k]) < 1e-6, "state %zu var %d", i, k);
  fclose(fp);
}

// ---- scene dump reader ----
struct Scene {
  b2c_ctx *ctx;
  int robot, object, neg, njoints;
  transf7 ref;
  std::vector<DofDesc> dofs;
  std::vector<int> jointDof;
  std::vector<double> jointRatio;
};

static bool loadScene(const char *path, Scene *S, int device = 0)
{
  std::ifstream in(path);
  if (!in) return false;
  if (b2c_create(device, &S->ctx) != B2C_OK) return false;
  int nb; in >> nb;
  std::vector<int> ids(nb);
  for (int i = 0; i < nb; i++) {
    double q[4], t[3]; std::string file;
    in >> q[0] >> q[1] >> q[2] >> q[3] >> t[0] >> t[1] >> t[2] >> file;
    FILE *f = fopen(file.c_str(), "rb");
    if (!f) return false;
    int n = 0; if (fread(&n, 4, 1, f) != 1) return false;
    std::vector<float> tri(9 * (size_t)n);
    if (fread(tri.data(), 4, tri.size(), f) != tri.size()) return false;
    fclose(f);
    int mesh;
    if (b2c_mesh_create(S->ctx, tri.data(), n, &mesh) != B2C_OK || b2c_body_create(S->ctx, mesh, &ids[i]) != B2C_OK) return false;
    b2c_body_set_transform(S->ctx, ids[i], q, t);
  }
  int nd; in >> nd;
  for (int i = 0; i < nd; i++) { int a, b; in >> a >> b; b2c_pair_set_active(S->ctx, ids[a], ids[b], 0); }
  b2c_robot_desc d;
  memset(&d, 0, sizeof d);
  int base, mount; in >> base >> mount >> d.ndof;
  d.base_body = ids[base]; d.mount_body = mount >= 0 ? ids[mount] : -1;
  std::vector<double> dmin(d.ndof), dmax(d.ndof);
  for (int i = 0; i < d.ndof; i++) in >> dmin[i];
  for (int i = 0; i < d.ndof; i++) in >> dmax[i];
  d.dof_min = dmin.data(); d.dof_max = dmax.data();
  in >> d.nchains;
  std::vector<b2c_chain_desc> chains(d.nchains);
  for (auto &c : chains) in >> c.tran_q[0] >> c.tran_q[1] >> c.tran_q[2] >> c.tran_q[3] >> c.tran_t[0] >> c.tran_t[1] >> c.tran_t[2] >> c.first_joint >> c.njoints >> c.first_link >> c.nlinks;
  d.chains = chains.data();
  in >> d.njoints;
  std::vector<b2c_joint_desc> joints(d.njoints);
  for (auto &j : joints) in >> j.dof >> j.revolute >> j.ratio >> j.theta >> j.d >> j.a >> j.alpha;
  d.joints = joints.data();
  in >> d.nlinks;
  std::vector<int> lb(d.nlinks), ll(d.nlinks);
  for (auto &x : lb) { int b; in >> b; x = ids[b]; }
  for (auto &x : ll) in >> x;
  d.link_body = lb.data(); d.link_last_joint = ll.data();
  d.parent_robot = -1; d.parent_offset_q[0] = 1;
  in >> d.approach_q[0] >> d.approach_q[1] >> d.approach_q[2] >> d.approach_q[3] >> d.approach_t[0] >> d.approach_t[1] >> d.approach_t[2];
  in >> d.neg;
  std::vector<double> eg((size_t)d.neg * d.ndof), eo(d.ndof), en(d.ndof);
  for (auto &x : eg) in >> x;
  for (auto &x : eo) in >> x;
  for (auto &x : en) in >> x;
  d.eg = eg.data(); d.eg_origin = eo.data(); d.eg_norm = en.data();
  in >> d.nvcontacts;
  std::vector<b2c_vcontact_desc> vcs(d.nvcontacts);
  for (auto &v : vcs) { int b; in >> b >> v.loc[0] >> v.loc[1] >> v.loc[2] >> v.normal[0] >> v.normal[1] >> v.normal[2]; v.body = ids[b]; }
  d.vcontacts = vcs.data();
  int obj; in >> obj;
  in >> S->ref.q[0] >> S->ref.q[1] >> S->ref.q[2] >> S->ref.q[3] >> S->ref.t[0] >> S->ref.t[1] >> S->ref.t[2];
  S->dofs.resize(d.ndof);
  for (int i = 0; i < d.ndof; i++) { S->dofs[i].min = dmin[i]; S->dofs[i].max = dmax[i]; in >> S->dofs[i].defaultVelocity; }
  for (int i = 0; i < d.ndof; i++) { int b; in >> b; S->dofs[i].breakAway = b != 0; }
  S->njoints = d.njoints;
  for (auto &j : joints) { S->jointDof.push_back(j.dof); S->jointRatio.push_back(j.ratio); }
  if (!in) return false;
  S->object = ids[obj]; S->neg = d.neg;
  return b2c_robot_define(S->ctx, &d, &S->robot) == B2C_OK;
}

static void gpuChecks(const char *scenePath)
{
  Scene S;
  if (!loadScene(scenePath, &S)) { printf("FAIL cannot load the scene (%s)\n", S.ctx ? b2c_last_error(S.ctx) : "no device"); g_fail++; return; }
  // (declared before the planner: the context must outlive it -- the planner's destructor releases its annealer there)
  struct CtxGuard { b2c_ctx *c; ~CtxGuard() { b2c_destroy(c); } } guard{S.ctx};
  std::vector<SearchVariable> vars = axisAngleEigenVariables(S.neg);
  const int K = 4096, nv = (int)vars.size();
  std::mt19937_64 rng(5);
  std::vector<float> init((size_t)K * nv);
  for (int c = 0; c < K; c++) for (int i = 0; i < nv; i++) init[(size_t)c * nv + i] = (float)std::uniform_real_distribution<double>(vars[i].mMinVal, vars[i].mMaxVal)(rng);
  BatchedSimAnnPlanner planner(S.ctx, S.robot, S.object, S.ref, vars, SimAnnParams::ANNEAL_DEFAULT(), init, 1234);
  CHECK(planner.valid(), "%s", planner.error().c_str());
  CHECK(planner.getCurrentStep() == 30000, "starts at DEF_K0");
  for (int it = 0; it < 6; it++) CHECK(planner.run(25), "%s", planner.error().c_str());
  CHECK(planner.getCurrentStep() == 30150, "step counter %d", planner.getCurrentStep());
  const std::vector<PlannerState> &best = planner.getBestList();
  CHECK(!best.empty() && (int)best.size() <= BEST_LIST_SIZE, "%zu solutions", best.size());
  for (size_t i = 1; i < best.size(); i++) CHECK(best[i - 1].energy <= best[i].energy, "sorted");
  for (size_t i = 0; i < best.size(); i++) for (size_t j = i + 1; j < best.size(); j++)
    CHECK(stateDistance(statePose(best[i], S.ref), statePose(best[j], S.ref), 200.0) >= 0.2, "solutions %zu and %zu are not distinct", i, j);
  // every listed solution is a real legal state with that energy (re-evaluated through ListPlanner's batch call)
  std::vector<unsigned char> legal; std::vector<float> energy; std::vector<PlannerState> lbest;
  CHECK(listPlannerRun(S.ctx, S.robot, S.object, S.ref, best, false, &legal, &energy, &lbest, S.neg), "%s", b2c_last_error(S.ctx));
  for (size_t i = 0; i < best.size(); i++) CHECK(legal[i] && std::fabs(energy[i] - best[i].energy) <= 1e-5 * std::fabs(best[i].energy), "solution %zu: legal %d energy %g vs %g", i, (int)legal[i], energy[i], best[i].energy);
  CHECK(lbest.size() == best.size(), "ListPlanner keeps all %zu legal inputs", best.size());
  // ListPlanner on random states: the list is the 20 lowest legal energies in input order of ties
  std::vector<PlannerState> rnd;
  for (int c = 0; c < 2000; c++) { PlannerState s; s.energy = 0; s.itNumber = 0; s.vars.assign(init.begin() + (size_t)c * nv, init.begin() + (size_t)(c + 1) * nv); rnd.push_back(s); }
  CHECK(listPlannerRun(S.ctx, S.robot, S.object, S.ref, rnd, false, &legal, &energy, &lbest, S.neg), "%s", b2c_last_error(S.ctx));
  std::vector<float> le;
  for (size_t i = 0; i < rnd.size(); i++) if (legal[i]) le.push_back(energy[i]);
  std::sort(le.begin(), le.end());
  CHECK(lbest.size() == std::min<size_t>(BEST_LIST_SIZE, le.size()) && !le.empty(), "%zu listed of %zu legal", lbest.size(), le.size());
  for (size_t i = 0; i < lbest.size(); i++) CHECK(lbest[i].energy == le[i], "rank %zu", i);
  // TimeTester, batched: perturbations of one legal state, legal + illegal = evaluated, deterministic for a seed
  {
    size_t pick = 0;
    while (pick < rnd.size() && !legal[pick]) pick++;
    CHECK(pick < rnd.size(), "a legal random state exists");
    BatchedTimeTester tt(S.ctx, S.robot, S.object, S.ref, vars, rnd[pick], S.neg, 4096), tt2(S.ctx, S.robot, S.object, S.ref, vars, rnd[pick], S.neg, 4096);
    for (int it = 0; it < 3; it++) CHECK(tt.mainLoop() && tt2.mainLoop(), "%s", b2c_last_error(S.ctx));
    CHECK(tt.getCount() + tt.getIllegalCount() == 3 * 4096 && tt.getLoops() == 3, "%lld + %lld", tt.getCount(), tt.getIllegalCount());
    CHECK(tt.getCount() == tt2.getCount() && tt.getCount() > 0, "deterministic: %lld vs %lld legal", tt.getCount(), tt2.getCount());
    printf("BatchedTimeTester: %lld legal, %lld illegal\n", tt.getCount(), tt.getIllegalCount());
  }
  // Hand::autoGrasp over a batch: open hands at the hand poses of the legal random states
  {
    std::vector<transf7> poses; std::vector<double> dofv;
    const int nd = (int)S.dofs.size();
    for (size_t i = 0; i < rnd.size() && poses.size() < 512; i++) {
      if (!legal[i]) continue;
      poses.push_back(statePose(rnd[i], S.ref));
      for (int d = 0; d < nd; d++) dofv.push_back(d == 0 ? 0.5 : 0.0);
    }
    std::vector<double> closed, cj; std::vector<int> st;
    CHECK(autoGraspBatch(S.ctx, S.robot, poses, dofv, S.dofs, 1.0, false, &closed, &st, &cj, S.njoints), "%s", b2c_last_error(S.ctx));
    int moved = 0, errors = 0, coupled = 0;
    std::vector<float> rows; std::vector<int> rowOf;
    for (size_t i = 0; i < poses.size(); i++) {
      if (st[i] & B2C_AG_INTERP_ERROR) { errors++; continue; }
      if (st[i] & B2C_AG_MOVED) moved++;
      bool in = true, same = true;
      for (int d = 0; d < nd; d++) in = in && closed[i * nd + d] >= S.dofs[d].min - 1e-9 && closed[i * nd + d] <= S.dofs[d].max + 1e-9;
      CHECK(in, "pose %zu: DOF outside its limits", i);
      CHECK(closed[i * nd] == (double)(float)0.5, "pose %zu: the spread DOF (velocity 0) moved", i);
      for (int j = 0; j < S.njoints; j++) same = same && std::fabs(cj[i * S.njoints + j] - closed[i * nd + S.jointDof[j]] * S.jointRatio[j]) < 1e-12;
      if (!same) continue;            // a finger in break-away: joints no longer equal ratio x DOF
      coupled++;
      rowOf.push_back((int)i);
      for (int k = 0; k < 4; k++) rows.push_back((float)poses[i].q[k]);
      for (int k = 0; k < 3; k++) rows.push_back((float)poses[i].t[k]);
      for (int d = 0; d < nd; d++) rows.push_back((float)closed[i * nd + d]);
    }
    CHECK(poses.size() >= 100 && moved + errors == (int)poses.size() && moved > 50, "%zu poses, %d moved, %d errors", poses.size(), moved, errors);
    // the closed posture is collision-free (the bisection stops 0 < d < THRESHOLD / 2 short of touching)
    std::vector<unsigned char> lg(rowOf.size()); std::vector<float> en(rowOf.size());
    b2c_eval_args a; memset(&a, 0, sizeof a);
    a.robot = S.robot; a.object = S.object; a.npose = (int)rowOf.size(); a.input_mode = B2C_INPUT_RAW; a.states = rows.data(); a.stride = 7 + nd;
    a.ref_q[0] = 1; a.flags = B2C_EVAL_PRESET; a.legal = lg.data(); a.energy = en.data();
    CHECK(coupled > 20 && b2c_eval_batch(S.ctx, &a) == B2C_OK, "%d coupled postures; %s", coupled, b2c_last_error(S.ctx));
    int nlegal = 0; for (unsigned char v : lg) nlegal += v;
    CHECK(nlegal >= (int)(0.99 * lg.size()), "%d of %zu closed postures are collision-free", nlegal, lg.size());
    printf("autoGraspBatch: %zu poses, %d moved, %d start in collision, %d coupled, %d legal\n", poses.size(), moved, errors, coupled, nlegal);
  }
}

// One process, several contexts (one per device; on a single-GPU box both on device 0): the chains are split over them,
// trajectories equal the single-context run chain by chain, the best list is fed by the collective-free exchange.
static void multiGpuChecks(const char *scenePath)
{
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  const int world = 2;
  Scene S[2], One;
  for (int r = 0; r < world; r++)
    if (!loadScene(scenePath, &S[r], ndev > 1 ? r : 0)) { printf("FAIL cannot load the scene on rank %d\n", r); g_fail++; return; }
  if (!loadScene(scenePath, &One, 0)) { g_fail++; return; }
  std::vector<SearchVariable> vars = axisAngleEigenVariables(One.neg);
  const int K = 4096, nv = (int)vars.size();
  std::mt19937_64 rng(5);
  std::vector<float> init((size_t)K * nv);
  for (int c = 0; c < K; c++) for (int i = 0; i < nv; i++) init[(size_t)c * nv + i] = (float)std::uniform_real_distribution<double>(vars[i].mMinVal, vars[i].mMaxVal)(rng);
  {
    MultiGpuSimAnnPlanner multi({S[0].ctx, S[1].ctx}, {S[0].robot, S[1].robot}, {S[0].object, S[1].object}, One.ref, vars,
                                SimAnnParams::ANNEAL_DEFAULT(), init, 1234);
    BatchedSimAnnPlanner single(One.ctx, One.robot, One.object, One.ref, vars, SimAnnParams::ANNEAL_DEFAULT(), init, 1234);
    CHECK(multi.valid() && single.valid(), "%s %s", multi.error().c_str(), single.error().c_str());
    for (int it = 0; it < 3; it++) CHECK(multi.run(20), "%s", multi.error().c_str());
    CHECK(single.run(60), "%s", single.error().c_str());
    CHECK(multi.getCurrentStep() == 30060, "step counter %d", multi.getCurrentStep());
    CHECK(multi.stepsSeen()[0] == 61 && multi.stepsSeen()[1] == 61, "rows seen from steps %d / %d", multi.stepsSeen()[0], multi.stepsSeen()[1]);
    // chain by chain the same trajectories: best energies of the two halves == the single run's
    std::vector<float> be(K), half(K / 2);
    CHECK(b2c_anneal_read(One.ctx, single.annealer(), NULL, NULL, NULL, be.data(), NULL, NULL) == B2C_OK, "read");
    int same = 0;
    for (int r = 0; r < world; r++) {
      CHECK(b2c_anneal_read(S[r].ctx, multi.annealer(r), NULL, NULL, NULL, half.data(), NULL, NULL) == B2C_OK, "read rank %d", r);
      for (int c = 0; c < K / 2; c++) same += half[c] == be[(size_t)r * (K / 2) + c];
    }
    CHECK(same == K, "%d of %d chains end with the single-GPU best energy", same, K);
    // the merged best list: sorted, distinct, and its head is the global minimum over all chains
    const std::vector<PlannerState> &best = multi.getBestList();
    CHECK(!best.empty() && (int)best.size() <= BEST_LIST_SIZE, "%zu solutions", best.size());
    for (size_t i = 1; i < best.size(); i++) CHECK(best[i - 1].energy <= best[i].energy, "sorted");
    float gmin = 3.0e38f;
    for (float e : be) gmin = std::min(gmin, e);
    CHECK(!best.empty() && (float)best[0].energy == gmin, "head %g vs global minimum %g", best.empty() ? 0.0 : best[0].energy, gmin);
    printf("MultiGpuSimAnnPlanner: %d contexts on %d device(s), %zu solutions, best %g\n", world, ndev > 1 ? 2 : 1, best.size(), best.empty() ? 0.0 : best[0].energy);
  }
  for (int r = 0; r < world; r++) b2c_destroy(S[r].ctx);
  b2c_destroy(One.ctx);
}

// a fault in the harness or the library prints where it happened (the GPU boxes have no debugger)
static void onFault(int sig)
{
  void *frames[64];
  const int n = backtrace(frames, 64);
  const char msg[] = "b200planner_check: fatal signal, backtrace:\n";
  if (write(2, msg, sizeof msg - 1) < 0) {}
  backtrace_symbols_fd(frames, n, 2);
  (void)sig;
  _exit(70);                 // (an exit code, not a re-raised signal: the GPU pool backs off after killed commands)
}

int main(int argc, char **argv)
{
  signal(SIGSEGV, onFault);
  signal(SIGABRT, onFault);
  if (argc < 2) { fprintf(stderr, "usage: %s cpu | gpu scene.txt\n", argv[0]); return 2; }
  cpuChecks();
  if (!strcmp(argv[1], "gpu")) {
    if (argc < 3) return 2;
    gpuChecks(argv[2]);
    multiGpuChecks(argv[2]);
  }
  printf("%s: %d checks, %d failed\n", g_fail ? "FAILED" : "PASSED", g_checks, g_fail);
  return g_fail ? 1 : 0;
}
