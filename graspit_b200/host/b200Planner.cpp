// See b200Planner.h.  Host C++ only; the geometry and the annealing run in libb2c.so.
#include "b200Planner.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

// the CUDA runtime is reached only through these four calls, so GraspIt! needs no CUDA headers to build this file
extern "C" {
int cudaMalloc(void **p, size_t n);
int cudaFree(void *p);
int cudaMemcpy(void *dst, const void *src, size_t n, int kind);
int cudaStreamSynchronize(void *stream);
}

namespace b200 {

static SimAnnParams mk(double yc, double hc, int yd, int hd, double nbr, double err, double t0, int k0)
{
  SimAnnParams p;
  p.YC = yc; p.HC = hc; p.YDIMS = yd; p.HDIMS = hd; p.NBR_ADJ = nbr; p.ERR_ADJ = err; p.DEF_T0 = t0; p.DEF_K0 = k0;
  return p;
}
SimAnnParams SimAnnParams::ANNEAL_DEFAULT() {return mk(7.0, 7.0, 8, 8, 1.0, 1.0e-6, 1.0e6, 30000);}
SimAnnParams SimAnnParams::ANNEAL_LOOP() {return mk(7.0, 7.0, 8, 8, 1.0, 1.0e-6, 1.0e6, 35000);}
SimAnnParams SimAnnParams::ANNEAL_MODIFIED() {return mk(2.38, 2.38, 8, 8, 1.0e-3, 1.0e-1, 1.0e3, 0);}
SimAnnParams SimAnnParams::ANNEAL_STRICT() {return mk(0.72, 0.22, 2, 2, 1.0, 1.0, 10.0, 0);}
SimAnnParams SimAnnParams::ANNEAL_ONLINE() {return mk(0.24, 0.54, 2, 3, 1.0, 1.0e-2, 1.0e1, 100);}

double SearchVariable::getRange() const {return std::fabs(mMaxVal - mMinVal);}

static SearchVariable var(const char *name, double mn, double mx, double def, double jump, bool circ = false)
{
  SearchVariable v;
  v.mName = name; v.mMinVal = mn; v.mMaxVal = mx; v.mDefaultValue = def; v.mMaxJump = jump; v.mCircular = circ;
  return v;
}

std::vector<SearchVariable> axisAngleEigenVariables(int nEigenGrasps)
{
  std::vector<SearchVariable> v;
  v.push_back(var("Tx", -250, 250, 0, 150));
  v.push_back(var("Ty", -250, 250, 0, 150));
  v.push_back(var("Tz", -250, 350, 350, 150));
  v.push_back(var("theta", 0, M_PI, 0, M_PI / 5));
  v.push_back(var("phi", -M_PI, M_PI, 0, M_PI / 2, true));
  v.push_back(var("alpha", 0, M_PI, M_PI / 2, M_PI / 2));
  for (int i = 0; i < nEigenGrasps; i++) {
    char name[32];
    snprintf(name, sizeof name, "EG %d", i);
    // limits not predefined in the eigengrasp file: -4..4, jump = range / 4 (searchStateImpl.cpp:66-76)
    v.push_back(var(name, -4.0, 4.0, -4.0, 2.0));
  }
  return v;
}

static void quatMul(const double a[4], const double b[4], double r[4])
{
  r[0] = a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3];
  r[1] = a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2];
  r[2] = a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3];
  r[3] = a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1];
}

transf7 statePose(const PlannerState &s, const transf7 &ref)
{
  const double tx = s.vars[0], ty = s.vars[1], tz = s.vars[2], th = s.vars[3], ph = s.vars[4], al = s.vars[5];
  double ax[3] = {std::sin(th) * std::cos(ph), std::sin(th) * std::sin(ph), std::cos(th)};
  double q[4] = {1, 0, 0, 0};
  if (al * al >= 1.0e-6) {                       // transf::AXIS_ANGLE_ROTATION snaps tiny angles to identity
    double n = std::sqrt(ax[0] * ax[0] + ax[1] * ax[1] + ax[2] * ax[2]);
    double sh = std::sin(al / 2);
    q[0] = std::cos(al / 2); q[1] = sh * ax[0] / n; q[2] = sh * ax[1] / n; q[3] = sh * ax[2] / n;
  }
  transf7 out;
  quatMul(ref.q, q, out.q);
  // ref.rot * t + ref.t   (v + w 2(u x v) + u x 2(u x v))
  const double w = ref.q[0], u[3] = {ref.q[1], ref.q[2], ref.q[3]}, v[3] = {tx, ty, tz};
  double uv[3] = {2 * (u[1] * v[2] - u[2] * v[1]), 2 * (u[2] * v[0] - u[0] * v[2]), 2 * (u[0] * v[1] - u[1] * v[0])};
  double c[3] = {u[1] * uv[2] - u[2] * uv[1], u[2] * uv[0] - u[0] * uv[2], u[0] * uv[1] - u[1] * uv[0]};
  for (int k = 0; k < 3; k++) {out.t[k] = ref.t[k] + v[k] + w * uv[k] + c[k];}
  return out;
}

double stateDistance(const transf7 &a, const transf7 &b, double maxRadius)
{
  double d = std::sqrt((a.t[0] - b.t[0]) * (a.t[0] - b.t[0]) + (a.t[1] - b.t[1]) * (a.t[1] - b.t[1]) +
                       (a.t[2] - b.t[2]) * (a.t[2] - b.t[2])) / maxRadius;
  double binv[4] = {b.q[0], -b.q[1], -b.q[2], -b.q[3]}, q[4];
  quatMul(a.q, binv, q);
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  double w = q[0] / n, vn = std::sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]) / n;
  if (w < 0) {w = -w;}                             // Eigen::AngleAxisd(Quaternion): angle in [0, pi]
  double angle = 2.0 * std::atan2(vn, w);
  return std::max(d, 0.5 * std::fabs(angle) / M_PI);
}

BestList::BestList(const transf7 &refTran, double maxRadius, int size, double distance) :
  mRefTran(refTran), mMaxRadius(maxRadius), mDistance(distance), mSize(size) {}

double BestList::worstEnergy() const
{
  return (int)mList.size() < mSize ? 1.0e5 : mList.back().energy;
}

bool BestList::addToListOfUniqueSolutions(const PlannerState &s)
{
  transf7 ps = statePose(s, mRefTran);
  bool add = true;
  size_t i = 0;
  while (i < mList.size()) {
    double d = stateDistance(ps, statePose(mList[i], mRefTran), mMaxRadius);
    if (std::fabs(d) < mDistance) {
      if (s.energy < mList[i].energy) {
        mList.erase(mList.begin() + i);            // new state is better; remove old one from list
      } else {
        add = false;                               // old state is better; we don't want to add the new one
        break;
      }
    } else {
      i++;
    }
  }
  if (add) {mList.push_back(s);}
  return add;
}

static bool compareStates(const PlannerState &a, const PlannerState &b) {return a.energy < b.energy;}

bool BestList::offer(const PlannerState &s)
{
  if (!(s.energy < worstEnergy())) {return false;}
  if (!addToListOfUniqueSolutions(s)) {return false;}
  std::stable_sort(mList.begin(), mList.end(), compareStates);
  while ((int)mList.size() > mSize) {mList.pop_back();}
  return true;
}

bool writeStatesToFile(FILE *fp, const std::vector<PlannerState> &states, int nEG)
{
  for (size_t i = 0; i < states.size(); i++) {
    const std::vector<float> &v = states[i].vars;
    if ((int)v.size() < 6 + nEG) {return false;}
    fprintf(fp, "%d ", 4);                                          // POSE_EIGEN
    for (int k = 0; k < nEG; k++) {fprintf(fp, "%f ", (double)v[6 + k]);}
    fprintf(fp, "\n");
    fprintf(fp, "%d ", 1);                                          // SPACE_AXIS_ANGLE
    for (int k = 0; k < 6; k++) {fprintf(fp, "%f ", (double)v[k]);}
    fprintf(fp, "\n");
  }
  return true;
}

bool readStatesFromFile(FILE *fp, int nEG, std::vector<PlannerState> *states)
{
  char line[10000];
  while (true) {
    int type;
    if (fscanf(fp, "%d", &type) <= 0) {break;}
    if (type != 4) {return false;}                                  // only POSE_EIGEN + SPACE_AXIS_ANGLE here
    if (!fgets(line, sizeof line, fp)) {return false;}
    PlannerState s;
    s.vars.assign(6 + nEG, 0.f); s.energy = 0; s.itNumber = 0;
    char *p = line;
    for (int k = 0; k < nEG; k++) {
      char *e; double v = strtod(p, &e);
      if (e == p) {return false;}                                   // "Line to short to read all state variables"
      s.vars[6 + k] = (float)v; p = e;
    }
    if (fscanf(fp, "%d", &type) <= 0 || type != 1) {return false;}
    if (!fgets(line, sizeof line, fp)) {return false;}
    p = line;
    for (int k = 0; k < 6; k++) {
      char *e; double v = strtod(p, &e);
      if (e == p) {return false;}
      s.vars[k] = (float)v; p = e;
    }
    states->push_back(s);
  }
  return true;
}

BatchedSimAnnPlanner::BatchedSimAnnPlanner(b2c_ctx *ctx, int robot, int object, const transf7 &refTran,
                                           const std::vector<SearchVariable> &variables, const SimAnnParams &params,
                                           const std::vector<float> &initialStates, unsigned long long seed,
                                           int chainOffset, bool contactLive, double maxRadius) :
  mCtx(ctx), mAnnealer(-1), mNumChains(0), mNumVars((int)variables.size()), mCurrentStep(params.DEF_K0),
  mRowsDevice(NULL), mBestList(refTran, maxRadius)
{
  if (mNumVars <= 0 || initialStates.size() % mNumVars != 0) {mError = "initial states do not match the variable count"; return;}
  mNumChains = (int)(initialStates.size() / mNumVars);
  std::vector<double> vmin(mNumVars), vmax(mNumVars), vjump(mNumVars);
  std::vector<unsigned char> circ(mNumVars);
  for (int i = 0; i < mNumVars; i++) {
    vmin[i] = variables[i].mMinVal; vmax[i] = variables[i].mMaxVal; vjump[i] = variables[i].mMaxJump;
    circ[i] = variables[i].mCircular ? 1 : 0;
  }
  b2c_anneal_desc d;
  memset(&d, 0, sizeof d);
  d.robot = robot; d.object = object; d.nchain = mNumChains; d.nvar = mNumVars;
  d.vmin = vmin.data(); d.vmax = vmax.data(); d.vjump = vjump.data(); d.circular = circ.data();
  for (int k = 0; k < 4; k++) {d.ref_q[k] = refTran.q[k];}
  for (int k = 0; k < 3; k++) {d.ref_t[k] = refTran.t[k];}
  d.YC = params.YC; d.HC = params.HC; d.NBR_ADJ = params.NBR_ADJ; d.ERR_ADJ = params.ERR_ADJ; d.T0 = params.DEF_T0;
  d.YDIMS = params.YDIMS; d.HDIMS = params.HDIMS; d.K0 = params.DEF_K0;
  d.seed = seed; d.chain_offset = chainOffset; d.flags = contactLive ? B2C_EVAL_LIVE : B2C_EVAL_PRESET;
  if (b2c_anneal_create(ctx, &d, initialStates.data(), &mAnnealer) != B2C_OK) {
    mError = b2c_last_error(ctx); mAnnealer = -1; return;
  }
  mRows.assign((size_t)BEST_LIST_SIZE * (mNumVars + 1), 0.f);
  if (cudaMalloc(&mRowsDevice, mRows.size() * sizeof(float)) != 0) {mError = "cudaMalloc failed"; mRowsDevice = NULL;}
}

BatchedSimAnnPlanner::~BatchedSimAnnPlanner()
{
  if (mAnnealer >= 0) {b2c_anneal_destroy(mCtx, mAnnealer);}
  if (mRowsDevice) {cudaFree(mRowsDevice);}
}

void BatchedSimAnnPlanner::mergeRows(const float *rows, int n)
{
  // candidates in ascending energy: the order a single planner would have met the better ones
  std::vector<int> order(n);
  for (int i = 0; i < n; i++) {order[i] = i;}
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
    return rows[(size_t)a * (mNumVars + 1) + mNumVars] < rows[(size_t)b * (mNumVars + 1) + mNumVars];
  });
  for (int k = 0; k < n; k++) {
    const float *r = rows + (size_t)order[k] * (mNumVars + 1);
    if (!(r[mNumVars] < 1.0e7f)) {continue;}          // empty slot / chain without a legal state yet
    PlannerState s;
    s.vars.assign(r, r + mNumVars); s.energy = r[mNumVars]; s.itNumber = mCurrentStep;
    mBestList.offer(s);
  }
}

bool BatchedSimAnnPlanner::run(int nSteps)
{
  if (mAnnealer < 0 || !mRowsDevice) {return false;}
  if (b2c_anneal_step(mCtx, mAnnealer, nSteps) != B2C_OK) {mError = b2c_last_error(mCtx); return false;}
  if (b2c_anneal_best_rows(mCtx, mAnnealer, BEST_LIST_SIZE, mRowsDevice) != B2C_OK) {mError = b2c_last_error(mCtx); return false;}
  cudaStreamSynchronize(b2c_stream(mCtx));
  if (cudaMemcpy(mRows.data(), mRowsDevice, mRows.size() * sizeof(float), 2 /* cudaMemcpyDeviceToHost */) != 0) {
    mError = "cudaMemcpy failed"; return false;
  }
  mCurrentStep += nSteps;
  mergeRows(mRows.data(), BEST_LIST_SIZE);
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------
static bool makeAnnealDesc(b2c_anneal_desc *d, int robot, int object, int nchain, const transf7 &refTran,
                           const std::vector<SearchVariable> &variables, const SimAnnParams &params, unsigned long long seed,
                           int chainOffset, bool contactLive, std::vector<double> *vmin, std::vector<double> *vmax,
                           std::vector<double> *vjump, std::vector<unsigned char> *circ)
{
  const int nv = (int)variables.size();
  vmin->resize(nv); vmax->resize(nv); vjump->resize(nv); circ->resize(nv);
  for (int i = 0; i < nv; i++) {
    (*vmin)[i] = variables[i].mMinVal; (*vmax)[i] = variables[i].mMaxVal; (*vjump)[i] = variables[i].mMaxJump;
    (*circ)[i] = variables[i].mCircular ? 1 : 0;
  }
  memset(d, 0, sizeof *d);
  d->robot = robot; d->object = object; d->nchain = nchain; d->nvar = nv;
  d->vmin = vmin->data(); d->vmax = vmax->data(); d->vjump = vjump->data(); d->circular = circ->data();
  for (int k = 0; k < 4; k++) {d->ref_q[k] = refTran.q[k];}
  for (int k = 0; k < 3; k++) {d->ref_t[k] = refTran.t[k];}
  d->YC = params.YC; d->HC = params.HC; d->NBR_ADJ = params.NBR_ADJ; d->ERR_ADJ = params.ERR_ADJ; d->T0 = params.DEF_T0;
  d->YDIMS = params.YDIMS; d->HDIMS = params.HDIMS; d->K0 = params.DEF_K0;
  d->seed = seed; d->chain_offset = chainOffset; d->flags = contactLive ? B2C_EVAL_LIVE : B2C_EVAL_PRESET;
  return true;
}

MultiGpuSimAnnPlanner::MultiGpuSimAnnPlanner(const std::vector<b2c_ctx *> &contexts, const std::vector<int> &robots,
                                             const std::vector<int> &objects, const transf7 &refTran,
                                             const std::vector<SearchVariable> &variables, const SimAnnParams &params,
                                             const std::vector<float> &initialStates, unsigned long long seed, bool contactLive,
                                             double maxRadius)
  : mCtx(contexts), mNumVars((int)variables.size()), mCurrentStep(params.DEF_K0), mChainsPerRank(0),
    mBestList(refTran, maxRadius), mValid(false)
{
  const int world = (int)contexts.size();
  if (world < 1 || (int)robots.size() != world || (int)objects.size() != world || mNumVars <= 0) {mError = "one robot and object id per context"; return;}
  const size_t total = initialStates.size() / mNumVars;
  if (total == 0 || total % world != 0) {mError = "the chain count must be a multiple of the context count"; return;}
  mChainsPerRank = (int)(total / world);
  mAnnealers.assign(world, -1);
  std::vector<void *> rings(world, NULL);
  std::vector<int> devices(world, 0);
  for (int r = 0; r < world; r++) {
    b2c_anneal_desc d;
    std::vector<double> vmin, vmax, vjump; std::vector<unsigned char> circ;
    makeAnnealDesc(&d, robots[r], objects[r], mChainsPerRank, refTran, variables, params, seed, r * mChainsPerRank, contactLive,
                   &vmin, &vmax, &vjump, &circ);
    if (b2c_anneal_create(mCtx[r], &d, initialStates.data() + (size_t)r * mChainsPerRank * mNumVars, &mAnnealers[r]) != B2C_OK ||
        b2c_exchange_create(mCtx[r], r, world, 4, BEST_LIST_SIZE, mNumVars + 1, NULL, &rings[r]) != B2C_OK) {
      mError = b2c_last_error(mCtx[r]); return;
    }
    devices[r] = b2c_device(mCtx[r]);
  }
  for (int r = 0; r < world; r++) {
    if (b2c_exchange_connect(mCtx[r], NULL, rings.data(), devices.data()) != B2C_OK) {mError = b2c_last_error(mCtx[r]); return;}
  }
  mRows.assign((size_t)world * BEST_LIST_SIZE * (mNumVars + 1), 0.f);
  mSteps.assign(world, 0);
  mValid = true;
}

MultiGpuSimAnnPlanner::~MultiGpuSimAnnPlanner()
{
  for (size_t r = 0; r < mCtx.size(); r++) {
    if (r < mAnnealers.size() && mAnnealers[r] >= 0) {b2c_anneal_destroy(mCtx[r], mAnnealers[r]);}
    b2c_exchange_destroy(mCtx[r]);
  }
}

bool MultiGpuSimAnnPlanner::collectAndMerge()
{
  if (b2c_exchange_collect(mCtx[0], mRows.data(), mSteps.data(), 0) != B2C_OK) {mError = b2c_last_error(mCtx[0]); return false;}
  const int n = (int)mCtx.size() * BEST_LIST_SIZE, w = mNumVars + 1;
  std::vector<int> order(n);
  for (int i = 0; i < n; i++) {order[i] = i;}
  std::stable_sort(order.begin(), order.end(), [&](int a, int b) {return mRows[(size_t)a * w + mNumVars] < mRows[(size_t)b * w + mNumVars];});
  for (int k = 0; k < n; k++) {
    const float *r = &mRows[(size_t)order[k] * w];
    if (!(r[mNumVars] < 1.0e7f) || !(r[mNumVars] < mBestList.worstEnergy())) {continue;}
    PlannerState s;
    s.vars.assign(r, r + mNumVars); s.energy = r[mNumVars]; s.itNumber = mCurrentStep;
    mBestList.offer(s);
  }
  return true;
}

bool MultiGpuSimAnnPlanner::run(int nSteps)
{
  if (!mValid) {return false;}
  for (int s = 0; s < nSteps; s++) {
    // enqueue on every device (asynchronous: the GPUs run the step concurrently), each publishes into all rings
    for (size_t r = 0; r < mCtx.size(); r++) {
      if (b2c_anneal_step(mCtx[r], mAnnealers[r], 1) != B2C_OK || b2c_anneal_publish(mCtx[r], mAnnealers[r], NULL) != B2C_OK) {
        mError = b2c_last_error(mCtx[r]); return false;
      }
    }
    mCurrentStep++;
    // merge what has landed on context 0 (waits for ITS stream only; peers that are behind are picked up next step)
    if (!collectAndMerge()) {return false;}
  }
  // the end of a run is a synchronisation point: every GPU's last rows are in
  for (size_t r = 0; r < mCtx.size(); r++) {
    if (b2c_check_overflow(mCtx[r]) != B2C_OK) {mError = b2c_last_error(mCtx[r]); return false;}
  }
  return collectAndMerge();
}

bool listPlannerRun(b2c_ctx *ctx, int robot, int object, const transf7 &refTran, const std::vector<PlannerState> &states,
                    bool contactLive, std::vector<unsigned char> *legal, std::vector<float> *energy,
                    std::vector<PlannerState> *bestList, int nEG)
{
  const int n = (int)states.size(), nv = 6 + nEG;
  legal->assign(n, 0); energy->assign(n, 0.f); bestList->clear();
  if (n == 0) {return true;}
  std::vector<float> x((size_t)n * nv);
  for (int i = 0; i < n; i++) {
    if ((int)states[i].vars.size() < nv) {return false;}
    for (int k = 0; k < nv; k++) {x[(size_t)i * nv + k] = states[i].vars[k];}
  }
  b2c_eval_args a;
  memset(&a, 0, sizeof a);
  a.robot = robot; a.object = object; a.npose = n; a.input_mode = B2C_INPUT_AA_EIGEN; a.states = x.data(); a.stride = nv;
  for (int k = 0; k < 4; k++) {a.ref_q[k] = refTran.q[k];}
  for (int k = 0; k < 3; k++) {a.ref_t[k] = refTran.t[k];}
  a.flags = contactLive ? B2C_EVAL_LIVE : B2C_EVAL_PRESET;
  a.legal = legal->data(); a.energy = energy->data();
  if (b2c_eval_batch(ctx, &a) != B2C_OK) {return false;}
  for (int i = 0; i < n; i++) {
    if (!(*legal)[i]) {continue;}
    double worst = (int)bestList->size() < BEST_LIST_SIZE ? 1.0e5 : bestList->back().energy;
    if ((*energy)[i] < worst) {
      PlannerState s = states[i];
      s.energy = (*energy)[i]; s.itNumber = i;
      bestList->push_back(s);
      std::stable_sort(bestList->begin(), bestList->end(), compareStates);
      while ((int)bestList->size() > BEST_LIST_SIZE) {bestList->pop_back();}
    }
  }
  return true;
}


BatchedTimeTester::BatchedTimeTester(b2c_ctx *ctx, int robot, int object, const transf7 &refTran,
                                     const std::vector<SearchVariable> &variables, const PlannerState &current, int nEG,
                                     int batch, bool contactLive, unsigned long long seed)
  : mCtx(ctx), mRobot(robot), mObject(object), mNumEG(nEG), mBatch(batch), mLive(contactLive), mRefTran(refTran),
    mVars(variables), mCurrentState(current), mSeed(seed), mCount(0), mIllegalCount(0), mLoops(0) {}

bool BatchedTimeTester::mainLoop()
{
  const int nv = 6 + mNumEG;
  if (mBatch <= 0 || (int)mCurrentState.vars.size() < nv || (int)mVars.size() < nv) {return false;}
  std::vector<float> x((size_t)mBatch * nv);
  for (int i = 0; i < mBatch; i++) {
    // splitmix64 keyed by (seed, loop, i): r = rand() / RAND_MAX - 0.5 of the reference, one r per perturbed state
    unsigned long long z = mSeed + 0x9e3779b97f4a7c15ULL * (unsigned long long)(mLoops * (long long)mBatch + i + 1);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL; z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL; z ^= z >> 31;
    const double r = (double)(z >> 11) / 9007199254740992.0 - 0.5;
    for (int k = 0; k < nv; k++) {
      double v = mCurrentState.vars[k];
      v += r * mVars[k].mMaxJump * 0.1;            // no variable of this layout is fixed (SearchVariable::isFixed)
      x[(size_t)i * nv + k] = (float)v;
    }
  }
  std::vector<unsigned char> legal(mBatch);
  std::vector<float> energy(mBatch);
  b2c_eval_args a;
  memset(&a, 0, sizeof a);
  a.robot = mRobot; a.object = mObject; a.npose = mBatch; a.input_mode = B2C_INPUT_AA_EIGEN; a.states = x.data(); a.stride = nv;
  for (int k = 0; k < 4; k++) {a.ref_q[k] = mRefTran.q[k];}
  for (int k = 0; k < 3; k++) {a.ref_t[k] = mRefTran.t[k];}
  a.flags = mLive ? B2C_EVAL_LIVE : B2C_EVAL_PRESET;
  a.legal = legal.data(); a.energy = energy.data();
  if (b2c_eval_batch(mCtx, &a) != B2C_OK) {return false;}
  for (int i = 0; i < mBatch; i++) {if (legal[i]) {mCount++;} else {mIllegalCount++;}}
  mLoops++;
  return true;
}

bool autoGraspBatch(b2c_ctx *ctx, int robot, const std::vector<transf7> &handPoses, const std::vector<double> &dofVals,
                    const std::vector<DofDesc> &dofs, double speedFactor, bool stopAtContact,
                    std::vector<double> *closedDofs, std::vector<int> *status, std::vector<double> *closedJoints, int nJoints)
{
  const int n = (int)handPoses.size(), nd = (int)dofs.size();
  if (!ctx || !closedDofs || nd <= 0 || dofVals.size() != (size_t)n * nd) {return false;}
  closedDofs->assign((size_t)n * nd, 0.0);
  std::vector<int> st((size_t)n * 2, 0);
  if (status) {status->assign(n, 0);}
  if (closedJoints) {closedJoints->assign((size_t)n * nJoints, 0.0);}
  if (n == 0) {return true;}
  const int stride = 7 + nd;
  std::vector<float> rows((size_t)n * stride);
  for (int i = 0; i < n; i++) {
    float *r = &rows[(size_t)i * stride];
    for (int k = 0; k < 4; k++) {r[k] = (float)handPoses[i].q[k];}
    for (int k = 0; k < 3; k++) {r[4 + k] = (float)handPoses[i].t[k];}
    for (int d = 0; d < nd; d++) {r[7 + d] = (float)dofVals[(size_t)i * nd + d];}
  }
  // Hand::autoGrasp: target = max or min by the sign of speedFactor * defaultVelocity, step = velocity * speedFactor * 0.01
  std::vector<double> desired(nd), steps(nd);
  std::vector<unsigned char> brk(nd);
  for (int d = 0; d < nd; d++) {
    desired[d] = speedFactor * dofs[d].defaultVelocity >= 0 ? dofs[d].max : dofs[d].min;
    steps[d] = dofs[d].defaultVelocity * speedFactor * 0.01;
    brk[d] = dofs[d].breakAway ? 1 : 0;
  }
  b2c_autograsp_args a;
  memset(&a, 0, sizeof a);
  a.robot = robot; a.npose = n; a.states = rows.data(); a.stride = stride;
  a.desired = desired.data(); a.steps = steps.data(); a.dof_breakaway = brk.data();
  a.stop_at_contact = stopAtContact ? 1 : 0;
  a.dof_out = closedDofs->data(); a.status = st.data();
  if (closedJoints && nJoints > 0) {a.joint_out = closedJoints->data();}
  if (b2c_autograsp_batch(ctx, &a) != B2C_OK) {return false;}
  if (status) {for (int i = 0; i < n; i++) {(*status)[i] = st[2 * (size_t)i];}}
  return true;
}

}  // namespace b200
