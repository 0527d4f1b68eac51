// Batched EGPlanner host side in C++ over the b2c C ABI: the counterpart, inside GraspIt!, of
//   SimAnnParams presets          reference src/EGPlanner/simAnnParams.cpp
//   SearchVariable tables         src/EGPlanner/searchStateImpl.cpp:50-80,147-155 (PositionStateAA, PostureStateEigen)
//   SimAnnPlanner::mainLoop       src/EGPlanner/simAnnPlanner.cpp:111-148   (best list of BEST_LIST_SIZE states)
//   EGPlanner::addToListOfUniqueSolutions / stateDistance   src/EGPlanner/egPlanner.cpp:520-557
//   HandObjectState::distance     src/EGPlanner/searchState.cpp:572-611
//   ListPlanner::mainLoop         src/EGPlanner/listPlanner.cpp:126-170
//   HandObjectState::writeToFile / readFromFile   src/EGPlanner/searchState.cpp:480-513, VariableSet :224-262
// with the annealing itself (K chains) running on the GPU through b2c_anneal_*.  No Qt, no CUDA headers: this file
// and b200Planner.cpp compile with any C++11 compiler against <b2c.h>.  Names follow the reference's.
#ifndef B200_PLANNER_H
#define B200_PLANNER_H

#include <cstdio>
#include <string>
#include <vector>

#include "b2c.h"

namespace b200 {

static const int BEST_LIST_SIZE = 20;             // simAnnPlanner.cpp:36

struct SimAnnParams {
  double YC, HC;
  int YDIMS, HDIMS;
  double NBR_ADJ, ERR_ADJ, DEF_T0;
  int DEF_K0;
  static SimAnnParams ANNEAL_DEFAULT();
  static SimAnnParams ANNEAL_LOOP();
  static SimAnnParams ANNEAL_MODIFIED();
  static SimAnnParams ANNEAL_STRICT();
  static SimAnnParams ANNEAL_ONLINE();
};

struct SearchVariable {
  std::string mName;
  double mMinVal, mMaxVal, mDefaultValue, mMaxJump;
  bool mCircular;
  double getRange() const;
};

//! SPACE_AXIS_ANGLE position + POSE_EIGEN posture, the planner dialog's default layout (6 + nEigenGrasps variables)
std::vector<SearchVariable> axisAngleEigenVariables(int nEigenGrasps);

struct transf7 {           // q wxyz, t
  double q[4], t[3];
};

//! one planner state in the SPACE_AXIS_ANGLE + POSE_EIGEN layout
struct PlannerState {
  std::vector<float> vars;
  double energy;
  int itNumber;
};

//! total hand transform % approach transform of a state: mRefTran % T(tx,ty,tz) % AxisAngle(alpha, axis(theta, phi))
transf7 statePose(const PlannerState &s, const transf7 &refTran);
//! HandObjectState::distance on two such poses: max(|dt| / maxRadius, 0.5 |angle| / pi)
double stateDistance(const transf7 &a, const transf7 &b, double maxRadius);

//! mBestList with the SimAnnPlanner rules
class BestList
{
    transf7 mRefTran;
    double mMaxRadius, mDistance;
    int mSize;
    std::vector<PlannerState> mList;
  public:
    BestList(const transf7 &refTran, double maxRadius = 200.0, int size = BEST_LIST_SIZE, double distance = 0.2);
    double worstEnergy() const;                   //!< 1.0e5 while the list has room (simAnnPlanner.cpp:128-130)
    bool addToListOfUniqueSolutions(const PlannerState &s);
    //! the JUMP branch of SimAnnPlanner::mainLoop: offer, sort by energy, trim
    bool offer(const PlannerState &s);
    const std::vector<PlannerState> &list() const {return mList;}
};

//! VariableSet wire format: "<type> v0 v1 ... \n", posture line then position line per state
bool writeStatesToFile(FILE *fp, const std::vector<PlannerState> &states, int nEigenGrasps);
bool readStatesFromFile(FILE *fp, int nEigenGrasps, std::vector<PlannerState> *states);

//! K annealing chains on one GPU + the planner's best list on the host
class BatchedSimAnnPlanner
{
    b2c_ctx *mCtx;
    int mAnnealer, mNumChains, mNumVars, mCurrentStep;
    std::vector<float> mRows;           // host copy of the context's best rows
    void *mRowsDevice;
    BestList mBestList;
    std::string mError;
  public:
    // ctx is borrowed and must outlive the planner (the destructor releases the annealer in it)
    BatchedSimAnnPlanner(b2c_ctx *ctx, int robot, int object, const transf7 &refTran,
                         const std::vector<SearchVariable> &variables, const SimAnnParams &params,
                         const std::vector<float> &initialStates, unsigned long long seed = 1234,
                         int chainOffset = 0, bool contactLive = false, double maxRadius = 200.0);
    ~BatchedSimAnnPlanner();
    bool valid() const {return mAnnealer >= 0;}
    const std::string &error() const {return mError;}
    //! n annealing steps of every chain, then the best-list merge of SimAnnPlanner::mainLoop
    bool run(int nSteps);
    int getCurrentStep() const {return mCurrentStep;}
    const std::vector<PlannerState> &getBestList() const {return mBestList.list();}
    //! this GPU's BEST_LIST_SIZE best rows in device memory (payload of the multi-GPU all-gather), refreshed by run()
    const void *bestRowsDevice() const {return mRowsDevice;}
    //! merges rows [n][nvars + 1] gathered from other GPUs into the best list
    void mergeRows(const float *rows, int n);
    int annealer() const {return mAnnealer;}
};

//! SimAnnPlanner over several GPUs driven from ONE process: one context per device, the chains
//! block-partitioned over them (chain g keeps its Philox stream wherever it runs, so the trajectories equal the
//! single-GPU ones), no data-path collective.  After every step each context publishes its BEST_LIST_SIZE best rows
//! straight into its peers' HBM rings (b2c_anneal_publish: peer access over NVLink, no NCCL, no rendezvous); the host
//! merges what has landed on the first context into mBestList with the reference's rules
//! (simAnnPlanner.cpp:111-148, egPlanner.cpp:527-557).  The contexts must hold the same world (bodies, robot, object).
class MultiGpuSimAnnPlanner
{
    std::vector<b2c_ctx *> mCtx;
    std::vector<int> mAnnealers;
    int mNumVars, mCurrentStep, mChainsPerRank;
    std::vector<float> mRows;           // [world][BEST_LIST_SIZE][nvars + 1] as collected on context 0
    std::vector<int> mSteps;
    BestList mBestList;
    std::string mError;
    bool mValid;
    bool collectAndMerge();
  public:
    //! robots[i] / objects[i]: the ids inside contexts[i]; initialStates: all chains (a multiple of the context count)
    MultiGpuSimAnnPlanner(const std::vector<b2c_ctx *> &contexts, const std::vector<int> &robots, const std::vector<int> &objects,
                          const transf7 &refTran, const std::vector<SearchVariable> &variables, const SimAnnParams &params,
                          const std::vector<float> &initialStates, unsigned long long seed = 1234, bool contactLive = false,
                          double maxRadius = 200.0);
    ~MultiGpuSimAnnPlanner();
    bool valid() const {return mValid;}
    const std::string &error() const {return mError;}
    //! n annealing steps on every GPU (enqueued device by device, running concurrently), exchange + merge every step
    bool run(int nSteps);
    int getCurrentStep() const {return mCurrentStep;}
    const std::vector<PlannerState> &getBestList() const {return mBestList.list();}
    //! annealing steps each source rank had completed when its rows were last taken (+ 1; 0 = none seen)
    const std::vector<int> &stepsSeen() const {return mSteps;}
    int annealer(int rank) const {return mAnnealers[rank];}
};

//! ListPlanner: every input state evaluated in one call; legal ones enter the list while it has room or they beat its
//! worst entry (no similarity pruning, as in the reference)
bool listPlannerRun(b2c_ctx *ctx, int robot, int object, const transf7 &refTran, const std::vector<PlannerState> &states,
                    bool contactLive, std::vector<unsigned char> *legal, std::vector<float> *energy,
                    std::vector<PlannerState> *bestList, int nEigenGrasps);

//! TimeTester (reference src/EGPlanner/timeTest.cpp:35-64): the reference's in-tree throughput probe.  Each mainLoop
//! there perturbs the current state once (every free variable by r * mMaxJump * 0.1, one r in [-0.5, 0.5) per call) and
//! counts legal / illegal analyzeState results; here a loop perturbs it `batch` times and evaluates the batch in one call.
class BatchedTimeTester
{
    b2c_ctx *mCtx;
    int mRobot, mObject, mNumEG, mBatch;
    bool mLive;
    transf7 mRefTran;
    std::vector<SearchVariable> mVars;
    PlannerState mCurrentState;
    unsigned long long mSeed;
    long long mCount, mIllegalCount, mLoops;
  public:
    BatchedTimeTester(b2c_ctx *ctx, int robot, int object, const transf7 &refTran, const std::vector<SearchVariable> &variables,
                      const PlannerState &current, int nEigenGrasps, int batch, bool contactLive = false,
                      unsigned long long seed = 1234);
    //! one batched mainLoop; false on an engine error
    bool mainLoop();
    long long getCount() const {return mCount;}                // legal states so far (TimeTester::mCount)
    long long getIllegalCount() const {return mIllegalCount;}
    long long getLoops() const {return mLoops;}
};

//! one DOF as Hand::autoGrasp sees it (reference dof.h: getMin / getMax / getDefaultVelocity, DOF::getType)
struct DofDesc {
  double min, max, defaultVelocity;
  bool breakAway;              // <dof type="b"> (BreakAwayDOF); false: RigidDOF
};

//! Hand::autoGrasp(renderIt = false, speedFactor, stopAtContact), static branch (reference robot.cpp:2427-2462), for a
//! batch of hand states in one b2c_autograsp_batch call.  handPoses[i] = hand->getTran(), dofVals = [n][nDOF] row-major
//! (the contact-free posture HandObjectState::execute left).  closedDofs gets what hand->getDOFVals() would return after
//! the call, status the B2C_AG_* flags (B2C_AG_MOVED = autoGrasp's return value).
bool autoGraspBatch(b2c_ctx *ctx, int robot, const std::vector<transf7> &handPoses, const std::vector<double> &dofVals,
                    const std::vector<DofDesc> &dofs, double speedFactor, bool stopAtContact,
                    std::vector<double> *closedDofs, std::vector<int> *status, std::vector<double> *closedJoints = NULL,
                    int nJoints = 0);

}  // namespace b200

#endif
