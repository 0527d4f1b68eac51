// SimAnnPlanner with K annealing chains on the B200 engine, behind GraspIt!'s own planner interface
// (reference include/graspit/EGPlanner/simAnnPlanner.h, egPlanner.h).  Compiles against the reference's unmodified headers.
//
// SimAnnPlanner::mainLoop (src/EGPlanner/simAnnPlanner.cpp:111-148) runs ONE SimAnn::iterate -- one neighbour, one
// SearchEnergy::analyzeState -- and offers the new state to mBestList.  Here one mainLoop is one annealing step of
// numChains independent chains on the device (b2c_anneal_step: neighbour generation, FK, legality, CONTACT_ENERGY,
// Metropolis rule; src/EGPlanner/simAnn.cpp:96-261), followed by the reference's own best-list logic on the states the
// device reports as its best: the BEST_LIST_SIZE best (state || energy) rows of this GPU and, with several GPUs, of its
// peers (b2c_anneal_publish / b2c_exchange_collect).  Everything else -- EGPlanner's state machine, startPlanner /
// pausePlanner, getGrasp(i), addToListOfUniqueSolutions, stateDistance, the search-state classes -- is inherited unchanged.
#ifndef B200_SIM_ANN_PLANNER_H
#define B200_SIM_ANN_PLANNER_H

#include <vector>
#include <string>

#include "graspit/EGPlanner/simAnnPlanner.h"
#include "graspit/EGPlanner/simAnnParams.h"
#include "b2c.h"

class B200SimAnnPlanner : public SimAnnPlanner
{
    b2c_ctx *mCtx;
    int mRobotId, mObjectId, mNumChains, mAnnealer;
    int mRank, mWorld;
    unsigned long long mSeed;
    SimAnnParams mParams;
    std::vector<float> mRows;
    std::vector<int> mStepsSeen;
    std::string mError;
    bool createAnnealer();
    //! offers the rows collected from every GPU to mBestList (the JUMP branch of the reference's mainLoop, per row)
    void mergeRows();
  protected:
    virtual void mainLoop();
    virtual void resetParameters();
  public:
    //! rank / world: this planner's place among the GPUs that share the search (chains [rank * numChains, ...) of
    //! world * numChains; connect the contexts' exchanges with b2c_exchange_connect before the first step when world > 1)
    B200SimAnnPlanner(Hand *h, b2c_ctx *ctx, int robotId, int objectBodyId, int numChains, unsigned long long seed = 1234,
                      int rank = 0, int world = 1);
    ~B200SimAnnPlanner();
    void setAnnealingParameters(SimAnnParams params);
    virtual bool initialized();
    //! one mainLoop() for callers that drive the planner themselves instead of through EGPlanner's thread / idle sensor
    bool step();
    int annealer() const {return mAnnealer;}
    int getNumChains() const {return mNumChains;}
    const std::string &error() const {return mError;}
};

#endif
