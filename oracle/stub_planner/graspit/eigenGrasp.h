// TEST INFRASTRUCTURE ONLY (oracle stub) -- never linked into the product.
// Shadows include/graspit/eigenGrasp.h with the members the search-state code uses.  getDOF restates
// EigenGraspInterface::toDOFSpace for a rigid interface (src/eigenGrasp.cpp:567,638-646): dof = (E^T amp) * norm + origin.
#ifndef B2C_ORACLE_STUB_EIGENGRASP_H
#define B2C_ORACLE_STUB_EIGENGRASP_H
#include <vector>
class EigenGrasp {
  public:
    double mMin, mMax;
    bool mPredefinedLimits;
    std::vector<double> mVals;
    EigenGrasp() : mMin(0), mMax(0), mPredefinedLimits(false) {}
};
class EigenGraspInterface {
    std::vector<EigenGrasp> mGrasps;
    std::vector<double> mOrigin, mNorm;
    bool mRigid;
  public:
    EigenGraspInterface() : mRigid(false) {}
    void configure(int neg, int ndof, const double *eg, const double *origin, const double *norm) {
      mGrasps.assign(neg, EigenGrasp());
      for (int e = 0; e < neg; e++) mGrasps[e].mVals.assign(eg + (size_t)e * ndof, eg + (size_t)(e + 1) * ndof);
      mOrigin.assign(origin, origin + ndof); mNorm.assign(norm, norm + ndof);
    }
    int getSize() const { return (int)mGrasps.size(); }
    const EigenGrasp *getGrasp(int i) const { return &mGrasps[i]; }
    bool isRigid() const { return mRigid; }
    void setRigid(bool r) { mRigid = r; }
    void getDOF(const double *amp, double *dof) const {
      for (size_t d = 0; d < mOrigin.size(); d++) {
        double x = 0.0;
        for (size_t e = 0; e < mGrasps.size(); e++) x += mGrasps[e].mVals[d] * amp[e];
        dof[d] = x * mNorm[d] + mOrigin[d];
      }
    }
    void getAmp(double *amp, const double *) const { for (size_t e = 0; e < mGrasps.size(); e++) amp[e] = 0.0; }
};
#endif
