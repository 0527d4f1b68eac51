// TEST INFRASTRUCTURE ONLY (oracle stub) -- the collision oracle's Body stand-in plus isA (searchState.cpp:598).
#ifndef B2C_ORACLE_STUB_PLANNER_BODY_H
#define B2C_ORACLE_STUB_PLANNER_BODY_H
#include "../../stub/graspit/body.h"
class GraspableBody : public Body {
  public:
    GraspableBody(const std::string &n = "object") : Body(n) {}
    bool isA(const char *) const { return true; }
    double getMaxRadius() const { return 200.0; }
    void setTran(const transf &t) { setTranRaw(t); }
};
#endif
