// TEST INFRASTRUCTURE ONLY (oracle stub) -- never linked into the product.
#ifndef B2C_ORACLE_STUB_GRASPITCORE_H
#define B2C_ORACLE_STUB_GRASPITCORE_H
#include "graspit/world.h"
class GraspitCore {
    World *mWorld;
  public:
    GraspitCore() : mWorld(0) {}
    World *getWorld() const { return mWorld; }
    void setWorld(World *w) { mWorld = w; }
};
extern GraspitCore *graspitCore;
#endif
