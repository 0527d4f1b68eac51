// TEST INFRASTRUCTURE ONLY (oracle stub) -- never linked into the product.
#ifndef B2C_ORACLE_STUB_WORLD_H
#define B2C_ORACLE_STUB_WORLD_H
#include <Inventor/nodes/SoSeparator.h>
class CollisionInterface;
class World {
    SoSeparator mRoot;
    CollisionInterface *mCollisionInterface;
  public:
    World() : mCollisionInterface(0) {}
    SoSeparator *getIVRoot() { return &mRoot; }
    CollisionInterface *getCollisionInterface() const { return mCollisionInterface; }
    void setCollisionInterface(CollisionInterface *c) { mCollisionInterface = c; }
    // element management of EGPlanner's clone handling (egPlanner.cpp:85-95,205-233): nothing to manage here
    template <typename T> void destroyElement(T *, bool = true) {}
    template <typename T> void removeElementFromSceneGraph(T *) {}
    template <typename T> void addElementToSceneGraph(T *) {}
    template <typename T> void addRobot(T *, bool = true) {}
    template <typename A, typename B> void toggleCollisions(bool, A *, B *) {}
    template <typename A> void toggleCollisions(bool, A *) {}
};
#endif
