// B200Collision -- GraspIt! CollisionInterface backend on top of the b2c C ABI (include/b2c.h).
//
// Drop-in for `class GraspitCollision` (reference include/graspit/Collision/Graspit/graspitCollision.h:45-113):
// it derives from the reference's own abstract `CollisionInterface`
// (include/graspit/Collision/collisionInterface.h:44-185), implements every pure virtual with the reference's
// argument meaning, return values and error behaviour ("GCOL: model not found" on stderr, then false / 0 / void),
// and forwards the work to the CUDA engine in libb2c.so.  This file and b200Collision.cpp are the only sources a
// GraspIt! maintainer adds to the tree; they see GraspIt!'s headers and <b2c.h>, never CUDA.
//
// Beyond the reference interface the class exposes the engine context (context(), bodyId()) so that a batched
// SearchEnergy (b200SearchEnergy.h) can evaluate whole batches of hand postures against the same bodies.
#ifndef B200_COLLISION_H
#define B200_COLLISION_H

#include <map>
#include <vector>

#include "graspit/Collision/collisionInterface.h"

struct b2c_ctx;

class B200Collision : public CollisionInterface
{
  private:
    b2c_ctx *mCtx;
    //! Body -> engine body id (the reference keeps map<const Body*, CollisionModel*>, graspitCollision.h:54)
    std::map<const Body *, int> mBodyIds;
    //! engine body id -> Body (reports hand Body pointers back, graspitCollision.h:56)
    std::map<int, Body *> mBodies;
    //! engine mesh id of every body; a clone shares its original's mesh
    std::map<const Body *, int> mMeshIds;
    //! how many bodies reference a mesh (the mesh is destroyed with its last user)
    std::map<int, int> mMeshUsers;

    int getId(const Body *body) const;
    int uploadGeometry(Body *body);
    void releaseMesh(int meshId);
    void convertInterestList(const std::vector<Body *> *inList, std::vector<int> *out) const;

  public:
    //! device < 0 selects the device named by the environment variable B2C_DEVICE (default 0)
    explicit B200Collision(int device = -1);
    virtual ~B200Collision();

    //! false when no CUDA device could be opened; every query then fails like an unknown body (no CPU path)
    bool valid() const {return mCtx != NULL;}
    b2c_ctx *context() const {return mCtx;}
    //! engine id of a registered body, -1 when unknown
    int bodyId(const Body *body) const {return getId(body);}

    // adding and moving bodies
    virtual bool addBody(Body *body, bool ExpectEmpty = false);
    virtual bool updateBodyGeometry(Body *body, bool ExpectEmpty = false);
    virtual void removeBody(Body *body);
    virtual void cloneBody(Body *clone, const Body *original);
    virtual void setBodyTransform(Body *body, const transf &t);

    // enable / disable collision
    virtual void activateBody(const Body *body, bool active);
    virtual void activatePair(const Body *body1, const Body *body2, bool active);
    virtual bool isActive(const Body *body1, const Body *body2 = NULL);

    // collision detection
    virtual int allCollisions(DetectionType type, CollisionReport *report,
                              const std::vector<Body *> *interestList);

    // contact
    virtual int allContacts(CollisionReport *report, double threshold,
                            const std::vector<Body *> *interestList);
    virtual int contact(ContactReport *report, double threshold,
                        const Body *body1, const Body *body2);

    // point-to-body and body-to-body distances
    virtual double pointToBodyDistance(const Body *body1, position point,
                                       position &closestPoint, vec3 &closestNormal);
    virtual double bodyToBodyDistance(const Body *body1, const Body *body2,
                                      position &p1, position &p2);

    // region on a body around a point
    virtual void bodyRegion(const Body *body, position point, vec3 normal,
                            double radius, Neighborhood *neighborhood);

    virtual void getBoundingVolumes(const Body *body, int depth, std::vector<BoundingBox> *bvs);
};

#endif
