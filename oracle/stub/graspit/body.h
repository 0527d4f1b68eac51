// TEST INFRASTRUCTURE ONLY (oracle stub) -- never linked into the product.
// Shadows include/graspit/body.h (Qt/Coin-bound, cannot compile here) with the three
// members src/Collision/Graspit/graspitCollision.cpp actually uses:
// getGeometryTriangles (:211-212), getTran (:447,452,474-475) and getName (debug prints).
#ifndef B2C_ORACLE_STUB_BODY_H
#define B2C_ORACLE_STUB_BODY_H
#include <vector>
#include <string>
#include <QString>
#include "graspit/matvec3D.h"
#include "graspit/triangle.h"

class WorldElement;
class Body {
    std::vector<Triangle> tris_;
    transf tran_;
    std::string name_;
  public:
    Body(const std::string &name = "body") : name_(name) {}
    virtual ~Body() {}
    void setTriangles(const std::vector<Triangle> &t) { tris_ = t; }
    void getGeometryTriangles(std::vector<Triangle> *out) const { *out = tris_; }
    const transf &getTran() const { return tran_; }
    void setTranRaw(const transf &t) { tran_ = t; }
    QString getName() const { return QString(name_); }
    // Body::getOwner (include/graspit/body.h): the robot a link belongs to; a free body owns itself.  Last in the class so
    // that the stand-in's layout is the same for every translation unit (oracle/ref_autograsp.cpp overrides it in Link)
    virtual WorldElement *getOwner() { return 0; }
};
#endif
