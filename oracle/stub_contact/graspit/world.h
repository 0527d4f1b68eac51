// TEST INFRASTRUCTURE ONLY (oracle stub) -- never linked into the product.
// The two World services the reference's contact code asks for: the friction table (Contact::updateCof,
// src/contact/contact.cpp:401-408) and FindRegion (findSoftNeighborhoods, src/contactSetting.cpp:108-133), which here
// only records the radius it was asked for.
#ifndef B2C_ORACLE_STUB_CONTACT_WORLD_H
#define B2C_ORACLE_STUB_CONTACT_WORLD_H
#include "graspit/Collision/collisionStructures.h"
class Body;
class World {
  public:
    double cof_, kcof_, lastRegionRadius_;
    World() : cof_(0.0), kcof_(0.0), lastRegionRadius_(0.0) {}
    double getCOF(int, int) { return cof_; }
    double getKCOF(int, int) { return kcof_; }
    void FindRegion(const Body *, position, vec3, double rad, Neighborhood *) { lastRegionRadius_ = rad; }
};
#endif
