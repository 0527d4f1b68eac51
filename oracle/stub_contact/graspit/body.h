// TEST INFRASTRUCTURE ONLY (oracle stub) -- never linked into the product.
// Shadows include/graspit/body.h for oracle/ref_contact.cpp with the members the reference's contact code touches:
// Contact::Contact / updateCof / computeWrenches / wrenchFromFrictionEdge (src/contact/contact.cpp:67-165),
// SoftContact (src/contact/softContact.cpp), contactSetting.cpp:39-191.
#ifndef B2C_ORACLE_STUB_CONTACT_BODY_H
#define B2C_ORACLE_STUB_CONTACT_BODY_H
#include <list>
#include <string>
#include <QString>
#include "graspit/matvec3D.h"
class World;
class Contact;
class Body {
  public:
    transf tran_;
    position cog_;
    double maxRadius_, youngs_;
    int material_;
    World *world_;
    std::string name_;
    Body() : cog_(0, 0, 0), maxRadius_(1.0), youngs_(-1.0), material_(0), world_(0), name_("body") {}
    virtual ~Body() {}
    const transf &getTran() const { return tran_; }
    position const &getCoG() const { return cog_; }
    double getMaxRadius() const { return maxRadius_; }
    double getYoungs() { return youngs_; }
    int getMaterial() const { return material_; }
    World *getWorld() const { return world_; }
    QString getName() const { return QString(name_); }
    void setContactsChanged() {}
    void removeContact(Contact *) {}
};
class GraspableBody : public Body {};
#endif
