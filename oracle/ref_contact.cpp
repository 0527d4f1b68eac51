// TEST INFRASTRUCTURE ONLY -- the oracle.  Never imported, linked or executed by the product path.
//
// Reference-pinned contact set-up: contact frames, friction edges, boundary wrenches of point and soft
// contacts, the normal check and the soft-neighbourhood merge.
//
// The class definitions are the reference's own headers (include/graspit/contact/contact.h, pointContact.h,
// softContact.h, FitParabola.h), compiled against stand-in Body / World headers (oracle/stub_contact).  The member
// functions and the free functions of contactSetting.cpp are the reference's OWN SOURCE LINES, cut out of the files where
// they lie under /root/reference by oracle/Makefile (sed line ranges, checked by grep) into oracle/_ref/obj/extract/*.inc
// at build time and #included below.  Their translation units as a whole need Coin (visual indicators), the Matrix class
// (LAPACK) and the real Body / World, and cannot be compiled here.  Nothing is copied into the repo.
//   contact.cpp:67-94      Contact::Contact(Body*, Body*, position, vec3)        (frame = transf::COORDINATE(...))
//   contact.cpp:124-142    Contact::wrenchFromFrictionEdge
//   contact.cpp:150-163    Contact::computeWrenches
//   contact.cpp:180-215    Contact::setUpFrictionEllipsoid
//   contact.cpp:401-408    Contact::updateCof
//   pointContact.cpp:18-33, 44-56    PointContact::PointContact, PointContact::setUpFrictionEdges
//   softContact.cpp:22-58, 69-122, 130-179, 185-359
//                          SoftContact::SoftContact, setUpFrictionEdges, computeWrenches, FitPoints, CalcRelPhi,
//                          CalcRprimes, CalcContact_Mattress
//   contactSetting.cpp:40-186   contactDistance, checkContactNormals, neighborhoodsOverlap, mergeNeighborhoods,
//                               findSoftNeighborhoods, mergeSoftNeighborhoods
// What IS written here: destructors without the world bookkeeping, null visual indicators, and the C entry points.
#include <cstdio>
#include <cstring>
#include <cmath>
#include <vector>
#include <list>

#include "graspit/matvec3D.h"
#include "graspit/contact/contact.h"
#include "graspit/contact/pointContact.h"
#include "graspit/contact/softContact.h"
#include "graspit/world.h"
#include "graspit/body.h"
#include "graspit/mytools.h"
#include "graspit/FitParabola.h"
#include "graspit/debug.h"

#include "extract/contact_constants.inc"        // contact.cpp:58-60: THRESHOLD, INHERITANCE_THRESHOLD, INHERITANCE_ANGULAR_THRESHOLD

#include "extract/contact_ctor.inc"
#include "extract/contact_wrenchFromFrictionEdge.inc"
#include "extract/contact_computeWrenches.inc"
#include "extract/contact_setUpFrictionEllipsoid.inc"
#include "extract/contact_updateCof.inc"
Contact::~Contact() {
  if (optimalCoeffs) { delete [] optimalCoeffs; }
  if (wrench) { delete [] wrench; }
  if (prevBetas) { delete [] prevBetas; }
}

#include "extract/pointContact_ctor.inc"
#include "extract/pointContact_setUpFrictionEdges.inc"
PointContact::~PointContact() {}
SoSeparator *PointContact::getVisualIndicator() { return NULL; }

#include "extract/softContact_ctor.inc"
#include "extract/softContact_setUpFrictionEdges.inc"
#include "extract/softContact_computeWrenches.inc"
#include "extract/softContact_fit_and_mattress.inc"
SoftContact::~SoftContact() { delete [] bodyNghbd; }
SoSeparator *SoftContact::getVisualIndicator() { return NULL; }

#include "extract/contactSetting_40_186.inc"

namespace {
transf mk7(const double *q7) { return transf(Quaternion(q7[0], q7[1], q7[2], q7[3]), vec3(q7[4], q7[5], q7[6])); }
void put7(const transf &t, double *o) {
  o[0] = t.rotation().w(); o[1] = t.rotation().x(); o[2] = t.rotation().y(); o[3] = t.rotation().z();
  o[4] = t.translation().x(); o[5] = t.translation().y(); o[6] = t.translation().z();
}
struct Scene {
  World w;
  GraspableBody b1, b2;
  Scene(const double *t1, const double *t2, double cof, const double *cog, double maxRadius, double y1, double y2) {
    w.cof_ = cof; w.kcof_ = cof;
    b1.world_ = &w; b2.world_ = &w;
    if (t1) b1.tran_ = mk7(t1);
    if (t2) b2.tran_ = mk7(t2);
    if (cog) b1.cog_ = position(cog[0], cog[1], cog[2]);
    b1.maxRadius_ = maxRadius; b1.youngs_ = y1; b2.youngs_ = y2;
  }
};
// access to what the reference keeps protected
struct PointContactX : public PointContact {
  PointContactX(Body *a, Body *b, position p, vec3 n) : PointContact(a, b, p, n) {}
  int edges() const { return numFrictionEdges; }
  const double *edge(int i) const { return &frictionEdges[6 * i]; }
  const Wrench &w(int i) const { return wrench[i]; }
  int nw() const { return numFCWrenches; }
};
struct SoftContactX : public SoftContact {
  SoftContactX(Body *a, Body *b, position p, vec3 n, Neighborhood *bn) : SoftContact(a, b, p, n, bn) {}
  const Wrench &w(int i) const { return wrench[i]; }
  int nw() const { return numFCWrenches; }
  void info(double *o) const { o[0] = majorAxis; o[1] = minorAxis; o[2] = relPhi; o[3] = r1prime; o[4] = r2prime; o[5] = a; o[6] = b; o[7] = c; o[8] = r1; o[9] = r2; }
};
} // namespace

extern "C" {

// PointContact on body 1: frame (Contact::Contact), the 8 friction edges (6 numbers each), the 8 boundary wrenches
// (force, torque).  pos / normal in body-1 coordinates.  Returns the number of wrenches.
int ref_point_contact(const double *pos, const double *normal, double cof, const double *cog, double maxRadius,
                      double *frame7, double *edges48, double *wrenches48) {
  Scene S(NULL, NULL, cof, cog, maxRadius, -1, -1);
  PointContactX c(&S.b1, &S.b2, position(pos[0], pos[1], pos[2]), vec3(normal[0], normal[1], normal[2]));
  c.computeWrenches();
  if (frame7) put7(c.getContactFrame(), frame7);
  for (int i = 0; i < c.edges() && edges48; i++) memcpy(edges48 + 6 * i, c.edge(i), 6 * sizeof(double));
  for (int i = 0; i < c.nw() && wrenches48; i++) {
    const Wrench &w = c.w(i);
    for (int k = 0; k < 3; k++) { wrenches48[6 * i + k] = w.force[k]; wrenches48[6 * i + 3 + k] = w.torque[k]; }
  }
  return c.nw();
}

// checkContactNormals (contactSetting.cpp:54-81) on one contact row (b1_pos, b2_pos, b1_normal, b2_normal), in place
int ref_check_contact_normals(const double *t1, const double *t2, double *row12) {
  Scene S(t1, t2, 0, NULL, 1, -1, -1);
  ContactData c(position(row12[0], row12[1], row12[2]), position(row12[3], row12[4], row12[5]),
                vec3(row12[6], row12[7], row12[8]), vec3(row12[9], row12[10], row12[11]));
  const bool ok = checkContactNormals(&S.b1, &S.b2, &c);
  for (int k = 0; k < 3; k++) { row12[6 + k] = c.b1_normal[k]; row12[9 + k] = c.b2_normal[k]; }
  return ok ? 1 : 0;
}

// the region radius findSoftNeighborhoods asks World::FindRegion for (contactSetting.cpp:108-133)
double ref_soft_neighborhood_radius(double y1, double y2) {
  Scene S(NULL, NULL, 0, NULL, 1, y1, y2);
  ContactReport set;
  set.push_back(ContactData());
  findSoftNeighborhoods(&S.b1, &S.b2, set);
  return S.w.lastRegionRadius_;
}

// mergeSoftNeighborhoods (contactSetting.cpp:139-191).  rows: n x 12; neighbourhoods as flat xyz with n + 1 offsets (in
// points) per side.  Outputs: surviving original indices (recognised by their distSq tag), merged neighbourhood sizes.
// Merged neighbourhoods are returned flat in out_n1 / out_n2 (caller sizes them for the sum of all inputs).
int ref_merge_soft_neighborhoods(const double *t1, const double *t2, int n, const double *rows,
                                 const double *n1, const int *off1, const double *n2, const int *off2,
                                 int *keep, double *out_n1, int *out_off1, double *out_n2, int *out_off2) {
  Scene S(t1, t2, 0, NULL, 1, -1, -1);
  ContactReport set;
  for (int i = 0; i < n; i++) {
    const double *r = rows + 12 * i;
    ContactData c(position(r[0], r[1], r[2]), position(r[3], r[4], r[5]), vec3(r[6], r[7], r[8]), vec3(r[9], r[10], r[11]), (double)i);
    for (int k = off1[i]; k < off1[i + 1]; k++) c.nghbd1.push_back(position(n1[3 * k], n1[3 * k + 1], n1[3 * k + 2]));
    for (int k = off2[i]; k < off2[i + 1]; k++) c.nghbd2.push_back(position(n2[3 * k], n2[3 * k + 1], n2[3 * k + 2]));
    set.push_back(c);
  }
  mergeSoftNeighborhoods(&S.b1, &S.b2, set);
  out_off1[0] = out_off2[0] = 0;
  for (size_t i = 0; i < set.size(); i++) {
    keep[i] = (int)set[i].distSq;
    int a = out_off1[i], b = out_off2[i];
    for (size_t k = 0; k < set[i].nghbd1.size(); k++, a++) for (int d = 0; d < 3; d++) out_n1[3 * a + d] = set[i].nghbd1[k][d];
    for (size_t k = 0; k < set[i].nghbd2.size(); k++, b++) for (int d = 0; d < 3; d++) out_n2[3 * b + d] = set[i].nghbd2[k][d];
    out_off1[i + 1] = a; out_off2[i + 1] = b;
  }
  return (int)set.size();
}

// A SoftContact and its mate (addContacts, contactSetting.cpp:203-221: c1 on body 1 with neighbourhood 1, c2 on body 2
// with neighbourhood 2, mates set, setUpFrictionEdges + computeWrenches on both).  Positions, normals and neighbourhoods
// in the respective body's coordinates.  Outputs for side 1: the wrenches (force, torque) and
// info10 = majorAxis, minorAxis, relPhi, r1prime, r2prime, a, b, c, r1, r2; info10_mate the same for side 2.
int ref_soft_contact_pair(const double *t1, const double *t2, const double *pos1, const double *nrm1, int np1, const double *ng1,
                          const double *pos2, const double *nrm2, int np2, const double *ng2, double y1, double y2,
                          double cof, const double *cog1, double maxRadius1, double *wrenches, double *info10, double *info10_mate) {
  Scene S(t1, t2, cof, cog1, maxRadius1, y1, y2);
  Neighborhood a, b;
  for (int k = 0; k < np1; k++) a.push_back(position(ng1[3 * k], ng1[3 * k + 1], ng1[3 * k + 2]));
  for (int k = 0; k < np2; k++) b.push_back(position(ng2[3 * k], ng2[3 * k + 1], ng2[3 * k + 2]));
  SoftContactX *c1 = new SoftContactX(&S.b1, &S.b2, position(pos1[0], pos1[1], pos1[2]), vec3(nrm1[0], nrm1[1], nrm1[2]), &a);
  SoftContactX *c2 = new SoftContactX(&S.b2, &S.b1, position(pos2[0], pos2[1], pos2[2]), vec3(nrm2[0], nrm2[1], nrm2[2]), &b);
  c1->setMate(c2); c2->setMate(c1);
  c1->setUpFrictionEdges();
  c1->computeWrenches();
  const int nw = c1->nw();
  for (int i = 0; i < nw && wrenches; i++) {
    const Wrench &w = c1->w(i);
    for (int k = 0; k < 3; k++) { wrenches[6 * i + k] = w.force[k]; wrenches[6 * i + 3 + k] = w.torque[k]; }
  }
  if (info10) c1->info(info10);
  if (info10_mate) c2->info(info10_mate);
  c1->setMate(NULL); c2->setMate(NULL);
  delete c1; delete c2;
  return nw;
}

} // extern "C"
