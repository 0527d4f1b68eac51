// TEST INFRASTRUCTURE ONLY -- the oracle.  Never imported, linked or executed by the product
// path (graspit_b200/, include/); only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load the library this file is built into.
//
// C-ABI driver around the reference's OWN GraspitCollision backend.  The reference sources
// (/root/reference/src/Collision/Graspit/*.cpp, src/Collision/collisionInterface.cpp,
// src/{bBox,triangle,transform}.cpp, qhull/src/libqhull/*.c) are compiled verbatim from where
// they lie by oracle/Makefile against oracle/shim (mini-Eigen, Qt/Coin stubs) and linked with
// this file into oracle/_ref/libgraspit_ref.so.  Nothing below re-implements collision
// geometry: every query forwards to the reference class.  The only logic written here is
//   * the Body stub plumbing (ids <-> Body*),
//   * counter-exposing subclasses of the reference callbacks (collisionAlgorithms.h:91-98),
//   * ref_eval_batch(): the caller loop of SearchEnergy::analyzeState / ContactEnergy::energy
//     (src/EGPlanner/energy/searchEnergy.cpp:148-188, contactEnergy.cpp:9-55,
//     src/world.cpp:1478-1508,1539-1549,1618-1651, src/contact/virtualContact.cpp:75-124),
//     which is Qt-bound in the reference and therefore restated.
#include <vector>
#include <array>
#include <map>
#include <thread>
#include <cstring>
#include <cmath>
#include <cstdio>
#include <unistd.h>
#include <fcntl.h>

#include "graspit/body.h"            // oracle/stub
#define private public               // reach GraspitCollision::getModel for counter queries
#include "graspit/Collision/Graspit/graspitCollision.h"
#undef private
#include "graspit/Collision/Graspit/collisionModel.h"
#include "graspit/Collision/Graspit/collisionAlgorithms.h"
#include "graspit/bBox.h"
#include "graspit/FitParabola.h"      // the reference's own paraboloid fit (header-only)

// symbols the reference expects from translation units we do not compile
QMutex qhull_mutex;                                                   // src/gws.cpp
SbRotation QuaterniontoSbRotation(Quaternion q) {                      // src/matvec.cpp
  return SbRotation((float)q.x(), (float)q.y(), (float)q.z(), (float)q.w());
}
SbVec3f toSbVec3f(vec3 v) { return SbVec3f((float)v.x(), (float)v.y(), (float)v.z()); }

namespace {

struct StatCollision : public Collision::CollisionCallback {
  StatCollision(const CollisionModel *a, const CollisionModel *b) : CollisionCallback(a, b) {}
  void stats(int *s) const { s[0] = mNumQuickTests; s[1] = mNumLeafTests; s[2] = mNumTriangleTests; }
};
struct StatDistance : public Collision::DistanceCallback {
  StatDistance(const CollisionModel *a, const CollisionModel *b) : DistanceCallback(a, b) {}
  void stats(int *s) const { s[0] = mNumQuickTests; s[1] = mNumLeafTests; s[2] = mNumTriangleTests; }
};
struct StatClosest : public Collision::ClosestPtCallback {
  StatClosest(const CollisionModel *a, const position &p) : ClosestPtCallback(a, p) {}
  void stats(int *s) const { s[0] = mNumQuickTests; s[1] = mNumLeafTests; s[2] = mNumTriangleTests; }
};
struct StatContact : public Collision::ContactCallback {
  StatContact(double thr, const CollisionModel *a, const CollisionModel *b) : ContactCallback(thr, a, b) {}
  void stats(int *s) const { s[0] = mNumQuickTests; s[1] = mNumLeafTests; s[2] = mNumTriangleTests; }
};

// Records which leaf pairs the reference's contact recursion reaches, in order, and how many contacts each appended
// (ContactCallback::leafTest, collisionAlgorithms_inl.h:96-133): the ground truth for the product's visit-order replay.
struct TraceContact : public Collision::ContactCallback {
  struct Visit { const Collision::Leaf *l1, *l2; int added; };
  std::vector<Visit> visits;
  TraceContact(double thr, const CollisionModel *a, const CollisionModel *b) : ContactCallback(thr, a, b) {}
  virtual void leafTest(const Collision::Leaf *l1, const Collision::Leaf *l2) {
    const size_t before = getReport().size();
    Collision::ContactCallback::leafTest(l1, l2);
    visits.push_back(Visit{l1, l2, (int)(getReport().size() - before)});
  }
};

struct RefWorld {
  GraspitCollision gc;
  std::vector<Body *> bodies;          // id -> Body* (NULL once removed)
  std::map<const Body *, int> ids;
  ~RefWorld() {
    // clones share the original's tree (collisionModel.cpp:466-474): remove clones first
    for (int pass = 0; pass < 2; pass++)
      for (size_t i = 0; i < bodies.size(); i++) {
        Body *b = bodies[i];
        if (!b) continue;
        bool isClone = clone[i];
        if ((pass == 0) == isClone) { gc.removeBody(b); delete b; bodies[i] = NULL; }
      }
  }
  std::vector<bool> clone;
  Body *get(int id) const { return (id >= 0 && id < (int)bodies.size()) ? bodies[id] : NULL; }
};

transf makeTransf(const double *q, const double *t) {
  return transf(Quaternion(q[0], q[1], q[2], q[3]), vec3(t[0], t[1], t[2]));
}

void writeContact(const ContactData &c, double *o) {
  for (int k = 0; k < 3; k++) {
    o[k] = c.b1_pos[k]; o[3 + k] = c.b2_pos[k];
    o[6 + k] = c.b1_normal[k]; o[9 + k] = c.b2_normal[k];
  }
  o[12] = c.distSq;
}

// Node::getBVRecurse prints to stdout on every call (collisionModel.cpp:439,450): silence it.
struct StdoutMute {
  int saved;
  StdoutMute() { fflush(stdout); saved = dup(1); int n = open("/dev/null", O_WRONLY); dup2(n, 1); close(n); }
  ~StdoutMute() { std::cout.flush(); fflush(stdout); dup2(saved, 1); close(saved); }
};

} // namespace

extern "C" {

void *ref_create() { return new RefWorld(); }
void ref_destroy(void *w) { delete (RefWorld *)w; }

// tri9: ntri x 9 doubles (v1 v2 v3).  Returns the new body id, or -1.
int ref_add_body(void *wv, const double *tri9, int ntri) {
  RefWorld *w = (RefWorld *)wv;
  std::vector<Triangle> tris;
  tris.reserve(ntri);
  for (int i = 0; i < ntri; i++) {
    const double *p = tri9 + 9 * i;
    tris.push_back(Triangle(position(p[0], p[1], p[2]), position(p[3], p[4], p[5]), position(p[6], p[7], p[8])));
  }
  char name[32]; snprintf(name, sizeof name, "body%d", (int)w->bodies.size());
  Body *b = new Body(name);
  b->setTriangles(tris);
  if (!w->gc.addBody(b)) { delete b; return -1; }
  w->bodies.push_back(b); w->clone.push_back(false);
  w->ids[b] = (int)w->bodies.size() - 1;
  return (int)w->bodies.size() - 1;
}

int ref_clone_body(void *wv, int original) {
  RefWorld *w = (RefWorld *)wv;
  Body *o = w->get(original);
  if (!o) return -1;
  Body *b = new Body("clone");
  w->gc.cloneBody(b, o);
  w->bodies.push_back(b); w->clone.push_back(true);
  w->ids[b] = (int)w->bodies.size() - 1;
  return (int)w->bodies.size() - 1;
}

void ref_remove_body(void *wv, int id) {
  RefWorld *w = (RefWorld *)wv;
  Body *b = w->get(id);
  if (!b) return;
  w->gc.removeBody(b);
  w->ids.erase(b);
  delete b;
  w->bodies[id] = NULL;
}

// mirrors Body::setTran (src/body.cpp:927-942): store on the body, push to the backend
void ref_set_transform(void *wv, int id, const double *q_wxyz, const double *t) {
  RefWorld *w = (RefWorld *)wv;
  Body *b = w->get(id);
  if (!b) return;
  transf tr = makeTransf(q_wxyz, t);
  b->setTranRaw(tr);
  w->gc.setBodyTransform(b, tr);
}

void ref_activate_body(void *wv, int id, int active) {
  RefWorld *w = (RefWorld *)wv;
  if (w->get(id)) w->gc.activateBody(w->get(id), active != 0);
}
void ref_activate_pair(void *wv, int a, int b, int active) {
  RefWorld *w = (RefWorld *)wv;
  if (w->get(a) && w->get(b)) w->gc.activatePair(w->get(a), w->get(b), active != 0);
}
int ref_is_active(void *wv, int a, int b) {
  RefWorld *w = (RefWorld *)wv;
  if (!w->get(a)) return 0;
  return w->gc.isActive(w->get(a), b >= 0 ? w->get(b) : NULL) ? 1 : 0;
}

// allCollisions; pairs_out receives (id1,id2) per colliding pair (ALL_COLLISIONS only).
int ref_all_collisions(void *wv, int fast, const int *interest, int ninterest, int *pairs_out, int cap) {
  RefWorld *w = (RefWorld *)wv;
  std::vector<Body *> il;
  for (int i = 0; i < ninterest; i++) if (w->get(interest[i])) il.push_back(w->get(interest[i]));
  CollisionReport report;
  int n = w->gc.allCollisions(fast ? CollisionInterface::FAST_COLLISION : CollisionInterface::ALL_COLLISIONS,
                              fast ? NULL : &report, interest ? &il : NULL);
  if (!fast && pairs_out)
    for (int i = 0; i < (int)report.size() && i < cap; i++) {
      pairs_out[2 * i] = w->ids[report[i].first];
      pairs_out[2 * i + 1] = w->ids[report[i].second];
    }
  return n;
}

double ref_body_distance(void *wv, int a, int b, double *p1o, double *p2o) {
  RefWorld *w = (RefWorld *)wv;
  if (!w->get(a) || !w->get(b)) return 0.0;
  position p1, p2;
  double d = w->gc.bodyToBodyDistance(w->get(a), w->get(b), p1, p2);
  for (int k = 0; k < 3; k++) { p1o[k] = p1[k]; p2o[k] = p2[k]; }
  return d;
}

double ref_point_distance(void *wv, int body, const double *p, double *cp, double *cn) {
  RefWorld *w = (RefWorld *)wv;
  if (!w->get(body)) return 0.0;
  position c; vec3 n;
  double d = w->gc.pointToBodyDistance(w->get(body), position(p[0], p[1], p[2]), c, n);
  for (int k = 0; k < 3; k++) { cp[k] = c[k]; cn[k] = n[k]; }
  return d;
}

// contact(): full reference post-processing (duplicate removal + compactContactSet/qhull)
int ref_contact(void *wv, int a, int b, double thr, double *out13, int cap) {
  RefWorld *w = (RefWorld *)wv;
  if (!w->get(a) || !w->get(b)) return 0;
  ContactReport rep;
  int n = w->gc.contact(&rep, thr, w->get(a), w->get(b));
  for (int i = 0; i < n && i < cap; i++) writeContact(rep[i], out13 + 13 * i);
  return n;
}

// raw ContactCallback output before any post-processing; stats3 optional
int ref_contact_raw(void *wv, int a, int b, double thr, double *out13, int cap, int *stats3) {
  RefWorld *w = (RefWorld *)wv;
  if (!w->get(a) || !w->get(b)) return 0;
  StatContact cc(thr, w->gc.getModel(w->get(a)), w->gc.getModel(w->get(b)));
  Collision::startRecursion(w->gc.getModel(w->get(a)), w->gc.getModel(w->get(b)), &cc);
  const ContactReport &rep = cc.getReport();
  for (int i = 0; i < (int)rep.size() && i < cap; i++) writeContact(rep[i], out13 + 13 * i);
  if (stats3) cc.stats(stats3);
  return (int)rep.size();
}

// Leaf pairs in the order the reference's ContactCallback visits them, restricted to those that appended contacts:
// visits3 = (triangle of a, triangle of b, contacts appended) per visit, triangles identified by their position in the
// lists the bodies were created from (tri9_a / tri9_b: the same arrays given to ref_add_body).  out13 (optional): the raw
// report in the same order.  Returns the number of such visits.
int ref_contact_trace(void *wv, int a, int b, double thr, const double *tri9_a, int ntri_a, const double *tri9_b, int ntri_b,
                      int *visits3, int cap_visits, double *out13, int cap_contacts) {
  RefWorld *w = (RefWorld *)wv;
  if (!w->get(a) || !w->get(b)) return 0;
  typedef std::array<double, 9> Key;
  std::map<Key, int> ida, idb;
  for (int i = ntri_a - 1; i >= 0; i--) { Key k; for (int j = 0; j < 9; j++) k[j] = tri9_a[9 * i + j]; ida[k] = i; }
  for (int i = ntri_b - 1; i >= 0; i--) { Key k; for (int j = 0; j < 9; j++) k[j] = tri9_b[9 * i + j]; idb[k] = i; }
  TraceContact cc(thr, w->gc.getModel(w->get(a)), w->gc.getModel(w->get(b)));
  Collision::startRecursion(w->gc.getModel(w->get(a)), w->gc.getModel(w->get(b)), &cc);
  auto key_of = [](const Collision::Leaf *l) {
    const Triangle &t = l->getTriangles()->front();
    Key k = {t.v1.x(), t.v1.y(), t.v1.z(), t.v2.x(), t.v2.y(), t.v2.z(), t.v3.x(), t.v3.y(), t.v3.z()};
    return k;
  };
  int n = 0;
  for (auto &v : cc.visits) {
    if (v.added <= 0) continue;
    if (n < cap_visits) {
      auto fa = ida.find(key_of(v.l1)), fb = idb.find(key_of(v.l2));
      visits3[3 * n] = fa == ida.end() ? -1 : fa->second;
      visits3[3 * n + 1] = fb == idb.end() ? -1 : fb->second;
      visits3[3 * n + 2] = v.added;
    }
    n++;
  }
  const ContactReport &rep = cc.getReport();
  if (out13) for (int i = 0; i < (int)rep.size() && i < cap_contacts; i++) writeContact(rep[i], out13 + 13 * i);
  return n;
}

// Exhaustive body-body distance: the reference's own leaf arithmetic (DistanceCallback::leafTest,
// collisionAlgorithms_inl.h:176-204: t1.applyTransform(mTran1To2), triangleTriangleDistanceSq) over EVERY triangle pair,
// no hierarchy and no pruning -- the arbiter when the engine's and the reference's tree searches disagree about which
// pair is closest.  p1 (body-a frame) and p2 (body-b frame) are the callback's mP1 / mP2, i.e. WITHOUT the extra inverse
// bodyToBodyDistance applies afterwards; tri_pair = the winning triangles.  Returns the distance, -1 on intersection.
double ref_brute_distance(void *wv, int a, int b, double *p1o, double *p2o, int *tri_pair) {
  RefWorld *w = (RefWorld *)wv;
  if (!w->get(a) || !w->get(b)) return 0.0;
  std::vector<Triangle> ta, tb;
  w->get(a)->getGeometryTriangles(&ta);
  w->get(b)->getGeometryTriangles(&tb);
  const transf t1 = w->get(a)->getTran(), t2 = w->get(b)->getTran();
  const transf tran2To1 = t1.inverse() % t2, tran1To2 = t2.inverse() % t1;     // RecursionCallback ctor
  double best = 1.0e300;
  position bp1(0, 0, 0), bp2(0, 0, 0);
  int bi = -1, bj = -1;
  for (size_t i = 0; i < ta.size() && best >= 0.0; i++) {
    Triangle t(ta[i]);
    t.applyTransform(tran1To2);
    for (size_t j = 0; j < tb.size() && best >= 0.0; j++) {
      position p1(0, 0, 0), p2(0, 0, 0);
      double d = triangleTriangleDistanceSq(t, tb[j], p1, p2);
      if (d < best) { best = d; bp1 = tran2To1 * p1; bp2 = p2; bi = (int)i; bj = (int)j; }
    }
  }
  if (best < 0.0) { bp1 = bp2 = position(0, 0, 0); }
  for (int k = 0; k < 3; k++) { if (p1o) p1o[k] = bp1[k]; if (p2o) p2o[k] = bp2[k]; }
  if (tri_pair) { tri_pair[0] = bi; tri_pair[1] = bj; }
  return best < 0.0 ? -1.0 : sqrt(best);
}

// allContacts(): pair_ids gets (id1,id2) per reported pair, counts the contacts per pair,
// out13 the contacts of all pairs back to back.  Returns number of pairs.
int ref_all_contacts(void *wv, double thr, const int *interest, int ninterest,
                     int *pair_ids, int *counts, int cap_pairs, double *out13, int cap_contacts) {
  RefWorld *w = (RefWorld *)wv;
  std::vector<Body *> il;
  for (int i = 0; i < ninterest; i++) if (w->get(interest[i])) il.push_back(w->get(interest[i]));
  CollisionReport report;
  int n = w->gc.allContacts(&report, thr, interest ? &il : NULL);
  int k = 0;
  for (int i = 0; i < n && i < cap_pairs; i++) {
    pair_ids[2 * i] = w->ids[report[i].first];
    pair_ids[2 * i + 1] = w->ids[report[i].second];
    counts[i] = (int)report[i].contacts.size();
    for (size_t c = 0; c < report[i].contacts.size(); c++, k++)
      if (k < cap_contacts) writeContact(report[i].contacts[c], out13 + 13 * k);
  }
  return n;
}

int ref_region(void *wv, int body, const double *p, const double *n, double radius, double *out3, int cap) {
  RefWorld *w = (RefWorld *)wv;
  if (!w->get(body)) return 0;
  Neighborhood nb;
  w->gc.bodyRegion(w->get(body), position(p[0], p[1], p[2]), vec3(n[0], n[1], n[2]), radius, &nb);
  for (int i = 0; i < (int)nb.size() && i < cap; i++)
    for (int k = 0; k < 3; k++) out3[3 * i + k] = nb[i][k];
  return (int)nb.size();
}

// out10 per box: translation(3), quaternion wxyz(4), halfSize(3)
int ref_bounding_volumes(void *wv, int body, int depth, double *out10, int cap) {
  RefWorld *w = (RefWorld *)wv;
  if (!w->get(body)) return 0;
  std::vector<BoundingBox> bvs;
  { StdoutMute mute; w->gc.getBoundingVolumes(w->get(body), depth, &bvs); }
  for (int i = 0; i < (int)bvs.size() && i < cap; i++) {
    const transf &t = bvs[i].getTran();
    double *o = out10 + 10 * i;
    for (int k = 0; k < 3; k++) { o[k] = t.translation()[k]; o[7 + k] = bvs[i].halfSize[k]; }
    o[3] = t.rotation().w(); o[4] = t.rotation().x(); o[5] = t.rotation().y(); o[6] = t.rotation().z();
  }
  return (int)bvs.size();
}

// ---- single-pair queries that also return the reference's own test counters ----
int ref_collide_pair_stats(void *wv, int a, int b, int *stats3) {
  RefWorld *w = (RefWorld *)wv;
  StatCollision cc(w->gc.getModel(w->get(a)), w->gc.getModel(w->get(b)));
  Collision::startRecursion(w->gc.getModel(w->get(a)), w->gc.getModel(w->get(b)), &cc);
  if (stats3) cc.stats(stats3);
  return cc.isCollision() ? 1 : 0;
}
double ref_distance_pair_stats(void *wv, int a, int b, int *stats3) {
  RefWorld *w = (RefWorld *)wv;
  StatDistance dc(w->gc.getModel(w->get(a)), w->gc.getModel(w->get(b)));
  Collision::startRecursion(w->gc.getModel(w->get(a)), w->gc.getModel(w->get(b)), &dc);
  if (stats3) dc.stats(stats3);
  return dc.getMin();
}
double ref_point_stats(void *wv, int body, const double *p, int *stats3) {
  RefWorld *w = (RefWorld *)wv;
  Body *b = w->get(body);
  StatClosest pc(w->gc.getModel(b), b->getTran().inverse() * position(p[0], p[1], p[2]));
  Collision::startRecursion(w->gc.getModel(b), NULL, &pc);
  if (stats3) pc.stats(stats3);
  return pc.getMin();
}

// ---- transf helpers so tests can pin the shim against test/transform_tests.cpp ----
// out7 = (q wxyz, t) of a % b
void ref_transf_compose(const double *qa, const double *ta, const double *qb, const double *tb, double *out7) {
  transf r = makeTransf(qa, ta) % makeTransf(qb, tb);
  out7[0] = r.rotation().w(); out7[1] = r.rotation().x(); out7[2] = r.rotation().y(); out7[3] = r.rotation().z();
  for (int k = 0; k < 3; k++) out7[4 + k] = r.translation()[k];
}
void ref_transf_inverse(const double *q, const double *t, double *out7) {
  transf r = makeTransf(q, t).inverse();
  out7[0] = r.rotation().w(); out7[1] = r.rotation().x(); out7[2] = r.rotation().y(); out7[3] = r.rotation().z();
  for (int k = 0; k < 3; k++) out7[4 + k] = r.translation()[k];
}
void ref_transf_apply(const double *q, const double *t, const double *p, double *out3) {
  vec3 r = makeTransf(q, t) * vec3(p[0], p[1], p[2]);
  for (int k = 0; k < 3; k++) out3[k] = r[k];
}
void ref_transf_rotate(const double *q, const double *t, const double *v, double *out3) {
  vec3 r = makeTransf(q, t).affine() * vec3(v[0], v[1], v[2]);
  for (int k = 0; k < 3; k++) out3[k] = r[k];
}
void ref_transf_axis_angle(double angle, const double *axis, double *out7) {
  transf r = transf::AXIS_ANGLE_ROTATION(angle, vec3(axis[0], axis[1], axis[2]));
  out7[0] = r.rotation().w(); out7[1] = r.rotation().x(); out7[2] = r.rotation().y(); out7[3] = r.rotation().z();
  for (int k = 0; k < 3; k++) out7[4 + k] = r.translation()[k];
}
// transf::COORDINATE (src/transform.cpp:40-50): the frame constructor used by Contact::Contact; out7 = (q wxyz, t), out9 = affine()
void ref_transf_coordinate(const double *o, const double *x, const double *y, double *out7, double *out9) {
  transf r = transf::COORDINATE(position(o[0], o[1], o[2]), vec3(x[0], x[1], x[2]), vec3(y[0], y[1], y[2]));
  out7[0] = r.rotation().w(); out7[1] = r.rotation().x(); out7[2] = r.rotation().y(); out7[3] = r.rotation().z();
  for (int k = 0; k < 3; k++) out7[4 + k] = r.translation()[k];
  if (out9) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) out9[3 * i + j] = r.affine()(i, j);
}
void ref_transf_jacobian(const double *q, const double *t, double *out36) {
  makeTransf(q, t).jacobian(out36);
}
// quaternion -> matrix -> quaternion round trip (test/eigen_tests.cpp)
void ref_quat_roundtrip(const double *q_wxyz, double *mat9_rowmajor, double *q_out_wxyz) {
  Quaternion q(q_wxyz[0], q_wxyz[1], q_wxyz[2], q_wxyz[3]);
  mat3 m = q.toRotationMatrix();
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) mat9_rowmajor[3 * r + c] = m(r, c);
  Quaternion q2(m);
  q_out_wxyz[0] = q2.w(); q_out_wxyz[1] = q2.x(); q_out_wxyz[2] = q2.y(); q_out_wxyz[3] = q2.z();
}

// Soft-finger contact surface fit.  FitParaboloid / RotateParaboloid are the reference's own (include/graspit/FitParabola.h,
// header-only, compiled verbatim above).  Restated around them (softContact.cpp and contact.cpp are Qt/Coin-bound):
// the contact frame of Contact::Contact (contact.cpp:77-90) built with the reference's transf::COORDINATE, and the
// neighbourhood mapping of SoftContact::SoftContact (softContact.cpp:33-45), which applies frame.inverse().affine(),
// i.e. the ROTATION only -- the contact location is not subtracted (the corrected line is commented out upstream).
// out15 = a, b, c, r1, r2, rotAngle, fitRot (row-major 3x3)
void ref_soft_fit(const double *pos, const double *nrm, const double *nghbd_xyz, int npts, double *out15) {
  vec3 normal = vec3(nrm[0], nrm[1], nrm[2]).normalized();
  vec3 tangentX, tangentY;
  if (fabs(normal.dot(vec3(1, 0, 0))) > 1.0 - MACHINE_ZERO) tangentX = normal.cross(vec3(0, 0, 1)).normalized();
  else tangentX = normal.cross(vec3(1, 0, 0)).normalized();
  tangentY = (normal.cross(tangentX)).normalized();
  transf frame = transf::COORDINATE(position(pos[0], pos[1], pos[2]), tangentX, tangentY);
  std::vector<vec3> pts(npts > 0 ? npts : 1);
  for (int i = 0; i < npts; i++) {
    vec3 temp(nghbd_xyz[3 * i], nghbd_xyz[3 * i + 1], nghbd_xyz[3 * i + 2]);
    position posit;
    posit = frame.inverse().affine() * (temp);
    pts[i] = posit;
  }
  double coeffs[3], r1 = 0, r2 = 0, ang = 0;
  mat3 rot = mat3::Identity();
  FitParaboloid(pts.data(), npts, coeffs);
  out15[0] = coeffs[0]; out15[1] = coeffs[1]; out15[2] = coeffs[2];
  RotateParaboloid(coeffs, &r1, &r2, &rot, &ang);
  out15[3] = r1; out15[4] = r2; out15[5] = ang;
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) out15[6 + 3 * r + c] = rot(r, c);
}

// ---------------------------------------------------------------------------------------
// Batched pose evaluation = the reference's per-state caller loop around the backend.
//   hand_ids[nhand]    bodies of the hand in World::findVirtualContacts order:
//                      finger links (chain-major) then the palm (world.cpp:1623-1639);
//                      the same set is the interest list of noCollision (world.cpp:1491-1503)
//   tf7                npose x nhand x 7 doubles (q wxyz, t): world-from-body link poses
//                      from the FK restatement (oracle/port/fk.py)
//   vc_*               PRESET virtual contacts (robot.cpp:435-475): index into hand_ids,
//                      location and normal in the link frame
//   mode               0 = CONTACT_PRESET, 1 = CONTACT_LIVE
// Outputs: legal[npose] (1 = no collision), energy[npose] (0 when illegal,
// searchEnergy.cpp:172-175), optional link_dist[npose x nhand] = bodyToBodyDistance(link, object)
// (only when want_dist != 0; costs one extra distance query per link).
// ---------------------------------------------------------------------------------------
static void evalRange(RefWorld *w, int lo, int hi, int nhand, const int *hand_ids, int object_id,
                      const double *tf7, int nvc, const int *vc_link, const double *vc_loc,
                      const double *vc_normal, int mode, int want_dist,
                      unsigned char *legal, double *energy, double *link_dist) {
  std::vector<Body *> interest;
  for (int i = 0; i < nhand; i++) interest.push_back(w->get(hand_ids[i]));
  Body *object = w->get(object_id);
  for (int p = lo; p < hi; p++) {
    const double *tf = tf7 + (size_t)p * nhand * 7;
    for (int i = 0; i < nhand; i++) {
      transf tr = makeTransf(tf + 7 * i, tf + 7 * i + 4);
      interest[i]->setTranRaw(tr);
      w->gc.setBodyTransform(interest[i], tr);
    }
    // SearchEnergy::legal -> World::noCollision(hand) -> allCollisions(FAST, NULL, &interestList)
    int col = w->gc.allCollisions(CollisionInterface::FAST_COLLISION, NULL, &interest);
    legal[p] = col ? 0 : 1;
    if (link_dist && want_dist) {
      for (int i = 0; i < nhand; i++) {
        position p1, p2;
        link_dist[(size_t)p * nhand + i] = w->gc.bodyToBodyDistance(interest[i], object, p1, p2);
      }
    }
    if (col) { energy[p] = 0.0; continue; }
    double total = 0.0;
    int ncontacts = 0;
    if (mode == 0) {
      for (int c = 0; c < nvc; c++) {
        Body *link = interest[vc_link[c]];
        position loc = link->getTran() * position(vc_loc[3 * c], vc_loc[3 * c + 1], vc_loc[3 * c + 2]);
        vec3 cn = link->getTran().affine() * vec3(vc_normal[3 * c], vc_normal[3 * c + 1], vc_normal[3 * c + 2]);
        position cp; vec3 n_unused;
        w->gc.pointToBodyDistance(object, loc, cp, n_unused);
        vec3 pv = cp - loc;                       // World::pointDistanceToBody
        double dist = pv.norm();
        total += fabs(dist);
        vec3 n = pv.normalized();
        total += (1 - cn.dot(n)) * 100.0 / 2.0;
        ncontacts++;
      }
    } else {
      for (int i = 0; i < nhand; i++) {
        Body *link = interest[i];
        position p1, p2;
        w->gc.bodyToBodyDistance(link, object, p1, p2);          // World::findVirtualContact
        vec3 n = link->getTran() * (p1) - object->getTran() * (p2);
        n = n.normalized();
        n = link->getTran().inverse().affine() * (n);
        // PointContact(loc=p1, normal=n) -> VirtualContact: normal = -n (virtualContact.cpp:92)
        vec3 vcn = -1 * n;
        position loc = link->getTran() * p1;
        vec3 cn = link->getTran().affine() * vcn;
        position cp; vec3 n_unused;
        w->gc.pointToBodyDistance(object, loc, cp, n_unused);
        vec3 pv = cp - loc;
        double dist = pv.norm();
        total += fabs(dist);
        vec3 nn = pv.normalized();
        total += (1 - cn.dot(nn)) * 100.0 / 2.0;
        ncontacts++;
      }
    }
    energy[p] = total / ncontacts;
  }
}

// worlds[nthreads]: identically built worlds, one per thread (the reference backend keeps the
// body transform inside the model, so threads cannot share an instance).
void ref_eval_batch(void **worlds, int nthreads, int npose, int nhand, const int *hand_ids,
                    int object_id, const double *tf7, int nvc, const int *vc_link,
                    const double *vc_loc, const double *vc_normal, int mode, int want_dist,
                    unsigned char *legal, double *energy, double *link_dist) {
  if (nthreads <= 1) {
    evalRange((RefWorld *)worlds[0], 0, npose, nhand, hand_ids, object_id, tf7, nvc, vc_link,
              vc_loc, vc_normal, mode, want_dist, legal, energy, link_dist);
    return;
  }
  std::vector<std::thread> th;
  for (int t = 0; t < nthreads; t++) {
    int lo = (int)((long long)npose * t / nthreads), hi = (int)((long long)npose * (t + 1) / nthreads);
    th.emplace_back(evalRange, (RefWorld *)worlds[t], lo, hi, nhand, hand_ids, object_id, tf7, nvc,
                    vc_link, vc_loc, vc_normal, mode, want_dist, legal, energy, link_dist);
  }
  for (auto &t : th) t.join();
}

int ref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

} // extern "C"
