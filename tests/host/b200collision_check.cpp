// TEST INFRASTRUCTURE -- A/B harness for the drop-in boundary.  Built by oracle/Makefile into
// oracle/_ref/b200collision_check (it links the verbatim-compiled reference objects, so it lives with them).
//
// Drives `B200Collision` (graspit_b200/host) and the reference's own `GraspitCollision` through the SAME
// abstract `CollisionInterface*`, with the same Body objects, and compares every virtual of the interface
// (collisionInterface.h:102-184).  It also replays the reference's three unit tests
// (test/graspit_collision_tests.cpp:26-140) against B200Collision with their golden values and SLOP.
//
// usage: b200collision_check mug.tri palm.tri [link.tri]     (.tri = int32 count + count*9 float32)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <set>
#include <string>
#include <vector>

#include "graspit/body.h"            // oracle/stub
#include "graspit/bBox.h"
#include "graspit/Collision/Graspit/graspitCollision.h"
#include "b200Collision.h"

static int g_fail = 0, g_checks = 0;
#define CHECK(cond, ...) do { g_checks++; if (!(cond)) { g_fail++; printf("FAIL %s:%d  %s  ", __FILE__, __LINE__, #cond); printf(__VA_ARGS__); printf("\n"); } } while (0)
#define NEAR(a, b, tol) (std::fabs((a) - (b)) <= (tol))
static const double SLOP = 0.001;     // test/graspit_collision_tests.cpp:24

static bool loadTris(const char *path, std::vector<Triangle> &out)
{
  FILE *f = fopen(path, "rb");
  if (!f) { return false; }
  int n = 0;
  if (fread(&n, 4, 1, f) != 1) { fclose(f); return false; }
  std::vector<float> v(9 * (size_t)n);
  size_t got = fread(v.data(), 4, v.size(), f);
  fclose(f);
  if (got != v.size()) { return false; }
  for (int i = 0; i < n; i++) {
    const float *p = &v[9 * (size_t)i];
    out.push_back(Triangle(position(p[0], p[1], p[2]), position(p[3], p[4], p[5]), position(p[6], p[7], p[8])));
  }
  return true;
}

struct Side {
  CollisionInterface *ci;
  Body *mug, *palm, *link;
  Side(CollisionInterface *c, const std::vector<Triangle> &tm, const std::vector<Triangle> &tp,
       const std::vector<Triangle> &tl) : ci(c) {
    mug = new Body("mug"); palm = new Body("palm"); link = new Body("link");
    mug->setTriangles(tm); palm->setTriangles(tp); link->setTriangles(tl);
  }
  void set(Body *b, const transf &t) { b->setTranRaw(t); ci->setBodyTransform(b, t); }
};

static transf randomPose(std::mt19937_64 &rng, double spread)
{
  std::normal_distribution<double> g(0.0, 1.0);
  std::uniform_real_distribution<double> u(-spread, spread);
  double q[4] = {g(rng), g(rng), g(rng), g(rng)};
  double n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  return transf(Quaternion(q[0] / n, q[1] / n, q[2] / n, q[3] / n), vec3(u(rng), u(rng), 105.0 + u(rng)));
}

static std::set<std::pair<std::string, std::string> > pairSet(const CollisionReport &r)
{
  std::set<std::pair<std::string, std::string> > s;
  for (size_t i = 0; i < r.size(); i++) {
    std::string a = r[i].first->getName().latin1(), b = r[i].second->getName().latin1();
    s.insert(a < b ? std::make_pair(a, b) : std::make_pair(b, a));
  }
  return s;
}

int main(int argc, char **argv)
{
  if (argc < 3) { fprintf(stderr, "usage: %s mug.tri palm.tri [link.tri]\n", argv[0]); return 2; }
  std::vector<Triangle> tm, tp, tl;
  if (!loadTris(argv[1], tm) || !loadTris(argv[2], tp)) { fprintf(stderr, "cannot read triangle files\n"); return 2; }
  if (argc > 3) { loadTris(argv[3], tl); } else { tl = tp; }

  B200Collision *b200 = new B200Collision(0);
  if (!b200->valid()) { fprintf(stderr, "no CUDA device\n"); return 3; }
  Side B(b200, tm, tp, tl);
  Side R(new GraspitCollision(), tm, tp, tl);
  Side *sides[2] = {&B, &R};

  // ---- registration (addBody twice -> false; unknown body -> defaults) ----
  for (int s = 0; s < 2; s++) {
    Side &S = *sides[s];
    CHECK(S.ci->addBody(S.mug), "side %d", s);
    CHECK(S.ci->addBody(S.palm), "side %d", s);
    CHECK(!S.ci->addBody(S.palm), "second addBody must fail, side %d", s);
    Body stranger("stranger");
    position p1, p2; vec3 n;
    CHECK(S.ci->bodyToBodyDistance(S.mug, &stranger, p1, p2) == 0, "side %d", s);
    CHECK(S.ci->pointToBodyDistance(&stranger, position(0, 0, 0), p1, n) == 0, "side %d", s);
    CHECK(!S.ci->isActive(&stranger), "side %d", s);
    CHECK(S.ci->isActive(S.mug) && S.ci->isActive(S.mug, S.palm), "side %d", s);
  }

  // ---- the reference's own unit tests, replayed against B200Collision ----
  {
    transf rot = transf::AXIS_ANGLE_ROTATION(10, vec3(1, 0, 0));
    transf translate = transf::TRANSLATION(vec3(1, 2, 3));
    for (int s = 0; s < 2; s++) { sides[s]->set(sides[s]->mug, translate % rot); sides[s]->set(sides[s]->palm, transf::IDENTITY); }
    position p1, p2;
    double dist = B.ci->bodyToBodyDistance(B.mug, B.palm, p1, p2);             // TEST_BODY_TO_BODY_DISTANCE
    CHECK(NEAR(dist, -1, SLOP), "dist %g", dist);
    CHECK(NEAR(p1.x(), -1, SLOP) && NEAR(p1.y(), 3.31021, SLOP) && NEAR(p1.z(), 1.42917, SLOP), "p1 %g %g %g", p1.x(), p1.y(), p1.z());
    CHECK(NEAR(p2.x(), 0, SLOP) && NEAR(p2.y(), 0, SLOP) && NEAR(p2.z(), 0, SLOP), "p2 %g %g %g", p2.x(), p2.y(), p2.z());
    position cp; vec3 cn;
    dist = B.ci->pointToBodyDistance(B.mug, position(10, 11, 12), cp, cn);      // TEST_POINT_TO_BODY_DISTANCE
    CHECK(NEAR(dist, 33.8332, SLOP), "dist %g", dist);
    CHECK(NEAR(cn.x(), 0.951057, SLOP) && NEAR(cn.y(), -0.168109, SLOP) && NEAR(cn.z(), 0.259287, SLOP), "normal %g %g %g", cn.x(), cn.y(), cn.z());
    CHECK(NEAR(cp.x(), 42.1773, SLOP) && NEAR(cp.y(), 5.31235, SLOP) && NEAR(cp.z(), 20.7725, SLOP), "closest %g %g %g", cp.x(), cp.y(), cp.z());
    std::vector<BoundingBox> bv;                                               // TEST_GET_BOUNDING_VOLUMES
    B.ci->getBoundingVolumes(B.mug, 0, &bv);
    CHECK(bv.size() == 1, "%zu boxes", bv.size());
    if (bv.size() == 1) {
      transf t = bv[0].getTran();
      CHECK(NEAR(t.translation().x(), -11.662, SLOP) && NEAR(t.translation().y(), -8.98983, SLOP) && NEAR(t.translation().z(), 0.000190189, SLOP), "mug box t");
      CHECK(NEAR(t.rotation().w(), 0.999999, SLOP) && NEAR(t.rotation().x(), -9.15765e-06, SLOP) && NEAR(t.rotation().y(), -1.47701e-06, SLOP) && NEAR(t.rotation().z(), 0.00157698, SLOP), "mug box q");
    }
    bv.clear();
    B.ci->getBoundingVolumes(B.palm, 0, &bv);
    CHECK(bv.size() == 1, "%zu boxes", bv.size());
    if (bv.size() == 1) {
      transf t = bv[0].getTran();
      CHECK(NEAR(t.translation().x(), 0.013166, SLOP) && NEAR(t.translation().y(), -9.91786, SLOP) && NEAR(t.translation().z(), -38.5818, SLOP), "palm box t");
      CHECK(NEAR(t.rotation().w(), 0.6664499, SLOP) && NEAR(t.rotation().x(), 0.236398, SLOP) && NEAR(t.rotation().y(), -0.66585779, SLOP) && NEAR(t.rotation().z(), -0.23789390, SLOP), "palm box q");
    }
    // deeper levels agree box for box with the reference hierarchy
    for (int depth = 1; depth <= 4; depth++) {
      std::vector<BoundingBox> b1, b2;
      B.ci->getBoundingVolumes(B.mug, depth, &b1); R.ci->getBoundingVolumes(R.mug, depth, &b2);
      CHECK(b1.size() == b2.size(), "depth %d: %zu vs %zu", depth, b1.size(), b2.size());
      for (size_t i = 0; i < b1.size() && i < b2.size(); i++) {
        vec3 dt = b1[i].getTran().translation() - b2[i].getTran().translation();
        vec3 dh = b1[i].halfSize - b2[i].halfSize;
        CHECK(dt.norm() < 1e-9 && dh.norm() < 1e-9, "depth %d box %zu", depth, i);
      }
    }
  }

  // ---- random poses: every query of the interface, side by side ----
  std::mt19937_64 rng(1234);
  int ncoll = 0, nfree = 0, ncontactPoses = 0;
  for (int it = 0; it < 300; it++) {
    transf tp_ = randomPose(rng, it % 2 ? 60.0 : 150.0);
    transf tmug = it % 3 == 0 ? randomPose(rng, 5.0) : transf::IDENTITY;
    for (int s = 0; s < 2; s++) { sides[s]->set(sides[s]->palm, tp_); sides[s]->set(sides[s]->mug, tmug); }
    CollisionReport cb, cr;
    int nb = B.ci->allCollisions(CollisionInterface::ALL_COLLISIONS, &cb, NULL);
    int nr = R.ci->allCollisions(CollisionInterface::ALL_COLLISIONS, &cr, NULL);
    CHECK(nb == nr && pairSet(cb) == pairSet(cr), "pose %d: %d vs %d collisions", it, nb, nr);
    int fb = B.ci->allCollisions(CollisionInterface::FAST_COLLISION, NULL, NULL);
    CHECK(fb == (nr > 0 ? 1 : 0), "pose %d fast %d", it, fb);
    position b1, b2, r1, r2;
    double db = B.ci->bodyToBodyDistance(B.palm, B.mug, b1, b2);
    double dr = R.ci->bodyToBodyDistance(R.palm, R.mug, r1, r2);
    CHECK((dr < 0) == (db < 0), "pose %d: dist %g vs %g", it, db, dr);
    if (dr >= 0 && db >= 0) {
      nfree++;
      CHECK(NEAR(db, dr, 1e-4 * std::max(1.0, dr)), "pose %d: dist %.9g vs %.9g", it, db, dr);
      // closest points are unique only up to ties; the separation they realise must match the distance
      position wb1 = tp_ * (tp_ * b1), wb2 = tmug * (tmug * b2);     // undo the reference's double inverse
      CHECK(NEAR((wb1 - wb2).norm(), dr, 1e-4 * std::max(1.0, dr)), "pose %d: |p1-p2| %.9g vs %.9g", it, (wb1 - wb2).norm(), dr);
    } else {
      ncoll++;
      CHECK((b1 - r1).norm() < 1e-6 && (b2 - r2).norm() < 1e-6, "pose %d: collision-case outputs", it);
    }
    position q(tp_.translation().x(), tp_.translation().y(), tp_.translation().z()), cpb, cpr;
    vec3 cnb, cnr;
    double pb = B.ci->pointToBodyDistance(B.mug, q, cpb, cnb), pr = R.ci->pointToBodyDistance(R.mug, q, cpr, cnr);
    CHECK(NEAR(pb, pr, 1e-6 * std::max(1.0, pr)), "pose %d: point dist %.9g vs %.9g", it, pb, pr);
    CHECK(NEAR((cpb - q).norm(), pr, 1e-6 * std::max(1.0, pr)), "pose %d: closest point", it);
  }
  CHECK(ncoll > 20 && nfree > 20, "coverage: %d colliding, %d free", ncoll, nfree);

  // ---- contacts: palm pushed to 0.1 mm from the mug along the closest-point direction ----
  for (int it = 0; it < 40; it++) {
    transf tp_ = randomPose(rng, 150.0);
    for (int s = 0; s < 2; s++) { sides[s]->set(sides[s]->palm, tp_); sides[s]->set(sides[s]->mug, transf::IDENTITY); }
    position r1, r2;
    double dr = R.ci->bodyToBodyDistance(R.palm, R.mug, r1, r2);
    if (dr <= 0.2) { continue; }
    position w1 = tp_ * (tp_ * r1), w2 = r2;       // mug at identity
    vec3 dir = (w2 - w1) / (w2 - w1).norm();
    transf moved(tp_.rotation(), tp_.translation() + dir * (dr - 0.1));
    for (int s = 0; s < 2; s++) { sides[s]->set(sides[s]->palm, moved); }
    ContactReport kb, kr;
    int nb = B.ci->contact(&kb, 0.2, B.palm, B.mug), nr = R.ci->contact(&kr, 0.2, R.palm, R.mug);
    CHECK(nr > 0 && nb > 0, "contact pose %d: %d vs %d contacts", it, nb, nr);
    // every reference contact has a B200 contact within the 3 mm duplicate radius and vice versa
    for (size_t i = 0; i < kr.size(); i++) {
      double best = 1e30;
      for (size_t j = 0; j < kb.size(); j++) { best = std::min(best, (kr[i].b1_pos - kb[j].b1_pos).norm()); }
      CHECK(best <= 3.0, "contact pose %d: reference contact %zu unmatched (%g)", it, i, best);
    }
    for (size_t j = 0; j < kb.size(); j++) {
      double best = 1e30;
      for (size_t i = 0; i < kr.size(); i++) { best = std::min(best, (kr[i].b1_pos - kb[j].b1_pos).norm()); }
      CHECK(best <= 3.0, "contact pose %d: B200 contact %zu unmatched (%g)", it, j, best);
      CHECK(NEAR(kb[j].b1_normal.norm(), 1.0, 1e-9), "unit normal");
    }
    CollisionReport ab, ar;
    int pb = B.ci->allContacts(&ab, 0.2, NULL), pr = R.ci->allContacts(&ar, 0.2, NULL);
    CHECK(pb == pr && pairSet(ab) == pairSet(ar), "allContacts pairs %d vs %d", pb, pr);
    if (!kr.empty()) {
      ncontactPoses++;
      Neighborhood hb, hr;
      B.ci->bodyRegion(B.mug, kr[0].b2_pos, kr[0].b2_normal, 3.5, &hb);
      R.ci->bodyRegion(R.mug, kr[0].b2_pos, kr[0].b2_normal, 3.5, &hr);
      CHECK(hb.size() == hr.size(), "region %zu vs %zu", hb.size(), hr.size());
    }
  }
  CHECK(ncontactPoses >= 20, "contact coverage %d", ncontactPoses);

  // ---- activation, clones, removal ----
  {
    transf rot = transf::AXIS_ANGLE_ROTATION(10, vec3(1, 0, 0));
    for (int s = 0; s < 2; s++) {
      Side &S = *sides[s];
      S.set(S.mug, transf::TRANSLATION(vec3(1, 2, 3)) % rot); S.set(S.palm, transf::IDENTITY);   // interpenetrating
      CHECK(S.ci->allCollisions(CollisionInterface::FAST_COLLISION, NULL, NULL) == 1, "side %d", s);
      S.ci->activatePair(S.mug, S.palm, false);
      CHECK(!S.ci->isActive(S.mug, S.palm) && !S.ci->isActive(S.palm, S.mug), "side %d", s);
      CHECK(S.ci->allCollisions(CollisionInterface::FAST_COLLISION, NULL, NULL) == 0, "side %d", s);
      S.ci->activatePair(S.palm, S.mug, true);
      CHECK(S.ci->isActive(S.mug, S.palm), "side %d", s);
      S.ci->activateBody(S.palm, false);
      CHECK(!S.ci->isActive(S.palm) && !S.ci->isActive(S.mug, S.palm), "side %d", s);
      CHECK(S.ci->allCollisions(CollisionInterface::ALL_COLLISIONS, NULL, NULL) == 0, "side %d", s);
      S.ci->activateBody(S.palm, true);
      // interest list: only pairs touching a listed body
      S.ci->addBody(S.link);
      S.set(S.link, transf::TRANSLATION(vec3(0, 0, 1000)));
      std::vector<Body *> interest(1, S.link);
      CollisionReport rep;
      CHECK(S.ci->allCollisions(CollisionInterface::ALL_COLLISIONS, &rep, &interest) == 0, "side %d", s);
      interest[0] = S.mug;
      CHECK(S.ci->allCollisions(CollisionInterface::ALL_COLLISIONS, &rep, &interest) == 1 && rep.size() == 1, "side %d", s);
      // a clone shares geometry, moves independently
      Body *clone = new Body("clone");
      clone->setTriangles(std::vector<Triangle>());
      S.ci->cloneBody(clone, S.palm);
      S.set(clone, transf::TRANSLATION(vec3(0, 0, 500)));
      position p1, p2;
      CHECK(S.ci->bodyToBodyDistance(clone, S.mug, p1, p2) > 100, "side %d", s);
      S.set(clone, transf::IDENTITY);
      CHECK(S.ci->bodyToBodyDistance(clone, S.mug, p1, p2) == -1, "side %d", s);
      S.ci->removeBody(clone);
      CHECK(S.ci->bodyToBodyDistance(clone, S.mug, p1, p2) == 0, "removed clone is unknown, side %d", s);
      delete clone;
      // updateBodyGeometry keeps the body's identity and disabled pairs
      S.ci->activatePair(S.link, S.mug, false);
      S.link->setTriangles(S.palm == NULL ? std::vector<Triangle>() : std::vector<Triangle>());
      std::vector<Triangle> newGeom;
      S.palm->getGeometryTriangles(&newGeom);
      S.link->setTriangles(newGeom);
      CHECK(S.ci->updateBodyGeometry(S.link), "side %d", s);
      S.set(S.link, transf::IDENTITY);
      CHECK(!S.ci->isActive(S.link, S.mug), "side %d", s);
      rep.clear();
      interest[0] = S.link;
      int n = S.ci->allCollisions(CollisionInterface::ALL_COLLISIONS, &rep, &interest);
      CHECK(n == 1 && pairSet(rep).count(std::make_pair(std::string("link"), std::string("palm"))), "side %d: %d", s, n);
      S.ci->removeBody(S.link);
    }
  }

  printf("%s: %d checks, %d failed\n", g_fail ? "FAILED" : "PASSED", g_checks, g_fail);
  return g_fail ? 1 : 0;
}
