// B200Collision: see b200Collision.h.  Host C++ only -- all geometry runs in libb2c.so (CUDA, sm_100a).
#include "b200Collision.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

#include "graspit/body.h"
#include "graspit/bBox.h"
#include "graspit/triangle.h"

#include "b2c.h"

// the reference reports errors with DBGA(...) on stderr and carries on (graspitCollision.cpp:106,118,...)
#define B200_DBGA(msg) fprintf(stderr, "%s\n", msg)

B200Collision::B200Collision(int device) : mCtx(NULL)
{
  if (device < 0) {
    const char *env = getenv("B2C_DEVICE");
    device = env ? atoi(env) : 0;
  }
  if (b2c_create(device, &mCtx) != B2C_OK) {
    B200_DBGA("B200COL: no usable CUDA device; the backend has no CPU path and stays disabled");
    mCtx = NULL;
  }
}

B200Collision::~B200Collision()
{
  if (mCtx) { b2c_destroy(mCtx); }
}

int B200Collision::getId(const Body *body) const
{
  std::map<const Body *, int>::const_iterator it = mBodyIds.find(body);
  return it == mBodyIds.end() ? -1 : it->second;
}

/*! Pulls the triangle soup from the body exactly as GraspitCollision::addBody does
  (graspitCollision.cpp:210-217) and hands it to the engine as fp32: the coordinates GraspIt! receives from
  Coin are fp32 values widened to double (body.cpp:1225-1249), so the narrowing is exact. */
int B200Collision::uploadGeometry(Body *body)
{
  std::vector<Triangle> triangles;
  body->getGeometryTriangles(&triangles);
  std::vector<float> tri9(9 * triangles.size() + 9);
  for (size_t i = 0; i < triangles.size(); i++) {
    const Triangle &t = triangles[i];
    float *o = &tri9[9 * i];
    o[0] = (float)t.v1.x(); o[1] = (float)t.v1.y(); o[2] = (float)t.v1.z();
    o[3] = (float)t.v2.x(); o[4] = (float)t.v2.y(); o[5] = (float)t.v2.z();
    o[6] = (float)t.v3.x(); o[7] = (float)t.v3.y(); o[8] = (float)t.v3.z();
  }
  int meshId = -1;
  if (b2c_mesh_create(mCtx, &tri9[0], (int)triangles.size(), &meshId) != B2C_OK) {
    B200_DBGA(b2c_last_error(mCtx));
    return -1;
  }
  return meshId;
}

void B200Collision::releaseMesh(int meshId)
{
  std::map<int, int>::iterator it = mMeshUsers.find(meshId);
  if (it == mMeshUsers.end()) { return; }
  if (--it->second <= 0) {
    b2c_mesh_destroy(mCtx, meshId);
    mMeshUsers.erase(it);
  }
}

bool B200Collision::addBody(Body *body, bool)
{
  if (!mCtx) { B200_DBGA("B200COL: backend disabled"); return false; }
  if (getId(body) >= 0) {
    B200_DBGA("GCOL: body already present");
    return false;
  }
  int meshId = uploadGeometry(body);
  if (meshId < 0) { return false; }
  int id = -1;
  if (b2c_body_create(mCtx, meshId, &id) != B2C_OK) {
    B200_DBGA(b2c_last_error(mCtx));
    b2c_mesh_destroy(mCtx, meshId);
    return false;
  }
  // bodies belong to the thread that adds them (CollisionModel(getThreadId()), graspitCollision.cpp:207)
  b2c_body_set_thread(mCtx, id, getThreadId());
  mBodyIds[body] = id;
  mBodies[id] = body;
  mMeshIds[body] = meshId;
  mMeshUsers[meshId] = 1;
  return true;
}

bool B200Collision::updateBodyGeometry(Body *body, bool)
{
  int id = getId(body);
  if (id < 0) {
    B200_DBGA("GCOL: body not found for geometry update");
    return false;
  }
  int meshId = uploadGeometry(body);
  if (meshId < 0) { return false; }
  // disabled pairs, activity and transform stay intact (collisionInterface.h:104)
  if (b2c_body_set_mesh(mCtx, id, meshId) != B2C_OK) {
    B200_DBGA(b2c_last_error(mCtx));
    b2c_mesh_destroy(mCtx, meshId);
    return false;
  }
  releaseMesh(mMeshIds[body]);
  mMeshIds[body] = meshId;
  mMeshUsers[meshId] = 1;
  return true;
}

void B200Collision::removeBody(Body *body)
{
  int id = getId(body);
  if (id < 0) {
    B200_DBGA("GCOL: model not found");
    return;
  }
  b2c_body_remove(mCtx, id);
  releaseMesh(mMeshIds[body]);
  mMeshIds.erase(body);
  mBodyIds.erase(body);
  mBodies.erase(id);
}

/*! A clone is a second body on the original's mesh: it shares the hierarchy and owns its transform
  (CollisionModel::cloneModel, collisionModel.cpp:466-474).  Unlike the reference the shared mesh is
  reference-counted, so deleting the original first is safe. */
void B200Collision::cloneBody(Body *clone, const Body *original)
{
  int orig = getId(original);
  if (orig < 0) {
    B200_DBGA("GCOL: model not found");
    return;
  }
  if (getId(clone) >= 0) {
    B200_DBGA("GCOL: clone already in collision detection system");
    return;
  }
  int meshId = mMeshIds[original];
  int id = -1;
  if (b2c_body_create(mCtx, meshId, &id) != B2C_OK) {
    B200_DBGA(b2c_last_error(mCtx));
    return;
  }
  b2c_body_set_thread(mCtx, id, getThreadId());
  mBodyIds[clone] = id;
  mBodies[id] = clone;
  mMeshIds[clone] = meshId;
  mMeshUsers[meshId]++;
}

void B200Collision::setBodyTransform(Body *body, const transf &t)
{
  int id = getId(body);
  if (id < 0) {
    B200_DBGA("GCOL: model not found");
    return;
  }
  const Quaternion &r = t.rotation();
  const vec3 &v = t.translation();
  double q[4] = {r.w(), r.x(), r.y(), r.z()};
  double tr[3] = {v.x(), v.y(), v.z()};
  b2c_body_set_transform(mCtx, id, q, tr);
}

void B200Collision::activateBody(const Body *body, bool active)
{
  int id = getId(body);
  if (id < 0) {
    B200_DBGA("GCOL: model not found");
    return;
  }
  b2c_body_set_active(mCtx, id, active ? 1 : 0);
}

void B200Collision::activatePair(const Body *body1, const Body *body2, bool active)
{
  int a = getId(body1), b = getId(body2);
  if (a < 0 || b < 0) {
    B200_DBGA("GCOL: model not found");
    return;
  }
  if (a == b) { B200_DBGA("GCOL Warning: insertion collision pair is actually one body"); }
  b2c_pair_set_active(mCtx, a, b, active ? 1 : 0);
}

bool B200Collision::isActive(const Body *body1, const Body *body2)
{
  int a = getId(body1);
  if (a < 0) {
    B200_DBGA("GCOL: model not found");
    return false;
  }
  int b = -1;
  if (body2) {
    b = getId(body2);
    if (b < 0) {
      B200_DBGA("GCOL: model not found");
      return false;
    }
  }
  int active = 0;
  b2c_pair_is_active(mCtx, a, b, &active);
  return active != 0;
}

void B200Collision::convertInterestList(const std::vector<Body *> *inList, std::vector<int> *out) const
{
  for (size_t i = 0; i < inList->size(); i++) {
    int id = getId((*inList)[i]);
    if (id < 0) {
      B200_DBGA("GCOL: interest list model not found");
    } else {
      out->push_back(id);
    }
  }
}

int B200Collision::allCollisions(DetectionType type, CollisionReport *report,
                                 const std::vector<Body *> *interestList)
{
  if (!mCtx) { return 0; }
  std::vector<int> interest;
  if (interestList) { convertInterestList(interestList, &interest); }
  b2c_set_thread(mCtx, getThreadId());
  int nb = (int)mBodyIds.size();
  std::vector<int> pairs(std::max(2, nb * (nb - 1)));
  int n = 0;
  int rc = b2c_all_collisions(mCtx, type == FAST_COLLISION ? 1 : 0,
                              interestList ? (interest.empty() ? pairs.data() : interest.data()) : NULL,
                              (int)interest.size(), pairs.data(), (int)pairs.size() / 2, &n);
  if (rc != B2C_OK) {
    B200_DBGA(b2c_last_error(mCtx));
    return 0;
  }
  // FAST_COLLISION returns at the first hit and leaves the report alone (graspitCollision.cpp:351-353)
  if (type == FAST_COLLISION) { return n; }
  if (report) {
    for (int i = 0; i < n; i++) {
      report->push_back(CollisionData(mBodies[pairs[2 * i]], mBodies[pairs[2 * i + 1]]));
    }
  }
  return n;
}

static ContactData toContactData(const b2c_contact &c)
{
  return ContactData(position(c.b1_pos[0], c.b1_pos[1], c.b1_pos[2]),
                     position(c.b2_pos[0], c.b2_pos[1], c.b2_pos[2]),
                     vec3(c.b1_normal[0], c.b1_normal[1], c.b1_normal[2]),
                     vec3(c.b2_normal[0], c.b2_normal[1], c.b2_normal[2]), c.distSq);
}

int B200Collision::allContacts(CollisionReport *report, double threshold,
                               const std::vector<Body *> *interestList)
{
  if (!mCtx) { return 0; }
  std::vector<int> interest;
  if (interestList) { convertInterestList(interestList, &interest); }
  b2c_set_thread(mCtx, getThreadId());
  int nb = (int)mBodyIds.size();
  int capPairs = std::max(1, nb * (nb - 1) / 2);
  std::vector<int> pairIds(2 * capPairs), counts(capPairs);
  std::vector<b2c_contact> contacts(1024);
  int np = 0, nc = 0;
  for (int attempt = 0; attempt < 2; attempt++) {
    int rc = b2c_all_contacts(mCtx, threshold,
                              interestList ? (interest.empty() ? pairIds.data() : interest.data()) : NULL,
                              (int)interest.size(), pairIds.data(), counts.data(), capPairs, &np,
                              contacts.data(), (int)contacts.size(), &nc);
    if (rc != B2C_OK) {
      B200_DBGA(b2c_last_error(mCtx));
      return 0;
    }
    if (nc <= (int)contacts.size()) { break; }
    contacts.resize(nc);
  }
  // one CollisionData per contacting pair, appended to the caller's report (graspitCollision.cpp:388-392);
  // the return value is the number of contacting pairs
  int at = 0;
  for (int i = 0; i < np; i++) {
    if (report) {
      report->push_back(CollisionData(mBodies[pairIds[2 * i]], mBodies[pairIds[2 * i + 1]]));
      for (int k = 0; k < counts[i]; k++) { report->back().contacts.push_back(toContactData(contacts[at + k])); }
    }
    at += counts[i];
  }
  return np;
}

int B200Collision::contact(ContactReport *report, double threshold,
                           const Body *body1, const Body *body2)
{
  int a = getId(body1), b = getId(body2);
  if (a < 0 || b < 0) {
    B200_DBGA("GCOL: model not found");
    return 0;
  }
  std::vector<b2c_contact> contacts(256);
  int n = 0;
  for (int attempt = 0; attempt < 2; attempt++) {
    if (b2c_contacts(mCtx, a, b, threshold, 0, contacts.data(), (int)contacts.size(), &n) != B2C_OK) {
      B200_DBGA(b2c_last_error(mCtx));
      return 0;
    }
    if (n <= (int)contacts.size()) { break; }
    contacts.resize(n);
  }
  // contact() overwrites the report (graspitCollision.cpp:422)
  report->clear();
  for (int i = 0; i < n; i++) { report->push_back(toContactData(contacts[i])); }
  return (int)report->size();
}

double B200Collision::pointToBodyDistance(const Body *body1, position point,
                                          position &closestPoint, vec3 &closestNormal)
{
  int id = getId(body1);
  if (id < 0) {
    B200_DBGA("GCOL: model not found");
    return 0;
  }
  double p[3] = {point.x(), point.y(), point.z()}, d = 0, c[3], n[3];
  if (b2c_point_distance(mCtx, id, p, &d, c, n) != B2C_OK) {
    B200_DBGA(b2c_last_error(mCtx));
    return 0;
  }
  closestPoint = position(c[0], c[1], c[2]);
  closestNormal = vec3(n[0], n[1], n[2]);
  return d;
}

double B200Collision::bodyToBodyDistance(const Body *body1, const Body *body2,
                                         position &p1, position &p2)
{
  int a = getId(body1), b = getId(body2);
  if (a < 0 || b < 0) {
    B200_DBGA("GCOL: model not found");
    return 0;
  }
  double d = 0, q1[3], q2[3];
  if (b2c_body_distance(mCtx, a, b, &d, q1, q2, NULL, NULL) != B2C_OK) {
    B200_DBGA(b2c_last_error(mCtx));
    return 0;
  }
  p1 = position(q1[0], q1[1], q1[2]);
  p2 = position(q2[0], q2[1], q2[2]);
  return d;
}

void B200Collision::bodyRegion(const Body *body, position point, vec3 normal,
                               double radius, Neighborhood *neighborhood)
{
  int id = getId(body);
  if (id < 0) {
    B200_DBGA("GCOL: model not found");
    return;
  }
  double p[3] = {point.x(), point.y(), point.z()}, nrm[3] = {normal.x(), normal.y(), normal.z()};
  std::vector<double> xyz(3 * 512);
  int n = 0;
  for (int attempt = 0; attempt < 2; attempt++) {
    if (b2c_region(mCtx, id, p, nrm, radius, xyz.data(), (int)xyz.size() / 3, &n) != B2C_OK) {
      B200_DBGA(b2c_last_error(mCtx));
      return;
    }
    if (3 * n <= (int)xyz.size()) { break; }
    xyz.resize(3 * n);
  }
  // bodyRegion overwrites the neighbourhood (graspitCollision.cpp:493)
  neighborhood->clear();
  for (int i = 0; i < n; i++) { neighborhood->push_back(position(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2])); }
}

void B200Collision::getBoundingVolumes(const Body *body, int depth, std::vector<BoundingBox> *bvs)
{
  int id = getId(body);
  if (id < 0) {
    B200_DBGA("GCOL: model not found");
    return;
  }
  std::vector<b2c_obb> boxes(64);
  int n = 0;
  for (int attempt = 0; attempt < 2; attempt++) {
    if (b2c_bounding_volumes(mCtx, id, depth, boxes.data(), (int)boxes.size(), &n) != B2C_OK) {
      B200_DBGA(b2c_last_error(mCtx));
      return;
    }
    if (n <= (int)boxes.size()) { break; }
    boxes.resize(n);
  }
  for (int i = 0; i < n; i++) {
    const b2c_obb &o = boxes[i];
    transf t(Quaternion(o.q_wxyz[0], o.q_wxyz[1], o.q_wxyz[2], o.q_wxyz[3]), vec3(o.t[0], o.t[1], o.t[2]));
    bvs->push_back(BoundingBox(t, vec3(o.half[0], o.half[1], o.half[2])));
  }
}
