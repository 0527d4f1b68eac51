// Soft-finger contacts (SFCL): neighbourhood merge, analytic surface fit, relative curvature, mattress-model contact
// ellipse and the 20-edge friction ellipsoid -> boundary wrenches.  Host-side closed forms (a few hundred flops per
// contact) behind the C ABI; the neighbourhoods themselves come from the device (region_kernel via b2c_region).
// Reference: src/contactSetting.cpp:83-191, src/contact/softContact.cpp:22-125,186-358, include/graspit/FitParabola.h.
#include "contacts_host.h"

#include <cmath>
#include <vector>

namespace b2c {

namespace {

struct V { double x, y, z; };
inline V sub(V a, V b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V add(V a, V b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V mul(V a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(V a, V b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V cross(V a, V b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V a) { return std::sqrt(dot(a, a)); }

// rotation matrix (row-major) of a quaternion after normalisation: transf::set(Quaternion, vec3) (transform.h:47-49)
void quat_matrix(const double qin[4], double R[3][3]) {
  double n = std::sqrt(qin[0] * qin[0] + qin[1] * qin[1] + qin[2] * qin[2] + qin[3] * qin[3]);
  double w = qin[0], x = qin[1], y = qin[2], z = qin[3];
  if (n > 0) { w /= n; x /= n; y /= n; z /= n; }
  const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
  const double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y,
               tzz = tz * z;
  R[0][0] = 1 - (tyy + tzz); R[0][1] = txy - twz; R[0][2] = txz + twy;
  R[1][0] = txy + twz; R[1][1] = 1 - (txx + tzz); R[1][2] = tyz - twx;
  R[2][0] = txz - twy; R[2][1] = tyz + twx; R[2][2] = 1 - (txx + tyy);
}
inline V mat_vec(const double R[3][3], V v) {
  return {R[0][0] * v.x + R[0][1] * v.y + R[0][2] * v.z, R[1][0] * v.x + R[1][1] * v.y + R[1][2] * v.z,
          R[2][0] * v.x + R[2][1] * v.y + R[2][2] * v.z};
}
inline V transf_apply(const double q[4], const double t[3], V p) {
  double R[3][3];
  quat_matrix(q, R);
  return add(mat_vec(R, p), {t[0], t[1], t[2]});
}

// closed-form 3x3 inverse by cofactors of the first column (the fixed-size inverse of the linear-algebra library the
// reference uses for mat3::inverse)
void inverse3(const double m[3][3], double r[3][3]) {
  auto cof = [&](int i, int j) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    return m[i1][j1] * m[i2][j2] - m[i1][j2] * m[i2][j1];
  };
  double c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
  double det = c0 * m[0][0] + c1 * m[1][0] + c2 * m[2][0];
  double invdet = 1.0 / det;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) r[j][i] = cof(i, j) * invdet;
}

constexpr double FIT_TINY = 1.0e-6;   // FitParabola.h:22

// FitParaboloid (FitParabola.h:25-85): least squares z = a x^2 + b y^2 + c xy through the normal equations
void fit_paraboloid(const std::vector<V> &p, double coeffs[3]) {
  double A[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}}, Ai[3][3];
  V z = {0, 0, 0};
  for (const V &q : p) {
    z = add(z, {q.x * q.x * q.z, q.y * q.y * q.z, q.x * q.y * q.z});
    const double x4 = q.x * q.x * q.x * q.x, x2y2 = q.x * q.x * q.y * q.y, x3y = q.x * q.x * q.x * q.y;
    const double y4 = q.y * q.y * q.y * q.y, xy3 = q.x * q.y * q.y * q.y;
    A[0][0] += x4;   A[1][0] += x2y2; A[2][0] += x3y;
    A[0][1] += x2y2; A[1][1] += y4;   A[2][1] += xy3;
    A[0][2] += x3y;  A[1][2] += xy3;  A[2][2] += x2y2;
  }
  inverse3(A, Ai);
  V x = mat_vec(Ai, z);
  coeffs[0] = x.x; coeffs[1] = x.y; coeffs[2] = x.z;
  for (int i = 0; i < 3; i++)
    if (std::fabs(coeffs[i]) < FIT_TINY) coeffs[i] = 0.0;   // almost planar
}

// RotateParaboloid (FitParabola.h:87-166): rotation about z that removes the xy term; radii of curvature, -1 = flat
void rotate_paraboloid(const double c[3], double *R1, double *R2, double *angle) {
  if (std::fabs(c[2]) > FIT_TINY) {
    double theta = (c[0] != c[1]) ? std::atan(c[2] / (c[1] - c[0])) / 2.0 : M_PI / 2;
    *angle = theta;
    double st = std::sin(theta), ct = std::cos(theta);
    double t1 = (c[0] * ct * ct + c[1] * st * st - c[2] * st * ct) * 2;
    double t2 = (c[0] * st * st + c[1] * ct * ct + c[2] * st * ct) * 2;
    *R1 = std::fabs(t1) > FIT_TINY ? 1 / t1 : -1.0;
    *R2 = std::fabs(t2) > FIT_TINY ? 1 / t2 : -1.0;
  } else {
    *R1 = std::fabs(c[0]) > FIT_TINY ? 1 / (2 * c[0]) : -1.0;
    *R2 = std::fabs(c[1]) > FIT_TINY ? 1 / (2 * c[1]) : -1.0;
    *angle = 0;
  }
}

// main curvature direction of one side in world coordinates (SoftContact::CalcRelPhi, softContact.cpp:208-221):
// body rotation * contact-frame rotation * fitRot * (1,0,0); fitRot's first column is (cos, sin, 0)
V curvature_dir_world(const double tangents[6], V normal_z, double rot_angle, const double body_q[4]) {
  V tx = {tangents[0], tangents[1], tangents[2]}, ty = {tangents[3], tangents[4], tangents[5]};
  double cx = rot_angle != 0 ? std::cos(rot_angle) : 1.0, cy = rot_angle != 0 ? std::sin(rot_angle) : 0.0;
  // frame.affine() has columns (tangentX, tangentY, z); the third component of fitRot's column is 0
  V in_body = {tx.x * cx + ty.x * cy + normal_z.x * 0.0, tx.y * cx + ty.y * cy + normal_z.y * 0.0,
               tx.z * cx + ty.z * cy + normal_z.z * 0.0};
  double R[3][3];
  quat_matrix(body_q, R);
  return mat_vec(R, in_body);
}

} // namespace

double soft_neighborhood_radius(double youngs1, double youngs2) {
  // contactSetting.cpp:129-133; the "hack" condition there (rad <= 3 && rad >= 10) can never hold
  double rad = std::pow(1 / (youngs1 > youngs2 ? youngs1 : youngs2), 0.333) * 1000.0 * 0.4;
  if (rad <= 3.0 && rad >= 10.0) rad = 5.0;
  return rad;
}

int soft_contact_merge(std::vector<b2c_contact> &contacts, std::vector<std::vector<double>> &n1,
                       std::vector<std::vector<double>> &n2, const double q1[4], const double t1[3], const double q2[4],
                       const double t2[3]) {
  auto overlap = [](const std::vector<double> &a, const std::vector<double> &b) {
    for (size_t i = 0; i + 2 < a.size(); i += 3)
      for (size_t j = 0; j + 2 < b.size(); j += 3)
        if (a[i] == b[j] && a[i + 1] == b[j + 1] && a[i + 2] == b[j + 2]) return true;
    return false;
  };
  auto merge = [](std::vector<double> &a, const std::vector<double> &b) {
    for (size_t j = 0; j + 2 < b.size(); j += 3) {
      bool present = false;
      for (size_t i = 0; i + 2 < a.size(); i += 3)
        if (a[i] == b[j] && a[i + 1] == b[j + 1] && a[i + 2] == b[j + 2]) { present = true; break; }
      if (!present) a.insert(a.end(), b.begin() + j, b.begin() + j + 3);
    }
  };
  auto cdist = [&](const b2c_contact &c) {   // contactDistance (contactSetting.cpp:39-47)
    V a = transf_apply(q1, t1, {c.b1_pos[0], c.b1_pos[1], c.b1_pos[2]});
    V b = transf_apply(q2, t2, {c.b2_pos[0], c.b2_pos[1], c.b2_pos[2]});
    return norm(sub(a, b));
  };
  int merges = 0;
  bool performed = true;
  while (performed) {          // mergeSoftNeighborhoods (contactSetting.cpp:141-184): restart after every merge
    performed = false;
    for (size_t r = 0; r < contacts.size() && !performed; r++)
      for (size_t o = 0; o < contacts.size(); o++) {
        if (o == r) continue;
        if (!overlap(n1[r], n1[o])) continue;    // only the neighbourhoods on body 1 are compared
        size_t keep = r, drop = o;
        if (!(cdist(contacts[r]) < cdist(contacts[o]))) { keep = o; drop = r; }
        merge(n1[keep], n1[drop]);
        merge(n2[keep], n2[drop]);
        contacts.erase(contacts.begin() + drop);
        n1.erase(n1.begin() + drop);
        n2.erase(n2.begin() + drop);
        performed = true;
        merges++;
        break;
      }
  }
  return merges;
}

int soft_contact_fit(const b2c_contact *c, int n, int side, const double *nghbd_xyz, const int *offsets, b2c_soft_fit *fit) {
  std::vector<double> frames(7 * (size_t)(n > 0 ? n : 1));
  contact_frames(c, n, side, frames.data(), nullptr);
  for (int i = 0; i < n; i++) {
    // SoftContact::SoftContact (softContact.cpp:33-45): frame.inverse().affine() * p -- the rotation of the inverse
    // frame only; the contact location is NOT subtracted (reference behaviour, kept for parity)
    const double *q = &frames[7 * (size_t)i];
    double n2 = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
    double qi[4] = {q[0], -q[1], -q[2], -q[3]};
    if (n2 > 0) for (double &v : qi) v /= n2; else for (double &v : qi) v = 0;
    double Ri[3][3];
    quat_matrix(qi, Ri);
    std::vector<V> pts;
    for (int k = offsets[i]; k < offsets[i + 1]; k++)
      pts.push_back(mat_vec(Ri, {nghbd_xyz[3 * (size_t)k], nghbd_xyz[3 * (size_t)k + 1], nghbd_xyz[3 * (size_t)k + 2]}));
    double co[3];
    fit_paraboloid(pts, co);
    fit[i].a = co[0]; fit[i].b = co[1]; fit[i].c = co[2];
    rotate_paraboloid(co, &fit[i].r1, &fit[i].r2, &fit[i].rot_angle);
  }
  return n;
}

int soft_contact_wrenches(const b2c_contact *c, int n, int side, const b2c_soft_fit *fit_this, const b2c_soft_fit *fit_mate,
                          const double q_this[4], const double q_mate[4], double youngs, double cof, const double cog[3],
                          double max_radius, double *wrench120, double *ellipse6) {
  std::vector<double> frames(7 * (size_t)(n > 0 ? n : 1)), tang(6 * (size_t)(n > 0 ? n : 1)), tang_m(6 * (size_t)(n > 0 ? n : 1));
  contact_frames(c, n, side, frames.data(), tang.data());
  contact_frames(c, n, side == 1 ? 2 : 1, frames.data(), tang_m.data());
  for (int i = 0; i < n; i++) {
    const double *pos = side == 1 ? c[i].b1_pos : c[i].b2_pos;
    const double *nrm = side == 1 ? c[i].b1_normal : c[i].b2_normal;
    V normal = {nrm[0], nrm[1], nrm[2]};
    { double l2 = dot(normal, normal); if (l2 > 0) normal = mul(normal, 1.0 / std::sqrt(l2)); }
    // CalcRelPhi (softContact.cpp:204-232): angle between the main curvature directions of the two bodies, world frame
    V R11 = curvature_dir_world(&tang[6 * (size_t)i], normal, fit_this[i].rot_angle, q_this);
    V R12 = curvature_dir_world(&tang_m[6 * (size_t)i], normal, fit_mate[i].rot_angle, q_mate);
    double relPhi = std::acos(dot(R11, R12) / (norm(R11) * norm(R12)));
    // CalcRprimes (softContact.cpp:238-314): relative radii of curvature (Johnson, Contact Mechanics, ch. 4)
    double w = fit_this[i].r1 < 0.0 ? 0.0 : 1 / fit_this[i].r1;
    double x = fit_this[i].r2 < 0.0 ? 0.0 : 1 / fit_this[i].r2;
    double y = fit_mate[i].r1 < 0.0 ? 0.0 : 1 / fit_mate[i].r1;
    double z = fit_mate[i].r2 < 0.0 ? 0.0 : 1 / fit_mate[i].r2;
    double ApB = 0.5 * (w + x + y + z);
    double AmB = 0.5 * std::sqrt((w - x) * (w - x) + (y - z) * (y - z) + 2 * std::cos(2 * relPhi) * (w - x) * (y - z));
    double r1p = 0, r2p = 0;          // the values the constructor leaves when the curvature is rejected
    if (!(AmB > ApB)) {
      r1p = (ApB + AmB != 0) ? 1 / (ApB + AmB) : -1.0;
      r2p = (ApB == AmB) ? -1.0 : 1 / (ApB - AmB);
    }
    // CalcContact_Mattress(5) (softContact.cpp:322-358): contact ellipse for a 5 N normal force, 3 mm mattress
    if (r1p < 0) r1p = 20;
    if (r2p < 0) r2p = 20;
    const double h = 0.003, nForce = 5;
    double delta = std::sqrt(nForce * h / (youngs * M_PI * std::sqrt(r1p * 0.001 * r2p * 0.001)));
    double major = std::sqrt(2 * delta * r1p * 0.001), minor = std::sqrt(2 * delta * r2p * 0.001);
    double torquedivN = 8 * std::sqrt(major * minor) / 15;
    if (ellipse6) {
      double *e = ellipse6 + 6 * (size_t)i;
      e[0] = major; e[1] = minor; e[2] = relPhi; e[3] = r1p; e[4] = r2p; e[5] = torquedivN;
    }
    // setUpFrictionEdges (softContact.cpp:69-117): 3-D friction ellipsoid, latitudes {1,5,8,5,1}
    const int ndirs[5] = {1, 5, 8, 5, 1};
    const double phi[5] = {M_PI_2, M_PI_2 * 0.50, 0.0, -M_PI_2 * 0.50, -M_PI_2};
    const double eccen[3] = {1, 1, torquedivN * 1000};
    std::vector<double> edges;
    friction_ellipsoid_edges(5, ndirs, phi, eccen, edges);
    // Contact::computeWrenches (contact.cpp:153-165; the in-plane moment variant of SoftContact is disabled upstream)
    V tx = {tang[6 * (size_t)i], tang[6 * (size_t)i + 1], tang[6 * (size_t)i + 2]};
    V ty = {tang[6 * (size_t)i + 3], tang[6 * (size_t)i + 4], tang[6 * (size_t)i + 5]};
    V radius = sub({pos[0], pos[1], pos[2]}, {cog[0], cog[1], cog[2]});
    for (int k = 0; k < 20; k++) {
      const double *e = &edges[6 * k];
      V force = add(add(normal, mul(tx, cof * e[0])), mul(ty, cof * e[1]));
      V torque = mul(add(cross(radius, force), mul(normal, cof * e[5])), 1.0 / max_radius);
      double *wr = wrench120 + 120 * (size_t)i + 6 * k;
      wr[0] = force.x; wr[1] = force.y; wr[2] = force.z; wr[3] = torque.x; wr[4] = torque.y; wr[5] = torque.z;
    }
  }
  return n;
}

} // namespace b2c
