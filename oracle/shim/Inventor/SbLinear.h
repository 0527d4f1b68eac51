// TEST INFRASTRUCTURE ONLY (oracle shim): tiny stand-ins for Coin3D's SbVec3f / SbRotation,
// which include/graspit/matvec3D.h names.  Coin3D is absent from this image.
#ifndef B2C_ORACLE_SBLINEAR_H
#define B2C_ORACLE_SBLINEAR_H
struct SbVec3f {
  float v[3];
  SbVec3f() { v[0] = v[1] = v[2] = 0.f; }
  SbVec3f(float x, float y, float z) { v[0] = x; v[1] = y; v[2] = z; }
  float operator[](int i) const { return v[i]; }
  float &operator[](int i) { return v[i]; }
  void getValue(float &x, float &y, float &z) const { x = v[0]; y = v[1]; z = v[2]; }
};
struct SbRotation {
  float q[4];  // x y z w
  SbRotation() { q[0] = q[1] = q[2] = 0.f; q[3] = 1.f; }
  SbRotation(float x, float y, float z, float w) { q[0] = x; q[1] = y; q[2] = z; q[3] = w; }
  void getValue(float &x, float &y, float &z, float &w) const { x = q[0]; y = q[1]; z = q[2]; w = q[3]; }
};
#endif
