// TEST INFRASTRUCTURE ONLY (oracle shim): stand-in for Coin3D's SoTransform node.
#ifndef B2C_ORACLE_SOTRANSFORM_H
#define B2C_ORACLE_SOTRANSFORM_H
#include <Inventor/SbLinear.h>
struct SoSFRotationShim {
  SbRotation r;
  void setValue(const SbRotation &v) { r = v; }
  void setValue(const SbVec3f &, float) {}          // axis / angle form (visual markers only)
  const SbRotation &getValue() const { return r; }
};
struct SoSFVec3fShim {
  SbVec3f t;
  void setValue(const SbVec3f &v) { t = v; }
  const SbVec3f &getValue() const { return t; }
};
struct SoTransform {
  SoSFRotationShim rotation;
  SoSFVec3fShim translation;
  void ref() {}          // Coin reference counting (Joint::Joint, joint.h:206): nothing to count here
  void unref() {}
};
#endif
