// TEST INFRASTRUCTURE ONLY -- symbols the verbatim-compiled reference objects expect from translation units
// that cannot compile here (src/gws.cpp, src/matvec.cpp); used by binaries that link the reference objects
// without oracle/ref_driver.cpp (which defines the same symbols for libgraspit_ref.so).
#include <QMutex>
#include "graspit/matvec3D.h"
#include <Inventor/SbLinear.h>

QMutex qhull_mutex;                                                   // src/gws.cpp
SbRotation QuaterniontoSbRotation(Quaternion q) {                      // src/matvec.cpp
  return SbRotation((float)q.x(), (float)q.y(), (float)q.z(), (float)q.w());
}
SbVec3f toSbVec3f(vec3 v) { return SbVec3f((float)v.x(), (float)v.y(), (float)v.z()); }
