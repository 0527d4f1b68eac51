// TEST INFRASTRUCTURE ONLY (oracle stub): stand-ins for the Coin3D nodes the reference's searchState.cpp builds its
// visual markers from.  Nothing is drawn; the calls only have to compile and do nothing.
#ifndef B2C_ORACLE_SOSEPARATOR_H
#define B2C_ORACLE_SOSEPARATOR_H
#include <Inventor/SbLinear.h>
struct SoNodeStub {
  void ref() {}
  void unref() {}
};
template <typename T> struct SoFieldStub {
  T v;
  SoFieldStub &operator=(const T &x) { v = x; return *this; }
  void setValue(const T &x) { v = x; }
  template <typename A, typename B> void setValue(const A &, const B &) {}
  template <typename A, typename B, typename C> void setValue(const A &, const B &, const C &) {}
};
struct SoSeparator : SoNodeStub {
  void addChild(void *) {}
  int findChild(const void *) const { return -1; }
  void removeChild(int) {}
};
typedef SoSeparator SoGroup;
#endif
