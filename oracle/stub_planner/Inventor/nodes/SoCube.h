// TEST INFRASTRUCTURE ONLY (oracle stub)
#ifndef B2C_ORACLE_SOCUBE_H
#define B2C_ORACLE_SOCUBE_H
#include <Inventor/nodes/SoSeparator.h>
struct SoCube : SoNodeStub { SoFieldStub<float> width, height, depth; };
#endif
