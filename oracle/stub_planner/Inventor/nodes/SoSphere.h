// TEST INFRASTRUCTURE ONLY (oracle stub)
#ifndef B2C_ORACLE_SOSPHERE_H
#define B2C_ORACLE_SOSPHERE_H
#include <Inventor/nodes/SoSeparator.h>
struct SoSphere : SoNodeStub { SoFieldStub<float> radius; };
#endif
