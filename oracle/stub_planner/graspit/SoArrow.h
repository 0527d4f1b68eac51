// TEST INFRASTRUCTURE ONLY (oracle stub)
#ifndef B2C_ORACLE_SOARROW_H
#define B2C_ORACLE_SOARROW_H
#include <Inventor/nodes/SoSeparator.h>
struct SoArrow : SoNodeStub { SoFieldStub<float> height, cylRadius, coneRadius, coneHeight; };
#endif
