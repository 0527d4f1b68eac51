// TEST INFRASTRUCTURE ONLY (oracle stub)
#ifndef B2C_ORACLE_SOMATERIAL_H
#define B2C_ORACLE_SOMATERIAL_H
#include <Inventor/nodes/SoSeparator.h>
struct SbColor { float r, g, b; SbColor(double R = 0, double G = 0, double B = 0) : r((float)R), g((float)G), b((float)B) {} };
struct SoMaterial : SoNodeStub { SoFieldStub<SbColor> diffuseColor, ambientColor; SoFieldStub<float> transparency; };
#endif
