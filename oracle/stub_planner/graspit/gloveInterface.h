// TEST INFRASTRUCTURE ONLY (oracle stub): CyberGlove input (EGPlanner::processInput); never used by the harness.
#ifndef B2C_ORACLE_STUB_GLOVE_H
#define B2C_ORACLE_STUB_GLOVE_H
class GloveInterface {
  public:
    bool isDOFControlled(int) const { return false; }
    float getDOFValue(int) const { return 0.f; }
};
#endif
