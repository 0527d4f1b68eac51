// TEST INFRASTRUCTURE ONLY (oracle stub): Coin's idle sensor (EGPlanner's single-threaded mode); never scheduled here.
#ifndef B2C_ORACLE_SOIDLESENSOR_H
#define B2C_ORACLE_SOIDLESENSOR_H
class SoSensor {
  public:
    void schedule() {}
    void unschedule() {}
};
typedef void SoSensorCB(void *, SoSensor *);
class SoIdleSensor : public SoSensor {
  public:
    SoIdleSensor(SoSensorCB * = 0, void * = 0) {}
    void schedule() {}
    void unschedule() {}
    bool isScheduled() const { return false; }
};
#endif
