// TEST INFRASTRUCTURE ONLY (oracle stub): the planner only asks isA("...") of these hand types
#include "graspit/robot.h"
