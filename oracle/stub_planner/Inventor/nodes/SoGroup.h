// TEST INFRASTRUCTURE ONLY (oracle stub)
#include <Inventor/nodes/SoSeparator.h>
