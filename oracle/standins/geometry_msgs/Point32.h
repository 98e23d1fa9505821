/* TEST INFRASTRUCTURE ONLY: stand-in for <geometry_msgs/Point32.h>, see oracle/standins/standins.h (found through -Istandins) */
#include "standins.h"
