/* TEST INFRASTRUCTURE ONLY: stand-in for <geometry_msgs/Twist.h>, see oracle/standins/standins.h (found through -Istandins) */
#include "standins.h"
