/* TEST INFRASTRUCTURE ONLY: stand-in for <sensor_msgs/LaserScan.h>, see oracle/standins/standins.h (found through -Istandins) */
#include "standins.h"
