/* TEST INFRASTRUCTURE ONLY: stand-in for <sensor_msgs/Joy.h>, see oracle/standins/standins.h (found through -Istandins) */
#include "standins.h"
