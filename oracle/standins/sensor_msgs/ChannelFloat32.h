/* TEST INFRASTRUCTURE ONLY: stand-in for <sensor_msgs/ChannelFloat32.h>, see oracle/standins/standins.h (found through -Istandins) */
#include "standins.h"
