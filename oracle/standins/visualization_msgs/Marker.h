/* TEST INFRASTRUCTURE ONLY: stand-in for <visualization_msgs/Marker.h>, see oracle/standins/standins.h (found through -Istandins) */
#include "standins.h"
