/* TEST INFRASTRUCTURE ONLY: stand-in for <ros/ros.h>, see oracle/standins/standins.h (found through -Istandins) */
#include "standins.h"
