/* TEST INFRASTRUCTURE ONLY: stand-in for <jackal_nav/JackalPose.h>, see oracle/standins/standins.h (found through -Istandins) */
#include "standins.h"
