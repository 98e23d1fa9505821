/* TEST INFRASTRUCTURE ONLY: stand-in for <cv_bridge/cv_bridge.h>, see oracle/standins/standins.h (found through -Istandins) */
#include "standins.h"
