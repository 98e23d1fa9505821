/* TEST INFRASTRUCTURE ONLY: stand-in for <opencv2/calib3d/calib3d.hpp>, see oracle/standins/standins.h (found through -Istandins) */
#include "standins.h"
