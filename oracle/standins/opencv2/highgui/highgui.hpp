/* TEST INFRASTRUCTURE ONLY: stand-in for <opencv2/highgui/highgui.hpp>, see oracle/standins/standins.h (found through -Istandins) */
#include "standins.h"
