/* TEST INFRASTRUCTURE ONLY: stand-in for <popt.h>, see oracle/standins/standins.h (found through -Istandins) */
#include "standins.h"
