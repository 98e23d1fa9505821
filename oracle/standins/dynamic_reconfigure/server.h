/* TEST INFRASTRUCTURE ONLY: stand-in for <dynamic_reconfigure/server.h>, see oracle/standins/standins.h (found through -Istandins) */
#include "standins.h"
