/* TEST INFRASTRUCTURE ONLY: stand-in for <image_transport/image_transport.h>, see oracle/standins/standins.h (found through -Istandins) */
#include "standins.h"
