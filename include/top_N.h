/* top_N.h -- forwarding header so reference callers that include "top_N.h" compile
 * unchanged against libmaveric_b200.so; all declarations live in one place. */
#include "maveric_slam_compat.h"
