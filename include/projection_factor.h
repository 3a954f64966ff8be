/* projection_factor.h -- forwarding header so reference callers that include "projection_factor.h" compile
 * unchanged against libmaveric_b200.so; all declarations live in one place. */
#include "maveric_slam_compat.h"
