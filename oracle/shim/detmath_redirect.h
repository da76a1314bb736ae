// oracle/shim/detmath_redirect.h -- force-included (gcc -include) into the reference translation units of the pin build: their
// calls to sin / cos / atan2 go to the deterministic implementations of include/ualm_detmath.h, the ones the oracle and the CUDA
// path use (glibc and CUDA libm differ in the last bit).  sqrt, floor, fabs stay the IEEE library ones.
#pragma once
#include <cmath>
#include <math.h>
#include "ualm_detmath.h"
namespace ualm_redirect {
inline double sin(double x) { return ualm_sin(x); }
inline double cos(double x) { return ualm_cos(x); }
inline double atan2(double y, double x) { return ualm_atan2(y, x); }
}
#define sin(x) ualm_redirect::sin(x)
#define cos(x) ualm_redirect::cos(x)
#define atan2(y, x) ualm_redirect::atan2(y, x)
