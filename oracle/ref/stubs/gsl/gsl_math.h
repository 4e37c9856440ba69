#ifndef G2_STUB_GSL_MATH_H
#define G2_STUB_GSL_MATH_H
#include <math.h>
#endif
