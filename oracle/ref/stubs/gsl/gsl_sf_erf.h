#ifndef G2_STUB_GSL_SF_ERF_H
#define G2_STUB_GSL_SF_ERF_H
#include <math.h>
static inline double gsl_sf_erfc(double x) { return erfc(x); }
#endif
