/* Stand-in for GSL's rng header (GSL is not installed).  Only get_random_number()'s table uses
 * it (NOTREERND coincident-particle randomisation, FORCETEST sampling); oracle builds use
 * -DNOTREERND and tests avoid coincident particles, so no result depends on this generator. */
#ifndef G2_STUB_GSL_RNG_H
#define G2_STUB_GSL_RNG_H
#include <stddef.h>
typedef struct { unsigned long long s; } gsl_rng;
typedef struct { int dummy; } gsl_rng_type;
extern const gsl_rng_type *gsl_rng_ranlxd1;
gsl_rng *gsl_rng_alloc(const gsl_rng_type *t);
void gsl_rng_set(gsl_rng *r, unsigned long seed);
double gsl_rng_uniform(gsl_rng *r);
void *gsl_rng_state(const gsl_rng *r);
size_t gsl_rng_size(const gsl_rng *r);
#endif
