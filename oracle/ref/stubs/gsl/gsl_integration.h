/* oracle/ref/stubs/gsl/gsl_integration.h — TEST INFRASTRUCTURE.  driftfac.c builds its look-up tables with
 * gsl_integration_qag() (comoving integration only; every shipped configuration runs with ComovingIntegrationOn 0,
 * where the tables are never read).  A composite Simpson rule stands in for QAG. */
#ifndef G2REF_GSL_INTEGRATION_H
#define G2REF_GSL_INTEGRATION_H
#include <stdlib.h>
#include "gsl_math.h"
#ifndef GSL_INTEG_GAUSS41
#define GSL_INTEG_GAUSS41 4
#endif
typedef struct { int n; } gsl_integration_workspace;
#ifndef G2REF_GSL_FUNCTION
#define G2REF_GSL_FUNCTION
typedef struct { double (*function) (double x, void *params); void *params; } gsl_function;
#endif
static inline gsl_integration_workspace *gsl_integration_workspace_alloc(size_t n)
{ gsl_integration_workspace *w = malloc(sizeof(*w)); w->n = (int) n; return w; }
static inline void gsl_integration_workspace_free(gsl_integration_workspace * w) { free(w); }
static inline int gsl_integration_qag(const gsl_function * f, double a, double b, double epsabs, double epsrel, size_t limit, int key,
				      gsl_integration_workspace * w, double *result, double *abserr)
{
  const int n = 2000;
  double h = (b - a) / n, s = f->function(a, f->params) + f->function(b, f->params);
  int i;
  for(i = 1; i < n; i++)
    s += (i & 1 ? 4.0 : 2.0) * f->function(a + i * h, f->params);
  *result = s * h / 3.0;
  *abserr = 0;
  (void) epsabs; (void) epsrel; (void) limit; (void) key; (void) w;
  return 0;
}
#endif
