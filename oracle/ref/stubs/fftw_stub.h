/* Stand-in for FFTW-2's 1-D complex interface (dfftw.h / sfftw.h / fftw.h are not installed).
 * fftw_one() is implemented in stubs.c with a Bluestein chirp-z transform so that the reference's
 * performConvolution (ngravs_core.c:72-159) runs unmodified for n = 589682. */
#ifndef G2_STUB_FFTW_H
#define G2_STUB_FFTW_H
typedef double fftw_real;
typedef struct { fftw_real re, im; } fftw_complex;
typedef struct g2_fftw_plan_s { int n; int dir; } *fftw_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_ESTIMATE 0
#define FFTW_MEASURE 1
fftw_plan fftw_create_plan(int n, int dir, int flags);
void fftw_destroy_plan(fftw_plan p);
void fftw_one(fftw_plan p, fftw_complex *in, fftw_complex *out);
#endif
