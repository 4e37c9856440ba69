/* oracle/ref/stubs.c — TEST INFRASTRUCTURE ONLY.
 * Single-rank implementations of the MPI-1 calls the reference hot-path closure makes, a tiny
 * uniform RNG behind the GSL names, and an O(n log n) arbitrary-length complex DFT (Bluestein)
 * behind FFTW-2's fftw_one(), so that the unmodified reference sources link without MPI/GSL/FFTW. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <sys/time.h>
#include "mpi.h"
#include "gsl/gsl_rng.h"
#include "fftw_stub.h"

int MPI_Init(int *argc, char ***argv) { (void) argc; (void) argv; return 0; }
int MPI_Finalize(void) { return 0; }
int MPI_Abort(MPI_Comm c, int err) { (void) c; fprintf(stderr, "g2ref: MPI_Abort(%d)\n", err); abort(); return 0; }
int MPI_Comm_rank(MPI_Comm c, int *r) { (void) c; *r = 0; return 0; }
int MPI_Comm_size(MPI_Comm c, int *s) { (void) c; *s = 1; return 0; }
int MPI_Barrier(MPI_Comm c) { (void) c; return 0; }
double MPI_Wtime(void) { struct timeval tv; gettimeofday(&tv, NULL); return tv.tv_sec + 1e-6 * tv.tv_usec; }
int MPI_Bcast(void *b, int n, MPI_Datatype t, int root, MPI_Comm c) { (void) b; (void) n; (void) t; (void) root; (void) c; return 0; }
static void cp(const void *s, void *r, size_t bytes) { if(s != r && bytes) memmove(r, s, bytes); }
int MPI_Allgather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, MPI_Comm c)
{ (void) nr; (void) tr; (void) c; cp(s, r, (size_t) ns * ts); return 0; }
int MPI_Allgatherv(const void *s, int ns, MPI_Datatype ts, void *r, const int *nr, const int *displ, MPI_Datatype tr, MPI_Comm c)
{ (void) nr; (void) c; cp(s, (char *) r + (size_t) displ[0] * tr, (size_t) ns * ts); return 0; }
int MPI_Gather(const void *s, int ns, MPI_Datatype ts, void *r, int nr, MPI_Datatype tr, int root, MPI_Comm c)
{ (void) nr; (void) tr; (void) root; (void) c; cp(s, r, (size_t) ns * ts); return 0; }
int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, MPI_Comm c)
{ (void) op; (void) c; cp(s, r, (size_t) n * t); return 0; }
int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype t, MPI_Op op, int root, MPI_Comm c)
{ (void) op; (void) root; (void) c; cp(s, r, (size_t) n * t); return 0; }
static int no_p2p(const char *w) { fprintf(stderr, "g2ref: %s called in a single-rank oracle\n", w); abort(); return 0; }
int MPI_Sendrecv(const void *s, int ns, MPI_Datatype ts, int dest, int stag, void *r, int nr, MPI_Datatype tr, int src, int rtag, MPI_Comm c, MPI_Status *st)
{ (void) s; (void) ns; (void) ts; (void) dest; (void) stag; (void) r; (void) nr; (void) tr; (void) src; (void) rtag; (void) c; (void) st; return no_p2p("MPI_Sendrecv"); }
int MPI_Ssend(const void *s, int n, MPI_Datatype t, int dest, int tag, MPI_Comm c)
{ (void) s; (void) n; (void) t; (void) dest; (void) tag; (void) c; return no_p2p("MPI_Ssend"); }
int MPI_Send(const void *s, int n, MPI_Datatype t, int dest, int tag, MPI_Comm c)
{ (void) s; (void) n; (void) t; (void) dest; (void) tag; (void) c; return no_p2p("MPI_Send"); }
int MPI_Recv(void *r, int n, MPI_Datatype t, int src, int tag, MPI_Comm c, MPI_Status *st)
{ (void) r; (void) n; (void) t; (void) src; (void) tag; (void) c; (void) st; return no_p2p("MPI_Recv"); }

/* ---- GSL names ---- */
static const gsl_rng_type ranlx = { 0 };
const gsl_rng_type *gsl_rng_ranlxd1 = &ranlx;
gsl_rng *gsl_rng_alloc(const gsl_rng_type *t) { gsl_rng *r = malloc(sizeof(gsl_rng)); (void) t; r->s = 88172645463325252ULL; return r; }
void gsl_rng_set(gsl_rng *r, unsigned long seed) { r->s = seed * 2685821657736338717ULL + 1442695040888963407ULL; }
double gsl_rng_uniform(gsl_rng *r)
{ r->s ^= r->s << 13; r->s ^= r->s >> 7; r->s ^= r->s << 17; return (double) (r->s >> 11) / 9007199254740992.0; }
void *gsl_rng_state(const gsl_rng *r) { return (void *) r; }
size_t gsl_rng_size(const gsl_rng *r) { (void) r; return sizeof(gsl_rng); }

/* ---- FFTW-2 1-D complex DFT of arbitrary n: out[k] = sum_j in[j] exp(dir*2*pi*i*j*k/n) ---- */
typedef struct { double re, im; } cpx;
static void fft_pow2(cpx *a, int n, int dir)
{
  int i, j, len;
  for(i = 1, j = 0; i < n; i++)
    {
      int bit = n >> 1;
      for(; j & bit; bit >>= 1)
	j ^= bit;
      j ^= bit;
      if(i < j) { cpx t = a[i]; a[i] = a[j]; a[j] = t; }
    }
  for(len = 2; len <= n; len <<= 1)
    {
      int half = len >> 1, k;
      for(k = 0; k < half; k++)
	{
	  double ang = dir * 2.0 * M_PI * k / len;
	  double wr = cos(ang), wi = sin(ang);
	  for(i = k; i < n; i += len)
	    {
	      cpx u = a[i], v = a[i + half], t;
	      t.re = v.re * wr - v.im * wi;
	      t.im = v.re * wi + v.im * wr;
	      a[i].re = u.re + t.re; a[i].im = u.im + t.im;
	      a[i + half].re = u.re - t.re; a[i + half].im = u.im - t.im;
	    }
	}
    }
}

fftw_plan fftw_create_plan(int n, int dir, int flags)
{ fftw_plan p = malloc(sizeof(*p)); (void) flags; p->n = n; p->dir = dir; return p; }
void fftw_destroy_plan(fftw_plan p) { free(p); }

void fftw_one(fftw_plan p, fftw_complex *in, fftw_complex *out)
{
  int n = p->n, dir = p->dir, m = 1, k;
  cpx *w, *a, *b;
  while(m < 2 * n - 1)
    m <<= 1;
  w = malloc(sizeof(cpx) * n);
  a = calloc(m, sizeof(cpx));
  b = calloc(m, sizeof(cpx));
  for(k = 0; k < n; k++)
    {				/* chirp exp(dir*i*pi*k^2/n), k^2 reduced mod 2n to keep the phase accurate */
      long long k2 = ((long long) k * k) % (2LL * n);
      double ang = dir * M_PI * (double) k2 / n;
      w[k].re = cos(ang);
      w[k].im = sin(ang);
    }
  for(k = 0; k < n; k++)
    {
      a[k].re = in[k].re * w[k].re - in[k].im * w[k].im;
      a[k].im = in[k].re * w[k].im + in[k].im * w[k].re;
    }
  b[0].re = w[0].re; b[0].im = -w[0].im;
  for(k = 1; k < n; k++)
    {
      b[k].re = b[m - k].re = w[k].re;
      b[k].im = b[m - k].im = -w[k].im;
    }
  fft_pow2(a, m, -1);
  fft_pow2(b, m, -1);
  for(k = 0; k < m; k++)
    {
      cpx t;
      t.re = a[k].re * b[k].re - a[k].im * b[k].im;
      t.im = a[k].re * b[k].im + a[k].im * b[k].re;
      a[k] = t;
    }
  fft_pow2(a, m, +1);
  for(k = 0; k < n; k++)
    {
      double re = a[k].re / m, im = a[k].im / m;
      out[k].re = re * w[k].re - im * w[k].im;
      out[k].im = re * w[k].im + im * w[k].re;
    }
  free(b);
  free(a);
  free(w);
}

/* ---- FFTW-2 MPI real 3-D transforms, single rank (pm_periodic.c:65-73, 465, 531) ---- */
#include "rfftw_mpi_stub.h"
rfftwnd_mpi_plan rfftw3d_mpi_create_plan(MPI_Comm comm, int nx, int ny, int nz, int dir, int flags)
{
  rfftwnd_mpi_plan p = malloc(sizeof(*p));
  (void) comm; (void) flags;
  if((nx & (nx - 1)) || (ny & (ny - 1)) || (nz & (nz - 1)))
    { fprintf(stderr, "g2ref: the rfftwnd_mpi stand-in needs power-of-two mesh sizes\n"); abort(); }
  p->nx = nx; p->ny = ny; p->nz = nz; p->dir = dir;
  return p;
}
void rfftwnd_mpi_destroy_plan(rfftwnd_mpi_plan p) { free(p); }
void rfftwnd_mpi_local_sizes(rfftwnd_mpi_plan p, int *local_nx, int *local_x_start, int *local_ny_after_transpose,
			     int *local_y_start_after_transpose, int *total_local_size)
{
  *local_nx = p->nx; *local_x_start = 0; *local_ny_after_transpose = p->ny; *local_y_start_after_transpose = 0;
  *total_local_size = p->nx * p->ny * 2 * (p->nz / 2 + 1);
}

/* in-place 3-D complex DFT of c[x][y][z], unnormalised */
static void fft3(cpx *c, int nx, int ny, int nz, int dir)
{
  int x, y, z, nmax = nx > ny ? nx : ny;
  cpx *line = malloc(sizeof(cpx) * nmax);
  for(x = 0; x < nx; x++)
    for(y = 0; y < ny; y++)
      fft_pow2(c + ((size_t) x * ny + y) * nz, nz, dir);
  for(x = 0; x < nx; x++)
    for(z = 0; z < nz; z++)
      {
	for(y = 0; y < ny; y++) line[y] = c[((size_t) x * ny + y) * nz + z];
	fft_pow2(line, ny, dir);
	for(y = 0; y < ny; y++) c[((size_t) x * ny + y) * nz + z] = line[y];
      }
  for(y = 0; y < ny; y++)
    for(z = 0; z < nz; z++)
      {
	for(x = 0; x < nx; x++) line[x] = c[((size_t) x * ny + y) * nz + z];
	fft_pow2(line, nx, dir);
	for(x = 0; x < nx; x++) c[((size_t) x * ny + y) * nz + z] = line[x];
      }
  free(line);
}

void rfftwnd_mpi(rfftwnd_mpi_plan p, int n_fields, fftw_real *data, fftw_real *work, fftwnd_mpi_output_order order)
{
  const int nx = p->nx, ny = p->ny, nz = p->nz, nzh = nz / 2 + 1, nz2 = 2 * nzh;
  int x, y, z;
  cpx *c = malloc(sizeof(cpx) * (size_t) nx * ny * nz);
  fftw_complex *cd = (fftw_complex *) data;
  (void) work;
  if(n_fields != 1 || order != FFTW_TRANSPOSED_ORDER)
    { fprintf(stderr, "g2ref: rfftwnd_mpi stand-in: unsupported call\n"); abort(); }
  if(p->dir == FFTW_REAL_TO_COMPLEX)
    {
      for(x = 0; x < nx; x++)
	for(y = 0; y < ny; y++)
	  for(z = 0; z < nz; z++)
	    {
	      c[((size_t) x * ny + y) * nz + z].re = data[((size_t) x * ny + y) * nz2 + z];
	      c[((size_t) x * ny + y) * nz + z].im = 0.0;
	    }
      fft3(c, nx, ny, nz, -1);
      for(y = 0; y < ny; y++)
	for(x = 0; x < nx; x++)
	  for(z = 0; z < nzh; z++)
	    {
	      cd[((size_t) y * nx + x) * nzh + z].re = c[((size_t) x * ny + y) * nz + z].re;
	      cd[((size_t) y * nx + x) * nzh + z].im = c[((size_t) x * ny + y) * nz + z].im;
	    }
    }
  else
    {
      /* rebuild the full spectrum from the stored half (Hermitian symmetry of a real field) */
      for(x = 0; x < nx; x++)
	for(y = 0; y < ny; y++)
	  for(z = 0; z < nz; z++)
	    {
	      cpx v;
	      if(z < nzh)
		{
		  v.re = cd[((size_t) y * nx + x) * nzh + z].re;
		  v.im = cd[((size_t) y * nx + x) * nzh + z].im;
		}
	      else
		{
		  int xm = (nx - x) % nx, ym = (ny - y) % ny, zm = nz - z;
		  v.re = cd[((size_t) ym * nx + xm) * nzh + zm].re;
		  v.im = -cd[((size_t) ym * nx + xm) * nzh + zm].im;
		}
	      c[((size_t) x * ny + y) * nz + z] = v;
	    }
      fft3(c, nx, ny, nz, +1);
      for(x = 0; x < nx; x++)
	for(y = 0; y < ny; y++)
	  for(z = 0; z < nz; z++)
	    data[((size_t) x * ny + y) * nz2 + z] = c[((size_t) x * ny + y) * nz + z].re;
    }
  free(c);
}
