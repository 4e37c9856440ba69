// g2_pm.cu — the periodic particle-mesh long-range force of the TreePM split: pmforce_periodic (pm_periodic.c:204-790), called by
// long_range_force (longrange.c:56-141) before gravity_tree (accel.c:36-46).  It is the complement of the short-range walk of
// g2_walk.cu.
//
// Reference algorithm, per ORDERED species pair (nA sources, nB targets):  CIC mass assignment of species nA on a PMGRID^3 mesh
// (:285-316) -> forward FFT (:465) -> multiply by GreensFxns[nA][nB](k) * (-exp(-k^2 asmth2)) / (sinc sinc sinc)^4, i.e. Gaussian
// long-range filter and twofold CIC deconvolution (:468-523), k = 0 zeroed (:525) -> inverse FFT (:531) -> 4-point finite
// differences (:726-737) -> trilinear (CIC) interpolation to the particles of species nB (:739-781), GravPM += .
//
// Here: D forward transforms (one per source species), then per TARGET species the filtered spectra of all sources are summed
// in k-space (the transform is linear), ONE inverse transform, and differencing + interpolation fused in one kernel that reads
// the potential mesh directly (no force meshes): D + D transforms instead of the reference's 2 D^2.  All mesh arithmetic is FP64
// like the reference's (fftw_real = double with DOUBLEPRECISION_FFTW); the float position scaling `to_slab_fac * Pos` is
// reproduced in FLOAT.  The 3-D FFT itself is cuFFT (a plain library transform); everything else is hand-written and HBM-bound:
//   deposit   32 B read per particle + 8 atomic FP64 adds (random, L2-resident mesh)
//   filter    (D + 1) x 16 B per k-space cell per target species
//   gather    32 B read per particle + 56 distinct potential reads (8 corners x (1 + 2 x 3 neighbours), L2) + 12 B written
#include <cufft.h>
#include "g2_common.cuh"

struct PMArgs
{
  const G2PRec *__restrict__ rec;
  int n, N;
  float to_slab_fac;		// FLOAT to_slab_fac = PMGRID / All.BoxSize (pm_periodic.c:46, 123)
  unsigned int t2g_packed;
};

// pm_periodic.c:286-310: slab index and offset inside the cell, in the reference's FLOAT arithmetic
__device__ __forceinline__ void pm_cell(const PMArgs &A, float pos, int &slab, double &d)
{
  const float u = __fmul_rn(A.to_slab_fac, pos);
  slab = (int) u;
  if(slab >= A.N)
    slab = A.N - 1;
  d = (double) __fsub_rn(u, (float) slab);
}

// one pass over the particles: every particle adds its mass to the mesh of its own species (mesh g at rho + g * N^3)
__global__ void __launch_bounds__(256) pm_deposit_kernel(PMArgs A, double *__restrict__ rho_all)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= A.n)
    return;
  const G2PRec p = A.rec[i];
  double *__restrict__ rho = rho_all + (size_t) ((A.t2g_packed >> (4 * min(max(p.type, 0), 5))) & 7u) * A.N * A.N * A.N;
  int sx, sy, sz;
  double dx, dy, dz;
  pm_cell(A, p.x, sx, dx);
  pm_cell(A, p.y, sy, dy);
  pm_cell(A, p.z, sz, dz);
  const int N = A.N;
  const int sxx = sx + 1 == N ? 0 : sx + 1, syy = sy + 1 == N ? 0 : sy + 1, szz = sz + 1 == N ? 0 : sz + 1;
  const double m = (double) p.m;
#define G2_RHO(X, Y, Z) (rho + ((size_t) (X) * N + (Y)) * N + (Z))
  // products left to right as in pm_periodic.c:312-319
  atomicAdd(G2_RHO(sx, sy, sz), m * (1.0 - dx) * (1.0 - dy) * (1.0 - dz));
  atomicAdd(G2_RHO(sx, syy, sz), m * (1.0 - dx) * dy * (1.0 - dz));
  atomicAdd(G2_RHO(sx, sy, szz), m * (1.0 - dx) * (1.0 - dy) * dz);
  atomicAdd(G2_RHO(sx, syy, szz), m * (1.0 - dx) * dy * dz);
  atomicAdd(G2_RHO(sxx, sy, sz), m * dx * (1.0 - dy) * (1.0 - dz));
  atomicAdd(G2_RHO(sxx, syy, sz), m * dx * dy * (1.0 - dz));
  atomicAdd(G2_RHO(sxx, sy, szz), m * dx * (1.0 - dy) * dz);
  atomicAdd(G2_RHO(sxx, syy, szz), m * dx * dy * dz);
#undef G2_RHO
}

struct PMGreens
{
  int id[G2GPU_MAX_GRAVS];	// GreensFxns[nA][nB] for the current target nB, nA = 0..D-1
  double par[G2GPU_MAX_GRAVS];
};

// k-space Green's functions of ngravs.c in mesh units (k integer): pgdelta :390, neg_pgdelta :407, pgyukawa :869, pgcoloyuk :831
__device__ __forceinline__ double pm_greens(int id, double par, double k2, double asmth2)
{
  switch (id)
    {
    case G2GPU_GREENS_NEWTON: return 1.0 / k2;
    case G2GPU_GREENS_NEG_NEWTON: return -1.0 / k2;
    case G2GPU_GREENS_YUKAWA: return 1.0 / (k2 + par * par) * exp(-par * par * asmth2);	// par = YUKAWA_IMASS / (2 pi)
    case G2GPU_GREENS_COLOYUK: return 1.0 / (k2 + par * par) * exp(-par * par * asmth2) + 1.0 / k2;
    default: return 0.0;
    }
}

// tab[i] = exp(-k_i^2 asmth2) / sinc(pi k_i / N)^4 for mesh index i (k_i = i > N/2 ? i - N : i): the Gaussian long-range filter and the
// twofold CIC deconvolution of pm_periodic.c:487-513 factorise over the three dimensions; the table is filled on the host in double.
// POT (pmpotential_periodic, pm_periodic.c:1014-1063): the same filter times `scale` = G / (pi L), and the k = 0 mode is KEPT (:1033-1035).
// The reference resets that mode only when its real part is NaN (:1060-1063); with the stock 1/k^2 it is -inf and every potential
// comes out infinite, so here a k = 0 mode that is not finite is set to zero -- what the comment at :1058 says the reset is for.
template <int D, bool POT>
__global__ void __launch_bounds__(256) pm_filter_kernel(int N, double asmth2, PMGreens Gf, const double *__restrict__ tab, const double2 *__restrict__ rk,
							 size_t kstride, double2 *__restrict__ potk, double scale)
{
  const int nzh = N / 2 + 1;
  const size_t idx = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if(idx >= (size_t) N * N * nzh)
    return;
  const int z = (int) (idx % nzh), y = (int) ((idx / nzh) % N), x = (int) (idx / ((size_t) nzh * N));
  const double kx = x > N / 2 ? x - N : x, ky = y > N / 2 ? y - N : y, kz = z;	// pm_periodic.c:472-485
  const double k2 = kx * kx + ky * ky + kz * kz;
  double2 out = make_double2(0.0, 0.0);
  if(POT || k2 > 0)
    {
      const double filt = -(__ldg(tab + x) * __ldg(tab + y) * __ldg(tab + z));	// :513
#pragma unroll
      for(int nA = 0; nA < D; nA++)
	{
	  if(POT && Gf.id[nA] == G2GPU_GREENS_NONE)
	    continue;		// (0 x an infinite density mode would be NaN)
	  const double smth = POT ? pm_greens(Gf.id[nA], Gf.par[nA], k2, asmth2) * filt * scale : pm_greens(Gf.id[nA], Gf.par[nA], k2, asmth2) * filt;
	  const double2 r = rk[(size_t) nA * kstride + idx];
	  out.x += r.x * smth;
	  out.y += r.y * smth;
	}
      if(POT && idx == 0 && !(isfinite(out.x) && isfinite(out.y)))
	out = make_double2(0.0, 0.0);
    }
  potk[idx] = out;		// force: k = 0 is zero (:525-526)
}

// pmpotential_periodic :1236-1267: trilinear (CIC) interpolation of the potential mesh of the particle's own species
__global__ void __launch_bounds__(256) pm_potgather_kernel(PMArgs A, const double *__restrict__ phi_all, float *__restrict__ pot)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= A.n)
    return;
  const G2PRec p = A.rec[i];
  const int N = A.N;
  const double *__restrict__ phi = phi_all + (size_t) ((A.t2g_packed >> (4 * min(max(p.type, 0), 5))) & 7u) * N * N * N;
  int sx, sy, sz;
  double dx, dy, dz;
  pm_cell(A, p.x, sx, dx);
  pm_cell(A, p.y, sy, dy);
  pm_cell(A, p.z, sz, dz);
  const int sxx = sx + 1 == N ? 0 : sx + 1, syy = sy + 1 == N ? 0 : sy + 1, szz = sz + 1 == N ? 0 : sz + 1;
#define G2_PHI(X, Y, Z) __ldg(phi + ((size_t) (X) * N + (Y)) * N + (Z))
  // corner order and products as in pm_periodic.c:1254-1267
  double acc = G2_PHI(sx, sy, sz) * (1.0 - dx) * (1.0 - dy) * (1.0 - dz);
  acc += G2_PHI(sx, syy, sz) * (1.0 - dx) * dy * (1.0 - dz);
  acc += G2_PHI(sx, sy, szz) * (1.0 - dx) * (1.0 - dy) * dz;
  acc += G2_PHI(sx, syy, szz) * (1.0 - dx) * dy * dz;
  acc += G2_PHI(sxx, sy, sz) * dx * (1.0 - dy) * (1.0 - dz);
  acc += G2_PHI(sxx, syy, sz) * dx * dy * (1.0 - dz);
  acc += G2_PHI(sxx, sy, szz) * dx * (1.0 - dy) * dz;
  acc += G2_PHI(sxx, syy, szz) * dx * dy * dz;
#undef G2_PHI
  pot[i] = (float) acc;		// P[i].Potential is a FLOAT
}

// Finite differences (:726-737) + trilinear interpolation (:739-781) fused: the 4-point difference along one dimension at the 8 corners of
// the particle's cell needs the potential at 6 mesh planes of that dimension (cell index - 2 .. + 3, periodic) and 2 x 2 of the others.
// One pass over the particles: every particle reads the potential mesh of its own species (mesh g at phi + g * N^3).
__global__ void __launch_bounds__(256) pm_gather_kernel(PMArgs A, const double *__restrict__ phi_all, double fac, float *__restrict__ gravpm)
{
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= A.n)
    return;
  const G2PRec p = A.rec[i];
  const int N = A.N;
  const double *__restrict__ phi = phi_all + (size_t) ((A.t2g_packed >> (4 * min(max(p.type, 0), 5))) & 7u) * N * N * N;
  int s[3];
  double d[3];
  pm_cell(A, p.x, s[0], d[0]);
  pm_cell(A, p.y, s[1], d[1]);
  pm_cell(A, p.z, s[2], d[2]);
  // element offsets of the 6 planes s-2 .. s+3 of every dimension (periodic wrap without integer division)
  unsigned int off[3][6];
  const unsigned int stride[3] = { (unsigned int) N * (unsigned int) N, (unsigned int) N, 1u };
#pragma unroll
  for(int dim = 0; dim < 3; dim++)
#pragma unroll
    for(int k = 0; k < 6; k++)
      {
	int c = s[dim] - 2 + k;
	c = c < 0 ? c + N : (c >= N ? c - N : c);
	off[dim][k] = (unsigned int) c * stride[dim];
      }
  const double wx[2] = { 1.0 - d[0], d[0] }, wy[2] = { 1.0 - d[1], d[1] }, wz[2] = { 1.0 - d[2], d[2] };
#pragma unroll
  for(int dim = 0; dim < 3; dim++)
    {
      double acc = 0.0;
      // corner order and products as in pm_periodic.c:769-778: (x,y,z), (x,yy,z), (x,y,zz), (x,yy,zz), (xx,y,z), ...
      const int order[8][3] = { {0, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 1, 1}, {1, 0, 0}, {1, 1, 0}, {1, 0, 1}, {1, 1, 1} };
#pragma unroll
      for(int q = 0; q < 8; q++)
	{
	  const int a = order[q][0], b = order[q][1], cz = order[q][2];
	  const int cc[3] = { a, b, cz };
	  // offsets of the two other dimensions at this corner, and the four planes along `dim`
	  unsigned int base = 0;
#pragma unroll
	  for(int e = 0; e < 3; e++)
	    if(e != dim)
	      base += off[e][2 + cc[e]];
	  const int k0 = cc[dim];	// corner plane index is 2 + k0: l = 1 + k0, r = 3 + k0, ll = k0, rr = 4 + k0
	  const double f = (4.0 / 3) * (__ldg(phi + base + off[dim][1 + k0]) - __ldg(phi + base + off[dim][3 + k0]))
	    - (1.0 / 6) * (__ldg(phi + base + off[dim][k0]) - __ldg(phi + base + off[dim][4 + k0]));
	  const double term = fac * f * wx[a] * wy[b] * wz[cz];
	  acc = q == 0 ? term : acc + term;
	}
      gravpm[3 * (size_t) i + dim] = (float) acc;	// P[i].GravPM is a FLOAT
    }
}

static void pm_release(g2gpu_ctx *c)
{
  if(c->pm_plans_valid)
    {
      cufftDestroy((cufftHandle) c->pm_fwd);
      cufftDestroy((cufftHandle) c->pm_inv);
    }
  c->pm_plans_valid = 0;
  if(c->pm_rho)
    cudaFree(c->pm_rho);
  if(c->pm_rk)
    cudaFree(c->pm_rk);
  if(c->pm_potk)
    cudaFree(c->pm_potk);
  if(c->pm_tab)
    cudaFree(c->pm_tab);
  c->pm_tab = nullptr;
  c->pm_rho = nullptr;
  c->pm_rk = nullptr;
  c->pm_potk = nullptr;
  c->pm_grid = 0;
}

void g2_pm_destroy(g2gpu_ctx *c) { pm_release(c); }

static int pm_reserve(g2gpu_ctx *c, int N)
{
  if(c->pm_grid == N)
    return 0;
  pm_release(c);
  const size_t nreal = (size_t) N * N * N, ncpx = (size_t) N * N * (N / 2 + 1);
  if(cudaMalloc(&c->pm_rho, sizeof(double) * nreal * c->D) != cudaSuccess || cudaMalloc(&c->pm_tab, sizeof(double) * N) != cudaSuccess || cudaMalloc(&c->pm_rk, sizeof(double2) * ncpx * c->D) != cudaSuccess
     || cudaMalloc(&c->pm_potk, sizeof(double2) * ncpx) != cudaSuccess)
    {
      pm_release(c);
      return g2_fail(G2GPU_ERR_NOMEM, "PM: device allocation of the %d^3 meshes failed", N);
    }
  cufftHandle f, b;
  if(cufftPlan3d(&f, N, N, N, CUFFT_D2Z) != CUFFT_SUCCESS)
    {
      pm_release(c);
      return g2_fail(G2GPU_ERR_CUDA, "PM: cufftPlan3d(D2Z, %d^3) failed", N);
    }
  if(cufftPlan3d(&b, N, N, N, CUFFT_Z2D) != CUFFT_SUCCESS)
    {
      cufftDestroy(f);
      pm_release(c);
      return g2_fail(G2GPU_ERR_CUDA, "PM: cufftPlan3d(Z2D, %d^3) failed", N);
    }
  cufftSetStream(f, c->stream);
  cufftSetStream(b, c->stream);
  c->pm_fwd = (int) f;
  c->pm_inv = (int) b;
  c->pm_plans_valid = 1;
  c->pm_grid = N;
  return 0;
}

template <int D>
static void launch_filter(g2gpu_ctx *c, int N, double asmth2, const PMGreens &Gf, bool pot, double scale)
{
  const size_t ncpx = (size_t) N * N * (N / 2 + 1);
  if(pot)
    pm_filter_kernel<D, true><<<g2_cdiv((long long) ncpx, 256), 256, 0, c->stream>>>(N, asmth2, Gf, (const double *) c->pm_tab, (const double2 *) c->pm_rk, ncpx,
											     (double2 *) c->pm_potk, scale);
  else
    pm_filter_kernel<D, false><<<g2_cdiv((long long) ncpx, 256), 256, 0, c->stream>>>(N, asmth2, Gf, (const double *) c->pm_tab, (const double2 *) c->pm_rk, ncpx,
											      (double2 *) c->pm_potk, scale);
}

// pmforce_periodic (potential == nullptr) or pmpotential_periodic (potential = host array of n floats, upload order)
static int pm_run(g2gpu_ctx *c, const g2gpu_pm_params *pp, float *potential)
{
  const bool pot = potential != nullptr;
  if(c->stage < 1)
    return g2_fail(G2GPU_ERR_STATE, "PM: no particles uploaded");
  const int N = pp->pmgrid, n = c->npart, D = c->D;
  if(N < 4 || N > 1024 || pp->boxsize <= 0 || pp->asmth <= 0)
    return g2_fail(G2GPU_ERR_ARG, "PM: bad parameters (PMGRID=%d BoxSize=%g Asmth=%g)", N, pp->boxsize, pp->asmth);
  for(int i = 0; i < D * D; i++)
    if(pp->greens_id[i] < G2GPU_GREENS_NONE || pp->greens_id[i] > G2GPU_GREENS_COLOYUK)
      return g2_fail(G2GPU_ERR_LAW, "PM: unknown k-space Green's function id %d for pair %d", pp->greens_id[i], i);
  G2_TRY(pm_reserve(c, N));
  // the result becomes the GravPM input of the next domain stage / dynamic update: both the upload-order and the current-order array
  // must exist (gather_kernel and dyn_particles_kernel write the latter)
  if(!pot)
    G2_TRY(g2_ensure_optional_inputs(c, 0, 1));
  cudaStream_t st = c->stream;
  G2_CUDA(cudaEventRecord(c->ev[13], st));

  PMArgs A;
  A.rec = c->in_rec; A.n = n; A.N = N;
  A.to_slab_fac = (float) (N / pp->boxsize);
  A.t2g_packed = 0;
  for(int t = 0; t < 6; t++)
    A.t2g_packed |= (unsigned int) c->type_to_grav[t] << (4 * t);
  double asmth2 = (2 * M_PI) * pp->asmth / pp->boxsize;	// pm_periodic.c:232-233
  asmth2 *= asmth2;
  double fac = pp->G / (M_PI * pp->boxsize);	// :236-237 (force), :832 (potential)
  const double potfac = fac;
  fac *= 1 / (2 * pp->boxsize / N);
  const size_t nreal = (size_t) N * N * N, ncpx = (size_t) N * N * (N / 2 + 1);

  // filter table (host, double, libm like the reference): exp(-k^2 asmth2) / sinc^4 per mesh index
  if(c->pm_tab_asmth2 != asmth2 || c->pm_tab_n != N)
    {
      double *tab = (double *) malloc(sizeof(double) * N);
      for(int i = 0; i < N; i++)
	{
	  const double k = i > N / 2 ? i - N : i;
	  double f = 1;
	  if(k != 0)
	    {
	      f = (M_PI * k) / N;
	      f = sin(f) / f;
	    }
	  const double ff = 1 / f;
	  tab[i] = exp(-k * k * asmth2) * ff * ff * ff * ff;
	}
      cudaError_t e = cudaMemcpyAsync(c->pm_tab, tab, sizeof(double) * N, cudaMemcpyHostToDevice, st);
      if(e == cudaSuccess)
	e = cudaStreamSynchronize(st);
      free(tab);
      if(e != cudaSuccess)
	return g2_fail(G2GPU_ERR_CUDA, "PM: %s", cudaGetErrorString(e));
      c->pm_tab_asmth2 = asmth2;
      c->pm_tab_n = N;
    }

  G2_CUDA(cudaMemsetAsync(c->pm_rho, 0, sizeof(double) * nreal * D, st));
  pm_deposit_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(A, (double *) c->pm_rho);
  c->launches++;
  for(int nA = 0; nA < D; nA++)
    if(cufftExecD2Z((cufftHandle) c->pm_fwd, (cufftDoubleReal *) c->pm_rho + (size_t) nA * nreal, (cufftDoubleComplex *) c->pm_rk + (size_t) nA * ncpx) != CUFFT_SUCCESS)
      return g2_fail(G2GPU_ERR_CUDA, "PM: forward FFT failed");
  // the density meshes are free now: mesh nB becomes the potential acting on species nB
  for(int nB = 0; nB < D; nB++)
    {
      PMGreens Gf;
      bool any = false;
      for(int nA = 0; nA < D; nA++)
	{
	  Gf.id[nA] = pp->greens_id[nA * D + nB];	// GreensFxns[nA][nB] (pm_periodic.c:512)
	  Gf.par[nA] = pp->greens_par[nA * D + nB];
	  any = any || Gf.id[nA] != G2GPU_GREENS_NONE;
	}
      if(!any)
	{
	  G2_CUDA(cudaMemsetAsync((double *) c->pm_rho + (size_t) nB * nreal, 0, sizeof(double) * nreal, st));
	  continue;
	}
      switch (D)
	{
	case 1: launch_filter<1>(c, N, asmth2, Gf, pot, potfac); break;
	case 2: launch_filter<2>(c, N, asmth2, Gf, pot, potfac); break;
	case 3: launch_filter<3>(c, N, asmth2, Gf, pot, potfac); break;
	case 4: launch_filter<4>(c, N, asmth2, Gf, pot, potfac); break;
	case 5: launch_filter<5>(c, N, asmth2, Gf, pot, potfac); break;
	case 6: launch_filter<6>(c, N, asmth2, Gf, pot, potfac); break;
	default: return g2_fail(G2GPU_ERR_ARG, "unsupported N_GRAVS %d", D);
	}
      c->launches++;
      if(cufftExecZ2D((cufftHandle) c->pm_inv, (cufftDoubleComplex *) c->pm_potk, (cufftDoubleReal *) c->pm_rho + (size_t) nB * nreal) != CUFFT_SUCCESS)
	return g2_fail(G2GPU_ERR_CUDA, "PM: inverse FFT failed");
    }
  if(pot)
    {
      // the spectra are consumed: their storage takes the per-particle result until it is copied out
      float *d_pot = (float *) c->pm_rk;
      if(sizeof(float) * (size_t) n > sizeof(double2) * ncpx * D)
	return g2_fail(G2GPU_ERR_NOMEM, "PM potential: %d particles need more staging than the %d^3 spectra provide", n, N);
      pm_potgather_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(A, (const double *) c->pm_rho, d_pot);
      c->launches++;
      G2_CUDA(cudaEventRecord(c->ev[14], st));
      G2_CUDA(cudaGetLastError());
      G2_CUDA(cudaMemcpyAsync(potential, d_pot, sizeof(float) * (size_t) n, cudaMemcpyDeviceToHost, st));
      G2_CUDA(cudaStreamSynchronize(st));
      c->d2h_bytes = sizeof(float) * (size_t) n;
      return 0;
    }
  pm_gather_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(A, (const double *) c->pm_rho, fac, c->in_gravpm);	// writes every particle (longrange.c:67)
  c->launches++;
  G2_CUDA(cudaEventRecord(c->ev[14], st));
  G2_CUDA(cudaGetLastError());
  c->have_gravpm = 1;		// the next g2gpu_domain carries GravPM along for the OldAcc term of the walk (gravtree.c:318-331)
  c->pm_done = 1;
  return 0;
}

int g2_pm_periodic(g2gpu_ctx *c, const g2gpu_pm_params *pp) { return pm_run(c, pp, nullptr); }

int g2_pm_potential_periodic(g2gpu_ctx *c, const g2gpu_pm_params *pp, float *potential)
{
  if(!potential)
    return g2_fail(G2GPU_ERR_ARG, "PM potential: null output array");
  return pm_run(c, pp, potential);
}

int g2_pm_download(g2gpu_ctx *c, float *gravpm)
{
  if(!c->pm_done)
    return g2_fail(G2GPU_ERR_STATE, "PM: g2gpu_pm_periodic has not run");
  const size_t bytes = sizeof(float) * 3 * (size_t) c->npart;
  G2_CUDA(cudaMemcpyAsync(gravpm, c->in_gravpm, bytes, cudaMemcpyDeviceToHost, c->stream));
  G2_CUDA(cudaStreamSynchronize(c->stream));
  c->d2h_bytes = bytes;
  return 0;
}
