// g2_lattice.cu — the lattice-sum correction walk of a periodic box WITHOUT PM: force_treeevaluate_lattice_correction
// (forcetree.c:2077-2455), which force_treeevaluate runs after its own walk (forcetree.c:1606-1608), and the table it interpolates
// (lattice_init, forcetree.c:3611-3790; ewald_force, ngravs.c:1170-1236).  SURVEY.md 8f-3.
//
// Same traversal as g2_walk.cu (one cursor per 32 tree-adjacent targets over the depth-first cell records, per-lane decisions, warp
// vote to descend) with this walk's own rule: a cell is used unless the force walk's opening criterion fires AND (the cell straddles
// the periodic boundary as seen from the target, or is longer than 0.2 BoxSize).  Every used cell contributes, per species with mass,
// mass * sign * trilinear(fcorr[tgt][src]) of the nearest-image displacement; every particle of an opened cell likewise.  Each used
// cell / particle counts 1 (the return value; added to GravCost, forcetree.c:2438).  The correction is summed in FP64 (table values
// FP32, 8 x 16-byte corner loads per term, L2 resident) and handed to the walk kernel, which adds it to the FLOAT GravAccel before
// the gravity_tree epilogue exactly where the reference does (forcetree.c:2435-2437).
#include "g2_walk_common.cuh"

// resident CTAs of 128 threads per SM.  MEASURED (B200, periodic 128^3 without PM, profiles/experiments/r1_pot_lattice_occupancy.sh):
// 6 (80 registers) -> 78.6 ms, 7 (72) -> 79.0 ms, 8 (64, 44 bytes spilled) -> 85.8 ms
#ifndef LATTICE_MINBLOCKS
#define LATTICE_MINBLOCKS 6
#endif

struct LatticeArgs
{
  const float4 *__restrict__ cells;
  const float4 *__restrict__ wpart;
  const unsigned int *__restrict__ targets;	// particle index of every active target, merged Peano-Hilbert order
  const int *__restrict__ slice;
  const G2PRec *__restrict__ prec;
  const float4 *__restrict__ tables;	// unique tables, (EN+1)^3 float4 (fx, fy, fz, 0) each, already divided by BoxSize^2
  float *__restrict__ latt;		// 3n, current particle order
  float *__restrict__ lattcost;	// n
  unsigned int *__restrict__ work_counter;
  unsigned int *__restrict__ sm_counter;	// ChunkDealer
  int nsm;
  int numnodes, en;
  float theta2, errtol, boxsize, boxinv, fac_intp;
  int t2g[6];
  unsigned char tabmap[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];
};

// one term: forcetree.c:2270-2352
__device__ __forceinline__ void lattice_term(const LatticeArgs &A, int ij, float m, float dx, float dy, float dz, double &ax, double &ay, double &az)
{
  const float sx = dx < 0.0f ? 1.0f : -1.0f, sy = dy < 0.0f ? 1.0f : -1.0f, sz = dz < 0.0f ? 1.0f : -1.0f;
  float u = fabsf(dx) * A.fac_intp, v = fabsf(dy) * A.fac_intp, w = fabsf(dz) * A.fac_intp;
  const int en = A.en, n1 = en + 1;
  int i = min((int) u, en - 1), j = min((int) v, en - 1), k = min((int) w, en - 1);
  u -= (float) i; v -= (float) j; w -= (float) k;
  const float4 *t = A.tables + (size_t) A.tabmap[ij] * n1 * n1 * n1 + ((size_t) i * n1 + j) * n1 + k;
  const float4 c000 = __ldg(t), c001 = __ldg(t + 1), c010 = __ldg(t + n1), c011 = __ldg(t + n1 + 1);
  t += (size_t) n1 * n1;
  const float4 c100 = __ldg(t), c101 = __ldg(t + 1), c110 = __ldg(t + n1), c111 = __ldg(t + n1 + 1);
  const float f1 = (1 - u) * (1 - v) * (1 - w), f2 = (1 - u) * (1 - v) * w, f3 = (1 - u) * v * (1 - w), f4 = (1 - u) * v * w;
  const float f5 = u * (1 - v) * (1 - w), f6 = u * (1 - v) * w, f7 = u * v * (1 - w), f8 = u * v * w;
  const float gx = c000.x * f1 + c001.x * f2 + c010.x * f3 + c011.x * f4 + c100.x * f5 + c101.x * f6 + c110.x * f7 + c111.x * f8;
  const float gy = c000.y * f1 + c001.y * f2 + c010.y * f3 + c011.y * f4 + c100.y * f5 + c101.y * f6 + c110.y * f7 + c111.y * f8;
  const float gz = c000.z * f1 + c001.z * f2 + c010.z * f3 + c011.z * f4 + c100.z * f5 + c101.z * f6 + c110.z * f7 + c111.z * f8;
  ax += (double) (m * sx * gx);
  ay += (double) (m * sy * gy);
  az += (double) (m * sz * gz);
}

template <int D>
__global__ void __launch_bounds__(WALK_THREADS, LATTICE_MINBLOCKS) lattice_kernel(const LatticeArgs A)
{
  __shared__ unsigned int s_chunk[WALK_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int R = 2 + D;
  const int lo = A.slice[G2_SLICE_LO], hi = A.slice[G2_SLICE_HI];
  const int nchunks = (hi - lo + 31) >> 5;
  const unsigned int end = (unsigned int) A.numnodes;

  unsigned int steal = 0u;	// chunk_dealer_next
  while(true)
    {
      if(lane == 0)
	s_chunk[warp] = chunk_dealer_next(A.sm_counter, A.work_counter, A.nsm, (unsigned int) nchunks, steal);
      __syncwarp();
      const unsigned int chunk = s_chunk[warp];
      __syncwarp();
      if(chunk >= (unsigned int) nchunks)
	break;
      const int ti = lo + (int) chunk * 32 + lane;
      const bool valid = ti < hi;
      unsigned int idx = 0;
      float px = 0, py = 0, pz = 0, aold = 0;
      int ptype = 1;
      if(valid)
	{
	  idx = A.targets[ti];
	  const G2PRec p = A.prec[idx];
	  px = p.x; py = p.y; pz = p.z;
	  ptype = p.type;
	  aold = A.errtol * p.oldacc;	// forcetree.c:1289, handed on at 1607
	}
      const int tg = A.t2g[ptype];
      double ax = 0.0, ay = 0.0, az = 0.0;
      int cost = 0;
      unsigned int skip_until = valid ? 0u : 0xffffffffu;
      unsigned int cur = 0u;

      while(cur < end)
	{
	  const float4 *rec = A.cells + (size_t) cur * R;
	  const float4 q0 = __ldg(rec);
	  const uint4 w = __ldg((const uint4 *) (rec + 1 + D));
	  bool open = false;
	  if(cur >= skip_until)
	    {
	      float dx[D], dy[D], dz[D], mass[D];
	      float r2min = 3.0e38f, summass = 0.0f;
	      const float len = q0.x;
	      const float cxr = q0.y - px, cyr = q0.z - py, czr = q0.w - pz;
#pragma unroll
	      for(int g = 0; g < D; g++)
		{
		  const float4 q = __ldg(rec + 1 + g);
		  mass[g] = q.w;
		  summass += q.w;
		  dx[g] = nearest<true>(q.x - px, A.boxsize, A.boxinv);
		  dy[g] = nearest<true>(q.y - py, A.boxsize, A.boxinv);
		  dz[g] = nearest<true>(q.z - pz, A.boxsize, A.boxinv);
		  r2min = fminf(r2min, dx[g] * dx[g] + dy[g] * dy[g] + dz[g] * dz[g]);
		}
	      bool openflag = false;
	      if(A.theta2 > 0.0f)
		openflag = len * len > r2min * A.theta2;	// forcetree.c:2179-2185
	      else
		openflag = (summass * len * len > r2min * r2min * aold)
		  || (fabsf(cxr) < 0.60f * len && fabsf(cyr) < 0.60f * len && fabsf(czr) < 0.60f * len);	// forcetree.c:2186-2208
	      if(openflag)
		{		// forcetree.c:2211-2256: can the cell be used nevertheless?
		  const float lim = 0.5f * (A.boxsize - len);
		  const float u0 = fabsf(nearest<true>(cxr, A.boxsize, A.boxinv)), u1 = fabsf(nearest<true>(cyr, A.boxsize, A.boxinv)),
		    u2 = fabsf(nearest<true>(czr, A.boxsize, A.boxinv));
		  open = u0 > lim || u1 > lim || u2 > lim || len > 0.20f * A.boxsize;
		}
	      if(!open)
		{
		  skip_until = w.x;
#pragma unroll
		  for(int g = 0; g < D; g++)
		    if(mass[g] != 0.0f)	// forcetree.c:2268
		      lattice_term(A, tg * D + g, mass[g], dx[g], dy[g], dz[g], ax, ay, az);
		  cost++;	// forcetree.c:2430
		}
	    }
	  const unsigned int ball = __ballot_sync(0xffffffffu, open);
	  if(ball != 0u)
	    {
	      const unsigned int np = w.z & 15u;
	      for(unsigned int j = 0; j < np; j++)
		{
		  const float4 p = __ldg(A.wpart + w.y + j);
		  if(open)
		    {
		      const int sg = A.t2g[(w.z >> (4 + 3 * j)) & 7];
		      lattice_term(A, tg * D + sg, p.w, nearest<true>(p.x - px, A.boxsize, A.boxinv), nearest<true>(p.y - py, A.boxsize, A.boxinv),
				   nearest<true>(p.z - pz, A.boxsize, A.boxinv), ax, ay, az);
		      cost++;
		    }
		}
	      cur = cur + 1u;
	    }
	  else
	    cur = w.x;
	}
      if(valid)
	{
	  A.latt[3 * (size_t) idx + 0] = (float) ax;
	  A.latt[3 * (size_t) idx + 1] = (float) ay;
	  A.latt[3 * (size_t) idx + 2] = (float) az;
	  A.lattcost[idx] = (float) cost;
	}
    }
}

// runs before the walk kernel of the same g2_stage_walk call (targets compacted, slice set); fills c->latt / c->lattcost
int g2_stage_lattice(g2gpu_ctx *c, const g2gpu_walk_params *wp)
{
  const int D = c->D;
  cudaStream_t st = c->stream;
  if(!c->latt)
    {
      G2_CUDA(cudaMalloc((void **) &c->latt, sizeof(float) * 3 * (size_t) c->cfg.max_part));
      G2_CUDA(cudaMalloc((void **) &c->lattcost, sizeof(float) * (size_t) c->cfg.max_part));
    }
  LatticeArgs A;
  memset(&A, 0, sizeof(A));
  A.cells = c->wcells; A.wpart = c->wpart; A.targets = c->w_targets; A.slice = c->d_slice; A.prec = c->prec;
  A.tables = (const float4 *) c->d_lattice; A.latt = c->latt; A.lattcost = c->lattcost;
  A.work_counter = (unsigned int *) (c->d_counters + 5);
  A.sm_counter = c->walk_sm_local ? c->d_smcount : nullptr; A.nsm = c->nsm;
  A.numnodes = c->numnodes; A.en = c->lattice_en;
  A.theta2 = (float) (wp->theta * wp->theta);
  A.errtol = (float) wp->errtol_force_acc;
  A.boxsize = (float) wp->boxsize; A.boxinv = (float) (1.0 / wp->boxsize);
  A.fac_intp = (float) (2.0 * c->lattice_en / wp->boxsize);	// forcetree.c:3749
  for(int t = 0; t < 6; t++)
    A.t2g[t] = c->type_to_grav[t];
  memcpy(A.tabmap, c->lattice_tabmap, sizeof(A.tabmap));
  int grid = c->nsm * LATTICE_MINBLOCKS, need = g2_cdiv(g2_cdiv(c->npart, 32), WALK_WARPS);	// the target count stays on the device
  if(grid > need)
    grid = need;
  G2_CUDA(cudaMemsetAsync(A.work_counter, 0, sizeof(unsigned int), st));
  G2_CUDA(cudaMemsetAsync(c->d_smcount, 0, sizeof(unsigned int) * G2_CHUNK_COUNTERS, st));
  G2_CUDA(cudaEventRecord(c->ev[18], st));
  if(grid > 0)
    {
      switch (D)
	{
#ifndef G2_FAST_BUILD
	case 1: lattice_kernel<1><<<grid, WALK_THREADS, 0, st>>>(A); break;
	case 3: lattice_kernel<3><<<grid, WALK_THREADS, 0, st>>>(A); break;
	case 5: lattice_kernel<5><<<grid, WALK_THREADS, 0, st>>>(A); break;
	case 6: lattice_kernel<6><<<grid, WALK_THREADS, 0, st>>>(A); break;
#endif
	case 2: lattice_kernel<2><<<grid, WALK_THREADS, 0, st>>>(A); break;
	case 4: lattice_kernel<4><<<grid, WALK_THREADS, 0, st>>>(A); break;
	default: return g2_fail(G2GPU_ERR_ARG, "unsupported N_GRAVS %d", D);
	}
      c->launches++;
    }
  G2_CUDA(cudaEventRecord(c->ev[19], st));
  G2_CUDA(cudaGetLastError());
  return 0;
}

// ---- ewald_force (ngravs.c:1170-1236) on the grid of lattice_init (forcetree.c:3700-3706): x = 0.5 (i,j,k)/EN, dimensionless ----
__global__ void __launch_bounds__(128) ewald_table_kernel(int en, double *__restrict__ out)
{
  const int n1 = en + 1, q = blockIdx.x * blockDim.x + threadIdx.x;
  if(q >= n1 * n1 * n1)
    return;
  const int i = q / (n1 * n1), j = (q / n1) % n1, k = q % n1;
  double f[3] = { 0.0, 0.0, 0.0 };
  if(q != 0)
    {
      const double alpha = 2.0, x[3] = { 0.5 * i / en, 0.5 * j / en, 0.5 * k / en };
      const double r2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
      for(int a = 0; a < 3; a++)
	f[a] += x[a] / (r2 * sqrt(r2));
      for(int n0 = -4; n0 <= 4; n0++)
	for(int n1_ = -4; n1_ <= 4; n1_++)
	  for(int n2 = -4; n2 <= 4; n2++)
	    {
	      const double d[3] = { x[0] - n0, x[1] - n1_, x[2] - n2 };
	      const double r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
	      const double val = (erfc(alpha * r) + 2 * alpha * r / sqrt(M_PI) * exp(-alpha * alpha * r * r)) / (r * r * r);
	      for(int a = 0; a < 3; a++)
		f[a] -= d[a] * val;
	    }
      for(int h0 = -4; h0 <= 4; h0++)
	for(int h1 = -4; h1 <= 4; h1++)
	  for(int h2_ = -4; h2_ <= 4; h2_++)
	    {
	      const int h2 = h0 * h0 + h1 * h1 + h2_ * h2_;
	      if(h2 > 0)
		{
		  const double val = 2.0 / (double) h2 * exp(-M_PI * M_PI * h2 / (alpha * alpha)) * sin(2 * M_PI * (x[0] * h0 + x[1] * h1 + x[2] * h2_));
		  f[0] -= h0 * val;
		  f[1] -= h1 * val;
		  f[2] -= h2_ * val;
		}
	    }
    }
  const size_t n3 = (size_t) n1 * n1 * n1;
  out[q] = f[0];
  out[n3 + q] = f[1];
  out[2 * n3 + q] = f[2];
}

int g2_make_ewald_table(g2gpu_ctx *c, int en, double *out)
{
  if(en < 1 || en > 256)
    return g2_fail(G2GPU_ERR_ARG, "ewald table: EN must be in [1, 256]");
  const size_t n3 = (size_t) (en + 1) * (en + 1) * (en + 1);
  double *d;
  G2_CUDA(cudaMalloc((void **) &d, sizeof(double) * 3 * n3));
  ewald_table_kernel<<<g2_cdiv((int) n3, 128), 128, 0, c->stream>>>(en, d);
  c->launches++;
  cudaError_t e = cudaMemcpyAsync(out, d, sizeof(double) * 3 * n3, cudaMemcpyDeviceToHost, c->stream);
  if(e == cudaSuccess)
    e = cudaStreamSynchronize(c->stream);
  cudaFree(d);
  if(e != cudaSuccess)
    return g2_fail(G2GPU_ERR_CUDA, "ewald table: %s", cudaGetErrorString(e));
  return 0;
}

// ---- ewald_psi (ngravs.c:761-816) on the same grid: potcorr of lattice_init for the stock wiring, before the division by BoxSize;
//      the origin holds LatticeZero (forcetree.c:3699-3702) ----
__global__ void __launch_bounds__(128) ewald_pot_table_kernel(int en, double latticezero, double *__restrict__ out)
{
  const int n1 = en + 1, q = blockIdx.x * blockDim.x + threadIdx.x;
  if(q >= n1 * n1 * n1)
    return;
  const int i = q / (n1 * n1), j = (q / n1) % n1, k = q % n1;
  if(q == 0)
    {
      out[0] = latticezero;
      return;
    }
  const double alpha = 2.0, x[3] = { 0.5 * i / en, 0.5 * j / en, 0.5 * k / en };
  double sum1 = 0.0, sum2 = 0.0;
  for(int n0 = -4; n0 <= 4; n0++)
    for(int n1_ = -4; n1_ <= 4; n1_++)
      for(int n2 = -4; n2 <= 4; n2++)
	{
	  const double d[3] = { x[0] - n0, x[1] - n1_, x[2] - n2 };
	  const double r = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
	  sum1 += erfc(alpha * r) / r;
	}
  for(int h0 = -4; h0 <= 4; h0++)
    for(int h1 = -4; h1 <= 4; h1++)
      for(int h2_ = -4; h2_ <= 4; h2_++)
	{
	  const int h2 = h0 * h0 + h1 * h1 + h2_ * h2_;
	  if(h2 > 0)
	    sum2 += 1 / (M_PI * h2) * exp(-M_PI * M_PI * h2 / (alpha * alpha)) * cos(2 * M_PI * (x[0] * h0 + x[1] * h1 + x[2] * h2_));
	}
  out[q] = M_PI / (alpha * alpha) - sum1 - sum2 + 1 / sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
}

int g2_make_ewald_pot_table(g2gpu_ctx *c, int en, double latticezero, double *out)
{
  if(en < 1 || en > 256)
    return g2_fail(G2GPU_ERR_ARG, "ewald table: EN must be in [1, 256]");
  const size_t n3 = (size_t) (en + 1) * (en + 1) * (en + 1);
  double *d;
  G2_CUDA(cudaMalloc((void **) &d, sizeof(double) * n3));
  ewald_pot_table_kernel<<<g2_cdiv((int) n3, 128), 128, 0, c->stream>>>(en, latticezero, d);
  c->launches++;
  cudaError_t e = cudaMemcpyAsync(out, d, sizeof(double) * n3, cudaMemcpyDeviceToHost, c->stream);
  if(e == cudaSuccess)
    e = cudaStreamSynchronize(c->stream);
  cudaFree(d);
  if(e != cudaSuccess)
    return g2_fail(G2GPU_ERR_CUDA, "ewald potential table: %s", cudaGetErrorString(e));
  return 0;
}
