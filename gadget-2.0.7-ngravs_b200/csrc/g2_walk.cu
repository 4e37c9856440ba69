// g2_walk.cu — stage 3 driver: active targets in merged Peano-Hilbert order, this rank's slice, launch of the walk kernel
// (g2_walk_kernel.cuh, instantiated per N_GRAVS in g2_walk_dN.cu); plus the direct-summation accuracy oracle.
#include "g2_walk_common.cuh"

// ---------------------------------------------------------------- active target list ---------------------------
// phorder[j] = particle index (current order) of the particle with rank j along the Peano-Hilbert curve of ALL species
// (captured by the stage-1 sort, g2_domain.cu).  flags[j] = that particle is active (gravtree.c:113: Ti_endstep == Ti_Current).
__global__ void __launch_bounds__(256) target_flag_kernel(const unsigned int *__restrict__ phorder, const G2PRec *__restrict__ prec, int n,
							  unsigned int *__restrict__ flags)
{
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if(j < n)
    flags[j] = prec[phorder[j]].active ? 1u : 0u;
}

__global__ void __launch_bounds__(256) target_compact_kernel(const unsigned int *__restrict__ scan, const unsigned int *__restrict__ phorder, int n,
							     unsigned int *__restrict__ targets)
{
  int j = blockIdx.x * blockDim.x + threadIdx.x;
  if(j < n && scan[j + 1] != scan[j])	// exclusive scan of the 0/1 flags: a step marks an active target
    targets[scan[j]] = phorder[j];
}

// This rank's slice [lo, hi) of the ntargets active targets, boundaries rounded to multiples of 32 so that the 32-target
// groups -- and with them every result bit -- do not depend on the number of ranks.  frac_lo/frac_hi: the rank's share as
// fractions of the target list (equal counts: rank/nranks; cost-weighted: from the last step's GravCost profile, domain.c:859-862).
__global__ void slice_kernel(const unsigned int *__restrict__ count, double frac_lo, double frac_hi, int last, int *__restrict__ slice,
			     unsigned long long *__restrict__ counters, const WalkExactParams ex, WalkExactParams *__restrict__ ex_out)
{
  if(threadIdx.x == 31 && blockIdx.x == 0 && ex_out)
    *ex_out = ex;
  if(threadIdx.x == 0 && blockIdx.x == 0)
    {
      const int nt = (int) *count;
      long long lo = (long long) (frac_lo * (double) nt), hi = (long long) (frac_hi * (double) nt);
      lo = (lo + 16) / 32 * 32;
      hi = last ? nt : (hi + 16) / 32 * 32;
      if(lo > nt) lo = nt;
      if(hi > nt) hi = nt;
      if(hi < lo) hi = lo;
      slice[G2_SLICE_NTARGETS] = nt;
      slice[G2_SLICE_LO] = (int) lo;
      slice[G2_SLICE_HI] = (int) hi;
    }
  if(threadIdx.x < 8 && blockIdx.x == 0)
    counters[threadIdx.x] = 0ull;
}

int g2_stage_targets(g2gpu_ctx *c, const g2gpu_walk_params *wp)
{
  const int n = c->npart;
  cudaStream_t st = c->stream;
  target_flag_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(c->phorder, c->prec, n, c->w_flags);
  G2_TRY(g2_scan_exclusive_u32(c, c->w_flags, c->w_flags, (size_t) n));
  target_compact_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(c->w_flags, c->phorder, n, c->w_targets);
  c->launches += 2;
  double flo, fhi;
  int last;
  if(c->slice_explicit)
    {
      flo = c->slice_frac[0]; fhi = c->slice_frac[1];
      last = fhi >= 1.0;
    }
  else
    {
      const int nr = c->cfg.nranks > 0 ? c->cfg.nranks : 1, rk = c->cfg.rank;
      flo = (double) rk / nr; fhi = (double) (rk + 1) / nr;
      last = rk == nr - 1;
    }
  WalkExactParams ex;
  memset(&ex, 0, sizeof(ex));
  if(wp)
    {
      ex.boxsize = wp->boxsize; ex.rcut = wp->rcut; ex.rcut2 = wp->rcut * wp->rcut; ex.theta = wp->theta; ex.errtol = wp->errtol_force_acc;
      ex.asmthfac = wp->asmth > 0 ? 0.5 / wp->asmth * (c->cfg.ntab / 3.0) : 0.0;	// forcetree.c:1708
      ex.ntab = c->cfg.ntab;
      for(int t = 0; t < 6; t++)
	ex.fsoft[t] = c->force_softening[t];
    }
  slice_kernel<<<1, 32, 0, st>>>(c->w_flags + n, flo, fhi, last, c->d_slice, c->d_counters, ex, wp ? (WalkExactParams *) c->d_exact : nullptr);
  c->launches++;
  G2_CUDA(cudaMemcpyAsync(c->h_slice, c->d_slice, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  c->slice_pending = 1;
  return 0;
}

// the slice of the last g2_stage_targets, on the host (waits for the stream if it has not been read yet)
int g2_fetch_slice(g2gpu_ctx *c)
{
  if(c->slice_pending)
    {
      G2_CUDA(cudaStreamSynchronize(c->stream));
      c->w_ntargets = c->h_slice[G2_SLICE_NTARGETS];
      c->w_lo = c->h_slice[G2_SLICE_LO];
      c->w_hi = c->h_slice[G2_SLICE_HI];
      c->slice_pending = 0;
    }
  return 0;
}

int g2_stage_walk(g2gpu_ctx *c, const g2gpu_walk_params *wp)
{
  if(c->stage < 3)
    return g2_fail(G2GPU_ERR_STATE, "walk: tree has not been built");
  if(!c->laws_set)
    return g2_fail(G2GPU_ERR_LAW, "walk: pair force laws not set (g2gpu_set_laws)");
  const int D = c->D;
  const bool sr = c->cfg.shortrange != 0, per = c->cfg.periodic != 0;
  if(sr && !c->srtable_set)
    return g2_fail(G2GPU_ERR_STATE, "walk: short-range table not set (g2gpu_set_srtable)");
  if((per || sr) && !(wp->boxsize > 0))
    return g2_fail(G2GPU_ERR_ARG, "walk: boxsize must be positive in a periodic configuration");
  if(sr && (!(wp->asmth > 0) || !(wp->rcut > 0)))
    return g2_fail(G2GPU_ERR_ARG, "walk: asmth and rcut must be positive under the TreePM split");
  cudaStream_t st = c->stream;
  G2_CUDA(cudaEventRecord(c->ev[6], st));
  G2_TRY(g2_stage_targets(c, wp));

  WalkArgs A;
  memset(&A, 0, sizeof(A));
  A.cells = c->wcells; A.wpart = c->wpart; A.targets = c->w_targets; A.slice = c->d_slice; A.prec = c->prec;
  A.cnt = (c->accumulator && c->counts_valid) ? c->wcnt : nullptr;
  A.gravpm = c->gravpm; A.srtable = c->d_srtable_f; A.acc = c->acc; A.cost = c->cost;
  A.oldacc_out = c->oldacc_out; A.counters = c->d_counters; A.work_counter = (unsigned int *) (c->d_counters + 3);
  A.cres = c->compact ? c->cres : nullptr;
  A.zc_acc = c->zc_acc; A.zc_cost = c->zc_cost; A.zc_oldacc = c->zc_oldacc; A.zc_aos = c->zc_aos;
  // the scan scratch is free once the slice is known: the list of targets to walk again in FP64
  A.redo_list = c->w_flags; A.redo_count = (unsigned int *) (c->d_counters + 6); A.redo_cap = (unsigned int) c->cfg.max_part;
  A.numnodes = c->numnodes; A.ntab = c->cfg.ntab;
  A.theta2 = (float) (wp->theta * wp->theta);
  A.errtol = (float) wp->errtol_force_acc;
  A.ex = (const WalkExactParams *) c->d_exact;
  A.boxsize = (float) wp->boxsize; A.boxinv = wp->boxsize > 0 ? (float) (1.0 / wp->boxsize) : 0.0f;
  // (the guard bands of the FP32 decisions are compile-time constants of g2_walk_kernel.cuh: G2_TOL_*)
  A.exact = c->walk_exact;
  A.sm_counter = c->walk_sm_local ? c->d_smcount : nullptr; A.nsm = c->nsm;	// (cleared right before the launch: the lattice walk uses the same counters)
  A.flush_mask = (unsigned int) c->walk_flush_mask;
  if(sr)
    {
      A.rcut = (float) wp->rcut; A.rcut2 = (float) (wp->rcut * wp->rcut);
      const double asmthfac_d = 0.5 / wp->asmth * (c->cfg.ntab / 3.0);	// forcetree.c:1708
      A.asmthfac = (float) asmthfac_d;
      A.utor2wpi = (float) (1.0 / (M_PI * 4 * wp->asmth * wp->asmth));	// forcetree.c:1711
      const double rmax = c->cfg.ntab / asmthfac_d;
      A.rmax2 = (float) (rmax * rmax);
      A.ntabm1f = (float) (c->cfg.ntab - 1);
      A.rmax2_border = (float) (rmax * rmax * 6.0e-7);
      // a tree that follows the host's drifted nodes (g2gpu_update_tree) may hold centres of mass outside their cubes: neither the
      // geometric cull shortcut nor the per-cell image shift may then be used
      A.shift_len_max = c->tree_dynamic ? 0.0f : (float) (0.499 * wp->boxsize - 1.001 * wp->rcut);
      A.cull_margin = c->tree_dynamic ? 3.0e38f : (float) (1.0e-5 * wp->boxsize);
    }
  A.G = wp->G; A.pos_fac_pre_g = wp->pos_fac_pre_g; A.pos_fac_post_g = wp->pos_fac_post_g;
  A.use_gravpm = wp->use_gravpm && c->have_gravpm;
  for(int t = 0; t < 6; t++)
    {
      A.fsoft[t] = (float) c->force_softening[t];
      A.t2g[t] = c->type_to_grav[t];
    }
  A.laws = c->laws;
  A.ntables = c->sr_ntables;
  memcpy(A.tabmap, c->sr_tabmap, sizeof(A.tabmap));
  bool stock = true;
  for(int i = 0; i < D * D; i++)
    if(c->laws.accel[i] != G2GPU_LAW_NEWTONIAN || c->laws.spline[i] != G2GPU_SPLINE_PLUMMER)
      stock = false;

  const size_t smem = sr ? sizeof(float) * (size_t) A.ntables * A.ntab : 0;
  // every CTA resident: the chunk counter balances the load.  The target count stays on the device, so the grid is sized for
  // the upper bound (all particles active); surplus CTAs find the counter exhausted and leave.
  int grid = c->nsm * WALK_KBLOCKS(D >= WALK_WIDE_D ? WALK_MINBLOCKS_WIDE : WALK_MINBLOCKS);
  const int need = g2_cdiv(g2_cdiv(c->npart, 32), WALK_KWARPS);
  if(grid > need)
    grid = need;
  // periodic box without PM: the lattice-sum correction walk first (its result enters the epilogue of the walk kernel)
  const bool lattice = per && !sr && c->lattice_set;
  if(lattice)
    {
      G2_TRY(g2_stage_lattice(c, wp));
      A.latt = c->latt; A.lattcost = c->lattcost;
    }
  G2_CUDA(cudaMemsetAsync(c->d_smcount, 0, sizeof(unsigned int) * G2_CHUNK_COUNTERS, st));
  G2_CUDA(cudaEventRecord(c->ev[7], st));
  {
    const bool uneq = c->cfg.unequal_softenings != 0;
    const int accd = c->acc_double, stats = c->walk_stats;
    int rc;
    switch (D)
      {
      case 1: rc = g2_launch_walk_d1(c, A, grid, smem, sr, per, uneq, stock, accd, stats); break;
      case 2: rc = g2_launch_walk_d2(c, A, grid, smem, sr, per, uneq, stock, accd, stats); break;
      case 3: rc = g2_launch_walk_d3(c, A, grid, smem, sr, per, uneq, stock, accd, stats); break;
      case 4: rc = g2_launch_walk_d4(c, A, grid, smem, sr, per, uneq, stock, accd, stats); break;
      case 5: rc = g2_launch_walk_d5(c, A, grid, smem, sr, per, uneq, stock, accd, stats); break;
      case 6: rc = g2_launch_walk_d6(c, A, grid, smem, sr, per, uneq, stock, accd, stats); break;
      default: return g2_fail(G2GPU_ERR_ARG, "unsupported N_GRAVS %d", D);
      }
    if(rc)
      return rc;
    c->launches++;
  }
  G2_CUDA(cudaEventRecord(c->ev[8], st));
  G2_CUDA(cudaGetLastError());
  c->stage = 4;
  return 0;
}

// ---------------------------------------------------------------- direct summation (accuracy oracle) -------------
// force_treeevaluate_direct (forcetree.c:3428-3548; gravity_forcetest, gravtree_forcetest.c:28-356): O(N) sum over all particles for a
// few targets, h = max of the two softenings, same pair laws, FP64 displacements and sums.  One CTA per target.  With the TreePM
// split (shortrange != 0) the sum is the SHORT-RANGE force the tree walk approximates (same table, cut at tabindex >= NTAB); periodic
// boxes use the nearest image (the reference's lattice-correction tables are outside the path).
struct DirectArgs
{
  const G2PRec *__restrict__ prec;
  const float *__restrict__ srtable;
  const int *__restrict__ targets;
  double *__restrict__ out;
  int n, D, ntab, sr, periodic, unequal, stock, ewald;
  double boxsize, asmthfac, utor2wpi;
  float fsoft[6];
  int t2g[6];
  unsigned char tabmap[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];
  G2LawTable laws;
};

// Exact lattice (Ewald) correction of a periodic box: the acceleration due to all periodic images of a unit source minus the nearest-image
// Newtonian term, for a nearest-image displacement (dx,dy,dz) = source - target in units of the box size; multiply by m / L^2.  This is
// what the reference tabulates on a 65^3 grid (ewald_force, ngravs.c:1170; lattice_init, forcetree.c:3611) and interpolates tri-linearly in
// lattice_corr (forcetree.c:3803) for gravity_forcetest; here the sums are evaluated directly in FP64 (alpha = 2, |n| <= 3, |h|^2 < 10,
// truncation error < 1e-11), so it is exact ground truth for Newtonian pairs, at ~2e4 flop per pair (accuracy tests only).
__device__ void ewald_correction(double x, double y, double z, double *out)
{
  const double alpha = 2.0, r2 = x * x + y * y + z * z, r = sqrt(r2);
  double fx = -x / (r2 * r), fy = -y / (r2 * r), fz = -z / (r2 * r);
  for(int i = -3; i <= 3; i++)
    for(int j = -3; j <= 3; j++)
      for(int k = -3; k <= 3; k++)
	{
	  const double dx = x - i, dy = y - j, dz = z - k;
	  const double d2 = dx * dx + dy * dy + dz * dz, d = sqrt(d2);
	  const double g = (erfc(alpha * d) + 2.0 * alpha * d / sqrt(M_PI) * exp(-alpha * alpha * d2)) / (d2 * d);
	  fx += dx * g;
	  fy += dy * g;
	  fz += dz * g;
	}
  for(int i = -3; i <= 3; i++)
    for(int j = -3; j <= 3; j++)
      for(int k = -3; k <= 3; k++)
	{
	  const int h2 = i * i + j * j + k * k;
	  if(h2 == 0 || h2 >= 10)
	    continue;
	  const double v = 2.0 / h2 * exp(-M_PI * M_PI * h2 / (alpha * alpha)) * sin(2.0 * M_PI * (i * x + j * y + k * z));
	  fx += i * v;
	  fy += j * v;
	  fz += k * v;
	}
  out[0] = fx;
  out[1] = fy;
  out[2] = fz;
}

__global__ void __launch_bounds__(256) direct_kernel(const DirectArgs A)
{
  __shared__ double s_red[3][256];
  const G2PRec p = A.prec[A.targets[blockIdx.x]];
  const int tg = A.t2g[p.type];
  const double hp = (double) A.fsoft[p.type];
  double ax = 0.0, ay = 0.0, az = 0.0;
  for(int i = threadIdx.x; i < A.n; i += blockDim.x)
    {
      const G2PRec q = A.prec[i];
      double dx = (double) q.x - (double) p.x, dy = (double) q.y - (double) p.y, dz = (double) q.z - (double) p.z;
      if(A.periodic)
	{
	  dx -= A.boxsize * rint(dx / A.boxsize);
	  dy -= A.boxsize * rint(dy / A.boxsize);
	  dz -= A.boxsize * rint(dz / A.boxsize);
	}
      const double r2 = dx * dx + dy * dy + dz * dz;
      if(r2 == 0.0 || q.m == 0.0f)
	continue;
      const double r = sqrt(r2);
      double h = hp;
      if(A.unequal)
	h = fmax(h, (double) A.fsoft[q.type]);
      const int sg = A.t2g[q.type], ij = tg * A.D + sg;
      int tabindex = 0;
      if(A.sr)
	{
	  tabindex = (int) (A.asmthfac * r);
	  if(tabindex >= A.ntab)
	    continue;
	}
      double fac;
      if(r >= h)
	{
	  double a;		// |a| of the pair law
	  if(A.stock)
	    a = (double) q.m / r2;
	  else
	    a = (double) accel_over_r(A.laws.accel[ij], A.laws.par[ij], p.m, q.m, (float) r2, (float) r, (float) (1.0 / r), 1.0f) * r;
	  if(A.sr)
	    a -= (double) q.m * A.utor2wpi * (double) A.srtable[(int) A.tabmap[ij] * A.ntab + tabindex];
	  fac = a / r;
	}
      else if(A.stock)
	{			// ngravs.c:420-434 in FP64
	  const double hinv = 1.0 / h, u = r * hinv, h3 = hinv * hinv * hinv;
	  fac = (double) q.m * h3 * (u < 0.5 ? 10.666666666667 + u * u * (32.0 * u - 38.4)
				     : 21.333333333333 - 48.0 * u + 38.4 * u * u - 10.666666666667 * u * u * u - 0.066666666667 / (u * u * u));
	}
      else
	fac = (double) accel_spline(A.laws.spline[ij], A.laws.par[ij], p.m, q.m, (float) h, (float) r, 1.0f);
      ax += dx * fac;
      ay += dy * fac;
      az += dz * fac;
      if(A.ewald)
	{			// forcetree.c:3513-3522: the images matter even inside the softening
	  double c[3];
	  const double linv = 1.0 / A.boxsize, mf = (double) q.m * linv * linv;
	  ewald_correction(dx * linv, dy * linv, dz * linv, c);
	  ax += mf * c[0];
	  ay += mf * c[1];
	  az += mf * c[2];
	}
    }
  s_red[0][threadIdx.x] = ax;
  s_red[1][threadIdx.x] = ay;
  s_red[2][threadIdx.x] = az;
  __syncthreads();
  for(int o = 128; o > 0; o >>= 1)
    {
      if(threadIdx.x < o)
	for(int k = 0; k < 3; k++)
	  s_red[k][threadIdx.x] += s_red[k][threadIdx.x + o];
      __syncthreads();
    }
  if(threadIdx.x < 3)
    A.out[3 * (size_t) blockIdx.x + threadIdx.x] = s_red[threadIdx.x][0];
}

int g2_direct_sum(g2gpu_ctx *c, const g2gpu_walk_params *wp, int ntargets, const int *targets, double *acc)
{
  if(c->stage < 2)
    return g2_fail(G2GPU_ERR_STATE, "direct summation needs g2gpu_domain (particles in current order)");
  if(!c->laws_set)
    return g2_fail(G2GPU_ERR_LAW, "pair force laws not set (g2gpu_set_laws)");
  const int sr = c->cfg.shortrange && wp->asmth > 0;
  if(sr && !c->srtable_set)
    return g2_fail(G2GPU_ERR_STATE, "direct: short-range table not set (g2gpu_set_srtable)");
  for(int i = 0; i < ntargets; i++)
    if(targets[i] < 0 || targets[i] >= c->npart)
      return g2_fail(G2GPU_ERR_ARG, "direct: target %d outside [0, %d)", targets[i], c->npart);
  DirectArgs A;
  memset(&A, 0, sizeof(A));
  int *d_t;
  double *d_out;
  G2_CUDA(cudaMalloc(&d_t, sizeof(int) * (size_t) ntargets));
  G2_CUDA(cudaMalloc(&d_out, sizeof(double) * 3 * (size_t) ntargets));
  G2_CUDA(cudaMemcpyAsync(d_t, targets, sizeof(int) * (size_t) ntargets, cudaMemcpyHostToDevice, c->stream));
  A.prec = c->prec; A.srtable = c->d_srtable_f; A.targets = d_t; A.out = d_out;
  A.n = c->npart; A.D = c->D; A.ntab = c->cfg.ntab; A.sr = sr; A.periodic = c->cfg.periodic != 0 && wp->boxsize > 0;
  A.unequal = c->cfg.unequal_softenings != 0;
  A.boxsize = wp->boxsize;
  A.ewald = c->direct_ewald && A.periodic && !sr;
  if(sr)
    {
      A.asmthfac = 0.5 / wp->asmth * (c->cfg.ntab / 3.0);
      A.utor2wpi = 1.0 / (M_PI * 4 * wp->asmth * wp->asmth);
    }
  A.stock = 1;
  for(int i = 0; i < c->D * c->D; i++)
    if(c->laws.accel[i] != G2GPU_LAW_NEWTONIAN || c->laws.spline[i] != G2GPU_SPLINE_PLUMMER)
      A.stock = 0;
  for(int t = 0; t < 6; t++)
    {
      A.fsoft[t] = (float) c->force_softening[t];
      A.t2g[t] = c->type_to_grav[t];
    }
  A.laws = c->laws;
  memcpy(A.tabmap, c->sr_tabmap, sizeof(A.tabmap));
  direct_kernel<<<ntargets, 256, 0, c->stream>>>(A);
  c->launches++;
  cudaError_t e = cudaMemcpyAsync(acc, d_out, sizeof(double) * 3 * (size_t) ntargets, cudaMemcpyDeviceToHost, c->stream);
  if(e == cudaSuccess)
    e = cudaStreamSynchronize(c->stream);
  cudaFree(d_t);
  cudaFree(d_out);
  if(e != cudaSuccess)
    return g2_fail(G2GPU_ERR_CUDA, "direct: %s", cudaGetErrorString(e));
  return 0;
}

// ---------------------------------------------------------------- stand-alone pair evaluation (tests) -----------
__global__ void eval_pairs_kernel(int n, int accel_id, int spline_id, float p0, float p1, const float *__restrict__ pm, const float *__restrict__ m,
				  const float *__restrict__ r, const float *__restrict__ h, const int *__restrict__ nn, float *__restrict__ fac)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n)
    return;
  float par[4] = { p0, p1, 0, 0 };
  float rr = r[i], r2 = rr * rr, rinv = rsqrtf(fmaxf(r2, 1.0e-37f));
  float N = nn ? (float) nn[i] : 1.0f;
  fac[i] = rr >= h[i] ? accel_over_r(accel_id, par, pm[i], m[i], r2, rr, rinv, N) : accel_spline(spline_id, par, pm[i], m[i], h[i], rr, N);
}

int g2_eval_pairs_standalone(g2gpu_ctx *c, int n, int tgt, int src, const float *pm, const float *m, const float *r, const float *h,
			     const int *nn, float *fac)
{
  if(!c->laws_set)
    return g2_fail(G2GPU_ERR_LAW, "pair force laws not set");
  float *d;
  int *dn = nullptr;
  size_t fb = sizeof(float) * (size_t) n;
  G2_CUDA(cudaMalloc(&d, 5 * fb));
  G2_CUDA(cudaMemcpy(d, pm, fb, cudaMemcpyHostToDevice));
  G2_CUDA(cudaMemcpy(d + n, m, fb, cudaMemcpyHostToDevice));
  G2_CUDA(cudaMemcpy(d + 2 * (size_t) n, r, fb, cudaMemcpyHostToDevice));
  G2_CUDA(cudaMemcpy(d + 3 * (size_t) n, h, fb, cudaMemcpyHostToDevice));
  if(nn)
    {
      G2_CUDA(cudaMalloc(&dn, sizeof(int) * (size_t) n));
      G2_CUDA(cudaMemcpy(dn, nn, sizeof(int) * (size_t) n, cudaMemcpyHostToDevice));
    }
  int ij = tgt * c->D + src;
  eval_pairs_kernel<<<g2_cdiv(n, 256), 256, 0, c->stream>>>(n, c->laws.accel[ij], c->laws.spline[ij], c->laws.par[ij][0], c->laws.par[ij][1], d, d + n,
							    d + 2 * (size_t) n, d + 3 * (size_t) n, dn, d + 4 * (size_t) n);
  c->launches++;
  G2_CUDA(cudaStreamSynchronize(c->stream));
  G2_CUDA(cudaMemcpy(fac, d + 4 * (size_t) n, fb, cudaMemcpyDeviceToHost));
  cudaFree(d);
  if(dn)
    cudaFree(dn);
  return 0;
}
