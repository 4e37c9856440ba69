// g2_pot.cu — the tree potential walks of compute_potential (potential.c:22-354): force_treeevaluate_potential_shortrange
// (forcetree.c:2789-3163, TreePM) and force_treeevaluate_potential (forcetree.c:2467-2776, no PM; in a periodic box every term also
// gets mass * lattice_pot_corr, forcetree.c:2736-2738, 2765-2767, 3895-3941: instantiation LATT).  SURVEY.md 8f-3.
//
// Same traversal as g2_walk.cu (one cursor per 32 tree-adjacent targets over the depth-first cell records, per-lane decisions, warp
// vote to descend), with the potential walk's own rules, reproduced as the reference has them:
//   * every particle is a target (potential.c:86), the target itself included: its self term PotentialSplines(m, h, 0) is removed by
//     the caller afterwards (potential.c:250-254);
//   * TreePM cull: the box test ALONE (|NEAREST(center - pos)| > rcut + len/2 on some axis, forcetree.c:2988-3015), not preceded by
//     the r2min > rcut^2 test of the force walk;
//   * a term counts if tabindex = (int)(r * asmthfac) < NTAB (forcetree.c:3108-3112, 3128-3132);
//   * particle term, r >= h:  pot -= PotentialFxns(m, r) - utorwpi * shortrange_fourier_pot[tgt][src][tabindex]   (forcetree.c:3115: the
//     table term carries NO mass factor in the reference); r < h: pot += PotentialSplines(m, h, r);
//   * node term: the table term is present only in a -DNGRAVS_ACCUMULATOR build (forcetree.c:3134-3140) -- option "accumulator";
//     species with zero mass in the node are skipped (forcetree.c:3123).
// Terms are evaluated in FP32 and summed into an FP64 accumulator whenever the warp descends (like the force walk); the result is
// stored as FLOAT like P[].Potential.  Decisions that fall within float rounding of the cull boundary are re-taken in FP64, which is
// what the reference computes in.
#include "g2_walk_common.cuh"

// resident CTAs of 128 threads per SM.  MEASURED (B200, periodic 128^3 TreePM, D = 2, profiles/experiments/r1_pot_lattice_occupancy.sh):
// 6 -> 14.5 ms, 8 -> 12.0 ms, 9 -> 11.3 ms (54 registers as written, no spills at the 56-register cap); D >= 3 needs up to 64 registers
#ifndef POT_MINBLOCKS
#define POT_MINBLOCKS 9
#endif
#ifndef POT_MINBLOCKS_WIDE
#define POT_MINBLOCKS_WIDE 8
#endif
#define POT_BLOCKS(D) ((D) >= WALK_WIDE_D ? POT_MINBLOCKS_WIDE : POT_MINBLOCKS)

struct PotArgs
{
  const float4 *__restrict__ cells;
  const float4 *__restrict__ wpart;
  const unsigned int *__restrict__ order;	// particle index by rank along the Peano-Hilbert curve of all species (the walk's target order)
  const G2PRec *__restrict__ prec;
  const float *__restrict__ pottable;	// unique tables, NTAB floats each
  float *__restrict__ pot;
  unsigned int *__restrict__ work_counter;
  unsigned int *__restrict__ sm_counter;	// ChunkDealer
  int nsm;
  int lo, hi, numnodes, ntab, ntables, node_table_term;
  float theta2, errtol, boxsize, boxinv, rcut, asmthfac, utorwpi;
  float shift_len_max;		// -DG2_POT_CELLSHIFT: cells smaller than this take the periodic image of their centre for all their points
  double rcut_d, boxsize_d;
  float fsoft[6];
  int t2g[6];
  unsigned char tabmap[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];
  int potfxn[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS], potspline[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];
  float bam_eps[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];	// BAM_EPSILON (ngravs.c:45), from the parameters of g2gpu_set_laws
  const unsigned int *__restrict__ cnt;	// particle counts per species of every cell (NGRAVS_ACCUMULATOR), or null
  // LATT: potcorr[target][source] of lattice_init (forcetree.c:52, 3697-3702, 3759), unique tables of (EN+1)^3 doubles
  const double *__restrict__ potcorr;
  double fac_intp;		// 2 EN / BoxSize (forcetree.c:3750)
  int potcorr_en;
  unsigned char potcorr_map[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];
};

// lattice_pot_corr (forcetree.c:3895-3941): trilinear interpolation in the octant table, in double like the reference
__device__ __forceinline__ double latt_pot_corr(const PotArgs &A, float dxf, float dyf, float dzf, int ij)
{
  const int en = A.potcorr_en, n1 = en + 1;
  const double *__restrict__ t = A.potcorr + (size_t) A.potcorr_map[ij] * n1 * n1 * n1;
  double u = fabs((double) dxf) * A.fac_intp, v = fabs((double) dyf) * A.fac_intp, w = fabs((double) dzf) * A.fac_intp;
  int i = (int) u, j = (int) v, k = (int) w;
  if(i >= en) i = en - 1;
  if(j >= en) j = en - 1;
  if(k >= en) k = en - 1;
  u -= i; v -= j; w -= k;
  const double *__restrict__ b = t + ((size_t) i * n1 + j) * n1 + k;
  const size_t sj = (size_t) n1, si = (size_t) n1 * n1;
  return __ldg(b) * ((1 - u) * (1 - v) * (1 - w)) + __ldg(b + 1) * ((1 - u) * (1 - v) * w) + __ldg(b + sj) * ((1 - u) * v * (1 - w)) +
    __ldg(b + sj + 1) * ((1 - u) * v * w) + __ldg(b + si) * (u * (1 - v) * (1 - w)) + __ldg(b + si + 1) * (u * (1 - v) * w) +
    __ldg(b + si + sj) * (u * v * (1 - w)) + __ldg(b + si + sj + 1) * (u * v * w);
}

// plummer_pot, ngravs.c:459-471 (u = r/h)
__device__ __forceinline__ float pot_plummer(float m, float h, float r)
{
  const float hinv = __frcp_rn(h), u = r * hinv;
  if(u < 0.5f)
    return m * hinv * (-2.8f + u * u * (5.333333333333f + u * u * (6.4f * u - 9.6f)));
  return m * hinv * (-3.2f + 0.066666666667f / u + u * u * (10.666666666667f + u * (-16.0f + u * (9.6f - 2.133333333333f * u))));
}

// BAM family, ngravs.c:672-760 (bambam_pot, sourcebaryonbam_pot, sourcebambaryon_pot): rho * atan(r eta) / r with a 7th-order Taylor
// form for r eta < 0.1; eta depends on which side is the BAM halo.  Wired as PotentialFxns AND PotentialSplines (ngravs.c:198-200).
// Evaluated in FP64 like the BAM force laws (g2_laws.cuh) and rounded once.
__device__ __forceinline__ float pot_bam(int id, float eps, float pm, float m, float r, float nn)
{
  const double rho = 2.0 * (double) pm * (double) m / 3.14159265358979323846;
  double eta;
  if(id == G2GPU_POT_BAMBAM)
    eta = 4.0 * 3.14159265358979323846 * (double) eps / ((double) pm + (double) m / (double) nn);	// ngravs.c:685
  else if(id == G2GPU_POT_SOURCEBARYONBAM)
    eta = 4.0 * (double) eps * 3.14159265358979323846 * (double) nn / (double) pm;	// ngravs.c:715
  else
    eta = 4.0 * (double) eps * 3.14159265358979323846 * (double) nn / (double) m;	// ngravs.c:741
  const double reta = (double) r * eta, reta2 = reta * reta, reta4 = reta2 * reta2;
  if(reta < 0.1)
    return (float) (rho * eta * (1.0 - reta2 / 3.0 + reta4 / 5.0 - reta2 * reta4 / 7.0));
  return (float) (rho * atan(reta) / (double) r);
}

// one species term; returns the amount to ADD to pot
// STOCKP: every pair is wired to newtonian_pot / plummer_pot (the stock wiring, ngravs.c:149-150): no dispatch on the pair's law per term
template <bool SR, bool STOCKP = false>
__device__ __forceinline__ float pot_term(const PotArgs &A, const float *__restrict__ s_tab, int ij, float m, float r2, float h, bool table_term,
					  float pm = 0.0f, float nn = 1.0f)
{
  const float rinv = fast_rsqrt(fmaxf(r2, 1.0e-37f));
  const float r = r2 * rinv;
  int tabindex = 0;
  if(SR)
    {
      tabindex = (int) (r * A.asmthfac);
      if(tabindex >= A.ntab)
	return 0.0f;
    }
  if(STOCKP)
    {
      if(r >= h)
	{
	  float p = m * rinv;
	  if(SR && table_term)
	    p -= A.utorwpi * s_tab[(int) A.tabmap[ij] * A.ntab + tabindex];
	  return -p;
	}
      return pot_plummer(m, h, r);
    }
  if(r >= h)
    {
      float p;
      switch (A.potfxn[ij])
	{
	case G2GPU_POT_NEWTONIAN: p = m * rinv; break;	// ngravs.c:368
	case G2GPU_POT_NEG_NEWTONIAN: p = -m * rinv; break;	// ngravs.c:375
	case G2GPU_POT_BAMBAM:
	case G2GPU_POT_SOURCEBARYONBAM:
	case G2GPU_POT_SOURCEBAMBARYON: p = pot_bam(A.potfxn[ij], A.bam_eps[ij], pm, m, r, nn); break;
	default: p = 0.0f; break;	// none, ngravs.c:344
	}
      if(SR && table_term)
	p -= A.utorwpi * s_tab[(int) A.tabmap[ij] * A.ntab + tabindex];
      return -p;
    }
  switch (A.potspline[ij])
    {
    case G2GPU_POTSPLINE_PLUMMER: return pot_plummer(m, h, r);	// ngravs.c:459
    case G2GPU_POTSPLINE_NEG_PLUMMER: return -pot_plummer(m, h, r);	// ngravs.c:476
    case G2GPU_POTSPLINE_BAMBAM:
    case G2GPU_POTSPLINE_SOURCEBARYONBAM:
    case G2GPU_POTSPLINE_SOURCEBAMBARYON:	// the reference ADDS the "spline" (forcetree.c:2734, 3118), also when it is wired to a *_pot function
      return pot_bam(A.potspline[ij] - G2GPU_POTSPLINE_BAMBAM + G2GPU_POT_BAMBAM, A.bam_eps[ij], pm, m, r, nn);
    default: return 0.0f;
    }
}

__device__ __forceinline__ double nearest_d(double x, double box)
{
  if(x > 0.5 * box)
    x -= box;
  if(x < -0.5 * box)
    x += box;
  return x;
}

template <int D, bool SR, bool UNEQUAL, bool LATT, bool STOCKP>
__global__ void __launch_bounds__(WALK_THREADS, LATT ? 4 : POT_BLOCKS(D)) pot_kernel(const PotArgs A)
{
  constexpr bool PERIODIC = SR || LATT;	// the reference's TreePM potential walk is the periodic one; LATT: periodic box without PM
  extern __shared__ float s_tab[];
  __shared__ unsigned int s_chunk[WALK_WARPS];
  if(SR)
    {
      for(int i = threadIdx.x; i < A.ntables * A.ntab; i += WALK_THREADS)
	s_tab[i] = A.pottable[i];
      __syncthreads();
    }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int R = 2 + D;
  const int nchunks = (A.hi - A.lo + 31) >> 5;
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
      const int ti = A.lo + (int) chunk * 32 + lane;
      const bool valid = ti < A.hi;
      unsigned int idx = 0;
      float px = 0, py = 0, pz = 0, aold = 0, pmass = 0;
      int ptype = 1;
      if(valid)
	{
	  idx = A.order[ti];
	  const G2PRec p = A.prec[idx];
	  px = p.x; py = p.y; pz = p.z; pmass = p.m;
	  ptype = p.type;
	  aold = A.errtol * p.oldacc;	// forcetree.c:2830
	}
      const int tg = A.t2g[ptype];
      const float hself = A.fsoft[ptype];
      double pot = 0.0;
      float fpot = 0.0f;
      unsigned int skip_until = valid ? 0u : 0xffffffffu;
      unsigned int cur = 0u;

      while(cur < end)
	{
	  const float4 *rec = A.cells + (size_t) cur * R;
	  const float4 q0 = __ldg(rec);
	  const uint4 w = __ldg((const uint4 *) (rec + 1 + D));
	  bool open = false;
#ifdef G2_POT_CELLSHIFT
	  // EXPERIMENT (off; not yet measured): the per-cell image shift of the force walk (g2_walk.cu).  A cell that is not culled lies within
	  // rcut + len/2 of the target on every axis, so for len < L/2 - rcut all its points share the periodic image of its centre: 3 rint
	  // per cell instead of 3 per point.  Same results bit for bit (box * k is exact for k = -1, 0, 1).
	  const bool small_cell = SR && q0.x < A.shift_len_max;
	  float shx = 0.0f, shy = 0.0f, shz = 0.0f;
#define POT_WRAP(v, sh) (small_cell ? (v) - (sh) : nearest<PERIODIC>((v), A.boxsize, A.boxinv))
#else
#define POT_WRAP(v, sh) nearest<PERIODIC>((v), A.boxsize, A.boxinv)
#endif
	  if(cur >= skip_until)
	    {
	      float r2[D], mass[D], dsp[LATT ? D : 1][3];
	      float r2min = 3.0e38f, r2max = -1.0f, summass = 0.0f;
	      const float len = q0.x;
	      const float cxr = q0.y - px, cyr = q0.z - py, czr = q0.w - pz;
	      bool done = false;
	      if(SR)
		{		// forcetree.c:2988-3015
		  const float eff = A.rcut + 0.5f * len;
#ifdef G2_POT_CELLSHIFT
		  shx = A.boxsize * rint_small(cxr * A.boxinv);
		  shy = A.boxsize * rint_small(cyr * A.boxinv);
		  shz = A.boxsize * rint_small(czr * A.boxinv);
		  const float d0 = fabsf(cxr - shx), d1 = fabsf(cyr - shy), d2 = fabsf(czr - shz);
#else
		  const float d0 = fabsf(nearest<PERIODIC>(cxr, A.boxsize, A.boxinv)), d1 = fabsf(nearest<PERIODIC>(cyr, A.boxsize, A.boxinv)),
		    d2 = fabsf(nearest<PERIODIC>(czr, A.boxsize, A.boxinv));
#endif
		  const float dmax = fmaxf(fmaxf(d0, d1), d2);
		  done = dmax > eff;
		  if(fabsf(dmax - eff) < 4.0e-6f * A.boxsize)
		    {		// within float rounding of the boundary: decide as the reference does, in double
		      const double effd = A.rcut_d + 0.5 * (double) len;
		      const double e0 = fabs(nearest_d((double) q0.y - (double) px, A.boxsize_d)), e1 = fabs(nearest_d((double) q0.z - (double) py, A.boxsize_d)),
			e2 = fabs(nearest_d((double) q0.w - (double) pz, A.boxsize_d));
		      done = e0 > effd || e1 > effd || e2 > effd;
		    }
		}
	      if(!done)
		{
#pragma unroll
		  for(int g = 0; g < D; g++)
		    {
		      const float4 q = __ldg(rec + 1 + g);
		      mass[g] = q.w;
		      summass += q.w;
		      const float dx = POT_WRAP(q.x - px, shx);
		      const float dy = POT_WRAP(q.y - py, shy);
		      const float dz = POT_WRAP(q.z - pz, shz);
		      r2[g] = dx * dx + dy * dy + dz * dz;
		      if(LATT)
			{
			  dsp[g][0] = dx; dsp[g][1] = dy; dsp[g][2] = dz;
			}
		      r2min = fminf(r2min, r2[g]);
		      r2max = fmaxf(r2max, r2[g]);
		    }
		  if(A.theta2 > 0.0f)
		    {		// forcetree.c:3017-3025
		      if(len * len > r2min * A.theta2)
			open = true;
		    }
		  else
		    {		// forcetree.c:3026-3047
		      if(summass * len * len > r2min * r2min * aold)
			open = true;
		      else if(fabsf(cxr) < 0.60f * len && fabsf(cyr) < 0.60f * len && fabsf(czr) < 0.60f * len)
			open = true;
		    }
		}
	      float h = hself;
	      if(UNEQUAL && !done && !open)
		{		// forcetree.c:3049-3077 (the record carries ForceSoftening[maxsofttype]; +inf and the mixed bit for an empty node)
		  const float hnode = __uint_as_float(w.w);
		  if(h < hnode)
		    {
		      h = hnode;
		      if(r2max < h * h && ((w.z >> 28) & 1))
			open = true;
		    }
		}
	      if(!open)
		{
		  skip_until = w.x;
		  if(!done)
		    {
#pragma unroll
		      for(int g = 0; g < D; g++)
			if(mass[g] != 0.0f)	// forcetree.c:3123
			  {
			    fpot += pot_term<SR, STOCKP>(A, s_tab, tg * D + g, mass[g], r2[g], h, A.node_table_term != 0, pmass,
						 A.cnt ? (float) __ldg(A.cnt + (size_t) cur * D + g) : 1.0f);
			    if(LATT)	// forcetree.c:2765-2767
			      pot += (double) mass[g] * latt_pot_corr(A, dsp[g][0], dsp[g][1], dsp[g][2], tg * D + g);
			  }
		    }
		}
	    }
	  const unsigned int ball = __ballot_sync(0xffffffffu, open);
	  if(ball != 0u)
	    {
	      if((cur & 7u) == 0u)	// (the FP32 partial sum goes into the FP64 accumulator at every eighth cell index, like in the force walk)
		{
		  pot += (double) fpot;
		  fpot = 0.0f;
		}
	      const unsigned int np = w.z & 15u;
	      for(unsigned int j = 0; j < np; j++)
		{
		  const float4 p = __ldg(A.wpart + w.y + j);
		  if(open)
		    {
		      const int stype = (w.z >> (4 + 3 * j)) & 7;
		      const int sg = A.t2g[stype];
		      float h = hself;
		      if(UNEQUAL)
			h = fmaxf(h, A.fsoft[stype]);	// forcetree.c:2958-2961
		      const float dx = POT_WRAP(p.x - px, shx);
		      const float dy = POT_WRAP(p.y - py, shy);
		      const float dz = POT_WRAP(p.z - pz, shz);
		      fpot += pot_term<SR, STOCKP>(A, s_tab, tg * D + sg, p.w, dx * dx + dy * dy + dz * dz, h, true, pmass, 1.0f);
		      if(LATT)	// forcetree.c:2736-2738
			pot += (double) p.w * latt_pot_corr(A, dx, dy, dz, tg * D + sg);
		    }
		}
	      cur = cur + 1u;
	    }
	  else
	    cur = w.x;
	}
      pot += (double) fpot;
      if(valid)
	A.pot[idx] = (float) pot;	// forcetree.c:3158
    }
}

#undef POT_WRAP

template <int D>
static void launch_pot(g2gpu_ctx *c, const PotArgs &A, int grid, size_t smem, bool sr, bool unequal, bool latt)
{
  bool stockp = true;
  for(int i = 0; i < D * D; i++)
    if(A.potfxn[i] != G2GPU_POT_NEWTONIAN || A.potspline[i] != G2GPU_POTSPLINE_PLUMMER)
      stockp = false;
#define G2_P(SRv, UNEv, LATv) do { \
    if(stockp) { \
      if(smem > 48 * 1024) cudaFuncSetAttribute(pot_kernel<D, SRv, UNEv, LATv, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem); \
      pot_kernel<D, SRv, UNEv, LATv, true><<<grid, WALK_THREADS, smem, c->stream>>>(A); \
    } else { \
      if(smem > 48 * 1024) cudaFuncSetAttribute(pot_kernel<D, SRv, UNEv, LATv, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem); \
      pot_kernel<D, SRv, UNEv, LATv, false><<<grid, WALK_THREADS, smem, c->stream>>>(A); } } while(0)
  if(sr) { if(unequal) G2_P(true, true, false); else G2_P(true, false, false); }
  else if(latt) { if(unequal) G2_P(false, true, true); else G2_P(false, false, true); }
  else   { if(unequal) G2_P(false, true, false); else G2_P(false, false, false); }
#undef G2_P
}

int g2_stage_potential(g2gpu_ctx *c, const g2gpu_walk_params *wp)
{
  if(c->stage < 3)
    return g2_fail(G2GPU_ERR_STATE, "potential: tree has not been built");
  if(!c->potlaws_set)
    return g2_fail(G2GPU_ERR_LAW, "potential: pair potential laws not set (g2gpu_set_potential_laws)");
  const int n = c->npart, D = c->D;
  const bool sr = c->cfg.shortrange != 0;
  if(sr && !c->pottable_set)
    return g2_fail(G2GPU_ERR_STATE, "potential: short-range potential table not set (g2gpu_set_srpot_table)");
  const bool latt = !sr && c->cfg.periodic;
  if(latt && !c->potcorr_set)
    return g2_fail(G2GPU_ERR_STATE, "potential: a periodic box without PM needs the lattice-sum potential tables (g2gpu_set_lattice_pot_tables; lattice_pot_corr, forcetree.c:3895)");
  if(latt && !(wp->boxsize > 0))
    return g2_fail(G2GPU_ERR_ARG, "potential: boxsize must be positive in a periodic configuration");
  cudaStream_t st = c->stream;
  if(!c->pot)
    {
      G2_CUDA(cudaMalloc((void **) &c->pot, sizeof(float) * (size_t) c->cfg.max_part));
      G2_CUDA(cudaMemsetAsync(c->pot, 0, sizeof(float) * (size_t) c->cfg.max_part, st));
    }
  PotArgs A;
  memset(&A, 0, sizeof(A));
  const int nr = c->cfg.nranks > 0 ? c->cfg.nranks : 1, rk = c->cfg.rank;
  A.cells = c->wcells; A.wpart = c->wpart; A.order = c->phorder; A.prec = c->prec; A.pottable = c->d_pottable_f; A.pot = c->pot;
  A.work_counter = (unsigned int *) (c->d_counters + 5);
  A.sm_counter = c->walk_sm_local ? c->d_smcount : nullptr; A.nsm = c->nsm;
  // slice boundaries at multiples of 32: the 32-target groups, and with them the result bits, do not depend on the number of ranks
  A.lo = (int) (((long long) n * rk / nr + 16) / 32 * 32); A.hi = rk == nr - 1 ? n : (int) (((long long) n * (rk + 1) / nr + 16) / 32 * 32);
  if(A.lo > n) A.lo = n;
  if(A.hi > n) A.hi = n;
  if(A.hi < A.lo) A.hi = A.lo;
  A.numnodes = c->numnodes; A.ntab = c->cfg.ntab; A.ntables = c->pot_ntables;
  A.node_table_term = (c->accumulator != 0);
  A.theta2 = (float) (wp->theta * wp->theta);
  A.errtol = (float) wp->errtol_force_acc;
  A.boxsize = (float) wp->boxsize; A.boxinv = wp->boxsize > 0 ? (float) (1.0 / wp->boxsize) : 0.0f;
  A.boxsize_d = wp->boxsize;
  if(sr)
    {
      if(!(wp->asmth > 0) || !(wp->boxsize > 0))
	return g2_fail(G2GPU_ERR_ARG, "potential: asmth and boxsize must be positive under the TreePM split");
      A.rcut = (float) wp->rcut; A.rcut_d = wp->rcut;
      A.asmthfac = (float) (0.5 / wp->asmth * (c->cfg.ntab / 3.0));	// forcetree.c:2862
      A.utorwpi = (float) (1.0 / (2 * M_PI * wp->asmth));	// forcetree.c:2865
      A.shift_len_max = (float) (0.499 * wp->boxsize - 1.001 * wp->rcut);
    }
  for(int t = 0; t < 6; t++)
    {
      A.fsoft[t] = (float) c->force_softening[t];
      A.t2g[t] = c->type_to_grav[t];
    }
  memcpy(A.tabmap, c->pot_tabmap, sizeof(A.tabmap));
  memcpy(A.potfxn, c->potfxn, sizeof(A.potfxn));
  memcpy(A.potspline, c->potspline, sizeof(A.potspline));
  for(int i = 0; i < D * D; i++)
    A.bam_eps[i] = c->laws_set ? c->laws.par[i][1] : 1.31e-6f;
  A.cnt = (c->accumulator && c->counts_valid) ? c->wcnt : nullptr;
  if(latt)
    {
      A.potcorr = c->d_potcorr;
      A.potcorr_en = c->potcorr_en;
      A.fac_intp = 2.0 * c->potcorr_en / wp->boxsize;
      memcpy(A.potcorr_map, c->potcorr_tabmap, sizeof(A.potcorr_map));
    }
  const size_t smem = sr ? sizeof(float) * (size_t) A.ntables * A.ntab : 0;
  int grid = c->nsm * (latt ? 4 : POT_BLOCKS(D)), need = g2_cdiv(g2_cdiv(A.hi - A.lo, 32), WALK_WARPS);
  if(grid > need)
    grid = need;
  G2_CUDA(cudaMemsetAsync(A.work_counter, 0, sizeof(unsigned int), st));
  G2_CUDA(cudaMemsetAsync(c->d_smcount, 0, sizeof(unsigned int) * G2_CHUNK_COUNTERS, st));
  G2_CUDA(cudaEventRecord(c->ev[16], st));
  if(grid > 0)
    {
      const bool uneq = c->cfg.unequal_softenings != 0;
      switch (D)
	{
#ifndef G2_FAST_BUILD
	case 1: launch_pot<1>(c, A, grid, smem, sr, uneq, latt); break;
	case 3: launch_pot<3>(c, A, grid, smem, sr, uneq, latt); break;
	case 5: launch_pot<5>(c, A, grid, smem, sr, uneq, latt); break;
	case 6: launch_pot<6>(c, A, grid, smem, sr, uneq, latt); break;
#endif
	case 2: launch_pot<2>(c, A, grid, smem, sr, uneq, latt); break;
	case 4: launch_pot<4>(c, A, grid, smem, sr, uneq, latt); break;
	default: return g2_fail(G2GPU_ERR_ARG, "unsupported N_GRAVS %d", D);
	}
      c->launches++;
    }
  G2_CUDA(cudaEventRecord(c->ev[17], st));
  G2_CUDA(cudaGetLastError());
  G2_CUDA(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, c->ev[16], c->ev[17]);
  c->pot_ms = ms;
  c->pot_valid = 1;
  return 0;
}

// ---------------------------------------------------------------- stand-alone pair potentials (tests) -----------
// out[i] = what the potential walk ADDS to pot for one pair without the TreePM table term: -PotentialFxns[tgt][src](pm, m, h, r, N) for
// r >= h, +PotentialSplines[tgt][src](pm, m, h, r, N) for r < h (forcetree.c:2732-2734, 3115-3118), in the kernel's own arithmetic.
__global__ void eval_pot_kernel(const PotArgs A, int n, int ij, const float *__restrict__ pm, const float *__restrict__ m, const float *__restrict__ r,
				const float *__restrict__ h, const int *__restrict__ nn, float *__restrict__ out)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n)
    out[i] = pot_term<false>(A, nullptr, ij, m[i], r[i] * r[i], h[i], false, pm[i], nn ? (float) nn[i] : 1.0f);
}

int g2_eval_potentials_standalone(g2gpu_ctx *c, int n, int tgt, int src, const float *pm, const float *m, const float *r, const float *h,
				  const int *nn, float *out)
{
  if(!c->potlaws_set)
    return g2_fail(G2GPU_ERR_LAW, "pair potential laws not set (g2gpu_set_potential_laws)");
  PotArgs A;
  memset(&A, 0, sizeof(A));
  memcpy(A.potfxn, c->potfxn, sizeof(A.potfxn));
  memcpy(A.potspline, c->potspline, sizeof(A.potspline));
  for(int i = 0; i < c->D * c->D; i++)
    A.bam_eps[i] = c->laws_set ? c->laws.par[i][1] : 1.31e-6f;
  float *d;
  int *dn = nullptr;
  const size_t fb = sizeof(float) * (size_t) n;
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
  eval_pot_kernel<<<g2_cdiv(n, 256), 256, 0, c->stream>>>(A, n, tgt * c->D + src, d, d + n, d + 2 * (size_t) n, d + 3 * (size_t) n, dn, d + 4 * (size_t) n);
  c->launches++;
  G2_CUDA(cudaStreamSynchronize(c->stream));
  G2_CUDA(cudaMemcpy(out, d + 4 * (size_t) n, fb, cudaMemcpyDeviceToHost));
  cudaFree(d);
  if(dn)
    cudaFree(dn);
  return 0;
}
