// g2_walk_common.cuh — what the walk kernel (g2_walk_kernel.cuh) and its driver (g2_walk.cu) share: kernel arguments, the periodic
// wrap, and one species term of an interaction.
#pragma once
#include "g2_common.cuh"
#include "g2_laws.cuh"


#define WALK_THREADS 128
#define WALK_WARPS (WALK_THREADS / 32)
#ifndef WALK_MINBLOCKS
#define WALK_MINBLOCKS 8	// 64 registers: measured 4 % faster than 9 blocks / 56 registers once the table base and the no-wrap bound stay in registers
#endif
#ifndef WALK_WIDE_D
#define WALK_WIDE_D 3		// from this many species on the per-species registers no longer fit 64 registers without spills
#endif
#ifndef WALK_MINBLOCKS_WIDE
#define WALK_MINBLOCKS_WIDE 8
#endif
// walk_kernel's own CTA size (WALK_THREADS is shared with the potential and lattice walks).  Its warps never meet at a barrier after the table load,
// so the CTA size only sets how many copies of the short-range table and how much shared memory an SM holds: the rest of the 256 KB array is
// L1 for cell records.  The resident warps per SM stay the same (WALK_MINBLOCKS x WALK_THREADS threads).
#ifndef WALK_KT
#define WALK_KT 1024	// MEASURED (B200, 256^3, profiles/experiments/r2_run31.sh): 128 / 256 / 512 / 1024 threads -> walk 177.8 / 177.7 / 175.6 / 172.9 ms
#endif
#define WALK_KWARPS (WALK_KT / 32)
#define WALK_KBLOCKS(minblocks) ((minblocks) * WALK_THREADS / WALK_KT)

// Work distribution of the walks (walk_kernel, pot_kernel, lattice_kernel): one chunk of 32 consecutive targets per warp.  The chunks are dealt out in
// one contiguous block per SM, so that the warps resident on an SM walk neighbouring targets at the same time and find each other's cell records in
// L1 (MEASURED, B200, 256^3: L1 hit rate 43 -> 64 %, walk 185.6 -> 177.7 ms); an SM that has finished its block takes chunks from the blocks of the
// following SMs (each block has its own counter).  sm_counter == nullptr: one global counter (option walk_sm_local = 0).
#define G2_CHUNK_COUNTERS 1024	// >= the largest %nsmid
// next chunk of the calling warp (lane 0 calls), >= nchunks when all are taken.  `steal` (how many blocks this warp has seen exhausted) is the only
// state a warp keeps between calls: everything else is re-derived per chunk, once per ~60 000 warp instructions of walking.
__device__ __forceinline__ unsigned int chunk_dealer_next(unsigned int *sm_counter, unsigned int *global_counter, int nsm_i, unsigned int nchunks, unsigned int &steal)
{
  if(!sm_counter)
    return atomicAdd(global_counter, 1u);
  unsigned int smid;
  asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
  const unsigned int nsm = (unsigned int) nsm_i, mysm = smid % nsm, blk = (nchunks + nsm - 1u) / nsm;
  for(; steal < nsm; steal++)
    {
      const unsigned int sm = mysm + steal < nsm ? mysm + steal : mysm + steal - nsm;
      const unsigned int t = atomicAdd(sm_counter + sm, 1u);
      if(t < blk && sm * blk + t < nchunks)
	return sm * blk + t;
    }
  return 0xffffffffu;
}

// (G2_SLICE_*: g2_common.cuh)

// what the FP64 re-evaluation of a borderline decision needs (walk_decide_exact): kept in device memory, so that the rarely called
// non-inlined functions take one pointer instead of a copy of the kernel arguments
struct WalkExactParams
{
  double boxsize, rcut, rcut2, theta, errtol, asmthfac;
  double fsoft[6];
  int ntab;
};

struct WalkArgs
{
  const float4 *__restrict__ cells;
  const float4 *__restrict__ wpart;
  const unsigned int *__restrict__ targets;	// particle index (current order) of every active target, in merged Peano-Hilbert order
  const int *__restrict__ slice;		// [G2_SLICE_*]: number of targets, this rank's [lo, hi)
  const G2PRec *__restrict__ prec;
  const float *__restrict__ gravpm;
  const unsigned int *__restrict__ cnt;	// particle counts per species of every cell [U][D] (NGRAVS_ACCUMULATOR), or null
  const float *__restrict__ srtable;		// unique tables, NTAB floats each
  const float *__restrict__ latt;		// lattice-sum correction of every target (periodic box without PM, g2_lattice.cu), or null
  const float *__restrict__ lattcost;
  float *__restrict__ acc;			// results by particle index (current order) ...
  float *__restrict__ cost;
  float *__restrict__ oldacc_out;
  float *__restrict__ cres;			// ... or, when not null, compact: 5 floats (acc[3], cost, oldacc) per target of the slice, target order
  // zero-copy results (g2gpu_group_gravity_tree with pinned result arrays): device-visible pointers to the CALLER's host arrays, indexed by
  // particle like acc / cost / oldacc_out; the epilogue stores there as well, so no download and no host scatter follow the walk
  float *__restrict__ zc_acc, *__restrict__ zc_cost, *__restrict__ zc_oldacc;
  G2ZcAos zc_aos;		// ... or to the caller's array of structures (the reference's P[]): base == nullptr when unused
  unsigned long long *__restrict__ counters;
  unsigned int *__restrict__ work_counter;
  unsigned int *__restrict__ sm_counter;	// one chunk counter per SM (ChunkDealer), nullptr: the global work_counter only
  int nsm;
  int numnodes;
  int ntab;
  int ntables;			// unique short-range tables held in shared memory
  float theta2;			// ErrTolTheta^2, 0 => relative criterion
  float errtol;			// ErrTolForceAcc
  float boxsize, boxinv;
  float rcut, rcut2, asmthfac, utor2wpi;
  float rmax2;			// (NTAB / asmthfac)^2: a term counts while r^2 is below it (tabindex < NTAB, forcetree.c:1962-1967)
  float ntabm1f;		// (float) (NTAB - 1)
  float cull_margin;		// absolute safety margin of the geometric cull shortcut (float rounding of positions); +inf disables the shortcut
  float shift_len_max;		// TreePM: cells smaller than this take the periodic image of their centre for all their points
  // guard bands of the FP32 decisions: G2_TOL_* in g2_walk_kernel.cuh (relative), and
  float rmax2_border;		// absolute, around rmax2
  double G, pos_fac_pre_g, pos_fac_post_g;
  const WalkExactParams *__restrict__ ex;
  unsigned int *__restrict__ redo_list;	// slice ordinals of the targets whose FP32 walk met a comparison inside its guard band
  unsigned int *__restrict__ redo_count;
  unsigned int redo_cap;
  int use_gravpm;
  int exact;			// 0: no FP64 re-evaluation (FP32 decisions only)
  unsigned int flush_mask;	// the FP32 partial sums are flushed into the accumulators when the warp descends at a cell with (index & mask) == 0
  float fsoft[6];
  int t2g[6];
  unsigned char tabmap[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];	// [tgt*D+src] -> unique table
  G2LawTable laws;
};

// NEAREST(x) (forcetree.c:43): x > L/2 -> x - L, x < -L/2 -> x + L.  For |x| < 1.5 L this equals x - L*rint(x/L)
// (round-half-even leaves x = +-L/2 untouched, like the strict comparisons of the macro): 3 instructions.
// rint() for |t| < 2^22 on the FMA pipe (two adds with 1.5*2^23) instead of FRND, which runs on the quarter-rate XU pipe
__device__ __forceinline__ float rint_small(float t)
{
  return __fadd_rn(__fadd_rn(t, 12582912.0f), -12582912.0f);
}

template <bool PERIODIC>
__device__ __forceinline__ float nearest(float x, float boxsize, float boxinv)
{
  if(PERIODIC)
    x = fmaf(-boxsize, rint_small(x * boxinv), x);
  return x;
}

// NEAREST() as the reference evaluates it, in double
__device__ __forceinline__ double nearest_dd(double x, double box)
{
  const double half = 0.5 * box;
  return x > half ? x - box : (x < -half ? x + box : x);
}

__device__ __forceinline__ float fast_rsqrt(float x)
{
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ float lds_f32(unsigned int saddr)
{
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(saddr));
  return v;
}

__device__ __forceinline__ float fmax3(float a, float b, float c) { return fmaxf(fmaxf(a, b), c); }

// one species term of an interaction: adds d*fac to the lane's FP32 partial sums.  `inrange` = the term lies inside the short-range
// table (tabindex < NTAB, forcetree.c:1962-1967), decided by the caller; `m != 0` is checked by the caller for node terms
// (forcetree.c:1553, 1992).  Stock wiring (Newton + Plummer spline for every pair, ngravs.c:109-116): branch-free except for the
// rare r < h case.
// SPL = false: the caller knows that r >= h (no softening spline needed).  asmthfac, ntabm1: A.asmthfac and A.ntabm1f, which the walk kernel keeps
// in registers (re-reading them from the constant bank costs an instruction per term).
template <bool SR, bool STOCK, bool SPL = true>
__device__ __forceinline__ void pair_term(const WalkArgs &A, const float *__restrict__ s_tab, unsigned int s_tab_addr, int tg, int sg, int D,
					  float pmass, float m, float dx, float dy, float dz, float r2, float h, bool counted, float &fx, float &fy, float &fz,
					  float asmthfac, float ntabm1, float nn = 1.0f)	// nn: N of forcetree.c:1563-1577 (only the non-stock laws read it)
{
  const float rinv = fast_rsqrt(fmaxf(r2, 1.0e-37f));
  const float r = r2 * rinv;
  if(STOCK)
    {
      float fac;
      if(SR)
	{
	  const int tabindex = (int) fminf(asmthfac * r, ntabm1);	// forcetree.c:1962 (clamped: terms beyond the table are not counted)
	  const float t = lds_f32(s_tab_addr + 4u * (unsigned int) tabindex);
	  // (m/r^2 - m*utor2wpi*tab) / r   (forcetree.c:1972-1974); the stock kernel's copy of the table in shared memory holds utor2wpi * tab
	  fac = m * rinv * (rinv * rinv - t);
	}
      else
	fac = m * rinv * rinv * rinv;
      if(SPL && r < h)		// inside the softening: spline (rare)
	fac = law_plummer(m, h, r);
      fac = counted ? fac : 0.0f;
      fx = fmaf(dx, fac, fx);
      fy = fmaf(dy, fac, fy);
      fz = fmaf(dz, fac, fz);
    }
  else
    {
      if(!counted)
	return;
      float fac;
      const int ij = tg * D + sg;
      if(SR)
	{
	  const int tabindex = min((int) (A.asmthfac * r), A.ntab - 1);
	  if(r >= h)
	    {
	      float a = accel_over_r(A.laws.accel[ij], A.laws.par[ij], pmass, m, r2, r, rinv, nn) * r;
	      float t = s_tab[(int) A.tabmap[ij] * A.ntab + tabindex];
	      fac = (a - m * A.utor2wpi * t) * rinv;
	    }
	  else
	    fac = accel_spline(A.laws.spline[ij], A.laws.par[ij], pmass, m, h, r, nn);
	}
      else
	{
	  if(r >= h)
	    fac = accel_over_r(A.laws.accel[ij], A.laws.par[ij], pmass, m, r2, r, rinv, nn);
	  else
	    fac = accel_spline(A.laws.spline[ij], A.laws.par[ij], pmass, m, h, r, nn);
	}
      fx = fmaf(dx, fac, fx);
      fy = fmaf(dy, fac, fy);
      fz = fmaf(dz, fac, fz);
    }
}

// per-D launchers (g2_walk_dN.cu, one translation unit per N_GRAVS so that they compile in parallel)
int g2_launch_walk_d1(g2gpu_ctx *c, const WalkArgs &A, int grid, size_t smem, bool sr, bool periodic, bool unequal, bool stock, int accd, int stats);
int g2_launch_walk_d2(g2gpu_ctx *c, const WalkArgs &A, int grid, size_t smem, bool sr, bool periodic, bool unequal, bool stock, int accd, int stats);
int g2_launch_walk_d3(g2gpu_ctx *c, const WalkArgs &A, int grid, size_t smem, bool sr, bool periodic, bool unequal, bool stock, int accd, int stats);
int g2_launch_walk_d4(g2gpu_ctx *c, const WalkArgs &A, int grid, size_t smem, bool sr, bool periodic, bool unequal, bool stock, int accd, int stats);
int g2_launch_walk_d5(g2gpu_ctx *c, const WalkArgs &A, int grid, size_t smem, bool sr, bool periodic, bool unequal, bool stock, int accd, int stats);
int g2_launch_walk_d6(g2gpu_ctx *c, const WalkArgs &A, int grid, size_t smem, bool sr, bool periodic, bool unequal, bool stock, int accd, int stats);

