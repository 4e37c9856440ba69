// g2_walk_common.cuh — what the two walk kernels (g2_walk.cu: one cursor per 32 targets; g2_walkb.cu: one warp per target)
// share: kernel arguments, the periodic wrap, and one species term of an interaction.
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

struct WalkArgs
{
  const float4 *__restrict__ cells;
  const float4 *__restrict__ wpart;
  const unsigned int *__restrict__ targets;	// sorted positions of active targets
  const unsigned int *__restrict__ tq;		// particle index of a sorted position
  const G2PRec *__restrict__ prec;
  const float *__restrict__ gravpm;
  const unsigned int *__restrict__ cnt;	// particle counts per species of every cell [U][D] (NGRAVS_ACCUMULATOR), or null
  const float *__restrict__ srtable;		// unique tables, NTAB floats each
  const float *__restrict__ latt;		// lattice-sum correction of every target (periodic box without PM, g2_lattice.cu), or null
  const float *__restrict__ lattcost;
  float *__restrict__ acc;
  float *__restrict__ cost;
  float *__restrict__ oldacc_out;
  unsigned long long *__restrict__ counters;
  unsigned int *__restrict__ work_counter;
  int lo, hi;			// slice of targets
  int numnodes;
  int ntab;
  int ntables;			// unique short-range tables held in shared memory
  float theta2;			// ErrTolTheta^2, 0 => relative criterion
  float errtol;			// ErrTolForceAcc
  float boxsize, boxinv;
  float rcut, rcut2, asmthfac, utor2wpi;
  float cull_margin;		// absolute safety margin of the geometric cull shortcut (float rounding of positions)
  float shift_len_max;		// TreePM: cells smaller than this take the periodic image of their centre for all their points
  double G, pos_fac_pre_g, pos_fac_post_g;
  int use_gravpm;
  float fsoft[6];
  int t2g[6];
  unsigned char tabmap[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];	// [tgt*D+src] -> unique table
  G2LawTable laws;
  // level-order (children of a cell are contiguous) structure-of-arrays copy of the walk records, for g2_walkb.cu
  const float4 *__restrict__ bq0;	// [V] len, centre
  const float4 *__restrict__ bs;	// [g * bstride + V] centre of mass and mass of species g
  const uint4 *__restrict__ bw;		// [V] first child V | #child cells << 28,  particle offset | #particles << 28,  pinfo,  hmax
  const unsigned char *__restrict__ bptype;	// particle type of every record of wpart
  unsigned int bstride;
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

// one species term of an interaction: adds d*fac to the lane's FP32 partial sums; returns whether it counted
// (mass != 0 and, for TreePM, r inside the short-range table: forcetree.c:1553-1582, 1958-2026).
// Stock wiring (Newton + Plummer spline for every pair, ngravs.c:109-116): branch-free except for the rare r < h case.
template <bool SR, bool STOCK>
__device__ __forceinline__ bool pair_term(const WalkArgs &A, const float *__restrict__ s_tab, unsigned int s_tab_addr, int tg, int sg, int D,
					  float pmass, float m, float dx, float dy, float dz, float r2, float h, float &fx, float &fy, float &fz,
					  float nn = 1.0f)	// N of forcetree.c:1563-1577 (only the non-stock laws read it)
{
  const float rinv = fast_rsqrt(fmaxf(r2, 1.0e-37f));
  const float r = r2 * rinv;
  if(STOCK)
    {
      bool counted = m != 0.0f;
      float fac;
      if(SR)
	{
	  int tabindex = (int) (A.asmthfac * r);	// forcetree.c:1962
	  counted = counted && tabindex < A.ntab;
	  tabindex = min(tabindex, A.ntab - 1);
	  const float t = lds_f32(s_tab_addr + 4u * (unsigned int) tabindex);
	  // (m/r^2 - m*utor2wpi*tab) / r   (forcetree.c:1972-1974)
	  fac = m * rinv * fmaf(-A.utor2wpi, t, rinv * rinv);
	}
      else
	fac = m * rinv * rinv * rinv;
      if(r < h)			// inside the softening: spline (rare)
	fac = law_plummer(m, h, r);
      fac = counted ? fac : 0.0f;
      fx = fmaf(dx, fac, fx);
      fy = fmaf(dy, fac, fy);
      fz = fmaf(dz, fac, fz);
      return counted;
    }
  else
    {
      if(m == 0.0f)
	return false;
      float fac;
      const int ij = tg * D + sg;
      if(SR)
	{
	  int tabindex = (int) (A.asmthfac * r);
	  if(tabindex >= A.ntab)
	    return false;
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
      return true;
    }
}

// g2_walkb.cu
int g2_launch_walkb(g2gpu_ctx *c, const WalkArgs &A, bool sr, bool periodic, bool unequal, bool stock);
