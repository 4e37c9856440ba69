// g2_walk_kernel.cuh — the tree walk of force_treeevaluate (forcetree.c:1244-1610) and force_treeevaluate_shortrange
// (forcetree.c:1623-2052), with the gravity_tree epilogue (gravtree.c:304-358).  Included by g2_walk_dN.cu (one per N_GRAVS).
//
// One warp takes 32 targets that are adjacent along the Peano-Hilbert curve of ALL species (hence compact in space: 10 % fewer
// cell visits than 32 neighbours in depth-first tree order) and walks the depth-first
// cell array once for all of them: at every cell each lane takes ITS OWN decision (cull / accept / open) with the reference's
// criteria; the warp descends if any lane opens (ballot), otherwise jumps to the cell's sibling.  A lane that accepted or culled a
// cell sleeps until the warp's cursor has left that cell's subtree (cells are in depth-first order, so this is one integer
// compare).  Direct particle children of an opened cell sit in a contiguous group and are applied by the opening lanes only.
// Per-lane interaction lists are therefore those of the reference walk; only the summation order differs.
//
// Decisions are taken in FP32.  The reference decides in double (forcetree.c:1628-1631), so a target that meets a comparison whose
// two sides are closer than the FP32 rounding of their inputs allows (guard bands in WalkArgs) is flagged -- a running minimum per
// visit --, and walk_redo_kernel walks the tree again for the flagged targets (about one per thousand) in the reference's own double
// arithmetic, operation by operation: every target's interaction list, and with it GravCost, is then the reference's exactly.
// Keeping the FP64 path out of this kernel keeps its 64 registers free of spills.
//
// Node records are (2+D) x 16 B, fetched with 128-bit loads (all lanes read the same address: one L1 broadcast).
#pragma once
#include "g2_walk_common.cuh"

// guard bands of the FP32 decisions, each a few times the FP32 rounding its comparison can accumulate: a coordinate difference against a length
// 2e-7 -> 4e-7; an r^2 (three products of rounded differences) 3.6e-7 -> 6e-7; the opening criteria (r^4 a against M len^2) 1.1e-6 -> 1.5e-6.
// Where NEAREST is applied per point the raw difference is rounded at box scale: 1e-4.  Compile-time constants: immediates in the FFMAs.
#define G2_TOL_POS 4.0e-7f
#define G2_TOL_R2 6.0e-7f
#define G2_TOL_CRIT 1.5e-6f
#define G2_TOL_WIDE 1.0e-4f
#define G2_DEC_CULL 0
#define G2_DEC_ACCEPT 1
#define G2_DEC_OPEN 2

// ---- the reference's decision on one node for one target, in its own arithmetic (double locals, FLOAT node fields):
//      forcetree.c:1790-1926 (TreePM) and 1355-1501 (tree only).  Rare path: called for borderline comparisons only. ----
template <int D, bool SR, bool PERIODIC, bool UNEQUAL>
__device__ __forceinline__ int walk_decide_exact(const WalkExactParams *__restrict__ E, const float4 *__restrict__ rec, const G2PRec *__restrict__ target)
{
  const G2PRec tp = *target;
  const float oldacc = tp.oldacc;
  const int ptype = tp.type;
  const double px = (double) tp.x, py = (double) tp.y, pz = (double) tp.z, box = E->boxsize;
  const float4 q0 = __ldg(rec);
  const uint4 w = __ldg((const uint4 *) (rec + 1 + D));
  const double len = (double) q0.x;
  double r2min = 1.0e300, r2max = -1.0e300, summass = 0.0;
#pragma unroll
  for(int g = 0; g < D; g++)
    {
      const float4 q = __ldg(rec + 1 + g);
      double dx = (double) q.x - px, dy = (double) q.y - py, dz = (double) q.z - pz;
      if(PERIODIC)
	{
	  dx = nearest_dd(dx, box);
	  dy = nearest_dd(dy, box);
	  dz = nearest_dd(dz, box);
	}
      const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
      summass = __dadd_rn(summass, (double) q.w);
      if(r2 < r2min)
	r2min = r2;
      if(r2 > r2max)
	r2max = r2;
    }
  if(SR && r2min > E->rcut2)
    {				// forcetree.c:1828-1862
      const double eff = __dadd_rn(E->rcut, __dmul_rn(0.5, len));
      double d0 = (double) q0.y - px, d1 = (double) q0.z - py, d2 = (double) q0.w - pz;
      if(PERIODIC)
	{
	  d0 = nearest_dd(d0, box);
	  d1 = nearest_dd(d1, box);
	  d2 = nearest_dd(d2, box);
	}
      if(d0 < -eff || d0 > eff || d1 < -eff || d1 > eff || d2 < -eff || d2 > eff)
	return G2_DEC_CULL;
    }
  const double theta = E->theta;
  if(theta != 0.0)
    {				// forcetree.c:1865-1873
      if((double) __fmul_rn(q0.x, q0.x) > __dmul_rn(__dmul_rn(r2min, theta), theta))	// nop->len * nop->len is a FLOAT product (forcetree.c:1867)
	return G2_DEC_OPEN;
    }
  else
    {				// forcetree.c:1874-1897
      const double aold = __dmul_rn(E->errtol, (double) oldacc);
      if(__dmul_rn(__dmul_rn(summass, len), len) > __dmul_rn(__dmul_rn(r2min, r2min), aold))
	return G2_DEC_OPEN;
      const double lim = __dmul_rn(0.60, len);
      if(fabs((double) q0.y - px) < lim && fabs((double) q0.z - py) < lim && fabs((double) q0.w - pz) < lim)
	return G2_DEC_OPEN;
    }
  if(UNEQUAL)
    {				// forcetree.c:1899-1926
      const int maxsofttype = (int) (w.z >> 29);
      if(maxsofttype == 7)
	return G2_DEC_OPEN;
      double h = E->fsoft[ptype];
      if(h < E->fsoft[maxsofttype])
	{
	  h = E->fsoft[maxsofttype];
	  if(r2max < __dmul_rn(h, h) && ((w.z >> 28) & 1))
	    return G2_DEC_OPEN;
	}
    }
  return G2_DEC_ACCEPT;
}

// tabindex = (int) (asmthfac * sqrt(r2)) < NTAB in the reference's arithmetic (forcetree.c:1958-1967, 1996-2000)
template <bool PERIODIC>
__device__ __forceinline__ bool term_in_range_exact(const WalkExactParams *__restrict__ E, float sx, float sy, float sz, float pxf, float pyf, float pzf)
{
  double dx = (double) sx - (double) pxf, dy = (double) sy - (double) pyf, dz = (double) sz - (double) pzf;
  if(PERIODIC)
    {
      dx = nearest_dd(dx, E->boxsize);
      dy = nearest_dd(dy, E->boxsize);
      dz = nearest_dd(dz, E->boxsize);
    }
  const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  return (int) __dmul_rn(E->asmthfac, __dsqrt_rn(r2)) < E->ntab;
}

// periodic image shift of one cell for one target, split so that both partial differences are exact (walk_visit_cell)
struct WalkShift
{
  float sx, sy, sz;		// subtracted from the source coordinate: max(shift, 0)
  float tx, ty, tz;		// target coordinate + min(shift, 0)
};

// per-lane state of a walk
struct WalkLane
{
  float px, py, pz, pmass, aold, hself;
  int tg;
  float fx, fy, fz;		// FP32 partial sums, flushed into the accumulators whenever the warp descends
  int ninter, nterms, ndec;
  float bmin;			// EXACT: smallest distance of any comparison of this walk from its guard band (negative: inside)
  unsigned int skip_until;
  unsigned int bits;		// DEFER: ring entries (staged sources) this lane still has to evaluate
  // kernel arguments the visit code reads at every term, copied once per chunk through an addition ptxas cannot fold away: it then keeps them in
  // uniform registers instead of re-reading the constant bank at every use (LDCU + MOV per term)
  float k_rmax2, k_rmax2b, k_asmthfac, k_ntabm1;	// (ptxas keeps these three in uniform registers; for the others it re-reads the constant AND adds the zero)
  // (MEASURED, profiles/r2_walk_microvariants.txt: the same for rcut, rcut^2 and the cull margin removes four LDC per decision and is 1.4 ms SLOWER)
};

// DEFER (stock pair laws): the terms a lane accepts are not evaluated where they are found -- a handful of lanes at a time -- but staged.
// The SOURCE records (species centres of mass of an accepted cell, direct particles of an opened cell) go once per warp into a ring of
// WALK_RING entries in shared memory; a lane only sets a bit per entry it has to evaluate.  When the ring is nearly full every lane
// walks its own bits (ring_flush): lanes are busy for as many passes as the busiest lane has bits, about 45 % of the lane slots
// against 20 % where terms are evaluated in place (counts in DESIGN.md, section 5).  Summation order is a
// function of the traversal only, so results stay reproducible bit for bit.
#define WALK_RING 32
struct WalkRing
{
  float4 *src;			// shared memory, this warp's WALK_RING entries: x, y, z, mass
  float *hsrc;			// (UNEQUAL) softening that comes with the source: ForceSoftening[type] of a particle, of maxsofttype for a cell
  unsigned int fill;		// entries staged                                        } warp-uniform
  unsigned int isp;		// bit e: entry e is a particle (its in-range test and count are taken at evaluation)
  bool wrap;			// an entry was staged by a visit that has to consider periodic images
};

// One visit, first half: this lane's decision on the cell at the cursor and, if it accepts the cell, the D species terms.
// WRAP = false: no periodic image can matter for this cell and this warp's targets (non-periodic run, or a TreePM cell smaller
// than the warp's no-wrap bound): plain differences.  WRAP = true: per-cell image shift for small TreePM cells, NEAREST per point
// otherwise.  Returns whether the lane opens the cell; shx/shy/shz/small_cell are handed on to the particle half.
// EXACT: comparisons inside their guard band flag the target for walk_redo_kernel (L.bmin).
template <int D, bool SR, bool PERIODIC, bool UNEQUAL, bool STOCK, bool WRAP, bool STATS, bool EXACT, bool DEFER = false, int CRIT = 0>
__device__ __forceinline__ bool walk_visit_cell(const WalkArgs &A, const float *__restrict__ s_tab, unsigned int s_tab_addr, const float4 *__restrict__ rec,
						const float4 q0, const uint4 w, unsigned int cur, WalkLane &L, WalkShift &S, bool &small_cell, unsigned int *newbits = nullptr)
{
  bool open = false, done = false, outside = false;
  // EXACT: smallest distance of a comparison of this visit from its guard band (negative: inside).  Kept as a running minimum instead of
  // predicates so that the flagging stays straight-line code (FADD + FMNMX per comparison).
  float bm = 3.0e38f;
  if(STATS)
    L.ndec++;
  // guard bands (EXACT): a few times the FP32 rounding a comparison can accumulate.  Wide where NEAREST is applied per
  // point (cells too large for the per-cell image shift, and periodic walks without PM): there the raw coordinate difference is
  // rounded at box scale before the wrap.
  const bool wide = WRAP && !(SR && q0.x < A.shift_len_max);
  const float tol_pos = wide ? G2_TOL_WIDE : G2_TOL_POS, tol_r2 = wide ? G2_TOL_WIDE : G2_TOL_R2, tol_crit = wide ? G2_TOL_WIDE : G2_TOL_CRIT;
  float dx[D], dy[D], dz[D], r2[D], mass[D];
  float r2min = 3.0e38f, r2max = -1.0f, summass = 0.0f;
  const float len = q0.x;
  const float cxr = q0.y - L.px, cyr = q0.z - L.py, czr = q0.w - L.pz;
  // TreePM: a cell that can interact with (or must be opened by) a target lies within rcut + len of it, so for
  // len < L/2 - rcut every point of the cell has the same periodic image as the cell centre; points of cells that
  // are culled anyway can only look farther away.  The image shift is then computed once per cell, not per point.
  small_cell = WRAP && SR && len < A.shift_len_max;	// uniform (a property of the cell)
  if(WRAP && SR)
    {
      // nearest image = source - target - shift, shift in {-L, 0, +L}.  Applied as (source - max(shift, 0)) - (target + min(shift, 0)):
      // a shift of +L belongs to a source near the upper face, one of -L to a target near it, so both partial results are small
      // and EXACT in FP32, and the final difference is rounded once to half an ulp of ITSELF -- like every unwrapped difference.
      const float shx = A.boxsize * rint_small(cxr * A.boxinv), shy = A.boxsize * rint_small(cyr * A.boxinv), shz = A.boxsize * rint_small(czr * A.boxinv);
      S.sx = fmaxf(shx, 0.0f); S.sy = fmaxf(shy, 0.0f); S.sz = fmaxf(shz, 0.0f);
      S.tx = L.px + fminf(shx, 0.0f); S.ty = L.py + fminf(shy, 0.0f); S.tz = L.pz + fminf(shz, 0.0f);
    }
  if(SR)
    {
      // forcetree.c:1828-1862 culls a node if r2min > rcut^2 AND the target is farther than rcut + len/2 from the
      // node centre on some axis.  All mass of a node lies inside its cube (up to float rounding of positions), so a
      // target that clears the cube by a small margin on one axis is farther than rcut from every centre of mass:
      // r2min > rcut^2 is then certain and the per-species distances need not be computed at all.
      const float d0 = fabsf(WRAP ? (q0.y - S.sx) - S.tx : cxr), d1 = fabsf(WRAP ? (q0.z - S.sy) - S.ty : cyr), d2 = fabsf(WRAP ? (q0.w - S.sz) - S.tz : czr);
      const float eff = fmaf(0.5f, len, A.rcut), t = fmax3(d0, d1, d2) - eff;
      outside = t > 0.0f;
      done = t > fmaf(1.0e-3f, len, A.cull_margin);
      if(EXACT)
	bm = fmaf(-tol_pos, eff, fabsf(t));
    }
  if(!done)
    {
#pragma unroll
      for(int g = 0; g < D; g++)
	{
	  const float4 q = __ldg(rec + 1 + g);
	  mass[g] = q.w;
	  summass += q.w;
	  if(!WRAP)
	    {
	      dx[g] = q.x - L.px;
	      dy[g] = q.y - L.py;
	      dz[g] = q.z - L.pz;
	    }
	  else if(small_cell)
	    {
	      dx[g] = (q.x - S.sx) - S.tx;
	      dy[g] = (q.y - S.sy) - S.ty;
	      dz[g] = (q.z - S.sz) - S.tz;
	    }
	  else
	    {
	      dx[g] = nearest<PERIODIC>(q.x - L.px, A.boxsize, A.boxinv);
	      dy[g] = nearest<PERIODIC>(q.y - L.py, A.boxsize, A.boxinv);
	      dz[g] = nearest<PERIODIC>(q.z - L.pz, A.boxsize, A.boxinv);
	    }
	  r2[g] = dx[g] * dx[g] + dy[g] * dy[g] + dz[g] * dz[g];
	  r2min = fminf(r2min, r2[g]);
	  r2max = fmaxf(r2max, r2[g]);
	}
      if(SR)
	{
	  const float u = r2min - A.rcut2;
	  if(outside)
	    done = u > 0.0f;
	  if(EXACT)		// (the box test only matters when r2min > rcut^2, the r2min test only outside the box: a superset is flagged)
	    bm = fminf(bm, outside ? fmaf(-tol_r2, A.rcut2, fabsf(u)) : 3.0e38f);
	}
    }
  if(!done)
    {
      // The difference of two FP32 coordinates is rounded once, to half an ulp OF THE DIFFERENCE (it is exact when the operands share a
      // binade), so an r^2 built from such differences, and the products compared below, carry a few 1e-7 of relative error whatever
      // the magnitude of the coordinates.
      if(CRIT == 0 && A.theta2 > 0.0f)	// (CRIT == 1: an instantiation for the relative criterion alone, without the run-time switch)
	{			// Barnes-Hut, forcetree.c:1437-1445
	  const float lhs = len * len, v = fmaf(-r2min, A.theta2, lhs);
	  open = v > 0.0f;
	  if(EXACT)
	    bm = fminf(bm, fmaf(-tol_crit, lhs, fabsf(v)));
	}
      else
	{			// relative criterion, forcetree.c:1446-1472
	  const float lhs = summass * len * len, v = fmaf(-(r2min * r2min), L.aold, lhs);
	  const float wv = fmaf(-0.60f, len, fmax3(fabsf(cxr), fabsf(cyr), fabsf(czr)));	// < 0: the target lies inside 0.6 len of the centre on every axis
	  open = v > 0.0f || wv < 0.0f;
	  if(EXACT)
	    bm = fminf(bm, fminf(fmaf(-tol_crit, lhs, fabsf(v)), fmaf(-tol_pos, len, fabsf(wv))));
	}
    }
  float h = L.hself;
  if(UNEQUAL && !done && !open)
    {				// forcetree.c:1475-1501; the record carries ForceSoftening[maxsofttype] (+inf and the
				// mixed-softening bit for an empty node, maxsofttype == 7, which is always opened)
      const float hnode = __uint_as_float(w.w);
      if(h < hnode)
	{
	  h = hnode;
	  if((w.z >> 28) & 1)
	    {
	      const float hv = fmaf(-h, h, r2max);
	      open = hv < 0.0f;
	      if(EXACT)
		bm = fminf(bm, fmaf(-tol_r2, r2max, fabsf(hv)));
	    }
	}
    }
  if(!open)
    {
      L.skip_until = w.x;	// sleep until the cursor leaves this subtree
      if(!done)
	{
	  bool any = false, cnt[D];
#pragma unroll
	  for(int g = 0; g < D; g++)
	    {
	      bool counted = mass[g] != 0.0f;	// forcetree.c:1553 / 1992
	      if(SR)
		{
		  const float tv = r2[g] - L.k_rmax2;
		  if(EXACT)
		    bm = fminf(bm, fabsf(tv) - L.k_rmax2b);
		  counted = counted && tv < 0.0f;
		}
	      cnt[g] = counted;
	      if(DEFER)
		*newbits |= (counted ? 1u : 0u) << g;	// evaluated from the ring (ring_flush)
	      any |= counted;
	      if(STATS)
		L.nterms += counted;
	    }
	  if(!SR || any)
	    L.ninter++;		// forcetree.c:1585 resp. 2031-2032
	  if(!DEFER)
	    {
	      // stock laws, TreePM: ONE test per visit whether any species term can fall inside the softening (r2min < h^2, rare) instead of one
	      // branch per term.  (Not for the walks without PM: nvcc 12.9 compiles the spline-free term of that instantiation to fac = 0 -- SASS checked,
	      // caught by tests/test_gpu_tree_walk.py --, so they keep the test per term.)
	      if(STOCK && SR && r2min >= h * h)
		{
#pragma unroll
		  for(int g = 0; g < D; g++)
		    // (no branch around a species without mass: cnt[g] is false for it, so its term is exactly zero -- the centre of an empty species
		    // is the finite cell centre, forcetree.c:678-683 -- and the branch, which the compiler cannot know to be warp-uniform, cost more
		    // than the term: MEASURED, B200, 256^3, profiles/r2_walk_microvariants.txt: 167.44 -> 161.17 ms, results bit-identical;
		    // -DG2_WALK_MASSBR restores it)
#ifdef G2_WALK_MASSBR
		    if(mass[g] != 0.0f)
#endif
		      pair_term<SR, STOCK, false>(A, s_tab, s_tab_addr, L.tg, g, D, L.pmass, mass[g], dx[g], dy[g], dz[g], r2[g], h, cnt[g], L.fx, L.fy, L.fz, L.k_asmthfac, L.k_ntabm1);
		}
	      else
		{
#pragma unroll
		  for(int g = 0; g < D; g++)
		    if(mass[g] != 0.0f)
		      {
			const float nn = (!STOCK && A.cnt) ? (float) __ldg(A.cnt + (size_t) cur * D + g) : 1.0f;
			pair_term<SR, STOCK>(A, s_tab, s_tab_addr, L.tg, g, D, L.pmass, mass[g], dx[g], dy[g], dz[g], r2[g], h, cnt[g], L.fx, L.fy, L.fz, L.k_asmthfac, L.k_ntabm1, nn);
		      }
		}
	    }
	}
    }
  if(EXACT)
    L.bmin = fminf(L.bmin, bm);	// negative: some comparison of this walk fell inside its guard band and the target is walked again in FP64
  return open;
}

// One visit, second half: the direct particle children of a cell, for the lanes that opened it.
template <int D, bool SR, bool PERIODIC, bool UNEQUAL, bool STOCK, bool WRAP, bool STATS, bool EXACT>
__device__ __forceinline__ void walk_visit_particles(const WalkArgs &A, const float *__restrict__ s_tab, unsigned int s_tab_addr, const uint4 w, bool open,
						     unsigned int t2g_packed, WalkLane &L, const WalkShift &S, bool small_cell)
{
  const unsigned int np = w.z & 15u;
  if(!open)
    return;			// (one branch for the whole group instead of one per particle)
  for(unsigned int j = 0; j < np; j++)
    {
      const float4 p = __ldg(A.wpart + w.y + j);
      int sg = 0;
      float h = L.hself;
      if(UNEQUAL || !STOCK)
	{
	  const int stype = (w.z >> (4 + 3 * j)) & 7;
	  sg = (t2g_packed >> (4 * stype)) & 7;
	  if(UNEQUAL)
	    h = fmaxf(h, A.fsoft[stype]);   // forcetree.c:1412-1415
	}
      float ddx, ddy, ddz;
      if(!WRAP)
	{
	  ddx = p.x - L.px;
	  ddy = p.y - L.py;
	  ddz = p.z - L.pz;
	}
      else if(small_cell)
	{
	  ddx = (p.x - S.sx) - S.tx;
	  ddy = (p.y - S.sy) - S.ty;
	  ddz = (p.z - S.sz) - S.tz;
	}
      else
	{
	  ddx = nearest<PERIODIC>(p.x - L.px, A.boxsize, A.boxinv);
	  ddy = nearest<PERIODIC>(p.y - L.py, A.boxsize, A.boxinv);
	  ddz = nearest<PERIODIC>(p.z - L.pz, A.boxsize, A.boxinv);
	}
      const float rr2 = ddx * ddx + ddy * ddy + ddz * ddz;
      bool counted = true;  // a particle term counts whatever its mass (forcetree.c:1958-1987)
      if(SR)
	{
	  const float tv = rr2 - L.k_rmax2;
	  counted = tv < 0.0f;
	  if(EXACT)
	    L.bmin = fminf(L.bmin, fabsf(tv) - L.k_rmax2b);
	}
      pair_term<SR, STOCK>(A, s_tab, s_tab_addr, L.tg, sg, D, L.pmass, p.w, ddx, ddy, ddz, rr2, h, counted, L.fx, L.fy, L.fz, L.k_asmthfac, L.k_ntabm1);
      L.ninter += counted;
      if(STATS)
	L.nterms += counted;
    }
}

// One whole visit of the warp at cell `cur` (decisions, vote, flush of the partial sums, particles of an opened cell); returns the
// next cursor.  The walk loop calls the WRAP or the wrap-free instantiation by a warp-uniform test, so that the wrap-free path carries
// no image-shift state at all.  ACC accumulators: FP64 in shared memory (acc_sh, one slot per thread) or FP32 in registers (acc_rg).
template <int D, bool SR, bool PERIODIC, bool UNEQUAL, bool STOCK, bool WRAP, bool STATS, bool EXACT, typename ACC, int CRIT = 0>
__device__ __forceinline__ unsigned int walk_visit(const WalkArgs &A, const float *__restrict__ s_tab, unsigned int s_tab_addr, unsigned int cur, const float4 *__restrict__ rec,
						   const float4 q0, const uint4 w, unsigned int t2g_packed, WalkLane &L, ACC (*acc_sh)[WALK_KT], ACC &ax, ACC &ay, ACC &az)
{
  bool open = false, small_cell = false;
  WalkShift S;
  if(cur >= L.skip_until)
    open = walk_visit_cell<D, SR, PERIODIC, UNEQUAL, STOCK, WRAP, STATS, EXACT, false, CRIT>(A, s_tab, s_tab_addr, rec, q0, w, cur, L, S, small_cell);
  // The openers as a warp-uniform ballot mask, so that no per-lane flag has to survive the vote (as a bool it lived byte-packed in a register:
  // three PRMT per visit).  MEASURED (B200, 256^3, profiles/r2_walk_microvariants.txt): 173.27 -> 170.68 ms.
  const unsigned int openers = __ballot_sync(0xffffffffu, open);
  if(openers == 0u)
    return w.x;			// nobody opens the cell: on to its sibling (the lanes that are awake stay awake: skip_until <= cur < sibling)
  open = (openers >> (threadIdx.x & 31u)) & 1u;
  small_cell = WRAP && SR && q0.x < A.shift_len_max;
  // FP32 partial sums go into the accumulators when the warp descends at a cell whose index has its low bits clear (A.flush_mask; with 0
  // at every descent, about every third visit): few conversions, bounded error, and flush points that depend on the traversal only
  // (=> reproducible bits)
  if((cur & A.flush_mask) == 0u)
    {
      if(sizeof(ACC) == 8)
	{
	  acc_sh[0][threadIdx.x] += (ACC) L.fx; acc_sh[1][threadIdx.x] += (ACC) L.fy; acc_sh[2][threadIdx.x] += (ACC) L.fz;
	}
      else
	{
	  ax += (ACC) L.fx; ay += (ACC) L.fy; az += (ACC) L.fz;
	}
      L.fx = L.fy = L.fz = 0.0f;
    }
  if((w.z & 15u) != 0u)
    walk_visit_particles<D, SR, PERIODIC, UNEQUAL, STOCK, WRAP, STATS, EXACT>(A, s_tab, s_tab_addr, w, open, t2g_packed, L, S, small_cell);
  return cur + 1u;
}

// DEFER: every lane evaluates the ring entries it holds a bit for.  WRAPF: some entry may need a periodic image; the image shift is
// taken per point and applied in two exact steps like in walk_visit_cell, so the difference is rounded once, at its own magnitude.
template <bool SR, bool UNEQUAL, bool STATS, bool EXACT, bool WRAPF>
__device__ __forceinline__ void ring_eval(const WalkArgs &A, const float *__restrict__ s_tab, unsigned int s_tab_addr, WalkLane &L, const WalkRing &R)
{
  unsigned int b = L.bits;
  while(b)
    {
      const int e = 31 - __clz((int) b);
      b ^= 1u << e;
      const float4 s = R.src[e];
      float h = L.hself;
      if(UNEQUAL)
	h = fmaxf(h, R.hsrc[e]);	// forcetree.c:1412-1415 (particles), 1475-1501 (cells)
      float dx, dy, dz;
      if(WRAPF)
	{
	  const float shx = A.boxsize * rint_small((s.x - L.px) * A.boxinv), shy = A.boxsize * rint_small((s.y - L.py) * A.boxinv),
	    shz = A.boxsize * rint_small((s.z - L.pz) * A.boxinv);
	  dx = (s.x - fmaxf(shx, 0.0f)) - (L.px + fminf(shx, 0.0f));
	  dy = (s.y - fmaxf(shy, 0.0f)) - (L.py + fminf(shy, 0.0f));
	  dz = (s.z - fmaxf(shz, 0.0f)) - (L.pz + fminf(shz, 0.0f));
	}
      else
	{
	  dx = s.x - L.px;
	  dy = s.y - L.py;
	  dz = s.z - L.pz;
	}
      const float r2 = dx * dx + dy * dy + dz * dz;
      bool counted = true;
      if(SR)
	{			// a species term of a cell was tested when it was staged; a particle term is tested here (forcetree.c:1958-1967)
	  const bool isp = (R.isp >> e) & 1u;
	  const float tv = r2 - L.k_rmax2;
	  counted = !isp || tv < 0.0f;
	  if(EXACT)
	    L.bmin = fminf(L.bmin, isp ? fabsf(tv) - L.k_rmax2b : 3.0e38f);
	  L.ninter += isp && counted;	// forcetree.c:2031
	  if(STATS)
	    L.nterms += isp && counted;
	}
      pair_term<SR, true>(A, s_tab, s_tab_addr, L.tg, 0, 1, L.pmass, s.w, dx, dy, dz, r2, h, counted, L.fx, L.fy, L.fz, L.k_asmthfac, L.k_ntabm1);
    }
  L.bits = 0u;
}

template <bool SR, bool UNEQUAL, bool STATS, bool EXACT, typename ACC>
__device__ __forceinline__ void ring_flush(const WalkArgs &A, const float *__restrict__ s_tab, unsigned int s_tab_addr, WalkLane &L, WalkRing &R,
					   ACC (*acc_sh)[WALK_KT], ACC &ax, ACC &ay, ACC &az)
{
  __syncwarp();			// the entries were written by other lanes
  if(R.wrap)
    ring_eval<SR, UNEQUAL, STATS, EXACT, true>(A, s_tab, s_tab_addr, L, R);
  else
    ring_eval<SR, UNEQUAL, STATS, EXACT, false>(A, s_tab, s_tab_addr, L, R);
  // the FP32 partial sums of one ring (at most WALK_RING terms per lane) go into the accumulators
  if(sizeof(ACC) == 8)
    {
      acc_sh[0][threadIdx.x] += (ACC) L.fx; acc_sh[1][threadIdx.x] += (ACC) L.fy; acc_sh[2][threadIdx.x] += (ACC) L.fz;
    }
  else
    {
      ax += (ACC) L.fx; ay += (ACC) L.fy; az += (ACC) L.fz;
    }
  L.fx = L.fy = L.fz = 0.0f;
  R.fill = 0u;
  R.isp = 0u;
  R.wrap = false;
  __syncwarp();			// nobody overwrites an entry that another lane still reads
}

// One whole visit with deferred terms: decisions, staging of the sources, vote; returns the next cursor.  The caller guarantees room for
// D + 8 entries in the ring.
template <int D, bool SR, bool PERIODIC, bool UNEQUAL, bool WRAP, bool STATS, bool EXACT>
__device__ __forceinline__ unsigned int walk_visit_defer(const WalkArgs &A, unsigned int cur, const float4 *__restrict__ rec, const float4 q0, const uint4 w,
							 int lane, WalkLane &L, WalkRing &R)
{
  bool open = false, small_cell = false;
  unsigned int newbits = 0u;
  WalkShift S;
  if(cur >= L.skip_until)
    open = walk_visit_cell<D, SR, PERIODIC, UNEQUAL, true, WRAP, STATS, EXACT, true>(A, nullptr, 0u, rec, q0, w, cur, L, S, small_cell, &newbits);
  if(__any_sync(0xffffffffu, newbits != 0u))
    {				// some lane uses the cell: stage its species records (lanes 0..D-1 copy one each)
      if(lane < D)
	{
	  R.src[R.fill + lane] = __ldg(rec + 1 + lane);
	  if(UNEQUAL)
	    R.hsrc[R.fill + lane] = __uint_as_float(w.w);
	}
      L.bits |= newbits << R.fill;
      R.fill += D;
      if(WRAP)
	R.wrap = true;
    }
  if(__ballot_sync(0xffffffffu, open) == 0u)
    return w.x;			// nobody opens the cell: on to its sibling
  const unsigned int np = w.z & 15u;
  if(np != 0u)
    {				// the direct particle children of the cell, for the lanes that opened it
      if((unsigned int) lane < np)
	{
	  R.src[R.fill + lane] = __ldg(A.wpart + w.y + lane);
	  if(UNEQUAL)
	    R.hsrc[R.fill + lane] = A.fsoft[(w.z >> (4 + 3 * lane)) & 7];
	}
      const unsigned int m = ((1u << np) - 1u) << R.fill;
      if(open)
	{
	  L.bits |= m;
	  if(!SR)
	    {			// without PM every particle of an opened cell counts (forcetree.c:1585)
	      L.ninter += (int) np;
	      if(STATS)
		L.nterms += (int) np;
	    }
	}
      R.isp |= m;
      R.fill += np;
      if(WRAP)
	R.wrap = true;
    }
  return cur + 1u;
}

// gravity_tree epilogue for one target: GravAccel is stored as FLOAT (forcetree.c:1592-1594), then gravtree.c:304-358
// The epilogue is a real CALL (not inlined), so that its branches -- zero-copy targets, AoS layouts, compact slices -- cannot change the register
// allocation and scheduling of the walk loop; it runs once per 32 targets, i.e. once per ~2 000 cell visits.  The kernel argument is
// __grid_constant__ so that the call can take its address without a local copy.  MEASURED (B200, 256^3, profiles/r2_walk_microvariants.txt):
// inlined 175.38 ms, as a call 173.35 ms (two runs each, +-0.03 ms); __grid_constant__ alone changes nothing.  -DG2_STORE_INLINE restores the inlined form.
#ifndef G2_STORE_INLINE
#define G2_STORE_ATTR __noinline__
#define G2_STORE_GRIDCONST __grid_constant__
#else
#define G2_STORE_ATTR __forceinline__
#define G2_STORE_GRIDCONST
#endif
template <bool SR, bool PERIODIC>
__device__ G2_STORE_ATTR void walk_store_result(const WalkArgs &A, unsigned int idx, int tloc, float px, float py, float pz, float fx, float fy, float fz, float ninter)
{
  if(PERIODIC && !SR && A.latt)
    {				// force_treeevaluate_lattice_correction adds to the FLOAT result and to GravCost (forcetree.c:2435-2438)
      fx = (float) ((double) fx + (double) A.latt[3 * (size_t) idx + 0]);
      fy = (float) ((double) fy + (double) A.latt[3 * (size_t) idx + 1]);
      fz = (float) ((double) fz + (double) A.latt[3 * (size_t) idx + 2]);
      if(A.lattcost)
	ninter += A.lattcost[idx];
    }
  if(A.pos_fac_pre_g != 0.0)
    {
      fx = (float) ((double) fx + A.pos_fac_pre_g * (double) px);
      fy = (float) ((double) fy + A.pos_fac_pre_g * (double) py);
      fz = (float) ((double) fz + A.pos_fac_pre_g * (double) pz);
    }
  double sx = (double) fx, sy = (double) fy, sz = (double) fz;
  if(A.use_gravpm)
    {
      sx += (double) A.gravpm[3 * (size_t) idx + 0] / A.G;
      sy += (double) A.gravpm[3 * (size_t) idx + 1] / A.G;
      sz += (double) A.gravpm[3 * (size_t) idx + 2] / A.G;
    }
  const float oldacc_new = (float) sqrt(sx * sx + sy * sy + sz * sz);
  fx = (float) ((double) fx * A.G);
  fy = (float) ((double) fy * A.G);
  fz = (float) ((double) fz * A.G);
  if(A.pos_fac_post_g != 0.0)
    {
      fx = (float) ((double) fx + A.pos_fac_post_g * (double) px);
      fy = (float) ((double) fy + A.pos_fac_post_g * (double) py);
      fz = (float) ((double) fz + A.pos_fac_post_g * (double) pz);
    }
  if(A.zc_acc)
    {				// straight into the caller's pinned host arrays (posted writes over PCIe, spread over the whole kernel)
      A.zc_acc[3 * (size_t) idx + 0] = fx;
      A.zc_acc[3 * (size_t) idx + 1] = fy;
      A.zc_acc[3 * (size_t) idx + 2] = fz;
      if(A.zc_cost)
	A.zc_cost[idx] = ninter;
      if(A.zc_oldacc)
	A.zc_oldacc[idx] = oldacc_new;
    }
  if(A.zc_aos.base)
    {				// ... or into the caller's array of structures
      char *q = A.zc_aos.base + (size_t) A.zc_aos.perm[idx] * A.zc_aos.stride;
      if(A.zc_aos.float_bytes == 4)
	{
	  float *ga = (float *) (q + A.zc_aos.off_acc);
	  ga[0] = fx; ga[1] = fy; ga[2] = fz;
	  if(A.zc_aos.off_old >= 0)
	    *(float *) (q + A.zc_aos.off_old) = oldacc_new;
	}
      else
	{
	  double *ga = (double *) (q + A.zc_aos.off_acc);
	  ga[0] = (double) fx; ga[1] = (double) fy; ga[2] = (double) fz;
	  if(A.zc_aos.off_old >= 0)
	    *(double *) (q + A.zc_aos.off_old) = (double) oldacc_new;
	}
      if(A.zc_aos.off_cost >= 0)
	*(float *) (q + A.zc_aos.off_cost) = ninter;	// (a float in either build, allvars.h:572)
    }
  if(A.cres)
    {				// compact results of the slice in target order (multi-GPU: only the slice travels to the host)
      float *o = A.cres + 5 * (size_t) tloc;
      o[0] = fx; o[1] = fy; o[2] = fz; o[3] = ninter; o[4] = oldacc_new;
    }
  else
    {
      A.oldacc_out[idx] = oldacc_new;
      A.acc[3 * (size_t) idx + 0] = fx;
      A.acc[3 * (size_t) idx + 1] = fy;
      A.acc[3 * (size_t) idx + 2] = fz;
      A.cost[idx] = ninter;
    }
}

template <int D, bool SR, bool PERIODIC, bool UNEQUAL, bool STOCK, typename ACC, bool STATS, bool EXACT, bool DEFER, int CRIT = 0>
__global__ void __launch_bounds__(WALK_KT, WALK_KBLOCKS(D >= WALK_WIDE_D ? WALK_MINBLOCKS_WIDE : WALK_MINBLOCKS)) walk_kernel(const G2_STORE_GRIDCONST WalkArgs A)
{
  extern __shared__ float s_tab[];
  __shared__ unsigned int s_chunk[WALK_KWARPS];
  __shared__ float4 s_ring[DEFER ? WALK_KWARPS * WALK_RING : 1];
  __shared__ float s_ringh[(DEFER && UNEQUAL) ? WALK_KWARPS * WALK_RING : 1];
  // the FP64 accumulators of a lane live in shared memory (touched only when the warp descends, about every third visit): six
  // registers less keeps the kernel at 64 registers without spills
  __shared__ ACC s_acc[sizeof(ACC) == 8 ? 3 : 1][WALK_KT];
  if(SR)
    {
      for(int i = threadIdx.x; i < A.ntables * A.ntab; i += WALK_KT)
	s_tab[i] = STOCK ? A.utor2wpi * A.srtable[i] : A.srtable[i];	// stock laws: the factor of forcetree.c:1972 goes into the table once
      __syncthreads();
    }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int RW = 2 + D;
  unsigned int s_tab_addr = (unsigned int) __cvta_generic_to_shared(s_tab);
  asm volatile("" : "+r"(s_tab_addr));	// keep the table base in a register: re-deriving it costs 4 uniform instructions per pair term
  unsigned int t2g_packed = 0;	// TypeToGrav as 6 nibbles
#pragma unroll
  for(int t = 0; t < 6; t++)
    t2g_packed |= (unsigned int) A.t2g[t] << (4 * t);
  const int lo = A.slice[G2_SLICE_LO], hi = A.slice[G2_SLICE_HI];
  const int nchunks = (hi - lo + 31) >> 5;
  const unsigned int end = (unsigned int) A.numnodes;
  unsigned long long tot_inter = 0, tot_visits = 0, tot_terms = 0, tot_dec = 0;

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
      const int tloc = (int) chunk * 32 + lane;	// target ordinal inside the slice
      const bool valid = lo + tloc < hi;
      unsigned int idx = 0;
      WalkLane L;
      L.px = L.py = L.pz = L.pmass = L.aold = 0.0f;
      int ptype = 1;
      if(valid)
	{
	  idx = A.targets[lo + tloc];
	  const G2PRec p = A.prec[idx];
	  L.px = p.x; L.py = p.y; L.pz = p.z; L.pmass = p.m;
	  ptype = p.type;
	  L.aold = A.errtol * p.oldacc;	// forcetree.c:1289
	}
      L.tg = (t2g_packed >> (4 * ptype)) & 7;
      L.hself = A.fsoft[ptype];
      L.fx = L.fy = L.fz = 0.0f;
      L.ninter = L.nterms = L.ndec = 0;
      L.bmin = 3.0e38f;
      // (a value ptxas cannot trace back to the constant bank -- threadIdx.x >> 10 is zero, which it does not know -- or it re-reads the
      // constant at every use: LDCU + MOV per term)
      const float opaque_zero = (float) (threadIdx.x >> 10);
      L.k_rmax2 = A.rmax2 + opaque_zero;
      L.k_rmax2b = A.rmax2_border + opaque_zero;
      L.k_asmthfac = A.asmthfac + opaque_zero;
      L.k_ntabm1 = A.ntabm1f + opaque_zero;
      L.bits = 0u;
      L.skip_until = valid ? 0u : 0xffffffffu;
      WalkRing R;
      R.src = s_ring + (DEFER ? warp * WALK_RING : 0);
      R.hsrc = s_ringh + ((DEFER && UNEQUAL) ? warp * WALK_RING : 0);
      R.fill = 0u;
      R.isp = 0u;
      R.wrap = false;
      if(sizeof(ACC) == 8)
	{
	  s_acc[0][threadIdx.x] = 0; s_acc[1][threadIdx.x] = 0; s_acc[2][threadIdx.x] = 0;
	}
      unsigned int iter = 0;
      unsigned int cur = __any_sync(0xffffffffu, valid) ? 0u : end;
      // TreePM: a target farther than rcut + len/2 (+ margins) from every face of the box needs no periodic image of a cell of size len:
      // cells across a face are culled with the raw distance as well (raw >= nearest-image distance >= distance to the face).  Cells
      // smaller than nowrap_len (warp minimum over the 32 targets, so uniform) take the no-wrap code path; results are unchanged.
      float nowrap_len = 0.0f;
      if(SR && PERIODIC)
	{
	  float m = valid ? fminf(fminf(fminf(L.px, A.boxsize - L.px), fminf(L.py, A.boxsize - L.py)), fminf(L.pz, A.boxsize - L.pz)) : 3.0e38f;
#pragma unroll
	  for(int o = 16; o > 0; o >>= 1)
	    m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
	  nowrap_len = fminf(A.shift_len_max, 1.99f * (m - A.rcut - 2.0f * A.cull_margin));
	}

      ACC rx = 0, ry = 0, rz = 0;	// (register accumulators of the FP32 variant)
      while(true)
	{
	  if(DEFER)
	    {			// one flush site: before a visit that might not find room for its D species records and 8 particles, and at the end
	      if(cur >= end || R.fill > (unsigned int) (WALK_RING - (D + 8)))
		{
		  ring_flush<SR, UNEQUAL, STATS, EXACT, ACC>(A, s_tab, s_tab_addr, L, R, s_acc, rx, ry, rz);
		  if(cur >= end)
		    break;
		}
	    }
	  else if(cur >= end)
	    break;
	  // (MEASURED, round 2, profiles/experiments/r2_run22.sh: a prefetch.global.L1 of the sibling's record here costs 1 ms at 256^3, and the
	  // mere presence of the option, a uniform branch, 5 ms; a shared-memory carveout of 100 KB instead of the driver's 132 KB gains 0.5 %.)
	  const float4 *rec = A.cells + (size_t) cur * RW;
	  const float4 q0 = __ldg(rec);
	  const uint4 w = __ldg((const uint4 *) (rec + 1 + D));
	  if(STATS)
	    iter++;
	  if(DEFER)
	    {
	      if(PERIODIC && !(SR && q0.x < nowrap_len))	// uniform
		cur = walk_visit_defer<D, SR, PERIODIC, UNEQUAL, PERIODIC, STATS, EXACT>(A, cur, rec, q0, w, lane, L, R);
	      else
		cur = walk_visit_defer<D, SR, PERIODIC, UNEQUAL, false, STATS, EXACT>(A, cur, rec, q0, w, lane, L, R);
	    }
	  else if(PERIODIC && !(SR && q0.x < nowrap_len))	// uniform
	    cur = walk_visit<D, SR, PERIODIC, UNEQUAL, STOCK, PERIODIC, STATS, EXACT, ACC, CRIT>(A, s_tab, s_tab_addr, cur, rec, q0, w, t2g_packed, L, s_acc, rx, ry, rz);
	  else
	    cur = walk_visit<D, SR, PERIODIC, UNEQUAL, STOCK, false, STATS, EXACT, ACC, CRIT>(A, s_tab, s_tab_addr, cur, rec, q0, w, t2g_packed, L, s_acc, rx, ry, rz);
	}

      const ACC ax = (sizeof(ACC) == 8 ? s_acc[0][threadIdx.x] : rx) + (ACC) L.fx, ay = (sizeof(ACC) == 8 ? s_acc[1][threadIdx.x] : ry) + (ACC) L.fy,
	az = (sizeof(ACC) == 8 ? s_acc[2][threadIdx.x] : rz) + (ACC) L.fz;
      if(STATS && lane == 0)
	tot_visits += iter;	// one cursor per warp: visits = loop trips
      if(valid)
	{
	  walk_store_result<SR, PERIODIC>(A, idx, tloc, L.px, L.py, L.pz, (float) ax, (float) ay, (float) az, (float) L.ninter);
	  tot_inter += (unsigned long long) L.ninter;
	  if(EXACT && L.bmin < 0.0f)
	    {			// to walk_redo_kernel
	      const unsigned int slot = atomicAdd(A.redo_count, 1u);
	      if(slot < A.redo_cap)
		A.redo_list[slot] = (unsigned int) tloc;
	    }
	  if(STATS)
	    {
	      tot_terms += (unsigned long long) L.nterms;
	      tot_dec += (unsigned long long) L.ndec;
	    }
	}
    }
  // statistics: interactions (= sum of GravCost), and with STATS cell visits (per cursor), species terms, decisions
#pragma unroll
  for(int o = 16; o > 0; o >>= 1)
    {
      tot_inter += __shfl_xor_sync(0xffffffffu, tot_inter, o);
      if(STATS)
	{
	  tot_terms += __shfl_xor_sync(0xffffffffu, tot_terms, o);
	  tot_dec += __shfl_xor_sync(0xffffffffu, tot_dec, o);
	  tot_visits += __shfl_xor_sync(0xffffffffu, tot_visits, o);
	}
    }
  if(lane == 0)
    {
      atomicAdd(&A.counters[0], tot_inter);
      if(STATS)
	{
	  atomicAdd(&A.counters[1], tot_visits);
	  atomicAdd(&A.counters[2], tot_terms);
	  atomicAdd(&A.counters[4], tot_dec);
	}
    }
}

// ---- the walk of a target whose FP32 walk met a comparison inside its guard band, again: ONE WARP per target, in the
//      reference's own arithmetic (double locals, FLOAT node and particle fields, decisions by walk_decide_exact, tabindex =
//      (int) (asmthfac * r): forcetree.c:1244-1610 and 1623-2052 restated over the depth-first records).  The depth-first array is
//      walked in disjoint index ranges, one per lane: lane 0 starts with the whole tree, and whenever a lane is idle a busy lane hands
//      it the part of its range that lies behind the sibling of its current cell.  Partial sums are FP64 and reduced at the end. ----
#define WALK_REDO_WARPS 4
template <int D, bool SR, bool PERIODIC, bool UNEQUAL, bool STOCK>
__global__ void __launch_bounds__(32 * WALK_REDO_WARPS) walk_redo_kernel(const G2_STORE_GRIDCONST WalkArgs A)
{
  extern __shared__ float s_tab[];
  if(SR)
    {
      for(int i = threadIdx.x; i < A.ntables * A.ntab; i += 32 * WALK_REDO_WARPS)
	s_tab[i] = A.srtable[i];
      __syncthreads();
    }
  const WalkExactParams *__restrict__ E = A.ex;
  const int R = 2 + D, lane = threadIdx.x & 31;
  const int lo = A.slice[G2_SLICE_LO];
  const unsigned int nredo = min(*A.redo_count, A.redo_cap), nodes = (unsigned int) A.numnodes;
  const unsigned int nwarps = gridDim.x * WALK_REDO_WARPS;
  for(unsigned int i = blockIdx.x * WALK_REDO_WARPS + (threadIdx.x >> 5); i < nredo; i += nwarps)
    {
      const int tloc = (int) A.redo_list[i];
      const unsigned int idx = A.targets[lo + tloc];
      const G2PRec tp = A.prec[idx];
      const double px = (double) tp.x, py = (double) tp.y, pz = (double) tp.z, box = E->boxsize;
      const int tg = A.t2g[tp.type];
      const double hself = E->fsoft[tp.type];
      double ax = 0.0, ay = 0.0, az = 0.0;
      int ninter = 0;
      unsigned int cur = lane == 0 ? 0u : nodes, end = nodes;	// this lane's range of the depth-first array
      while(true)
	{
	  const unsigned int busy = __ballot_sync(0xffffffffu, cur < end);
	  if(busy == 0u)
	    break;
	  uint4 w = make_uint4(end, 0u, 0u, 0u);
	  const float4 *rec = A.cells + (size_t) (cur < end ? cur : 0u) * R;
	  if(cur < end)
	    w = __ldg((const uint4 *) (rec + 1 + D));
	  // work hand-over: the k-th idle lane takes [sibling, end) from the k-th busy lane that has cells behind the sibling of its
	  // current cell; that lane keeps [cur, sibling)
	  const bool isbusy = cur < end, can_give = isbusy && w.x < end;
	  const unsigned int givers = __ballot_sync(0xffffffffu, can_give), idle = ~busy;
	  const int npairs = min(__popc(givers), __popc(idle));
	  if(npairs > 0)
	    {
	      const unsigned int below = (1u << lane) - 1u;
	      const int grank = __popc(givers & below), irank = __popc(idle & below);
	      const bool takes = !isbusy && irank < npairs;
	      unsigned int src_lane = 0;
	      if(takes)
		{
		  unsigned int m = givers;
		  for(int r = 0; r < irank; r++)
		    m &= m - 1u;
		  src_lane = (unsigned int) (__ffs(m) - 1);
		}
	      const unsigned int give_lo = __shfl_sync(0xffffffffu, w.x, src_lane), give_hi = __shfl_sync(0xffffffffu, end, src_lane);
	      if(takes)
		{
		  cur = give_lo;
		  end = give_hi;
		}
	      else if(can_give && grank < npairs)
		end = w.x;
	    }
	  if(!isbusy)
	    continue;		// idle lanes, also those that just took work: they load their record in the next round
	  const int dec = walk_decide_exact<D, SR, PERIODIC, UNEQUAL>(E, rec, A.prec + idx);
	  if(dec == G2_DEC_CULL)
	    {
	      cur = w.x;
	      continue;
	    }
	  const bool isnode = dec == G2_DEC_ACCEPT;
	  const int nsrc = isnode ? D : (int) (w.z & 15u);
	  double h = hself;
	  if(UNEQUAL && isnode)
	    {			// forcetree.c:1475-1501: the node's largest softening if it exceeds the target's
	      const int mst = (int) (w.z >> 29);
	      if(mst < 6 && h < E->fsoft[mst])
		h = E->fsoft[mst];
	    }
	  bool any = false;
	  for(int j = 0; j < nsrc; j++)
	    {
	      const float4 q = isnode ? __ldg(rec + 1 + j) : __ldg(A.wpart + w.y + j);
	      int sg = j;
	      double hh = h;
	      if(!isnode)
		{
		  const int stype = (w.z >> (4 + 3 * j)) & 7;
		  sg = A.t2g[stype];
		  if(UNEQUAL && hh < E->fsoft[stype])
		    hh = E->fsoft[stype];	// forcetree.c:1412-1415
		}
	      else if(q.w == 0.0f)
		continue;	// forcetree.c:1553 / 1992
	      double dx = (double) q.x - px, dy = (double) q.y - py, dz = (double) q.z - pz;
	      if(PERIODIC)
		{
		  dx = nearest_dd(dx, box);
		  dy = nearest_dd(dy, box);
		  dz = nearest_dd(dz, box);
		}
	      const double r2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
	      const double r = __dsqrt_rn(r2);
	      int tabindex = 0;
	      if(SR)
		{
		  tabindex = (int) __dmul_rn(E->asmthfac, r);	// forcetree.c:1962
		  if(tabindex >= E->ntab)
		    continue;
		}
	      else if(!isnode)
		ninter++;	// without PM every particle of an opened node counts (forcetree.c:1585)
	      const int ij = tg * D + sg;
	      const double m = (double) q.w;
	      const float nn = (isnode && !STOCK && A.cnt) ? (float) __ldg(A.cnt + (size_t) cur * D + j) : 1.0f;
	      double fac = 0.0;
	      if(r >= hh)
		{
		  double a = STOCK ? m / r2 : (double) accel_over_r(A.laws.accel[ij], A.laws.par[ij], tp.m, q.w, (float) r2, (float) r, (float) (1.0 / r), nn) * r;
		  if(SR)
		    a -= m * (double) A.utor2wpi * (double) s_tab[(int) A.tabmap[ij] * A.ntab + tabindex];
		  fac = a / r;
		}
	      else
		fac = STOCK ? (double) law_plummer(q.w, (float) hh, (float) r) : (double) accel_spline(A.laws.spline[ij], A.laws.par[ij], tp.m, q.w, (float) hh, (float) r, nn);
	      ax += dx * fac;
	      ay += dy * fac;
	      az += dz * fac;
	      any = true;
	      if(SR && !isnode)
		ninter++;	// a particle inside the table counts (forcetree.c:2031)
	    }
	  if(isnode && (!SR || any))
	    ninter++;		// forcetree.c:1585 resp. 2031-2032
	  cur = isnode ? w.x : cur + 1u;
	}
#pragma unroll
      for(int o = 16; o > 0; o >>= 1)
	{
	  ax += __shfl_xor_sync(0xffffffffu, ax, o);
	  ay += __shfl_xor_sync(0xffffffffu, ay, o);
	  az += __shfl_xor_sync(0xffffffffu, az, o);
	  ninter += __shfl_xor_sync(0xffffffffu, ninter, o);
	}
      if(lane == 0)
	{
	  // replace the FP32 walk's result (and its share of the interaction total)
	  const float old_inter = A.cres ? A.cres[5 * (size_t) tloc + 3] : A.cost[idx];
	  float lattc = 0.0f;
	  if(PERIODIC && !SR && A.latt && A.lattcost)
	    lattc = A.lattcost[idx];
	  walk_store_result<SR, PERIODIC>(A, idx, tloc, tp.x, tp.y, tp.z, (float) ax, (float) ay, (float) az, (float) ninter);
	  const long long delta = (long long) ninter - (long long) (old_inter - lattc);
	  atomicAdd(&A.counters[0], (unsigned long long) delta);
	}
    }
  if(blockIdx.x == 0 && threadIdx.x == 0)
    {
      A.counters[5] = (unsigned long long) nredo;
      A.counters[7] = (unsigned long long) *A.redo_count;	// flagged (> nredo: the list was full, the rest keep their FP32 walk)
    }
}

template <int D, bool SR, bool PERIODIC, bool UNEQUAL, bool STOCK, typename ACC, bool STATS, bool EXACT, bool DEFER, int CRIT = 0>
static int launch_one(g2gpu_ctx *c, const WalkArgs &A, int grid, size_t smem)
{
  if(smem > 48 * 1024)
    G2_CUDA(cudaFuncSetAttribute(walk_kernel<D, SR, PERIODIC, UNEQUAL, STOCK, ACC, STATS, EXACT, DEFER, CRIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  if(c->walk_carveout >= 0)	// share of the 256 KB array kept as shared memory, in percent (the rest is L1); -1: the driver's choice
    G2_CUDA(cudaFuncSetAttribute(walk_kernel<D, SR, PERIODIC, UNEQUAL, STOCK, ACC, STATS, EXACT, DEFER, CRIT>, cudaFuncAttributePreferredSharedMemoryCarveout, c->walk_carveout));
  walk_kernel<D, SR, PERIODIC, UNEQUAL, STOCK, ACC, STATS, EXACT, DEFER, CRIT><<<grid, WALK_KT, smem, c->stream>>>(A);
  if(EXACT)
    {
      if(smem > 48 * 1024)
	G2_CUDA(cudaFuncSetAttribute(walk_redo_kernel<D, SR, PERIODIC, UNEQUAL, STOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      walk_redo_kernel<D, SR, PERIODIC, UNEQUAL, STOCK><<<c->nsm * 6, 32 * WALK_REDO_WARPS, smem, c->stream>>>(A);
      c->launches++;
    }
  return 0;
}

template <int D, bool SR, bool PERIODIC, bool UNEQUAL, bool STOCK>
static int launch_walk(g2gpu_ctx *c, const WalkArgs &A, int grid, size_t smem, int acc_double, int stats)
{
  // exact (default): FP64 accumulators + guard bands + redo pass; otherwise the plain FP32-decision kernel with FP32 accumulators.
  // -DG2_WALK_DEFER builds the deferred-term variant as well (option walk_defer / G2GPU_WALK_DEFER = 1 selects it for the stock pair laws).
  // MEASURED (B200, round 2, profiles/r2_walk_defer.txt): parity green, 4 % faster at 128^3, 4 % slower at 256^3, 57 % slower on the tree-only
  // Hernquist walk -- the ring is evaluated at 10 of 32 lanes and ~50 instructions per pass, which is no better than evaluating in place.
#ifdef G2_WALK_DEFER
  if(STOCK && c->walk_defer)
    {
      if(stats)
	return launch_one<D, SR, PERIODIC, UNEQUAL, STOCK, double, true, true, STOCK>(c, A, grid, smem);
      if(!acc_double || !A.exact)
	return launch_one<D, SR, PERIODIC, UNEQUAL, STOCK, float, false, false, STOCK>(c, A, grid, smem);
      return launch_one<D, SR, PERIODIC, UNEQUAL, STOCK, double, false, true, STOCK>(c, A, grid, smem);
    }
#endif
  if(stats)			// instrumented instantiation (visits, species terms, decisions)
    return launch_one<D, SR, PERIODIC, UNEQUAL, STOCK, double, true, true, false>(c, A, grid, smem);
#ifdef G2_WALK_ACC_MATRIX
  if(!acc_double && A.exact)
    return launch_one<D, SR, PERIODIC, UNEQUAL, STOCK, float, false, true, false>(c, A, grid, smem);
  if(acc_double && !A.exact)
    return launch_one<D, SR, PERIODIC, UNEQUAL, STOCK, double, false, false, false>(c, A, grid, smem);
#endif
  if(!acc_double || !A.exact)
    return launch_one<D, SR, PERIODIC, UNEQUAL, STOCK, float, false, false, false>(c, A, grid, smem);
  // The relative criterion has its own instantiation of the default kernel, without the run-time criterion switch per decision (LDCU + FSETP +
  // BRA).  MEASURED (B200, 256^3, profiles/r2_walk_microvariants.txt): 170.68 -> 167.38 ms.  -DG2_WALK_NO_RELCRIT builds the generic kernel only.
#ifndef G2_WALK_NO_RELCRIT
  if(!(A.theta2 > 0.0f))
    return launch_one<D, SR, PERIODIC, UNEQUAL, STOCK, double, false, true, false, 1>(c, A, grid, smem);
#endif
  return launch_one<D, SR, PERIODIC, UNEQUAL, STOCK, double, false, true, false>(c, A, grid, smem);
}

template <int D>
static int dispatch_walk(g2gpu_ctx *c, const WalkArgs &A, int grid, size_t smem, bool sr, bool periodic, bool unequal, bool stock, int accd, int stats)
{
#define G2_W(SRv, PERv, UNEv, STv) return launch_walk<D, SRv, PERv, UNEv, STv>(c, A, grid, smem, accd, stats)
  if(sr)
    {				// TreePM implies PERIODIC and equal softenings are not required; keep both UNEQUAL variants
      if(unequal) { if(stock) G2_W(true, true, true, true); else G2_W(true, true, true, false); }
      else        { if(stock) G2_W(true, true, false, true); else G2_W(true, true, false, false); }
    }
  else if(periodic)
    {
      if(unequal) { if(stock) G2_W(false, true, true, true); else G2_W(false, true, true, false); }
      else        { if(stock) G2_W(false, true, false, true); else G2_W(false, true, false, false); }
    }
  else
    {
      if(unequal) { if(stock) G2_W(false, false, true, true); else G2_W(false, false, true, false); }
      else        { if(stock) G2_W(false, false, false, true); else G2_W(false, false, false, false); }
    }
#undef G2_W
}
