// g2_walk.cu — stage 3: the tree walk of force_treeevaluate (forcetree.c:1244-1610) and
// force_treeevaluate_shortrange (forcetree.c:1623-2052), with the gravity_tree epilogue (gravtree.c:304-358).
//
// One warp takes 32 targets that are adjacent along the tree order (hence in space) and walks the depth-first
// cell array once for all of them: at every cell each lane takes ITS OWN decision (cull / accept / open) with the
// reference's criteria; the warp descends if any lane opens (ballot), otherwise jumps to the cell's sibling.  A
// lane that accepted or culled a cell sleeps until the warp's cursor has left that cell's subtree (cells are in
// depth-first order, so this is one integer compare).  Direct particle children of an opened cell sit in a
// contiguous group and are applied by the opening lanes only.  Per-lane interaction lists are therefore those of
// the reference walk; only the summation order differs.
//
// Node records are (2+D) x 16 B, fetched with 128-bit loads (all lanes read the same address: one L1 broadcast).
#include "g2_walk_common.cuh"

template <int D, bool SR, bool PERIODIC, bool UNEQUAL, bool STOCK, typename ACC, int G>
__global__ void __launch_bounds__(WALK_THREADS, (D >= WALK_WIDE_D ? WALK_MINBLOCKS_WIDE : WALK_MINBLOCKS)) walk_kernel(const WalkArgs A)
{
  // G = targets per cursor.  G == 32: the whole warp shares one cursor (every node record is one broadcast load).
  // G < 32: the warp's 32 consecutive targets form 32/G sub-groups with their own cursors; a sub-group's union of
  // interaction regions is tighter, so fewer lanes sleep through a visit, at the price of 32/G distinct record
  // addresses per load instruction.  All sub-groups execute the same instruction stream (loads, decision, vote,
  // particle loop), only on different nodes.  MEASURED (B200, round 1): G = 16/8/4 need 1.2-1.7x fewer iterations per
  // cursor but run 10-30 % SLOWER than G = 32 on all three bench workloads (the divergent record loads cost more than
  // the saved visits), and software prefetch of the next record did not help either -- the kernel is issue bound.
  // Only G = 32 (and G = 8 with -DG2_WALK_SUBGROUPS, for experiments) is instantiated.
  extern __shared__ float s_tab[];
  __shared__ unsigned int s_chunk[WALK_WARPS];
  if(SR)
    {
      for(int i = threadIdx.x; i < A.ntables * A.ntab; i += WALK_THREADS)
	s_tab[i] = A.srtable[i];
      __syncthreads();
    }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int R = 2 + D;
  const unsigned int gmask = (G == 32) ? 0xffffffffu : (((1u << (G & 31)) - 1u) << (lane & ~(G - 1)));
  unsigned int s_tab_addr = (unsigned int) __cvta_generic_to_shared(s_tab);
#ifndef G2_WALK_NO_TABREG
  asm volatile("" : "+r"(s_tab_addr));	// keep the table base in a register: re-deriving it costs 4 uniform instructions per pair term
#endif
  unsigned int t2g_packed = 0;	// TypeToGrav as 6 nibbles
#pragma unroll
  for(int t = 0; t < 6; t++)
    t2g_packed |= (unsigned int) A.t2g[t] << (4 * t);
  const int nchunks = (A.hi - A.lo + 31) >> 5;
  unsigned long long tot_inter = 0, tot_visits = 0, tot_terms = 0, tot_dec = 0;

  while(true)
    {
      // dynamic work distribution: one chunk of 32 consecutive targets per warp
      if(lane == 0)
	s_chunk[warp] = atomicAdd(A.work_counter, 1u);
      __syncwarp();
      const unsigned int chunk = s_chunk[warp];
      __syncwarp();
      if(chunk >= (unsigned int) nchunks)
	break;
      const int ti = A.lo + (int) chunk * 32 + lane;
      const bool valid = ti < A.hi;
      unsigned int idx = 0;
      float px = 0, py = 0, pz = 0, pmass = 0, aold = 0;
      int ptype = 1;
      if(valid)
	{
	  idx = A.tq[A.targets[ti]];
	  const G2PRec p = A.prec[idx];
	  px = p.x; py = p.y; pz = p.z; pmass = p.m;
	  ptype = p.type;
	  aold = A.errtol * p.oldacc;	// forcetree.c:1289
	}
      const int tg = (t2g_packed >> (4 * ptype)) & 7;
      const float hself = A.fsoft[ptype];
      ACC ax = 0, ay = 0, az = 0;
      float fx = 0.0f, fy = 0.0f, fz = 0.0f;
      int ninter = 0, nterms = 0, ndec = 0;
      unsigned int skip_until = valid ? 0u : 0xffffffffu;
      unsigned int iter = 0;	// per chunk, so that the FP32 flush points (and hence the result bits) do not depend on scheduling
      const unsigned int end = (unsigned int) A.numnodes;
      // a sub-group without any valid target has nothing to walk
      unsigned int cur = (__ballot_sync(0xffffffffu, valid) & gmask) ? 0u : end;
      // TreePM: a target farther than rcut + len/2 (+ margins) from every face of the box needs no periodic image of a cell of size len:
      // cells across a face are culled with the raw distance as well (raw >= nearest-image distance >= distance to the face).  Cells
      // smaller than nowrap_len (warp minimum over the 32 targets, so uniform) skip the image-shift arithmetic; results are unchanged.
      float nowrap_len = 0.0f;
#ifndef G2_WALK_NO_NOWRAP
      if(SR && PERIODIC && G == 32)
#else
      if(false)
#endif
	{
	  float m = valid ? fminf(fminf(fminf(px, A.boxsize - px), fminf(py, A.boxsize - py)), fminf(pz, A.boxsize - pz)) : 3.0e38f;
#pragma unroll
	  for(int o = 16; o > 0; o >>= 1)
	    m = fminf(m, __shfl_xor_sync(0xffffffffu, m, o));
	  nowrap_len = fminf(A.shift_len_max, 1.99f * (m - A.rcut - 2.0f * A.cull_margin));
	}

      while(true)
	{
	  const bool live = cur < end;
	  if(G == 32)
	    {
	      if(!live)
		break;
	    }
	  else if(!__any_sync(0xffffffffu, live))
	    break;
	  const float4 *rec = A.cells + (size_t) (live ? cur : 0u) * R;
	  const float4 q0 = __ldg(rec);
	  const uint4 w = __ldg((const uint4 *) (rec + 1 + D));
#ifdef G2_WALK_PREFETCH
	  // EXPERIMENT (off): the next record is either the first child (cur + 1, adjacent) or the sibling; asking L1 for the sibling's line
	  // now (prefetch.global.L1, no destination register) would hide its L2 latency behind this visit.  MEASURED (B200, round 1): 208.3 ->
	  // 215.3 ms at 256^3, 3.69 -> 4.16 ms Hernquist 1 M: the kernel is issue bound, the extra instructions cost more than the latency.
	  if(G == 32 && live && w.x < end)
	    asm volatile("prefetch.global.L1 [%0];" :: "l"(A.cells + (size_t) w.x * R));
#endif
	  bool open = false;
	  // TreePM: a cell that can interact with (or must be opened by) a target lies within rcut + len of it, so for
	  // len < L/2 - rcut every point of the cell has the same periodic image as the cell centre; points of cells that
	  // are culled anyway can only look farther away.  The image shift is then computed once per cell, not per point.
	  const bool small_cell = SR && PERIODIC && q0.x < A.shift_len_max;	// uniform within a sub-group (a property of the cell)
	  float shx = 0.0f, shy = 0.0f, shz = 0.0f;
	  if(live && cur >= skip_until)
	    {
	      ndec++;
	      float dx[D], dy[D], dz[D], r2[D], mass[D];
	      float r2min = 3.0e38f, r2max = -1.0f, summass = 0.0f;
	      const float len = q0.x;
	      const float cxr = q0.y - px, cyr = q0.z - py, czr = q0.w - pz;
	      bool done = false;	// culled: skip the subtree without interaction
	      bool outside = false;
	      if(SR && PERIODIC && !(len < nowrap_len))
		{
		  shx = A.boxsize * rint_small(cxr * A.boxinv);
		  shy = A.boxsize * rint_small(cyr * A.boxinv);
		  shz = A.boxsize * rint_small(czr * A.boxinv);
		}
	      if(SR)
		{
		  // forcetree.c:1828-1862 culls a node if r2min > rcut^2 AND the target is farther than rcut + len/2 from the
		  // node centre on some axis.  All mass of a node lies inside its cube (up to float rounding of positions), so a
		  // target that clears the cube by a small margin on one axis is farther than rcut from every centre of mass:
		  // r2min > rcut^2 is then certain and the per-species distances need not be computed at all.
		  const float eff = A.rcut + 0.5f * len;
		  const float d0 = fabsf(PERIODIC ? cxr - shx : cxr), d1 = fabsf(PERIODIC ? cyr - shy : cyr), d2 = fabsf(PERIODIC ? czr - shz : czr);
		  const float dmax = fmaxf(fmaxf(d0, d1), d2);
		  outside = dmax > eff;
		  done = dmax > eff + A.cull_margin + 1.0e-3f * len;
		}
	      if(!done)
		{
#pragma unroll
		  for(int g = 0; g < D; g++)
		    {
		      const float4 q = __ldg(rec + 1 + g);
		      mass[g] = q.w;
		      summass += q.w;
		      if(small_cell)
			{
			  dx[g] = (q.x - px) - shx;
			  dy[g] = (q.y - py) - shy;
			  dz[g] = (q.z - pz) - shz;
			}
		      else
			{
			  dx[g] = nearest<PERIODIC>(q.x - px, A.boxsize, A.boxinv);
			  dy[g] = nearest<PERIODIC>(q.y - py, A.boxsize, A.boxinv);
			  dz[g] = nearest<PERIODIC>(q.z - pz, A.boxsize, A.boxinv);
			}
		      r2[g] = dx[g] * dx[g] + dy[g] * dy[g] + dz[g] * dz[g];
		      r2min = fminf(r2min, r2[g]);
		      r2max = fmaxf(r2max, r2[g]);
		    }
		  if(SR && outside && r2min > A.rcut2)
		    done = true;
		}
	      if(!done)
		{
		  if(A.theta2 > 0.0f)
		    {		// Barnes-Hut, forcetree.c:1437-1445
		      if(len * len > r2min * A.theta2)
			open = true;
		    }
		  else
		    {		// relative criterion, forcetree.c:1446-1472
		      if(summass * len * len > r2min * r2min * aold)
			open = true;
		      else if(fabsf(cxr) < 0.60f * len && fabsf(cyr) < 0.60f * len && fabsf(czr) < 0.60f * len)
			open = true;
		    }
		}
	      float h = hself;
	      if(UNEQUAL && !done && !open)
		{		// forcetree.c:1475-1501; the record carries ForceSoftening[maxsofttype] (+inf and the
				// mixed-softening bit for an empty node, maxsofttype == 7, which is always opened)
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
		  skip_until = w.x;	// sleep until the cursor leaves this subtree
		  if(!done)
		    {
		      bool any = false;
#pragma unroll
		      for(int g = 0; g < D; g++)
			{
			  const float nn = (!STOCK && A.cnt) ? (float) __ldg(A.cnt + (size_t) cur * D + g) : 1.0f;
			  bool cnt = pair_term<SR, STOCK>(A, s_tab, s_tab_addr, tg, g, D, pmass, mass[g], dx[g], dy[g], dz[g], r2[g], h, fx, fy, fz, nn);
			  any |= cnt;
			  nterms += cnt;
			}
		      if(!SR || any)
			ninter++;	// forcetree.c:1585 resp. 2031-2032
		    }
		}
	    }
	  iter++;
	  const unsigned int ball = __ballot_sync(0xffffffffu, open);
	  const bool gopen = (ball & gmask) != 0u;	// some target of my sub-group opens the node
	  if(G != 32 && live && (lane & (G - 1)) == 0)
	    tot_visits++;
	  if(ball != 0u)
	    {
	      // FP32 partial sums go into the (FP64) accumulators whenever the warp descends (about every third visit): few
	      // conversions, bounded error, and flush points that depend on the traversal only (=> reproducible bits)
	      ax += (ACC) fx; ay += (ACC) fy; az += (ACC) fz;
	      fx = fy = fz = 0.0f;
	      // direct particle children of the opened cell, for the lanes that opened it
	      const unsigned int np = (live && gopen) ? (w.z & 15u) : 0u;
	      const unsigned int npmax = (G == 32) ? np : __reduce_max_sync(0xffffffffu, np);
	      for(unsigned int j = 0; j < npmax; j++)
		{
		  if(j < np)
		    {
		      const float4 p = __ldg(A.wpart + w.y + j);
		      if(open)
			{
			  const int stype = (w.z >> (4 + 3 * j)) & 7;
			  const int sg = (t2g_packed >> (4 * stype)) & 7;
			  float h = hself;
			  if(UNEQUAL)
			    h = fmaxf(h, A.fsoft[stype]);	// forcetree.c:1412-1415
			  float ddx, ddy, ddz;
			  if(small_cell)
			    {
			      ddx = (p.x - px) - shx;
			      ddy = (p.y - py) - shy;
			      ddz = (p.z - pz) - shz;
			    }
			  else
			    {
			      ddx = nearest<PERIODIC>(p.x - px, A.boxsize, A.boxinv);
			      ddy = nearest<PERIODIC>(p.y - py, A.boxsize, A.boxinv);
			      ddz = nearest<PERIODIC>(p.z - pz, A.boxsize, A.boxinv);
			    }
			  float rr2 = ddx * ddx + ddy * ddy + ddz * ddz;
			  bool counted = pair_term<SR, STOCK>(A, s_tab, s_tab_addr, tg, sg, D, pmass, p.w, ddx, ddy, ddz, rr2, h, fx, fy, fz);
			  nterms += counted;
			  if(!SR || counted)
			    ninter++;
			}
		    }
		}
	    }
	  if(live)
	    cur = gopen ? cur + 1u : w.x;
	}

      ax += (ACC) fx; ay += (ACC) fy; az += (ACC) fz;
      if(G == 32 && lane == 0)
	tot_visits += iter;	// one cursor per warp: visits = loop trips
      if(valid)
	{
	  // gravity_tree epilogue: GravAccel is stored as FLOAT (forcetree.c:1592-1594), then gravtree.c:304-358
	  fx = (float) ax; fy = (float) ay; fz = (float) az;
	  if(PERIODIC && !SR && A.latt)
	    {			// force_treeevaluate_lattice_correction adds to the FLOAT result and to GravCost (forcetree.c:2435-2438)
	      fx = (float) ((double) fx + (double) A.latt[3 * (size_t) idx + 0]);
	      fy = (float) ((double) fy + (double) A.latt[3 * (size_t) idx + 1]);
	      fz = (float) ((double) fz + (double) A.latt[3 * (size_t) idx + 2]);
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
	  A.oldacc_out[idx] = (float) sqrt(sx * sx + sy * sy + sz * sz);
	  fx = (float) ((double) fx * A.G);
	  fy = (float) ((double) fy * A.G);
	  fz = (float) ((double) fz * A.G);
	  if(A.pos_fac_post_g != 0.0)
	    {
	      fx = (float) ((double) fx + A.pos_fac_post_g * (double) px);
	      fy = (float) ((double) fy + A.pos_fac_post_g * (double) py);
	      fz = (float) ((double) fz + A.pos_fac_post_g * (double) pz);
	    }
	  A.acc[3 * (size_t) idx + 0] = fx;
	  A.acc[3 * (size_t) idx + 1] = fy;
	  A.acc[3 * (size_t) idx + 2] = fz;
	  A.cost[idx] = (PERIODIC && !SR && A.lattcost) ? (float) ninter + A.lattcost[idx] : (float) ninter;
	  tot_inter += (unsigned long long) ninter;
	  tot_terms += (unsigned long long) nterms;
	  tot_dec += (unsigned long long) ndec;
	}
    }
  // statistics: interactions (= sum of GravCost), cell visits (per cursor), species terms
#pragma unroll
  for(int o = 16; o > 0; o >>= 1)
    {
      tot_inter += __shfl_xor_sync(0xffffffffu, tot_inter, o);
      tot_terms += __shfl_xor_sync(0xffffffffu, tot_terms, o);
      tot_dec += __shfl_xor_sync(0xffffffffu, tot_dec, o);
      tot_visits += __shfl_xor_sync(0xffffffffu, tot_visits, o);
    }
  if(lane == 0)
    {
      atomicAdd(&A.counters[0], tot_inter);
      atomicAdd(&A.counters[1], tot_visits);
      atomicAdd(&A.counters[2], tot_terms);
      atomicAdd(&A.counters[4], tot_dec);
    }
}

// ---------------------------------------------------------------- active target list ---------------------------
__global__ void __launch_bounds__(256) target_flag_kernel(const unsigned int *__restrict__ tq, const G2PRec *__restrict__ prec, int n,
							  unsigned int *__restrict__ flags)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if(p < n)
    flags[p] = prec[tq[p]].active ? 1u : 0u;
}

__global__ void __launch_bounds__(256) target_compact_kernel(const unsigned int *__restrict__ scan, int n, unsigned int *__restrict__ targets)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if(p < n && scan[p + 1] != scan[p])	// exclusive scan of the 0/1 flags: a step marks an active target
    targets[scan[p]] = (unsigned int) p;
}

template <int D, bool SR, bool PERIODIC, bool UNEQUAL, bool STOCK, typename ACC, int G>
static int launch_one(g2gpu_ctx *c, const WalkArgs &A, int grid, size_t smem)
{
  if(smem > 48 * 1024)
    G2_CUDA(cudaFuncSetAttribute(walk_kernel<D, SR, PERIODIC, UNEQUAL, STOCK, ACC, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
  walk_kernel<D, SR, PERIODIC, UNEQUAL, STOCK, ACC, G><<<grid, WALK_THREADS, smem, c->stream>>>(A);
  return 0;
}

template <int D, bool SR, bool PERIODIC, bool UNEQUAL, bool STOCK>
static int launch_walk(g2gpu_ctx *c, const WalkArgs &A, int grid, size_t smem, int acc_double, int group)
{
  if(!acc_double)
    return launch_one<D, SR, PERIODIC, UNEQUAL, STOCK, float, 32>(c, A, grid, smem);
  switch (group)
    {
#ifdef G2_WALK_SUBGROUPS
    case 8: return launch_one<D, SR, PERIODIC, UNEQUAL, STOCK, double, 8>(c, A, grid, smem);
#endif
    default: return launch_one<D, SR, PERIODIC, UNEQUAL, STOCK, double, 32>(c, A, grid, smem);
    }
}

template <int D>
static int dispatch_walk(g2gpu_ctx *c, const WalkArgs &A, int grid, size_t smem, bool sr, bool periodic, bool unequal, bool stock, int accd, int group)
{
#define G2_W(SRv, PERv, UNEv, STv) return launch_walk<D, SRv, PERv, UNEv, STv>(c, A, grid, smem, accd, group)
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

int g2_stage_walk(g2gpu_ctx *c, const g2gpu_walk_params *wp)
{
  if(c->stage < 3)
    return g2_fail(G2GPU_ERR_STATE, "walk: tree has not been built");
  if(!c->laws_set)
    return g2_fail(G2GPU_ERR_LAW, "walk: pair force laws not set (g2gpu_set_laws)");
  const int n = c->npart, D = c->D;
  const bool sr = c->cfg.shortrange != 0;
  if(sr && !c->srtable_set)
    return g2_fail(G2GPU_ERR_STATE, "walk: short-range table not set (g2gpu_set_srtable)");
  cudaStream_t st = c->stream;
  if(c->walk_mode == 1 && !c->bfs_valid && !(c->accumulator && c->counts_valid) && c->walk_group == 32)
    G2_TRY(g2_stage_bfs(c));	// normally done by the build; here only if the option was set after it
  G2_CUDA(cudaEventRecord(c->ev[6], st));

  // active targets in tree order (gravtree.c:113: Ti_endstep == Ti_Current), split into nranks equal slices
  target_flag_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(c->tq, c->prec, n, c->w_flags);
  c->launches++;
  G2_TRY(g2_scan_exclusive_u32(c, c->w_flags, c->w_flags, (size_t) n));
  target_compact_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(c->w_flags, n, c->w_targets);
  c->launches++;
  G2_CUDA(cudaMemcpyAsync(&c->h_err[4], c->w_flags + n, sizeof(int), cudaMemcpyDeviceToHost, st));
  G2_CUDA(cudaMemsetAsync(c->d_counters, 0, 8 * sizeof(unsigned long long), st));
  G2_CUDA(cudaStreamSynchronize(st));
  c->w_ntargets = c->h_err[4];
  const int nr = c->cfg.nranks > 0 ? c->cfg.nranks : 1, rk = c->cfg.rank;
  c->w_lo = (int) ((long long) c->w_ntargets * rk / nr);
  c->w_hi = (int) ((long long) c->w_ntargets * (rk + 1) / nr);

  WalkArgs A;
  memset(&A, 0, sizeof(A));
  A.cells = c->wcells; A.wpart = c->wpart; A.targets = c->w_targets; A.tq = c->tq; A.prec = c->prec;
  A.cnt = (c->accumulator && c->counts_valid) ? c->wcnt : nullptr;
  A.gravpm = c->gravpm; A.srtable = c->d_srtable_f; A.acc = c->acc; A.cost = c->cost;
  A.oldacc_out = c->oldacc_out; A.counters = c->d_counters; A.work_counter = (unsigned int *) (c->d_counters + 3);
  A.lo = c->w_lo; A.hi = c->w_hi; A.numnodes = c->numnodes; A.ntab = c->cfg.ntab;
  A.theta2 = (float) (wp->theta * wp->theta);
  A.errtol = (float) wp->errtol_force_acc;
  A.boxsize = (float) wp->boxsize; A.boxinv = wp->boxsize > 0 ? (float) (1.0 / wp->boxsize) : 0.0f;
  if(sr)
    {
      A.rcut = (float) wp->rcut; A.rcut2 = (float) (wp->rcut * wp->rcut);
      A.asmthfac = (float) (0.5 / wp->asmth * (c->cfg.ntab / 3.0));	// forcetree.c:1708
      A.utor2wpi = (float) (1.0 / (M_PI * 4 * wp->asmth * wp->asmth));	// forcetree.c:1711
      A.shift_len_max = (float) (0.499 * wp->boxsize - 1.001 * wp->rcut);
      A.cull_margin = (float) (1.0e-5 * wp->boxsize);
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

  size_t smem = sr ? sizeof(float) * (size_t) A.ntables * A.ntab : 0;
  const int ntgt = c->w_hi - c->w_lo;
  int grid = c->nsm * (D >= WALK_WIDE_D ? WALK_MINBLOCKS_WIDE : WALK_MINBLOCKS);	// every CTA resident: the chunk counter balances the load
  int need = g2_cdiv(g2_cdiv(ntgt, 32), WALK_WARPS);
  if(grid > need)
    grid = need;
  // walk mode 1 (one warp per target) needs the level-order records; pair laws that take the per-cell particle counts
  // (NGRAVS_ACCUMULATOR) and the sub-group experiments stay on the cursor walk
  // periodic box without PM: the lattice-sum correction walk first (its result enters the epilogue of the walk kernel)
  const bool lattice = c->cfg.periodic && !sr && c->lattice_set && ntgt > 0;
  if(lattice)
    {
      if(!(wp->boxsize > 0))
	return g2_fail(G2GPU_ERR_ARG, "walk: boxsize must be positive for the lattice-sum correction");
      G2_TRY(g2_stage_lattice(c, wp));
      A.latt = c->latt; A.lattcost = c->lattcost;
    }
  const bool mode_b = c->walk_mode == 1 && c->bfs_valid && A.cnt == nullptr && c->walk_group == 32 && !lattice;
  G2_CUDA(cudaEventRecord(c->ev[7], st));
  if(ntgt > 0 && mode_b)
    {
      A.bq0 = c->b_q0; A.bs = c->b_s; A.bw = c->b_w; A.bptype = c->b_ptype; A.bstride = (unsigned int) c->numnodes;
      G2_TRY(g2_launch_walkb(c, A, sr, c->cfg.periodic != 0, c->cfg.unequal_softenings != 0, stock));
      c->launches++;
    }
  else if(ntgt > 0)
    {
      const bool per = c->cfg.periodic != 0, uneq = c->cfg.unequal_softenings != 0;
      int rc;
      switch (D)
	{
#ifndef G2_FAST_BUILD
	case 1: rc = dispatch_walk<1>(c, A, grid, smem, sr, per, uneq, stock, c->acc_double, c->walk_group); break;
#endif
	case 2: rc = dispatch_walk<2>(c, A, grid, smem, sr, per, uneq, stock, c->acc_double, c->walk_group); break;
#ifndef G2_FAST_BUILD
	case 3: rc = dispatch_walk<3>(c, A, grid, smem, sr, per, uneq, stock, c->acc_double, c->walk_group); break;
#endif
	case 4: rc = dispatch_walk<4>(c, A, grid, smem, sr, per, uneq, stock, c->acc_double, c->walk_group); break;
#ifndef G2_FAST_BUILD
	case 5: rc = dispatch_walk<5>(c, A, grid, smem, sr, per, uneq, stock, c->acc_double, c->walk_group); break;
	case 6: rc = dispatch_walk<6>(c, A, grid, smem, sr, per, uneq, stock, c->acc_double, c->walk_group); break;
#endif
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
