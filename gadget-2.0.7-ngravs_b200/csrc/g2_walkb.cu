// g2_walkb.cu — the tree walk of force_treeevaluate (forcetree.c:1244-1610) and force_treeevaluate_shortrange
// (forcetree.c:1623-2052) with ONE WARP PER TARGET: the 32 lanes work on 32 different cells (or 32 source particles)
// of the same target, so no lane waits for the decisions of a neighbouring target.
//
// The warp keeps three small LIFOs in shared memory:
//   * cells to examine, as ranges "children of an opened cell" (first V | count << 28) -- the level-order record
//     layout of g2_stage_bfs makes the child cells of a cell contiguous, so one word describes up to 8 of them;
//   * source particles to apply, as ranges "direct particles of an opened cell" (offset in wpart | count << 28);
//   * accepted cells (V), whose D species monopoles are applied in dense batches of 32.
// A step pops 32 cells (ranges are split across lanes by a warp scan), each lane takes the reference's decision for
// its cell -- cull (forcetree.c:1828-1862), open (Barnes-Hut 1865-1873 / relative criterion 1874-1897 / softening rule
// 1899-1926) or accept -- and the three outcomes are appended to the LIFOs with ballots.  Interaction lists are those
// of the reference walk for every target (GravCost identical); only the order of the summation differs, and it does
// not depend on scheduling (every target is walked by one warp in a fixed order => reproducible bits).
//
// Stack bound: a "wide" step pops >= 1 range and pushes <= 32; it is only taken while sp <= WB_STACK_WIDE.  Otherwise
// the warp pops ONE range (<= 8 cells, <= 8 pushes), i.e. a depth-first descent that can add at most 7 entries per
// tree level: sp <= WB_STACK_WIDE + 32 + 7 * G2_MAXDEPTH < WB_STACK always.
#include "g2_walk_common.cuh"

#define WB_THREADS 256
#define WB_WARPS (WB_THREADS / 32)
#ifndef WB_MINBLOCKS
#define WB_MINBLOCKS 4
#endif
#define WB_STACK 384
#define WB_STACK_WIDE (WB_STACK - 7 * G2_MAXDEPTH - 8 - 32)
#define WB_PQ 64
#define WB_AQ 64
#define WB_WARP_WORDS (WB_STACK + WB_PQ + WB_AQ)
#define WB_FULL 0xffffffffu
#ifndef WB_CHUNK
#define WB_CHUNK 16		// consecutive targets a warp takes at a time
#endif

// Pops up to 32 items from a LIFO of ranges (first | count << 28, count >= 1).  Lane l receives item l (returns whether
// it exists).  Ranges are consumed from the top; the last one touched may be split (its remainder stays on the stack).
__device__ __forceinline__ bool pop_items(unsigned int *__restrict__ stk, int &sp, int lane, bool narrow, unsigned int &item)
{
  const int e = sp - 1 - lane;
  const bool have = e >= 0 && (!narrow || lane == 0);
  const unsigned int ent = have ? stk[e] : 0u;
  const unsigned int c = ent >> 28;
  // exclusive prefix sum of the 4-bit counts from four independent ballots (shorter dependency chain than a shuffle scan)
  const unsigned int lt = (1u << lane) - 1u;
  const unsigned int b0 = __ballot_sync(WB_FULL, c & 1u), b1 = __ballot_sync(WB_FULL, c & 2u), b2 = __ballot_sync(WB_FULL, c & 4u),
    b3 = __ballot_sync(WB_FULL, c & 8u);
  const unsigned int excl = __popc(b0 & lt) + 2u * __popc(b1 & lt) + 4u * __popc(b2 & lt) + 8u * __popc(b3 & lt);
  const unsigned int incl = excl + c;
  const unsigned int total = min((unsigned int) (__popc(b0) + 2 * __popc(b1) + 4 * __popc(b2) + 8 * __popc(b3)), 32u);
  const unsigned int starts = __reduce_or_sync(WB_FULL, (c > 0u && excl < 32u) ? (1u << excl) : 0u);
  const int j = (__popc(starts & (WB_FULL >> (31 - lane))) - 1) & 31;	// the range item `lane` belongs to
  const unsigned int first_j = __shfl_sync(WB_FULL, ent & 0x0fffffffu, j);
  const unsigned int excl_j = __shfl_sync(WB_FULL, excl, j);
  item = first_j + ((unsigned int) lane - excl_j);
  const int nfull = __popc(__ballot_sync(WB_FULL, have && incl <= 32u));
  if(have && excl < 32u && incl > 32u)	// split: keep the remainder
    stk[e] = ((ent & 0x0fffffffu) + (32u - excl)) | ((incl - 32u) << 28);
  sp -= nfull;
  __syncwarp();
  return (unsigned int) lane < total;
}

template <int D, bool SR, bool PERIODIC, bool UNEQUAL, bool STOCK, typename ACC>
__global__ void __launch_bounds__(WB_THREADS, WB_MINBLOCKS) walkb_kernel(const WalkArgs A)
{
  extern __shared__ float s_tab[];
  __shared__ unsigned int s_chunk[WB_WARPS];
  const int ntabf = SR ? A.ntables * A.ntab : 0;
  if(SR)
    {
      for(int i = threadIdx.x; i < ntabf; i += WB_THREADS)
	s_tab[i] = A.srtable[i];
      __syncthreads();
    }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned int *const s_stack = (unsigned int *) (s_tab + ntabf) + warp * WB_WARP_WORDS;
  unsigned int *const s_pq = s_stack + WB_STACK;
  unsigned int *const s_aq = s_pq + WB_PQ;
  const unsigned int s_tab_addr = (unsigned int) __cvta_generic_to_shared(s_tab);
  const unsigned int lt_mask = (1u << lane) - 1u;
  unsigned int t2g_packed = 0;
#pragma unroll
  for(int t = 0; t < 6; t++)
    t2g_packed |= (unsigned int) A.t2g[t] << (4 * t);
  const int nchunks = (A.hi - A.lo + WB_CHUNK - 1) / WB_CHUNK;

  while(true)
    {
      if(lane == 0)
	s_chunk[warp] = atomicAdd(A.work_counter, 1u);
      __syncwarp();
      const unsigned int chunk = s_chunk[warp];
      __syncwarp();
      if(chunk >= (unsigned int) nchunks)
	break;
      const int t0 = A.lo + (int) chunk * WB_CHUNK, t1 = min(t0 + WB_CHUNK, A.hi);
      unsigned int c_inter = 0, c_steps = 0, c_terms = 0, c_dec = 0;

      for(int ti = t0; ti < t1; ti++)
	{
	  // the target (same addresses in all lanes: broadcast loads)
	  const unsigned int idx = __ldg(A.tq + __ldg(A.targets + ti));
	  const float4 pp = __ldg((const float4 *) (A.prec + idx));
	  const int ptype = __ldg(&A.prec[idx].type);
	  const float px = pp.x, py = pp.y, pz = pp.z, pmass = pp.w;
	  const float aold = A.errtol * __ldg(&A.prec[idx].oldacc);	// forcetree.c:1289
	  const int tg = (t2g_packed >> (4 * ptype)) & 7;
	  const float hself = A.fsoft[ptype];
	  float fx = 0.0f, fy = 0.0f, fz = 0.0f;
	  int ninter = 0, nterms = 0, ndec = 0, nsteps = 0;
	  int sp = 1, npq = 0, naq = 0, pq_total = 0;
	  if(lane == 0)
	    s_stack[0] = 0u | (1u << 28);	// the root is examined like every other cell (forcetree.c:1343)
	  __syncwarp();

	  // Software pipeline of one iteration: pop the next 32 cells and ISSUE their record loads, then evaluate the batches of
	  // accepted cells / source particles queued so far (their loads overlap the cell loads), then take the decisions.
	  while(true)
	    {
	      const bool drain = sp == 0;	// no cells left: empty the two queues and finish
	      // ---- 1. next 32 cells, loads in flight ----
	      unsigned int V = 0;
	      bool take = false;
	      float4 q0 = make_float4(0.f, 0.f, 0.f, 0.f), qs[D];
	      uint4 w = make_uint4(0u, 0u, 0u, 0u);
	      if(!drain)
		{
		  take = pop_items(s_stack, sp, lane, sp > WB_STACK_WIDE, V);
		  nsteps++;
		  if(take)
		    {
		      q0 = __ldg(A.bq0 + V);
		      w = __ldg(A.bw + V);
#pragma unroll
		      for(int g = 0; g < D; g++)
			qs[g] = __ldg(A.bs + (size_t) g * A.bstride + V);
		    }
		}
	      // ---- 2. queued interactions: accepted cells (D species terms each) and source particles, 32 at a time ----
	      const int lim = drain ? 1 : 32;
	      while(naq >= lim || pq_total >= lim)
		{
		  const bool do_aq = naq >= lim, do_pq = pq_total >= lim;
		  unsigned int pi = 0, C = 0;
		  bool hp = false, ha = false;
		  if(do_pq)
		    {
		      hp = pop_items(s_pq, npq, lane, false, pi);
		      pq_total = max(pq_total - 32, 0);
		    }
		  if(do_aq)
		    {
		      const int e = naq - 1 - lane;
		      naq = max(naq - 32, 0);
		      ha = e >= 0;
		      if(ha)
			C = s_aq[e];
		    }
		  float4 p = make_float4(0.f, 0.f, 0.f, 0.f), qa[D];
		  int stype = 1;
		  float ha_h = hself;
		  if(hp)
		    {
		      p = __ldg(A.wpart + pi);
		      stype = (int) __ldg(A.bptype + pi);
		    }
		  if(ha)
		    {
#pragma unroll
		      for(int g = 0; g < D; g++)
			qa[g] = __ldg(A.bs + (size_t) g * A.bstride + C);
		      if(UNEQUAL)
			ha_h = fmaxf(hself, __uint_as_float(__ldg(&A.bw[C].w)));
		    }
		  if(hp)
		    {
		      const int sg = (t2g_packed >> (4 * stype)) & 7;
		      float h = hself;
		      if(UNEQUAL)
			h = fmaxf(h, A.fsoft[stype]);	// forcetree.c:1412-1415
		      const float dx = nearest<PERIODIC>(p.x - px, A.boxsize, A.boxinv);
		      const float dy = nearest<PERIODIC>(p.y - py, A.boxsize, A.boxinv);
		      const float dz = nearest<PERIODIC>(p.z - pz, A.boxsize, A.boxinv);
		      const float r2 = dx * dx + dy * dy + dz * dz;
		      const bool counted = pair_term<SR, STOCK>(A, s_tab, s_tab_addr, tg, sg, D, pmass, p.w, dx, dy, dz, r2, h, fx, fy, fz);
		      nterms += counted;
		      if(!SR || counted)
			ninter++;
		    }
		  if(ha)
		    {
		      bool any = false;
#pragma unroll
		      for(int g = 0; g < D; g++)
			{
			  const float dx = nearest<PERIODIC>(qa[g].x - px, A.boxsize, A.boxinv);
			  const float dy = nearest<PERIODIC>(qa[g].y - py, A.boxsize, A.boxinv);
			  const float dz = nearest<PERIODIC>(qa[g].z - pz, A.boxsize, A.boxinv);
			  const float r2 = dx * dx + dy * dy + dz * dz;
			  const bool cnt = pair_term<SR, STOCK>(A, s_tab, s_tab_addr, tg, g, D, pmass, qa[g].w, dx, dy, dz, r2, ha_h, fx, fy, fz);
			  any |= cnt;
			  nterms += cnt;
			}
		      if(!SR || any)
			ninter++;	// forcetree.c:1585 resp. 2031-2032
		    }
		  __syncwarp();
		}
	      if(drain)
		break;
	      // ---- 3. the decisions ----
	      bool open = false, accept = false;
	      if(take)
		{
		  ndec++;
		  float r2min = 3.0e38f, r2max = -1.0f, summass = 0.0f;
		  const float len = q0.x;
		  const float cxr = q0.y - px, cyr = q0.z - py, czr = q0.w - pz;
		  bool done = false, outside = false;
		  float shx = 0.0f, shy = 0.0f, shz = 0.0f;
		  const bool small_cell = SR && PERIODIC && len < A.shift_len_max;
		  if(SR && PERIODIC)
		    {
		      shx = A.boxsize * rint_small(cxr * A.boxinv);
		      shy = A.boxsize * rint_small(cyr * A.boxinv);
		      shz = A.boxsize * rint_small(czr * A.boxinv);
		    }
		  if(SR)
		    {		// geometric shortcut of the cull test, see g2_walk.cu
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
			  const float4 q = qs[g];
			  summass += q.w;
			  float dx, dy, dz;
			  if(small_cell)
			    {
			      dx = (q.x - px) - shx;
			      dy = (q.y - py) - shy;
			      dz = (q.z - pz) - shz;
			    }
			  else
			    {
			      dx = nearest<PERIODIC>(q.x - px, A.boxsize, A.boxinv);
			      dy = nearest<PERIODIC>(q.y - py, A.boxsize, A.boxinv);
			      dz = nearest<PERIODIC>(q.z - pz, A.boxsize, A.boxinv);
			    }
			  const float r2 = dx * dx + dy * dy + dz * dz;
			  r2min = fminf(r2min, r2);
			  r2max = fmaxf(r2max, r2);
			}
		      if(SR && outside && r2min > A.rcut2)
			done = true;
		    }
		  if(!done)
		    {
		      if(A.theta2 > 0.0f)
			{	// Barnes-Hut, forcetree.c:1437-1445
			  if(len * len > r2min * A.theta2)
			    open = true;
			}
		      else
			{	// relative criterion, forcetree.c:1446-1472
			  if(summass * len * len > r2min * r2min * aold)
			    open = true;
			  else if(fabsf(cxr) < 0.60f * len && fabsf(cyr) < 0.60f * len && fabsf(czr) < 0.60f * len)
			    open = true;
			}
		      if(UNEQUAL && !open)
			{	// forcetree.c:1475-1501
			  const float hnode = __uint_as_float(w.w);
			  if(hself < hnode && r2max < hnode * hnode && ((w.z >> 28) & 1))
			    open = true;
			}
		      accept = !open;
		    }
		}
	      // ---- 4. append the outcomes ----
	      {
		const bool pc = open && (w.x >> 28) != 0u;	// child cells
		const unsigned int np = open ? (w.y >> 28) : 0u;	// direct particles
		const unsigned int bc = __ballot_sync(WB_FULL, pc);
		const unsigned int bp = __ballot_sync(WB_FULL, np != 0u);
		const unsigned int ba = __ballot_sync(WB_FULL, accept);
		if(pc)
		  s_stack[sp + __popc(bc & lt_mask)] = w.x;
		sp += __popc(bc);
		if(np != 0u)
		  s_pq[npq + __popc(bp & lt_mask)] = w.y;
		npq += __popc(bp);
		if(bp)
		  pq_total += (int) __reduce_add_sync(WB_FULL, np);
		if(accept)
		  s_aq[naq + __popc(ba & lt_mask)] = V;
		naq += __popc(ba);
		__syncwarp();
	      }
	    }


	  // ---- the target's sums ----
	  ACC sx = (ACC) fx, sy = (ACC) fy, sz = (ACC) fz;
#pragma unroll
	  for(int o = 16; o > 0; o >>= 1)
	    {
	      sx += __shfl_xor_sync(WB_FULL, sx, o);
	      sy += __shfl_xor_sync(WB_FULL, sy, o);
	      sz += __shfl_xor_sync(WB_FULL, sz, o);
	    }
	  ninter = (int) __reduce_add_sync(WB_FULL, (unsigned int) ninter);
	  c_inter += (unsigned int) ninter;
	  c_terms += (unsigned int) nterms;
	  c_dec += (unsigned int) ndec;
	  c_steps += (unsigned int) nsteps;
	  if(lane == 0)
	    {
	      // gravity_tree epilogue: GravAccel is stored as FLOAT (forcetree.c:1592-1594), then gravtree.c:304-358
	      fx = (float) sx; fy = (float) sy; fz = (float) sz;
	      if(A.pos_fac_pre_g != 0.0)
		{
		  fx = (float) ((double) fx + A.pos_fac_pre_g * (double) px);
		  fy = (float) ((double) fy + A.pos_fac_pre_g * (double) py);
		  fz = (float) ((double) fz + A.pos_fac_pre_g * (double) pz);
		}
	      double dsx = (double) fx, dsy = (double) fy, dsz = (double) fz;
	      if(A.use_gravpm)
		{
		  dsx += (double) A.gravpm[3 * (size_t) idx + 0] / A.G;
		  dsy += (double) A.gravpm[3 * (size_t) idx + 1] / A.G;
		  dsz += (double) A.gravpm[3 * (size_t) idx + 2] / A.G;
		}
	      A.oldacc_out[idx] = (float) sqrt(dsx * dsx + dsy * dsy + dsz * dsz);
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
	      A.cost[idx] = (float) ninter;
	    }
	}
      // statistics: interactions (= sum of GravCost), warp steps, species terms, decisions
      c_terms = __reduce_add_sync(WB_FULL, c_terms);
      c_dec = __reduce_add_sync(WB_FULL, c_dec);
      if(lane == 0)
	{
	  atomicAdd(&A.counters[0], (unsigned long long) c_inter);
	  atomicAdd(&A.counters[1], (unsigned long long) c_steps);
	  atomicAdd(&A.counters[2], (unsigned long long) c_terms);
	  atomicAdd(&A.counters[4], (unsigned long long) c_dec);
	}
    }
}

template <int D, bool SR, bool PERIODIC, bool UNEQUAL, bool STOCK>
static int launch_b(g2gpu_ctx *c, const WalkArgs &A, int acc_double)
{
  const size_t smem = (SR ? sizeof(float) * (size_t) A.ntables * A.ntab : 0) + sizeof(unsigned int) * WB_WARPS * WB_WARP_WORDS;
  const int ntgt = A.hi - A.lo;
  int grid = c->nsm * WB_MINBLOCKS;
  const int need = g2_cdiv(g2_cdiv(ntgt, WB_CHUNK), WB_WARPS);
  if(grid > need)
    grid = need;
  if(acc_double)
    {
      if(smem > 48 * 1024)
	G2_CUDA(cudaFuncSetAttribute(walkb_kernel<D, SR, PERIODIC, UNEQUAL, STOCK, double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      walkb_kernel<D, SR, PERIODIC, UNEQUAL, STOCK, double><<<grid, WB_THREADS, smem, c->stream>>>(A);
    }
  else
    {
      if(smem > 48 * 1024)
	G2_CUDA(cudaFuncSetAttribute(walkb_kernel<D, SR, PERIODIC, UNEQUAL, STOCK, float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem));
      walkb_kernel<D, SR, PERIODIC, UNEQUAL, STOCK, float><<<grid, WB_THREADS, smem, c->stream>>>(A);
    }
  return 0;
}

template <int D>
static int dispatch_b(g2gpu_ctx *c, const WalkArgs &A, bool sr, bool periodic, bool unequal, bool stock, int accd)
{
#define G2_W(SRv, PERv, UNEv, STv) return launch_b<D, SRv, PERv, UNEv, STv>(c, A, accd)
  if(sr)
    {
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

int g2_launch_walkb(g2gpu_ctx *c, const WalkArgs &A, bool sr, bool periodic, bool unequal, bool stock)
{
  switch (c->D)
    {
#ifndef G2_FAST_BUILD
    case 1: return dispatch_b<1>(c, A, sr, periodic, unequal, stock, c->acc_double);
#endif
    case 2: return dispatch_b<2>(c, A, sr, periodic, unequal, stock, c->acc_double);
#ifndef G2_FAST_BUILD
    case 3: return dispatch_b<3>(c, A, sr, periodic, unequal, stock, c->acc_double);
#endif
    case 4: return dispatch_b<4>(c, A, sr, periodic, unequal, stock, c->acc_double);
#ifndef G2_FAST_BUILD
    case 5: return dispatch_b<5>(c, A, sr, periodic, unequal, stock, c->acc_double);
    case 6: return dispatch_b<6>(c, A, sr, periodic, unequal, stock, c->acc_double);
#endif
    default: return g2_fail(G2GPU_ERR_ARG, "unsupported N_GRAVS %d", c->D);
    }
}
