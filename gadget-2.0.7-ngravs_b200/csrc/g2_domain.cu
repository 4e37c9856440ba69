// g2_domain.cu — stage 1 of the hot path on the device:
//   domain_findExtent            (domain.c:882-924)   bounding cube -> DomainCorner/Center/Len/Fac
//   key loop                     (domain.c:938-944)   Peano-Hilbert keys, peano_hilbert_key (peano.c:356-398)
//   peano_hilbert_order          (peano.c:36-185)     gas block, then species-major, PH within a block
//   domain_topsplit_local/topsplit/walktoptree (domain.c:1019-1138, 802-816)   TopNodes[]
//   force_create_empty_nodes     (forcetree.c:292-336) the force-tree nodes of the top-level tree
// Integer / FP64-exact work; results are bit-identical to the reference for the same float32 inputs.
#include "g2_common.cuh"
#include "../../include/g2_ph_table.h"

__constant__ unsigned char c_ph_table[G2_PH_NSTATES * 8] = G2_PH_TABLE_INIT;

// ---------------------------------------------------------------- extent ---------------------------------
__device__ __forceinline__ unsigned int f2ord(float f)
{
  unsigned int u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned int u)
{
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void extent_init_kernel(unsigned int *mm)
{
  if(threadIdx.x < 3)
    mm[threadIdx.x] = 0xffffffffu;	// min
  else if(threadIdx.x < 6)
    mm[threadIdx.x] = 0u;		// max
}

__global__ void __launch_bounds__(256) extent_kernel(const G2PRec *__restrict__ rec, int n, unsigned int *mm)
{
  float mn[3] = { 3.0e38f, 3.0e38f, 3.0e38f }, mx[3] = { -3.0e38f, -3.0e38f, -3.0e38f };
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
      float4 p = *((const float4 *) &rec[i]);
      mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
      mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
      mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
#pragma unroll
  for(int k = 0; k < 3; k++)
    {
#pragma unroll
      for(int o = 16; o > 0; o >>= 1)
	{
	  mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
	  mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
	}
    }
  if((threadIdx.x & 31) == 0)
    {
#pragma unroll
      for(int k = 0; k < 3; k++)
	{
	  atomicMin(&mm[k], f2ord(mn[k]));
	  atomicMax(&mm[3 + k], f2ord(mx[k]));
	}
    }
}

// domain.c:909-923 in the reference's double arithmetic (explicit round-to-nearest ops: no contraction)
__global__ void domain_params_kernel(const unsigned int *mm, double *dom)
{
  if(threadIdx.x != 0 || blockIdx.x != 0)
    return;
  double xmin[3], xmax[3], len = 0;
  for(int j = 0; j < 3; j++)
    {
      xmin[j] = (double) ord2f(mm[j]);
      xmax[j] = (double) ord2f(mm[3 + j]);
      double e = __dsub_rn(xmax[j], xmin[j]);
      if(e > len)
	len = e;
    }
  len = __dmul_rn(len, 1.001);
  for(int j = 0; j < 3; j++)
    {
      double mid = __dmul_rn(0.5, __dadd_rn(xmin[j], xmax[j]));
      dom[3 + j] = mid;						// DomainCenter
      dom[j] = __dsub_rn(mid, __dmul_rn(0.5, len));		// DomainCorner
    }
  dom[6] = len;							// DomainLen
  dom[7] = __dmul_rn(__ddiv_rn(1.0, len), (double) (1LL << G2_PH_BITS));	// DomainFac = 1.0/len * 2^18
}

// ---------------------------------------------------------------- keys -----------------------------------
// tab: the 24x8 state machine; threads of a warp index it divergently, so callers pass a SHARED-memory copy (the
// constant cache would serialise the 32 different addresses of every level)
__device__ __forceinline__ long long ph_key(const unsigned char *__restrict__ tab, int x, int y, int z, int bits)
{
  unsigned int st = 0;
  long long key = 0;
  for(int l = bits - 1; l >= 0; l--)
    {
      unsigned int o = (((x >> l) & 1) << 2) | (((y >> l) & 1) << 1) | ((z >> l) & 1);
      unsigned int e = tab[st * 8 + o];
      key = (key << 3) | (e & 7);
      st = e >> 3;
    }
  return key;
}

__device__ __forceinline__ void ph_table_to_shared(unsigned char *s_tab)
{
  for(int i = threadIdx.x; i < G2_PH_NSTATES * 8; i += blockDim.x)
    s_tab[i] = c_ph_table[i];
  __syncthreads();
}

struct G2TypeMap { int t2g[6]; };

// sort key = (block << 54) | PH key; block 0 = gas (type 0, peano.c:47-67), block 1+g = species g (peano.c:90-133)
__global__ void __launch_bounds__(256) keys_kernel(const G2PRec *__restrict__ rec, int n,
						   const double *__restrict__ dom, G2TypeMap tm,
						   unsigned long long *__restrict__ skey, unsigned int *__restrict__ sval, int *__restrict__ err)
{
  __shared__ unsigned char s_ph[G2_PH_NSTATES * 8];
  ph_table_to_shared(s_ph);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n)
    return;
  const G2PRec p = rec[i];
  double fac = dom[7];
  // (P[i].Pos[k] - DomainCorner[k]) * DomainFac, converted to int by truncation (domain.c:940-943)
  int x = __double2int_rz(__dmul_rn(__dsub_rn((double) p.x, dom[0]), fac));
  int y = __double2int_rz(__dmul_rn(__dsub_rn((double) p.y, dom[1]), fac));
  int z = __double2int_rz(__dmul_rn(__dsub_rn((double) p.z, dom[2]), fac));
  unsigned long long key = (unsigned long long) ph_key(s_ph, x, y, z, G2_PH_BITS);
  int t = p.type;
  if(t < 0 || t > 5)
    {				// particle types are 0..5 (allvars.h:571)
      atomicExch(&err[2], G2GPU_ERR_ARG);
      t = t < 0 ? 0 : 5;
    }
  unsigned long long block = (t == 0) ? 0ull : (unsigned long long) (1 + tm.t2g[t]);
  skey[i] = (block << (3 * G2_PH_BITS)) | key;
  sval[i] = (unsigned int) i;
}

__global__ void __launch_bounds__(256) gather_kernel(int n, const unsigned long long *__restrict__ skey, const unsigned int *__restrict__ sval,
						     const G2PRec *__restrict__ in_rec, const float *__restrict__ in_vel, const float *__restrict__ in_gravpm,
						     G2PRec *__restrict__ prec, float *__restrict__ vel, float *__restrict__ gravpm,
						     long long *__restrict__ phkey, int *__restrict__ perm)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n)
    return;
  unsigned int src = sval[i];
  G2PRec r = in_rec[src];
  r.type = r.type < 0 ? 0 : (r.type > 5 ? 5 : r.type);
  r.active = r.active != 0;
  prec[i] = r;
  if(in_vel)
    for(int k = 0; k < 3; k++)
      vel[3 * (size_t) i + k] = in_vel[3 * (size_t) src + k];
  if(in_gravpm)
    for(int k = 0; k < 3; k++)
      gravpm[3 * (size_t) i + k] = in_gravpm[3 * (size_t) src + k];
  phkey[i] = (long long) (skey[i] & ((1ull << (3 * G2_PH_BITS)) - 1ull));
  perm[i] = (int) src;
}

// block boundaries in the sorted (block, key) list: bstart[b] = first index with block >= b, b = 0..nblocks
__global__ void block_starts_kernel(const unsigned long long *__restrict__ skey, int n, int nblocks, int *bstart)
{
  int b = threadIdx.x;
  if(b > nblocks)
    return;
  unsigned long long target = (unsigned long long) b << (3 * G2_PH_BITS);
  int lo = 0, hi = n;
  while(lo < hi)
    {
      int mid = (lo + hi) >> 1;
      if(skey[mid] < target)
	lo = mid + 1;
      else
	hi = mid;
    }
  bstart[b] = lo;
}

// ---------------------------------------------------------------- top-level tree ---------------------------
// first index in [lo,hi) whose key is >= target (keys ascending): warp-cooperative 32-ary search, all 32 lanes
// must call; a handful of rounds instead of ~20 dependent loads.
__device__ __forceinline__ int warp_lower_bound(const long long *__restrict__ phkey, int lo, int hi, long long target)
{
  const int lane = threadIdx.x & 31;
  while(hi - lo > 32)
    {
      int step = (hi - lo + 31) >> 5;
      int idx = lo + (lane + 1) * step - 1;	// last element of this lane's chunk
      bool below = idx < hi && phkey[idx] < target;
      int t = __popc(__ballot_sync(0xffffffffu, below));	// chunks entirely below the target (a prefix of the lanes)
      lo = lo + t * step;
      int nh = lo + step;
      hi = nh < hi ? nh : hi;
    }
  int idx = lo + lane;
  bool below = idx < hi && phkey[idx] < target;
  return lo + __popc(__ballot_sync(0xffffffffu, below));
}

#define TT_THREADS 1024
#define TT_CHUNK 32		// frontier nodes refined per round
#define TT_FAST 4096		// refinement trees up to this size are numbered out of shared memory
#define TT_DYN_SMEM (TT_FAST * (8 + 5 * 4))
#define TT_MAXBLOCKS (G2GPU_MAX_GRAVS + 1)
// scratch of the breadth-first refinement (temporary numbering)
struct G2TopScratch
{
  long long start[G2_MAXTOP];
  int shift[G2_MAXTOP];
  int count[G2_MAXTOP];
  int child[G2_MAXTOP];		// first of 8 children in temporary numbering, -1 = not split
  int frontier[2][G2_MAXTOP / 8 + 8];
};

__global__ void __launch_bounds__(TT_THREADS) toptree_kernel(const long long *__restrict__ phkey, const int *__restrict__ bstart, int nblocks, int n,
							      const double *__restrict__ dom, G2TopTree *tt, G2TopScratch *ts)
{
  __shared__ int s_nfront, s_nnext, s_ntmp, s_err;
  __shared__ int s_lb[TT_CHUNK * 9 * TT_MAXBLOCKS];	// lower bounds of the 9 child boundaries per block
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const double thr = __ddiv_rn((double) n, 20.0);	// All.TotNumPart / (TOPNODEFACTOR * NTask * NTask), domain.c:1071

  if(tid == 0)
    {
      ts->start[0] = 0;
      ts->shift[0] = 3 * G2_PH_BITS;
      ts->count[0] = n;
      ts->child[0] = -1;
      s_ntmp = 1;
      s_nfront = 1;		// the root is always split (domain.c:965 calls topsplit_local unconditionally)
      ts->frontier[0][0] = 0;
      s_err = 0;
    }
  __syncthreads();
  int cur = 0;
  while(true)
    {
      const int nfront = s_nfront;
      if(nfront == 0)
	break;
      const int base = s_ntmp;
      if(base + 8 * nfront > G2_MAXTOP)
	{
	  if(tid == 0)
	    s_err = 1;
	  __syncthreads();
	  break;
	}
      if(tid == 0)
	s_nnext = 0;
      __syncthreads();
      for(int c0 = 0; c0 < nfront; c0 += TT_CHUNK)
	{
	  const int nc = nfront - c0 < TT_CHUNK ? nfront - c0 : TT_CHUNK;
	  // number of particles with PH key in a child's key range: the reference counts them in KeySorted (all
	  // species together, domain.c:1052-1066); keys are sorted per block here, so search every block.
	  const int ntask = nc * 9 * nblocks;
	  for(int task = warp; task < ntask; task += TT_THREADS / 32)
	    {
	      int fi = task / (9 * nblocks), rem = task - fi * 9 * nblocks;
	      int ci = rem / nblocks, b = rem - ci * nblocks;
	      int f = ts->frontier[cur][c0 + fi];
	      int sh = ts->shift[f] - 3;
	      long long key = ts->start[f] + ((long long) ci << sh);
	      int r = warp_lower_bound(phkey, bstart[b], bstart[b + 1], key);
	      if(lane == 0)
		s_lb[task] = r;
	    }
	  __syncthreads();
	  for(int it = tid; it < 8 * nc; it += TT_THREADS)
	    {
	      int fi = it >> 3, ci = it & 7;
	      int f = ts->frontier[cur][c0 + fi];
	      int id = base + 8 * (c0 + fi) + ci;
	      int sh = ts->shift[f] - 3;
	      int cnt = 0;
	      for(int b = 0; b < nblocks; b++)
		cnt += s_lb[(fi * 9 + ci + 1) * nblocks + b] - s_lb[(fi * 9 + ci) * nblocks + b];
	      ts->start[id] = ts->start[f] + ((long long) ci << sh);
	      ts->shift[id] = sh;
	      ts->count[id] = cnt;
	      ts->child[id] = -1;
	      if(ci == 0)
		ts->child[f] = id;
	      if((double) cnt > thr && sh >= 3)
		{
		  int slot = atomicAdd(&s_nnext, 1);
		  ts->frontier[cur ^ 1][slot] = id;
		}
	    }
	  __syncthreads();
	}
      if(tid == 0)
	{
	  s_ntmp = base + 8 * nfront;
	  s_nfront = s_nnext;
	}
      cur ^= 1;
      __syncthreads();
    }

  // ---- the numbering passes are sequential (they ARE the reference's recursion order); run them on one thread, but
  //      out of shared memory: the refinement tree is staged by all threads first (global loads on a single
  //      thread cost ~0.5 us each and dominated this kernel).  Larger trees fall back to global memory.
  extern __shared__ unsigned char tt_dyn[];
  long long *m_start = (long long *) tt_dyn;	// TT_FAST
  int *m_child = (int *) (m_start + TT_FAST);
  int *m_count = m_child + TT_FAST;
  int *m_shift = m_count + TT_FAST;
  int *m_daughter = m_shift + TT_FAST;
  int *m_leaf = m_daughter + TT_FAST;
  const int ntmp = s_ntmp;
  const bool fast = ntmp <= TT_FAST;
  if(fast)
    for(int i = tid; i < ntmp; i += TT_THREADS)
      {
	m_start[i] = ts->start[i];
	m_child[i] = ts->child[i];
	m_count[i] = ts->count[i];
	m_shift[i] = ts->shift[i];
      }
  __syncthreads();
#define TS_START(i) (fast ? m_start[i] : ts->start[i])
#define TS_CHILD(i) (fast ? m_child[i] : ts->child[i])
#define TS_COUNT(i) (fast ? m_count[i] : ts->count[i])
#define TS_SHIFT(i) (fast ? m_shift[i] : ts->shift[i])
#define TT_DAUGHTER(i) (fast ? m_daughter[i] : tt->daughter[i])
#define TT_LEAF(i) (fast ? m_leaf[i] : tt->leaf[i])
#define SET_DAUGHTER(i, v) do { const int v__ = (v); tt->daughter[i] = v__; if(fast) m_daughter[i] = v__; } while(0)
#define SET_LEAF(i, v) do { const int v__ = (v); tt->leaf[i] = v__; if(fast) m_leaf[i] = v__; } while(0)

  if(tid == 0)
    {
      tt->err = s_err;
      if(!s_err)
	{
	  int sp = 0;
	  // ---- depth-first numbering exactly as the recursion of domain_topsplit_local (domain.c:1019-1075): a node that
	  //      is split allocates its 8 daughters when it is entered, then enters the daughters that exceed the threshold
	  int stack_node[32], stack_tmp[32], stack_i[32];
	  int ntop = 1;
	  SET_DAUGHTER(0, -1);
	  SET_LEAF(0, -1);
	  tt->shift[0] = TS_SHIFT(0); tt->startkey[0] = 0; tt->count[0] = n;
	  {
	    int pend_node = 0, pend_tmp = 0;
	    bool pending = true;
	    while(pending || sp > 0)
	      {
		if(pending)
		  {		// enter (pend_node, pend_tmp)
		    pending = false;
		    int c0 = TS_CHILD(pend_tmp);
		    if(c0 >= 0)
		      {
			SET_DAUGHTER(pend_node, ntop);
			for(int i = 0; i < 8; i++)
			  {
			    SET_DAUGHTER(ntop + i, -1);
			    SET_LEAF(ntop + i, -1);
			    tt->shift[ntop + i] = TS_SHIFT(c0 + i); tt->startkey[ntop + i] = TS_START(c0 + i); tt->count[ntop + i] = TS_COUNT(c0 + i);
			  }
			stack_node[sp] = ntop; stack_tmp[sp] = c0; stack_i[sp] = 0; sp++;	// daughters base, tmp children base
			ntop += 8;
		      }
		    continue;
		  }
		int i = stack_i[sp - 1];
		if(i == 8)
		  {
		    sp--;
		    continue;
		  }
		stack_i[sp - 1] = i + 1;
		if(TS_CHILD(stack_tmp[sp - 1] + i) >= 0)
		  {
		    pend_node = stack_node[sp - 1] + i;
		    pend_tmp = stack_tmp[sp - 1] + i;
		    pending = true;
		  }
	      }
	  }
	  tt->ntopnodes = ntop;

	  // ---- leaves in Peano-Hilbert order: domain_walktoptree (domain.c:802-816)
	  int nleaves = 0;
	  {
	    int st_node[32], st_i[32];
	    sp = 0;
	    if(TT_DAUGHTER(0) == -1)
	      SET_LEAF(0, nleaves++);
	    else
	      {
		st_node[0] = 0; st_i[0] = 0; sp = 1;
	      }
	    while(sp > 0)
	      {
		int node = st_node[sp - 1], i = st_i[sp - 1];
		if(i == 8)
		  {
		    sp--;
		    continue;
		  }
		st_i[sp - 1] = i + 1;
		int sub = TT_DAUGHTER(node) + i;
		if(TT_DAUGHTER(sub) == -1)
		  SET_LEAF(sub, nleaves++);
		else
		  {
		    st_node[sp] = sub; st_i[sp] = 0; sp++;
		  }
	      }
	  }
	  tt->ntopleaves = nleaves;

	  // ---- force-tree nodes of the top-level tree: root (forcetree.c:103-110) + force_create_empty_nodes
	  //      (forcetree.c:292-336): children created in loop order i(x) outer, k(z) inner, each followed at once by
	  //      its own subtree; the TopNodes daughter of octant (i,j,k) is the last PH digit of the child's cell.
	  //      The parent's geometry travels on the stack (no read-back from global memory).
	  {
	    int nf = 1, topmaxdepth = 0;
	    int st_k[32], st_top[32], st_state[32], st_c[32], st_depth[32];
	    float st_len[32], st_cx[32], st_cy[32], st_cz[32];
	    unsigned long long st_mort[32];
	    tt->flen[0] = (float) dom[6];
	    tt->fcx[0] = (float) dom[3]; tt->fcy[0] = (float) dom[4]; tt->fcz[0] = (float) dom[5];
	    tt->fdepth[0] = 0; tt->fmorton[0] = 0; tt->ffather[0] = -1; tt->fnode[0] = 0;
	    for(int s = 0; s < 8; s++)
	      tt->fsuns[0][s] = -1;
	    tt->fisleaf[0] = (TT_DAUGHTER(0) == -1);
	    if(TT_DAUGHTER(0) == -1)
	      tt->dni[TT_LEAF(0)] = 0;
	    st_k[0] = 0; st_top[0] = 0; st_state[0] = 0; st_c[0] = 0; st_depth[0] = 0; st_mort[0] = 0;
	    st_len[0] = (float) dom[6]; st_cx[0] = (float) dom[3]; st_cy[0] = (float) dom[4]; st_cz[0] = (float) dom[5];
	    sp = 1;
	    while(sp > 0)
	      {
		const int f = sp - 1;
		const int k = st_k[f], top = st_top[f], cidx = st_c[f];
		const int dbase = TT_DAUGHTER(top);
		if(dbase < 0 || cidx == 8)
		  {
		    sp--;
		    continue;
		  }
		st_c[f] = cidx + 1;
		const int i = (cidx >> 2) & 1, j = (cidx >> 1) & 1, kk = cidx & 1;	// loop order i, j, k
		const unsigned int e = c_ph_table[st_state[f] * 8 + (i << 2 | j << 1 | kk)];
		const int sub = e & 7;
		const int slot = i + 2 * j + 4 * kk;
		const int nn = nf++;
		tt->fsuns[k][slot] = nn;
		const float plen = st_len[f];
		const double q = __dmul_rn(0.25, (double) plen);
		const float nlen = 0.5f * plen;
		const float ncx = (float) __dadd_rn((double) st_cx[f], i ? q : -q);
		const float ncy = (float) __dadd_rn((double) st_cy[f], j ? q : -q);
		const float ncz = (float) __dadd_rn((double) st_cz[f], kk ? q : -q);
		const unsigned long long nm = (st_mort[f] << 3) | (unsigned long long) slot;
		tt->flen[nn] = nlen; tt->fcx[nn] = ncx; tt->fcy[nn] = ncy; tt->fcz[nn] = ncz;
		tt->fdepth[nn] = st_depth[f] + 1;
		if(st_depth[f] + 1 > topmaxdepth)
		  topmaxdepth = st_depth[f] + 1;
		tt->fmorton[nn] = nm;
		tt->ffather[nn] = k;
		for(int s = 0; s < 8; s++)
		  tt->fsuns[nn][s] = -1;
		const int tsub = dbase + sub;
		tt->fnode[tsub] = nn;
		const bool isleaf = TT_DAUGHTER(tsub) == -1;
		tt->fisleaf[nn] = isleaf;
		if(isleaf)
		  tt->dni[TT_LEAF(tsub)] = nn;
		st_k[sp] = nn; st_top[sp] = tsub; st_state[sp] = (int) (e >> 3); st_c[sp] = 0; st_depth[sp] = st_depth[f] + 1; st_mort[sp] = nm;
		st_len[sp] = nlen; st_cx[sp] = ncx; st_cy[sp] = ncy; st_cz[sp] = ncz;
		sp++;
	      }
	    tt->pad = topmaxdepth;	// depth of the deepest top-level node
	  }
	}
      __threadfence();
    }
  __syncthreads();
  if(s_err)
    return;

  // ---- slot-order DFS rank (the order the walk visits top-level nodes: children 0..7, forcetree.c:526-540), in
  //      parallel: node A precedes B iff padded path(A) < padded path(B), or the paths tie and A is shallower.
  {
    const int ntop = *((volatile int *) &tt->ntopnodes);
    unsigned long long *m_key = (unsigned long long *) m_start;	// reuse the staging area
    int *m_depth = m_child;
    const bool fastd = ntop <= TT_FAST;
    if(fastd)
      for(int k = tid; k < ntop; k += TT_THREADS)
	{
	  int d = __ldcg(&tt->fdepth[k]);
	  m_depth[k] = d;
	  m_key[k] = d == 0 ? 0ull : (__ldcg(&tt->fmorton[k]) << (3 * (G2_MAXDEPTH - d)));
	}
    __syncthreads();
    for(int k = tid; k < ntop; k += TT_THREADS)
      {
	int dk = fastd ? m_depth[k] : __ldcg(&tt->fdepth[k]);
	unsigned long long kk = fastd ? m_key[k] : (dk == 0 ? 0ull : (__ldcg(&tt->fmorton[k]) << (3 * (G2_MAXDEPTH - dk))));
	int rank = 0;
	for(int j = 0; j < ntop; j++)
	  {
	    int dj = fastd ? m_depth[j] : __ldcg(&tt->fdepth[j]);
	    unsigned long long kj = fastd ? m_key[j] : (dj == 0 ? 0ull : (__ldcg(&tt->fmorton[j]) << (3 * (G2_MAXDEPTH - dj))));
	    rank += (kj < kk || (kj == kk && dj < dk)) ? 1 : 0;
	  }
	tt->fdfs[k] = rank;
	tt->fdfs_inv[rank] = k;
      }
  }
#undef TS_START
#undef TS_CHILD
#undef TS_COUNT
#undef TS_SHIFT
#undef TT_DAUGHTER
#undef TT_LEAF
#undef SET_DAUGHTER
#undef SET_LEAF
}

// ---------------------------------------------------------------- stand-alone key kernel (tests) ---------------
__global__ void peano_keys_kernel(const int *__restrict__ xyz, int n, int bits, long long *__restrict__ keys)
{
  __shared__ unsigned char s_ph[G2_PH_NSTATES * 8];
  ph_table_to_shared(s_ph);
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n)
    keys[i] = ph_key(s_ph, xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], bits);
}

int g2_peano_keys_standalone(g2gpu_ctx *c, int n, const int *xyz, int bits, long long *keys)
{
  int *d_xyz;
  long long *d_keys;
  G2_CUDA(cudaMalloc(&d_xyz, sizeof(int) * 3 * (size_t) n));
  G2_CUDA(cudaMalloc(&d_keys, sizeof(long long) * (size_t) n));
  G2_CUDA(cudaMemcpyAsync(d_xyz, xyz, sizeof(int) * 3 * (size_t) n, cudaMemcpyHostToDevice, c->stream));
  peano_keys_kernel<<<g2_cdiv(n, 256), 256, 0, c->stream>>>(d_xyz, n, bits, d_keys);
  c->launches++;
  G2_CUDA(cudaMemcpyAsync(keys, d_keys, sizeof(long long) * (size_t) n, cudaMemcpyDeviceToHost, c->stream));
  G2_CUDA(cudaStreamSynchronize(c->stream));
  cudaFree(d_xyz);
  cudaFree(d_keys);
  return 0;
}

// ---------------------------------------------------------------- stage driver -----------------------------
int g2_stage_domain(g2gpu_ctx *c)
{
  const int n = c->npart;
  if(n <= 0)
    return g2_fail(G2GPU_ERR_STATE, "domain: no particles uploaded");
  cudaStream_t st = c->stream;
  G2_CUDA(cudaEventRecord(c->ev[0], st));
  G2_CUDA(cudaMemsetAsync(c->d_err, 0, 4 * sizeof(int), st));

  extent_init_kernel<<<1, 32, 0, st>>>((unsigned int *) c->d_minmax);
  int nb = c->nsm * 8;
  if(nb > g2_cdiv(n, 256))
    nb = g2_cdiv(n, 256);
  extent_kernel<<<nb, 256, 0, st>>>(c->in_rec, n, (unsigned int *) c->d_minmax);
  domain_params_kernel<<<1, 32, 0, st>>>((const unsigned int *) c->d_minmax, c->d_domain);
  G2TypeMap tm;
  for(int t = 0; t < 6; t++)
    tm.t2g[t] = c->type_to_grav[t];
  keys_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(c->in_rec, n, c->d_domain, tm, c->skey[0], c->sval[0], c->d_err);
  c->launches += 4;

  // blocks 0 (gas) .. D (species D-1): 54 key bits + block bits
  int nblocks = c->D + 1;
  int blockbits = 1;
  while((1 << blockbits) < nblocks)
    blockbits++;
  unsigned long long *k = c->skey[0];
  unsigned int *v = c->sval[0];
  G2_CUDA(cudaEventRecord(c->ev[1], st));
  // the last pass (the block bits, starting at bit 54 = 6 x 9) finds the pairs in Peano-Hilbert order of ALL species and records where
  // each goes: phorder[j] = index in the final (species-major) order of the particle with PH rank j -- the walk's target order
  if(n > 1)
    G2_TRY(g2_radix_sort_pairs(c, n, &k, &v, c->skey[1], c->sval[1], 0, 3 * G2_PH_BITS + blockbits, 3 * G2_PH_BITS, c->phorder));
  else
    G2_CUDA(cudaMemsetAsync(c->phorder, 0, sizeof(unsigned int), st));
  G2_CUDA(cudaEventRecord(c->ev[2], st));

  gather_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(n, k, v, c->in_rec, c->have_vel ? c->in_vel : nullptr, c->have_gravpm ? c->in_gravpm : nullptr,
						  c->prec, c->vel, c->gravpm, c->phkey, c->perm);
  block_starts_kernel<<<1, 32, 0, st>>>(k, n, nblocks, c->d_species_start);
  if(!c->d_topscratch)
    G2_CUDA(cudaMalloc(&c->d_topscratch, sizeof(G2TopScratch)));
  G2_CUDA(cudaFuncSetAttribute(toptree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TT_DYN_SMEM));
  toptree_kernel<<<1, TT_THREADS, TT_DYN_SMEM, st>>>(c->phkey, c->d_species_start, nblocks, n, c->d_domain, c->d_top, (G2TopScratch *) c->d_topscratch);
  c->launches += 3;
  G2_CUDA(cudaMemcpyAsync(c->h_domain, c->d_domain, 8 * sizeof(double), cudaMemcpyDeviceToHost, st));	// read after the build's host synchronisation
  G2_CUDA(cudaEventRecord(c->ev[3], st));
  G2_CUDA(cudaGetLastError());
  c->stage = 2;
  return 0;
}
