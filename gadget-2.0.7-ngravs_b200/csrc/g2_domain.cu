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

__global__ void __launch_bounds__(256) extent_kernel(const float4 *__restrict__ pm, int n, unsigned int *mm)
{
  float mn[3] = { 3.0e38f, 3.0e38f, 3.0e38f }, mx[3] = { -3.0e38f, -3.0e38f, -3.0e38f };
  for(int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
      float4 p = pm[i];
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
__device__ __forceinline__ long long ph_key(int x, int y, int z, int bits)
{
  unsigned int st = 0;
  long long key = 0;
  for(int l = bits - 1; l >= 0; l--)
    {
      unsigned int o = (((x >> l) & 1) << 2) | (((y >> l) & 1) << 1) | ((z >> l) & 1);
      unsigned int e = c_ph_table[st * 8 + o];
      key = (key << 3) | (e & 7);
      st = e >> 3;
    }
  return key;
}

struct G2TypeMap { int t2g[6]; };

// sort key = (block << 54) | PH key; block 0 = gas (type 0, peano.c:47-67), block 1+g = species g (peano.c:90-133)
__global__ void __launch_bounds__(256) keys_kernel(const float4 *__restrict__ pm, const int *__restrict__ type, int n,
						   const double *__restrict__ dom, G2TypeMap tm,
						   unsigned long long *__restrict__ skey, unsigned int *__restrict__ sval, int *__restrict__ err)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n)
    return;
  float4 p = pm[i];
  double fac = dom[7];
  // (P[i].Pos[k] - DomainCorner[k]) * DomainFac, converted to int by truncation (domain.c:940-943)
  int x = __double2int_rz(__dmul_rn(__dsub_rn((double) p.x, dom[0]), fac));
  int y = __double2int_rz(__dmul_rn(__dsub_rn((double) p.y, dom[1]), fac));
  int z = __double2int_rz(__dmul_rn(__dsub_rn((double) p.z, dom[2]), fac));
  unsigned long long key = (unsigned long long) ph_key(x, y, z, G2_PH_BITS);
  int t = type[i];
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
						     const float4 *__restrict__ in_pm, const int *__restrict__ in_type,
						     const float *__restrict__ in_oldacc, const unsigned char *__restrict__ in_active,
						     const float *__restrict__ in_vel, const float *__restrict__ in_gravpm,
						     float4 *__restrict__ pm, unsigned char *__restrict__ ptype, float *__restrict__ oldacc,
						     unsigned char *__restrict__ active, float *__restrict__ vel, float *__restrict__ gravpm,
						     long long *__restrict__ phkey, int *__restrict__ perm)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n)
    return;
  unsigned int src = sval[i];
  pm[i] = in_pm[src];
  int t = in_type[src];
  ptype[i] = (unsigned char) (t < 0 ? 0 : (t > 5 ? 5 : t));
  oldacc[i] = in_oldacc[src];
  active[i] = in_active[src];
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

  if(tid != 0)
    return;
  tt->err = s_err;
  if(s_err)
    return;

  // ---- depth-first numbering exactly as the recursion of domain_topsplit_local (domain.c:1019-1075)
  int stack_node[32], stack_tmp[32], stack_i[32];
  int ntop = 1;
  tt->daughter[0] = -1; tt->shift[0] = ts->shift[0]; tt->startkey[0] = 0; tt->count[0] = n; tt->leaf[0] = -1;
  int sp = 0;
  // "enter" a node: if split, allocate its 8 children
  {
    int node = 0, tmp = 0;
    if(ts->child[tmp] >= 0)
      {
	tt->daughter[node] = ntop;
	for(int i = 0; i < 8; i++)
	  {
	    int ct = ts->child[tmp] + i;
	    tt->daughter[ntop + i] = -1; tt->shift[ntop + i] = ts->shift[ct]; tt->startkey[ntop + i] = ts->start[ct];
	    tt->count[ntop + i] = ts->count[ct]; tt->leaf[ntop + i] = -1;
	  }
	ntop += 8;
	stack_node[0] = node; stack_tmp[0] = tmp; stack_i[0] = 0; sp = 1;
      }
  }
  while(sp > 0)
    {
      int node = stack_node[sp - 1], tmp = stack_tmp[sp - 1], i = stack_i[sp - 1];
      if(i == 8)
	{
	  sp--;
	  continue;
	}
      stack_i[sp - 1] = i + 1;
      int sub = tt->daughter[node] + i, ct = ts->child[tmp] + i;
      if(ts->child[ct] >= 0)
	{
	  tt->daughter[sub] = ntop;
	  for(int j = 0; j < 8; j++)
	    {
	      int c2 = ts->child[ct] + j;
	      tt->daughter[ntop + j] = -1; tt->shift[ntop + j] = ts->shift[c2]; tt->startkey[ntop + j] = ts->start[c2];
	      tt->count[ntop + j] = ts->count[c2]; tt->leaf[ntop + j] = -1;
	    }
	  ntop += 8;
	  stack_node[sp] = sub; stack_tmp[sp] = ct; stack_i[sp] = 0; sp++;
	}
    }
  tt->ntopnodes = ntop;

  // ---- leaves in Peano-Hilbert order: domain_walktoptree (domain.c:802-816)
  int nleaves = 0;
  {
    int st_node[32], st_i[32];
    sp = 0;
    if(tt->daughter[0] == -1)
      tt->leaf[0] = nleaves++;
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
	int sub = tt->daughter[node] + i;
	if(tt->daughter[sub] == -1)
	  tt->leaf[sub] = nleaves++;
	else
	  {
	    st_node[sp] = sub; st_i[sp] = 0; sp++;
	  }
      }
  }
  tt->ntopleaves = nleaves;

  // ---- force-tree nodes of the top-level tree: root (forcetree.c:103-110) + force_create_empty_nodes
  //      (forcetree.c:292-336): children created in loop order i(x) outer, k(z) inner, each followed at once by
  //      its own subtree; TopNodes daughter of octant (i,j,k) is the last PH digit of the child's cell.
  {
    int nf = 0;
    tt->flen[0] = (float) dom[6];
    tt->fcx[0] = (float) dom[3]; tt->fcy[0] = (float) dom[4]; tt->fcz[0] = (float) dom[5];
    tt->fdepth[0] = 0; tt->fmorton[0] = 0; tt->ffather[0] = -1; tt->fnode[0] = 0;
    for(int s = 0; s < 8; s++)
      tt->fsuns[0][s] = -1;
    tt->fisleaf[0] = (tt->daughter[0] == -1);
    if(tt->fisleaf[0])
      tt->dni[tt->leaf[0]] = 0;
    nf = 1;
    int st_k[32], st_top[32], st_state[32], st_c[32];
    sp = 0;
    st_k[0] = 0; st_top[0] = 0; st_state[0] = 0; st_c[0] = 0; sp = 1;
    while(sp > 0)
      {
	int k = st_k[sp - 1], top = st_top[sp - 1], state = st_state[sp - 1], cidx = st_c[sp - 1];
	if(tt->daughter[top] < 0 || cidx == 8)
	  {
	    sp--;
	    continue;
	  }
	st_c[sp - 1] = cidx + 1;
	int i = (cidx >> 2) & 1, j = (cidx >> 1) & 1, kk = cidx & 1;	// loop order i, j, k
	unsigned int e = c_ph_table[state * 8 + (i << 2 | j << 1 | kk)];
	int sub = e & 7;
	int slot = i + 2 * j + 4 * kk;
	int nn = nf++;
	tt->fsuns[k][slot] = nn;
	float plen = tt->flen[k];
	tt->flen[nn] = 0.5f * plen;
	double q = __dmul_rn(0.25, (double) plen);
	tt->fcx[nn] = (float) __dadd_rn((double) tt->fcx[k], i ? q : -q);
	tt->fcy[nn] = (float) __dadd_rn((double) tt->fcy[k], j ? q : -q);
	tt->fcz[nn] = (float) __dadd_rn((double) tt->fcz[k], kk ? q : -q);
	tt->fdepth[nn] = tt->fdepth[k] + 1;
	tt->fmorton[nn] = (tt->fmorton[k] << 3) | (unsigned long long) slot;
	tt->ffather[nn] = k;
	for(int s = 0; s < 8; s++)
	  tt->fsuns[nn][s] = -1;
	int tsub = tt->daughter[top] + sub;
	tt->fnode[tsub] = nn;
	tt->fisleaf[nn] = (tt->daughter[tsub] == -1);
	if(tt->fisleaf[nn])
	  tt->dni[tt->leaf[tsub]] = nn;
	st_k[sp] = nn; st_top[sp] = tsub; st_state[sp] = (int) (e >> 3); st_c[sp] = 0; sp++;
      }
    // nf == ntop by construction
  }

  // ---- slot-order DFS rank (the order the walk visits top-level nodes: children 0..7, forcetree.c:526-540)
  {
    int st_k[32], st_s[32];
    int rank = 0;
    sp = 0;
    tt->fdfs[0] = rank; tt->fdfs_inv[rank] = 0; rank++;
    st_k[0] = 0; st_s[0] = 0; sp = 1;
    while(sp > 0)
      {
	int k = st_k[sp - 1], s = st_s[sp - 1];
	if(s == 8)
	  {
	    sp--;
	    continue;
	  }
	st_s[sp - 1] = s + 1;
	int ch = tt->fsuns[k][s];
	if(ch >= 0)
	  {
	    tt->fdfs[ch] = rank; tt->fdfs_inv[rank] = ch; rank++;
	    st_k[sp] = ch; st_s[sp] = 0; sp++;
	  }
      }
  }
}

// ---------------------------------------------------------------- stand-alone key kernel (tests) ---------------
__global__ void peano_keys_kernel(const int *__restrict__ xyz, int n, int bits, long long *__restrict__ keys)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n)
    keys[i] = ph_key(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], bits);
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
  extent_kernel<<<nb, 256, 0, st>>>(c->in_pm, n, (unsigned int *) c->d_minmax);
  domain_params_kernel<<<1, 32, 0, st>>>((const unsigned int *) c->d_minmax, c->d_domain);
  G2TypeMap tm;
  for(int t = 0; t < 6; t++)
    tm.t2g[t] = c->type_to_grav[t];
  keys_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(c->in_pm, c->in_type, n, c->d_domain, tm, c->skey[0], c->sval[0], c->d_err);
  c->launches += 4;

  // blocks 0 (gas) .. D (species D-1): 54 key bits + block bits
  int nblocks = c->D + 1;
  int blockbits = 1;
  while((1 << blockbits) < nblocks)
    blockbits++;
  unsigned long long *k = c->skey[0];
  unsigned int *v = c->sval[0];
  G2_CUDA(cudaEventRecord(c->ev[1], st));
  G2_TRY(g2_radix_sort_pairs(c, n, &k, &v, c->skey[1], c->sval[1], 0, 3 * G2_PH_BITS + blockbits));
  G2_CUDA(cudaEventRecord(c->ev[2], st));

  gather_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(n, k, v, c->in_pm, c->in_type, c->in_oldacc, c->in_active,
						  c->have_vel ? c->in_vel : nullptr, c->have_gravpm ? c->in_gravpm : nullptr,
						  c->pm, c->ptype, c->oldacc, c->active, c->vel, c->gravpm, c->phkey, c->perm);
  block_starts_kernel<<<1, 32, 0, st>>>(k, n, nblocks, c->d_species_start);
  if(!c->d_topscratch)
    G2_CUDA(cudaMalloc(&c->d_topscratch, sizeof(G2TopScratch)));
  toptree_kernel<<<1, TT_THREADS, 0, st>>>(c->phkey, c->d_species_start, nblocks, n, c->d_domain, c->d_top, (G2TopScratch *) c->d_topscratch);
  c->launches += 3;
  G2_CUDA(cudaEventRecord(c->ev[3], st));
  G2_CUDA(cudaGetLastError());
  c->stage = 2;
  return 0;
}
