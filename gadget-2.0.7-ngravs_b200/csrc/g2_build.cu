// g2_build.cu — stage 2: the oct-tree of force_treebuild (forcetree.c:61-281) built in parallel.
//
// The reference inserts particles one by one (forcetree.c:136-257).  The resulting tree is a pure function
// of the particle set: every octree cell below a top-level leaf that holds >= 2 particles is a node, and the
// octant of a particle in a cell is decided by `Pos > center` on FLOAT centres obtained by repeated
// center +- 0.25*len (forcetree.c:160-165, 188-206).  We reproduce exactly that geometry:
//   1. tree key = octant path of the particle (3 bits/level, G2_MAXDEPTH levels): the top-level part comes from
//      the integer PH key (forcetree.c:144-153), the rest from the float comparisons;
//   2. radix sort by tree key -> depth-first (slot-order) sequence of particles;
//   3. common-prefix depths of neighbours give every cell with >= 2 particles (no path compression);
//   4. a top-down pass finds each cell's children, a bottom-up pass (last-arriving child continues upward)
//      accumulates per-species monopoles in FP64 exactly like force_update_node_recursive (forcetree.c:451-743),
//      children in slot order 0..7, children's values re-read as rounded FLOATs;
//   5. cells are laid out in depth-first order (index U) for the walk; the reference's insertion-order node
//      numbers are derived on request (g2_stage_renumber) from the second-smallest particle index per cell.
#include "g2_common.cuh"

#define KEYBITS (3 * G2_MAXDEPTH)

struct G2Soft
{
  int t2g[6];
  double fsoft[6];		// All.ForceSoftening
  int unequal;
};

// ---------------------------------------------------------------- helpers ----------------------------------
__device__ __forceinline__ int top_leaf_of_key(const G2TopTree *__restrict__ tt, long long key)
{
  // forcetree.c:148-150
  int no = 0;
  while(tt->daughter[no] >= 0)
    no = tt->daughter[no] + (int) ((key - tt->startkey[no]) >> (tt->shift[no] - 3));
  return no;
}

__device__ __forceinline__ unsigned int digit_at(unsigned long long key, int level)	// level 1..G2_MAXDEPTH
{
  return (unsigned int) (key >> (3 * (G2_MAXDEPTH - level))) & 7u;
}

__device__ __forceinline__ unsigned long long prefix_of(unsigned long long key, int depth)
{
  return depth >= G2_MAXDEPTH ? key : (key >> (3 * (G2_MAXDEPTH - depth)));
}

// child geometry: forcetree.c:188-206 (lenhalf is double, centres are stored FLOAT)
__device__ __forceinline__ void descend(float &cx, float &cy, float &cz, float &len, unsigned int sub)
{
  double q = __dmul_rn(0.25, (double) len);
  cx = (float) __dadd_rn((double) cx, (sub & 1) ? q : -q);
  cy = (float) __dadd_rn((double) cy, (sub & 2) ? q : -q);
  cz = (float) __dadd_rn((double) cz, (sub & 4) ? q : -q);
  len = 0.5f * len;
}

// ---------------------------------------------------------------- 1. tree keys -------------------------------
__global__ void __launch_bounds__(256) treekey_kernel(const G2PRec *__restrict__ prec, const long long *__restrict__ phkey, int n,
						      const G2TopTree *__restrict__ tt, unsigned long long *__restrict__ tkey,
						      unsigned int *__restrict__ tval, unsigned short *__restrict__ ptl)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n)
    return;
  int leaf = top_leaf_of_key(tt, phkey[i]);
  int k = tt->fnode[leaf];
  float cx = tt->fcx[k], cy = tt->fcy[k], cz = tt->fcz[k], len = tt->flen[k];
  int t = tt->fdepth[k];
  unsigned long long path = tt->fmorton[k];
  const float4 p = *((const float4 *) &prec[i]);
  // `(FLOAT) (center +- 0.25*len)` of forcetree.c:190-206 is a double sum rounded to float.  When the exponents of the
  // centre and of len/4 at the deepest level differ by less than 29 bits the double sum is exact, so a plain float
  // addition (one rounding of the same exact value) gives the identical result without FP64 conversions.
  const float cmax = fmaxf(fmaxf(fabsf(cx), fabsf(cy)), fabsf(cz)) + len;
  const bool float_exact = cmax < len * (float) (1 << (t + 5 > 30 ? 30 : t + 5));	// |c| / (len_leaf / 2^(23-t)) < 2^28
  if(float_exact)
    for(int lvl = t + 1; lvl <= G2_MAXDEPTH; lvl++)
      {
	unsigned int sub = (p.x > cx ? 1u : 0u) | (p.y > cy ? 2u : 0u) | (p.z > cz ? 4u : 0u);
	path = (path << 3) | sub;
	const float q = 0.25f * len;
	cx = __fadd_rn(cx, (sub & 1) ? q : -q);
	cy = __fadd_rn(cy, (sub & 2) ? q : -q);
	cz = __fadd_rn(cz, (sub & 4) ? q : -q);
	len = 0.5f * len;
      }
  else
    for(int lvl = t + 1; lvl <= G2_MAXDEPTH; lvl++)
      {
	unsigned int sub = (p.x > cx ? 1u : 0u) | (p.y > cy ? 2u : 0u) | (p.z > cz ? 4u : 0u);
	path = (path << 3) | sub;
	descend(cx, cy, cz, len, sub);
      }
  tkey[i] = path;
  tval[i] = (unsigned int) i;
  ptl[i] = (unsigned short) k;
}

// ---------------------------------------------------------------- 2. neighbour depths ------------------------
// tm[p] = number of common levels of sorted positions p and p+1, or 255 if they lie in different top-level
// leaves (or p is the last particle); tl[p] = top node of position p.
__global__ void __launch_bounds__(256) pair_depth_kernel(const unsigned long long *__restrict__ tkey, const unsigned int *__restrict__ tq,
							 const unsigned short *__restrict__ ptl, int n, unsigned char *__restrict__ tm,
							 unsigned short *__restrict__ tl, int *__restrict__ err)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if(p >= n)
    return;
  unsigned short k = ptl[tq[p]];
  tl[p] = k;
  unsigned char m = 255;
  if(p + 1 < n && ptl[tq[p + 1]] == k)
    {
      unsigned long long x = tkey[p] ^ tkey[p + 1];
      if(x == 0)
	{
	  // Particles that share all G2_MAXDEPTH levels end up together in one cell of that depth (a bucket of up to 8 direct
	  // particles, see find_children).  The reference would go on subdividing -- or, at this scale, pick random subnodes
	  // (forcetree.c:208-232: len < 1e-3 ForceSoftening) -- so its tree is not defined by geometry there either.
	  m = G2_MAXDEPTH;
	  int k = 1;
	  while(k < 8 && p - k >= 0 && tkey[p - k] == tkey[p])
	    k++;
	  if(k >= 8)
	    atomicExch(&err[0], G2GPU_ERR_TREE_DEPTH);	// 9 or more: does not fit the 8 child slots of a node
	}
      else
	m = (unsigned char) ((__clzll((long long) x) - (64 - KEYBITS)) / 3);
    }
  tm[p] = m;
}

// number of cells owned by pair p = cells whose leftmost particle is p:
// depths max(m[p-1], t)+1 .. m[p]   (t = depth of the top-level leaf)
__device__ __forceinline__ int first_depth(const unsigned char *__restrict__ tm, const unsigned short *__restrict__ tl,
					   const G2TopTree *__restrict__ tt, int p)
{
  int t = tt->fdepth[tl[p]];
  int mprev = -1;
  if(p > 0)
    {
      unsigned char v = tm[p - 1];
      if(v != 255)
	mprev = v;
    }
  return (mprev > t ? mprev : t) + 1;
}

__global__ void __launch_bounds__(256) cell_count_kernel(const unsigned char *__restrict__ tm, const unsigned short *__restrict__ tl,
							 const G2TopTree *__restrict__ tt, int n, unsigned int *__restrict__ cnt)
{
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if(p >= n)
    return;
  int c = 0;
  unsigned char m = tm[p];
  if(m != 255)
    {
      c = (int) m - first_depth(tm, tl, tt, p) + 1;
      if(c < 0)
	c = 0;
    }
  cnt[p] = (unsigned int) c;
}

// ---------------------------------------------------------------- 3. cells -----------------------------------
__global__ void __launch_bounds__(256) cell_init_kernel(const unsigned char *__restrict__ tm, const unsigned short *__restrict__ tl,
							const G2TopTree *__restrict__ tt, const unsigned int *__restrict__ tbase, int n,
							int max_cells, unsigned int *__restrict__ c_a, unsigned char *__restrict__ c_d,
							int *__restrict__ err, unsigned int *__restrict__ depth_count)
{
  __shared__ unsigned int s_hist[32];
  if(threadIdx.x < 32)
    s_hist[threadIdx.x] = 0;
  __syncthreads();
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if(p < n)
    {
      unsigned int b0 = tbase[p], b1 = tbase[p + 1];
      if(b1 > b0)
	{
	  if(b1 > (unsigned int) max_cells)
	    atomicExch(&err[1], G2GPU_ERR_MAXNODES);
	  else
	    {
	      int fd = first_depth(tm, tl, tt, p);
	      for(unsigned int c = b0; c < b1; c++)
		{
		  int d = fd + (int) (c - b0);
		  c_a[c] = (unsigned int) p;
		  c_d[c] = (unsigned char) d;
		  atomicAdd(&s_hist[d & 31], 1u);
		}
	    }
	}
    }
  __syncthreads();
  if(threadIdx.x < 32 && s_hist[threadIdx.x])
    atomicAdd(&depth_count[threadIdx.x], s_hist[threadIdx.x]);	// cells per depth: the moment pass runs level by level
}

struct DepthStarts { unsigned int start[32]; };

// cell ids grouped by depth (order inside a depth is irrelevant): block-aggregated slot allocation
__global__ void __launch_bounds__(256) depth_scatter_kernel(const unsigned char *__restrict__ c_d, int ncells, DepthStarts ds, unsigned int *__restrict__ cursor,
							    unsigned int *__restrict__ depth_list)
{
  __shared__ unsigned int s_cnt[32], s_base[32];
  if(threadIdx.x < 32)
    s_cnt[threadIdx.x] = 0;
  __syncthreads();
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  int d = c < ncells ? (int) c_d[c] & 31 : -1;
  unsigned int rank = 0;
  if(d >= 0)
    rank = atomicAdd(&s_cnt[d], 1u);
  __syncthreads();
  if(threadIdx.x < 32 && s_cnt[threadIdx.x])
    s_base[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], s_cnt[threadIdx.x]);
  __syncthreads();
  if(d >= 0)
    depth_list[ds.start[d] + s_base[d] + rank] = (unsigned int) c;
}

// last position in [pos, hi] sharing the first `depth` digits with position pos (keys sorted ascending)
__device__ __forceinline__ unsigned int run_end(const unsigned long long *__restrict__ tkey, unsigned int pos, unsigned int hi, int depth)
{
  const unsigned long long pre = prefix_of(tkey[pos], depth);
  unsigned int step = 1, lo = pos;	// invariant: lo is inside the run
  while(lo + step <= hi && prefix_of(tkey[lo + step], depth) == pre)
    {
      lo += step;
      step <<= 1;
    }
  unsigned int top = lo + step - 1;	// first candidate known (or assumed) outside is lo+step
  if(top > hi)
    top = hi;
  // binary search in (lo, top] for the last inside
  while(lo < top)
    {
      unsigned int mid = lo + (top - lo + 1) / 2;
      if(prefix_of(tkey[mid], depth) == pre)
	lo = mid;
      else
	top = mid - 1;
    }
  return lo;
}

// enumerate the children of the cell [a,b] at depth d: suns[slot] = particle position (>= 0), -(cell+2), or -1
__device__ __forceinline__ void find_children(const unsigned long long *__restrict__ tkey, const unsigned char *__restrict__ tm,
					      const unsigned short *__restrict__ tl, const G2TopTree *__restrict__ tt,
					      const unsigned int *__restrict__ tbase, unsigned int a, unsigned int b, int d,
					      int self_code, int *__restrict__ suns, int *__restrict__ c_father, int *__restrict__ p_parent,
					      int &nchild, int &npart)
{
  nchild = 0;
  npart = 0;
#pragma unroll
  for(int s = 0; s < 8; s++)
    suns[s] = -1;
  unsigned int pos = a;
  if(d >= G2_MAXDEPTH)
    {				// bucket at the deepest level: its (at most 8, pair_depth_kernel) particles are direct children
      for(; pos <= b && pos - a < 8u; pos++)
	{
	  suns[pos - a] = (int) pos;
	  p_parent[pos] = self_code;
	  npart++;
	}
      return;
    }
  while(pos <= b)
    {
      unsigned int slot = digit_at(tkey[pos], d + 1);
      unsigned int e = run_end(tkey, pos, b, d + 1);
      if(e == pos)
	{
	  suns[slot] = (int) pos;
	  p_parent[pos] = self_code;
	  npart++;
	}
      else
	{
	  unsigned int cc = tbase[pos] + (unsigned int) ((d + 1) - first_depth(tm, tl, tt, (int) pos));
	  suns[slot] = -((int) cc + 2);
	  c_father[cc] = self_code;
	  nchild++;
	}
      pos = e + 1;
    }
}

// father / p_parent codes: >= 0 regular cell id, <= -2 top node k = -(code+2)
__global__ void __launch_bounds__(128) topdown_kernel(const unsigned long long *__restrict__ tkey, const unsigned char *__restrict__ tm,
						      const unsigned short *__restrict__ tl, const G2TopTree *__restrict__ tt,
						      const unsigned int *__restrict__ tbase, int n, int max_cells,
						      const unsigned int *__restrict__ c_a, const unsigned char *__restrict__ c_d,
						      unsigned int *__restrict__ c_b, int *__restrict__ c_suns, int *__restrict__ c_father,
						      int *__restrict__ p_parent, unsigned char *__restrict__ c_nchild,
						      unsigned char *__restrict__ c_npart, unsigned int *__restrict__ u_npart)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  int ncells = (int) tbase[n];
  if(c >= ncells || c >= max_cells)
    return;
  unsigned int a = c_a[c];
  int d = c_d[c];
  unsigned int b = run_end(tkey, a, (unsigned int) (n - 1), d);
  c_b[c] = b;
  int suns[8], nchild, npart;
  find_children(tkey, tm, tl, tt, tbase, a, b, d, c, suns, c_father, p_parent, nchild, npart);
#pragma unroll
  for(int s = 0; s < 8; s++)
    c_suns[8 * (size_t) c + s] = suns[s];
  c_nchild[c] = (unsigned char) nchild;
  c_npart[c] = (unsigned char) npart;
  int U = c + tt->fdfs[tl[a]] + 1;
  u_npart[U] = (unsigned int) npart;
}

// top-level nodes: particle range, children of top leaves, U index
__global__ void __launch_bounds__(128) top_children_kernel(const unsigned long long *__restrict__ tkey, const unsigned char *__restrict__ tm,
							   const unsigned short *__restrict__ tl, const G2TopTree *__restrict__ tt,
							   const unsigned int *__restrict__ tbase, int n, int *__restrict__ t_suns,
							   int *__restrict__ c_father, int *__restrict__ p_parent, unsigned int *__restrict__ t_first,
							   unsigned int *__restrict__ t_last, unsigned char *__restrict__ t_nchild,
							   unsigned char *__restrict__ t_npart, unsigned int *__restrict__ t_u,
							   unsigned int *__restrict__ u_npart)
{
  int k = blockIdx.x * blockDim.x + threadIdx.x;
  if(k >= tt->ntopnodes)
    return;
  int t = tt->fdepth[k];
  // positions whose key starts with this node's slot path
  unsigned long long lo_key = t == 0 ? 0ull : (tt->fmorton[k] << (3 * (G2_MAXDEPTH - t)));
  unsigned long long hi_key = t == 0 ? (1ull << KEYBITS) : ((tt->fmorton[k] + 1ull) << (3 * (G2_MAXDEPTH - t)));
  unsigned int lo = 0, hi = (unsigned int) n;
  while(lo < hi)
    {
      unsigned int mid = (lo + hi) >> 1;
      if(tkey[mid] < lo_key)
	lo = mid + 1;
      else
	hi = mid;
    }
  unsigned int first = lo;
  hi = (unsigned int) n;
  while(lo < hi)
    {
      unsigned int mid = (lo + hi) >> 1;
      if(tkey[mid] < hi_key)
	lo = mid + 1;
      else
	hi = mid;
    }
  unsigned int end = lo;	// one past the last
  t_first[k] = first;
  t_last[k] = end;		// exclusive end
  unsigned int U = (unsigned int) tt->fdfs[k] + tbase[first];
  t_u[k] = U;
  int nchild = 0, npart = 0;
  int suns[8];
#pragma unroll
  for(int s = 0; s < 8; s++)
    suns[s] = -1;
  if(tt->fisleaf[k] && end > first)
    find_children(tkey, tm, tl, tt, tbase, first, end - 1, t, -(k + 2), suns, c_father, p_parent, nchild, npart);
#pragma unroll
  for(int s = 0; s < 8; s++)
    t_suns[8 * k + s] = suns[s];
  t_nchild[k] = (unsigned char) nchild;
  t_npart[k] = (unsigned char) npart;
  u_npart[U] = (unsigned int) npart;
}

// ---------------------------------------------------------------- 4. moments, bottom-up ------------------------
// Walk record of a cell, (2+D) float4, index U (depth-first order):
//   [0]       len, center x, y, z                        (struct NODE len, center; allvars.h:620-621)
//   [1+g]     s[0][g], s[1][g], s[2][g], mass[g]         (u.d.s, u.d.mass; allvars.h:642-643)
//   [1+D]     as uint: sibling U | particle offset | pinfo | hmax
//             pinfo = count (4 bits) | type of direct particle j at bits 4+3j | mixed-softening bit 28 | maxsofttype
//             at bits 29-31 (u.d.bitflags content, forcetree.c:706); hmax = ForceSoftening[maxsofttype] as float bits
template <int D>
struct Moments
{
  double mass[D], sx[D], sy[D], sz[D];
  int maxsofttype, diffsoft;
  unsigned int min1, min2;
};

template <int D>
__device__ __forceinline__ void mom_init(Moments<D> &M)
{
#pragma unroll
  for(int g = 0; g < D; g++)
    M.mass[g] = M.sx[g] = M.sy[g] = M.sz[g] = 0.0;
  M.maxsofttype = 7;
  M.diffsoft = 0;
  M.min1 = M.min2 = 0xffffffffu;
}

__device__ __forceinline__ void min2_insert(unsigned int &m1, unsigned int &m2, unsigned int v)
{
  if(v < m1)
    {
      m2 = m1;
      m1 = v;
    }
  else if(v < m2)
    m2 = v;
}

// forcetree.c:571-597 / 629-646: running (maxsofttype, diffsoftflag) update with one more type
__device__ __forceinline__ void soft_update(int &maxsofttype, int &diffsoft, int type, const G2Soft &S)
{
  if(maxsofttype == 7)
    maxsofttype = type;
  else if(type != 7)
    {
      if(S.fsoft[type] > S.fsoft[maxsofttype])
	{
	  maxsofttype = type;
	  diffsoft = 1;
	}
      else if(S.fsoft[type] < S.fsoft[maxsofttype])
	diffsoft = 1;
    }
}

template <int D>
__device__ __forceinline__ void mom_add_particle(Moments<D> &M, float4 p, int type, unsigned int idx, const G2Soft &S)
{
  int g = S.t2g[type];
  double m = (double) p.w;
#pragma unroll
  for(int gg = 0; gg < D; gg++)
    if(gg == g)
      {
	// `pa->Mass * pa->Pos[k]` is a FLOAT*FLOAT product, rounded to float before it is added to the
	// double accumulator (forcetree.c:611-619)
	M.mass[gg] = __dadd_rn(M.mass[gg], m);
	M.sx[gg] = __dadd_rn(M.sx[gg], (double) __fmul_rn(p.w, p.x));
	M.sy[gg] = __dadd_rn(M.sy[gg], (double) __fmul_rn(p.w, p.y));
	M.sz[gg] = __dadd_rn(M.sz[gg], (double) __fmul_rn(p.w, p.z));
      }
  if(S.unequal)
    soft_update(M.maxsofttype, M.diffsoft, type, S);
  min2_insert(M.min1, M.min2, idx);
}

template <int D>
__device__ __forceinline__ void mom_add_cell(Moments<D> &M, const float4 *__restrict__ rec, unsigned int cmin1, unsigned int cmin2,
					     const G2Soft &S)
{
#pragma unroll
  for(int g = 0; g < D; g++)
    {
      float4 q = __ldcg(&rec[1 + g]);
      double m = (double) q.w;
      // FLOAT*FLOAT products as in forcetree.c:555-567
      M.mass[g] = __dadd_rn(M.mass[g], m);
      M.sx[g] = __dadd_rn(M.sx[g], (double) __fmul_rn(q.w, q.x));
      M.sy[g] = __dadd_rn(M.sy[g], (double) __fmul_rn(q.w, q.y));
      M.sz[g] = __dadd_rn(M.sz[g], (double) __fmul_rn(q.w, q.z));
    }
  if(S.unequal)
    {
      uint4 w = __ldcg((const uint4 *) &rec[1 + D]);
      int ctype = (w.z >> 29) & 7;
      if(ctype != 7)		// (an empty child carries the bit only as the walk's always-open marker)
	M.diffsoft |= (w.z >> 28) & 1;	// forcetree.c:571
      if(M.maxsofttype == 7)
	M.maxsofttype = ctype;
      else if(ctype != 7)
	{
	  if(S.fsoft[ctype] > S.fsoft[M.maxsofttype])
	    {
	      M.maxsofttype = ctype;
	      M.diffsoft = 1;
	    }
	  else if(S.fsoft[ctype] < S.fsoft[M.maxsofttype])
	    M.diffsoft = 1;
	}
    }
  min2_insert(M.min1, M.min2, cmin1);
  min2_insert(M.min1, M.min2, cmin2);
}

template <int D>
__device__ __forceinline__ void mom_store(const Moments<D> &M, float4 *__restrict__ rec, float len, float cx, float cy, float cz,
					  unsigned int sibU, unsigned int poff, unsigned int pinfo, unsigned int extra_flags,
					  const G2Soft &S)
{
  rec[0] = make_float4(len, cx, cy, cz);
#pragma unroll
  for(int g = 0; g < D; g++)
    {
      float4 q;
      if(M.mass[g] > 0)
	{			// forcetree.c:667-677
	  q.x = (float) __ddiv_rn(M.sx[g], M.mass[g]);
	  q.y = (float) __ddiv_rn(M.sy[g], M.mass[g]);
	  q.z = (float) __ddiv_rn(M.sz[g], M.mass[g]);
	}
      else
	{			// forcetree.c:678-683: empty species sits at the geometric centre
	  q.x = cx;
	  q.y = cy;
	  q.z = cz;
	}
      q.w = (float) M.mass[g];
      rec[1 + g] = q;
    }
  // word 2: pinfo | mixed-softening bit (28) | maxsofttype (29-31); word 3: ForceSoftening[maxsofttype] as float.
  // An empty node (maxsofttype 7) is always opened by the walk (forcetree.c:1478-1483): +inf and the mixed bit do that.
  unsigned int tbits = 0u, hbits = 0u;
  if(S.unequal)
    {
      if(M.maxsofttype == 7)
	{
	  tbits = (1u << 28) | (7u << 29);
	  hbits = 0x7f800000u;
	}
      else
	{
	  tbits = ((unsigned int) M.diffsoft << 28) | ((unsigned int) M.maxsofttype << 29);
	  hbits = __float_as_uint((float) S.fsoft[M.maxsofttype]);
	}
    }
  (void) extra_flags;
  uint4 w = make_uint4(sibU, poff, pinfo | tbits, hbits);
  *((uint4 *) &rec[1 + D]) = w;
}

struct BuildArrays
{
  const unsigned long long *tkey;
  const unsigned int *tq;
  const unsigned short *tl;
  const unsigned int *tbase;
  const G2PRec *prec;
  const unsigned int *c_a, *c_b;
  const unsigned char *c_d;
  const int *c_suns, *c_father;
  const unsigned char *c_nchild;
  unsigned int *c_ready;
  unsigned int *c_min1, *c_min2;
  const int *t_suns;
  const unsigned int *t_first, *t_last, *t_u;
  const unsigned char *t_nchild;
  unsigned int *t_ready;
  unsigned int *t_min1, *t_min2;
  const unsigned int *u_poff;	// exclusive scan of u_npart
  float4 *wcells;
  float4 *wpart;
  unsigned int *wsrc;		// particle index (current order) behind every record of wpart
  int n;
  int numnodes_cap;
};

// process one regular cell c; returns its father code
template <int D>
__device__ int process_cell(const BuildArrays &A, const G2TopTree *__restrict__ tt, const G2Soft &S, int c)
{
  const unsigned int a = A.c_a[c], b = A.c_b[c];
  const int d = A.c_d[c];
  const int leafk = A.tl[a];
  const unsigned int U = (unsigned int) c + (unsigned int) tt->fdfs[leafk] + 1u;
  // geometry from the top-level leaf down (forcetree.c:188-206)
  float cx = tt->fcx[leafk], cy = tt->fcy[leafk], cz = tt->fcz[leafk], len = tt->flen[leafk];
  const unsigned long long key = A.tkey[a];
  for(int lvl = tt->fdepth[leafk] + 1; lvl <= d; lvl++)
    descend(cx, cy, cz, len, digit_at(key, lvl));

  Moments<D> M;
  mom_init(M);
  unsigned int poff = A.u_poff[U], pinfo = 0, np = 0;
  for(int s = 0; s < 8; s++)	// slot order 0..7 (forcetree.c:526)
    {
      int v = A.c_suns[8 * (size_t) c + s];
      if(v == -1)
	continue;
      if(v >= 0)
	{
	  unsigned int idx = A.tq[v];
	  const float4 p = *((const float4 *) &A.prec[idx]);
	  const int type = A.prec[idx].type;
	  mom_add_particle<D>(M, p, type, idx, S);
	  A.wpart[poff + np] = p;
	  A.wsrc[poff + np] = idx;
	  pinfo |= (unsigned int) type << (4 + 3 * np);
	  np++;
	}
      else
	{
	  int cc = -(v + 2);
	  unsigned int Uc = (unsigned int) cc + (unsigned int) tt->fdfs[leafk] + 1u;
	  mom_add_cell<D>(M, A.wcells + (size_t) Uc * (2 + D), __ldcg(&A.c_min1[cc]), __ldcg(&A.c_min2[cc]), S);
	}
    }
  pinfo |= np;
  unsigned int sibU = U + (A.tbase[b + 1] - (unsigned int) c);
  mom_store<D>(M, A.wcells + (size_t) U * (2 + D), len, cx, cy, cz, sibU, poff, pinfo, 0u, S);
  A.c_min1[c] = M.min1;
  A.c_min2[c] = M.min2;
  return A.c_father[c];
}

// process top-level node k
template <int D>
__device__ int process_top(const BuildArrays &A, const G2TopTree *__restrict__ tt, const G2Soft &S, int k)
{
  const unsigned int U = A.t_u[k];
  Moments<D> M;
  mom_init(M);
  unsigned int poff = A.u_poff[U], pinfo = 0, np = 0;
  const int isleaf = tt->fisleaf[k];
  for(int s = 0; s < 8; s++)
    {
      if(isleaf)
	{
	  int v = A.t_suns[8 * k + s];
	  if(v == -1)
	    continue;
	  if(v >= 0)
	    {
	      unsigned int idx = A.tq[v];
	      const float4 p = *((const float4 *) &A.prec[idx]);
	      const int type = A.prec[idx].type;
	      mom_add_particle<D>(M, p, type, idx, S);
	      A.wpart[poff + np] = p;
	      A.wsrc[poff + np] = idx;
	      pinfo |= (unsigned int) type << (4 + 3 * np);
	      np++;
	    }
	  else
	    {
	      int cc = -(v + 2);
	      unsigned int Uc = (unsigned int) cc + (unsigned int) tt->fdfs[k] + 1u;
	      mom_add_cell<D>(M, A.wcells + (size_t) Uc * (2 + D), __ldcg(&A.c_min1[cc]), __ldcg(&A.c_min2[cc]), S);
	    }
	}
      else
	{
	  int ch = tt->fsuns[k][s];
	  if(ch < 0)
	    continue;
	  mom_add_cell<D>(M, A.wcells + (size_t) A.t_u[ch] * (2 + D), __ldcg(&A.t_min1[ch]), __ldcg(&A.t_min2[ch]), S);
	}
    }
  pinfo |= np;
  // next top-level node in walk order that is not a descendant: skip this node's top subtree
  unsigned int sibU;
  {
    // size of the top subtree = number of top nodes whose slot path starts with this node's path; they are
    // contiguous in fdfs order, so find the end by scanning fdfs successors' depth
    int r = tt->fdfs[k] + 1;
    const int t = tt->fdepth[k];
    while(r < tt->ntopnodes && tt->fdepth[tt->fdfs_inv[r]] > t)
      r++;
    sibU = (r < tt->ntopnodes) ? A.t_u[tt->fdfs_inv[r]] : (unsigned int) (tt->ntopnodes + (int) A.tbase[A.n]);
  }
  mom_store<D>(M, A.wcells + (size_t) U * (2 + D), tt->flen[k], tt->fcx[k], tt->fcy[k], tt->fcz[k], sibU, poff, pinfo, 1u << 8, S);
  A.t_min1[k] = M.min1;
  A.t_min2[k] = M.min2;
  return tt->ffather[k] >= 0 ? -(tt->ffather[k] + 2) : -1;
}

// Moment pass, one launch per depth (deepest first): every cell / top-level node of that depth is processed by one
// thread; its children were completed by the previous launches, so no flags, fences or atomics are needed.
template <int D>
__global__ void __launch_bounds__(128) level_kernel(BuildArrays A, const G2TopTree *__restrict__ tt, G2Soft S, int depth, unsigned int lstart, int lcount,
						    const unsigned int *__restrict__ depth_list)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < lcount)
    process_cell<D>(A, tt, S, (int) depth_list[lstart + i]);
  else
    {
      int k = i - lcount;
      if(k < tt->ntopnodes && tt->fdepth[k] == depth)
	process_top<D>(A, tt, S, k);
    }
}

// wvs (depth-first index U) -> reference numbering
__global__ void __launch_bounds__(128) vs_permute_kernel(BuildArrays A, const G2TopTree *__restrict__ tt, const unsigned int *__restrict__ c_refid,
							 int ncells, int D, const float *__restrict__ wvs, float *__restrict__ out)
{
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int ntop = tt->ntopnodes;
  unsigned int U;
  int ref;
  if(tid < ncells)
    {
      U = (unsigned int) tid + (unsigned int) tt->fdfs[A.tl[A.c_a[tid]]] + 1u;
      ref = (int) c_refid[tid];
    }
  else if(tid < ncells + ntop)
    {
      int k = tid - ncells;
      U = A.t_u[k];
      ref = k;
    }
  else
    return;
  for(int q = 0; q < 3 * D; q++)
    out[(size_t) ref * 3 * D + q] = wvs[(size_t) U * 3 * D + q];
}

// ---------------------------------------------------------------- velocity moments (Extnodes[].vs) -----------------
// forcetree.c:563-567, 617-619, 667-677, 692-694: vs = sum(m v) / sum(m) per species, FLOAT*FLOAT products accumulated in
// double, children in slot order.  Only needed for the host mirror (dynamic tree updates, SPH), so it is a separate
// on-demand pass over the same depth lists; wvs[U][3][D].
template <int D>
__global__ void __launch_bounds__(128) vs_level_kernel(BuildArrays A, const G2TopTree *__restrict__ tt, G2Soft S, const float *__restrict__ vel,
						       float *__restrict__ wvs, int depth, unsigned int lstart, int lcount,
						       const unsigned int *__restrict__ depth_list)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int *suns;
  unsigned int U;
  int leafk, istop = 0, k = -1;
  if(i < lcount)
    {
      int c = (int) depth_list[lstart + i];
      suns = A.c_suns + 8 * (size_t) c;
      leafk = A.tl[A.c_a[c]];
      U = (unsigned int) c + (unsigned int) tt->fdfs[leafk] + 1u;
    }
  else
    {
      k = i - lcount;
      if(k >= tt->ntopnodes || tt->fdepth[k] != depth)
	return;
      istop = 1;
      suns = A.t_suns + 8 * k;
      leafk = k;
      U = A.t_u[k];
    }
  double v[3][D], m[D];
#pragma unroll
  for(int g = 0; g < D; g++)
    v[0][g] = v[1][g] = v[2][g] = m[g] = 0.0;
  for(int s = 0; s < 8; s++)
    {
      unsigned int Uc = 0xffffffffu;
      if(istop && !tt->fisleaf[k])
	{
	  int ch = tt->fsuns[k][s];
	  if(ch < 0)
	    continue;
	  Uc = A.t_u[ch];
	}
      else
	{
	  int sv = suns[s];
	  if(sv == -1)
	    continue;
	  if(sv >= 0)
	    {
	      unsigned int idx = A.tq[sv];
	      const G2PRec r = A.prec[idx];
	      int g = S.t2g[r.type];
#pragma unroll
	      for(int gg = 0; gg < D; gg++)
		if(gg == g)
		  {
		    m[gg] = __dadd_rn(m[gg], (double) r.m);
		    for(int j = 0; j < 3; j++)
		      v[j][gg] = __dadd_rn(v[j][gg], (double) __fmul_rn(r.m, vel[3 * (size_t) idx + j]));
		  }
	      continue;
	    }
	  Uc = (unsigned int) (-(sv + 2)) + (unsigned int) tt->fdfs[leafk] + 1u;
	}
      const float4 *rec = A.wcells + (size_t) Uc * (2 + D);
#pragma unroll
      for(int g = 0; g < D; g++)
	{
	  float mc = rec[1 + g].w;
	  m[g] = __dadd_rn(m[g], (double) mc);
	  for(int j = 0; j < 3; j++)
	    v[j][g] = __dadd_rn(v[j][g], (double) __fmul_rn(mc, wvs[((size_t) Uc * 3 + j) * D + g]));
	}
    }
#pragma unroll
  for(int g = 0; g < D; g++)
    for(int j = 0; j < 3; j++)
      wvs[((size_t) U * 3 + j) * D + g] = m[g] > 0 ? (float) __ddiv_rn(v[j][g], m[g]) : (float) v[j][g];
}

static G2Soft make_soft(const g2gpu_ctx *c);
static BuildArrays make_arrays(g2gpu_ctx *c);

// Persistent staging of the on-demand exports (device buffer + pinned host mirror of it): one allocation, one D2H copy per call.
static int export_reserve(g2gpu_ctx *c, size_t bytes)
{
  if(c->export_bytes >= bytes)
    return 0;
  if(c->d_export)
    cudaFree(c->d_export);
  if(c->h_export)
    cudaFreeHost(c->h_export);
  c->d_export = c->h_export = nullptr;
  c->export_bytes = 0;
  bytes += bytes / 4;
  G2_CUDA(cudaMalloc((void **) &c->d_export, bytes));
  G2_CUDA(cudaMallocHost((void **) &c->h_export, bytes));
  c->export_bytes = bytes;
  return 0;
}

// ---------------------------------------------------------------- particle counts per species (NGRAVS_ACCUMULATOR) ----
// forcetree.c:557-559, 621-623: Nparticles[g] of a node = number of particles of species g below it; handed to the pair laws as
// N for particle-node interactions (forcetree.c:1563-1577).  A separate bottom-up pass over the same depth lists; wcnt[U][D].
template <int D>
__global__ void __launch_bounds__(128) cnt_level_kernel(BuildArrays A, const G2TopTree *__restrict__ tt, G2Soft S, unsigned int *__restrict__ wcnt,
							int depth, unsigned int lstart, int lcount, const unsigned int *__restrict__ depth_list)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int *suns;
  unsigned int U;
  int leafk, istop = 0, k = -1;
  if(i < lcount)
    {
      int c = (int) depth_list[lstart + i];
      suns = A.c_suns + 8 * (size_t) c;
      leafk = A.tl[A.c_a[c]];
      U = (unsigned int) c + (unsigned int) tt->fdfs[leafk] + 1u;
    }
  else
    {
      k = i - lcount;
      if(k >= tt->ntopnodes || tt->fdepth[k] != depth)
	return;
      istop = 1;
      suns = A.t_suns + 8 * k;
      leafk = k;
      U = A.t_u[k];
    }
  unsigned int cnt[D];
#pragma unroll
  for(int g = 0; g < D; g++)
    cnt[g] = 0;
  for(int s = 0; s < 8; s++)
    {
      unsigned int Uc;
      if(istop && !tt->fisleaf[k])
	{
	  int ch = tt->fsuns[k][s];
	  if(ch < 0)
	    continue;
	  Uc = A.t_u[ch];
	}
      else
	{
	  int sv = suns[s];
	  if(sv == -1)
	    continue;
	  if(sv >= 0)
	    {
	      int g = S.t2g[A.prec[A.tq[sv]].type];
#pragma unroll
	      for(int gg = 0; gg < D; gg++)
		cnt[gg] += (gg == g) ? 1u : 0u;
	      continue;
	    }
	  Uc = (unsigned int) (-(sv + 2)) + (unsigned int) tt->fdfs[leafk] + 1u;
	}
#pragma unroll
      for(int g = 0; g < D; g++)
	cnt[g] += wcnt[(size_t) Uc * D + g];
    }
#pragma unroll
  for(int g = 0; g < D; g++)
    wcnt[(size_t) U * D + g] = cnt[g];
}

// wcnt (depth-first index U) -> reference numbering, widened to the reference's `long`
__global__ void __launch_bounds__(128) cnt_permute_kernel(BuildArrays A, const G2TopTree *__restrict__ tt, const unsigned int *__restrict__ c_refid,
							  int ncells, int D, const unsigned int *__restrict__ wcnt, long long *__restrict__ out)
{
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int ntop = tt->ntopnodes;
  unsigned int U;
  int ref;
  if(tid < ncells)
    {
      U = (unsigned int) tid + (unsigned int) tt->fdfs[A.tl[A.c_a[tid]]] + 1u;
      ref = (int) c_refid[tid];
    }
  else if(tid < ncells + ntop)
    {
      int k = tid - ncells;
      U = A.t_u[k];
      ref = k;
    }
  else
    return;
  for(int g = 0; g < D; g++)
    out[(size_t) ref * D + g] = (long long) wcnt[(size_t) U * D + g];
}

int g2_stage_counts(g2gpu_ctx *c)
{
  if(c->stage < 3)
    return g2_fail(G2GPU_ERR_STATE, "tree has not been built");
  const int D = c->D, ntop = c->numnodes - c->ncells;
  cudaStream_t st = c->stream;
  if(!c->wcnt)
    G2_CUDA(cudaMalloc((void **) &c->wcnt, sizeof(unsigned int) * (size_t) D * (size_t) c->cfg.max_nodes));
  BuildArrays A = make_arrays(c);
  G2Soft S = make_soft(c);
  for(int d = c->maxdepth; d >= 0; d--)
    {
      const int lcount = c->depth_count[d];
      const int grid = g2_cdiv(lcount + ntop, 128);
      switch (D)
	{
	case 1: cnt_level_kernel<1><<<grid, 128, 0, st>>>(A, c->d_top, S, c->wcnt, d, c->depth_start[d], lcount, c->c_ready); break;
	case 2: cnt_level_kernel<2><<<grid, 128, 0, st>>>(A, c->d_top, S, c->wcnt, d, c->depth_start[d], lcount, c->c_ready); break;
	case 3: cnt_level_kernel<3><<<grid, 128, 0, st>>>(A, c->d_top, S, c->wcnt, d, c->depth_start[d], lcount, c->c_ready); break;
	case 4: cnt_level_kernel<4><<<grid, 128, 0, st>>>(A, c->d_top, S, c->wcnt, d, c->depth_start[d], lcount, c->c_ready); break;
	case 5: cnt_level_kernel<5><<<grid, 128, 0, st>>>(A, c->d_top, S, c->wcnt, d, c->depth_start[d], lcount, c->c_ready); break;
	default: cnt_level_kernel<6><<<grid, 128, 0, st>>>(A, c->d_top, S, c->wcnt, d, c->depth_start[d], lcount, c->c_ready); break;
	}
      c->launches++;
    }
  G2_CUDA(cudaGetLastError());
  c->counts_valid = 1;
  return 0;
}

int g2_export_nparticles(g2gpu_ctx *c, long long *out)
{
  if(c->stage < 3)
    return g2_fail(G2GPU_ERR_STATE, "tree has not been built");
  if(!c->counts_valid)
    G2_TRY(g2_stage_counts(c));
  G2_TRY(g2_stage_renumber(c));
  const int nn = c->numnodes, D = c->D;
  const size_t bytes = sizeof(long long) * (size_t) D * nn;
  G2_TRY(export_reserve(c, bytes));
  cudaStream_t st = c->stream;
  BuildArrays A = make_arrays(c);
  cnt_permute_kernel<<<g2_cdiv(nn, 128), 128, 0, st>>>(A, c->d_top, c->c_refid, c->ncells, D, c->wcnt, (long long *) c->d_export);
  c->launches++;
  cudaError_t e = cudaMemcpyAsync(c->h_export, c->d_export, bytes, cudaMemcpyDeviceToHost, st);
  if(e == cudaSuccess)
    e = cudaStreamSynchronize(st);
  if(e != cudaSuccess)
    return g2_fail(G2GPU_ERR_CUDA, "nparticles export: %s", cudaGetErrorString(e));
  memcpy(out, c->h_export, bytes);
  return 0;
}

int g2_export_extnodes(g2gpu_ctx *c, float *vs)
{
  if(c->stage < 3)
    return g2_fail(G2GPU_ERR_STATE, "tree has not been built");
  if(!c->have_vel)
    return g2_fail(G2GPU_ERR_STATE, "velocities were not uploaded (g2gpu_upload vel argument)");
  G2_TRY(g2_stage_renumber(c));
  const int nn = c->numnodes, D = c->D, ntop = nn - c->ncells;
  cudaStream_t st = c->stream;
  const size_t vbytes = sizeof(float) * 3 * (size_t) D * nn;
  G2_TRY(export_reserve(c, 2 * vbytes));
  float *wvs = (float *) (c->d_export + vbytes), *out = (float *) c->d_export;
  BuildArrays A = make_arrays(c);
  G2Soft S = make_soft(c);
  for(int d = c->maxdepth; d >= 0; d--)
    {
      const int lcount = c->depth_count[d];
      const int grid = g2_cdiv(lcount + ntop, 128);
      switch (D)
	{
	case 1: vs_level_kernel<1><<<grid, 128, 0, st>>>(A, c->d_top, S, c->vel, wvs, d, c->depth_start[d], lcount, c->c_ready); break;
	case 2: vs_level_kernel<2><<<grid, 128, 0, st>>>(A, c->d_top, S, c->vel, wvs, d, c->depth_start[d], lcount, c->c_ready); break;
	case 3: vs_level_kernel<3><<<grid, 128, 0, st>>>(A, c->d_top, S, c->vel, wvs, d, c->depth_start[d], lcount, c->c_ready); break;
	case 4: vs_level_kernel<4><<<grid, 128, 0, st>>>(A, c->d_top, S, c->vel, wvs, d, c->depth_start[d], lcount, c->c_ready); break;
	case 5: vs_level_kernel<5><<<grid, 128, 0, st>>>(A, c->d_top, S, c->vel, wvs, d, c->depth_start[d], lcount, c->c_ready); break;
	default: vs_level_kernel<6><<<grid, 128, 0, st>>>(A, c->d_top, S, c->vel, wvs, d, c->depth_start[d], lcount, c->c_ready); break;
	}
      c->launches++;
    }
  vs_permute_kernel<<<g2_cdiv(nn, 128), 128, 0, st>>>(A, c->d_top, c->c_refid, c->ncells, D, wvs, out);
  c->launches++;
  cudaError_t e = cudaMemcpyAsync(c->h_export, out, vbytes, cudaMemcpyDeviceToHost, st);
  if(e == cudaSuccess)
    e = cudaStreamSynchronize(st);
  if(e != cudaSuccess)
    return g2_fail(G2GPU_ERR_CUDA, "extnodes export: %s", cudaGetErrorString(e));
  memcpy(vs, c->h_export, vbytes);
  return 0;
}

// ---------------------------------------------------------------- 5. reference numbering -----------------------
// The reference creates the internal nodes of one insertion (particle i meeting particle j < i) as a chain of
// consecutive depths; these are exactly the cells whose second-smallest particle index is i (forcetree.c:183-247).
// Node number = NTopnodes + #cells with smaller second-min + (depth - shallowest depth of the chain).
__global__ void __launch_bounds__(256) renum_hist_kernel(const unsigned int *__restrict__ c_min2, const unsigned char *__restrict__ c_d,
							 const unsigned int *__restrict__ tbase, int n, unsigned int *__restrict__ hist,
							 unsigned int *__restrict__ dmin)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if(c >= (int) tbase[n])
    return;
  unsigned int i2 = c_min2[c];
  atomicAdd(&hist[i2], 1u);
  atomicMin(&dmin[i2], (unsigned int) c_d[c]);
}

__global__ void __launch_bounds__(256) renum_assign_kernel(const unsigned int *__restrict__ c_min2, const unsigned char *__restrict__ c_d,
							   const unsigned int *__restrict__ tbase, int n, const unsigned int *__restrict__ hscan,
							   const unsigned int *__restrict__ dmin, int ntop, unsigned int *__restrict__ c_refid)
{
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if(c >= (int) tbase[n])
    return;
  unsigned int i2 = c_min2[c];
  c_refid[c] = (unsigned int) ntop + hscan[i2] + ((unsigned int) c_d[c] - dmin[i2]);
}

__global__ void fill_u32_kernel(unsigned int *p, unsigned int v, size_t n)
{
  size_t i = (size_t) blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n)
    p[i] = v;
}

// ---------------------------------------------------------------- 6. export in the reference's form -------------
struct ExportArrays
{
  float *len, *center, *s, *mass;
  int *bitflags, *sibling, *nextnode, *father, *p_nextnode, *p_father;
};

// reference index (MaxPart-relative node number, or particle index) of the DFS successor of sorted position b's
// subtree end: the node visited after everything up to and including position b, given b ends a subtree whose
// enclosing cell has depth tm[b] (forcetree.c:499-512 `nextsib`).  Returns node numbers offset by maxpart.
__device__ int successor_after(const BuildArrays &A, const unsigned char *__restrict__ tm, const G2TopTree *__restrict__ tt,
			       const unsigned int *__restrict__ c_refid, const int *__restrict__ top_sibling, int maxpart, unsigned int b)
{
  unsigned char m = tm[b];
  if(m == 255)
    return top_sibling[A.tl[b]];	// last particle of its top-level leaf: continue after that leaf
  // the child of the depth-m cell that starts at b+1 has depth m+1
  unsigned char m2 = tm[b + 1];
  if(m2 != 255 && (int) m2 >= (int) m + 1)
    return maxpart + (int) c_refid[A.tbase[b + 1]];	// a cell: first cell owned by pair b+1
  return (int) A.tq[b + 1];	// a single particle
}

// reference-style sibling / nextnode of every top-level node (small, one thread)
__global__ void export_top_links_kernel(BuildArrays A, const unsigned char *__restrict__ tm, const G2TopTree *__restrict__ tt,
					const unsigned int *__restrict__ c_refid, int maxpart, int *__restrict__ top_sibling,
					int *__restrict__ top_nextnode)
{
  if(threadIdx.x != 0 || blockIdx.x != 0)
    return;
  const int ntop = tt->ntopnodes;
  // sibling: next existing slot under the father, else the father's sibling (root: -1); fdfs order guarantees
  // fathers come first
  for(int r = 0; r < ntop; r++)
    {
      int k = tt->fdfs_inv[r];
      int f = tt->ffather[k];
      int sib = -1;
      if(f >= 0)
	{
	  int myslot = (int) (tt->fmorton[k] & 7ull);
	  sib = top_sibling[f];
	  for(int s = myslot + 1; s < 8; s++)
	    if(tt->fsuns[f][s] >= 0)
	      {
		sib = maxpart + tt->fsuns[f][s];
		break;
	      }
	}
      top_sibling[k] = sib;
    }
  for(int k = 0; k < ntop; k++)
    {
      int nx;
      if(!tt->fisleaf[k])
	{
	  nx = -1;
	  for(int s = 0; s < 8; s++)
	    if(tt->fsuns[k][s] >= 0)
	      {
		nx = maxpart + tt->fsuns[k][s];
		break;
	      }
	}
      else
	{
	  unsigned int first = A.t_first[k], end = A.t_last[k];
	  if(end <= first)
	    nx = top_sibling[k];	// empty leaf: the next node visited is its successor
	  else
	    {
	      unsigned char m = tm[first];
	      if(m != 255 && (int) m >= tt->fdepth[k] + 1)
		nx = maxpart + (int) c_refid[A.tbase[first]];
	      else
		nx = (int) A.tq[first];
	    }
	}
      top_nextnode[k] = nx;
    }
}

template <int D>
__global__ void __launch_bounds__(128) export_kernel(BuildArrays A, const unsigned char *__restrict__ tm, const G2TopTree *__restrict__ tt,
						     const unsigned int *__restrict__ c_refid, const int *__restrict__ p_parent,
						     const int *__restrict__ top_sibling, const int *__restrict__ top_nextnode, int maxpart,
						     int unequal, ExportArrays E)
{
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int ncells = (int) A.tbase[A.n];
  const int ntop = tt->ntopnodes;
  if(tid < ncells + ntop)
    {
      int ref, sib, nxt, fat;
      unsigned int U;
      unsigned int extra = 0;
      if(tid < ncells)
	{
	  int c = tid;
	  ref = (int) c_refid[c];
	  unsigned int a = A.c_a[c], b = A.c_b[c];
	  int d = A.c_d[c];
	  U = (unsigned int) c + (unsigned int) tt->fdfs[A.tl[a]] + 1u;
	  sib = successor_after(A, tm, tt, c_refid, top_sibling, maxpart, b);
	  unsigned char m = tm[a];
	  if(m != 255 && (int) m >= d + 1)
	    nxt = maxpart + (int) c_refid[c + 1];	// first child is the next cell of the same chain
	  else
	    nxt = (int) A.tq[a];
	  int fc = A.c_father[c];
	  fat = fc >= 0 ? maxpart + (int) c_refid[fc] : maxpart + (-(fc + 2));
	}
      else
	{
	  int k = tid - ncells;
	  ref = k;
	  U = A.t_u[k];
	  sib = top_sibling[k];
	  nxt = top_nextnode[k];
	  fat = tt->ffather[k] >= 0 ? maxpart + tt->ffather[k] : -1;
	  extra = 3;		// force_flag_localnodes (forcetree.c:954-996): bits 0,1 on every top-level node
	}
      const float4 *rec = A.wcells + (size_t) U * (2 + D);
      float4 q0 = rec[0];
      if(E.len)
	E.len[ref] = q0.x;
      if(E.center)
	{
	  E.center[3 * (size_t) ref + 0] = q0.y;
	  E.center[3 * (size_t) ref + 1] = q0.z;
	  E.center[3 * (size_t) ref + 2] = q0.w;
	}
      for(int g = 0; g < D; g++)
	{
	  float4 q = rec[1 + g];
	  if(E.s)
	    {
	      E.s[(3 * (size_t) ref + 0) * D + g] = q.x;
	      E.s[(3 * (size_t) ref + 1) * D + g] = q.y;
	      E.s[(3 * (size_t) ref + 2) * D + g] = q.z;
	    }
	  if(E.mass)
	    E.mass[(size_t) ref * D + g] = q.w;
	}
      uint4 w = *((const uint4 *) &rec[1 + D]);
      if(E.bitflags)
	{
	  // an empty node carries the mixed bit only as a walk marker; the reference stores diffsoftflag = 0 for it
	  unsigned int mt = (w.z >> 29) & 7u, df = (mt == 7u) ? 0u : ((w.z >> 28) & 1u);
	  E.bitflags[ref] = unequal ? (int) (4u * mt + 32u * df + extra) : (int) extra;
	}
      if(E.sibling)
	E.sibling[ref] = sib;
      if(E.nextnode)
	E.nextnode[ref] = nxt;
      if(E.father)
	E.father[ref] = fat;
    }
  else if(tid < ncells + ntop + A.n)
    {
      unsigned int p = (unsigned int) (tid - ncells - ntop);
      unsigned int idx = A.tq[p];
      if(E.p_father)
	{
	  int pc = p_parent[p];
	  E.p_father[idx] = pc >= 0 ? maxpart + (int) c_refid[pc] : maxpart + (-(pc + 2));
	}
      if(E.p_nextnode)
	{
	  // a particle is a subtree of its own: successor at depth max(tm[p-1], tm[p]) ... use the right neighbour
	  // rule: the enclosing cell that ends at p has depth tm[p] (common levels with p+1)
	  E.p_nextnode[idx] = successor_after(A, tm, tt, c_refid, top_sibling, maxpart, p);
	}
    }
}

// ---------------------------------------------------------------- stage drivers -------------------------------
static G2Soft make_soft(const g2gpu_ctx *c)
{
  G2Soft S;
  for(int t = 0; t < 6; t++)
    {
      S.t2g[t] = c->type_to_grav[t];
      S.fsoft[t] = c->force_softening[t];
    }
  S.unequal = c->cfg.unequal_softenings;
  return S;
}

static BuildArrays make_arrays(g2gpu_ctx *c)
{
  BuildArrays A;
  A.tkey = c->tkey; A.tq = c->tq; A.tl = c->ttl; A.tbase = c->tbase; A.prec = c->prec;
  A.c_a = c->c_a; A.c_b = c->c_b; A.c_d = c->c_d; A.c_suns = c->c_suns; A.c_father = c->c_father;
  A.c_nchild = c->c_nchild; A.c_ready = c->c_ready; A.c_min1 = c->c_min1; A.c_min2 = c->c_min2;
  A.t_suns = c->t_suns; A.t_first = c->t_first; A.t_last = c->t_last; A.t_u = c->t_ubase; A.t_nchild = c->t_nchild;
  A.t_ready = c->t_ready; A.t_min1 = c->t_min1; A.t_min2 = c->t_min2; A.u_poff = c->c_poff;
  A.wcells = c->wcells; A.wpart = c->wpart; A.wsrc = c->wsrc; A.n = c->npart; A.numnodes_cap = c->cfg.max_nodes;
  return A;
}

int g2_stage_treebuild(g2gpu_ctx *c)
{
  if(c->stage < 2)
    return g2_fail(G2GPU_ERR_STATE, "treebuild: g2gpu_domain has not run");
  const int n = c->npart;
  const int maxcells = c->cfg.max_nodes;
  cudaStream_t st = c->stream;
  G2_CUDA(cudaEventRecord(c->ev[4], st));

  // 1. keys + sort (keys live in the sort ping-pong buffers)
  unsigned short *ptl = (unsigned short *) c->w_flags;	// per-particle top node, scratch reuse (n * 2 bytes)
  treekey_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(c->prec, c->phkey, n, c->d_top, c->skey[0], c->sval[0], ptl);
  c->launches++;
  unsigned long long *k = c->skey[0];
  unsigned int *v = c->sval[0];
  G2_TRY(g2_radix_sort_pairs(c, n, &k, &v, c->skey[1], c->sval[1], 0, KEYBITS));
  c->tkey = k;
  c->tq = v;

  // 2-3. cells
  pair_depth_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(c->tkey, c->tq, ptl, n, c->tm, c->ttl, c->d_err);
  cell_count_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(c->tm, c->ttl, c->d_top, n, c->tcnt);
  c->launches += 2;
  G2_TRY(g2_scan_exclusive_u32(c, c->tcnt, c->tbase, (size_t) n));
  G2_CUDA(cudaMemsetAsync(c->d_depth, 0, 64 * sizeof(unsigned int), st));
  cell_init_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(c->tm, c->ttl, c->d_top, c->tbase, n, maxcells, c->c_a, c->c_d, c->d_err, c->d_depth);
  c->launches++;

  // the cell count is needed on the host for grid sizes and the MaxNodes check (forcetree.c:249-255)
  G2_CUDA(cudaMemcpyAsync(&c->h_err[4], c->tbase + n, sizeof(int), cudaMemcpyDeviceToHost, st));
  G2_CUDA(cudaMemcpyAsync(&c->h_err[5], &c->d_top->ntopnodes, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  G2_CUDA(cudaMemcpyAsync(&c->h_err[0], c->d_err, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
  G2_CUDA(cudaMemcpyAsync(&c->h_err[16], c->d_depth, 32 * sizeof(int), cudaMemcpyDeviceToHost, st));
  G2_CUDA(cudaStreamSynchronize(st));
  if(c->h_err[2])
    return g2_fail(G2GPU_ERR_ARG, "a particle has a type outside 0..5");
  if(c->h_err[7])
    return g2_fail(G2GPU_ERR_TOPNODES, "top-level tree exceeds %d nodes", G2_MAXTOP);
  if(c->h_err[0])
    return g2_fail(G2GPU_ERR_TREE_DEPTH, "more than 8 particles share all %d octree levels (coincident particles?)", G2_MAXDEPTH);
  {				// largest |coordinate| of the domain cube (guard bands of the walk's FP32 decisions)
    double m = 0;
    for(int j = 0; j < 3; j++)
      {
	m = fmax(m, fabs(c->h_domain[j]));
	m = fmax(m, fabs(c->h_domain[j] + c->h_domain[6]));
      }
    c->coord_max = m;
  }
  c->ncells = c->h_err[4];
  const int ntop = c->h_err[5];
  c->numnodes = ntop + c->ncells;
  if(c->h_err[1] || c->numnodes >= c->cfg.max_nodes)
    return g2_fail(G2GPU_ERR_MAXNODES, "maximum number %d of tree-nodes reached (need %d)", c->cfg.max_nodes, c->numnodes);
  const int ncells = c->ncells;
  const int numnodes = c->numnodes;

  // 4. children, particle groups
  DepthStarts ds;
  int maxdepth = 0;
  {
    unsigned int run = 0;
    for(int d = 0; d < 32; d++)
      {
	ds.start[d] = run;
	c->depth_start[d] = run;
	c->depth_count[d] = c->h_err[16 + d];
	run += (unsigned int) c->h_err[16 + d];
	if(c->h_err[16 + d])
	  maxdepth = d;
      }
    if(c->h_err[8] > maxdepth)
      maxdepth = c->h_err[8];	// deepest top-level node
    c->maxdepth = maxdepth;
  }
  if(ncells > 0)
    {
      depth_scatter_kernel<<<g2_cdiv(ncells, 256), 256, 0, st>>>(c->c_d, ncells, ds, c->d_depth + 32, c->c_ready);
      c->launches++;
    }
  if(ncells > 0)
    topdown_kernel<<<g2_cdiv(ncells, 128), 128, 0, st>>>(c->tkey, c->tm, c->ttl, c->d_top, c->tbase, n, maxcells, c->c_a, c->c_d, c->c_b,
							  c->c_suns, c->c_father, c->p_parent, c->c_nchild, c->c_npart, c->tcnt);
  top_children_kernel<<<g2_cdiv(ntop, 128), 128, 0, st>>>(c->tkey, c->tm, c->ttl, c->d_top, c->tbase, n, c->t_suns, c->c_father, c->p_parent,
							   c->t_first, c->t_last, c->t_nchild, c->t_npart, c->t_ubase, c->tcnt);
  c->launches += 2;
  G2_TRY(g2_scan_exclusive_u32(c, c->tcnt, c->c_poff, (size_t) numnodes));

  // 5. moments
  BuildArrays A = make_arrays(c);
  G2Soft S = make_soft(c);
  for(int d = maxdepth; d >= 0; d--)
    {
      const int lcount = c->h_err[16 + d];
      const int nthreads = lcount + ntop;
      const int grid = g2_cdiv(nthreads, 128);
      const unsigned int *dl = c->c_ready;	// depth-grouped cell list
      switch (c->D)
	{
	case 1: level_kernel<1><<<grid, 128, 0, st>>>(A, c->d_top, S, d, ds.start[d], lcount, dl); break;
	case 2: level_kernel<2><<<grid, 128, 0, st>>>(A, c->d_top, S, d, ds.start[d], lcount, dl); break;
	case 3: level_kernel<3><<<grid, 128, 0, st>>>(A, c->d_top, S, d, ds.start[d], lcount, dl); break;
	case 4: level_kernel<4><<<grid, 128, 0, st>>>(A, c->d_top, S, d, ds.start[d], lcount, dl); break;
	case 5: level_kernel<5><<<grid, 128, 0, st>>>(A, c->d_top, S, d, ds.start[d], lcount, dl); break;
	case 6: level_kernel<6><<<grid, 128, 0, st>>>(A, c->d_top, S, d, ds.start[d], lcount, dl); break;
	default: return g2_fail(G2GPU_ERR_ARG, "unsupported N_GRAVS %d", c->D);
	}
      c->launches++;
    }
  c->stage = 3;
  c->tree_npart = n;
  c->tree_dynamic = 0;
  G2_CUDA(cudaEventRecord(c->ev[5], st));
  G2_CUDA(cudaGetLastError());
  c->renumbered = 0;
  c->counts_valid = 0;
  if(c->accumulator)
    G2_TRY(g2_stage_counts(c));
  return 0;
}

// ---------------------------------------------------------------- dynamic tree update ----------------------------
// Between two tree constructions the reference keeps its tree and lets it follow the particles: node centres of mass drift with the
// node velocities (predict.c:79-91), kicks are passed up the Father chain (timestep.c:329-344) and force_update_len (forcetree.c:1005-1122)
// enlarges nodes whose particles left them.  All of that runs on the host mirror; here the device tree of the LAST construction takes
// (a) the freshly uploaded particle records (same particles, same upload order) and (b) the host's len and s of every node.
__global__ void __launch_bounds__(256) dyn_particles_kernel(int n, const int *__restrict__ perm, const G2PRec *__restrict__ in_rec,
							     const float *__restrict__ in_gravpm, G2PRec *__restrict__ prec, float *__restrict__ gravpm)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n)
    return;
  const int src = perm[i];
  G2PRec r = in_rec[src];
  r.type = r.type < 0 ? 0 : (r.type > 5 ? 5 : r.type);
  r.active = r.active != 0;
  prec[i] = r;
  if(in_gravpm)
    for(int k = 0; k < 3; k++)
      gravpm[3 * (size_t) i + k] = in_gravpm[3 * (size_t) src + k];
}

__global__ void __launch_bounds__(256) dyn_wpart_kernel(int n, const unsigned int *__restrict__ wsrc, const G2PRec *__restrict__ prec,
							 float4 *__restrict__ wpart)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i < n)
    wpart[i] = *((const float4 *) &prec[wsrc[i]]);
}

template <int D>
__global__ void __launch_bounds__(256) dyn_nodes_kernel(BuildArrays A, const G2TopTree *__restrict__ tt, const unsigned int *__restrict__ c_refid,
							 int ncells, const float *__restrict__ len, const float *__restrict__ s)
{
  int tid = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned int U;
  int ref;
  if(tid < ncells)
    {
      U = (unsigned int) tid + (unsigned int) tt->fdfs[A.tl[A.c_a[tid]]] + 1u;
      ref = (int) c_refid[tid];
    }
  else if(tid - ncells < tt->ntopnodes)
    {
      ref = tid - ncells;
      U = A.t_u[ref];
    }
  else
    return;
  float4 *rec = A.wcells + (size_t) U * (2 + D);
  rec[0].x = len[ref];
#pragma unroll
  for(int g = 0; g < D; g++)
    {
      rec[1 + g].x = s[(3 * (size_t) ref + 0) * D + g];
      rec[1 + g].y = s[(3 * (size_t) ref + 1) * D + g];
      rec[1 + g].z = s[(3 * (size_t) ref + 2) * D + g];
    }
}

int g2_update_tree(g2gpu_ctx *c, const float *len, const float *s)
{
  if(!c->tree_npart)
    return g2_fail(G2GPU_ERR_STATE, "update_tree: no tree has been built");
  if(c->stage != 1 || c->npart != c->tree_npart)
    return g2_fail(G2GPU_ERR_STATE, "update_tree: upload the same %d particles first (stage %d, %d uploaded)", c->tree_npart, c->stage, c->npart);
  c->stage = 3;			// the arrays of the last construction are intact: an upload touches none of them
  G2_TRY(g2_stage_renumber(c));
  const int n = c->npart, nn = c->numnodes, D = c->D, ncells = c->ncells;
  cudaStream_t st = c->stream;
  const size_t lbytes = sizeof(float) * (size_t) nn, sbytes = sizeof(float) * 3 * (size_t) nn * D;
  G2_TRY(export_reserve(c, lbytes + sbytes));
  memcpy(c->h_export, len, lbytes);
  memcpy(c->h_export + lbytes, s, sbytes);
  G2_CUDA(cudaMemcpyAsync(c->d_export, c->h_export, lbytes + sbytes, cudaMemcpyHostToDevice, st));
  c->h2d_bytes += lbytes + sbytes;
  dyn_particles_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(n, c->perm, c->in_rec, c->have_gravpm ? c->in_gravpm : nullptr, c->prec, c->gravpm);
  dyn_wpart_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(n, c->wsrc, c->prec, c->wpart);
  BuildArrays A = make_arrays(c);
  const int total = nn;
  const float *dl = (const float *) c->d_export, *ds = (const float *) (c->d_export + lbytes);
  switch (D)
    {
#define G2_B(Dv) case Dv: dyn_nodes_kernel<Dv><<<g2_cdiv(total, 256), 256, 0, st>>>(A, c->d_top, c->c_refid, ncells, dl, ds); break
      G2_B(1); G2_B(2); G2_B(3); G2_B(4); G2_B(5); G2_B(6);
#undef G2_B
    default: return g2_fail(G2GPU_ERR_ARG, "unsupported N_GRAVS %d", D);
    }
  c->launches += 3;
  G2_CUDA(cudaGetLastError());
  G2_CUDA(cudaStreamSynchronize(st));	// h_export may be reused by the caller's next export
  c->tree_dynamic = 1;
  c->counts_valid = c->accumulator ? c->counts_valid : 0;
  return 0;
}

int g2_stage_renumber(g2gpu_ctx *c)
{
  if(c->stage < 3)
    return g2_fail(G2GPU_ERR_STATE, "tree has not been built");
  if(c->renumbered)
    return 0;
  const int n = c->npart, ncells = c->ncells;
  cudaStream_t st = c->stream;
  G2_CUDA(cudaMemsetAsync(c->hist2, 0, sizeof(unsigned int) * (size_t) (n + 1), st));
  fill_u32_kernel<<<g2_cdiv(n, 256), 256, 0, st>>>(c->dmin, 0xffffffffu, (size_t) n);
  c->launches++;
  if(ncells > 0)
    {
      renum_hist_kernel<<<g2_cdiv(ncells, 256), 256, 0, st>>>(c->c_min2, c->c_d, c->tbase, n, c->hist2, c->dmin);
      c->launches++;
    }
  G2_TRY(g2_scan_exclusive_u32(c, c->hist2, c->hist2_scan, (size_t) n));
  if(ncells > 0)
    {
      renum_assign_kernel<<<g2_cdiv(ncells, 256), 256, 0, st>>>(c->c_min2, c->c_d, c->tbase, n, c->hist2_scan, c->dmin,
								 c->numnodes - ncells, c->c_refid);
      c->launches++;
    }
  G2_CUDA(cudaGetLastError());
  c->renumbered = 1;
  return 0;
}

int g2_export_tree(g2gpu_ctx *c, float *len, float *center, float *s, float *mass, int *bitflags, int *sibling, int *nextnode,
		   int *father, int *p_nextnode, int *p_father)
{
  G2_TRY(g2_stage_renumber(c));
  const int n = c->npart, nn = c->numnodes, D = c->D, ntop = nn - c->ncells;
  cudaStream_t st = c->stream;
  // device staging
  size_t fbytes = sizeof(float) * (size_t) nn * (4 + 4 * (size_t) D);
  size_t ibytes = sizeof(int) * ((size_t) nn * 4 + (size_t) n * 2 + 2 * G2_MAXTOP);
  G2_TRY(export_reserve(c, fbytes + ibytes));
  char *buf = c->d_export;
  ExportArrays E;
  E.len = (float *) buf;
  E.center = E.len + nn;
  E.s = E.center + 3 * (size_t) nn;
  E.mass = E.s + 3 * (size_t) nn * D;
  E.bitflags = (int *) (buf + fbytes);
  E.sibling = E.bitflags + nn;
  E.nextnode = E.sibling + nn;
  E.father = E.nextnode + nn;
  E.p_nextnode = E.father + nn;
  E.p_father = E.p_nextnode + n;
  int *top_sibling = E.p_father + n;
  int *top_nextnode = top_sibling + G2_MAXTOP;
  BuildArrays A = make_arrays(c);
  export_top_links_kernel<<<1, 32, 0, st>>>(A, c->tm, c->d_top, c->c_refid, c->cfg.max_part, top_sibling, top_nextnode);
  int total = nn + n;
  int grid = g2_cdiv(total, 128);
  switch (D)
    {
    case 1: export_kernel<1><<<grid, 128, 0, st>>>(A, c->tm, c->d_top, c->c_refid, c->p_parent, top_sibling, top_nextnode, c->cfg.max_part, c->cfg.unequal_softenings, E); break;
    case 2: export_kernel<2><<<grid, 128, 0, st>>>(A, c->tm, c->d_top, c->c_refid, c->p_parent, top_sibling, top_nextnode, c->cfg.max_part, c->cfg.unequal_softenings, E); break;
    case 3: export_kernel<3><<<grid, 128, 0, st>>>(A, c->tm, c->d_top, c->c_refid, c->p_parent, top_sibling, top_nextnode, c->cfg.max_part, c->cfg.unequal_softenings, E); break;
    case 4: export_kernel<4><<<grid, 128, 0, st>>>(A, c->tm, c->d_top, c->c_refid, c->p_parent, top_sibling, top_nextnode, c->cfg.max_part, c->cfg.unequal_softenings, E); break;
    case 5: export_kernel<5><<<grid, 128, 0, st>>>(A, c->tm, c->d_top, c->c_refid, c->p_parent, top_sibling, top_nextnode, c->cfg.max_part, c->cfg.unequal_softenings, E); break;
    case 6: export_kernel<6><<<grid, 128, 0, st>>>(A, c->tm, c->d_top, c->c_refid, c->p_parent, top_sibling, top_nextnode, c->cfg.max_part, c->cfg.unequal_softenings, E); break;
    default: return g2_fail(G2GPU_ERR_ARG, "unsupported N_GRAVS %d", D);
    }
  c->launches += 2;
  (void) ntop;
  // one copy into the pinned mirror of the staging buffer, then plain memcpy into the caller's (pageable) arrays
  const size_t used = fbytes + sizeof(int) * ((size_t) nn * 4 + (size_t) n * 2);
  cudaError_t e = cudaMemcpyAsync(c->h_export, buf, used, cudaMemcpyDeviceToHost, st);
  if(e == cudaSuccess)
    e = cudaStreamSynchronize(st);
  if(e != cudaSuccess)
    return g2_fail(G2GPU_ERR_CUDA, "export: %s", cudaGetErrorString(e));
  const char *h = c->h_export;
#define G2_OUT(dst, src, count) if(dst) memcpy(dst, h + ((const char *) (src) - buf), sizeof(*(dst)) * (count))
  G2_OUT(len, E.len, (size_t) nn);
  G2_OUT(center, E.center, 3 * (size_t) nn);
  G2_OUT(s, E.s, 3 * (size_t) nn * D);
  G2_OUT(mass, E.mass, (size_t) nn * D);
  G2_OUT(bitflags, E.bitflags, (size_t) nn);
  G2_OUT(sibling, E.sibling, (size_t) nn);
  G2_OUT(nextnode, E.nextnode, (size_t) nn);
  G2_OUT(father, E.father, (size_t) nn);
  G2_OUT(p_nextnode, E.p_nextnode, (size_t) n);
  G2_OUT(p_father, E.p_father, (size_t) n);
#undef G2_OUT
  return 0;
}
