// g2_sort.cu — device-wide exclusive scan and stable LSD radix sort of (u64 key, u32 value) pairs.
//
// Replaces the reference's qsort() calls on the hot path: the Peano-Hilbert key sort of
// domain_determineTopTree (domain.c:946) and the two-level (species, key) sort of peano_hilbert_order
// (peano.c:100,124), and orders particles along the tree for the build.  HBM-bound integer work:
// per pass  tile histogram (8 B/key read)  ->  scan of the bin-major tile histogram  ->  stable scatter
// (12 B read + 12 B written per pair).  Digits are RADIX_BITS = 9 bits (512 bins) so that a 57..63-bit key
// needs 7 passes.
#include "g2_common.cuh"

#define RADIX_BITS 9
#define RADIX_BINS (1 << RADIX_BITS)
#define SORT_THREADS 256
#define SORT_ITEMS 16
#define SORT_TILE (SORT_THREADS * SORT_ITEMS)	// 4096 pairs per CTA
#define SORT_WARPS (SORT_THREADS / 32)

// ------------------------------------------------------------------ exclusive scan (u32) -----------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
#define SCAN_TILE (SCAN_THREADS * SCAN_ITEMS)

__device__ __forceinline__ unsigned int warp_incl_scan(unsigned int v)
{
#pragma unroll
  for(int o = 1; o < 32; o <<= 1)
    {
      unsigned int t = __shfl_up_sync(0xffffffffu, v, o);
      if((threadIdx.x & 31) >= o)
	v += t;
    }
  return v;
}

// block-wide exclusive scan of one value per thread; returns exclusive prefix, *total = block sum
__device__ __forceinline__ unsigned int block_excl_scan(unsigned int v, unsigned int *total, unsigned int *smem /* >= 33 */)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  unsigned int inc = warp_incl_scan(v);
  if(lane == 31)
    smem[warp] = inc;
  __syncthreads();
  if(warp == 0)
    {
      unsigned int w = (lane < nw) ? smem[lane] : 0;
      unsigned int winc = warp_incl_scan(w);
      smem[lane] = winc - w;
      if(lane == 31)
	smem[32] = winc;
    }
  __syncthreads();
  unsigned int res = inc - v + smem[warp];
  *total = smem[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_reduce_kernel(const unsigned int *__restrict__ in, unsigned int *__restrict__ bsum, size_t n)
{
  __shared__ unsigned int sm[33];
  size_t base = (size_t) blockIdx.x * SCAN_TILE;
  unsigned int s = 0;
#pragma unroll
  for(int i = 0; i < SCAN_ITEMS; i++)
    {
      size_t idx = base + (size_t) i * SCAN_THREADS + threadIdx.x;
      if(idx < n)
	s += in[idx];
    }
  unsigned int tot;
  block_excl_scan(s, &tot, sm);
  if(threadIdx.x == 0)
    bsum[blockIdx.x] = tot;
}

// single CTA: exclusive scan of nb block sums in place (nb arbitrary), writes grand total to bsum[nb]
__global__ void __launch_bounds__(1024) scan_blocksums_kernel(unsigned int *bsum, int nb)
{
  __shared__ unsigned int sm[33];
  __shared__ unsigned int carry;
  if(threadIdx.x == 0)
    carry = 0;
  __syncthreads();
  for(int start = 0; start < nb; start += 1024)
    {
      int i = start + threadIdx.x;
      unsigned int v = (i < nb) ? bsum[i] : 0;
      unsigned int tot;
      unsigned int ex = block_excl_scan(v, &tot, sm);
      unsigned int c = carry;
      if(i < nb)
	bsum[i] = ex + c;
      __syncthreads();
      if(threadIdx.x == 0)
	carry = c + tot;
      __syncthreads();
    }
  if(threadIdx.x == 0)
    bsum[nb] = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const unsigned int *__restrict__ in, unsigned int *__restrict__ out,
								   const unsigned int *__restrict__ bsum, size_t n, int nb)
{
  __shared__ unsigned int sm[33];
  size_t base = (size_t) blockIdx.x * SCAN_TILE + (size_t) threadIdx.x * SCAN_ITEMS;	// blocked arrangement
  unsigned int v[SCAN_ITEMS];
  unsigned int s = 0;
#pragma unroll
  for(int i = 0; i < SCAN_ITEMS; i++)
    {
      size_t idx = base + i;
      v[i] = (idx < n) ? in[idx] : 0;
      s += v[i];
    }
  unsigned int tot;
  unsigned int ex = block_excl_scan(s, &tot, sm) + bsum[blockIdx.x];
#pragma unroll
  for(int i = 0; i < SCAN_ITEMS; i++)
    {
      size_t idx = base + i;
      if(idx < n)
	out[idx] = ex;
      ex += v[i];
    }
  // out[n] = grand total (callers allocate n+1)
  if(blockIdx.x == nb - 1 && threadIdx.x == 0)
    out[n] = bsum[nb];
}

// out has n+1 entries: out[i] = sum_{j<i} in[j], out[n] = total.  in == out is allowed.
int g2_scan_exclusive_u32(g2gpu_ctx *c, const unsigned int *in, unsigned int *out, size_t n)
{
  if(n == 0)
    {
      G2_CUDA(cudaMemsetAsync(out, 0, sizeof(unsigned int), c->stream));
      return 0;
    }
  int nb = g2_cdiv((long long) n, SCAN_TILE);
  if((size_t) nb + 1 > c->scan_tmp_elems)
    return g2_fail(G2GPU_ERR_ARG, "scan: %zu elements exceed scratch", n);
  scan_reduce_kernel<<<nb, SCAN_THREADS, 0, c->stream>>>(in, c->scan_tmp, n);
  scan_blocksums_kernel<<<1, 1024, 0, c->stream>>>(c->scan_tmp, nb);
  scan_apply_kernel<<<nb, SCAN_THREADS, 0, c->stream>>>(in, out, c->scan_tmp, n, nb);
  c->launches += 3;
  G2_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------ radix sort ----------------------------
__device__ __forceinline__ unsigned int digit_of(unsigned long long k, int shift, unsigned int mask)
{
  return (unsigned int) (k >> shift) & mask;
}

// per-tile digit histogram, stored bin-major: tilehist[bin * ntiles + tile]
__global__ void __launch_bounds__(SORT_THREADS) sort_hist_kernel(const unsigned long long *__restrict__ keys, unsigned int *__restrict__ tilehist,
								  int n, int ntiles, int shift, unsigned int mask)
{
  __shared__ unsigned int h[RADIX_BINS];
  for(int i = threadIdx.x; i < RADIX_BINS; i += SORT_THREADS)
    h[i] = 0;
  __syncthreads();
  const int tile = blockIdx.x;
  const size_t base = (size_t) tile * SORT_TILE;
#pragma unroll 4
  for(int i = 0; i < SORT_ITEMS; i++)
    {
      size_t idx = base + (size_t) i * SORT_THREADS + threadIdx.x;
      if(idx < (size_t) n)
	atomicAdd(&h[digit_of(keys[idx], shift, mask)], 1u);
    }
  __syncthreads();
  for(int i = threadIdx.x; i < RADIX_BINS; i += SORT_THREADS)
    tilehist[(size_t) i * ntiles + tile] = h[i];
}

// stable scatter.  Element order inside a tile: warp w owns [w*512, (w+1)*512), iteration k takes 32 consecutive
// elements, so (warp, k, lane) is the input order.  The tile is first sorted by digit in shared memory, then written
// out in tile-sorted order: consecutive threads write consecutive addresses of one bin's run (coalesced), instead of
// 4096 scattered 8-byte stores.
#define SCATTER_SMEM (SORT_WARPS * RADIX_BINS * 4 + 2 * RADIX_BINS * 4 + 64 * 4 + SORT_TILE * 8 + SORT_TILE * 4)
__global__ void __launch_bounds__(SORT_THREADS, 3) sort_scatter_kernel(const unsigned long long *__restrict__ keys_in, const unsigned int *__restrict__ vals_in,
								     unsigned long long *__restrict__ keys_out, unsigned int *__restrict__ vals_out,
								     const unsigned int *__restrict__ tilescan, int n, int ntiles, int shift, unsigned int mask,
									     unsigned int *__restrict__ capture_dest)
{
  extern __shared__ unsigned long long sort_dyn[];
  unsigned long long *stage_k = sort_dyn;				// SORT_TILE keys
  unsigned int *stage_v = (unsigned int *) (stage_k + SORT_TILE);	// SORT_TILE values
  unsigned int (*wcnt)[RADIX_BINS] = (unsigned int (*)[RADIX_BINS]) (stage_v + SORT_TILE);	// per-warp digit counts -> warp offsets
  unsigned int *binstart = &wcnt[0][0] + SORT_WARPS * RADIX_BINS;	// tile-local exclusive scan of the tile histogram
  unsigned int *gdelta = binstart + RADIX_BINS;				// global base of (bin, tile) minus binstart
  unsigned int *scratch = gdelta + RADIX_BINS;				// 33 words for the block scan
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int tile = blockIdx.x;
  for(int i = threadIdx.x; i < SORT_WARPS * RADIX_BINS; i += SORT_THREADS)
    (&wcnt[0][0])[i] = 0;
  __syncthreads();

  const size_t tbase = (size_t) tile * SORT_TILE;
  const int tcount = (size_t) n - tbase < (size_t) SORT_TILE ? (int) ((size_t) n - tbase) : SORT_TILE;
  const size_t wbase = tbase + (size_t) warp * (32 * SORT_ITEMS);
  unsigned long long k[SORT_ITEMS];
  unsigned int r[SORT_ITEMS];
#pragma unroll
  for(int i = 0; i < SORT_ITEMS; i++)
    {
      size_t idx = wbase + (size_t) i * 32 + lane;
      k[i] = (idx < (size_t) n) ? keys_in[idx] : ~0ull;
    }
#pragma unroll
  for(int i = 0; i < SORT_ITEMS; i++)
    {
      size_t idx = wbase + (size_t) i * 32 + lane;
      bool valid = idx < (size_t) n;
      unsigned int d = valid ? digit_of(k[i], shift, mask) : RADIX_BINS;	// invalid lanes match only each other
      unsigned int peers = __match_any_sync(0xffffffffu, d);
      unsigned int below = __popc(peers & ((1u << lane) - 1u));
      unsigned int prev = 0;
      int leader = __ffs(peers) - 1;
      if(valid)
	prev = wcnt[warp][d];	// all peers read the same counter before the leader updates it
      __syncwarp();
      if(valid && lane == leader)
	wcnt[warp][d] = prev + __popc(peers);
      __syncwarp();
      r[i] = prev + below;
    }
  __syncthreads();
  // per bin: exclusive scan over warps (-> warp offsets) and the bin total; two consecutive bins per thread
  unsigned int tot2[RADIX_BINS / SORT_THREADS];
  unsigned int mysum = 0;
#pragma unroll
  for(int q = 0; q < RADIX_BINS / SORT_THREADS; q++)
    {
      int b = threadIdx.x * (RADIX_BINS / SORT_THREADS) + q;
      unsigned int run = 0;
#pragma unroll
      for(int w = 0; w < SORT_WARPS; w++)
	{
	  unsigned int t = wcnt[w][b];
	  wcnt[w][b] = run;
	  run += t;
	}
      tot2[q] = run;
      mysum += run;
    }
  unsigned int blocktotal;
  unsigned int ex = block_excl_scan(mysum, &blocktotal, scratch);
#pragma unroll
  for(int q = 0; q < RADIX_BINS / SORT_THREADS; q++)
    {
      int b = threadIdx.x * (RADIX_BINS / SORT_THREADS) + q;
      binstart[b] = ex;
      gdelta[b] = tilescan[(size_t) b * ntiles + tile] - ex;
      ex += tot2[q];
    }
  __syncthreads();
  // tile-sorted staging
#pragma unroll
  for(int i = 0; i < SORT_ITEMS; i++)
    {
      size_t idx = wbase + (size_t) i * 32 + lane;
      if(idx < (size_t) n)
	{
	  unsigned int d = digit_of(k[i], shift, mask);
	  unsigned int pos = binstart[d] + wcnt[warp][d] + r[i];
	  stage_k[pos] = k[i];
	  stage_v[pos] = vals_in[idx];
	  if(capture_dest)		// where the element at input position idx of THIS pass ends up (coalesced: consecutive lanes, consecutive idx)
	    capture_dest[idx] = gdelta[d] + pos;
	}
    }
  __syncthreads();
  for(int j = threadIdx.x; j < tcount; j += SORT_THREADS)
    {
      unsigned long long key = stage_k[j];
      unsigned int d = digit_of(key, shift, mask);
      unsigned int out = gdelta[d] + (unsigned int) j;
      keys_out[out] = key;
      vals_out[out] = stage_v[j];
    }
}

// ------------------------------------------------------------------ one-sweep radix sort ------------------
// One kernel per digit (9 bits): a tile loads its 4096 pairs ONCE, ranks them, learns where its run of every bin starts from the tiles
// before it by a decoupled look-back over per-(tile, bin) status words, and scatters -- 12 B read + 12 B written per pair and pass,
// instead of histogram + three scan kernels + scatter (32 B per pair and pass, five launches).  The digit histograms of ALL passes are
// counted by one kernel up front (they do not depend on the order of the pairs).  Tiles take their number from an atomic ticket, so
// every tile a look-back waits for is already running.
// lanes of the warp that hold the same digit (d <= OS_BINS: OS_BITS + 1 bits).  BALLOT: one vote per digit bit instead of MATCH.ANY.
template <bool BALLOT>
__device__ __forceinline__ unsigned int os_peers(unsigned int d)
{
  if(!BALLOT)
    return __match_any_sync(0xffffffffu, d);
  unsigned int peers = 0xffffffffu;
#pragma unroll
  for(int b = 0; b <= RADIX_BITS; b++)
    {
      const bool bit = (d >> b) & 1u;
      const unsigned int bal = __ballot_sync(0xffffffffu, bit);
      peers &= bit ? bal : ~bal;
    }
  return peers;
}

#define OS_BITS RADIX_BITS		// 9 bits: 7 passes for the 56-bit (block, Peano-Hilbert) keys and for the 63-bit tree keys
#define OS_BINS (1 << OS_BITS)
#define OS_MAXPASS 8
#define OS_FLAG_AGG 0x40000000u		// status word: count of this tile alone
#define OS_FLAG_INC 0x80000000u		// status word: count of this tile and all tiles before it
#define OS_VALUE_MASK 0x3fffffffu

// hist[p * OS_BINS + d] += number of keys whose p-th digit is d (all passes at once)
__global__ void __launch_bounds__(256) os_hist_kernel(const unsigned long long *__restrict__ keys, int n, int begin_bit, int nbits, int npass,
						      unsigned int *__restrict__ hist)
{
  __shared__ unsigned int h[OS_MAXPASS * OS_BINS];
  for(int i = threadIdx.x; i < npass * OS_BINS; i += 256)
    h[i] = 0;
  __syncthreads();
  const size_t stride = (size_t) gridDim.x * 256;
  for(size_t i = (size_t) blockIdx.x * 256 + threadIdx.x; i < (size_t) n; i += stride)
    {
      const unsigned long long k = keys[i] >> begin_bit;
      for(int p = 0; p < npass; p++)
	{
	  const int bits = nbits - p * OS_BITS < OS_BITS ? nbits - p * OS_BITS : OS_BITS;	// the last pass may be narrower
	  atomicAdd(&h[p * OS_BINS + (unsigned int) ((k >> (p * OS_BITS)) & ((1u << bits) - 1u))], 1u);
	}
    }
  __syncthreads();
  for(int i = threadIdx.x; i < npass * OS_BINS; i += 256)
    if(h[i])
      atomicAdd(&hist[i], h[i]);
}

// exclusive scan of every pass's bins, in place (one thread per bin); also resets the tile tickets
__global__ void __launch_bounds__(OS_BINS) os_scan_kernel(unsigned int *hist, int npass, unsigned int *tickets)
{
  __shared__ unsigned int sm[33];
  for(int p = 0; p < npass; p++)
    {
      unsigned int v = hist[p * OS_BINS + threadIdx.x], tot;
      unsigned int ex = block_excl_scan(v, &tot, sm);
      hist[p * OS_BINS + threadIdx.x] = ex;
    }
  if(threadIdx.x < OS_MAXPASS)
    tickets[threadIdx.x] = 0;
}

#define OS_SMEM(ITEMS) (SORT_WARPS * OS_BINS * 4 + 2 * OS_BINS * 4 + 64 * 4 + (SORT_THREADS * (ITEMS)) * 8 + (SORT_THREADS * (ITEMS)) * 4)
// ITEMS pairs per thread: 8 (2048-pair tiles, 4 CTAs per SM, values loaded together with the keys; default) or 16 (4096-pair tiles, 3 CTAs per SM;
// G2GPU_SORT_ITEMS=16).  The pass is latency bound (ncu, profiles/r2_ospass_p256.txt: 15 % issue slots, 17 long-scoreboard stalls per issue, DRAM 14 %).
// MEASURED (B200, 16.8 M pairs, 7 passes of the domain sort; profiles/r2_sort_variants.txt): serial look-back 2.25 ms (16 items) / 2.19 ms (8 items);
// windowed look-back 2.08 / 1.90 ms; a window of 8 tiles instead of 4 and ranking by warp votes instead of MATCH.ANY change nothing.
template <int ITEMS, int BLOCKS, bool BALLOT, int WINDOW>
__global__ void __launch_bounds__(SORT_THREADS, BLOCKS) os_pass_kernel(const unsigned long long *__restrict__ keys_in, const unsigned int *__restrict__ vals_in,
								    unsigned long long *__restrict__ keys_out, unsigned int *__restrict__ vals_out,
								    const unsigned int *__restrict__ binbase /* this pass's scanned histogram */,
								    volatile unsigned int *__restrict__ status /* [tile][bin] of this pass, zeroed */,
								    unsigned int *__restrict__ ticket, int n, int shift, unsigned int mask,
								    unsigned int *__restrict__ capture_dest)
{
  constexpr int TILE = SORT_THREADS * ITEMS;
  extern __shared__ unsigned long long sort_dyn[];
  unsigned long long *stage_k = sort_dyn;				// TILE keys
  unsigned int *stage_v = (unsigned int *) (stage_k + TILE);	// TILE values
  unsigned int (*wcnt)[OS_BINS] = (unsigned int (*)[OS_BINS]) (stage_v + TILE);	// per-warp digit counts -> warp offsets
  unsigned int *binstart = &wcnt[0][0] + SORT_WARPS * OS_BINS;	// tile-local exclusive scan of the tile histogram
  unsigned int *gdelta = binstart + OS_BINS;				// global start of the tile's run of a bin, minus binstart
  unsigned int *scratch = gdelta + OS_BINS;				// 33 words for the block scan + the ticket
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if(threadIdx.x == 0)
    scratch[40] = atomicAdd(ticket, 1u);	// (scratch[0..32] serve the block scan)
  for(int i = threadIdx.x; i < SORT_WARPS * OS_BINS; i += SORT_THREADS)
    (&wcnt[0][0])[i] = 0;
  __syncthreads();
  const int tile = (int) scratch[40];

  const size_t tbase = (size_t) tile * TILE;
  const int tcount = (size_t) n - tbase < (size_t) TILE ? (int) ((size_t) n - tbase) : TILE;
  const size_t wbase = tbase + (size_t) warp * (32 * ITEMS);
  unsigned long long k[ITEMS];
  unsigned int r[ITEMS], v[ITEMS <= 8 ? ITEMS : 1];
#pragma unroll
  for(int i = 0; i < ITEMS; i++)
    {
      size_t idx = wbase + (size_t) i * 32 + lane;
      k[i] = (idx < (size_t) n) ? keys_in[idx] : ~0ull;
      if(ITEMS <= 8)
	v[i] = (idx < (size_t) n) ? vals_in[idx] : 0u;
    }
#pragma unroll
  for(int i = 0; i < ITEMS; i++)
    {
      size_t idx = wbase + (size_t) i * 32 + lane;
      bool valid = idx < (size_t) n;
      unsigned int d = valid ? digit_of(k[i], shift, mask) : OS_BINS;	// invalid lanes match only each other
      unsigned int peers = os_peers<BALLOT>(d);
      unsigned int below = __popc(peers & ((1u << lane) - 1u));
      unsigned int prev = 0;
      int leader = __ffs(peers) - 1;
      if(valid)
	prev = wcnt[warp][d];	// all peers read the same counter before the leader updates it
      __syncwarp();
      if(valid && lane == leader)
	wcnt[warp][d] = prev + __popc(peers);
      __syncwarp();
      r[i] = prev + below;
    }
  __syncthreads();
  // two consecutive bins per thread: exclusive scan over warps (-> warp offsets), the tile's count, its start inside the tile
  unsigned int run2[OS_BINS / SORT_THREADS], before2[OS_BINS / SORT_THREADS];
  unsigned int mysum = 0;
#pragma unroll
  for(int q = 0; q < OS_BINS / SORT_THREADS; q++)
    {
      const int b = threadIdx.x * (OS_BINS / SORT_THREADS) + q;
      unsigned int run = 0;
#pragma unroll
      for(int w = 0; w < SORT_WARPS; w++)
	{
	  unsigned int t = wcnt[w][b];
	  wcnt[w][b] = run;
	  run += t;
	}
      run2[q] = run;
      mysum += run;
      // publish this tile's count of the bin
      status[(size_t) tile * OS_BINS + b] = run | (tile == 0 ? OS_FLAG_INC : OS_FLAG_AGG);
    }
  // add up the tiles before this one (decoupled look-back), then publish the inclusive count.  A thread owns two adjacent bins, i.e. one
  // 8-byte word of every tile's status row; it fetches the words of WINDOW predecessor tiles at once (independent loads: the chain of L2
  // round trips, which is what bounds the pass, becomes WINDOW times shorter) and consumes them in order up to the first one that is not
  // published yet.
  static_assert(OS_BINS / SORT_THREADS == 2, "two bins per thread");
  {
    const int b0 = threadIdx.x * 2;
    unsigned int before0 = 0, before1 = 0;
    if(tile > 0)
      {
	bool done0 = false, done1 = false;
	int t = tile - 1;
	while(!(done0 && done1))
	  {
	    unsigned long long sw[WINDOW];
#pragma unroll
	    for(int w = 0; w < WINDOW; w++)
	      sw[w] = t - w >= 0 ? *(const volatile unsigned long long *) (status + (size_t) (t - w) * OS_BINS + b0)
		: ((unsigned long long) OS_FLAG_INC << 32 | OS_FLAG_INC);	// (before tile 0: nothing)
	    int used = 0;
#pragma unroll
	    for(int w = 0; w < WINDOW; w++)
	      {
		const unsigned int lo = (unsigned int) sw[w], hi = (unsigned int) (sw[w] >> 32);
		const bool ready = (done0 || (lo & (OS_FLAG_INC | OS_FLAG_AGG))) && (done1 || (hi & (OS_FLAG_INC | OS_FLAG_AGG)));
		if(used == w && ready && !(done0 && done1))
		  {
		    if(!done0)
		      {
			before0 += lo & OS_VALUE_MASK;
			done0 = (lo & OS_FLAG_INC) != 0u;
		      }
		    if(!done1)
		      {
			before1 += hi & OS_VALUE_MASK;
			done1 = (hi & OS_FLAG_INC) != 0u;
		      }
		    used = w + 1;
		  }
	      }
	    t -= used;		// (used == 0: the nearest word is not published yet -- that tile holds a smaller ticket, so it is running: look again)
	  }
	status[(size_t) tile * OS_BINS + b0] = (before0 + run2[0]) | OS_FLAG_INC;
	status[(size_t) tile * OS_BINS + b0 + 1] = (before1 + run2[1]) | OS_FLAG_INC;
      }
    before2[0] = before0;
    before2[1] = before1;
  }
  unsigned int blocktotal;
  unsigned int ex = block_excl_scan(mysum, &blocktotal, scratch);
#pragma unroll
  for(int q = 0; q < OS_BINS / SORT_THREADS; q++)
    {
      const int b = threadIdx.x * (OS_BINS / SORT_THREADS) + q;
      binstart[b] = ex;
      gdelta[b] = binbase[b] + before2[q] - ex;
      ex += run2[q];
    }
  __syncthreads();
  // tile-sorted staging
#pragma unroll
  for(int i = 0; i < ITEMS; i++)
    {
      size_t idx = wbase + (size_t) i * 32 + lane;
      if(idx < (size_t) n)
	{
	  unsigned int d = digit_of(k[i], shift, mask);
	  unsigned int pos = binstart[d] + wcnt[warp][d] + r[i];
	  stage_k[pos] = k[i];
	  stage_v[pos] = ITEMS <= 8 ? v[ITEMS <= 8 ? i : 0] : vals_in[idx];
	  if(capture_dest)		// where the element at input position idx of THIS pass ends up (coalesced: consecutive lanes, consecutive idx)
	    capture_dest[idx] = gdelta[d] + pos;
	}
    }
  __syncthreads();
  for(int j = threadIdx.x; j < tcount; j += SORT_THREADS)
    {
      unsigned long long key = stage_k[j];
      unsigned int d = digit_of(key, shift, mask);
      unsigned int out = gdelta[d] + (unsigned int) j;
      keys_out[out] = key;
      vals_out[out] = stage_v[j];
    }
}

static int g2_onesweep_sort_pairs(g2gpu_ctx *c, int n, unsigned long long **keys_io, unsigned int **vals_io, unsigned long long *keys_alt,
				  unsigned int *vals_alt, int begin_bit, int end_bit, int capture_shift, unsigned int *capture_dest)
{
  const int items = c->sort_items == 8 ? 8 : 16;
  const int ntiles = g2_cdiv(n, SORT_THREADS * items);
  const int nbits = end_bit - begin_bit;
  const int npass = (nbits + OS_BITS - 1) / OS_BITS;
  if(npass > OS_MAXPASS)
    return g2_fail(G2GPU_ERR_ARG, "sort: %d key bits need more than %d passes", nbits, OS_MAXPASS);
  // scratch: [npass * 256] histograms | [8] tickets | [npass][ntiles][256] status words
  const size_t need = (size_t) OS_MAXPASS * OS_BINS + 8 + (size_t) npass * ntiles * OS_BINS;
  if(need > c->tilehist_elems)
    return g2_fail(G2GPU_ERR_ARG, "sort: %d pairs exceed scratch", n);
  unsigned int *hist = c->tilehist, *tickets = hist + OS_MAXPASS * OS_BINS, *status = tickets + 8;
  cudaStream_t st = c->stream;
  G2_CUDA(cudaMemsetAsync(c->tilehist, 0, sizeof(unsigned int) * need, st));
  int hb = c->nsm * 8;
  if(hb > g2_cdiv(n, 256))
    hb = g2_cdiv(n, 256);
  os_hist_kernel<<<hb, 256, 0, st>>>(*keys_io, n, begin_bit, nbits, npass, hist);
  os_scan_kernel<<<1, OS_BINS, 0, st>>>(hist, npass, tickets);
  c->launches += 2;
  unsigned long long *kin = *keys_io, *kout = keys_alt;
  unsigned int *vin = *vals_io, *vout = vals_alt;
  for(int p = 0; p < npass; p++)
    {
      const int shift = begin_bit + p * OS_BITS;
      const int bits = end_bit - shift < OS_BITS ? end_bit - shift : OS_BITS;
      const unsigned int mask = (1u << bits) - 1u;
#define G2_OS_LAUNCH(IT, BL, BA, WI) do { \
	  G2_CUDA(cudaFuncSetAttribute(os_pass_kernel<IT, BL, BA, WI>, cudaFuncAttributeMaxDynamicSharedMemorySize, OS_SMEM(IT))); \
	  os_pass_kernel<IT, BL, BA, WI><<<ntiles, SORT_THREADS, OS_SMEM(IT), st>>>(kin, vin, kout, vout, hist + p * OS_BINS, status + (size_t) p * ntiles * OS_BINS, \
	      tickets + p, n, shift, mask, shift == capture_shift ? capture_dest : nullptr); } while(0)
      const bool w8 = c->sort_window == 8;
      if(items == 8)
	{
	  if(w8) G2_OS_LAUNCH(8, 4, false, 8); else G2_OS_LAUNCH(8, 4, false, 4);
	}
      else if(c->sort_rank_ballot)
	G2_OS_LAUNCH(16, 3, true, 4);
      else
	{
	  if(w8) G2_OS_LAUNCH(16, 3, false, 8); else G2_OS_LAUNCH(16, 3, false, 4);
	}
#undef G2_OS_LAUNCH
      c->launches++;
      unsigned long long *tk = kin; kin = kout; kout = tk;
      unsigned int *tv = vin; vin = vout; vout = tv;
    }
  G2_CUDA(cudaGetLastError());
  *keys_io = kin;
  *vals_io = vin;
  return 0;
}

// Sorts n pairs by key bits [begin_bit, end_bit), stable.  *keys_io/*vals_io hold the input and are
// updated to point at the buffers holding the result (ping-pong with keys_alt/vals_alt).
// capture_shift >= 0: the pass that starts at that bit also writes capture_dest[i] = output position of its i-th input element.  An LSD
// sort is sorted by the bits below `shift` when that pass starts, so with the pass being the last one capture_dest lists, in the order
// of the low bits alone, where every element finally went (g2_domain.cu: Peano-Hilbert order of all species -> species-major order).
int g2_radix_sort_pairs(g2gpu_ctx *c, int n, unsigned long long **keys_io, unsigned int **vals_io,
			unsigned long long *keys_alt, unsigned int *vals_alt, int begin_bit, int end_bit, int capture_shift, unsigned int *capture_dest)
{
  if(n <= 1 || end_bit <= begin_bit)
    return 0;
  if(c->sort_onesweep && (capture_shift < 0 || (capture_shift - begin_bit) % OS_BITS == 0))
    return g2_onesweep_sort_pairs(c, n, keys_io, vals_io, keys_alt, vals_alt, begin_bit, end_bit, capture_shift, capture_dest);
  int ntiles = g2_cdiv(n, SORT_TILE);
  if((size_t) ntiles * RADIX_BINS + 1 > c->tilehist_elems)
    return g2_fail(G2GPU_ERR_ARG, "sort: %d pairs exceed scratch", n);
  G2_CUDA(cudaFuncSetAttribute(sort_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SCATTER_SMEM));
  unsigned long long *kin = *keys_io, *kout = keys_alt;
  unsigned int *vin = *vals_io, *vout = vals_alt;
  int nbits = end_bit - begin_bit;
  int npass = (nbits + RADIX_BITS - 1) / RADIX_BITS;
  for(int p = 0; p < npass; p++)
    {
      int shift = begin_bit + p * RADIX_BITS;
      int bits = end_bit - shift < RADIX_BITS ? end_bit - shift : RADIX_BITS;
      unsigned int mask = (1u << bits) - 1u;
      sort_hist_kernel<<<ntiles, SORT_THREADS, 0, c->stream>>>(kin, c->tilehist, n, ntiles, shift, mask);
      c->launches++;
      G2_TRY(g2_scan_exclusive_u32(c, c->tilehist, c->tilehist, (size_t) ntiles * RADIX_BINS));
      sort_scatter_kernel<<<ntiles, SORT_THREADS, SCATTER_SMEM, c->stream>>>(kin, vin, kout, vout, c->tilehist, n, ntiles, shift, mask,
									      shift == capture_shift ? capture_dest : nullptr);
      c->launches++;
      unsigned long long *tk = kin; kin = kout; kout = tk;
      unsigned int *tv = vin; vin = vout; vout = tv;
    }
  G2_CUDA(cudaGetLastError());
  *keys_io = kin;
  *vals_io = vin;
  return 0;
}
