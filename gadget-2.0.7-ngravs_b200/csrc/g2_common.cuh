// g2_common.cuh — shared declarations of the g2gpu CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/g2gpu.h"

#define G2_MAXTOP 32768		// cap on top-level tree nodes held on the device (reference MAXTOPNODES = 200000)
#define G2_MAXDEPTH 21		// 3*21 = 63 key bits: deepest supported octree level below the root
#define G2_PH_BITS 18		// BITS_PER_DIMENSION (allvars.h:34)
#define G2_NSM_FALLBACK 148
#ifndef G2_DEFAULT_WALK_MODE
#define G2_DEFAULT_WALK_MODE 0
#endif

extern __thread char g2_errbuf[512];	// per host thread: the group path (g2_group.cu) drives one context per thread
int g2_fail(int code, const char *fmt, ...);

#define G2_CUDA(call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if(e__ != cudaSuccess)                                                                              \
      return g2_fail(G2GPU_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
  } while(0)

#define G2_TRY(call)                                                                                    \
  do {                                                                                                  \
    int r__ = (call);                                                                                   \
    if(r__ != 0)                                                                                        \
      return r__;                                                                                       \
  } while(0)

static inline int g2_cdiv(long long a, long long b) { return (int) ((a + b - 1) / b); }

// 32-byte particle record (= g2gpu_particle of include/g2gpu.h): one random access fetches everything the tree code
// needs about a particle
struct __align__(16) G2PRec
{
  float x, y, z, m;
  int type;
  float oldacc;
  int active;
  int pad;
};

// ---- device copy of the top-level tree (domain.c TopNodes + forcetree.c top-level nodes) ----
struct G2TopTree
{
  int ntopnodes;		// NTopnodes
  int ntopleaves;		// NTopleaves
  int err;			// != 0: overflow
  int pad;
  // TopNodes[] (allvars.h:252-262)
  int daughter[G2_MAXTOP];
  int leaf[G2_MAXTOP];
  int shift[G2_MAXTOP];		// Size = 1 << shift  (Size is a power of 8)
  long long startkey[G2_MAXTOP];
  long long count[G2_MAXTOP];
  // force-tree node created for each TopNode by force_create_empty_nodes (forcetree.c:292-336)
  int fnode[G2_MAXTOP];		// reference node index - MaxPart (root = 0), indexed by TopNode
  // indexed by force-tree top node number k (0..ntopnodes-1, creation order = reference numbering)
  float flen[G2_MAXTOP];
  float fcx[G2_MAXTOP], fcy[G2_MAXTOP], fcz[G2_MAXTOP];
  int fdepth[G2_MAXTOP];	// depth below root
  unsigned long long fmorton[G2_MAXTOP];	// slot digits from the root, 3 bits per level (depth digits)
  int ffather[G2_MAXTOP];	// top node k's father (k index) or -1
  int fsuns[G2_MAXTOP][8];	// children (k index) by spatial slot, -1 = none (leaf)
  int fisleaf[G2_MAXTOP];	// 1 if it corresponds to a TopNodes leaf
  int fdfs[G2_MAXTOP];		// rank of top node k in the slot-order DFS (walk order)
  int fdfs_inv[G2_MAXTOP];	// inverse of fdfs
  int dni[G2_MAXTOP];		// DomainNodeIndex[leaf] - MaxPart  (k index)
};

// ---- walk constants ----
struct G2LawTable
{
  int accel[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];
  int spline[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];
  float par[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS][4];
};

// zero-copy results into an array of structures in pinned host memory (the reference's P[]): result of device-order particle i goes to
// base + perm[i] * stride, GravAccel[3] / OldAcc as float or double, GravCost as float (g2gpu_group_bind_results_aos)
struct G2ZcAos
{
  char *base;			// device-visible alias of the host array, nullptr: unused
  const int *perm;		// device order -> index in the host array
  unsigned long long stride;
  int off_acc, off_cost, off_old;	// byte offsets, < 0: not written
  int float_bytes;
};

// slice[] (device memory, written by slice_kernel before the walk): the walk never needs a host round trip for the target count
#define G2_SLICE_NTARGETS 0
#define G2_SLICE_LO 1
#define G2_SLICE_HI 2

struct g2gpu_ctx
{
  g2gpu_config cfg;
  int D;			// n_gravs
  int nsm;
  cudaStream_t stream;
  cudaEvent_t ev[20];

  int npart;			// current particle count
  int nactive;			// active particles (all ranks)
  int stage;			// 0 created, 1 uploaded, 2 domain done, 3 tree built, 4 walked
  int type_to_grav[6];
  double force_softening[6];
  G2LawTable laws;
  int laws_set;
  double *d_srtable;		// D*D*NTAB doubles
  float *d_srtable_f;		// same as float
  int srtable_set;
  int sr_ntables;		// unique short-range tables (identical pair tables are stored once)
  unsigned char sr_tabmap[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];
  // potential walk (g2_pot.cu)
  int potfxn[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS], potspline[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];	// PotentialFxns / PotentialSplines [tgt*D+src]
  int potlaws_set, pottable_set, pot_valid;
  float *d_pottable_f;		// unique shortrange_fourier_pot tables, NTAB floats each
  int pot_ntables;
  unsigned char pot_tabmap[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];
  float *pot;			// n, current particle order (allocated on first use)
  double pot_ms;
  // lattice-sum correction of a periodic box without PM (g2_lattice.cu)
  float *d_lattice;		// unique tables, (EN+1)^3 float4 (fx, fy, fz, 0) each
  int lattice_en, lattice_set, lattice_ntables;
  double *d_potcorr;		// unique lattice-sum potential tables, (EN+1)^3 doubles each (g2gpu_set_lattice_pot_tables)
  int potcorr_en, potcorr_set;
  unsigned char potcorr_tabmap[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];
  unsigned char lattice_tabmap[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];
  float *latt, *lattcost;	// 3n / n, current particle order (allocated on first use)
  int acc_double;		// accumulate accelerations in FP64 (default) or FP32
  int accumulator;		// NGRAVS_ACCUMULATOR: nodes carry the particle count per species (wcnt), laws receive it as N
  int counts_valid;		// wcnt belongs to the current tree
  unsigned int *wcnt;		// [U][D], filled by g2_stage_counts after the build when accumulator != 0
  int direct_ewald;		// g2gpu_direct adds the exact lattice correction of a periodic box (option "direct_ewald")
  int walk_stats;		// option "walk_stats": the instrumented walk kernel (visits, species terms, decisions in counters[2..4])
  int walk_exact;		// option "walk_exact" (default 1): borderline decisions are re-taken in FP64 (exact GravCost)
  int walk_defer;		// option "walk_defer" (G2GPU_WALK_DEFER; -DG2_WALK_DEFER builds only): stock pair laws evaluated from the per-warp source ring
  int walk_flush_mask;		// G2GPU_WALK_FLUSH_MASK (0, 1, 3, 7 = default): the FP32 partial sums of the walk go into the accumulators at descents with (cell index & mask) == 0
  int compact;			// option "compact": the walk writes its slice's results in target order (cres) instead of by particle index
  float *cres;			// 5 floats per target of the slice (acc[3], cost, oldacc), allocated with the option
  G2ZcAos zc_aos;		// the same for an array of structures
  float *zc_acc, *zc_cost, *zc_oldacc;	// set by the multi-GPU group for one walk: device-visible pointers to the caller's pinned result arrays (WalkArgs)
  int slice_explicit;		// slice_frac[] instead of rank/nranks (cost-weighted slices of the group path)
  double slice_frac[2];
  int tree_dynamic;		// the current tree came from g2gpu_update_tree (host-drifted nodes)

  // upload-order inputs
  G2PRec *in_rec;		// may point at caller-bound device memory (g2gpu_bind_inputs)
  G2PRec *own_in_rec;		// the library's own input buffer
  float *in_vel;		// 3n, optional
  float *in_gravpm;		// 3n, optional
  G2PRec *h_rec;		// pinned host staging of g2gpu_upload_aos (records | velocities | GravPM), allocated on first use
  float *h_vel, *h_gravpm;
  char *d_export, *h_export;	// staging of g2gpu_download_tree / g2gpu_download_extnodes (device, pinned host), grown on demand
  size_t export_bytes;
  float *in_raw;		// H2D landing zone: pos[3n] | mass[n] | type[n] | oldacc[n] | active[n]
  int inputs_bound;
  size_t h2d_bytes, d2h_bytes;
  int have_vel, have_gravpm;

  // current-order (species-major PH) particle arrays
  G2PRec *prec;
  float *vel;
  float *gravpm;
  long long *phkey;
  int *perm;
  unsigned int *phorder;	// particle index (current order) by rank along the Peano-Hilbert curve of ALL species
  double *h_domain;		// pinned copy of d_domain (valid after the build's host synchronisation)
  double coord_max;		// largest |coordinate| of the domain cube

  // sort scratch (ping-pong)
  unsigned long long *skey[2];
  unsigned int *sval[2];
  int sort_onesweep;		// option "sort_onesweep" (default 1): one kernel per radix digit with decoupled look-back
  int sort_items;		// pairs per thread of a one-sweep pass: 8 (default) or 16 (G2GPU_SORT_ITEMS)
  int sort_rank_ballot;	// rank the digits of a tile with one warp vote per digit bit instead of MATCH.ANY (G2GPU_SORT_RANK_BALLOT)
  int sort_window;		// predecessor tiles fetched per look-back step: 4 (default) or 8 (G2GPU_SORT_WINDOW)
  unsigned int *tilehist;	// NBINS * ntiles
  size_t tilehist_elems;
  unsigned int *scan_tmp;	// block sums for the device-wide scan
  size_t scan_tmp_elems;

  // domain
  double *d_domain;		// corner[3], center[3], len, fac  (+ scratch)
  float *d_minmax;		// 6 floats
  G2TopTree *d_top;
  void *d_topscratch;
  G2TopTree *h_top;		// pinned host copy (filled on demand)
  int *d_species_start;		// block boundaries in current order: D+2 ints (gas block, species blocks)

  // tree (sorted-by-treekey order p)
  unsigned long long *tkey;	// sorted tree keys
  unsigned int *tq;		// particle index at sorted position p
  unsigned char *tm;		// common depth of positions p,p+1 (levels from root), 255 = different top leaf
  unsigned short *ttl;		// top-tree node k (force top node index) holding position p
  unsigned int *tbase;		// exclusive scan of cells owned by pair p  (size n+1)
  unsigned int *tcnt;
  int ncells;			// regular (non-top) cells
  int numnodes;			// ntopnodes + ncells
  // per regular cell (provisional id c, DFS pre-order)
  unsigned int *c_a, *c_b;	// sorted-position range [a,b]
  unsigned char *c_d;		// depth
  int *c_suns;			// 8 per cell: -1 none; >=0 particle sorted position; <= -2: cell id = -(v+2)
  int *p_parent;		// per sorted position: parent code (same encoding as c_father)
  int *c_father;		// provisional: >=0 regular cell id, <= -2: top node k = -(v+2)
  unsigned int *c_ready;	// child-cell completion counters
  unsigned char *c_nchild;	// number of child cells
  unsigned char *c_npart;	// number of direct particle children
  unsigned int *c_min1, *c_min2;	// smallest / second smallest particle index in the cell
  unsigned int *c_poff;		// exclusive scan of c_npart (offset in the leaf-grouped particle array), incl. top nodes
  unsigned int *c_refid;	// reference node number (k index, MaxPart-relative) of regular cell c
  // top-node extras (size G2_MAXTOP)
  int *t_suns;			// children of top LEAF nodes among regular cells/particles (8 per top node)
  unsigned int *t_first, *t_last;	// sorted-position range of each top node's particles (first > last: empty)
  unsigned int *t_min1, *t_min2;
  unsigned int *t_ready;
  unsigned char *t_npart, *t_nchild;
  unsigned int *t_ubase;	// U index of top node k
  // moments in provisional indexing (U order): the walk records
  float4 *wcells;		// (2+D) float4 per cell, U order
  float4 *wpart;		// leaf-grouped particle records
  unsigned int *wsrc;		// particle index (current order) behind every record of wpart (dynamic tree update)
  int tree_npart;		// particles of the last tree construction (0: none)
  unsigned int *hist2, *hist2_scan;	// reference renumbering scratch (size n+1)
  unsigned int *dmin;
  int renumbered;
  unsigned int depth_start[32];	// cells grouped by depth (list in c_ready)
  int depth_count[32];
  int maxdepth;
  unsigned int *d_depth;	// [0..31] cells per depth, [32..63] scatter cursors
  int *d_err;			// device error flags (4 ints)
  int *h_err;			// pinned

  // walk
  unsigned int *w_targets;	// sorted positions of active targets (compacted)
  unsigned int *w_flags;	// scratch for compaction
  int w_ntargets;		// total active (all ranks)
  int w_lo, w_hi;		// this rank's slice of w_targets (host copies, valid after g2_fetch_slice)
  int *d_slice, *h_slice;	// [G2_SLICE_*] on the device / pinned
  void *d_exact;		// WalkExactParams of the current walk
  unsigned int *d_smcount;	// 1024 chunk counters of the walk's per-SM work distribution
  int walk_carveout;		// option / G2GPU_WALK_CARVEOUT: cudaFuncAttributePreferredSharedMemoryCarveout of walk_kernel in percent, -1 = driver default
  int walk_sm_local;		// option / G2GPU_WALK_SM_LOCAL (default 1): chunks dealt out in one contiguous block per SM
  int slice_pending;
  float *acc;			// 3n, current particle order
  float *cost;			// n
  float *oldacc_out;		// n
  unsigned long long *d_counters;	// [0] sum cost, [1] node visits, [2] particle visits
  unsigned long long *h_counters;

  // periodic PM long-range force (g2_pm.cu)
  int pm_grid, pm_fwd, pm_inv, pm_plans_valid, pm_done;	// mesh size the buffers/plans were made for; cuFFT handles
  void *pm_rho, *pm_rk, *pm_potk, *pm_tab;
  double pm_tab_asmth2;
  int pm_tab_n;	// real mesh (density, then potential), D spectra, filtered spectrum of one target species

  void *h_stage;		// pinned upload staging
  size_t h_stage_bytes;
  double ms[8];
  long long launches;
};

// ---- cross-file entry points ----
int g2_scan_exclusive_u32(g2gpu_ctx *c, const unsigned int *in, unsigned int *out, size_t n);
int g2_radix_sort_pairs(g2gpu_ctx *c, int n, unsigned long long **keys_io, unsigned int **vals_io,
			unsigned long long *keys_alt, unsigned int *vals_alt, int begin_bit, int end_bit,
			int capture_shift = -1, unsigned int *capture_dest = nullptr);
int g2_ensure_optional_inputs(g2gpu_ctx *c, int want_vel, int want_gravpm);
int g2_upload_soa_shard(g2gpu_ctx *c, int n_total, int lo, int cnt, const float *pos, const float *mass, const int *type, const float *oldacc,
			const int *active);
int g2_upload_aos_shard(g2gpu_ctx *c, int n_total, int lo, int cnt, const void *P, size_t stride, int float_bytes, int off_pos, int off_mass,
			int off_type, int off_oldacc, int off_vel, int off_gravpm, int off_ti_endstep, int ti_current, unsigned int max_threads);
int g2_stage_targets(g2gpu_ctx *c, const g2gpu_walk_params *wp);
int g2_fetch_slice(g2gpu_ctx *c);
int g2_stage_domain(g2gpu_ctx *c);
int g2_stage_treebuild(g2gpu_ctx *c);
int g2_stage_renumber(g2gpu_ctx *c);
int g2_stage_walk(g2gpu_ctx *c, const g2gpu_walk_params *wp);
int g2_stage_potential(g2gpu_ctx *c, const g2gpu_walk_params *wp);
int g2_stage_lattice(g2gpu_ctx *c, const g2gpu_walk_params *wp);
int g2_make_ewald_table(g2gpu_ctx *c, int en, double *out);
int g2_make_ewald_pot_table(g2gpu_ctx *c, int en, double latticezero, double *out);
int g2_stage_counts(g2gpu_ctx *c);
int g2_update_tree(g2gpu_ctx *c, const float *len, const float *s);
int g2_pm_periodic(g2gpu_ctx *c, const g2gpu_pm_params *pp);
int g2_pm_potential_periodic(g2gpu_ctx *c, const g2gpu_pm_params *pp, float *potential);
int g2_pm_download(g2gpu_ctx *c, float *gravpm);
void g2_pm_destroy(g2gpu_ctx *c);
int g2_direct_sum(g2gpu_ctx *c, const g2gpu_walk_params *wp, int ntargets, const int *targets, double *acc);
int g2_export_nparticles(g2gpu_ctx *c, long long *out);
int g2_export_extnodes(g2gpu_ctx *c, float *vs);
int g2_export_tree(g2gpu_ctx *c, float *len, float *center, float *s, float *mass, int *bitflags, int *sibling,
		   int *nextnode, int *father, int *p_nextnode, int *p_father);
int g2_peano_keys_standalone(g2gpu_ctx *c, int n, const int *xyz, int bits, long long *keys);
int g2_eval_potentials_standalone(g2gpu_ctx *c, int n, int tgt, int src, const float *pm, const float *m, const float *r, const float *h,
				  const int *nn, float *out);
int g2_eval_pairs_standalone(g2gpu_ctx *c, int n, int tgt, int src, const float *pm, const float *m, const float *r,
			     const float *h, const int *nn, float *fac);

#define G2_LAUNCH(c) ((c)->launches++)
