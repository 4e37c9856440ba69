// g2_group.cu — the tree force on several GPUs of one node, inside the library and behind the C ABI (include/g2gpu.h, g2gpu_group_*).
//
// It plays the role of the parallel driver inside the reference's gravity_tree() (gravtree.c:102-285: export particles, walk, import
// results), re-designed for one 8 x B200 NVSwitch box: ONE host process, one context + stream + host thread per device.
//   upload    every device copies ITS 1/N of the particle records from host memory (sharded H2D over its own PCIe link) and one
//             ncclAllGather per step replicates the 32-byte records over NVLink (communicators from ncclCommInitAll);
//   build     every device runs the same deterministic keys -> sort -> tree build on the full set (replicated tree);
//   walk      device i walks slice i of the Peano-Hilbert-ordered active targets; slices are cut at equal GravCost of the previous
//             force computation (the reference's own balance measure, domain.c:859-862), boundaries at multiples of 32 targets, so
//             every result bit is independent of the number of devices;
//   download  every device returns only its slice (20 B of results + 4 B particle index per target, pinned staging) and host
//             threads scatter the slices into the caller's arrays -- SoA, or straight into the reference's P[] (AoS).
// No data-path collective other than the all-gather; results need no reduction.  With one device the same code runs without NCCL.
// NCCL is resolved at run time (dlopen of libnccl.so.2), so a single-GPU host needs no NCCL installation.
#include "g2_common.cuh"
#include <dlfcn.h>
#include <algorithm>
#include <thread>
#include <vector>

#define G2_GROUP_MAX 16
#define G2_COST_BLOCK 1024	// granularity of the GravCost profile that balances the slices

typedef struct ncclComm *ncclComm_t;
typedef int ncclResult_t;
struct G2Nccl
{
  void *lib;
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, int /* ncclDataType_t */, ncclComm_t, cudaStream_t);
  const char *(*GetErrorString)(ncclResult_t);
};
#define G2_NCCL_INT8 0		/* ncclInt8 / ncclChar (nccl.h) */

struct g2gpu_group
{
  int n;
  int devices[G2_GROUP_MAX];
  g2gpu_ctx *ctx[G2_GROUP_MAX];
  G2Nccl nccl;
  ncclComm_t comm[G2_GROUP_MAX];
  int have_comm;
  int npart;
  int per;			// records per device in the padded all-gather
  int cost_weighted;
  double frac[G2_GROUP_MAX + 1];	// slice boundaries as fractions of the target list
  std::vector<double> profile;	// GravCost per block of G2_COST_BLOCK targets of the last walk (target order)
  int profile_ntargets;
  // per device: pinned staging of the slice results
  float *h_res[G2_GROUP_MAX];
  unsigned int *h_idx[G2_GROUP_MAX];
  size_t h_cap[G2_GROUP_MAX];
  int lo[G2_GROUP_MAX], hi[G2_GROUP_MAX], ntargets;
  // zero-copy results: GravCost profile of a device's slice, summed on the device (cost_profile_kernel)
  double *d_prof[G2_GROUP_MAX], *h_prof[G2_GROUP_MAX];
  // results bound to an array of structures in host memory (g2gpu_group_bind_results_aos)
  void *aos_host;		// the bound array (nullptr: none), registered by this library when aos_registered
  size_t aos_bytes;
  int aos_registered, aos_walk;	// aos_walk: the last g2gpu_group_walk stored its results there
  G2ZcAos aos[G2_GROUP_MAX];	// per device: mapped alias, layout
  int zero_copy;		// G2GPU_ZERO_COPY (default 1): pinned result arrays of g2gpu_group_gravity_tree are written by the walk kernel itself
  int last_zero_copy;		// the last g2gpu_group_gravity_tree took that path
  char errs[G2_GROUP_MAX][512];
  long long h2d_bytes, d2h_bytes, gather_bytes;
};

static int device_profile(g2gpu_group *g, int i);

// ---- one host thread per device ----
template <class F>
static int run_all(g2gpu_group *g, F f)
{
  int rc[G2_GROUP_MAX];
  auto body = [&](int i) {
    g2_errbuf[0] = 0;
    if(cudaSetDevice(g->devices[i]) != cudaSuccess)
      rc[i] = g2_fail(G2GPU_ERR_CUDA, "cudaSetDevice(%d) failed", g->devices[i]);
    else
      rc[i] = f(i);
    snprintf(g->errs[i], sizeof(g->errs[i]), "%s", g2_errbuf);
  };
  if(g->n == 1)
    body(0);
  else
    {
      std::vector<std::thread> pool;
      for(int i = 0; i < g->n; i++)
	pool.emplace_back(body, i);
      for(auto &t : pool)
	t.join();
    }
  for(int i = 0; i < g->n; i++)
    if(rc[i] != 0)
      {
	snprintf(g2_errbuf, sizeof(g2_errbuf), "device %d: %s", g->devices[i], g->errs[i]);
	return rc[i];
      }
  return 0;
}

static int load_nccl(G2Nccl *N)
{
  if(N->lib)
    return 0;
  const char *names[] = { "libnccl.so.2", "libnccl.so" };
  for(int k = 0; k < 2 && !N->lib; k++)
    N->lib = dlopen(names[k], RTLD_NOW | RTLD_GLOBAL);
  if(!N->lib)
    return g2_fail(G2GPU_ERR_CUDA, "multi-GPU group: libnccl.so.2 not found (%s)", dlerror());
  N->CommInitAll = (ncclResult_t (*)(ncclComm_t *, int, const int *)) dlsym(N->lib, "ncclCommInitAll");
  N->CommDestroy = (ncclResult_t (*)(ncclComm_t)) dlsym(N->lib, "ncclCommDestroy");
  N->AllGather = (ncclResult_t (*)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t)) dlsym(N->lib, "ncclAllGather");
  N->GetErrorString = (const char *(*)(ncclResult_t)) dlsym(N->lib, "ncclGetErrorString");
  if(!N->CommInitAll || !N->CommDestroy || !N->AllGather || !N->GetErrorString)
    return g2_fail(G2GPU_ERR_CUDA, "multi-GPU group: NCCL symbols missing");
  return 0;
}

extern "C" int g2gpu_group_create(g2gpu_group **out, const g2gpu_config *cfg, int ndev, const int *devices)
{
  if(!out || !cfg)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  *out = nullptr;
  const int have = g2gpu_device_count();
  if(have == 0)
    return g2_fail(G2GPU_ERR_CUDA, "no CUDA device (this library has no CPU fallback)");
  if(ndev <= 0)
    ndev = have;		// all visible devices
  if(ndev > G2_GROUP_MAX || ndev > have)
    return g2_fail(G2GPU_ERR_ARG, "%d devices requested, %d visible (at most %d per group)", ndev, have, G2_GROUP_MAX);
  g2gpu_group *g = new g2gpu_group();
  g->n = ndev;
  g->have_comm = 0;
  g->npart = 0;
  g->per = 0;
  g->cost_weighted = 1;
  g->profile_ntargets = 0;
  g->ntargets = 0;
  g->h2d_bytes = g->d2h_bytes = g->gather_bytes = 0;
  g->zero_copy = getenv("G2GPU_ZERO_COPY") ? atoi(getenv("G2GPU_ZERO_COPY")) != 0 : 1;
  g->last_zero_copy = 0;
  g->aos_host = nullptr; g->aos_bytes = 0; g->aos_registered = 0; g->aos_walk = 0;
  memset(g->aos, 0, sizeof(g->aos));
  memset(&g->nccl, 0, sizeof(g->nccl));
  for(int i = 0; i < G2_GROUP_MAX; i++)
    {
      g->ctx[i] = nullptr;
      g->comm[i] = nullptr;
      g->h_res[i] = nullptr;
      g->d_prof[i] = g->h_prof[i] = nullptr;
      g->h_idx[i] = nullptr;
      g->h_cap[i] = 0;
      g->lo[i] = g->hi[i] = 0;
    }
  for(int i = 0; i < ndev; i++)
    {
      g->devices[i] = devices ? devices[i] : i;
      g->frac[i] = (double) i / ndev;
    }
  g->frac[ndev] = 1.0;
  for(int i = 0; i < ndev; i++)
    {
      g2gpu_config c = *cfg;
      c.device = g->devices[i];
      c.rank = i;
      c.nranks = ndev;
      int rc = g2gpu_create(&g->ctx[i], &c);
      if(rc == 0)
	rc = g2gpu_set_option(g->ctx[i], "compact", 1);
      if(rc)
	{
	  g2gpu_group_destroy(g);
	  return rc;
	}
    }
  if(ndev > 1)
    {
      int rc = load_nccl(&g->nccl);
      if(rc == 0)
	{
	  ncclResult_t r = g->nccl.CommInitAll(g->comm, ndev, g->devices);
	  if(r != 0)
	    rc = g2_fail(G2GPU_ERR_CUDA, "ncclCommInitAll: %s", g->nccl.GetErrorString(r));
	  else
	    g->have_comm = 1;
	}
      if(rc)
	{
	  g2gpu_group_destroy(g);
	  return rc;
	}
    }
  *out = g;
  return 0;
}

extern "C" void g2gpu_group_destroy(g2gpu_group *g)
{
  if(!g)
    return;
  for(int i = 0; i < g->n; i++)
    {
      if(g->ctx[i])
	{
	  cudaSetDevice(g->devices[i]);
	  g2gpu_sync(g->ctx[i]);
	}
      if(g->have_comm && g->comm[i])
	g->nccl.CommDestroy(g->comm[i]);
      if(g->h_res[i])
	cudaFreeHost(g->h_res[i]);
      if(g->h_idx[i])
	cudaFreeHost(g->h_idx[i]);
      if(g->h_prof[i])
	cudaFreeHost(g->h_prof[i]);
      if(g->d_prof[i])
	cudaFree(g->d_prof[i]);
      if(g->ctx[i])
	g2gpu_destroy(g->ctx[i]);
    }
  if(g->aos_host && g->aos_registered)
    cudaHostUnregister(g->aos_host);
  delete g;
}

extern "C" int g2gpu_group_size(g2gpu_group *g) { return g ? g->n : 0; }
extern "C" int g2gpu_group_zero_copy(g2gpu_group *g) { return g ? g->last_zero_copy : 0; }
extern "C" g2gpu_ctx *g2gpu_group_ctx(g2gpu_group *g, int i) { return (g && i >= 0 && i < g->n) ? g->ctx[i] : nullptr; }

// ---- tables and options go to every device ----
extern "C" int g2gpu_group_set_species(g2gpu_group *g, const int type_to_grav[6], const double force_softening[6])
{
  if(!g)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  for(int i = 0; i < g->n; i++)
    G2_TRY(g2gpu_set_species(g->ctx[i], type_to_grav, force_softening));
  return 0;
}

extern "C" int g2gpu_group_set_laws(g2gpu_group *g, const int *accel_id, const int *spline_id, const double *params)
{
  if(!g)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  for(int i = 0; i < g->n; i++)
    G2_TRY(g2gpu_set_laws(g->ctx[i], accel_id, spline_id, params));
  return 0;
}

extern "C" int g2gpu_group_set_srtable(g2gpu_group *g, const double *table)
{
  if(!g)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  for(int i = 0; i < g->n; i++)
    G2_TRY(g2gpu_set_srtable(g->ctx[i], table));
  return 0;
}

extern "C" int g2gpu_group_set_lattice_tables(g2gpu_group *g, int en, const double *fcorr)
{
  if(!g)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  for(int i = 0; i < g->n; i++)
    G2_TRY(g2gpu_set_lattice_tables(g->ctx[i], en, fcorr));
  return 0;
}

extern "C" int g2gpu_group_set_option(g2gpu_group *g, const char *name, int value)
{
  if(!g || !name)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(strcmp(name, "cost_weighted") == 0)
    {
      g->cost_weighted = value != 0;
      if(!value)
	for(int i = 0; i <= g->n; i++)
	  g->frac[i] = (double) i / g->n;
      return 0;
    }
  if(strcmp(name, "rank") == 0 || strcmp(name, "nranks") == 0 || strcmp(name, "compact") == 0)
    return g2_fail(G2GPU_ERR_ARG, "option %s is managed by the group", name);
  for(int i = 0; i < g->n; i++)
    G2_TRY(g2gpu_set_option(g->ctx[i], name, value));
  return 0;
}

// ---- upload: one shard per device, then the all-gather ----
static int group_allgather(g2gpu_group *g, int i, int with_gravpm)
{
  if(g->n == 1)
    return 0;
  g2gpu_ctx *c = g->ctx[i];
  const size_t per = (size_t) g->per;
  ncclResult_t r = g->nccl.AllGather((const char *) (c->own_in_rec + (size_t) i * per), c->own_in_rec, per * sizeof(G2PRec), G2_NCCL_INT8, g->comm[i], c->stream);
  if(r == 0 && with_gravpm)
    r = g->nccl.AllGather((const char *) (c->in_gravpm + 3 * (size_t) i * per), c->in_gravpm, per * 12, G2_NCCL_INT8, g->comm[i], c->stream);
  if(r != 0)
    return g2_fail(G2GPU_ERR_CUDA, "ncclAllGather: %s", g->nccl.GetErrorString(r));
  return 0;
}

static void shard_of(const g2gpu_group *g, int npart, int i, int *lo, int *cnt)
{
  const long long per = g->per;
  long long a = per * i, b = per * (i + 1);
  if(a > npart) a = npart;
  if(b > npart) b = npart;
  *lo = (int) a;
  *cnt = (int) (b - a);
}

static int group_begin_upload(g2gpu_group *g, int npart)
{
  if(!g)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(npart < 1 || npart > g->ctx[0]->cfg.max_part)
    return g2_fail(G2GPU_ERR_ARG, "npart %d outside [1, MaxPart=%d]", npart, g->ctx[0]->cfg.max_part);
  g->npart = npart;
  g->per = (npart + g->n - 1) / g->n;
  return 0;
}

extern "C" int g2gpu_group_upload(g2gpu_group *g, int npart, const float *pos, const float *mass, const int *type, const float *oldacc,
				  const int *active)
{
  G2_TRY(group_begin_upload(g, npart));
  if(!pos || !mass || !type)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_TRY(run_all(g, [&](int i) {
    int lo, cnt;
    shard_of(g, npart, i, &lo, &cnt);
    G2_TRY(g2_upload_soa_shard(g->ctx[i], npart, lo, cnt, pos, mass, type, oldacc, active));
    return group_allgather(g, i, 0);
  }));
  g->h2d_bytes = 0;
  for(int i = 0; i < g->n; i++)
    g->h2d_bytes += (long long) g->ctx[i]->h2d_bytes;
  g->gather_bytes = g->n > 1 ? (long long) g->per * g->n * (long long) sizeof(G2PRec) : 0;
  return 0;
}

extern "C" int g2gpu_group_upload_aos(g2gpu_group *g, int npart, const void *P, size_t stride, int float_bytes, int off_pos, int off_mass,
				      int off_type, int off_oldacc, int off_vel, int off_gravpm, int off_ti_endstep, int ti_current)
{
  G2_TRY(group_begin_upload(g, npart));
  if(!P)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(off_vel >= 0 && g->n > 1)
    return g2_fail(G2GPU_ERR_ARG, "velocities (Extnodes[].vs of the host tree mirror) are served by a single-device group only");
  const unsigned int hw = std::max(1u, std::thread::hardware_concurrency());
  const unsigned int thr = std::max(1u, std::min(32u, hw) / (unsigned int) g->n);
  G2_TRY(run_all(g, [&](int i) {
    int lo, cnt;
    shard_of(g, npart, i, &lo, &cnt);
    G2_TRY(g2_upload_aos_shard(g->ctx[i], npart, lo, cnt, P, stride, float_bytes, off_pos, off_mass, off_type, off_oldacc, off_vel, off_gravpm,
			       off_ti_endstep, ti_current, thr));
    return group_allgather(g, i, off_gravpm >= 0);
  }));
  g->h2d_bytes = 0;
  for(int i = 0; i < g->n; i++)
    g->h2d_bytes += (long long) g->ctx[i]->h2d_bytes;
  g->gather_bytes = g->n > 1 ? (long long) g->per * g->n * (long long) (sizeof(G2PRec) + (off_gravpm >= 0 ? 12 : 0)) : 0;
  return 0;
}

// device-resident variant (benchmarks): every device already holds ITS shard of the records in its own input buffer
// (g2gpu_input_buffers(g2gpu_group_ctx(g, i)) + shard offset); only the all-gather runs
extern "C" int g2gpu_group_gather_resident(g2gpu_group *g, int npart)
{
  G2_TRY(group_begin_upload(g, npart));
  G2_TRY(run_all(g, [&](int i) {
    G2_TRY(g2gpu_inputs_ready(g->ctx[i], npart));
    return group_allgather(g, i, 0);
  }));
  g->gather_bytes = g->n > 1 ? (long long) g->per * g->n * (long long) sizeof(G2PRec) : 0;
  return 0;
}

extern "C" int g2gpu_group_shard(g2gpu_group *g, int npart, int i, int *lo, int *cnt)
{
  if(!g || i < 0 || i >= g->n || npart < 1)
    return g2_fail(G2GPU_ERR_ARG, "bad argument");
  const int per_save = g->per;
  g->per = (npart + g->n - 1) / g->n;
  int a, b;
  shard_of(g, npart, i, &a, &b);
  g->per = per_save;
  if(lo)
    *lo = a;
  if(cnt)
    *cnt = b;
  return 0;
}

// ---- stages ----
extern "C" int g2gpu_group_domain(g2gpu_group *g)
{
  if(!g)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  return run_all(g, [&](int i) { return g2_stage_domain(g->ctx[i]); });
}

extern "C" int g2gpu_group_treebuild(g2gpu_group *g, int *numnodes)
{
  if(!g)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_TRY(run_all(g, [&](int i) { return g2_stage_treebuild(g->ctx[i]); }));
  if(numnodes)
    *numnodes = g->ctx[0]->numnodes;
  return 0;
}

extern "C" int g2gpu_group_update_tree(g2gpu_group *g, const float *len, const float *s)
{
  if(!g || !len || !s)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  return run_all(g, [&](int i) { return g2_update_tree(g->ctx[i], len, s); });
}

extern "C" int g2gpu_group_walk(g2gpu_group *g, const g2gpu_walk_params *wp)
{
  if(!g || !wp)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  g->aos_walk = g->aos_host != nullptr;
  return run_all(g, [&](int i) {
    g2gpu_ctx *c = g->ctx[i];
    c->compact = 1;		// (the slice downloads and the GravCost profile read the compacted results)
    c->slice_explicit = 1;
    c->slice_frac[0] = g->frac[i];
    c->slice_frac[1] = g->frac[i + 1];
    if(g->aos_walk)
      {				// results also go straight into the bound array of structures
	c->zc_aos = g->aos[i];
	c->zc_aos.perm = c->perm;
      }
    const int rc = g2_stage_walk(c, wp);
    c->zc_aos.base = nullptr;
    return rc;
  });
}

// ---- download: every device returns its slice; host threads scatter ----
static int fetch_slice_results(g2gpu_group *g, int i)
{
  g2gpu_ctx *c = g->ctx[i];
  if(c->stage < 4)
    return g2_fail(G2GPU_ERR_STATE, "g2gpu_group_walk has not run");
  G2_TRY(g2_fetch_slice(c));
  g->lo[i] = c->w_lo;
  g->hi[i] = c->w_hi;
  const size_t cnt = (size_t) (c->w_hi - c->w_lo);
  if(cnt > g->h_cap[i])
    {
      if(g->h_res[i])
	cudaFreeHost(g->h_res[i]);
      if(g->h_idx[i])
	cudaFreeHost(g->h_idx[i]);
      g->h_res[i] = nullptr;
      g->h_idx[i] = nullptr;
      const size_t cap = cnt + cnt / 4 + 1024;
      G2_CUDA(cudaMallocHost((void **) &g->h_res[i], sizeof(float) * 5 * cap));
      G2_CUDA(cudaMallocHost((void **) &g->h_idx[i], sizeof(unsigned int) * cap));
      g->h_cap[i] = cap;
    }
  cudaStream_t st = c->stream;
  G2_CUDA(cudaEventRecord(c->ev[11], st));
  if(cnt > 0)
    {
      G2_CUDA(cudaMemcpyAsync(g->h_res[i], c->cres, sizeof(float) * 5 * cnt, cudaMemcpyDeviceToHost, st));
      G2_CUDA(cudaMemcpyAsync(g->h_idx[i], c->w_targets + c->w_lo, sizeof(unsigned int) * cnt, cudaMemcpyDeviceToHost, st));
    }
  G2_CUDA(cudaEventRecord(c->ev[12], st));
  G2_CUDA(cudaStreamSynchronize(st));
  c->d2h_bytes = cnt * 24;
  return 0;
}

template <class W>
static void scatter_threads(size_t cnt, unsigned int nthr, W work, size_t min_per_thread = 131072)
{
  nthr = std::max(1u, std::min(nthr, (unsigned int) (cnt / min_per_thread + 1)));
  if(nthr == 1)
    {
      work(0, cnt);
      return;
    }
  std::vector<std::thread> pool;
  for(unsigned int t = 0; t < nthr; t++)
    pool.emplace_back(work, cnt * t / nthr, cnt * (t + 1) / nthr);
  for(auto &t : pool)
    t.join();
}

// GravCost profile of the walk just downloaded -> slice boundaries of the next one (equal cost per device)
static void rebalance(g2gpu_group *g, bool device_profile = false)
{
  int nt = 0;
  for(int i = 0; i < g->n; i++)
    nt = std::max(nt, g->hi[i]);
  g->ntargets = g->ctx[0]->w_ntargets;
  if(!g->cost_weighted || g->n == 1 || nt <= 0)
    return;
  const int nb = (nt + G2_COST_BLOCK - 1) / G2_COST_BLOCK;
  g->profile.assign((size_t) nb, 0.0);
  const unsigned int hw = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  for(int i = 0; i < g->n; i++)
    {
      // blocks of the profile that overlap this device's slice, shared out over host threads (a block is summed by one thread; the
      // block that straddles two slices is visited once per slice, and the slices are taken one after the other)
      const float *r = g->h_res[i];
      const int lo = g->lo[i], hi = g->hi[i];
      if(hi <= lo)
	continue;
      const size_t b0 = (size_t) lo / G2_COST_BLOCK, b1 = (size_t) (hi - 1) / G2_COST_BLOCK + 1;
      double *prof = g->profile.data();
      if(device_profile)
	{			// summed by cost_profile_kernel: h_prof[i][k] belongs to block b0 + k
	  for(size_t blk = b0; blk < b1; blk++)
	    prof[blk] += g->h_prof[i][blk - b0];
	  continue;
	}
      scatter_threads(b1 - b0, hw, [=](size_t a, size_t b) {
	for(size_t blk = b0 + a; blk < b0 + b; blk++)
	  {
	    const size_t t0 = std::max((size_t) lo, blk * G2_COST_BLOCK), t1 = std::min((size_t) hi, (blk + 1) * G2_COST_BLOCK);
	    double s = 0.0;
	    for(size_t t = t0; t < t1; t++)
	      s += (double) r[5 * (t - (size_t) lo) + 3] + 1.0;	// + 1: a target costs something even without interactions
	    prof[blk] += s;
	  }
      }, 128);
    }
  g->profile_ntargets = nt;
  double total = 0;
  for(double v : g->profile)
    total += v;
  if(!(total > 0))
    return;
  double run = 0;
  int dev = 1;
  for(int b = 0; b < nb && dev < g->n; b++)
    {
      const double next = run + g->profile[(size_t) b];
      while(dev < g->n && next >= total * dev / g->n)
	{			// the boundary falls inside block b: interpolate
	  const double want = total * dev / g->n;
	  const double f = g->profile[(size_t) b] > 0 ? (want - run) / g->profile[(size_t) b] : 0.0;
	  double pos = ((double) b + f) * G2_COST_BLOCK;
	  if(pos > nt)
	    pos = nt;
	  g->frac[dev] = pos / nt;
	  dev++;
	}
      run = next;
    }
  for(; dev < g->n; dev++)
    g->frac[dev] = 1.0;
  g->frac[0] = 0.0;
  g->frac[g->n] = 1.0;
}

// acc[3n], cost[n], oldacc[n] in CURRENT (device) particle order, like g2gpu_download_acc; only active targets are written
extern "C" int g2gpu_group_download_acc(g2gpu_group *g, float *acc, float *cost, float *oldacc)
{
  if(!g)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  const unsigned int hw = std::max(1u, std::thread::hardware_concurrency());
  const unsigned int thr = std::max(1u, std::min(32u, hw) / (unsigned int) g->n);
  G2_TRY(run_all(g, [&](int i) {
    G2_TRY(fetch_slice_results(g, i));
    const float *r = g->h_res[i];
    const unsigned int *ix = g->h_idx[i];
    scatter_threads((size_t) (g->hi[i] - g->lo[i]), thr, [=](size_t a, size_t b) {
      for(size_t k = a; k < b; k++)
	{
	  const size_t p = ix[k];
	  if(acc)
	    {
	      acc[3 * p + 0] = r[5 * k + 0];
	      acc[3 * p + 1] = r[5 * k + 1];
	      acc[3 * p + 2] = r[5 * k + 2];
	    }
	  if(cost)
	    cost[p] = r[5 * k + 3];
	  if(oldacc)
	    oldacc[p] = r[5 * k + 4];
	}
    });
    return 0;
  }));
  g->d2h_bytes = 0;
  for(int i = 0; i < g->n; i++)
    g->d2h_bytes += (long long) g->ctx[i]->d2h_bytes;
  rebalance(g);
  return 0;
}

// The same, straight into the reference's AoS: P[perm[p]].GravAccel / GravCost / OldAcc for every active target p (device order);
// perm == NULL: P[] is in device order (after peano_hilbert_order).  float_bytes = sizeof(FLOAT) of GravAccel and OldAcc; GravCost is a
// float in either build (allvars.h:572).  *cost_sum = sum of GravCost.
extern "C" int g2gpu_group_download_aos(g2gpu_group *g, void *P, size_t stride, int float_bytes, int off_gravaccel, int off_gravcost, int off_oldacc,
					const int *perm, double *cost_sum)
{
  if(!g || !P)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(float_bytes != 4 && float_bytes != 8)
    return g2_fail(G2GPU_ERR_ARG, "float_bytes must be 4 or 8");
  double sums[G2_GROUP_MAX];
  if(g->aos_walk && P == g->aos_host && stride == (size_t) g->aos[0].stride && float_bytes == g->aos[0].float_bytes && off_gravaccel == g->aos[0].off_acc
     && off_gravcost == g->aos[0].off_cost && off_oldacc == g->aos[0].off_old)
    {				// the walk kernel has stored the results in P[] itself: wait for the devices, take their GravCost profiles
      G2_TRY(run_all(g, [&](int i) {
	G2_TRY(device_profile(g, i));
	double sum = 0.0;
	if(g->hi[i] > g->lo[i])
	  for(size_t k = 0, nb = (size_t) (g->hi[i] - 1) / G2_COST_BLOCK - (size_t) g->lo[i] / G2_COST_BLOCK + 1; k < nb; k++)
	    sum += g->h_prof[i][k];
	sums[i] = sum - (double) (g->hi[i] - g->lo[i]);	// (the profile counts + 1 per target)
	g->ctx[i]->d2h_bytes = (size_t) (g->hi[i] - g->lo[i]) * (size_t) (float_bytes == 4 ? 20 : 36) + sizeof(double) * (size_t) (g->ctx[i]->cfg.max_part / G2_COST_BLOCK + 2);
	return 0;
      }));
      if(cost_sum)
	{
	  *cost_sum = 0;
	  for(int i = 0; i < g->n; i++)
	    *cost_sum += sums[i];
	}
      g->d2h_bytes = 0;
      for(int i = 0; i < g->n; i++)
	g->d2h_bytes += (long long) g->ctx[i]->d2h_bytes;
      g->last_zero_copy = 1;
      rebalance(g, true);
      return 0;
    }
  g->last_zero_copy = 0;
  const unsigned int hw = std::max(1u, std::thread::hardware_concurrency());
  const unsigned int thr = std::max(1u, std::min(32u, hw) / (unsigned int) g->n);
  G2_TRY(run_all(g, [&](int i) {
    G2_TRY(fetch_slice_results(g, i));
    const float *r = g->h_res[i];
    const unsigned int *ix = g->h_idx[i];
    char *base = (char *) P;
    scatter_threads((size_t) (g->hi[i] - g->lo[i]), thr, [=](size_t a, size_t b) {
      for(size_t k = a; k < b; k++)
	{
	  const size_t p = perm ? (size_t) perm[ix[k]] : (size_t) ix[k];
	  char *q = base + p * stride;
	  if(float_bytes == 4)
	    {
	      float *ga = (float *) (q + off_gravaccel);
	      ga[0] = r[5 * k + 0]; ga[1] = r[5 * k + 1]; ga[2] = r[5 * k + 2];
	      if(off_gravcost >= 0)
		*(float *) (q + off_gravcost) = r[5 * k + 3];
	      if(off_oldacc >= 0)
		*(float *) (q + off_oldacc) = r[5 * k + 4];
	    }
	  else
	    {
	      double *ga = (double *) (q + off_gravaccel);
	      ga[0] = r[5 * k + 0]; ga[1] = r[5 * k + 1]; ga[2] = r[5 * k + 2];
	      if(off_gravcost >= 0)
		*(float *) (q + off_gravcost) = r[5 * k + 3];	/* float in either build */
	      if(off_oldacc >= 0)
		*(double *) (q + off_oldacc) = r[5 * k + 4];
	    }
	}
    });
    double s = 0;
    const size_t cnt = (size_t) (g->hi[i] - g->lo[i]);
    for(size_t k = 0; k < cnt; k++)
      s += (double) r[5 * k + 3];
    sums[i] = s;
    return 0;
  }));
  if(cost_sum)
    {
      *cost_sum = 0;
      for(int i = 0; i < g->n; i++)
	*cost_sum += sums[i];
    }
  g->d2h_bytes = 0;
  for(int i = 0; i < g->n; i++)
    g->d2h_bytes += (long long) g->ctx[i]->d2h_bytes;
  rebalance(g);
  return 0;
}

extern "C" int g2gpu_group_get_order(g2gpu_group *g, int *perm)
{
  if(!g)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  return g2gpu_get_order(g->ctx[0], perm);
}

extern "C" int g2gpu_group_sync(g2gpu_group *g)
{
  if(!g)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  return run_all(g, [&](int i) { return g2gpu_sync(g->ctx[i]); });
}

// ---- zero-copy results: pinned result arrays are written by the walk kernel itself ----
static float *mapped_host_ptr(const void *p);

// device-visible alias of a host pointer on the current device, or null when the memory is not pinned and mapped there
static float *mapped_host_ptr(const void *p)
{
  if(!p)
    return nullptr;
  cudaPointerAttributes a;
  if(cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
      cudaGetLastError();
      return nullptr;
    }
  if(a.type != cudaMemoryTypeHost || !a.devicePointer)
    return nullptr;
  return (float *) a.devicePointer;
}

// GravCost (+ 1 per target) of this device's slice per block of G2_COST_BLOCK targets, from the compact results: one warp per block;
// prof[k] belongs to block lo / G2_COST_BLOCK + k.  The slice bounds are read on the device (no host round trip before the launch).
__global__ void __launch_bounds__(256) cost_profile_kernel(const float *__restrict__ cres, const int *__restrict__ slice, int nwarps, double *__restrict__ prof)
{
  const int warp = (int) ((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if(warp >= nwarps)
    return;
  const long long lo = slice[G2_SLICE_LO], hi = slice[G2_SLICE_HI];
  const long long blk = lo / G2_COST_BLOCK + warp;
  const long long t0 = blk * G2_COST_BLOCK > lo ? blk * G2_COST_BLOCK : lo, t1 = (blk + 1) * G2_COST_BLOCK < hi ? (blk + 1) * G2_COST_BLOCK : hi;
  double s = 0.0;
  for(long long t = t0 + lane; t < t1; t += 32)
    s += (double) cres[5 * (t - lo) + 3] + 1.0;
#pragma unroll
  for(int o = 16; o > 0; o >>= 1)
    s += __shfl_xor_sync(0xffffffffu, s, o);
  if(lane == 0)
    prof[warp] = s;
}

// GravCost profile of device i's slice from its compact results, and the slice bounds; synchronises the device's stream
static int device_profile(g2gpu_group *g, int i)
{
  g2gpu_ctx *c = g->ctx[i];
  const int nbcap = c->cfg.max_part / G2_COST_BLOCK + 2;
  if(!g->d_prof[i])
    {
      G2_CUDA(cudaMalloc((void **) &g->d_prof[i], sizeof(double) * (size_t) nbcap));
      G2_CUDA(cudaMallocHost((void **) &g->h_prof[i], sizeof(double) * (size_t) nbcap));
    }
  cost_profile_kernel<<<g2_cdiv(nbcap * 32, 256), 256, 0, c->stream>>>(c->cres, (const int *) c->d_slice, nbcap, g->d_prof[i]);
  G2_CUDA(cudaMemcpyAsync(g->h_prof[i], g->d_prof[i], sizeof(double) * (size_t) nbcap, cudaMemcpyDeviceToHost, c->stream));
  G2_TRY(g2_fetch_slice(c));
  G2_CUDA(cudaStreamSynchronize(c->stream));
  G2_CUDA(cudaGetLastError());
  g->lo[i] = c->w_lo;
  g->hi[i] = c->w_hi;
  return 0;
}

// Binds the results of the following g2gpu_group_walk calls to an array of structures in host memory (the reference's P[], allvars.h:548-590):
// the walk kernel then stores GravAccel / GravCost / OldAcc of every active target straight into it (P[order[i]] for device-order particle i)
// and g2gpu_group_download_aos only waits for the devices.  The array is page-locked here (cudaHostRegister) unless it already is; when that
// fails, or with G2GPU_ZERO_COPY=0, nothing is bound and the download stages and scatters as before.  P == NULL removes the binding.
extern "C" int g2gpu_group_bind_results_aos(g2gpu_group *g, void *P, size_t nelem, size_t stride, int float_bytes, int off_gravaccel, int off_gravcost, int off_oldacc)
{
  if(!g)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(g->aos_host)
    {
      G2_TRY(g2gpu_group_sync(g));
      if(g->aos_registered)
	cudaHostUnregister(g->aos_host);
      g->aos_host = nullptr;
      g->aos_registered = 0;
    }
  if(!P || !g->zero_copy)
    return 0;
  if((float_bytes != 4 && float_bytes != 8) || off_gravaccel < 0 || stride == 0 || nelem == 0)
    return g2_fail(G2GPU_ERR_ARG, "bind_results_aos: bad layout");
  if(cudaSetDevice(g->devices[0]) != cudaSuccess)
    return g2_fail(G2GPU_ERR_CUDA, "cudaSetDevice(%d) failed", g->devices[0]);
  int registered = 0;
  if(!mapped_host_ptr(P))
    {
      if(cudaHostRegister(P, nelem * stride, cudaHostRegisterPortable | cudaHostRegisterMapped) != cudaSuccess)
	{
	  cudaGetLastError();
	  return 0;		// cannot be page-locked: staged downloads
	}
      registered = 1;
    }
  for(int i = 0; i < g->n; i++)
    {
      cudaSetDevice(g->devices[i]);
      g->aos[i].base = (char *) mapped_host_ptr(P);
      g->aos[i].perm = nullptr;
      g->aos[i].stride = (unsigned long long) stride;
      g->aos[i].off_acc = off_gravaccel; g->aos[i].off_cost = off_gravcost; g->aos[i].off_old = off_oldacc;
      g->aos[i].float_bytes = float_bytes;
      if(!g->aos[i].base)
	{
	  if(registered)
	    cudaHostUnregister(P);
	  return 0;
	}
    }
  g->aos_host = P;
  g->aos_bytes = nelem * stride;
  g->aos_registered = registered;
  return 0;
}

// Whole step with host buffers: sharded upload + all-gather -> domain -> treebuild -> walk of N slices -> slice downloads (the e2e path)
extern "C" int g2gpu_group_gravity_tree(g2gpu_group *g, int npart, const float *pos, const float *mass, const int *type, const float *oldacc,
					const int *active, const g2gpu_walk_params *wp, float *acc, float *cost, float *oldacc_out, int *perm)
{
  if(!g || !wp)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_TRY(group_begin_upload(g, npart));
  if(!pos || !mass || !type)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  // one device and every particle a target: the walk writes its results in particle order and they are copied straight into the caller's
  // arrays -- no compaction, no index list, no host scatter
  const bool direct = g->n == 1 && !active;
  // pinned result arrays (cudaHostAlloc / cudaHostRegister'd, mapped on every device of the group): the epilogue of the walk kernel stores each
  // target's results straight into them -- posted writes over PCIe spread over the whole walk -- so no download, no staging and no host scatter
  // follow the walk; what is left for the host is the GravCost profile of the balancer, summed on the device (128 KB per device)
  int zc = g->zero_copy && acc != nullptr;
  float *zacc[G2_GROUP_MAX], *zcost[G2_GROUP_MAX], *zold[G2_GROUP_MAX];
  for(int i = 0; i < g->n && zc; i++)
    {
      if(cudaSetDevice(g->devices[i]) != cudaSuccess)
	return g2_fail(G2GPU_ERR_CUDA, "cudaSetDevice(%d) failed", g->devices[i]);
      zacc[i] = mapped_host_ptr(acc); zcost[i] = mapped_host_ptr(cost); zold[i] = mapped_host_ptr(oldacc_out);
      if(!zacc[i] || (cost && !zcost[i]) || (oldacc_out && !zold[i]))
	zc = 0;
    }
  g->last_zero_copy = zc;
  const size_t nbcap = (size_t) (g->ctx[0]->cfg.max_part / G2_COST_BLOCK + 2);
  // one host thread per device drives its whole pipeline: no host barrier between the stages (the all-gather synchronises the devices)
  G2_TRY(run_all(g, [&](int i) {
    g2gpu_ctx *c = g->ctx[i];
    int lo, cnt;
    shard_of(g, npart, i, &lo, &cnt);
    c->compact = direct ? 0 : 1;
    G2_TRY(g2_upload_soa_shard(c, npart, lo, cnt, pos, mass, type, oldacc, active));
    G2_TRY(group_allgather(g, i, 0));
    G2_TRY(g2_stage_domain(c));
    G2_TRY(g2_stage_treebuild(c));
    c->slice_explicit = 1;
    c->slice_frac[0] = g->frac[i];
    c->slice_frac[1] = g->frac[i + 1];
    if(!zc)
      return g2_stage_walk(c, wp);
    c->zc_acc = zacc[i]; c->zc_cost = zcost[i]; c->zc_oldacc = zold[i];
    const int rc = g2_stage_walk(c, wp);
    c->zc_acc = c->zc_cost = c->zc_oldacc = nullptr;
    if(rc)
      return rc;
    const bool want_profile = g->cost_weighted && g->n > 1;
    if(want_profile)
      G2_TRY(device_profile(g, i));
    else
      {
	G2_TRY(g2_fetch_slice(c));	// (synchronises the stream: the results are in the caller's arrays)
	G2_CUDA(cudaStreamSynchronize(c->stream));
	G2_CUDA(cudaGetLastError());
	g->lo[i] = c->w_lo;
	g->hi[i] = c->w_hi;
      }
    c->d2h_bytes = (size_t) (c->w_hi - c->w_lo) * 20 + (want_profile ? sizeof(double) * (size_t) nbcap : 0);	// kernel stores into host memory + the profile
    return 0;
  }));
  g->h2d_bytes = 0;
  for(int i = 0; i < g->n; i++)
    g->h2d_bytes += (long long) g->ctx[i]->h2d_bytes;
  g->gather_bytes = g->n > 1 ? (long long) g->per * g->n * (long long) sizeof(G2PRec) : 0;
  if(zc)
    {
      g->ntargets = g->ctx[0]->w_ntargets;
      g->d2h_bytes = 0;
      for(int i = 0; i < g->n; i++)
	g->d2h_bytes += (long long) g->ctx[i]->d2h_bytes;
      rebalance(g, true);
    }
  else if(direct)
    {
      g2gpu_ctx *c = g->ctx[0];
      G2_TRY(g2_fetch_slice(c));
      g->lo[0] = c->w_lo;
      g->hi[0] = c->w_hi;
      g->ntargets = c->w_ntargets;
      int rc = g2gpu_download_acc(c, acc, cost, oldacc_out);
      c->compact = 1;
      if(rc)
	return rc;
      g->d2h_bytes = (long long) c->d2h_bytes;
    }
  else
    G2_TRY(g2gpu_group_download_acc(g, acc, cost, oldacc_out));
  if(direct)
    g->ctx[0]->compact = 1;
  if(perm)
    {
      // every device holds the same order: each one returns 1/N of it over its own link
      G2_TRY(run_all(g, [&](int i) {
	g2gpu_ctx *c = g->ctx[i];
	const size_t lo = (size_t) npart * i / g->n, hi = (size_t) npart * (i + 1) / g->n;
	if(hi > lo)
	  G2_CUDA(cudaMemcpyAsync(perm + lo, c->perm + lo, sizeof(int) * (hi - lo), cudaMemcpyDeviceToHost, c->stream));
	G2_CUDA(cudaStreamSynchronize(c->stream));
	return 0;
      }));
      g->d2h_bytes += (long long) sizeof(int) * npart;
    }
  return 0;
}

// One step with device-resident inputs (benchmarks: `value` of bench.py): all-gather of the resident shards -> domain -> treebuild ->
// walk, one host thread per device running its whole pipeline; results stay on the devices.
extern "C" int g2gpu_group_step_resident(g2gpu_group *g, int npart, const g2gpu_walk_params *wp)
{
  if(!g || !wp)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_TRY(group_begin_upload(g, npart));
  G2_TRY(run_all(g, [&](int i) {
    g2gpu_ctx *c = g->ctx[i];
    G2_TRY(g2gpu_inputs_ready(c, npart));
    c->compact = 1;
    G2_TRY(group_allgather(g, i, 0));
    G2_TRY(g2_stage_domain(c));
    G2_TRY(g2_stage_treebuild(c));
    c->slice_explicit = 1;
    c->slice_frac[0] = g->frac[i];
    c->slice_frac[1] = g->frac[i + 1];
    return g2_stage_walk(c, wp);
  }));
  g->gather_bytes = g->n > 1 ? (long long) g->per * g->n * (long long) sizeof(G2PRec) : 0;
  return 0;
}

// ms[8] / counters[8] like g2gpu_timings: times are the MAXIMUM over the devices, counters the SUM
extern "C" int g2gpu_group_timings(g2gpu_group *g, double ms[8], long long counters[8])
{
  if(!g)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  double m[G2_GROUP_MAX][8];
  long long cn[G2_GROUP_MAX][8];
  G2_TRY(run_all(g, [&](int i) { return g2gpu_timings(g->ctx[i], m[i], cn[i]); }));
  for(int k = 0; k < 8; k++)
    {
      if(ms)
	{
	  ms[k] = 0;
	  for(int i = 0; i < g->n; i++)
	    ms[k] = std::max(ms[k], m[i][k]);
	}
      if(counters)
	{
	  counters[k] = 0;
	  for(int i = 0; i < g->n; i++)
	    counters[k] += cn[i][k];
	}
    }
  return 0;
}

// out[0] host->device bytes (all devices), out[1] device->host bytes, out[2] bytes received per device by the all-gather;
// slices: lo/hi of every device's target slice of the last walk, fractions of the NEXT walk's slices
extern "C" int g2gpu_group_io_bytes(g2gpu_group *g, long long out[3])
{
  if(!g || !out)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  out[0] = g->h2d_bytes;
  out[1] = g->d2h_bytes;
  out[2] = g->gather_bytes;
  return 0;
}

extern "C" int g2gpu_group_slices(g2gpu_group *g, int *lo, int *hi, double *next_frac)
{
  if(!g)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  for(int i = 0; i < g->n; i++)
    {
      if(lo)
	lo[i] = g->lo[i];
      if(hi)
	hi[i] = g->hi[i];
    }
  if(next_frac)
    for(int i = 0; i <= g->n; i++)
      next_frac[i] = g->frac[i];
  return 0;
}
