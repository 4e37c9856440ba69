// g2_api.cu — the C ABI (include/g2gpu.h): context life cycle, uploads, downloads, stage entry points.
#include "g2_common.cuh"
#include <algorithm>
#include <thread>
#include <vector>
#include <stdarg.h>
#include <stdlib.h>

__thread char g2_errbuf[512] = "";

int g2_fail(int code, const char *fmt, ...)
{
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g2_errbuf, sizeof(g2_errbuf), fmt, ap);
  va_end(ap);
  return code;
}

extern "C" const char *g2gpu_last_error(void) { return g2_errbuf; }

extern "C" int g2gpu_device_count(void)
{
  int n = 0;
  if(cudaGetDeviceCount(&n) != cudaSuccess)
    return 0;
  return n;
}

template <typename T>
static int dalloc(T **p, size_t count)
{
  cudaError_t e = cudaMalloc((void **) p, sizeof(T) * (count ? count : 1));
  if(e != cudaSuccess)
    return g2_fail(G2GPU_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", sizeof(T) * count, cudaGetErrorString(e));
  return 0;
}

extern "C" int g2gpu_create(g2gpu_ctx **out, const g2gpu_config *cfg)
{
  if(!out || !cfg)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  *out = nullptr;
  if(cfg->n_gravs < 1 || cfg->n_gravs > G2GPU_MAX_GRAVS || cfg->max_part < 1 || cfg->max_nodes < 8)
    return g2_fail(G2GPU_ERR_ARG, "bad configuration (n_gravs=%d max_part=%d max_nodes=%d)", cfg->n_gravs, cfg->max_part, cfg->max_nodes);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if(e != cudaSuccess || ndev == 0)
    return g2_fail(G2GPU_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
  if(cfg->device < 0 || cfg->device >= ndev)
    return g2_fail(G2GPU_ERR_ARG, "device %d out of range (%d devices)", cfg->device, ndev);
  G2_CUDA(cudaSetDevice(cfg->device));
  cudaDeviceProp prop;
  G2_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
  if(prop.major < 10)
    return g2_fail(G2GPU_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);

  g2gpu_ctx *c = (g2gpu_ctx *) calloc(1, sizeof(g2gpu_ctx));
  if(!c)
    return g2_fail(G2GPU_ERR_NOMEM, "host allocation failed");
  c->cfg = *cfg;
  if(c->cfg.nranks < 1)
    c->cfg.nranks = 1;
  c->D = cfg->n_gravs;
  c->nsm = prop.multiProcessorCount > 0 ? prop.multiProcessorCount : G2_NSM_FALLBACK;
  c->acc_double = 1;
  c->walk_exact = getenv("G2GPU_WALK_EXACT") ? atoi(getenv("G2GPU_WALK_EXACT")) != 0 : 1;
  c->walk_defer = getenv("G2GPU_WALK_DEFER") ? atoi(getenv("G2GPU_WALK_DEFER")) != 0 : 0;	// only in a -DG2_WALK_DEFER build
  c->walk_carveout = getenv("G2GPU_WALK_CARVEOUT") ? atoi(getenv("G2GPU_WALK_CARVEOUT")) : -1;
  c->walk_sm_local = getenv("G2GPU_WALK_SM_LOCAL") ? atoi(getenv("G2GPU_WALK_SM_LOCAL")) != 0 : 1;
  // MEASURED (B200, 256^3, profiles/r2_exact_cost.txt): mask 0 / 1 / 3 / 7 -> walk 212.7 / 209.0 / 207.5 / 206.7 ms, median error against the reference
  // 1.38e-7 / - / 1.55e-7 / 1.79e-7, p99.9 and GravCost unchanged
  c->walk_flush_mask = getenv("G2GPU_WALK_FLUSH_MASK") ? atoi(getenv("G2GPU_WALK_FLUSH_MASK")) : 7;
  for(int t = 0; t < 6; t++)
    c->force_softening[t] = 1.0;
  bool ok = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) == cudaSuccess;
  for(int i = 0; i < 20 && ok; i++)
    ok = cudaEventCreate(&c->ev[i]) == cudaSuccess;
  if(!ok)
    {
      g2gpu_destroy(c);
      return g2_fail(G2GPU_ERR_CUDA, "stream / event creation failed: %s", cudaGetErrorString(cudaGetLastError()));
    }

  const size_t np = (size_t) cfg->max_part, nn = (size_t) cfg->max_nodes;
  const size_t cap = (np > nn ? np : nn) + G2_MAXTOP + 8;
  const int R = 2 + c->D;
  int rc = 0;
  rc |= dalloc(&c->in_rec, np + 64);	// + slack: the padded all-gather of the multi-GPU group may write up to 7 records past MaxPart
  rc |= dalloc(&c->in_raw, 7 * np);
  c->own_in_rec = c->in_rec;
  rc |= dalloc(&c->prec, np);
  rc |= dalloc(&c->phkey, np); rc |= dalloc(&c->perm, np); rc |= dalloc(&c->phorder, np);
  rc |= dalloc(&c->skey[0], np); rc |= dalloc(&c->skey[1], np); rc |= dalloc(&c->sval[0], np); rc |= dalloc(&c->sval[1], np);
  // radix-sort scratch: the one-sweep sort keeps a status word per (pass, tile, bin): 7 passes x tiles x 512 bins
  c->tilehist_elems = (size_t) 8 * 512 + 16 + (size_t) 7 * ((np + 2047) / 2048) * 512;	// (2048-pair tiles: G2GPU_SORT_ITEMS=8)
  c->sort_onesweep = getenv("G2GPU_SORT_ONESWEEP") ? atoi(getenv("G2GPU_SORT_ONESWEEP")) != 0 : 1;
  c->sort_items = getenv("G2GPU_SORT_ITEMS") ? atoi(getenv("G2GPU_SORT_ITEMS")) : 8;
  c->sort_rank_ballot = getenv("G2GPU_SORT_RANK_BALLOT") ? atoi(getenv("G2GPU_SORT_RANK_BALLOT")) != 0 : 0;
  c->sort_window = getenv("G2GPU_SORT_WINDOW") ? atoi(getenv("G2GPU_SORT_WINDOW")) : 4;
  rc |= dalloc(&c->tilehist, c->tilehist_elems);
  c->scan_tmp_elems = cap / 2048 + 16 + c->tilehist_elems / 2048;
  rc |= dalloc(&c->scan_tmp, c->scan_tmp_elems);
  rc |= dalloc(&c->d_domain, (size_t) 16); rc |= dalloc(&c->d_minmax, (size_t) 8); rc |= dalloc(&c->d_top, (size_t) 1);
  rc |= dalloc(&c->d_species_start, (size_t) 16);
  rc |= dalloc(&c->tm, np + 1); rc |= dalloc(&c->ttl, np + 1); rc |= dalloc(&c->tbase, np + 2); rc |= dalloc(&c->tcnt, cap + 2);
  rc |= dalloc(&c->c_a, nn); rc |= dalloc(&c->c_b, nn); rc |= dalloc(&c->c_d, nn); rc |= dalloc(&c->c_suns, nn * 8);
  rc |= dalloc(&c->c_father, nn); rc |= dalloc(&c->p_parent, np); rc |= dalloc(&c->c_ready, nn + 1); rc |= dalloc(&c->c_nchild, nn);
  rc |= dalloc(&c->c_npart, nn); rc |= dalloc(&c->c_min1, nn); rc |= dalloc(&c->c_min2, nn); rc |= dalloc(&c->c_poff, cap + 2);
  rc |= dalloc(&c->c_refid, nn);
  rc |= dalloc(&c->t_suns, (size_t) G2_MAXTOP * 8); rc |= dalloc(&c->t_first, (size_t) G2_MAXTOP); rc |= dalloc(&c->t_last, (size_t) G2_MAXTOP);
  rc |= dalloc(&c->t_min1, (size_t) G2_MAXTOP); rc |= dalloc(&c->t_min2, (size_t) G2_MAXTOP); rc |= dalloc(&c->t_ready, (size_t) G2_MAXTOP);
  rc |= dalloc(&c->t_npart, (size_t) G2_MAXTOP); rc |= dalloc(&c->t_nchild, (size_t) G2_MAXTOP); rc |= dalloc(&c->t_ubase, (size_t) G2_MAXTOP);
  rc |= dalloc(&c->wcells, (nn + G2_MAXTOP) * R); rc |= dalloc(&c->wpart, np); rc |= dalloc(&c->wsrc, np);
  rc |= dalloc(&c->hist2, np + 2); rc |= dalloc(&c->hist2_scan, np + 2); rc |= dalloc(&c->dmin, np + 2);
  rc |= dalloc(&c->d_err, (size_t) 8);
  rc |= dalloc(&c->d_depth, (size_t) 64);
  rc |= dalloc(&c->w_targets, np); rc |= dalloc(&c->w_flags, np + 2);
  rc |= dalloc(&c->acc, 3 * np); rc |= dalloc(&c->cost, np); rc |= dalloc(&c->oldacc_out, np);
  rc |= dalloc(&c->d_counters, (size_t) 8); rc |= dalloc(&c->d_slice, (size_t) 4); rc |= dalloc((char **) &c->d_exact, (size_t) 256); rc |= dalloc(&c->d_smcount, (size_t) 1024);
  if(rc)
    {
      g2gpu_destroy(c);
      return G2GPU_ERR_NOMEM;
    }
  if(cudaMallocHost((void **) &c->h_err, 64 * sizeof(int)) != cudaSuccess || cudaMallocHost((void **) &c->h_counters, 8 * sizeof(unsigned long long)) != cudaSuccess
     || cudaMallocHost((void **) &c->h_top, sizeof(G2TopTree)) != cudaSuccess || cudaMallocHost((void **) &c->h_slice, 4 * sizeof(int)) != cudaSuccess
     || cudaMallocHost((void **) &c->h_domain, 8 * sizeof(double)) != cudaSuccess)
    {
      g2gpu_destroy(c);
      return g2_fail(G2GPU_ERR_NOMEM, "pinned host allocation failed");
    }
  if(cudaMemset(c->acc, 0, sizeof(float) * 3 * np) != cudaSuccess || cudaMemset(c->cost, 0, sizeof(float) * np) != cudaSuccess
     || cudaMemset(c->oldacc_out, 0, sizeof(float) * np) != cudaSuccess || cudaMemset(c->d_slice, 0, 4 * sizeof(int)) != cudaSuccess)
    {
      g2gpu_destroy(c);
      return g2_fail(G2GPU_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
  *out = c;
  return 0;
}

extern "C" void g2gpu_destroy(g2gpu_ctx *c)
{
  if(!c)
    return;
  cudaSetDevice(c->cfg.device);
  if(c->stream)
    cudaStreamSynchronize(c->stream);
  void *ptrs[] = { c->own_in_rec, c->in_raw, c->in_vel, c->in_gravpm, c->prec, c->vel, c->gravpm,
    c->phkey, c->perm, c->skey[0], c->skey[1], c->sval[0], c->sval[1], c->tilehist, c->scan_tmp, c->d_domain, c->d_minmax, c->d_top, c->d_topscratch,
    c->d_species_start, c->tm, c->ttl, c->tbase, c->tcnt, c->c_a, c->c_b, c->c_d, c->c_suns, c->c_father, c->p_parent, c->c_ready, c->c_nchild, c->c_npart,
    c->c_min1, c->c_min2, c->c_poff, c->c_refid, c->t_suns, c->t_first, c->t_last, c->t_min1, c->t_min2, c->t_ready, c->t_npart, c->t_nchild, c->t_ubase,
    c->wcells, c->wpart, c->hist2, c->hist2_scan, c->dmin, c->d_err, c->d_depth, c->w_targets, c->w_flags, c->acc, c->cost, c->oldacc_out, c->d_counters,
    c->d_srtable, c->d_srtable_f, c->d_pottable_f, c->pot, c->d_lattice, c->d_potcorr, c->latt, c->lattcost, c->wcnt, c->wsrc, c->phorder, c->d_slice, c->d_exact, c->cres, c->d_smcount };
  for(size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++)
    if(ptrs[i])
      cudaFree(ptrs[i]);
  g2_pm_destroy(c);
  if(c->h_err)
    cudaFreeHost(c->h_err);
  if(c->h_counters)
    cudaFreeHost(c->h_counters);
  if(c->h_top)
    cudaFreeHost(c->h_top);
  if(c->h_slice)
    cudaFreeHost(c->h_slice);
  if(c->h_domain)
    cudaFreeHost(c->h_domain);
  if(c->h_stage)
    cudaFreeHost(c->h_stage);
  if(c->h_rec)
    cudaFreeHost(c->h_rec);
  if(c->h_vel)
    cudaFreeHost(c->h_vel);
  if(c->h_gravpm)
    cudaFreeHost(c->h_gravpm);
  if(c->h_export)
    cudaFreeHost(c->h_export);
  if(c->d_export)
    cudaFree(c->d_export);
  for(int i = 0; i < 20; i++)
    if(c->ev[i])
      cudaEventDestroy(c->ev[i]);
  if(c->stream)
    cudaStreamDestroy(c->stream);
  free(c);
}

extern "C" int g2gpu_set_species(g2gpu_ctx *c, const int type_to_grav[6], const double force_softening[6])
{
  if(!c || !type_to_grav || !force_softening)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  for(int t = 0; t < 6; t++)
    {
      if(type_to_grav[t] < 0 || type_to_grav[t] >= c->D)	// ngravs_core.c:270-279
	return g2_fail(G2GPU_ERR_ARG, "native interaction %d declared for type %d does not exist", type_to_grav[t], t);
      c->type_to_grav[t] = type_to_grav[t];
      c->force_softening[t] = force_softening[t];
    }
  return 0;
}

extern "C" int g2gpu_set_laws(g2gpu_ctx *c, const int *accel_id, const int *spline_id, const double *params)
{
  if(!c || !accel_id || !spline_id)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  const int D = c->D;
  for(int i = 0; i < D * D; i++)
    {
      int a = accel_id[i], s = spline_id[i];
      if(a < G2GPU_LAW_NONE || a > G2GPU_LAW_SOURCEBARYONBAM)	// ngravs_core.c:326-365: every slot must be wired
	return g2_fail(G2GPU_ERR_LAW, "AccelFxns[%d][%d] is not a registered law (%d)", i / D, i % D, a);
      if(s < G2GPU_SPLINE_NONE || s > G2GPU_SPLINE_SOURCEBARYONBAM)
	return g2_fail(G2GPU_ERR_LAW, "AccelSplines[%d][%d] is not a registered spline (%d)", i / D, i % D, s);
      c->laws.accel[i] = a;
      c->laws.spline[i] = s;
      for(int k = 0; k < 4; k++)
	c->laws.par[i][k] = params ? (float) params[4 * i + k] : 0.0f;
    }
  c->laws_set = 1;
  return 0;
}

extern "C" int g2gpu_set_srtable(g2gpu_ctx *c, const double *table)
{
  if(!c || !table)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  const int D = c->D, ntab = c->cfg.ntab;
  if(ntab <= 0)
    return g2_fail(G2GPU_ERR_ARG, "ntab not configured");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  // identical pair tables are stored once (stock wiring: all D*D tables are the Newtonian one)
  int nu = 0;
  int first[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];
  unsigned char tabmap[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];
  memset(tabmap, 0, sizeof(tabmap));
  for(int i = 0; i < D * D; i++)
    {
      int found = -1;
      for(int u = 0; u < nu && found < 0; u++)
	if(memcmp(table + (size_t) first[u] * ntab, table + (size_t) i * ntab, sizeof(double) * ntab) == 0)
	  found = u;
      if(found < 0)
	{
	  first[nu] = i;
	  found = nu++;
	}
      tabmap[i] = (unsigned char) found;
    }
  // checked before any state changes: a refused call leaves the previous tables in place
  if((size_t) nu * ntab * sizeof(float) > 200 * 1024)
    return g2_fail(G2GPU_ERR_ARG, "%d distinct short-range tables do not fit in shared memory", nu);
  memcpy(c->sr_tabmap, tabmap, sizeof(tabmap));
  c->srtable_set = 0;
  c->sr_ntables = nu;
  float *hf = (float *) malloc(sizeof(float) * (size_t) nu * ntab);
  if(!hf)
    return g2_fail(G2GPU_ERR_NOMEM, "host allocation failed");
  for(int u = 0; u < nu; u++)
    for(int k = 0; k < ntab; k++)
      hf[(size_t) u * ntab + k] = (float) table[(size_t) first[u] * ntab + k];
  if(c->d_srtable_f)
    cudaFree(c->d_srtable_f);
  c->d_srtable_f = nullptr;
  int rc = dalloc(&c->d_srtable_f, (size_t) nu * ntab);
  if(rc == 0 && cudaMemcpy(c->d_srtable_f, hf, sizeof(float) * (size_t) nu * ntab, cudaMemcpyHostToDevice) != cudaSuccess)
    rc = g2_fail(G2GPU_ERR_CUDA, "table upload failed");
  free(hf);
  if(rc)
    return rc;
  c->srtable_set = 1;
  return 0;
}

// ---- uploads --------------------------------------------------------------------------------------------------
static int ensure_opt(g2gpu_ctx *c, int want_vel, int want_gravpm)
{
  const size_t np = (size_t) c->cfg.max_part;
  if(want_vel && !c->in_vel)
    G2_TRY(dalloc(&c->in_vel, 3 * np));
  if(want_vel && !c->vel)
    G2_TRY(dalloc(&c->vel, 3 * np));
  if(want_gravpm && !c->in_gravpm)
    G2_TRY(dalloc(&c->in_gravpm, 3 * (np + 64)));
  if(want_gravpm && !c->gravpm)
    G2_TRY(dalloc(&c->gravpm, 3 * np));
  return 0;
}

int g2_ensure_optional_inputs(g2gpu_ctx *c, int want_vel, int want_gravpm) { return ensure_opt(c, want_vel, want_gravpm); }

// SoA arrays go to the device as they are (no host-side repacking); pack_inputs_kernel builds the 32-byte records.
__global__ void __launch_bounds__(256) pack_inputs_kernel(int n, const float *__restrict__ pos3, const float *__restrict__ mass, const int *__restrict__ type,
							   const float *__restrict__ oldacc, const int *__restrict__ active_i, G2PRec *__restrict__ rec)
{
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if(i >= n)
    return;
  G2PRec r;
  r.x = pos3[3 * (size_t) i];
  r.y = pos3[3 * (size_t) i + 1];
  r.z = pos3[3 * (size_t) i + 2];
  r.m = mass[i];
  r.type = type[i];
  r.oldacc = oldacc ? oldacc[i] : 0.0f;
  r.active = active_i ? (active_i[i] != 0) : 1;
  r.pad = 0;
  rec[i] = r;
}

// records [lo, lo + cnt) of an n_total-particle set: H2D of the SoA pieces + pack into own_in_rec[lo ...] (the multi-GPU group uploads
// one shard per device and all-gathers the rest over NVLink, g2_group.cu)
int g2_upload_soa_shard(g2gpu_ctx *c, int n_total, int lo, int cnt, const float *pos, const float *mass, const int *type, const float *oldacc,
			const int *active)
{
  if(n_total < 1 || n_total > c->cfg.max_part || lo < 0 || cnt < 0 || lo + cnt > n_total)
    return g2_fail(G2GPU_ERR_ARG, "npart %d outside [1, MaxPart=%d] (shard %d+%d)", n_total, c->cfg.max_part, lo, cnt);
  c->in_rec = c->own_in_rec;	// back to the library's own input buffer
  c->inputs_bound = 0;
  const size_t n = (size_t) cnt;
  cudaStream_t st = c->stream;
  float *d_pos3 = c->in_raw, *d_mass = c->in_raw + 3 * n, *d_old = c->in_raw + 5 * n;
  int *d_type = (int *) (c->in_raw + 4 * n), *d_act = (int *) (c->in_raw + 6 * n);
  G2_CUDA(cudaEventRecord(c->ev[9], st));
  if(cnt > 0)
    {
      G2_CUDA(cudaMemcpyAsync(d_pos3, pos + 3 * (size_t) lo, n * 12, cudaMemcpyHostToDevice, st));
      G2_CUDA(cudaMemcpyAsync(d_mass, mass + lo, n * 4, cudaMemcpyHostToDevice, st));
      G2_CUDA(cudaMemcpyAsync(d_type, type + lo, n * 4, cudaMemcpyHostToDevice, st));
      if(oldacc)
	G2_CUDA(cudaMemcpyAsync(d_old, oldacc + lo, n * 4, cudaMemcpyHostToDevice, st));
      if(active)
	G2_CUDA(cudaMemcpyAsync(d_act, active + lo, n * 4, cudaMemcpyHostToDevice, st));
      pack_inputs_kernel<<<g2_cdiv(cnt, 256), 256, 0, st>>>(cnt, d_pos3, d_mass, d_type, oldacc ? d_old : nullptr, active ? d_act : nullptr, c->own_in_rec + lo);
      c->launches++;
    }
  G2_CUDA(cudaEventRecord(c->ev[10], st));
  G2_CUDA(cudaGetLastError());
  c->h2d_bytes = n * (12 + 4 + 4 + (oldacc ? 4 : 0) + (active ? 4 : 0));
  c->have_vel = 0;
  c->have_gravpm = 0;
  c->pm_done = 0;
  c->npart = n_total;
  c->stage = 1;
  return 0;
}

extern "C" int g2gpu_upload(g2gpu_ctx *c, int npart, const float *pos, const float *mass, const int *type, const float *oldacc,
			    const float *vel, const float *gravpm, const int *active)
{
  if(!c || !pos || !mass || !type)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(npart < 1 || npart > c->cfg.max_part)
    return g2_fail(G2GPU_ERR_ARG, "npart %d outside [1, MaxPart=%d]", npart, c->cfg.max_part);
  G2_CUDA(cudaSetDevice(c->cfg.device));
  G2_TRY(ensure_opt(c, vel != nullptr, gravpm != nullptr));
  G2_TRY(g2_upload_soa_shard(c, npart, 0, npart, pos, mass, type, oldacc, active));
  const size_t n = (size_t) npart;
  cudaStream_t st = c->stream;
  if(vel)
    G2_CUDA(cudaMemcpyAsync(c->in_vel, vel, n * 12, cudaMemcpyHostToDevice, st));
  if(gravpm)
    G2_CUDA(cudaMemcpyAsync(c->in_gravpm, gravpm, n * 12, cudaMemcpyHostToDevice, st));
  G2_CUDA(cudaEventRecord(c->ev[10], st));
  c->h2d_bytes += n * ((vel ? 12 : 0) + (gravpm ? 12 : 0));
  c->have_vel = vel != nullptr;
  c->have_gravpm = gravpm != nullptr;
  return 0;
}

// The reference's AoS goes through ONE pass on the host (several threads) into a pinned staging area that already has the
// device layout -- 32-byte particle records (+ optional velocity / GravPM triples) -- and then to the device with one copy each.
// Records [lo, lo + cnt) of an n_total-particle set (a shard per device in the multi-GPU group; the whole set otherwise).
int g2_upload_aos_shard(g2gpu_ctx *c, int n_total, int lo, int cnt, const void *P, size_t stride, int float_bytes, int off_pos, int off_mass,
			int off_type, int off_oldacc, int off_vel, int off_gravpm, int off_ti_endstep, int ti_current, unsigned int max_threads)
{
  if(n_total < 1 || n_total > c->cfg.max_part || lo < 0 || cnt < 0 || lo + cnt > n_total)
    return g2_fail(G2GPU_ERR_ARG, "npart %d outside [1, MaxPart=%d] (shard %d+%d)", n_total, c->cfg.max_part, lo, cnt);
  if(float_bytes != 4 && float_bytes != 8)
    return g2_fail(G2GPU_ERR_ARG, "float_bytes must be 4 or 8");
  G2_TRY(ensure_opt(c, off_vel >= 0, off_gravpm >= 0));
  const size_t n = (size_t) cnt, np = (size_t) c->cfg.max_part;
  if(!c->h_rec)
    G2_CUDA(cudaMallocHost((void **) &c->h_rec, sizeof(G2PRec) * np));
  if(off_vel >= 0 && !c->h_vel)
    G2_CUDA(cudaMallocHost((void **) &c->h_vel, sizeof(float) * 3 * np));
  if(off_gravpm >= 0 && !c->h_gravpm)
    G2_CUDA(cudaMallocHost((void **) &c->h_gravpm, sizeof(float) * 3 * np));
  cudaStream_t st = c->stream;
  G2_CUDA(cudaStreamSynchronize(st));	// the staging area may still be the source of the previous upload
  const char *base = (const char *) P + (size_t) lo * stride;
  G2PRec *hrec = c->h_rec;
  float *hvel = c->h_vel, *hgpm = c->h_gravpm;
  auto convert = [=](size_t a, size_t b) {
#define G2_RD(off, k) (float_bytes == 4 ? ((const float *) (q + (off)))[k] : (float) ((const double *) (q + (off)))[k])
    for(size_t i = a; i < b; i++)
      {
	const char *q = base + i * stride;
	G2PRec r;
	r.x = G2_RD(off_pos, 0);
	r.y = G2_RD(off_pos, 1);
	r.z = G2_RD(off_pos, 2);
	r.m = G2_RD(off_mass, 0);
	r.type = *(const int *) (q + off_type);
	r.oldacc = off_oldacc >= 0 ? G2_RD(off_oldacc, 0) : 0.0f;
	r.active = off_ti_endstep >= 0 ? (*(const int *) (q + off_ti_endstep) == ti_current) : 1;
	r.pad = 0;
	hrec[i] = r;
	if(off_vel >= 0)
	  for(int k = 0; k < 3; k++)
	    hvel[3 * i + k] = G2_RD(off_vel, k);
	if(off_gravpm >= 0)
	  for(int k = 0; k < 3; k++)
	    hgpm[3 * i + k] = G2_RD(off_gravpm, k);
      }
#undef G2_RD
  };
  unsigned int nthr = std::thread::hardware_concurrency();
  nthr = std::max(1u, std::min(std::min(nthr, max_threads), (unsigned int) (n / 262144 + 1)));
  if(nthr == 1)
    convert(0, n);
  else
    {
      std::vector<std::thread> pool;
      for(unsigned int t = 0; t < nthr; t++)
	pool.emplace_back(convert, n * t / nthr, n * (t + 1) / nthr);
      for(auto &t : pool)
	t.join();
    }
  c->in_rec = c->own_in_rec;
  c->inputs_bound = 0;
  G2_CUDA(cudaEventRecord(c->ev[9], st));
  if(cnt > 0)
    {
      G2_CUDA(cudaMemcpyAsync(c->own_in_rec + lo, hrec, sizeof(G2PRec) * n, cudaMemcpyHostToDevice, st));
      if(off_vel >= 0)
	G2_CUDA(cudaMemcpyAsync(c->in_vel + 3 * (size_t) lo, hvel, n * 12, cudaMemcpyHostToDevice, st));
      if(off_gravpm >= 0)
	G2_CUDA(cudaMemcpyAsync(c->in_gravpm + 3 * (size_t) lo, hgpm, n * 12, cudaMemcpyHostToDevice, st));
    }
  G2_CUDA(cudaEventRecord(c->ev[10], st));
  c->h2d_bytes = n * (sizeof(G2PRec) + (off_vel >= 0 ? 12 : 0) + (off_gravpm >= 0 ? 12 : 0));
  c->have_vel = off_vel >= 0;
  c->have_gravpm = off_gravpm >= 0;
  c->pm_done = 0;
  c->npart = n_total;
  c->stage = 1;
  return 0;
}

extern "C" int g2gpu_upload_aos(g2gpu_ctx *c, int npart, const void *P, size_t stride, int float_bytes, int off_pos, int off_mass,
				int off_type, int off_oldacc, int off_vel, int off_gravpm, int off_ti_endstep, int ti_current)
{
  if(!c || !P)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_upload_aos_shard(c, npart, 0, npart, P, stride, float_bytes, off_pos, off_mass, off_type, off_oldacc, off_vel, off_gravpm, off_ti_endstep,
			     ti_current, 16u);
}

extern "C" int g2gpu_input_buffers(g2gpu_ctx *c, int npart, void **records)
{
  if(!c || !records)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(npart < 1 || npart > c->cfg.max_part)
    return g2_fail(G2GPU_ERR_ARG, "npart %d outside [1, MaxPart=%d]", npart, c->cfg.max_part);
  *records = c->own_in_rec;
  return 0;
}

extern "C" int g2gpu_bind_inputs(g2gpu_ctx *c, int npart, void *records)
{
  if(!c || !records)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(npart < 1 || npart > c->cfg.max_part)
    return g2_fail(G2GPU_ERR_ARG, "npart %d outside [1, MaxPart=%d]", npart, c->cfg.max_part);
  if(((size_t) records & 15) != 0)
    return g2_fail(G2GPU_ERR_ARG, "particle records must be 16-byte aligned");
  c->in_rec = (G2PRec *) records;
  c->inputs_bound = 1;
  c->have_vel = 0;
  c->have_gravpm = 0;
  c->pm_done = 0;
  c->npart = npart;
  c->stage = 1;
  return 0;
}

extern "C" int g2gpu_inputs_ready(g2gpu_ctx *c, int npart)
{
  if(!c)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(npart < 1 || npart > c->cfg.max_part)
    return g2_fail(G2GPU_ERR_ARG, "npart %d outside [1, MaxPart=%d]", npart, c->cfg.max_part);
  c->in_rec = c->own_in_rec;
  c->inputs_bound = 0;
  c->have_vel = 0;
  c->have_gravpm = 0;
  c->pm_done = 0;
  c->npart = npart;
  c->stage = 1;
  return 0;
}

// ---- stages ---------------------------------------------------------------------------------------------------
static float ev_ms(g2gpu_ctx *c, int a, int b)
{
  float ms = 0;
  if(cudaEventElapsedTime(&ms, c->ev[a], c->ev[b]) != cudaSuccess)
    {
      cudaGetLastError();
      return 0;
    }
  return ms;
}

extern "C" int g2gpu_domain(g2gpu_ctx *c)
{
  if(!c)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(c->stage < 1)
    return g2_fail(G2GPU_ERR_STATE, "domain: no particles uploaded");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_stage_domain(c);
}

extern "C" int g2gpu_treebuild(g2gpu_ctx *c, int *numnodes)
{
  if(!c)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  G2_TRY(g2_stage_treebuild(c));
  if(numnodes)
    *numnodes = c->numnodes;
  return 0;
}

extern "C" int g2gpu_walk(g2gpu_ctx *c, const g2gpu_walk_params *wp)
{
  if(!c || !wp)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_stage_walk(c, wp);
}

extern "C" int g2gpu_direct(g2gpu_ctx *c, const g2gpu_walk_params *wp, int ntargets, const int *targets, double *acc)
{
  if(!c || !wp || !targets || !acc || ntargets < 1)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_direct_sum(c, wp, ntargets, targets, acc);
}

extern "C" int g2gpu_sync(g2gpu_ctx *c)
{
  if(!c)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" void *g2gpu_stream(g2gpu_ctx *c) { return c ? (void *) c->stream : nullptr; }

extern "C" int g2gpu_update_tree(g2gpu_ctx *c, const float *len, const float *s)
{
  if(!c || !len || !s)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_update_tree(c, len, s);
}

extern "C" int g2gpu_pm_periodic(g2gpu_ctx *c, const g2gpu_pm_params *pp)
{
  if(!c || !pp)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_pm_periodic(c, pp);
}

extern "C" int g2gpu_pm_potential_periodic(g2gpu_ctx *c, const g2gpu_pm_params *pp, float *potential)
{
  if(!c || !pp || !potential)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_pm_potential_periodic(c, pp, potential);
}

extern "C" int g2gpu_download_gravpm(g2gpu_ctx *c, float *gravpm)
{
  if(!c || !gravpm)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_pm_download(c, gravpm);
}

extern "C" int g2gpu_set_option(g2gpu_ctx *c, const char *name, int value)
{
  if(!c || !name)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(strcmp(name, "acc_double") == 0)
    c->acc_double = value;
  else if(strcmp(name, "accumulator") == 0)
    c->accumulator = value != 0;	// takes effect at the next g2gpu_treebuild
  else if(strcmp(name, "direct_ewald") == 0)
    c->direct_ewald = value != 0;
  else if(strcmp(name, "walk_stats") == 0)
    c->walk_stats = value != 0;
  else if(strcmp(name, "sort_onesweep") == 0)
    c->sort_onesweep = value != 0;
  else if(strcmp(name, "walk_exact") == 0)
    c->walk_exact = value != 0;
  else if(strcmp(name, "walk_defer") == 0)
    c->walk_defer = value != 0;
  else if(strcmp(name, "compact") == 0)
    {
      if(value && !c->cres)
	{
	  G2_CUDA(cudaSetDevice(c->cfg.device));
	  G2_TRY(dalloc(&c->cres, 5 * (size_t) c->cfg.max_part));
	}
      c->compact = value != 0;
    }
  else if(strcmp(name, "rank") == 0)
    {
      c->cfg.rank = value;
      c->slice_explicit = 0;
    }
  else if(strcmp(name, "nranks") == 0)
    {
      c->cfg.nranks = value < 1 ? 1 : value;
      c->slice_explicit = 0;
    }
  else
    return g2_fail(G2GPU_ERR_ARG, "unknown option %s", name);
  return 0;
}

// ---- downloads --------------------------------------------------------------------------------------------------
extern "C" int g2gpu_get_domain(g2gpu_ctx *c, double out[8])
{
  if(!c || !out)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(c->stage < 2)
    return g2_fail(G2GPU_ERR_STATE, "g2gpu_domain has not run");
  G2_CUDA(cudaMemcpyAsync(out, c->d_domain, 8 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  G2_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int g2gpu_get_keys(g2gpu_ctx *c, long long *keys)
{
  if(!c || !keys)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(c->stage < 2)
    return g2_fail(G2GPU_ERR_STATE, "g2gpu_domain has not run");
  G2_CUDA(cudaMemcpyAsync(keys, c->phkey, sizeof(long long) * (size_t) c->npart, cudaMemcpyDeviceToHost, c->stream));
  G2_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int g2gpu_get_order(g2gpu_ctx *c, int *perm)
{
  if(!c || !perm)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(c->stage < 2)
    return g2_fail(G2GPU_ERR_STATE, "g2gpu_domain has not run");
  G2_CUDA(cudaMemcpyAsync(perm, c->perm, sizeof(int) * (size_t) c->npart, cudaMemcpyDeviceToHost, c->stream));
  G2_CUDA(cudaStreamSynchronize(c->stream));
  return 0;
}

extern "C" int g2gpu_get_topnodes(g2gpu_ctx *c, int *ntopnodes, int *ntopleaves, int *daughter, int *leaf, long long *startkey,
				  long long *size, long long *count, int *domain_node_index)
{
  if(!c)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(c->stage < 2)
    return g2_fail(G2GPU_ERR_STATE, "g2gpu_domain has not run");
  G2_CUDA(cudaMemcpyAsync(c->h_top, c->d_top, sizeof(G2TopTree), cudaMemcpyDeviceToHost, c->stream));
  G2_CUDA(cudaStreamSynchronize(c->stream));
  const G2TopTree *t = c->h_top;
  if(t->err)
    return g2_fail(G2GPU_ERR_TOPNODES, "top-level tree exceeds %d nodes", G2_MAXTOP);
  if(ntopnodes)
    *ntopnodes = t->ntopnodes;
  if(ntopleaves)
    *ntopleaves = t->ntopleaves;
  for(int i = 0; i < t->ntopnodes; i++)
    {
      if(daughter)
	daughter[i] = t->daughter[i];
      if(leaf)
	leaf[i] = t->leaf[i];
      if(startkey)
	startkey[i] = t->startkey[i];
      if(size)
	size[i] = 1LL << t->shift[i];
      if(count)
	count[i] = t->count[i];
    }
  if(domain_node_index)
    for(int i = 0; i < t->ntopleaves; i++)
      domain_node_index[i] = c->cfg.max_part + t->dni[i];
  return 0;
}

extern "C" int g2gpu_download_tree(g2gpu_ctx *c, float *len, float *center, float *s, float *mass, int *bitflags, int *sibling,
				   int *nextnode, int *father, int *p_nextnode, int *p_father)
{
  if(!c)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_export_tree(c, len, center, s, mass, bitflags, sibling, nextnode, father, p_nextnode, p_father);
}

extern "C" int g2gpu_download_extnodes(g2gpu_ctx *c, float *vs)
{
  if(!c || !vs)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_export_extnodes(c, vs);
}

extern "C" int g2gpu_download_nparticles(g2gpu_ctx *c, long long *nparticles)
{
  if(!c || !nparticles)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_export_nparticles(c, nparticles);
}

extern "C" int g2gpu_download_acc(g2gpu_ctx *c, float *acc, float *cost, float *oldacc)
{
  if(!c)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(c->stage < 4)
    return g2_fail(G2GPU_ERR_STATE, "g2gpu_walk has not run");
  const size_t n = (size_t) c->npart;
  cudaStream_t st = c->stream;
  G2_CUDA(cudaEventRecord(c->ev[11], st));
  if(acc)
    G2_CUDA(cudaMemcpyAsync(acc, c->acc, sizeof(float) * 3 * n, cudaMemcpyDeviceToHost, st));
  if(cost)
    G2_CUDA(cudaMemcpyAsync(cost, c->cost, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
  if(oldacc)
    G2_CUDA(cudaMemcpyAsync(oldacc, c->oldacc_out, sizeof(float) * n, cudaMemcpyDeviceToHost, st));
  G2_CUDA(cudaEventRecord(c->ev[12], st));
  G2_CUDA(cudaStreamSynchronize(st));
  c->d2h_bytes = n * ((acc ? 12 : 0) + (cost ? 4 : 0) + (oldacc ? 4 : 0));
  return 0;
}

// ---- potential walk (g2_pot.cu) ---------------------------------------------------------------------------------
extern "C" int g2gpu_set_potential_laws(g2gpu_ctx *c, const int *pot_id, const int *potspline_id)
{
  if(!c || !pot_id || !potspline_id)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  const int D = c->D;
  for(int i = 0; i < D * D; i++)
    {
      if(pot_id[i] < G2GPU_POT_NONE || pot_id[i] > G2GPU_POT_SOURCEBAMBARYON)	// ngravs_core.c:348-358: every slot must be wired
	return g2_fail(G2GPU_ERR_LAW, "PotentialFxns[%d][%d] is not a registered potential (%d)", i / D, i % D, pot_id[i]);
      if(potspline_id[i] < G2GPU_POTSPLINE_NONE || potspline_id[i] > G2GPU_POTSPLINE_SOURCEBAMBARYON)
	return g2_fail(G2GPU_ERR_LAW, "PotentialSplines[%d][%d] is not a registered potential spline (%d)", i / D, i % D, potspline_id[i]);
      c->potfxn[i] = pot_id[i];
      c->potspline[i] = potspline_id[i];
    }
  c->potlaws_set = 1;
  return 0;
}

extern "C" int g2gpu_set_srpot_table(g2gpu_ctx *c, const double *table)
{
  if(!c || !table)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  const int D = c->D, ntab = c->cfg.ntab;
  if(ntab <= 0)
    return g2_fail(G2GPU_ERR_ARG, "ntab not configured");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  int nu = 0;
  int first[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];
  for(int i = 0; i < D * D; i++)
    {				// identical pair tables are stored once
      int found = -1;
      for(int u = 0; u < nu && found < 0; u++)
	if(memcmp(table + (size_t) first[u] * ntab, table + (size_t) i * ntab, sizeof(double) * ntab) == 0)
	  found = u;
      if(found < 0)
	{
	  first[nu] = i;
	  found = nu++;
	}
      c->pot_tabmap[i] = (unsigned char) found;
    }
  if((size_t) nu * ntab * sizeof(float) > 200 * 1024)
    return g2_fail(G2GPU_ERR_ARG, "%d distinct short-range potential tables do not fit in shared memory", nu);
  float *hf = (float *) malloc(sizeof(float) * (size_t) nu * ntab);
  if(!hf)
    return g2_fail(G2GPU_ERR_NOMEM, "host allocation failed");
  for(int u = 0; u < nu; u++)
    for(int k = 0; k < ntab; k++)
      hf[(size_t) u * ntab + k] = (float) table[(size_t) first[u] * ntab + k];
  if(c->d_pottable_f)
    cudaFree(c->d_pottable_f);
  c->d_pottable_f = nullptr;
  int rc = dalloc(&c->d_pottable_f, (size_t) nu * ntab);
  if(rc == 0 && cudaMemcpy(c->d_pottable_f, hf, sizeof(float) * (size_t) nu * ntab, cudaMemcpyHostToDevice) != cudaSuccess)
    rc = g2_fail(G2GPU_ERR_CUDA, "table upload failed");
  free(hf);
  if(rc)
    return rc;
  c->pot_ntables = nu;
  c->pottable_set = 1;
  return 0;
}

extern "C" int g2gpu_potential(g2gpu_ctx *c, const g2gpu_walk_params *wp)
{
  if(!c || !wp)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_stage_potential(c, wp);
}

extern "C" int g2gpu_download_potential(g2gpu_ctx *c, float *pot, double *kernel_ms)
{
  if(!c || !pot)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  if(!c->pot_valid)
    return g2_fail(G2GPU_ERR_STATE, "g2gpu_potential has not run");
  G2_CUDA(cudaMemcpyAsync(pot, c->pot, sizeof(float) * (size_t) c->npart, cudaMemcpyDeviceToHost, c->stream));
  G2_CUDA(cudaStreamSynchronize(c->stream));
  if(kernel_ms)
    *kernel_ms = c->pot_ms;
  return 0;
}

// ---- lattice-sum correction (g2_lattice.cu) ----------------------------------------------------------------------
extern "C" int g2gpu_set_lattice_tables(g2gpu_ctx *c, int en, const double *fcorr)
{
  if(!c)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  if(!fcorr)
    {				// back to the nearest-image force
      c->lattice_set = 0;
      return 0;
    }
  if(!c->cfg.periodic || c->cfg.shortrange)
    return g2_fail(G2GPU_ERR_ARG, "lattice tables belong to a periodic box without PM (PERIODIC && !PMGRID)");
  if(en < 1 || en > 256)
    return g2_fail(G2GPU_ERR_ARG, "EN must be in [1, 256]");
  const int D = c->D;
  const size_t n3 = (size_t) (en + 1) * (en + 1) * (en + 1);
  // fcorr[comp][tgt][src][n3]; identical pair tables (all three components) are stored once
  int nu = 0, first[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];
  for(int i = 0; i < D * D; i++)
    {
      int found = -1;
      for(int u = 0; u < nu && found < 0; u++)
	{
	  bool same = true;
	  for(int cc = 0; cc < 3 && same; cc++)
	    same = memcmp(fcorr + ((size_t) cc * D * D + first[u]) * n3, fcorr + ((size_t) cc * D * D + i) * n3, sizeof(double) * n3) == 0;
	  if(same)
	    found = u;
	}
      if(found < 0)
	{
	  first[nu] = i;
	  found = nu++;
	}
      c->lattice_tabmap[i] = (unsigned char) found;
    }
  float *hf = (float *) malloc(sizeof(float) * 4 * n3 * (size_t) nu);
  if(!hf)
    return g2_fail(G2GPU_ERR_NOMEM, "host allocation failed");
  for(int u = 0; u < nu; u++)
    for(size_t q = 0; q < n3; q++)
      {
	for(int cc = 0; cc < 3; cc++)
	  hf[4 * ((size_t) u * n3 + q) + cc] = (float) fcorr[((size_t) cc * D * D + first[u]) * n3 + q];
	hf[4 * ((size_t) u * n3 + q) + 3] = 0.0f;
      }
  if(c->d_lattice)
    cudaFree(c->d_lattice);
  c->d_lattice = nullptr;
  int rc = dalloc(&c->d_lattice, 4 * n3 * (size_t) nu);
  if(rc == 0 && cudaMemcpy(c->d_lattice, hf, sizeof(float) * 4 * n3 * (size_t) nu, cudaMemcpyHostToDevice) != cudaSuccess)
    rc = g2_fail(G2GPU_ERR_CUDA, "table upload failed");
  free(hf);
  if(rc)
    return rc;
  c->lattice_en = en;
  c->lattice_ntables = nu;
  c->lattice_set = 1;
  return 0;
}

extern "C" int g2gpu_set_lattice_pot_tables(g2gpu_ctx *c, int en, const double *potcorr)
{
  if(!c)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  if(!potcorr)
    {
      c->potcorr_set = 0;
      return 0;
    }
  if(!c->cfg.periodic || c->cfg.shortrange)
    return g2_fail(G2GPU_ERR_ARG, "lattice potential tables belong to a periodic box without PM (PERIODIC && !PMGRID)");
  if(en < 1 || en > 256)
    return g2_fail(G2GPU_ERR_ARG, "EN must be in [1, 256]");
  const int D = c->D;
  const size_t n3 = (size_t) (en + 1) * (en + 1) * (en + 1);
  int nu = 0, first[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];	// identical pair tables are stored once
  for(int i = 0; i < D * D; i++)
    {
      int found = -1;
      for(int u = 0; u < nu && found < 0; u++)
	if(memcmp(potcorr + (size_t) first[u] * n3, potcorr + (size_t) i * n3, sizeof(double) * n3) == 0)
	  found = u;
      if(found < 0)
	{
	  first[nu] = i;
	  found = nu++;
	}
      c->potcorr_tabmap[i] = (unsigned char) found;
    }
  if(c->d_potcorr)
    cudaFree(c->d_potcorr);
  c->d_potcorr = nullptr;
  c->potcorr_set = 0;
  G2_CUDA(cudaMalloc((void **) &c->d_potcorr, sizeof(double) * n3 * (size_t) nu));
  for(int u = 0; u < nu; u++)
    G2_CUDA(cudaMemcpy(c->d_potcorr + (size_t) u * n3, potcorr + (size_t) first[u] * n3, sizeof(double) * n3, cudaMemcpyHostToDevice));
  c->potcorr_en = en;
  c->potcorr_set = 1;
  return 0;
}

extern "C" int g2gpu_make_ewald_pot_table(g2gpu_ctx *c, int en, double latticezero, double *out)
{
  if(!c || !out)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_make_ewald_pot_table(c, en, latticezero, out);
}

extern "C" int g2gpu_make_ewald_table(g2gpu_ctx *c, int en, double *out)
{
  if(!c || !out)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_make_ewald_table(c, en, out);
}

extern "C" int g2gpu_slice(g2gpu_ctx *c, int *lo, int *hi)
{
  if(!c)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_TRY(g2_fetch_slice(c));
  if(lo)
    *lo = c->w_lo;
  if(hi)
    *hi = c->w_hi;
  return 0;
}

extern "C" int g2gpu_gravity_tree(g2gpu_ctx *c, int npart, const float *pos, const float *mass, const int *type, const float *oldacc,
				  const int *active, const g2gpu_walk_params *wp, float *acc, float *cost, float *oldacc_out, int *perm)
{
  G2_TRY(g2gpu_upload(c, npart, pos, mass, type, oldacc, nullptr, nullptr, active));
  G2_TRY(g2gpu_domain(c));
  G2_TRY(g2gpu_treebuild(c, nullptr));
  G2_TRY(g2gpu_walk(c, wp));
  G2_TRY(g2gpu_download_acc(c, acc, cost, oldacc_out));
  if(perm)
    G2_TRY(g2gpu_get_order(c, perm));
  return 0;
}

extern "C" int g2gpu_timings(g2gpu_ctx *c, double ms[8], long long counters[8])
{
  if(!c)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  G2_CUDA(cudaStreamSynchronize(c->stream));
  if(ms)
    {
      ms[0] = c->stage >= 2 ? ev_ms(c, 0, 3) : 0;
      ms[1] = c->stage >= 3 ? ev_ms(c, 4, 5) : 0;
      ms[2] = c->stage >= 4 ? ev_ms(c, 6, 8) : 0;
      ms[3] = c->stage >= 4 ? ev_ms(c, 7, 8) : 0;
      ms[4] = c->stage >= 2 ? ev_ms(c, 1, 2) : 0;
      ms[5] = c->stage >= 1 ? ev_ms(c, 9, 10) : 0;
      ms[6] = c->stage >= 4 ? ev_ms(c, 11, 12) : 0;
      ms[7] = c->pm_done ? ev_ms(c, 13, 14) : 0;
    }
  if(counters)
    {
      counters[0] = c->launches;
      for(int i = 1; i < 8; i++)
	counters[i] = 0;
      if(c->stage >= 4)
	{
	  G2_CUDA(cudaMemcpy(c->h_counters, c->d_counters, 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
	  counters[1] = (long long) c->h_counters[0];
	  counters[2] = (long long) c->h_counters[1];
	  counters[3] = (long long) c->h_counters[2];
	  counters[4] = (long long) c->h_counters[4];
	  counters[5] = (long long) c->h_counters[5];	// targets walked again in FP64 (an FP32 decision differed from the reference's)
	  counters[6] = (long long) c->h_counters[7];	// targets flagged (equals [5] unless the list overflowed)
	}
    }
  return 0;
}

extern "C" int g2gpu_get_counts(g2gpu_ctx *c, int out[4])
{
  if(!c || !out)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  out[0] = c->npart;
  out[1] = c->stage >= 3 ? c->numnodes : 0;
  out[2] = 0;
  out[3] = c->stage;
  if(c->stage >= 4)
    {
      G2_TRY(g2_fetch_slice(c));
      out[2] = c->w_ntargets;
    }
  return 0;
}

extern "C" int g2gpu_io_bytes(g2gpu_ctx *c, long long out[2])
{
  if(!c || !out)
    return g2_fail(G2GPU_ERR_ARG, "null argument");
  out[0] = (long long) c->h2d_bytes;
  out[1] = (long long) c->d2h_bytes;
  return 0;
}

extern "C" void g2gpu_reset_counters(g2gpu_ctx *c)
{
  if(c)
    c->launches = 0;
}

// ---- stand-alone kernels for parity tests ------------------------------------------------------------------------
extern "C" int g2gpu_peano_keys(g2gpu_ctx *c, int n, const int *xyz, int bits, long long *keys)
{
  if(!c || !xyz || !keys || n < 1 || bits < 1 || bits > 21)
    return g2_fail(G2GPU_ERR_ARG, "bad argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_peano_keys_standalone(c, n, xyz, bits, keys);
}

extern "C" int g2gpu_sort_pairs(g2gpu_ctx *c, int n, unsigned long long *keys, unsigned int *vals, int begin_bit, int end_bit)
{
  if(!c || !keys || !vals || n < 1 || n > c->cfg.max_part || begin_bit < 0 || end_bit > 64 || end_bit < begin_bit)
    return g2_fail(G2GPU_ERR_ARG, "bad argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  cudaStream_t st = c->stream;
  G2_CUDA(cudaMemcpyAsync(c->skey[0], keys, sizeof(unsigned long long) * (size_t) n, cudaMemcpyHostToDevice, st));
  G2_CUDA(cudaMemcpyAsync(c->sval[0], vals, sizeof(unsigned int) * (size_t) n, cudaMemcpyHostToDevice, st));
  unsigned long long *k = c->skey[0];
  unsigned int *v = c->sval[0];
  G2_CUDA(cudaEventRecord(c->ev[13], st));
  G2_TRY(g2_radix_sort_pairs(c, n, &k, &v, c->skey[1], c->sval[1], begin_bit, end_bit));
  G2_CUDA(cudaEventRecord(c->ev[14], st));
  G2_CUDA(cudaMemcpyAsync(keys, k, sizeof(unsigned long long) * (size_t) n, cudaMemcpyDeviceToHost, st));
  G2_CUDA(cudaMemcpyAsync(vals, v, sizeof(unsigned int) * (size_t) n, cudaMemcpyDeviceToHost, st));
  G2_CUDA(cudaStreamSynchronize(st));
  c->stage = 0;			// the sort scratch held the particle order
  return 0;
}

extern "C" int g2gpu_eval_pairs(g2gpu_ctx *c, int n, int tgt, int src, const float *pm, const float *m, const float *r, const float *h,
				const int *npart_in_node, float *fac)
{
  if(!c || !pm || !m || !r || !h || !fac || n < 1 || tgt < 0 || src < 0 || tgt >= c->D || src >= c->D)
    return g2_fail(G2GPU_ERR_ARG, "bad argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_eval_pairs_standalone(c, n, tgt, src, pm, m, r, h, npart_in_node, fac);
}

extern "C" int g2gpu_eval_potentials(g2gpu_ctx *c, int n, int tgt, int src, const float *pm, const float *m, const float *r, const float *h,
				     const int *npart_in_node, float *out)
{
  if(!c || !pm || !m || !r || !h || !out || n < 1 || tgt < 0 || src < 0 || tgt >= c->D || src >= c->D)
    return g2_fail(G2GPU_ERR_ARG, "bad argument");
  G2_CUDA(cudaSetDevice(c->cfg.device));
  return g2_eval_potentials_standalone(c, n, tgt, src, pm, m, r, h, npart_in_node, out);
}
