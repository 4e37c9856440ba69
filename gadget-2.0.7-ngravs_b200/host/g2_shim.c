/* g2_shim.c — host-side drop-in for Gadget-2.0.7-ngravs: the reference's own C entry points, re-implemented on
 * top of the C ABI of include/g2gpu.h (hand-written sm_100a kernels in libg2gpu.so).
 *
 * It is compiled WITH the reference's headers (allvars.h, proto.h, ngravs.h) and replaces, symbol for symbol,
 *   gravtree.c : gravity_tree (27), set_softenings (468), grav_tree_compare_key (525)
 *   forcetree.c: force_treeallocate (3176), force_treefree (3411), force_treebuild (61),
 *                force_treeevaluate (1244; in a PERIODIC build without PMGRID together with force_treeevaluate_lattice_correction, 2077,
 *                from tables the shim tabulates like lattice_init, 3611), force_treeevaluate_shortrange (1623),
 *                force_treeevaluate_potential_shortrange (2789; PMGRID, unless -DG2_SHIM_KEEP_REFERENCE_POTENTIAL): the per-target
 *                function compute_potential() loops over (potential.c:86-97) is served from ONE device walk of all particles
 *   peano.c    : peano_hilbert_key (356), peano_hilbert_order (36), compare_key (190)
 *   pm_periodic.c (PMGRID && PERIODIC, unless -DG2_SHIM_KEEP_REFERENCE_PM): pm_init_periodic (53), pm_init_periodic_allocate (139),
 *                pm_init_periodic_free (187), pmforce_periodic (204); pmpotential_periodic (800) is not provided and ends the run
 *   gravtree_forcetest.c (-DFORCETEST, unless -DG2_SHIM_KEEP_REFERENCE_FORCETEST): gravity_forcetest (28), the direct-summation accuracy
 *                check; force_treeevaluate_direct (forcetree.c:3428) itself stays available on the host
 * so that accel.c, domain.c, init.c, run.c, ... call it unchanged.
 * All state stays in the reference's globals (P[], All, NumPart, TreeReconstructFlag, ...).  Errors of the GPU
 * library end the run through endrun() like the reference's own failures (endrun.c:24).  There is no CPU
 * fallback: without a B200 the first call to force_treeallocate() ends the run.
 */
#include <stddef.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include <unistd.h>
#include <mpi.h>

#include "allvars.h"
#include "proto.h"
#include "ngravs.h"

#include "g2gpu.h"
#include "g2_ph_table.h"

static void *g2_bound_P = NULL;	/* P[] as bound to the walk's zero-copy results (g2_bind_results) */
static int g2_bound_maxpart = 0;
static g2gpu_group *G2G = NULL;	/* all GPUs of the node (or G2GPU_NGPU of them): one context + host thread per device */
static g2gpu_ctx *G2 = NULL;	/* device 0's context: tree mirror, potential, PM, direct sums */
static int g2_ndev = 1;
static int g2_maxpart = 0, g2_maxnodes = 0;
static int *g2_perm = NULL;		/* device particle index -> index in P[] */
static int g2_perm_identity = 1;
static float *g2_acc = NULL, *g2_cost = NULL, *g2_oldacc = NULL;	/* per-target results of g2_evaluate_one (allocated on first use) */

/* ---- host loops over NumPart run on several threads (the reference's are single-threaded; at 16.8 M particles they would
 *      otherwise cost more than the device step) ---- */
struct g2_par_job { void (*fn) (long lo, long hi, void *arg); void *arg; long lo, hi; };
static void *g2_par_thread(void *v)
{
  struct g2_par_job *j = v;
  j->fn(j->lo, j->hi, j->arg);
  return NULL;
}
static void g2_parallel_for(long n, void (*fn) (long lo, long hi, void *arg), void *arg)
{
  long nt = sysconf(_SC_NPROCESSORS_ONLN), t;
  pthread_t th[16];
  int started[16];
  struct g2_par_job job[16];
  if(nt > 16)
    nt = 16;
  if(nt > n / 65536 + 1)
    nt = n / 65536 + 1;
  if(nt <= 1)
    {
      fn(0, n, arg);
      return;
    }
  for(t = 0; t < nt; t++)
    {
      job[t].fn = fn;
      job[t].arg = arg;
      job[t].lo = n * t / nt;
      job[t].hi = n * (t + 1) / nt;
      started[t] = 0;
    }
  for(t = 1; t < nt; t++)
    started[t] = pthread_create(&th[t], NULL, g2_par_thread, &job[t]) == 0;
  fn(job[0].lo, job[0].hi, arg);
  for(t = 1; t < nt; t++)
    {
      if(started[t])
	pthread_join(th[t], NULL);
      else
	fn(job[t].lo, job[t].hi, arg);	/* could not start a thread: do the slice here */
    }
}
static int g2_mirror = -1;		/* refresh the host mirror Nodes[]/Nextnode[]/Father[] after every build */
static int g2_tree_mirrored = 0;	/* Nodes[] holds the tree the device built last and P[] has not been reordered since */
static int g2_dynamic = -1;		/* follow the reference's dynamic tree updates instead of building a new tree (G2GPU_DYNAMIC_TREE) */

#ifdef PMGRID
static double g2_srtable[N_GRAVS][N_GRAVS][NTAB];
static double g2_srpot[N_GRAVS][N_GRAVS][NTAB];	/* shortrange_fourier_pot, forcetree.c:34 */
static int g2_srtable_done = 0;
#endif
static int g2_pot_valid = 0;	/* g2_pot_byP[] holds the tree potential of the particles as uploaded last */

static void g2_check(int rc, const char *what)
{
  if(rc != 0)
    {
      printf("task %d: g2gpu %s failed (%d): %s\n", ThisTask, what, rc, g2gpu_last_error());
      fflush(stdout);
      endrun(7000 - rc);
    }
}

/* ---------------------------------------------------------------- peano.c ---------------------------------- */
static const unsigned char g2_ph[G2_PH_NSTATES * 8] = G2_PH_TABLE_INIT;

/* peano.c:356.  Called by domain.c:940 for every particle and by forcetree code for top-level cells. */
peanokey peano_hilbert_key(int x, int y, int z, int bits)
{
  unsigned st = 0;
  peanokey key = 0;
  int l;
  for(l = bits - 1; l >= 0; l--)
    {
      unsigned o = (((x >> l) & 1) << 2) | (((y >> l) & 1) << 1) | ((z >> l) & 1);
      unsigned e = g2_ph[st * 8 + o];
      key = (key << 3) | (e & 7);
      st = e >> 3;
    }
  return key;
}

struct g2_keyindex { peanokey key; int index; };
int compare_key(const void *a, const void *b)	/* peano.c:190 */
{
  peanokey x = ((const struct g2_keyindex *) a)->key, y = ((const struct g2_keyindex *) b)->key;
  return (x > y) - (x < y);
}

static void g2_push_tables(void);
static void g2_fill_walk_params(g2gpu_walk_params * wp);

static void g2_upload(int npart)
{
  int off_gravpm = -1, off_vel = -1;
  g2_pot_valid = 0;
#ifdef PMGRID
  off_gravpm = (int) offsetof(struct particle_data, GravPM);
#endif
  if(g2_mirror > 0 || N_gas > 0 || (All.TreeDomainUpdateFrequency > 0.0 && TreeReconstructFlag))
    off_vel = (int) offsetof(struct particle_data, Vel);	/* Extnodes[].vs of the host mirror */
  if(off_vel >= 0 && g2_ndev > 1)
    {
      printf("g2gpu: the host tree mirror (gas particles, TreeDomainUpdateFrequency > 0 or G2GPU_HOST_MIRROR) needs node velocities, which a\n"
	     "multi-GPU group does not keep; run with G2GPU_NGPU=1.\n");
      endrun(7103);
    }
  /* every device copies its 1/N of P[] (packed by host threads into pinned 32-byte records), one NCCL all-gather replicates them */
  g2_check(g2gpu_group_upload_aos(G2G, npart, P, sizeof(struct particle_data), (int) sizeof(FLOAT),
				  (int) offsetof(struct particle_data, Pos), (int) offsetof(struct particle_data, Mass),
				  (int) offsetof(struct particle_data, Type), (int) offsetof(struct particle_data, OldAcc),
				  off_vel, off_gravpm, (int) offsetof(struct particle_data, Ti_endstep), All.Ti_Current), "upload");
}

static void g2_fetch_order(int npart)
{
  int i;
  g2_check(g2gpu_group_get_order(G2G, g2_perm), "get_order");
  g2_perm_identity = 1;
  for(i = 0; i < npart; i++)
    if(g2_perm[i] != i)
      {
	g2_perm_identity = 0;
	break;
      }
}

static void g2_copy_P(long lo, long hi, void *arg)
{
  memcpy((struct particle_data *) arg + lo, P + lo, sizeof(struct particle_data) * (size_t) (hi - lo));
}

static void g2_permute_P(long lo, long hi, void *arg)
{
  const struct particle_data *tmp = arg;
  long i;
  for(i = lo; i < hi; i++)
    P[i] = tmp[g2_perm[i]];
}

/* peano.c:36: gas block first, then species-major, Peano-Hilbert inside a block.  Keys and the sort run on the GPU;
 * the host only applies the permutation to P[] (and SphP[] for the leading gas block). */
void peano_hilbert_order(void)
{
  int i;
  struct particle_data *tmp;
  g2_tree_mirrored = 0;

  if(ThisTask == 0)
    printf("begin Peano-Hilbert order (GPU)...\n");
  if(NumPart == 0)
    return;
  g2_push_tables();		/* the sort needs TypeToGrav */
  g2_upload(NumPart);
  g2_check(g2gpu_group_domain(G2G), "domain");
  g2_fetch_order(NumPart);
  if(!g2_perm_identity)
    {
      if(!(tmp = malloc(sizeof(struct particle_data) * NumPart)))
	endrun(7101);
      g2_parallel_for(NumPart, g2_copy_P, tmp);
      g2_parallel_for(NumPart, g2_permute_P, tmp);
      free(tmp);
      if(N_gas > 0)
	{
	  struct sph_particle_data *stmp = malloc(sizeof(struct sph_particle_data) * N_gas);
	  if(!stmp)
	    endrun(7102);
	  memcpy(stmp, SphP, sizeof(struct sph_particle_data) * N_gas);
	  for(i = 0; i < N_gas; i++)
	    SphP[i] = stmp[g2_perm[i]];
	  free(stmp);
	}
    }
  if(ThisTask == 0)
    printf("Peano-Hilbert done.\n");
}

/* ---------------------------------------------------------------- forcetree.c ------------------------------- */
static int g2_law_id(gravity f)
{
  if(f == none) return G2GPU_LAW_NONE;
  if(f == newtonian) return G2GPU_LAW_NEWTONIAN;
  if(f == neg_newtonian) return G2GPU_LAW_NEG_NEWTONIAN;
  if(f == yukawa) return G2GPU_LAW_YUKAWA;
  if(f == coloyuk) return G2GPU_LAW_COLOYUK;
  if(f == bambam) return G2GPU_LAW_BAMBAM;
  if(f == sourcebambaryon) return G2GPU_LAW_SOURCEBAMBARYON;
  if(f == sourcebaryonbam) return G2GPU_LAW_SOURCEBARYONBAM;
  return -1;
}

static int g2_spline_id(gravity f)
{
  if(f == none) return G2GPU_SPLINE_NONE;
  if(f == plummer) return G2GPU_SPLINE_PLUMMER;
  if(f == neg_plummer) return G2GPU_SPLINE_NEG_PLUMMER;
  if(f == bambam_spline) return G2GPU_SPLINE_BAMBAM;
  if(f == sourcebambaryon_spline) return G2GPU_SPLINE_SOURCEBAMBARYON;
  if(f == sourcebaryonbam_spline) return G2GPU_SPLINE_SOURCEBARYONBAM;
  return -1;
}

#ifndef YUKAWA_IMASS
#define YUKAWA_IMASS 60.0	/* ngravs.c:41-43 */
#endif
#ifndef BAM_EPSILON
#define BAM_EPSILON 1.31e-6	/* ngravs.c:45-47 */
#endif

/* hands TypeToGrav, All.ForceSoftening and the D x D law tables of wire_grav_maps() (ngravs.c:64) to the device */
static void g2_push_tables(void)
{
  int accel[N_GRAVS * N_GRAVS], spline[N_GRAVS * N_GRAVS], i, j;
  double par[N_GRAVS * N_GRAVS * 4];
  for(i = 0; i < N_GRAVS; i++)
    for(j = 0; j < N_GRAVS; j++)
      {
	int a = g2_law_id(AccelFxns[i][j]), s = g2_spline_id(AccelSplines[i][j]);
	if(a < 0 || s < 0)
	  {
	    printf("ngravs/g2gpu: AccelFxns[%d][%d] or AccelSplines[%d][%d] has no device implementation.\n"
		   "Register it in include/g2gpu.h (enum g2gpu_law) and csrc/g2_laws.cuh.\n", i, j, i, j);
	    endrun(7201);
	  }
	accel[i * N_GRAVS + j] = a;
	spline[i * N_GRAVS + j] = s;
	par[4 * (i * N_GRAVS + j) + 0] = All.BoxSize > 0 ? YUKAWA_IMASS / All.BoxSize : 0.0;	/* ngravs.c:858 */
	par[4 * (i * N_GRAVS + j) + 1] = BAM_EPSILON;
	par[4 * (i * N_GRAVS + j) + 2] = par[4 * (i * N_GRAVS + j) + 3] = 0;
      }
  g2_check(g2gpu_group_set_species(G2G, TypeToGrav, All.ForceSoftening), "set_species");
  g2_check(g2gpu_group_set_laws(G2G, accel, spline, par), "set_laws");
#ifdef PMGRID
  {				/* PotentialFxns / PotentialSplines (allvars.h:147-148).  A wiring without a device implementation (the BAM potentials)
				 * is left unset: the force path does not need it, and a potential request then ends the run (G2GPU_ERR_LAW). */
    int pot[N_GRAVS * N_GRAVS], pots[N_GRAVS * N_GRAVS], ok = 1;
    for(i = 0; i < N_GRAVS; i++)
      for(j = 0; j < N_GRAVS; j++)
	{
	  gravity f = PotentialFxns[i][j], g = PotentialSplines[i][j];
	  pot[i * N_GRAVS + j] = f == none ? G2GPU_POT_NONE : f == newtonian_pot ? G2GPU_POT_NEWTONIAN : f == neg_newtonian_pot ? G2GPU_POT_NEG_NEWTONIAN :
	    f == bambam_pot ? G2GPU_POT_BAMBAM : f == sourcebaryonbam_pot ? G2GPU_POT_SOURCEBARYONBAM : f == sourcebambaryon_pot ? G2GPU_POT_SOURCEBAMBARYON : -1;
	  pots[i * N_GRAVS + j] = g == none ? G2GPU_POTSPLINE_NONE : g == plummer_pot ? G2GPU_POTSPLINE_PLUMMER : g == neg_plummer_pot ? G2GPU_POTSPLINE_NEG_PLUMMER :
	    g == bambam_pot ? G2GPU_POTSPLINE_BAMBAM : g == sourcebaryonbam_pot ? G2GPU_POTSPLINE_SOURCEBARYONBAM :
	    g == sourcebambaryon_pot ? G2GPU_POTSPLINE_SOURCEBAMBARYON : -1;
	  if(pot[i * N_GRAVS + j] < 0 || pots[i * N_GRAVS + j] < 0)
	    ok = 0;
	}
    if(ok)
      g2_check(g2gpu_set_potential_laws(G2, pot, pots), "set_potential_laws");
  }
#endif
}

#ifdef PMGRID
extern struct ngravsInterpolant *ngravsPeriodicTable;

/* forcetree.c:3274-3354: the reference's own FFT integration (ngravs_core.c:72-159) stays on the host, in double */
static void g2_build_srtable(void)
{
  int nA, nB, i;
  double temp[NTAB], tempI[NTAB], u;
  ngravsPeriodicTable = ngravsConvolutionInit(NTAB, 3, 8);
  for(nA = 0; nA < N_GRAVS; nA++)
    for(nB = 0; nB < N_GRAVS; nB++)
      {
	if(performConvolution(ngravsPeriodicTable, NormedGreensFxns[nB][nA], 0.5, temp, tempI))
	  endrun(1047);
	for(i = 0; i < NTAB; i++)
	  {
	    u = 3.0 / NTAB * (i + 0.5);
	    g2_srtable[nB][nA][i] = tempI[i] / (u * u) - temp[i] / u;
	    g2_srpot[nB][nA][i] = temp[i] / u;	/* forcetree.c:3342-3346 */
	  }
      }
  ngravsConvolutionFree(ngravsPeriodicTable);
  g2_srtable_done = 1;
}
#endif

#if defined(PERIODIC) && !defined(PMGRID)
/* lattice_init (forcetree.c:3611-3790) for the device: fcorrx/y/z[target][source] on the (NGRAVS_EN+1)^3 grid x = 0.5 (i,j,k)/NGRAVS_EN,
 * divided by BoxSize^2.  The reference's own tables are file-scope statics of forcetree.c, so they are tabulated again here: pairs wired
 * to ewald_force (the stock wiring, ngravs.c:131) on the device in FP64 (g2gpu_make_ewald_table, the same sums), any other LatticeForce
 * through the reference's function pointer on the host; every distinct function once. */
static void g2_push_lattice(void)
{
  const size_t n1 = NGRAVS_EN + 1, n3 = n1 * n1 * n1;
  double *tab = malloc(sizeof(double) * 3 * N_GRAVS * N_GRAVS * n3), *one = malloc(sizeof(double) * 3 * n3);
  int l, m, l2, m2, c, i, j, k;
  if(!tab || !one)
    endrun(7501);
  for(l = 0; l < N_GRAVS; l++)
    for(m = 0; m < N_GRAVS; m++)
      {
	int done = 0;
	for(l2 = 0; l2 <= l && !done; l2++)
	  for(m2 = 0; m2 < (l2 < l ? N_GRAVS : m) && !done; m2++)
	    if(LatticeForce[l2][m2] == LatticeForce[l][m])
	      {
		for(c = 0; c < 3; c++)
		  memcpy(tab + ((size_t) (c * N_GRAVS + l) * N_GRAVS + m) * n3, tab + ((size_t) (c * N_GRAVS + l2) * N_GRAVS + m2) * n3, sizeof(double) * n3);
		done = 1;
	      }
	if(done)
	  continue;
	if(LatticeForce[l][m] == ewald_force)
	  g2_check(g2gpu_make_ewald_table(G2, NGRAVS_EN, one), "make_ewald_table");
	else
	  for(i = 0; i <= NGRAVS_EN; i++)
	    for(j = 0; j <= NGRAVS_EN; j++)
	      for(k = 0; k <= NGRAVS_EN; k++)
		{
		  double x[3] = { 0.5 * i / NGRAVS_EN, 0.5 * j / NGRAVS_EN, 0.5 * k / NGRAVS_EN }, force[3];
		  (*LatticeForce[l][m]) (i, j, k, x, force);	/* forcetree.c:3700-3710 */
		  for(c = 0; c < 3; c++)
		    one[c * n3 + (i * n1 + j) * n1 + k] = force[c];
		}
	for(c = 0; c < 3; c++)
	  for(i = 0; i < (int) n3; i++)
	    tab[((size_t) (c * N_GRAVS + l) * N_GRAVS + m) * n3 + i] = one[c * n3 + i] / (All.BoxSize * All.BoxSize);	/* forcetree.c:3757-3761 */
      }
  g2_check(g2gpu_group_set_lattice_tables(G2G, NGRAVS_EN, tab), "set_lattice_tables");
  free(one);
  free(tab);
}
#endif

/* forcetree.c:3176.  The host arrays stay (restart.c, predict.c, ngb.c read them); the device context is created
 * once and survives force_treefree() because pm_periodic.c frees/reallocates the tree around every PM step. */
void force_treeallocate(int maxnodes, int maxpart)
{
  MaxNodes = maxnodes;
  if(!(Nodes_base = malloc((MaxNodes + 1) * sizeof(struct NODE))) || !(Extnodes_base = malloc((MaxNodes + 1) * sizeof(struct extNODE)))
     || !(Nextnode = malloc((maxpart + MAXTOPNODES) * sizeof(int))) || !(Father = malloc(maxpart * sizeof(int))))
    {
      printf("failed to allocate memory for %d tree-nodes.\n", MaxNodes);
      endrun(3);
    }
  Nodes = Nodes_base - All.MaxPart;
  Extnodes = Extnodes_base - All.MaxPart;

  if(NTask > 1)
    {				/* the shim replaces gravity_tree() without its particle export/import between MPI tasks */
      if(ThisTask == 0)
	printf("g2gpu: this drop-in serves ONE MPI task that owns all particles (it spreads the work over the GPUs of the node itself);\n"
	       "run with a single task.\n");
      endrun(7001);
    }
  if(G2G && (g2_maxpart != All.MaxPart || g2_maxnodes != maxnodes))
    {
      g2gpu_group_destroy(G2G);
      G2G = NULL;
      G2 = NULL;
      g2_bound_P = NULL;
    }
  if(!G2G)
    {
      g2gpu_config cfg;
      memset(&cfg, 0, sizeof(cfg));
      cfg.device = 0;
      cfg.n_gravs = N_GRAVS;
#ifdef PERIODIC
      cfg.periodic = 1;
#endif
#ifdef PMGRID
      cfg.shortrange = 1;
      cfg.ntab = NTAB;
#endif
#ifdef UNEQUALSOFTENINGS
      cfg.unequal_softenings = 1;
#endif
      cfg.max_part = All.MaxPart;
      cfg.max_nodes = maxnodes;
      /* all visible GPUs share the work (ndev = 0); G2GPU_NGPU limits them */
      g2_check(g2gpu_group_create(&G2G, &cfg, getenv("G2GPU_NGPU") ? atoi(getenv("G2GPU_NGPU")) : 0, NULL), "group_create");
      g2_ndev = g2gpu_group_size(G2G);
      G2 = g2gpu_group_ctx(G2G, 0);
#ifdef NGRAVS_ACCUMULATOR
      g2_check(g2gpu_group_set_option(G2G, "accumulator", 1), "set_option");	/* laws receive Nodes[].u.d.Nparticles as N */
#endif
      g2_maxpart = All.MaxPart;
      g2_maxnodes = maxnodes;
      g2_perm = realloc(g2_perm, sizeof(int) * All.MaxPart);
      if(ThisTask == 0)
	printf("\ng2gpu: %d device context%s for %d particles / %d tree nodes created.\n\n", g2_ndev, g2_ndev > 1 ? "s" : "", All.MaxPart, maxnodes);
#ifdef PMGRID
      if(!g2_srtable_done)
	g2_build_srtable();
      g2_check(g2gpu_group_set_srtable(G2G, &g2_srtable[0][0][0]), "set_srtable");
      g2_check(g2gpu_set_srpot_table(G2, &g2_srpot[0][0][0]), "set_srpot_table");
#endif
#if defined(PERIODIC) && !defined(PMGRID)
      g2_push_lattice();	/* gravity_tree() then includes the lattice-sum correction walk (forcetree.c:1606-1608) */
#endif
    }
  if(g2_mirror < 0)
    g2_mirror = getenv("G2GPU_HOST_MIRROR") ? atoi(getenv("G2GPU_HOST_MIRROR")) : 0;
}

void force_treefree(void)	/* forcetree.c:3411 */
{
  free(Father);
  free(Nextnode);
  free(Extnodes_base);
  free(Nodes_base);
}

/* ---------------------------------------------------------------- pm_periodic.c ----------------------------- */
#if defined(PMGRID) && defined(PERIODIC) && !defined(G2_SHIM_KEEP_REFERENCE_PM)
static int g2_greens_id(gravity f)
{
  if(f == none) return G2GPU_GREENS_NONE;
  if(f == pgdelta) return G2GPU_GREENS_NEWTON;
  if(f == neg_pgdelta) return G2GPU_GREENS_NEG_NEWTON;
  if(f == pgyukawa) return G2GPU_GREENS_YUKAWA;
  if(f == pgcoloyuk) return G2GPU_GREENS_COLOYUK;
  return -1;
}

void pm_init_periodic(void)	/* pm_periodic.c:53: no FFTW plans or slab tables to set up, cuFFT plans are made on first use */
{
  All.Asmth[0] = ASMTH * All.BoxSize / PMGRID;
  All.Rcut[0] = RCUT * All.Asmth[0];
}

void pm_init_periodic_allocate(int dimprod) { (void) dimprod; }	/* the meshes live on the device */
void pm_init_periodic_free(void) { }

static void g2_fill_pm_params(g2gpu_pm_params * pp)
{
  int nA, nB;
  memset(pp, 0, sizeof(*pp));
  pp->pmgrid = PMGRID;
  pp->boxsize = All.BoxSize;
  pp->asmth = All.Asmth[0];
  pp->G = All.G;
  for(nA = 0; nA < N_GRAVS; nA++)
    for(nB = 0; nB < N_GRAVS; nB++)
      {
	int id = g2_greens_id(GreensFxns[nA][nB]);
	if(id < 0)
	  {
	    printf("ngravs/g2gpu: GreensFxns[%d][%d] has no device implementation.\n"
		   "Register it in include/g2gpu.h (enum g2gpu_greens) and csrc/g2_pm.cu.\n", nA, nB);
	    endrun(7312);
	  }
	pp->greens_id[nA * N_GRAVS + nB] = id;
	pp->greens_par[nA * N_GRAVS + nB] = YUKAWA_IMASS / (2 * M_PI);	/* ngravs.c:871 */
      }
}

/* pm_periodic.c:204.  long_range_force() (longrange.c:56) calls it before gravity_tree() (accel.c:36-46) with P[] in whatever
 * order the last domain decomposition left; the device version does not need species blocks, so P[] is uploaded as it is and
 * P[].GravPM is read back in the same order. */
void pmforce_periodic(void)
{
  g2gpu_pm_params pp;
  float *gpm;
  int i;

  if(ThisTask == 0)
    {
      printf("Starting periodic PM calculation.\n");
      fflush(stdout);
    }
  if(!G2)
    {
      printf("g2gpu: pmforce_periodic() before force_treeallocate()\n");
      endrun(7311);
    }
  g2_fill_pm_params(&pp);
  g2_push_tables();
  g2_upload(NumPart);
  g2_check(g2gpu_pm_periodic(G2, &pp), "pm_periodic");
  gpm = malloc(sizeof(float) * 3 * (size_t) NumPart);
  g2_check(g2gpu_download_gravpm(G2, gpm), "download_gravpm");
  for(i = 0; i < NumPart; i++)
    {
      P[i].GravPM[0] = gpm[3 * i + 0];
      P[i].GravPM[1] = gpm[3 * i + 1];
      P[i].GravPM[2] = gpm[3 * i + 2];
    }
  free(gpm);
  /* the reference frees and re-allocates the tree storage around the PM step and therefore asks for a new domain decomposition
   * (pm_periodic.c:232, 781-783); the request is kept so that the sequence of decompositions and tree builds of a run is unchanged */
  All.NumForcesSinceLastDomainDecomp = 1 + All.TotNumPart * All.TreeDomainUpdateFrequency;
  TreeReconstructFlag = 1;
  if(ThisTask == 0)
    {
      printf("done PM.\n");
      fflush(stdout);
    }
}

/* pm_periodic.c:798: the long-range potential, added to P[].Potential (compute_potential, potential.c:271).  One device pass
 * (g2gpu_pm_potential_periodic); see include/g2gpu.h for the k = 0 mode. */
void pmpotential_periodic(void)
{
  g2gpu_pm_params pp;
  float *pot;
  int i;

  if(ThisTask == 0)
    {
      printf("Starting periodic PM calculation.\n");
      fflush(stdout);
    }
  if(!G2)
    {
      printf("g2gpu: pmpotential_periodic() before force_treeallocate()\n");
      endrun(7313);
    }
  g2_fill_pm_params(&pp);
  g2_push_tables();
  g2_upload(NumPart);
  pot = malloc(sizeof(float) * (size_t) NumPart);
  g2_check(g2gpu_pm_potential_periodic(G2, &pp, pot), "pm_potential_periodic");
  for(i = 0; i < NumPart; i++)
    P[i].Potential += pot[i];
  free(pot);
  All.NumForcesSinceLastDomainDecomp = 1 + All.TotNumPart * All.TreeDomainUpdateFrequency;	/* pm_periodic.c:1283 */
  TreeReconstructFlag = 1;
  if(ThisTask == 0)
    {
      printf("done PM-Potential.\n");
      fflush(stdout);
    }
}
#endif

/* ---------------------------------------------------------------- gravtree_forcetest.c ---------------------- */
#if defined(FORCETEST) && !defined(G2_SHIM_KEEP_REFERENCE_FORCETEST)
/* gravtree_forcetest.c:28.  accel.c:52-54 calls it right after gravity_tree(), so the device still holds the particles of this force
 * computation (current order, g2_perm).  The random fraction FORCETEST of the active particles is selected exactly like the reference does
 * (get_random_number(P[i].ID)); their direct sums run as ONE device call (one CTA per target, FP64).  In a periodic box the lattice
 * correction of all images is evaluated exactly (Ewald sums) where the reference interpolates its 65^3 table (lattice_corr, forcetree.c:3803):
 * Newtonian pairs only.  forcetest.txt lines and the DIRECT line of timings.txt keep the reference's formats. */
void gravity_forcetest(void)
{
  int i, j, k, nt = 0, *targets, *pidx, *inv;
  double *acc, tstart, tend, timetree, fac1;
  g2gpu_walk_params wp;
  char buf[200];

#ifdef PMGRID
  if(All.PM_Ti_endstep != All.Ti_Current)
    return;
#endif
  if(All.ComovingIntegrationOn)
    set_softenings();
  for(i = 0, NumForceUpdate = 0; i < NumPart; i++)
    if(P[i].Ti_endstep == All.Ti_Current && get_random_number(P[i].ID) < FORCETEST)
      {
	P[i].Ti_endstep = -P[i].Ti_endstep - 1;
	NumForceUpdate++;
      }
  nt = NumForceUpdate;
  targets = malloc(sizeof(int) * (nt + 1));
  pidx = malloc(sizeof(int) * (nt + 1));
  acc = malloc(sizeof(double) * 3 * (nt + 1));
  inv = malloc(sizeof(int) * (NumPart + 1));
  for(i = 0; i < NumPart; i++)
    inv[g2_perm[i]] = i;	/* index in P[] -> index on the device */
  for(i = 0, nt = 0; i < NumPart; i++)
    if(P[i].Ti_endstep < 0)
      {
	targets[nt] = inv[i];
	pidx[nt] = i;
	nt++;
      }
  tstart = second();
  g2_fill_walk_params(&wp);
  wp.asmth = wp.rcut = 0;	/* the complete force, not its short-range part */
#ifdef PERIODIC
  for(i = 0; i < N_GRAVS; i++)
    for(j = 0; j < N_GRAVS; j++)
      if(AccelFxns[i][j] != newtonian)
	{
	  printf("g2gpu: gravity_forcetest() in a periodic box knows the lattice sum of Newtonian pairs only (pair %d,%d)\n", i, j);
	  endrun(7321);
	}
  g2_check(g2gpu_set_option(G2, "direct_ewald", 1), "set_option");
#endif
  if(nt > 0)
    g2_check(g2gpu_direct(G2, &wp, nt, targets, acc), "direct");
  tend = second();
  timetree = timediff(tstart, tend);
  for(i = 0; i < nt; i++)
    for(k = 0; k < 3; k++)
      P[pidx[i]].GravAccelDirect[k] = acc[3 * i + k];
  /* gravtree_forcetest.c:241-268 */
  if(All.ComovingIntegrationOn)
    {
#ifndef PERIODIC
      fac1 = 0.5 * All.Hubble * All.Hubble * All.Omega0 / All.G;
      for(i = 0; i < NumPart; i++)
	if(P[i].Ti_endstep < 0)
	  for(j = 0; j < 3; j++)
	    P[i].GravAccelDirect[j] += fac1 * P[i].Pos[j];
#endif
    }
  for(i = 0; i < NumPart; i++)
    if(P[i].Ti_endstep < 0)
      for(j = 0; j < 3; j++)
	P[i].GravAccelDirect[j] *= All.G;
  if(All.ComovingIntegrationOn == 0)
    {
      fac1 = All.OmegaLambda * All.Hubble * All.Hubble;
      for(i = 0; i < NumPart; i++)
	if(P[i].Ti_endstep < 0)
	  for(j = 0; j < 3; j++)
	    P[i].GravAccelDirect[j] += fac1 * P[i].Pos[j];
    }
  sprintf(buf, "%s%s", All.OutputDir, "forcetest.txt");
  if(!(FdForceTest = fopen(buf, "a")))
    {
      printf("error in opening file '%s'\n", buf);
      endrun(17);
    }
  for(i = 0; i < NumPart; i++)
    if(P[i].Ti_endstep < 0)
      {
#ifndef PMGRID
	fprintf(FdForceTest, "%d %g %g %g %g %g %g %g %g %g %g %g %d\n", P[i].Type, All.Time, All.Time - TimeOfLastTreeConstruction,
		P[i].Pos[0], P[i].Pos[1], P[i].Pos[2], P[i].GravAccelDirect[0], P[i].GravAccelDirect[1], P[i].GravAccelDirect[2],
		P[i].GravAccel[0], P[i].GravAccel[1], P[i].GravAccel[2], P[i].ID);
#else
	fprintf(FdForceTest, "%d %f %f %f %f %f %.15e %.15e %.15e %.15e %.15e %.15e %.15e %.15e %.15e %d\n", P[i].Type, All.Time,
		All.Time - TimeOfLastTreeConstruction, P[i].Pos[0], P[i].Pos[1], P[i].Pos[2], P[i].GravAccelDirect[0], P[i].GravAccelDirect[1],
		P[i].GravAccelDirect[2], P[i].GravAccel[0], P[i].GravAccel[1], P[i].GravAccel[2], P[i].GravPM[0] + P[i].GravAccel[0],
		P[i].GravPM[1] + P[i].GravAccel[1], P[i].GravPM[2] + P[i].GravAccel[2], P[i].ID);
#endif
      }
  fclose(FdForceTest);
  for(i = 0; i < NumPart; i++)
    if(P[i].Ti_endstep < 0)
      P[i].Ti_endstep = -P[i].Ti_endstep - 1;
  if(ThisTask == 0)
    {
      fprintf(FdTimings, "DIRECT Nf= %d    part/sec=%g | %g  ia/part=%g \n", nt, nt / (timetree + 1.0e-20), nt / (timetree + 1.0e-20),
	      (double) NumPart);
      fprintf(FdTimings, "\n");
      fflush(FdTimings);
    }
  free(inv);
  free(acc);
  free(pidx);
  free(targets);
}
#endif

/* copies the device tree into the reference's host layout (struct NODE after the moment pass, allvars.h:618-660) */
static void g2_refresh_mirror(int npart)
{
  int nn = Numnodestree, k, j, g;
  float *len = malloc(sizeof(float) * nn), *center = malloc(sizeof(float) * 3 * nn), *s = malloc(sizeof(float) * 3 * nn * N_GRAVS),
    *mass = malloc(sizeof(float) * nn * N_GRAVS), *vs = malloc(sizeof(float) * 3 * nn * N_GRAVS);
  int *bf = malloc(sizeof(int) * nn * 4), *sib = bf + nn, *nxt = sib + nn, *fat = nxt + nn;
  if(!len || !center || !s || !mass || !bf || !vs)
    endrun(7301);
  g2_check(g2gpu_download_tree(G2, len, center, s, mass, bf, sib, nxt, fat, Nextnode, Father), "download_tree");
  g2_check(g2gpu_download_extnodes(G2, vs), "download_extnodes");
  for(k = 0; k < nn; k++)
    {
      struct NODE *nop = &Nodes[All.MaxPart + k];
      nop->len = len[k];
      for(j = 0; j < 3; j++)
	{
	  nop->center[j] = center[3 * k + j];
	  for(g = 0; g < N_GRAVS; g++)
	    nop->u.d.s[j][g] = s[(3 * k + j) * N_GRAVS + g];
	}
      for(g = 0; g < N_GRAVS; g++)
	nop->u.d.mass[g] = mass[k * N_GRAVS + g];
      for(j = 0; j < 3; j++)
	for(g = 0; g < N_GRAVS; g++)
	  Extnodes[All.MaxPart + k].vs[j][g] = vs[(3 * k + j) * N_GRAVS + g];
      Extnodes[All.MaxPart + k].hmax = 0;	/* SPH smoothing lengths are outside the path */
      nop->u.d.bitflags = bf[k];
      nop->u.d.sibling = sib[k];
      nop->u.d.nextnode = nxt[k];
      nop->u.d.father = fat[k];
    }
#ifdef NGRAVS_ACCUMULATOR
  {
    long long *cnt = malloc(sizeof(long long) * nn * N_GRAVS);
    if(!cnt)
      endrun(7301);
    g2_check(g2gpu_download_nparticles(G2, cnt), "download_nparticles");
    for(k = 0; k < nn; k++)
      for(g = 0; g < N_GRAVS; g++)
	Nodes[All.MaxPart + k].u.d.Nparticles[g] = (long) cnt[k * N_GRAVS + g];
    free(cnt);
  }
#endif
  (void) npart;
  free(vs);
  free(bf);
  free(mass);
  free(s);
  free(center);
  free(len);
}

/* P[] is not in the device's order (a construction was requested without a preceding peano_hilbert_order): the host-side
 * dynamic updates get a one-node tree to act on, so that every Father[] chain is valid memory. */
static void g2_neutral_mirror(int npart)
{
  struct NODE *root = &Nodes[All.MaxPart];
  int i, j;
  memset(root, 0, sizeof(*root));
  memset(&Extnodes[All.MaxPart], 0, sizeof(struct extNODE));
  root->len = DomainLen;
  for(j = 0; j < 3; j++)
    root->center[j] = DomainCenter[j];
  root->u.d.sibling = root->u.d.nextnode = root->u.d.father = -1;
  for(i = 0; i < npart; i++)
    {
      Father[i] = All.MaxPart;
      Nextnode[i] = -1;
    }
}

/* forcetree.c:61.  Uploads the first npart particles, runs the device domain stage (extent, keys, top-level tree)
 * and the parallel build; the tree always reflects the current P[]. */
int force_treebuild(int npart)
{
  int numnodes = 0;
  if(npart <= 0)
    {				/* ngb_treebuild() of a run without gas (init.c:151, ngb.c:408): nothing to build */
      g2_neutral_mirror(0);
      Numnodestree = 1;
      return Numnodestree;
    }
  g2_tree_mirrored = 0;
  g2_push_tables();
  g2_upload(npart);
  g2_check(g2gpu_group_domain(G2G), "domain");
  g2_fetch_order(npart);
  g2_check(g2gpu_group_treebuild(G2G, &numnodes), "treebuild");
  Numnodestree = numnodes;
  {
    /* Who reads the host arrays: ngb.c / density.c (gas), and -- when the run does not rebuild at every step
     * (TreeDomainUpdateFrequency > 0) -- predict.c:79-91, timestep.c:329-344 and force_update_len(), which drift / kick /
     * enlarge Nodes[], Extnodes[] through Father[] between two constructions.  They get the tree of the construction the
     * reference itself asked for (TreeReconstructFlag); the device tree is rebuilt from the current P[] regardless. */
    const int forced = g2_mirror > 0 || N_gas > 0;
    const int dynamic = All.TreeDomainUpdateFrequency > 0.0 && TreeReconstructFlag;
    if(forced || dynamic)
      {
	if(g2_perm_identity)
	  {
	    g2_refresh_mirror(npart);
	    g2_tree_mirrored = npart;
	  }
	else if(forced)
	  {
	    printf("g2gpu: host tree mirror requested but P[] is not in Peano-Hilbert order\n");
	    endrun(7302);
	  }
	else
	  g2_neutral_mirror(npart);
      }
  }
  TimeOfLastTreeConstruction = All.Time;
  return Numnodestree;
}

static void g2_fill_walk_params(g2gpu_walk_params * wp)
{
  memset(wp, 0, sizeof(*wp));
  wp->theta = All.ErrTolTheta;
  wp->errtol_force_acc = All.ErrTolForceAcc;
  wp->boxsize = All.BoxSize;
  wp->G = All.G;
#ifdef PMGRID
  wp->asmth = All.Asmth[0];
  wp->rcut = All.Rcut[0];
  wp->use_gravpm = 1;
#endif
#if !defined(PERIODIC) && !defined(PMGRID)
  if(All.ComovingIntegrationOn)
    wp->pos_fac_pre_g = 0.5 * All.Hubble * All.Hubble * All.Omega0 / All.G;	/* gravtree.c:304-316 */
  else
    wp->pos_fac_post_g = All.OmegaLambda * All.Hubble * All.Hubble;	/* gravtree.c:346-358 */
#endif
}

/* forcetree.c:1244 / 1623 for ONE target: kept for callers outside gravity_tree().  Runs a device walk with only
 * `target` active; returns its interaction count. */
static int g2_evaluate_one(int target, int mode)
{
  g2gpu_walk_params wp;
  int i, save, n = NumPart, cost;
  int *ti;
  if(mode != 0)
    {
      printf("g2gpu: force_treeevaluate(mode=1) (imported particles) does not exist without domain decomposition\n");
      endrun(7401);
    }
  ti = malloc(sizeof(int) * n);
  for(i = 0; i < n; i++)
    {
      ti[i] = P[i].Ti_endstep;
      P[i].Ti_endstep = All.Ti_Current + 1;
    }
  P[target].Ti_endstep = All.Ti_Current;
  save = TreeReconstructFlag;
  force_treebuild(NumPart);
  TreeReconstructFlag = save;
  g2_fill_walk_params(&wp);
  wp.G = 1.0;			/* the per-target functions return the pre-G acceleration (forcetree.c:1592) */
  wp.pos_fac_pre_g = wp.pos_fac_post_g = 0;
  if(!g2_acc)
    {
      g2_acc = malloc(sizeof(float) * 3 * (size_t) All.MaxPart);
      g2_cost = malloc(sizeof(float) * (size_t) All.MaxPart);
      g2_oldacc = malloc(sizeof(float) * (size_t) All.MaxPart);
      if(!g2_acc || !g2_cost || !g2_oldacc)
	endrun(7404);
    }
  g2_check(g2gpu_group_walk(G2G, &wp), "walk");
  g2_check(g2gpu_group_download_acc(G2G, g2_acc, g2_cost, g2_oldacc), "download_acc");
  for(i = 0; i < n; i++)
    if(g2_perm[i] == target)
      {
	P[target].GravAccel[0] = g2_acc[3 * i];
	P[target].GravAccel[1] = g2_acc[3 * i + 1];
	P[target].GravAccel[2] = g2_acc[3 * i + 2];
	P[target].GravCost = g2_cost[i];
	break;
      }
  cost = (int) P[target].GravCost;
  for(i = 0; i < n; i++)
    P[i].Ti_endstep = ti[i];
  free(ti);
  return cost;
}

int force_treeevaluate(int target, int mode, double *ewaldcountsum)
{
  (void) ewaldcountsum;
  return g2_evaluate_one(target, mode);
}

#ifdef PMGRID
int force_treeevaluate_shortrange(int target, int mode)
{
  return g2_evaluate_one(target, mode);
}
#endif

#if defined(PMGRID) && !defined(G2_SHIM_KEEP_REFERENCE_POTENTIAL)
/* forcetree.c:2789.  compute_potential() (potential.c:86-97) calls this once per particle.  The first call after the particles or
 * the opening parameters changed builds a tree from the current P[] and walks it for ALL particles in one device call
 * (g2gpu_potential); every call then stores its particle's value in P[target].Potential exactly as the reference's function does
 * (pre-G, self term included: potential.c:250-270 goes on from there unchanged). */
static float *g2_pot_byP = NULL;
static int g2_pot_n = 0;
static double g2_pot_theta = -1, g2_pot_errtol = -1;

void force_treeevaluate_potential_shortrange(int target, int mode)
{
  if(mode != 0)
    {
      printf("g2gpu: force_treeevaluate_potential_shortrange(mode=1) (imported particles) does not exist without domain decomposition\n");
      endrun(7402);
    }
  if(!g2_pot_valid || g2_pot_n != NumPart || g2_pot_theta != All.ErrTolTheta || g2_pot_errtol != All.ErrTolForceAcc)
    {
      g2gpu_walk_params wp;
      float *pot = malloc(sizeof(float) * (size_t) NumPart);
      int i, save = TreeReconstructFlag;
      if(!pot || !(g2_pot_byP = realloc(g2_pot_byP, sizeof(float) * (size_t) All.MaxPart)))
	endrun(7403);
      force_treebuild(NumPart);
      TreeReconstructFlag = save;
      g2_fill_walk_params(&wp);
      g2_check(g2gpu_set_option(G2, "nranks", 1), "set_option");	/* device 0 walks all particles (compute_potential is not on the benchmarked path) */
      g2_check(g2gpu_potential(G2, &wp), "potential");
      g2_check(g2gpu_download_potential(G2, pot, NULL), "download_potential");
      for(i = 0; i < NumPart; i++)
	g2_pot_byP[g2_perm[i]] = pot[i];
      free(pot);
      g2_pot_n = NumPart;
      g2_pot_theta = All.ErrTolTheta;
      g2_pot_errtol = All.ErrTolForceAcc;
      g2_pot_valid = 1;
    }
  P[target].Potential = g2_pot_byP[target];
}
#endif

/* ---------------------------------------------------------------- gravtree.c -------------------------------- */
void set_softenings(void)	/* gravtree.c:468 */
{
  const double soft[6] = { All.SofteningGas, All.SofteningHalo, All.SofteningDisk, All.SofteningBulge, All.SofteningStars, All.SofteningBndry };
  const double maxphys[6] = { All.SofteningGasMaxPhys, All.SofteningHaloMaxPhys, All.SofteningDiskMaxPhys, All.SofteningBulgeMaxPhys,
    All.SofteningStarsMaxPhys, All.SofteningBndryMaxPhys };
  int i;
  for(i = 0; i < 6; i++)
    {
      if(All.ComovingIntegrationOn && soft[i] * All.Time > maxphys[i])
	All.SofteningTable[i] = maxphys[i] / All.Time;
      else
	All.SofteningTable[i] = soft[i];
      All.ForceSoftening[i] = 2.8 * All.SofteningTable[i];
    }
  All.MinGasHsml = All.MinGasHsmlFractional * All.ForceSoftening[0];
}

int grav_tree_compare_key(const void *a, const void *b)	/* gravtree.c:525 */
{
  int x = ((const struct gravdata_index *) a)->Task, y = ((const struct gravdata_index *) b)->Task;
  return (x > y) - (x < y);
}

/* The reference does not build a new tree for every force computation: while NumForcesSinceLastDomainDecomp stays below
 * TotNumPart * TreeDomainUpdateFrequency it drifts the node centres of mass (predict.c:79-91), kicks the node velocities
 * (timestep.c:329-344) and enlarges nodes (force_update_len, forcetree.c:1005) -- all on the host arrays, which the shim filled from the
 * device tree at the last construction.  Here the device tree of that construction takes the current P[] and the host's len / s, so the
 * walk sees the tree the reference's walk would see. */
static void g2_dynamic_update(int npart)
{
  const int nn = Numnodestree;
  float *len = malloc(sizeof(float) * (size_t) nn), *s = malloc(sizeof(float) * 3 * (size_t) nn * N_GRAVS);
  int k, j, g;
  for(k = 0; k < nn; k++)
    {
      const struct NODE *nop = &Nodes[All.MaxPart + k];
      len[k] = nop->len;
      for(j = 0; j < 3; j++)
	for(g = 0; g < N_GRAVS; g++)
	  s[(3 * (size_t) k + j) * N_GRAVS + g] = nop->u.d.s[j][g];
    }
  g2_push_tables();
  g2_upload(npart);
  g2_check(g2gpu_group_update_tree(G2G, len, s), "update_tree");
  free(s);
  free(len);
}

/* The results of the walk go straight into P[] (zero-copy stores of the walk kernel into page-locked host memory) whenever P[] can be
 * page-locked; bound again when the reference has moved or resized P[]. */
static void g2_bind_results(void)
{
  if(g2_bound_P == (void *) P && g2_bound_maxpart == All.MaxPart)
    return;
  g2_check(g2gpu_group_bind_results_aos(G2G, P, (size_t) All.MaxPart, sizeof(struct particle_data), (int) sizeof(FLOAT),
					(int) offsetof(struct particle_data, GravAccel), (int) offsetof(struct particle_data, GravCost),
					(int) offsetof(struct particle_data, OldAcc)), "bind_results_aos");
  g2_bound_P = (void *) P;
  g2_bound_maxpart = All.MaxPart;
}

/* gravtree.c:27.  Build if flagged (always from the current P[]), walk every particle with
 * Ti_endstep == All.Ti_Current on the GPU, epilogue (OldAcc, G, cosmological terms) fused into the walk kernel,
 * results written back into P[], the reference's counters and timings.txt lines kept. */
void gravity_tree(void)
{
  double tstart, tend, timetree, costtotal = 0;
  long long ntot = NumForceUpdate;
  g2gpu_walk_params wp;
  int i, j;
  (void) i;
  (void) j;	/* used by the NOGRAVITY / SELECTIVE_NO_GRAVITY branches only */

  if(All.ComovingIntegrationOn)
    set_softenings();

  tstart = second();
  if(g2_dynamic < 0)
    g2_dynamic = getenv("G2GPU_DYNAMIC_TREE") ? atoi(getenv("G2GPU_DYNAMIC_TREE")) : 1;
  if(ThisTask == 0 && TreeReconstructFlag)
    printf("Tree construction.\n");
  if(!TreeReconstructFlag && g2_dynamic && g2_tree_mirrored == NumPart && All.TreeDomainUpdateFrequency > 0.0)
    g2_dynamic_update(NumPart);	/* the reference keeps its (drifted) tree: so does the device */
  else
    force_treebuild(NumPart);	/* a new device tree from the current P[] */
  TreeReconstructFlag = 0;
  tend = second();
  All.CPU_TreeConstruction += timediff(tstart, tend);

#ifndef NOGRAVITY
  if(ThisTask == 0)
    printf("Begin tree force.\n");
#ifdef SELECTIVE_NO_GRAVITY
  for(i = 0; i < NumPart; i++)
    if(((1 << P[i].Type) & (SELECTIVE_NO_GRAVITY)))
      P[i].Ti_endstep = -P[i].Ti_endstep - 1;
  force_treebuild(NumPart);	/* re-upload the changed active flags */
#endif
  tstart = second();
  g2_fill_walk_params(&wp);
  g2_bind_results();		/* P[] page-locked once: the walk kernel writes GravAccel / GravCost / OldAcc straight into it */
  g2_check(g2gpu_group_walk(G2G, &wp), "walk");
  /* every device returns its slice of the active targets; host threads write GravAccel, GravCost and OldAcc straight into P[] */
  g2_check(g2gpu_group_download_aos(G2G, P, sizeof(struct particle_data), (int) sizeof(FLOAT), (int) offsetof(struct particle_data, GravAccel),
				    (int) offsetof(struct particle_data, GravCost), (int) offsetof(struct particle_data, OldAcc),
				    g2_perm_identity ? NULL : g2_perm, &costtotal), "download_aos");
  tend = second();
  timetree = timediff(tstart, tend);

  if(All.TypeOfOpeningCriterion == 1)
    All.ErrTolTheta = 0;	/* gravtree.c:334-335 */
#ifdef SELECTIVE_NO_GRAVITY
  for(i = 0; i < NumPart; i++)
    if(P[i].Ti_endstep < 0)
      P[i].Ti_endstep = -P[i].Ti_endstep - 1;
#endif
  if(ThisTask == 0)
    printf("tree is done.\n");
#else
  timetree = 0;
  for(i = 0; i < NumPart; i++)
    if(P[i].Ti_endstep == All.Ti_Current)
      for(j = 0; j < 3; j++)
	P[i].GravAccel[j] = 0;
#endif

  if(ThisTask == 0)
    {				/* timings.txt, gravtree.c:404-452 */
      All.TotNumOfForces += ntot;
      fprintf(FdTimings, "Step= %d  t= %g  dt= %g \n", All.NumCurrentTiStep, All.Time, All.TimeStep);
      fprintf(FdTimings, "Nf= %d%09d  total-Nf= %d%09d  ex-frac= %g  iter= %d\n", (int) (ntot / 1000000000), (int) (ntot % 1000000000),
	      (int) (All.TotNumOfForces / 1000000000), (int) (All.TotNumOfForces % 1000000000), 0.0, 1);
      fprintf(FdTimings, "work-load balance: %g  max=%g avg=%g PE0=%g\n", 1.0, timetree, timetree, timetree);
      fprintf(FdTimings, "particle-load balance: %g\n", 1.0);
      fprintf(FdTimings, "max. nodes: %d, filled: %g\n", Numnodestree, Numnodestree / (All.TreeAllocFactor * All.MaxPart));
      fprintf(FdTimings, "part/sec=%g | %g  ia/part=%g (%g)\n", ntot / (timetree + 1.0e-20), ntot / (timetree + 1.0e-20),
	      ntot > 0 ? costtotal / ntot : 0.0, 0.0);
      fprintf(FdTimings, "\n");
      fflush(FdTimings);
      All.CPU_TreeWalk += timetree;
    }
}
