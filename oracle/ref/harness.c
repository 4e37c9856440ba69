/* oracle/ref/harness.c — TEST INFRASTRUCTURE ONLY (never linked into the product).
 *
 * Drives the UNMODIFIED reference sources (compiled where they lie under /root/reference by
 * oracle/ref/Makefile) as a single-rank CPU oracle.  It fills the reference's own globals
 * (`All`, `P[]`, ...) the way begrun()/init() would (begrun.c:28-140, init.c:20-160), calls the
 * reference's own entry points
 *     domain_Decomposition()  (domain.c:62)   keys + top tree + species-major PH order
 *     gravity_tree()          (gravtree.c:27) tree build + walk + G scaling
 * and copies results out of the reference's globals.  FLOAT quantities are returned as double
 * (exact for float).  All indices are the reference's own (particles [0,MaxPart), nodes
 * [MaxPart, MaxPart+MaxNodes)).
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <pthread.h>
#include <sys/time.h>
#include <unistd.h>

#include "allvars.h"
#include "proto.h"
#include "ngravs.h"

/* accessor defined in ref_forcetree_unit.c (needs file-scope statics of forcetree.c) */
int g2ref_srtable_copy(double *out);
int g2ref_srpot_copy(double *out);

static double wall(void)
{
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return tv.tv_sec + 1e-6 * tv.tv_usec;
}

static int g2ref_ready = 0;
static double t_domain, t_gravity, t_build, t_walk;

/* compile-time configuration of this oracle variant */
void g2ref_config(int *out)
{
  out[0] = N_GRAVS;
  out[1] = (int) sizeof(FLOAT);
#ifdef PERIODIC
  out[2] = 1;
#else
  out[2] = 0;
#endif
#ifdef PMGRID
  out[3] = PMGRID;
  out[4] = NTAB;
#else
  out[3] = 0;
  out[4] = 0;
#endif
#ifdef UNEQUALSOFTENINGS
  out[5] = 1;
#else
  out[5] = 0;
#endif
#ifdef FORCETEST
  out[6] = 1;
#else
  out[6] = 0;
#endif
  out[7] = (int) sizeof(struct particle_data);
  out[8] = (int) sizeof(struct NODE);
  out[9] = (int) sizeof(struct extNODE);
  out[10] = BITS_PER_DIMENSION;
}

/* par[0]=MaxPart par[1]=BoxSize par[2]=G par[3]=ErrTolTheta par[4]=ErrTolForceAcc
 * par[5]=TypeOfOpeningCriterion par[6..11]=Softening (Plummer-equivalent, per type)
 * par[12..17]=Gravity<Type> (species of each type)  par[18]=TreeAllocFactor par[19]=BufferSize(MB) */
int g2ref_setup(const double *par)
{
  int i;

  if(g2ref_ready)
    return -1;			/* globals of the reference are one-shot */

  ThisTask = 0;
  NTask = 1;
  PTask = 0;

  memset(&All, 0, sizeof(All));
  All.MaxPart = (int) par[0];
  All.MaxPartSph = (int) par[0];	/* reorder_gas() moves SphP[] when type-0 particles exist */
  All.BoxSize = par[1];
  All.G = par[2];
  All.ErrTolTheta = par[3];
  All.ErrTolForceAcc = par[4];
  All.TypeOfOpeningCriterion = (int) par[5];
  All.SofteningGas = par[6];
  All.SofteningHalo = par[7];
  All.SofteningDisk = par[8];
  All.SofteningBulge = par[9];
  All.SofteningStars = par[10];
  All.SofteningBndry = par[11];
  All.GravityGas = (int) par[12];
  All.GravityHalo = (int) par[13];
  All.GravityDisk = (int) par[14];
  All.GravityBulge = (int) par[15];
  All.GravityStars = (int) par[16];
  All.GravityBndry = (int) par[17];
  All.TreeAllocFactor = par[18];
  All.BufferSize = (int) par[19];
  All.PartAllocFactor = 1.0;
  All.ComovingIntegrationOn = 0;
  All.Time = 0;
  All.Ti_Current = 0;
  All.TreeDomainUpdateFrequency = 0.0;
  All.NumForcesSinceLastDomainDecomp = 1;	/* > TotNumPart*0 : always decompose */
#ifdef PERIODIC
  All.PeriodicBoundariesOn = 1;
#endif
#ifdef PMGRID
  /* pm_periodic.c:59-60 */
  All.Asmth[0] = ASMTH * All.BoxSize / PMGRID;
  All.Rcut[0] = RCUT * All.Asmth[0];
  All.PM_Ti_endstep = -1;	/* never equal to Ti_Current: no PM step in the oracle */
#endif

  FdTimings = fopen("/dev/null", "w");
  FdCPU = FdTimings;
  FdInfo = FdTimings;
  FdEnergy = FdTimings;

  for(i = 0; i < RNDTABLE; i++)
    RndTable[i] = 0.5;

  allocate_commbuffers();	/* allocate.c:19  */
  init_grav_maps();		/* ngravs_core.c:201 -> wire_grav_maps() ngravs.c:64 */
  allocate_memory();		/* allocate.c:98  */
  set_softenings();		/* gravtree.c:468 */
  force_treeallocate((int) (All.TreeAllocFactor * All.MaxPart), All.MaxPart);	/* init.c:151 */
#if defined(PERIODIC) && !defined(PMGRID)
  {
    /* begrun.c:47-49.  lattice_init() (forcetree.c:3611) reads its tables from / writes them to lattice_spc_table_*.dat in the working
     * directory (forcetree.c:3637-3745); G2REF_CACHE_DIR (set by refrun.py to oracle/_ref) keeps that cache next to the libraries. */
    char cwd[4096];
    const char *dir = getenv("G2REF_CACHE_DIR");
    int moved = dir && getcwd(cwd, sizeof(cwd)) && chdir(dir) == 0;
    lattice_init();
    if(moved && chdir(cwd) != 0)
      return -2;
  }
#endif

  g2ref_ready = 1;
  return 0;
}

/* pos n*3, vel n*3 (may be NULL), mass n, type n.  ID := index in this call, so that the
 * permutation applied by peano_hilbert_order() can be read back from the IDs. */
int g2ref_load(int n, const double *pos, const double *vel, const double *mass, const int *type)
{
  int i, k;

  if(!g2ref_ready || n > All.MaxPart)
    return -1;
  NumPart = n;
  N_gas = 0;
  All.TotNumPart = n;
  for(i = 0; i < n; i++)
    {
      memset(&P[i], 0, sizeof(struct particle_data));
      for(k = 0; k < 3; k++)
	{
	  P[i].Pos[k] = pos[3 * i + k];
	  P[i].Vel[k] = vel ? vel[3 * i + k] : 0;
	}
      P[i].Mass = mass[i];
      P[i].Type = type[i];
      P[i].ID = i;
      P[i].Ti_endstep = 0;
      P[i].Ti_begstep = 0;
      P[i].GravCost = 0;
      P[i].OldAcc = 0;
      if(type[i] == 0)
	N_gas++;
    }
  for(i = 0; i < N_gas; i++)
    if(P[i].Type != 0)
      return -2;		/* gas must sit at the head of P[] (peano.c:47-67) */
  All.TotN_gas = N_gas;
  if(N_gas)
    memset(SphP, 0, sizeof(struct sph_particle_data) * N_gas);
  NumForceUpdate = n;
  All.NumForcesSinceLastDomainDecomp = 1;
  return 0;
}

void g2ref_domain(void)
{
  double t0 = wall();
  All.NumForcesSinceLastDomainDecomp = 1 + All.TotNumPart;
  domain_Decomposition();
  t_domain = wall() - t0;
}

/* like domain_Decomposition() but keeps Key[] alive: the reference frees Key/KeySorted at the
 * end of domain_Decomposition (domain.c:149-150), so for key parity we recompute them with the
 * reference's own key function on the reference's own Domain* values. */
void g2ref_keys(long long *keys)
{
  int i;
  for(i = 0; i < NumPart; i++)
    keys[i] = peano_hilbert_key((P[i].Pos[0] - DomainCorner[0]) * DomainFac,
				(P[i].Pos[1] - DomainCorner[1]) * DomainFac,
				(P[i].Pos[2] - DomainCorner[2]) * DomainFac, BITS_PER_DIMENSION);
}

long long g2ref_peano_key(int x, int y, int z, int bits)
{
  return peano_hilbert_key(x, y, z, bits);
}

void g2ref_set_active(const int *active)	/* active[i]!=0 <=> Ti_endstep == Ti_Current */
{
  int i;
  NumForceUpdate = 0;
  for(i = 0; i < NumPart; i++)
    {
      P[i].Ti_endstep = active[i] ? All.Ti_Current : All.Ti_Current + 1;
      if(active[i])
	NumForceUpdate++;
    }
}

void g2ref_set_opening(double theta, double errtolforceacc, int type_of_criterion)
{
  All.ErrTolTheta = theta;
  All.ErrTolForceAcc = errtolforceacc;
  All.TypeOfOpeningCriterion = type_of_criterion;
}

void g2ref_force_rebuild(void)
{
  TreeReconstructFlag = 1;
}

void g2ref_gravity(void)
{
  double c0 = All.CPU_TreeConstruction, w0 = All.CPU_TreeWalk, t0 = wall();
  gravity_tree();
  t_gravity = wall() - t0;
  t_build = All.CPU_TreeConstruction - c0;
  t_walk = All.CPU_TreeWalk - w0;
}

/* only the tree build (force_treebuild, forcetree.c:61) */
int g2ref_treebuild(void)
{
  double t0 = wall();
  int n = force_treebuild(NumPart);
  t_build = wall() - t0;
  TreeReconstructFlag = 0;
  return n;
}

/* ---- multi-threaded walk of the UNMODIFIED per-target reference function.  The per-target
 * functions read only global tree/particle state and write only P[target] (forcetree.c:1590-1596,
 * 2036-2049), so slicing targets over threads mirrors the GPU scheme (whole tree, slice of targets).
 * Used as the K-core CPU baseline. */
struct slice { int lo, hi; double cost; };
static void *walk_slice(void *arg)
{
  struct slice *s = (struct slice *) arg;
  int i;
  double ewald = 0;
  s->cost = 0;
  for(i = s->lo; i < s->hi; i++)
    if(P[i].Ti_endstep == All.Ti_Current)
      {
#ifndef PMGRID
	s->cost += force_treeevaluate(i, 0, &ewald);
#else
	s->cost += force_treeevaluate_shortrange(i, 0);
#endif
      }
  return NULL;
}

/* walks targets [lo,hi) with nthreads threads; returns wall seconds, *cost = sum of interactions.
 * Results are the pre-G accelerations in P[].GravAccel (as inside gravity_tree before line 338). */
double g2ref_walk_threads(int lo, int hi, int nthreads, double *cost)
{
  pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
  struct slice *sl = malloc(sizeof(struct slice) * nthreads);
  int t, n = hi - lo;
  double t0 = wall(), dt;
  for(t = 0; t < nthreads; t++)
    {
      sl[t].lo = lo + (int) ((long long) n * t / nthreads);
      sl[t].hi = lo + (int) ((long long) n * (t + 1) / nthreads);
      pthread_create(&th[t], NULL, walk_slice, &sl[t]);
    }
  *cost = 0;
  for(t = 0; t < nthreads; t++)
    {
      pthread_join(th[t], NULL);
      *cost += sl[t].cost;
    }
  dt = wall() - t0;
  free(sl);
  free(th);
  return dt;
}

void g2ref_timings(double *out)
{
  out[0] = t_domain;
  out[1] = t_gravity;
  out[2] = t_build;
  out[3] = t_walk;
  out[4] = All.CPU_Peano;
  out[5] = All.CPU_Domain;
}

/* ---- read-back ---- */
int g2ref_numpart(void) { return NumPart; }
int g2ref_maxpart(void) { return All.MaxPart; }
int g2ref_maxnodes(void) { return MaxNodes; }
int g2ref_numnodes(void) { return Numnodestree; }
int g2ref_ntopnodes(void) { return NTopnodes; }
int g2ref_ntopleaves(void) { return NTopleaves; }

void g2ref_get_domain(double *out)	/* corner[3], center[3], len, fac */
{
  int k;
  for(k = 0; k < 3; k++)
    {
      out[k] = DomainCorner[k];
      out[3 + k] = DomainCenter[k];
    }
  out[6] = DomainLen;
  out[7] = DomainFac;
}

void g2ref_get_particles(double *pos, double *mass, int *type, unsigned int *id, double *acc, float *gravcost,
			 double *oldacc)
{
  int i, k;
  for(i = 0; i < NumPart; i++)
    {
      for(k = 0; k < 3; k++)
	{
	  if(pos)
	    pos[3 * i + k] = P[i].Pos[k];
	  if(acc)
	    acc[3 * i + k] = P[i].GravAccel[k];
	}
      if(mass)
	mass[i] = P[i].Mass;
      if(type)
	type[i] = P[i].Type;
      if(id)
	id[i] = P[i].ID;
      if(gravcost)
	gravcost[i] = P[i].GravCost;
      if(oldacc)
	oldacc[i] = P[i].OldAcc;
    }
}

void g2ref_set_vel(const double *vel)	/* velocities of the particles in their CURRENT order */
{
  int i, k;
  for(i = 0; i < NumPart; i++)
    for(k = 0; k < 3; k++)
      P[i].Vel[k] = vel[3 * i + k];
}

void g2ref_set_oldacc(const double *oldacc)
{
  int i;
  for(i = 0; i < NumPart; i++)
    P[i].OldAcc = oldacc[i];
}

/* TopNodes[] as 7 long long per node: Daughter,Pstart,Blocks,Leaf,Size,StartKey,Count;
 * DomainNodeIndex[NTopleaves] */
void g2ref_get_topnodes(long long *tn, int *domain_node_index)
{
  int i;
  for(i = 0; i < NTopnodes; i++)
    {
      tn[7 * i + 0] = TopNodes[i].Daughter;
      tn[7 * i + 1] = TopNodes[i].Pstart;
      tn[7 * i + 2] = TopNodes[i].Blocks;
      tn[7 * i + 3] = TopNodes[i].Leaf;
      tn[7 * i + 4] = TopNodes[i].Size;
      tn[7 * i + 5] = TopNodes[i].StartKey;
      tn[7 * i + 6] = TopNodes[i].Count;
    }
  for(i = 0; i < NTopleaves; i++)
    domain_node_index[i] = DomainNodeIndex[i];
}

/* tree after force_treebuild (post-moment form of the union): per node k in [0,Numnodestree):
 * len, center[3] -> geom[4k..]; s[3][D] -> s[(3k+j)*D+g]; mass[D]; vs[3][D]; bitflags,sibling,nextnode,father -> link[4k..] */
void g2ref_get_tree(double *geom, double *s, double *mass, double *vs, int *link, int *nextnode, int *father)
{
  int k, j, g;
  for(k = 0; k < Numnodestree; k++)
    {
      struct NODE *nop = &Nodes[All.MaxPart + k];
      geom[4 * k + 0] = nop->len;
      for(j = 0; j < 3; j++)
	geom[4 * k + 1 + j] = nop->center[j];
      for(j = 0; j < 3; j++)
	for(g = 0; g < N_GRAVS; g++)
	  {
	    s[(3 * k + j) * N_GRAVS + g] = nop->u.d.s[j][g];
	    if(vs)
	      vs[(3 * k + j) * N_GRAVS + g] = Extnodes[All.MaxPart + k].vs[j][g];
	  }
      for(g = 0; g < N_GRAVS; g++)
	mass[k * N_GRAVS + g] = nop->u.d.mass[g];
      link[4 * k + 0] = nop->u.d.bitflags;
      link[4 * k + 1] = nop->u.d.sibling;
      link[4 * k + 2] = nop->u.d.nextnode;
      link[4 * k + 3] = nop->u.d.father;
    }
  for(k = 0; k < NumPart; k++)
    {
      nextnode[k] = Nextnode[k];
      father[k] = Father[k];
    }
}

/* Nodes[].u.d.Nparticles[g] (allvars.h:645-648); returns 0 when the variant was built without -DNGRAVS_ACCUMULATOR */
int g2ref_get_nparticles(long long *out)
{
#ifdef NGRAVS_ACCUMULATOR
  int k, g;
  for(k = 0; k < Numnodestree; k++)
    for(g = 0; g < N_GRAVS; g++)
      out[k * N_GRAVS + g] = Nodes[All.MaxPart + k].u.d.Nparticles[g];
  return 1;
#else
  (void) out;
  return 0;
#endif
}

/* short-range table shortrange_fourier_force[tgt][src][NTAB] (forcetree.c:33, filled 3274-3354) */
int g2ref_get_srtable(double *out)
{
  return g2ref_srtable_copy(out);
}

int g2ref_lattice_copy(double *out);
/* fcorrx/y/z[tgt][src][EN+1][EN+1][EN+1] after lattice_init (already divided by BoxSize^2, forcetree.c:3757-3761) as double:
 * out[((c * D + tgt) * D + src) * (EN+1)^3 + ...], c = 0, 1, 2.  Returns EN + 1, or 0 where the variant has no such tables. */
int g2ref_get_lattice_tables(double *out)
{
  return g2ref_lattice_copy(out);
}

int g2ref_get_srpot_table(double *out)
{
  return g2ref_srpot_copy(out);
}

/* ---- tree potential of EVERY particle with the reference's own per-target walk, as the loop of compute_potential()
 * does (potential.c:86-97): force_treeevaluate_potential_shortrange (forcetree.c:2789) under PMGRID, else
 * force_treeevaluate_potential (forcetree.c:2467; its body exists only with -DOUTPUTPOTENTIAL).  The tree must be
 * current (g2ref_gravity / g2ref_treebuild).  out[i] = P[i].Potential straight after the walk: the pre-G sum that still
 * contains the self term (potential.c:250-268 removes it and applies G afterwards).  Returns -1 where the variant
 * has no potential walk. */
struct pslice { int lo, hi; };
static void *pot_slice(void *arg)
{
  struct pslice *s = (struct pslice *) arg;
  int i;
  for(i = s->lo; i < s->hi; i++)
    {
#ifdef PMGRID
      force_treeevaluate_potential_shortrange(i, 0);
#else
      force_treeevaluate_potential(i, 0);
#endif
    }
  return NULL;
}

int g2ref_potential(double *out, int nthreads)
{
#if defined(PMGRID) || defined(OUTPUTPOTENTIAL)
  pthread_t *th;
  struct pslice *sl;
  int t, i;
  if(nthreads < 1)
    nthreads = 1;
#ifdef G2_SHIM_HARNESS
  nthreads = 1;			/* the shim's entry points drive one device context: single caller, like the reference's own loop */
#endif
  th = malloc(sizeof(pthread_t) * nthreads);
  sl = malloc(sizeof(struct pslice) * nthreads);
  for(t = 0; t < nthreads; t++)
    {
      sl[t].lo = (int) ((long long) NumPart * t / nthreads);
      sl[t].hi = (int) ((long long) NumPart * (t + 1) / nthreads);
      pthread_create(&th[t], NULL, pot_slice, &sl[t]);
    }
  for(t = 0; t < nthreads; t++)
    pthread_join(th[t], NULL);
  free(sl);
  free(th);
  for(i = 0; i < NumPart; i++)
    out[i] = P[i].Potential;
  return 0;
#else
  (void) out;
  (void) nthreads;
  return -1;
#endif
}

/* the shipped pair potentials by name (ngravs.c:368, 459, 672, 708, 734): they are compiled in every variant, whereas the
 * PotentialFxns tables exist only under PMGRID / OUTPUTPOTENTIAL -- and the BAM wiring (ngravs.c:172-200) leaves NormedGreensFxns
 * unset, so a PMGRID build with that wiring ends in a null call during set-up and cannot be run. */
double g2ref_named_pot(int which, double pm, double m, double h, double r, long n)
{
  switch (which)
    {
    case 0: return bambam_pot(pm, m, h, r, n);
    case 1: return sourcebaryonbam_pot(pm, m, h, r, n);
    case 2: return sourcebambaryon_pot(pm, m, h, r, n);
    case 3: return newtonian_pot(pm, m, h, r, n);
    case 4: return plummer_pot(pm, m, h, r, n);
    default: return 0.0 / 0.0;
    }
}

/* potential pair-law probes (allvars.h:147-148) */
double g2ref_potfxn(int tgt, int src, double pm, double m, double h, double r, long n)
{
#if defined(PMGRID) || defined(OUTPUTPOTENTIAL)
  return (*PotentialFxns[tgt][src]) (pm, m, h, r, n);
#else
  return 0.0 / 0.0;
#endif
}

double g2ref_potspline(int tgt, int src, double pm, double m, double h, double r, long n)
{
#if defined(PMGRID) || defined(OUTPUTPOTENTIAL)
  return (*PotentialSplines[tgt][src]) (pm, m, h, r, n);
#else
  return 0.0 / 0.0;
#endif
}

void g2ref_get_pm_split(double *out)
{
#ifdef PMGRID
  out[0] = All.Asmth[0];
  out[1] = All.Rcut[0];
#else
  out[0] = out[1] = 0;
#endif
}

/* pair-law probes through the reference's own function-pointer tables (allvars.h:134-136) */
double g2ref_accel(int tgt, int src, double pm, double m, double r2, double r, long n)
{
  return (*AccelFxns[tgt][src]) (pm, m, r2, r, n);
}

double g2ref_spline(int tgt, int src, double pm, double m, double h, double r, long n)
{
  return (*AccelSplines[tgt][src]) (pm, m, h, r, n);
}

void g2ref_get_softening(double *out)
{
  int i;
  for(i = 0; i < 6; i++)
    out[i] = All.ForceSoftening[i];
}

/* direct summation with the reference's own routine (forcetree.c:3428), FORCETEST builds only.
 * Non-periodic variants only (PERIODIC adds lattice_corr which needs lattice_init tables). */
int g2ref_direct(int ntargets, const int *targets, double *acc)
{
#if defined(FORCETEST) && !defined(PERIODIC)
  int t, k;
  for(t = 0; t < ntargets; t++)
    {
      force_treeevaluate_direct(targets[t], 0);
      for(k = 0; k < 3; k++)
	acc[3 * t + k] = P[targets[t]].GravAccelDirect[k];
    }
  return 0;
#else
  return -1;
#endif
}

/* symbols the hot-path closure leaves unresolved: never reached by the oracle */
#ifdef PERIODIC
void do_box_wrapping(void)	/* predict.c:106 is outside the closure; the oracle is only fed in-box positions */
{
  int i, j;
  for(i = 0; i < NumPart; i++)
    for(j = 0; j < 3; j++)
      if(!(P[i].Pos[j] >= 0 && P[i].Pos[j] < All.BoxSize))
	{
	  fprintf(stderr, "g2ref: particle %d outside the periodic box; oracle inputs must be pre-wrapped\n", i);
	  abort();
	}
}
#endif
size_t my_fread(void *ptr, size_t size, size_t nmemb, FILE * stream) { return fread(ptr, size, nmemb, stream); }
size_t my_fwrite(void *ptr, size_t size, size_t nmemb, FILE * stream) { return fwrite(ptr, size, nmemb, stream); }

/* ---- periodic PM long-range force (pm_periodic.c:204-790), single rank through the rfftwnd_mpi stand-in ----
 * P[] must be in the order domain_Decomposition() left it (species blocks, NgravLocal[] set: pm_periodic.c:251-254).
 * out: GravPM of every particle (n x 3, as double), in the current order of P[]. */
static int pm_ready = 0;
int g2ref_pmforce(double *out)
{
#if defined(PMGRID) && defined(PERIODIC)
  int i, k;
  if(!pm_ready)
    {
      pm_init_periodic();	/* pm_periodic.c:53 */
      pm_ready = 1;
    }
  pmforce_periodic();		/* frees and re-allocates the tree storage like the reference does */
  TreeReconstructFlag = 1;
  for(i = 0; i < NumPart; i++)
    for(k = 0; k < 3; k++)
      out[3 * i + k] = P[i].GravPM[k];
  return 0;
#else
  (void) out;
  return -1;
#endif
}

/* ---- periodic PM long-range potential (pmpotential_periodic, pm_periodic.c:798-1300), single rank.  P[].Potential is zeroed first, so
 * out[i] is what the routine ADDS to the potential of particle i (current order of P[], species blocks). ---- */
int g2ref_pmpotential(double *out)
{
#if defined(PMGRID) && defined(PERIODIC)
  int i;
  if(!pm_ready)
    {
      pm_init_periodic();
      pm_ready = 1;
    }
  for(i = 0; i < NumPart; i++)
    P[i].Potential = 0;
  pmpotential_periodic();	/* frees and re-allocates the tree storage like pmforce_periodic */
  TreeReconstructFlag = 1;
  for(i = 0; i < NumPart; i++)
    out[i] = P[i].Potential;
  return 0;
#else
  (void) out;
  return -1;
#endif
}

/* lattice_pot_corr (forcetree.c:3895) called directly, and the table it interpolates.  force_treeevaluate_potential, its only caller,
 * does not compile in the reference (see the Makefile), so the look-up is what can be pinned. */
int g2ref_potcorr_copy(double *out);
int g2ref_get_potcorr(double *out) { return g2ref_potcorr_copy(out); }
double g2ref_lattice_pot_corr(double dx, double dy, double dz, int target, int source)
{
#if defined(PERIODIC) && !defined(PMGRID)
  return lattice_pot_corr(dx, dy, dz, target, source);
#else
  return 0.0 / 0.0;
#endif
}

/* gravity_forcetest() (gravtree_forcetest.c:28) after a gravity_tree(): direct sums of the selected particles (all active ones with
 * FORCETEST = 1.0, RndTable = 0.5).  out: P[].GravAccelDirect (n x 3, post-G), current order.  forcetest.txt goes to the working directory. */
int g2ref_forcetest(double *out)
{
#if defined(FORCETEST) && !defined(PERIODIC)
  int i, k;
  gravity_forcetest();
  for(i = 0; i < NumPart; i++)
    for(k = 0; k < 3; k++)
      out[3 * i + k] = P[i].GravAccelDirect[k];
  return 0;
#else
  (void) out;
  return -1;
#endif
}
