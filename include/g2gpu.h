/* g2gpu.h — C ABI of the B200-native tree-gravity hot path for Gadget-2-ngravs.
 *
 * This is the drop-in boundary: plain C, pointers and sizes only.  The reference's C entry points
 * (gravity_tree gravtree.c:27, force_treebuild forcetree.c:61, force_treeallocate forcetree.c:3176,
 * force_treefree forcetree.c:3411, peano_hilbert_order peano.c:36, peano_hilbert_key peano.c:356,
 * init_grav_maps ngravs_core.c:201) keep their signatures in the host shim (host/ *.c) and call the
 * functions below; see INTEGRATION.md.  There is no CPU fallback: every call that needs the GPU fails
 * with G2GPU_ERR_CUDA when no sm_100 device is present.
 *
 * Index conventions are the reference's (forcetree.c:76-92): particles are [0, NumPart), tree nodes
 * are [MaxPart, MaxPart+MaxNodes).  All functions return 0 on success or a negative G2GPU_ERR_* code
 * (the host shim turns non-zero into endrun(code), endrun.c:24); g2gpu_last_error() gives the text.
 */
#ifndef G2GPU_H
#define G2GPU_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define G2GPU_MAX_GRAVS 6	/* N_GRAVS upper bound: 6 particle types (allvars.h:132) */
#define G2GPU_NTYPES 6

enum g2gpu_error
{
  G2GPU_OK = 0,
  G2GPU_ERR_CUDA = -1,		/* CUDA runtime failure or no usable device */
  G2GPU_ERR_ARG = -2,		/* invalid argument / call order */
  G2GPU_ERR_NOMEM = -3,		/* device or host allocation failed */
  G2GPU_ERR_MAXNODES = -4,	/* more tree nodes than MaxNodes (forcetree.c:249-255 -> endrun(1)) */
  G2GPU_ERR_TREE_DEPTH = -5,	/* more than 8 particles not separable within the supported tree depth (21 levels) */
  G2GPU_ERR_TOPNODES = -6,	/* top-level tree larger than MAXTOPNODES (domain.c:1049 -> endrun(13213)) */
  G2GPU_ERR_LAW = -7,		/* unknown / unwired pair force law (ngravs_core.c:326-365) */
  G2GPU_ERR_STATE = -8		/* call made before its prerequisite stage */
};

/* Pair force laws shipped in ngravs.c, by registry id.  AccelFxns[tgt][src] takes (pm, m, r2, r, N),
 * AccelSplines[tgt][src] takes (pm, m, h, r, N) (allvars.h:134-136, ngravs.c:332-861). */
enum g2gpu_law
{
  G2GPU_LAW_NONE = 0,		/* none()            ngravs.c:344 */
  G2GPU_LAW_NEWTONIAN = 1,	/* newtonian()       ngravs.c:351 */
  G2GPU_LAW_NEG_NEWTONIAN = 2,	/* neg_newtonian()   ngravs.c:359 */
  G2GPU_LAW_YUKAWA = 3,		/* yukawa()          ngravs.c:856, param[0] = YUKAWA_IMASS/BoxSize */
  G2GPU_LAW_COLOYUK = 4,	/* coloyuk()         ngravs.c:826 */
  G2GPU_LAW_BAMBAM = 5,		/* bambam()          ngravs.c:495, param[1] = BAM_EPSILON */
  G2GPU_LAW_SOURCEBAMBARYON = 6,	/* sourcebambaryon() ngravs.c:590 */
  G2GPU_LAW_SOURCEBARYONBAM = 7,	/* sourcebaryonbam() ngravs.c:646 */
  G2GPU_SPLINE_NONE = 16,	/* none()            as a spline */
  G2GPU_SPLINE_PLUMMER = 17,	/* plummer()         ngravs.c:420 */
  G2GPU_SPLINE_NEG_PLUMMER = 18,	/* neg_plummer()     ngravs.c:438 */
  G2GPU_SPLINE_BAMBAM = 19,	/* bambam_spline()   ngravs.c:531 */
  G2GPU_SPLINE_SOURCEBAMBARYON = 20,	/* sourcebambaryon_spline() ngravs.c:562 */
  G2GPU_SPLINE_SOURCEBARYONBAM = 21	/* sourcebaryonbam_spline() ngravs.c:616 */
};

/* Pair potentials shipped in ngravs.c for the tree potential walks: PotentialFxns[tgt][src](pm, m, h, r, N) for r >= h,
 * PotentialSplines[tgt][src](pm, m, h, r, N) for r < h (allvars.h:147-148).  The BAM potentials take BAM_EPSILON from params[1] of
 * g2gpu_set_laws; the reference wires them as function AND as "spline" (ngravs.c:198-200). */
enum g2gpu_potlaw
{
  G2GPU_POT_NONE = 32,		/* none()              ngravs.c:344 */
  G2GPU_POT_NEWTONIAN = 33,	/* newtonian_pot()     ngravs.c:368 */
  G2GPU_POT_NEG_NEWTONIAN = 34,	/* neg_newtonian_pot() ngravs.c:375 */
  G2GPU_POT_BAMBAM = 35,	/* bambam_pot()        ngravs.c:672 */
  G2GPU_POT_SOURCEBARYONBAM = 36,	/* sourcebaryonbam_pot() ngravs.c:708 */
  G2GPU_POT_SOURCEBAMBARYON = 37,	/* sourcebambaryon_pot() ngravs.c:734 */
  G2GPU_POTSPLINE_NONE = 48,	/* none()              as a spline */
  G2GPU_POTSPLINE_PLUMMER = 49,	/* plummer_pot()       ngravs.c:459 */
  G2GPU_POTSPLINE_NEG_PLUMMER = 50,	/* neg_plummer_pot()   ngravs.c:476 */
  G2GPU_POTSPLINE_BAMBAM = 51,	/* the three BAM potentials in the spline slot */
  G2GPU_POTSPLINE_SOURCEBARYONBAM = 52,
  G2GPU_POTSPLINE_SOURCEBAMBARYON = 53
};

/* k-space Green's functions of the periodic PM force shipped in ngravs.c (GreensFxns[source][target], allvars.h:140; mesh units). */
enum g2gpu_greens
{
  G2GPU_GREENS_NONE = 0,	/* none()      ngravs.c:344 */
  G2GPU_GREENS_NEWTON = 1,	/* pgdelta()   ngravs.c:390: 1/k^2 */
  G2GPU_GREENS_NEG_NEWTON = 2,	/* neg_pgdelta() ngravs.c:407 */
  G2GPU_GREENS_YUKAWA = 3,	/* pgyukawa()  ngravs.c:869, param = YUKAWA_IMASS/(2 pi) */
  G2GPU_GREENS_COLOYUK = 4	/* pgcoloyuk() ngravs.c:831 */
};

/* Compile-time switches of the reference (Makefile.reference:49-135) become run-time configuration. */
typedef struct g2gpu_config
{
  int device;			/* CUDA device ordinal */
  int n_gravs;			/* N_GRAVS (D), 1..G2GPU_MAX_GRAVS */
  int periodic;			/* PERIODIC: NEAREST() wrap in the walk (forcetree.c:43) */
  int shortrange;		/* PMGRID != 0: force_treeevaluate_shortrange instead of force_treeevaluate */
  int ntab;			/* NTAB, length of one short-range table (2048) */
  int unequal_softenings;	/* UNEQUALSOFTENINGS */
  int max_part;			/* All.MaxPart: index of the first tree node */
  int max_nodes;		/* MaxNodes = TreeAllocFactor*MaxPart (init.c:151) */
  int rank;			/* this process's rank among the GPUs sharing the particle set */
  int nranks;			/* number of such GPUs; targets are split into nranks equal PH slices */
} g2gpu_config;

/* Per-call parameters of a force computation: the fields of `All` that gravity_tree() and the walks read
 * (gravtree.c:27-460, forcetree.c:1244-2052). */
typedef struct g2gpu_walk_params
{
  double theta;			/* All.ErrTolTheta; 0 selects the relative criterion (forcetree.c:1439) */
  double errtol_force_acc;	/* All.ErrTolForceAcc */
  double boxsize;		/* All.BoxSize (PERIODIC) */
  double G;			/* All.G */
  double asmth;			/* All.Asmth[0]  (short-range) */
  double rcut;			/* All.Rcut[0]   (short-range) */
  double pos_fac_pre_g;		/* added as fac*Pos before the G scaling (gravtree.c:304-316), normally 0 */
  double pos_fac_post_g;	/* added as fac*Pos after the G scaling  (gravtree.c:346-358), normally 0 */
  int use_gravpm;		/* OldAcc includes GravPM/G (gravtree.c:321-325) */
} g2gpu_walk_params;

/* Parameters of the periodic PM long-range force: PMGRID, the fields of `All` pmforce_periodic() reads (pm_periodic.c:204-237) and
 * the k-space Green's function of every ORDERED species pair, indexed [source * D + target] like GreensFxns[nA][nB] (:512). */
typedef struct g2gpu_pm_params
{
  int pmgrid;			/* PMGRID (mesh cells per dimension) */
  double boxsize;		/* All.BoxSize */
  double asmth;			/* All.Asmth[0] = ASMTH * BoxSize / PMGRID (pm_periodic.c:59) */
  double G;			/* All.G */
  int greens_id[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];	/* enum g2gpu_greens */
  double greens_par[G2GPU_MAX_GRAVS * G2GPU_MAX_GRAVS];
} g2gpu_pm_params;

typedef struct g2gpu_ctx g2gpu_ctx;

/* ---- life cycle: force_treeallocate (forcetree.c:3176) / force_treefree (forcetree.c:3411) ---- */
int g2gpu_create(g2gpu_ctx **ctx, const g2gpu_config *cfg);
void g2gpu_destroy(g2gpu_ctx *ctx);
const char *g2gpu_last_error(void);
int g2gpu_device_count(void);

/* ---- plug-in tables: init_grav_maps/wire_grav_maps (ngravs_core.c:201, ngravs.c:64) ---- */
/* TypeToGrav[6] (allvars.h:132) and All.ForceSoftening[6] (gravtree.c:514). */
int g2gpu_set_species(g2gpu_ctx *ctx, const int type_to_grav[G2GPU_NTYPES], const double force_softening[G2GPU_NTYPES]);
/* accel_id / spline_id: D*D registry ids indexed [target][source] (ngravs.c:75-76); params: 4 doubles per pair. */
int g2gpu_set_laws(g2gpu_ctx *ctx, const int *accel_id, const int *spline_id, const double *params);
/* shortrange_fourier_force[target][source][NTAB] (forcetree.c:33, filled 3274-3354), as double. */
int g2gpu_set_srtable(g2gpu_ctx *ctx, const double *table);

/* ---- particle upload (struct particle_data, allvars.h:546-581) ---- */
/* SoA, float32: pos[3n], mass[n], type[n], and optional oldacc[n], vel[3n], gravpm[3n], active[n]
 * (active[i] != 0 <=> P[i].Ti_endstep == All.Ti_Current, gravtree.c:113; NULL = all active). */
int g2gpu_upload(g2gpu_ctx *ctx, int npart, const float *pos, const float *mass, const int *type,
		 const float *oldacc, const float *vel, const float *gravpm, const int *active);
/* Same from the reference's AoS: base pointer, stride = sizeof(struct particle_data), byte offsets of
 * Pos, Mass, Type, OldAcc, Vel, GravPM (-1 if absent), Ti_endstep; float_bytes = sizeof(FLOAT). */
int g2gpu_upload_aos(g2gpu_ctx *ctx, int npart, const void *P, size_t stride, int float_bytes, int off_pos,
		     int off_mass, int off_type, int off_oldacc, int off_vel, int off_gravpm, int off_ti_endstep,
		     int ti_current);
/* The device-side particle record (32 bytes): what the library builds from the SoA/AoS uploads, and what a caller that
 * fills device memory itself (multi-GPU all-gather) must provide. */
typedef struct g2gpu_particle
{
  float pos[3];			/* P[].Pos  */
  float mass;			/* P[].Mass */
  int type;			/* P[].Type, 0..5 */
  float oldacc;			/* P[].OldAcc */
  int active;			/* != 0 <=> P[].Ti_endstep == All.Ti_Current */
  int pad;
} g2gpu_particle;

/* Multi-GPU: *records = the library's own device input buffer (max_part records), to be filled by the caller (e.g. as the
 * receive buffer of an NCCL all-gather); call g2gpu_inputs_ready(npart) afterwards. */
int g2gpu_input_buffers(g2gpu_ctx *ctx, int npart, void **records);
int g2gpu_inputs_ready(g2gpu_ctx *ctx, int npart);
/* Zero-copy variant: use a caller-owned DEVICE array of npart g2gpu_particle records (16-byte aligned, e.g. the output of
 * an NCCL all-gather) as the particle input.  It must stay valid and unchanged until g2gpu_domain() has run.  The next
 * g2gpu_upload()/g2gpu_inputs_ready() unbinds it. */
int g2gpu_bind_inputs(g2gpu_ctx *ctx, int npart, void *records);

/* ---- stage 1: domain_findExtent + key loop + top tree + peano_hilbert_order
 *      (domain.c:882-924, 933-1138; peano.c:36-185, 356-398) ---- */
int g2gpu_domain(g2gpu_ctx *ctx);
int g2gpu_get_domain(g2gpu_ctx *ctx, double out[8]);	/* DomainCorner[3], DomainCenter[3], DomainLen, DomainFac */
int g2gpu_get_keys(g2gpu_ctx *ctx, long long *keys);	/* Key[] in CURRENT particle order */
int g2gpu_get_order(g2gpu_ctx *ctx, int *perm);	/* perm[i] = index at upload time of the particle now at i */
/* TopNodes[] (allvars.h:252-262) as 4 arrays + DomainNodeIndex[NTopleaves]; pass NULL to skip an output. */
int g2gpu_get_topnodes(g2gpu_ctx *ctx, int *ntopnodes, int *ntopleaves, int *daughter, int *leaf,
		       long long *startkey, long long *size, long long *count, int *domain_node_index);

/* ---- stage 2: force_treebuild (forcetree.c:61-281, 451-743, 954-996) ---- */
int g2gpu_treebuild(g2gpu_ctx *ctx, int *numnodes);
/* Host mirror in the reference's numbering and layout semantics (struct NODE after the moment pass,
 * allvars.h:618-660): per node k (index MaxPart+k): len[k], center[3k], s[(3k+j)*D+g], mass[k*D+g],
 * bitflags/sibling/nextnode/father[k]; per particle Nextnode[i], Father[i].  NULL skips an output. */
int g2gpu_download_tree(g2gpu_ctx *ctx, float *len, float *center, float *s, float *mass, int *bitflags,
			int *sibling, int *nextnode, int *father, int *p_nextnode, int *p_father);

/* Dynamic tree update: between two constructions the reference lets its tree follow the particles on the host (node drift
 * predict.c:79-91, node kicks timestep.c:329-344, force_update_len forcetree.c:1005-1122).  After a fresh g2gpu_upload*() of the SAME
 * particles in the SAME order as at the last g2gpu_treebuild, this call re-attaches the tree of that construction (instead of
 * g2gpu_domain + g2gpu_treebuild): particle records are refreshed from the upload, and every node takes the host's values
 * len[k] = Nodes[MaxPart+k].len and s[(3k+j)*D+g] = Nodes[MaxPart+k].u.d.s[j][g] (reference numbering, as g2gpu_download_tree
 * delivers them).  g2gpu_walk then sees what the reference's walk sees. */
int g2gpu_update_tree(g2gpu_ctx *ctx, const float *len, const float *s);

/* struct extNODE (allvars.h:667-677): vs[(3k+j)*D+g] = Extnodes[MaxPart+k].vs[j][g], the centre-of-mass velocity per
 * species (forcetree.c:563-567, 617-619, 667-694); needs the vel argument of g2gpu_upload.  hmax (SPH) is not computed. */
int g2gpu_download_extnodes(g2gpu_ctx *ctx, float *vs);

/* -DNGRAVS_ACCUMULATOR (allvars.h:645-648; forcetree.c:557-559, 621-623): nparticles[k*D+g] = Nodes[MaxPart+k].u.d.Nparticles[g],
 * the number of particles of species g below node k.  With the option "accumulator" the walk hands it to the pair laws as N for
 * particle-node interactions (forcetree.c:1563-1577); particle-particle interactions always pass 1. */
int g2gpu_download_nparticles(g2gpu_ctx *ctx, long long *nparticles);

/* ---- stage 3: the walk + gravity_tree epilogue (gravtree.c:102-358; forcetree.c:1244-2052) ---- */
int g2gpu_walk(g2gpu_ctx *ctx, const g2gpu_walk_params *wp);
/* force_treeevaluate_direct (forcetree.c:3428-3548), the accuracy oracle of gravity_forcetest (gravtree_forcetest.c:28-356): FP64
 * direct summation over all particles for ntargets particles given by their index in CURRENT order (after g2gpu_domain);
 * acc[3*i+k] is the pre-G acceleration (G = 1).  Under the TreePM split (config.shortrange and wp->asmth > 0) it is the short-range
 * force the tree walk approximates; periodic boxes take the nearest image.  With the option "direct_ewald" and wp->asmth == 0 a periodic
 * box gets the exact lattice (Ewald) correction of all images added (what the reference tabulates and interpolates in lattice_corr,
 * forcetree.c:3803, for gravity_forcetest), i.e. the complete periodic Newtonian force the TreePM sum (tree + PM) approximates. */
int g2gpu_direct(g2gpu_ctx *ctx, const g2gpu_walk_params *wp, int ntargets, const int *targets, double *acc);

/* ---- lattice-sum correction of a periodic box WITHOUT PM (config.periodic && !config.shortrange):
 *      force_treeevaluate_lattice_correction (forcetree.c:2077-2455), the second walk force_treeevaluate runs for every target
 *      (forcetree.c:1606-1608).  Once tables are set, g2gpu_walk runs that walk as well: the correction is added to GravAccel before the
 *      gravity_tree epilogue and its interaction count to GravCost (forcetree.c:2435-2438; counters[1] stays the tree walk's own sum,
 *      like costtotal in gravtree.c:120).  Without tables the force is the nearest-image tree force.
 *      fcorr[((c * D + target) * D + source) * (en+1)^3 + (i * (en+1) + j) * (en+1) + k], c = 0,1,2: the reference's fcorrx/y/z AFTER
 *      lattice_init (forcetree.c:3611-3790), i.e. divided by BoxSize^2; en = NGRAVS_EN (64).  NULL removes the tables. ---- */
int g2gpu_set_lattice_tables(g2gpu_ctx *ctx, int en, const double *fcorr);
/* The table lattice_init computes for the stock wiring: ewald_force (ngravs.c:1170-1236) at x = 0.5 (i,j,k)/en, FP64 on the device,
 * DIMENSIONLESS (divide by BoxSize^2 for g2gpu_set_lattice_tables); out[c * (en+1)^3 + ...], c = 0,1,2. */
int g2gpu_make_ewald_table(g2gpu_ctx *ctx, int en, double *out);
/* The potential of a periodic box without PM: force_treeevaluate_potential adds mass * lattice_pot_corr(dx, dy, dz, target, source) to every
 * term (forcetree.c:2736-2738, 2765-2767; look-up forcetree.c:3895-3941).  potcorr[(target * D + source) * (en+1)^3 + (i * (en+1) + j) *
 * (en+1) + k]: the reference's potcorr AFTER lattice_init (forcetree.c:3697-3702, 3759), i.e. LatticePotential at x = 0.5 (i,j,k)/en,
 * LatticeZero at the origin, divided by BoxSize.  Required by g2gpu_potential when config.periodic && !config.shortrange; NULL removes it. */
int g2gpu_set_lattice_pot_tables(g2gpu_ctx *ctx, int en, const double *potcorr);
/* The table lattice_init computes for the stock wiring: ewald_psi (ngravs.c:761-816) at x = 0.5 (i,j,k)/en, FP64 on the device,
 * latticezero (LatticeZero[l][m], 2.8372975 in ngravs.c:133) at the origin; NOT yet divided by BoxSize; out[(en+1)^3]. */
int g2gpu_make_ewald_pot_table(g2gpu_ctx *ctx, int en, double latticezero, double *out);

/* ---- tree potential of compute_potential (potential.c:22-354): force_treeevaluate_potential_shortrange
 *      (forcetree.c:2789-3163) when config.shortrange is set, else force_treeevaluate_potential (forcetree.c:2467-2776; non-periodic
 *      only -- the periodic variant needs the lattice-sum tables of lattice_pot_corr, which are not built: G2GPU_ERR_ARG).
 *      g2gpu_potential walks the current tree (after g2gpu_treebuild or g2gpu_update_tree) for EVERY particle (potential.c:86) with the
 *      opening criterion of wp (theta, errtol_force_acc with the uploaded OldAcc) and stores what the reference's walk stores in
 *      P[].Potential (forcetree.c:3158): the pre-G sum, self term included; the caller applies potential.c:250-270 (self term, G).
 *      The node terms carry the short-range table term only with the option "accumulator" (forcetree.c:3134-3140).  With nranks > 1
 *      a rank walks its 1/nranks slice of the tree-ordered particles. ---- */
/* pot_id / potspline_id: D*D ids of enum g2gpu_potlaw indexed [target][source]. */
int g2gpu_set_potential_laws(g2gpu_ctx *ctx, const int *pot_id, const int *potspline_id);
/* shortrange_fourier_pot[target][source][NTAB] (forcetree.c:34, filled at 3346), as double. */
int g2gpu_set_srpot_table(g2gpu_ctx *ctx, const double *table);
int g2gpu_potential(g2gpu_ctx *ctx, const g2gpu_walk_params *wp);
/* pot[n] = P[].Potential as left by the walk, in CURRENT particle order; with nranks > 1 only the entries of this rank's slice are written
 * by a call (the others keep what an earlier call left, zero initially).  *kernel_ms (optional) = CUDA-event time of the walk kernel. */
int g2gpu_download_potential(g2gpu_ctx *ctx, float *pot, double *kernel_ms);

/* ---- periodic PM long-range force: pmforce_periodic (pm_periodic.c:204-790), the caller of which is long_range_force
 *      (longrange.c:56-141; accel.c:36 runs it BEFORE gravity_tree).  Works on the particle records uploaded last, in UPLOAD
 *      order (no g2gpu_domain needed): CIC assignment per species, FFT, per-pair Green's function x Gaussian filter x CIC
 *      deconvolution, inverse FFT, 4-point differences, CIC interpolation; FP64 meshes.  The result also becomes the GravPM input
 *      of the next g2gpu_domain/g2gpu_walk (walk_params.use_gravpm), as if it had been passed to g2gpu_upload. ---- */
int g2gpu_pm_periodic(g2gpu_ctx *ctx, const g2gpu_pm_params *pp);
int g2gpu_download_gravpm(g2gpu_ctx *ctx, float *gravpm);	/* P[].GravPM, n x 3, upload order */
/* pmpotential_periodic (pm_periodic.c:798-1290; compute_potential calls it at potential.c:271, after the tree potential): the long-range
 * potential of the particle records uploaded last.  potential[n] (host, UPLOAD order) receives what the reference's routine ADDS to
 * P[].Potential: CIC assignment per species, FFT, -exp(-k^2 asmth^2) x GreensFxns[source][target](k) x G/(pi L) / sinc^4, inverse FFT, CIC
 * interpolation of the potential mesh.  The k = 0 mode is kept like the reference keeps it (:1033-1035) when it is finite (Yukawa-type
 * Green's functions); for 1/k^2 the reference leaves an infinite mode in (its reset at :1060-1063 only catches NaN) and returns infinite
 * potentials -- here that mode is set to zero.  GravPM and the state of the force path are untouched. */
int g2gpu_pm_potential_periodic(g2gpu_ctx *ctx, const g2gpu_pm_params *pp, float *potential);

/* acc[3n] = P[].GravAccel (after the G scaling), cost[n] = P[].GravCost, oldacc[n] = P[].OldAcc, in CURRENT
 * particle order; only entries of active particles of this rank's slice are written. */
int g2gpu_download_acc(g2gpu_ctx *ctx, float *acc, float *cost, float *oldacc);
int g2gpu_slice(g2gpu_ctx *ctx, int *lo, int *hi);	/* this rank's slice of the walk-ordered active list */
/* Whole step with host buffers: upload -> domain -> treebuild -> walk -> download (the e2e path). */
int g2gpu_gravity_tree(g2gpu_ctx *ctx, int npart, const float *pos, const float *mass, const int *type,
		       const float *oldacc, const int *active, const g2gpu_walk_params *wp, float *acc, float *cost,
		       float *oldacc_out, int *perm);

/* Run-time options: "rank", "nranks" (equal target slices, boundaries at multiples of 32),
 * "walk_exact" (1, default: targets that meet a decision within FP32 rounding of its threshold are walked again in the reference's
 * double arithmetic, so GravCost is the reference's exactly; 0: FP32 decisions only), "acc_double" (1: FP64 acceleration accumulators,
 * default; 0: FP32 accumulators, implies walk_exact 0), "walk_stats" (1: instrumented walk kernel, fills counters[2..4] of g2gpu_timings),
 * "compact" (1: the walk stores the results of this rank's slice in target order for g2gpu_download_slice instead of by particle index),
 * "direct_ewald" (1: g2gpu_direct adds the exact periodic lattice correction),
 * "accumulator" (1 = the reference built with -DNGRAVS_ACCUMULATOR; takes effect at the next g2gpu_treebuild). */
int g2gpu_set_option(g2gpu_ctx *ctx, const char *name, int value);

/* ---- several GPUs of one node behind the same entry points: the parallel driver inside gravity_tree() (gravtree.c:102-285: export,
 *      walk, import), as ONE host process with one context, stream and host thread per device (csrc/g2_group.cu).  Every device
 *      uploads its 1/N of the particle records, one ncclAllGather over NVLink replicates them, every device builds the same tree and
 *      walks one slice of the Peano-Hilbert-ordered active targets (slices of equal GravCost of the previous call, domain.c:859-862;
 *      boundaries at multiples of 32 targets, so results are bit-identical for any device count), and returns only its slice.
 *      ndev <= 0: all visible devices; devices == NULL: 0..ndev-1; cfg->device/rank/nranks are set per device.  With one device no
 *      NCCL is needed (it is resolved at run time).  Stage calls mirror the single-context ones; tables and options go to every
 *      device; anything not listed (tree mirror, potential, PM, direct sums) is served by g2gpu_group_ctx(grp, 0). ---- */
typedef struct g2gpu_group g2gpu_group;
int g2gpu_group_create(g2gpu_group **grp, const g2gpu_config *cfg, int ndev, const int *devices);
void g2gpu_group_destroy(g2gpu_group *grp);
int g2gpu_group_size(g2gpu_group *grp);
g2gpu_ctx *g2gpu_group_ctx(g2gpu_group *grp, int i);
int g2gpu_group_set_species(g2gpu_group *grp, const int type_to_grav[G2GPU_NTYPES], const double force_softening[G2GPU_NTYPES]);
int g2gpu_group_set_laws(g2gpu_group *grp, const int *accel_id, const int *spline_id, const double *params);
int g2gpu_group_set_srtable(g2gpu_group *grp, const double *table);
int g2gpu_group_set_lattice_tables(g2gpu_group *grp, int en, const double *fcorr);
/* the options of g2gpu_set_option on every device, plus "cost_weighted" (1, default: slices of equal GravCost; 0: equal counts) */
int g2gpu_group_set_option(g2gpu_group *grp, const char *name, int value);
/* sharded H2D (device i copies records [i*per, (i+1)*per), per = ceil(npart / N)) + ncclAllGather; arguments as g2gpu_upload[_aos] */
int g2gpu_group_upload(g2gpu_group *grp, int npart, const float *pos, const float *mass, const int *type, const float *oldacc, const int *active);
int g2gpu_group_upload_aos(g2gpu_group *grp, int npart, const void *P, size_t stride, int float_bytes, int off_pos, int off_mass, int off_type,
			   int off_oldacc, int off_vel, int off_gravpm, int off_ti_endstep, int ti_current);
/* device-resident inputs: every device already holds ITS shard (g2gpu_group_shard) inside its input buffer (g2gpu_input_buffers of
 * g2gpu_group_ctx(grp, i), record offset lo); only the all-gather runs */
int g2gpu_group_gather_resident(g2gpu_group *grp, int npart);
int g2gpu_group_shard(g2gpu_group *grp, int npart, int i, int *lo, int *cnt);
int g2gpu_group_domain(g2gpu_group *grp);
int g2gpu_group_treebuild(g2gpu_group *grp, int *numnodes);
int g2gpu_group_update_tree(g2gpu_group *grp, const float *len, const float *s);
int g2gpu_group_walk(g2gpu_group *grp, const g2gpu_walk_params *wp);
/* slice downloads + host scatter: SoA in CURRENT (device) order like g2gpu_download_acc, or into the reference's P[]: P[perm[p]] for device
 * index p (perm == NULL: P[] is in device order); float_bytes = sizeof(FLOAT) of GravAccel and OldAcc, GravCost is a float (allvars.h:572);
 * off_gravcost / off_oldacc may be -1; *cost_sum = sum of GravCost over the targets */
int g2gpu_group_download_acc(g2gpu_group *grp, float *acc, float *cost, float *oldacc);
int g2gpu_group_download_aos(g2gpu_group *grp, void *P, size_t stride, int float_bytes, int off_gravaccel, int off_gravcost, int off_oldacc,
			     const int *perm, double *cost_sum);
int g2gpu_group_get_order(g2gpu_group *grp, int *perm);
/* whole step with host buffers (the e2e path), arguments as g2gpu_gravity_tree; one host thread per device runs its whole pipeline */
int g2gpu_group_gravity_tree(g2gpu_group *grp, int npart, const float *pos, const float *mass, const int *type, const float *oldacc,
			     const int *active, const g2gpu_walk_params *wp, float *acc, float *cost, float *oldacc_out, int *perm);
/* one step from device-resident shards (see g2gpu_group_gather_resident): all-gather -> domain -> treebuild -> walk; results stay on the devices */
int g2gpu_group_step_resident(g2gpu_group *grp, int npart, const g2gpu_walk_params *wp);
int g2gpu_group_sync(g2gpu_group *grp);
int g2gpu_group_timings(g2gpu_group *grp, double ms[8], long long counters[8]);	/* times: max over devices; counters: sums */
int g2gpu_group_io_bytes(g2gpu_group *grp, long long out[3]);	/* H2D, D2H (all devices), bytes received per device by the all-gather */
/* 1 when the last g2gpu_group_gravity_tree found acc / cost / oldacc_out pinned and mapped on every device (cudaHostAlloc, cudaHostRegister) and let
 * the walk kernel store the results straight into them (no download, no host scatter; replaces the export / import of results of gravtree.c:102-285);
 * 0 when it staged and scattered them.  G2GPU_ZERO_COPY=0 forces the staged path. */
int g2gpu_group_zero_copy(g2gpu_group *grp);
/* Binds the results of the following g2gpu_group_walk calls to an array of structures in host memory -- the reference's P[] (allvars.h:548-590):
 * the walk kernel stores GravAccel[3] / GravCost / OldAcc of every active target straight into P[order[i]] (what gravtree.c:231-283 and 304-358
 * do with the imported results), and g2gpu_group_download_aos with the same arguments only waits for the devices.  The array is page-locked
 * here (cudaHostRegister, portable + mapped) unless it already is; if that fails, or with G2GPU_ZERO_COPY=0, nothing is bound and downloads are
 * staged.  nelem = allocated elements (All.MaxPart); P == NULL removes the binding (do so before the array is freed). */
int g2gpu_group_bind_results_aos(g2gpu_group *grp, void *P, size_t nelem, size_t stride, int float_bytes, int off_gravaccel, int off_gravcost, int off_oldacc);
int g2gpu_group_slices(g2gpu_group *grp, int *lo, int *hi, double *next_frac);	/* last walk's slices [N]; next walk's boundaries [N+1] */

/* ---- instrumentation ---- */
/* CUDA-event times (ms) of the last call of each stage: [0] domain [1] treebuild [2] walk
 * [3] walk kernel only [4] sort kernels of stage 1 [5] H2D [6] D2H [7] g2gpu_pm_periodic; counters[0] = kernel launches since
 * g2gpu_reset_counters, [1] = sum of GravCost of the last walk (this rank's targets), [2] = cell visits summed
 * over warps, [3] = species terms evaluated (one per particle interaction, <= D per node interaction), [4] = opening decisions
 * (node visits summed over the targets that were awake at the visit) -- [2..4] only with the option "walk_stats" --, [5] = targets
 * whose walk was repeated in FP64 because a decision was borderline in FP32 (option "walk_exact"), [6..7] reserved (0). */
int g2gpu_timings(g2gpu_ctx *ctx, double ms[8], long long counters[8]);
void g2gpu_reset_counters(g2gpu_ctx *ctx);
int g2gpu_get_counts(g2gpu_ctx *ctx, int out[4]);	/* particles, tree nodes (Numnodestree), active targets of the last walk, stage reached */
int g2gpu_io_bytes(g2gpu_ctx *ctx, long long out[2]);	/* host->device / device->host bytes of the last upload / download */
void *g2gpu_stream(g2gpu_ctx *ctx);	/* cudaStream_t all kernels are launched on */
int g2gpu_sync(g2gpu_ctx *ctx);

/* Stand-alone kernels exposed for parity tests. */
int g2gpu_peano_keys(g2gpu_ctx *ctx, int n, const int *xyz, int bits, long long *keys);	/* peano.c:356 */
int g2gpu_sort_pairs(g2gpu_ctx *ctx, int n, unsigned long long *keys, unsigned int *vals, int begin_bit, int end_bit);
/* AccelFxns[tgt][src](pm,m,r2,r,N)/r resp. AccelSplines[tgt][src](pm,m,h,r,N) for r>=h resp. r<h, in the
 * kernel's own arithmetic: out[i] = fac such that acc += d*fac (forcetree.c:1542-1544). */
int g2gpu_eval_pairs(g2gpu_ctx *ctx, int n, int tgt, int src, const float *pm, const float *m, const float *r,
		     const float *h, const int *npart_in_node, float *fac);

/* what the potential walk adds to pot for one pair without the TreePM table term: -PotentialFxns[tgt][src](pm, m, h, r, N) for r >= h,
 * +PotentialSplines[tgt][src](pm, m, h, r, N) for r < h (forcetree.c:2732-2734, 3115-3118), in the kernel's own arithmetic */
int g2gpu_eval_potentials(g2gpu_ctx *ctx, int n, int tgt, int src, const float *pm, const float *m, const float *r, const float *h,
			  const int *npart_in_node, float *out);

#ifdef __cplusplus
}
#endif
#endif
