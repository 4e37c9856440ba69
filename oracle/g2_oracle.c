/* oracle/g2_oracle.c — CPU RESTATEMENT of the reference's tree-gravity hot path.  TEST INFRASTRUCTURE ONLY:
 * nothing in the product (gadget-2.0.7-ngravs_b200/, include/) links, imports or executes this file; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may.
 *
 * Parity status: PINNED.  tests/test_oracle_port.py checks every stage of this file bit-for-bit (keys, order,
 * TopNodes, node records, Nextnode/Father, GravCost) and the accelerations to 1e-12 against the UNMODIFIED
 * reference compiled by oracle/ref/Makefile (oracle/_ref), and against the fixtures in tests/golden/ that were
 * generated from that reference build (tests/golden/make_golden.py).
 *
 * It is a plain sequential C program with the reference's compile-time switches (N_GRAVS, PERIODIC, PMGRID,
 * UNEQUALSOFTENINGS; Makefile.reference:49-135) turned into run-time fields, for FLOAT = float, one MPI rank
 * (NTask = 1: no pseudo-particles, forcetree.c:366-368) and -DNOTREERND.  Every function cites the reference
 * lines it restates.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <sys/time.h>

#include "../include/g2_ph_table.h"

#define MAXG 6
#define PH_BITS 18			/* BITS_PER_DIMENSION, allvars.h:34 */
#define MAXTOP 200000			/* MAXTOPNODES, allvars.h:29 */

typedef long long peanokey;

typedef struct
{
  int daughter, leaf;
  peanokey size, startkey;
  long long count;
} topnode;

typedef struct
{
  float len, center[3];
  int suns[8];				/* build phase (the reference overlays this with the fields below) */
  float s[3][MAXG], mass[MAXG];
  int bitflags, sibling, nextnode, father;
} node;

typedef struct
{
  float pos[3], mass, oldacc, gravpm[3];
  int type, id, active;
  peanokey key;
  float acc[3], cost;
  double accd[3];			/* walk result before rounding to FLOAT (diagnostics) */
} particle;

typedef struct g2o
{
  int D, periodic, shortrange, unequal, ntab;
  int maxpart, maxnodes;
  int t2g[6];
  double fsoft[6];			/* All.ForceSoftening */
  int accel_id[MAXG][MAXG], spline_id[MAXG][MAXG];
  double lawpar[MAXG][MAXG][4];
  double *srtab;			/* [D][D][ntab] */
  double *srpot;			/* shortrange_fourier_pot [D][D][ntab] */
  double *fcorr;			/* fcorrx/y/z after lattice_init: [3][D][D][(en+1)^3] (PERIODIC without PMGRID), or NULL */
  double *potcorr;		/* potcorr after lattice_init: [D][D][(en+1)^3] (PERIODIC without PMGRID), or NULL */
  int pot_en;
  int lattice_en;
  double lattice_cost;			/* sum of the lattice walk's return values (ewaldcount, gravtree.c:120) */
  int pot_id[MAXG][MAXG], potspline_id[MAXG][MAXG];	/* PotentialFxns / PotentialSplines, ids of include/g2gpu.h */
  int node_table_term;			/* -DNGRAVS_ACCUMULATOR build: node terms of the short-range potential carry the table term */
  float *pot;				/* P[].Potential after the potential walk */
  double boxsize, G, theta, errtol, asmth, rcut;
  int n;
  particle *P;
  double corner[3], center[3], len, fac;
  topnode *top;
  int ntop, ntopleaves;
  int *dni;				/* DomainNodeIndex */
  node *nodes;				/* nodes[k] is reference node MaxPart+k */
  int numnodes;
  int *nextnode, *father;		/* per particle */
  int last;
  double t_domain, t_build, t_walk;
} g2o;

static const unsigned char ph_table[G2_PH_NSTATES * 8] = G2_PH_TABLE_INIT;

static double wall(void)
{
  struct timeval tv;
  gettimeofday(&tv, NULL);
  return tv.tv_sec + 1e-6 * tv.tv_usec;
}

/* ---- peano_hilbert_key, peano.c:356-398, as the state machine of include/g2_ph_table.h ---- */
peanokey g2o_peano_key(int x, int y, int z, int bits)
{
  unsigned st = 0;
  peanokey key = 0;
  int l;
  for(l = bits - 1; l >= 0; l--)
    {
      unsigned o = (((x >> l) & 1) << 2) | (((y >> l) & 1) << 1) | ((z >> l) & 1);
      unsigned e = ph_table[st * 8 + o];
      key = (key << 3) | (e & 7);
      st = e >> 3;
    }
  return key;
}

/* ---- pair laws, ngravs.c:344-861 (registry ids of include/g2gpu.h) ---- */
static double plummer_law(double m, double h, double r)	/* ngravs.c:420-434 */
{
  double hinv = 1 / h, u = r * hinv;
  if(u < 0.5)
    return m * hinv * hinv * hinv * (10.666666666667 + u * u * (32.0 * u - 38.4));
  return m * hinv * hinv * hinv * (21.333333333333 - 48.0 * u + 38.4 * u * u - 10.666666666667 * u * u * u -
				   0.066666666667 / (u * u * u));
}

static double bam_accel(double rho, double eta, double r)	/* shared body of ngravs.c:495-529, 590-614, 646-670 */
{
  double reta = r * eta, reta2 = reta * reta, eta3 = eta * eta * eta;
  if(reta < 0.1)
    return rho * eta3 * (2.0 * r / 3.0 - 4.0 * reta2 * r / 5.0 + 6.0 * reta2 * reta2 * r / 7.0);
  return rho * eta3 * (atan(reta) / (reta2 * eta) - 1.0 / (reta * eta * (1 + reta2)));
}

static double bam_spline(double rho, double eta, double r)	/* ngravs.c:531-560, 562-588, 616-644 */
{
  double reta = r * eta, reta2 = reta * reta, eta3 = eta * eta * eta;
  if(reta < 0.1)
    return rho * eta3 * (2.0 / 3.0 - 4.0 * reta2 / 5.0 + 6.0 * reta2 * reta2 / 7.0);
  return rho * eta3 * (atan(reta) / (reta2 * reta) - 1.0 / (reta2 * (1 + reta2)));
}

/* AccelFxns[tg][sg](pm, m, r2, r, N) */
static double accel_fxn(const g2o * o, int tg, int sg, double pm, double m, double r2, double r, long N)
{
  const double *par = o->lawpar[tg][sg];
  switch (o->accel_id[tg][sg])
    {
    case 1: return m / r2;						/* newtonian      ngravs.c:351 */
    case 2: return -m / r2;						/* neg_newtonian  ngravs.c:359 */
    case 3: return m * exp(-r * par[0]) * (par[0] / r + 1.0 / r2);	/* yukawa         ngravs.c:856 */
    case 4: return m * exp(-r * par[0]) * (par[0] / r + 1.0 / r2) + m / r2;	/* coloyuk ngravs.c:826 */
    case 5: return bam_accel(2 * pm * m / M_PI, 4.0 * M_PI * par[1] / (pm + m / N), r);	/* bambam */
    case 6: return bam_accel(2 * pm * m / M_PI, 4.0 * M_PI * par[1] * N / m, r);	/* sourcebambaryon */
    case 7: return bam_accel(2 * pm * m / M_PI, 4.0 * M_PI * par[1] / pm, r);		/* sourcebaryonbam */
    default: return 0.0;						/* none           ngravs.c:344 */
    }
}

/* AccelSplines[tg][sg](pm, m, h, r, N) */
static double accel_spline(const g2o * o, int tg, int sg, double pm, double m, double h, double r, long N)
{
  const double *par = o->lawpar[tg][sg];
  switch (o->spline_id[tg][sg])
    {
    case 17: return plummer_law(m, h, r);
    case 18: return -plummer_law(m, h, r);
    case 19: return bam_spline(2 * pm * m / M_PI, 4.0 * M_PI * par[1] / (pm + m / N), r);
    case 20: return bam_spline(2 * pm * m / M_PI, 4.0 * M_PI * par[1] * N / m, r);
    case 21: return bam_spline(2 * pm * m / M_PI, 4.0 * M_PI * par[1] / pm, r);
    default: return 0.0;
    }
}

double g2o_accel(g2o * o, int tg, int sg, double pm, double m, double r2, double r, long N) { return accel_fxn(o, tg, sg, pm, m, r2, r, N); }
double g2o_spline(g2o * o, int tg, int sg, double pm, double m, double h, double r, long N) { return accel_spline(o, tg, sg, pm, m, h, r, N); }

/* ---- life cycle ---- */
g2o *g2o_create(int D, int periodic, int shortrange, int unequal, int ntab, int maxpart, int maxnodes)
{
  g2o *o = calloc(1, sizeof(g2o));
  int i, j;
  o->D = D; o->periodic = periodic; o->shortrange = shortrange; o->unequal = unequal; o->ntab = ntab;
  o->maxpart = maxpart; o->maxnodes = maxnodes;
  o->P = calloc(maxpart, sizeof(particle));
  o->top = calloc(MAXTOP, sizeof(topnode));
  o->dni = calloc(MAXTOP, sizeof(int));
  o->nodes = calloc((size_t) maxnodes + 1, sizeof(node));
  o->nextnode = calloc(maxpart, sizeof(int));
  o->father = calloc(maxpart, sizeof(int));
  for(i = 0; i < MAXG; i++)
    for(j = 0; j < MAXG; j++)
      {
	o->accel_id[i][j] = 1;
	o->spline_id[i][j] = 17;
      }
  o->G = 1; o->theta = 0.5; o->errtol = 0.005;
  return o;
}

void g2o_destroy(g2o * o)
{
  if(!o)
    return;
  free(o->P); free(o->top); free(o->dni); free(o->nodes); free(o->nextnode); free(o->father); free(o->srtab); free(o->srpot); free(o->pot); free(o->fcorr); free(o->potcorr); free(o);
}

void g2o_set_species(g2o * o, const int *t2g, const double *fsoft)
{
  int t;
  for(t = 0; t < 6; t++)
    {
      o->t2g[t] = t2g[t];
      o->fsoft[t] = fsoft[t];
    }
}

void g2o_set_laws(g2o * o, const int *accel_id, const int *spline_id, const double *par)
{
  int i, j, k;
  for(i = 0; i < o->D; i++)
    for(j = 0; j < o->D; j++)
      {
	o->accel_id[i][j] = accel_id[i * o->D + j];
	o->spline_id[i][j] = spline_id[i * o->D + j];
	for(k = 0; k < 4; k++)
	  o->lawpar[i][j][k] = par ? par[4 * (i * o->D + j) + k] : 0;
      }
}

void g2o_set_srtable(g2o * o, const double *tab)
{
  size_t n = (size_t) o->D * o->D * o->ntab;
  free(o->srtab);
  o->srtab = malloc(sizeof(double) * n);
  memcpy(o->srtab, tab, sizeof(double) * n);
}

void g2o_set_params(g2o * o, double boxsize, double G, double theta, double errtol, double asmth, double rcut)
{
  o->boxsize = boxsize; o->G = G; o->theta = theta; o->errtol = errtol; o->asmth = asmth; o->rcut = rcut;
}

int g2o_load(g2o * o, int n, const float *pos, const float *mass, const int *type, const float *oldacc, const int *active)
{
  int i, k;
  if(n > o->maxpart)
    return -1;
  o->n = n;
  for(i = 0; i < n; i++)
    {
      particle *p = &o->P[i];
      memset(p, 0, sizeof(*p));
      for(k = 0; k < 3; k++)
	p->pos[k] = pos[3 * i + k];
      p->mass = mass[i];
      p->type = type[i];
      p->id = i;
      p->oldacc = oldacc ? oldacc[i] : 0;
      p->active = active ? (active[i] != 0) : 1;
    }
  return 0;
}

/* ---- domain_findExtent, domain.c:882-924 ---- */
static void find_extent(g2o * o)
{
  double xmin[3], xmax[3], len = 0;
  int i, j;
  for(j = 0; j < 3; j++)
    {
      xmin[j] = 1e37;
      xmax[j] = -1e37;
    }
  for(i = 0; i < o->n; i++)
    for(j = 0; j < 3; j++)
      {
	if(xmin[j] > o->P[i].pos[j])
	  xmin[j] = o->P[i].pos[j];
	if(xmax[j] < o->P[i].pos[j])
	  xmax[j] = o->P[i].pos[j];
      }
  for(j = 0; j < 3; j++)
    if(xmax[j] - xmin[j] > len)
      len = xmax[j] - xmin[j];
  len *= 1.001;
  for(j = 0; j < 3; j++)
    {
      o->center[j] = 0.5 * (xmin[j] + xmax[j]);
      o->corner[j] = 0.5 * (xmin[j] + xmax[j]) - 0.5 * len;
    }
  o->len = len;
  o->fac = 1.0 / len * (((peanokey) 1) << PH_BITS);
}

static peanokey key_of(const g2o * o, const float *pos)	/* domain.c:940-943 == forcetree.c:144-146 */
{
  return g2o_peano_key((int) ((pos[0] - o->corner[0]) * o->fac), (int) ((pos[1] - o->corner[1]) * o->fac),
		       (int) ((pos[2] - o->corner[2]) * o->fac), PH_BITS);
}

static int cmp_key(const void *a, const void *b)
{
  peanokey x = *(const peanokey *) a, y = *(const peanokey *) b;
  return (x > y) - (x < y);
}

/* ---- domain_topsplit_local, domain.c:1019-1075 (NTask = 1: the global pass, 1086-1138, yields the same tree) ---- */
static void topsplit(g2o * o, const peanokey * sorted, int node, int pstart, peanokey startkey)
{
  int i, p, sub, pst[8];
  topnode *T = o->top;
  if(T[node].size < 8)
    return;
  T[node].daughter = o->ntop;
  for(i = 0; i < 8; i++)
    {
      if(o->ntop >= MAXTOP)
	{
	  fprintf(stderr, "g2o: out of top nodes\n");
	  abort();
	}
      sub = T[node].daughter + i;
      T[sub].size = T[node].size / 8;
      T[sub].count = 0;
      T[sub].daughter = -1;
      T[sub].leaf = -1;
      T[sub].startkey = startkey + i * T[sub].size;
      pst[i] = pstart;
      o->ntop++;
    }
  for(p = pstart; p < pstart + T[node].count; p++)
    {
      int bin = (int) ((sorted[p] - startkey) / (T[node].size / 8));
      sub = T[node].daughter + bin;
      if(T[sub].count == 0)
	pst[bin] = p;
      T[sub].count++;
    }
  for(i = 0; i < 8; i++)
    {
      sub = T[node].daughter + i;
      if(T[sub].count > o->n / 20.0)	/* TOPNODEFACTOR, domain.c:29,1071 */
	topsplit(o, sorted, sub, pst[i], T[sub].startkey);
    }
}

static void walk_toptree(g2o * o, int no)	/* domain_walktoptree, domain.c:802-816 */
{
  int i;
  if(o->top[no].daughter == -1)
    o->top[no].leaf = o->ntopleaves++;
  else
    for(i = 0; i < 8; i++)
      walk_toptree(o, o->top[no].daughter + i);
}

typedef struct { long long block; peanokey key; int index; } sortrec;
static int cmp_rec(const void *a, const void *b)
{
  const sortrec *x = a, *y = b;
  if(x->block != y->block)
    return (x->block > y->block) - (x->block < y->block);
  if(x->key != y->key)
    return (x->key > y->key) - (x->key < y->key);
  return (x->index > y->index) - (x->index < y->index);	/* the reference's qsort leaves ties unordered; we pick input order */
}

/* domain_Decomposition for one rank: extent, keys, top tree (domain.c:882-1138), then peano_hilbert_order
 * (peano.c:36-185): gas block first, then species-major, Peano-Hilbert within a block. */
void g2o_domain(g2o * o)
{
  double t0 = wall();
  int i, n = o->n;
  peanokey *sorted = malloc(sizeof(peanokey) * n);
  sortrec *rec = malloc(sizeof(sortrec) * n);
  particle *tmp = malloc(sizeof(particle) * n);
  find_extent(o);
  for(i = 0; i < n; i++)
    sorted[i] = o->P[i].key = key_of(o, o->P[i].pos);
  qsort(sorted, n, sizeof(peanokey), cmp_key);
  o->ntop = 1;
  o->top[0].daughter = -1;
  o->top[0].leaf = -1;
  o->top[0].size = ((peanokey) 1) << (3 * PH_BITS);
  o->top[0].startkey = 0;
  o->top[0].count = n;
  topsplit(o, sorted, 0, 0, 0);
  o->ntopleaves = 0;
  walk_toptree(o, 0);
  for(i = 0; i < n; i++)
    {
      rec[i].block = o->P[i].type == 0 ? 0 : 1 + o->t2g[o->P[i].type];
      rec[i].key = o->P[i].key;
      rec[i].index = i;
    }
  qsort(rec, n, sizeof(sortrec), cmp_rec);
  for(i = 0; i < n; i++)
    tmp[i] = o->P[rec[i].index];
  memcpy(o->P, tmp, sizeof(particle) * n);
  free(tmp);
  free(rec);
  free(sorted);
  o->t_domain = wall() - t0;
}

/* ---- force_create_empty_nodes, forcetree.c:292-336 ---- */
static void create_empty_nodes(g2o * o, int no, int topnode_, unsigned state, int *nextfree)
{
  int i, j, k, n;
  if(o->top[topnode_].daughter < 0)
    return;
  for(i = 0; i < 2; i++)
    for(j = 0; j < 2; j++)
      for(k = 0; k < 2; k++)
	{
	  unsigned e = ph_table[state * 8 + (i << 2 | j << 1 | k)];	/* last digit of the child's PH key = TopNodes daughter */
	  int sub = e & 7, count = i + 2 * j + 4 * k, nf = *nextfree;
	  node *nd = &o->nodes[nf], *pa = &o->nodes[no];
	  pa->suns[count] = o->maxpart + nf;
	  nd->len = 0.5 * pa->len;
	  nd->center[0] = pa->center[0] + (2 * i - 1) * 0.25 * pa->len;
	  nd->center[1] = pa->center[1] + (2 * j - 1) * 0.25 * pa->len;
	  nd->center[2] = pa->center[2] + (2 * k - 1) * 0.25 * pa->len;
	  for(n = 0; n < 8; n++)
	    nd->suns[n] = -1;
	  if(o->top[o->top[topnode_].daughter + sub].daughter == -1)
	    o->dni[o->top[o->top[topnode_].daughter + sub].leaf] = nf;
	  *nextfree = nf + 1;
	  if(*nextfree >= o->maxnodes)
	    {
	      fprintf(stderr, "g2o: maximum number of tree nodes reached\n");
	      abort();
	    }
	  create_empty_nodes(o, nf, o->top[topnode_].daughter + sub, e >> 3, nextfree);
	}
}

/* children are coded as in the reference but MaxPart-relative: v < 0 empty, v < maxpart particle, else node maxpart+k */
#define IS_NODE(o, v) ((v) >= (o)->maxpart)
#define NODE_OF(o, v) (&(o)->nodes[(v) - (o)->maxpart])

/* ---- force_update_node_recursive, forcetree.c:451-743 ---- */
static void set_next(g2o * o, int no)
{
  if(o->last >= 0)
    {
      if(IS_NODE(o, o->last))
	NODE_OF(o, o->last)->nextnode = no;
      else
	o->nextnode[o->last] = no;
    }
  o->last = no;
}

static void update_node(g2o * o, int no, int sib, int father)
{
  int j, jj, g, p, pp, nextsib, suns[8], D = o->D;
  double s[3][MAXG], mass[MAXG];
  int maxsofttype = 7, diffsoftflag = 0;
  node *nd;

  if(!IS_NODE(o, no))
    {
      set_next(o, no);
      o->father[no] = father;
      return;
    }
  nd = NODE_OF(o, no);
  for(j = 0; j < 8; j++)
    suns[j] = nd->suns[j];
  set_next(o, no);
  for(g = 0; g < D; g++)
    s[0][g] = s[1][g] = s[2][g] = mass[g] = 0;

  for(j = 0; j < 8; j++)
    {
      if((p = suns[j]) < 0)
	continue;
      for(jj = j + 1; jj < 8; jj++)
	if((pp = suns[jj]) >= 0)
	  break;
      nextsib = jj < 8 ? pp : sib;
      update_node(o, p, nextsib, no);
      if(IS_NODE(o, p))
	{
	  node *ch = NODE_OF(o, p);
	  int ct = (ch->bitflags >> 2) & 7;
	  for(g = 0; g < D; g++)
	    {
	      mass[g] += ch->mass[g];
	      s[0][g] += ch->mass[g] * ch->s[0][g];
	      s[1][g] += ch->mass[g] * ch->s[1][g];
	      s[2][g] += ch->mass[g] * ch->s[2][g];
	    }
	  if(o->unequal)
	    {			/* forcetree.c:570-597 */
	      diffsoftflag |= (ch->bitflags >> 5) & 1;
	      if(maxsofttype == 7)
		maxsofttype = ct;
	      else if(ct != 7)
		{
		  if(o->fsoft[ct] > o->fsoft[maxsofttype])
		    {
		      maxsofttype = ct;
		      diffsoftflag = 1;
		    }
		  else if(o->fsoft[ct] < o->fsoft[maxsofttype])
		    diffsoftflag = 1;
		}
	    }
	}
      else
	{
	  particle *pa = &o->P[p];
	  g = o->t2g[pa->type];
	  mass[g] += pa->mass;
	  s[0][g] += pa->mass * pa->pos[0];
	  s[1][g] += pa->mass * pa->pos[1];
	  s[2][g] += pa->mass * pa->pos[2];
	  if(o->unequal)
	    {			/* forcetree.c:628-646 */
	      if(maxsofttype == 7)
		maxsofttype = pa->type;
	      else if(o->fsoft[pa->type] > o->fsoft[maxsofttype])
		{
		  maxsofttype = pa->type;
		  diffsoftflag = 1;
		}
	      else if(o->fsoft[pa->type] < o->fsoft[maxsofttype])
		diffsoftflag = 1;
	    }
	}
    }
  for(g = 0; g < D; g++)
    {
      if(mass[g] > 0)
	{
	  s[0][g] /= mass[g];
	  s[1][g] /= mass[g];
	  s[2][g] /= mass[g];
	}
      else
	{
	  s[0][g] = nd->center[0];
	  s[1][g] = nd->center[1];
	  s[2][g] = nd->center[2];
	}
      nd->s[0][g] = s[0][g];
      nd->s[1][g] = s[1][g];
      nd->s[2][g] = s[2][g];
      nd->mass[g] = mass[g];
    }
  nd->bitflags = o->unequal ? 4 * maxsofttype + 32 * diffsoftflag : 0;
  nd->sibling = sib;
  nd->father = father;
}

/* ---- force_treebuild, forcetree.c:61-281 + force_flag_localnodes 954-996 ---- */
int g2o_treebuild(g2o * o)
{
  double t0 = wall();
  int i, j, nfree, numnodes, subnode = 0, parent = -1, th, nn, no;
  const int MP = o->maxpart;
  node *root = &o->nodes[0];

  root->len = o->len;
  for(j = 0; j < 3; j++)
    root->center[j] = o->center[j];
  for(j = 0; j < 8; j++)
    root->suns[j] = -1;
  nfree = 1;
  if(o->top[0].daughter == -1)
    o->dni[o->top[0].leaf] = 0;
  create_empty_nodes(o, 0, 0, 0, &nfree);
  numnodes = nfree;
  /* node indices below are reference-style (MaxPart-based); nfree counts MaxPart-relative */
  for(i = 0; i < o->n; i++)
    {
      const float *pos = o->P[i].pos;
      peanokey key = key_of(o, pos);
      no = 0;
      while(o->top[no].daughter >= 0)
	no = o->top[no].daughter + (int) ((key - o->top[no].startkey) / (o->top[no].size / 8));
      th = MP + o->dni[o->top[no].leaf];
      while(1)
	{
	  if(th >= MP)
	    {
	      node *t = NODE_OF(o, th);
	      subnode = (pos[0] > t->center[0]) + 2 * (pos[1] > t->center[1]) + 4 * (pos[2] > t->center[2]);
	      nn = t->suns[subnode];
	      if(nn >= 0)
		{
		  parent = th;
		  th = nn;
		}
	      else
		{
		  t->suns[subnode] = i;
		  break;
		}
	    }
	  else
	    {			/* a leaf holding particle th: make a node (forcetree.c:183-247) */
	      node *pa = NODE_OF(o, parent), *nw = &o->nodes[nfree];
	      double lenhalf = 0.25 * pa->len;
	      pa->suns[subnode] = MP + nfree;
	      nw->len = 0.5 * pa->len;
	      nw->center[0] = (subnode & 1) ? pa->center[0] + lenhalf : pa->center[0] - lenhalf;
	      nw->center[1] = (subnode & 2) ? pa->center[1] + lenhalf : pa->center[1] - lenhalf;
	      nw->center[2] = (subnode & 4) ? pa->center[2] + lenhalf : pa->center[2] - lenhalf;
	      for(j = 0; j < 8; j++)
		nw->suns[j] = -1;
	      subnode = (o->P[th].pos[0] > nw->center[0]) + 2 * (o->P[th].pos[1] > nw->center[1]) + 4 * (o->P[th].pos[2] > nw->center[2]);
	      nw->suns[subnode] = th;
	      th = MP + nfree;
	      numnodes++;
	      nfree++;
	      if(numnodes >= o->maxnodes)
		{
		  fprintf(stderr, "g2o: maximum number %d of tree-nodes reached\n", o->maxnodes);
		  abort();
		}
	    }
	}
    }
  o->last = -1;
  update_node(o, MP, -1, -1);
  if(IS_NODE(o, o->last))
    NODE_OF(o, o->last)->nextnode = -1;
  else
    o->nextnode[o->last] = -1;
  /* force_flag_localnodes: with one rank every top-level node gets bits 0 and 1 */
  for(i = 0; i < o->ntopleaves; i++)
    {
      no = MP + o->dni[i];
      while(no >= 0 && !(NODE_OF(o, no)->bitflags & 1))
	{
	  NODE_OF(o, no)->bitflags |= 3;
	  no = NODE_OF(o, no)->father;
	}
    }
  o->numnodes = numnodes;
  o->t_build = wall() - t0;
  return numnodes;
}

/* ---- force_treeevaluate (forcetree.c:1244-1610) and force_treeevaluate_shortrange (1623-2052) ---- */
#define NEAREST(x) (((x) > boxhalf) ? ((x) - boxsize) : (((x) < -boxhalf) ? ((x) + boxsize) : (x)))

static int tree_evaluate(g2o * o, int target)
{
  const int D = o->D, MP = o->maxpart, SR = o->shortrange, PER = o->periodic;
  const particle *tp = &o->P[target];
  double r2[MAXG], dx[MAXG], dy[MAXG], dz[MAXG], mass[MAXG], r, fac, h;
  double acc_x = 0, acc_y = 0, acc_z = 0;
  const double pos_x = tp->pos[0], pos_y = tp->pos[1], pos_z = tp->pos[2];
  const int ptype = tp->type, pg = o->t2g[ptype];
  const double pmass = tp->mass, aold = o->errtol * tp->oldacc;
  const double boxsize = o->boxsize, boxhalf = 0.5 * o->boxsize;
  const double rcut = o->rcut, rcut2 = rcut * rcut, asmth = o->asmth;
  const double asmthfac = SR ? 0.5 / asmth * (o->ntab / 3.0) : 0, utor2wpi = SR ? 1.0 / (M_PI * 4 * asmth * asmth) : 0;
  int ninteractions = 0, no = MP, sG, g;
  node *nop = 0;

  h = o->fsoft[ptype];
  while(no >= 0)
    {
      if(no < MP)
	{
	  const particle *sp = &o->P[no];
	  sG = o->t2g[sp->type];
	  mass[sG] = sp->mass;
	  dx[sG] = sp->pos[0] - pos_x;
	  dy[sG] = sp->pos[1] - pos_y;
	  dz[sG] = sp->pos[2] - pos_z;
	  if(PER)
	    {
	      dx[sG] = NEAREST(dx[sG]);
	      dy[sG] = NEAREST(dy[sG]);
	      dz[sG] = NEAREST(dz[sG]);
	    }
	  r2[sG] = dx[sG] * dx[sG] + dy[sG] * dy[sG] + dz[sG] * dz[sG];
	  if(o->unequal)
	    {
	      h = o->fsoft[ptype];
	      if(h < o->fsoft[sp->type])
		h = o->fsoft[sp->type];
	    }
	  no = o->nextnode[no];
	}
      else
	{
	  double r2min = INFINITY, r2max = -INFINITY, summass = 0;
	  nop = NODE_OF(o, no);
	  for(g = 0; g < D; g++)
	    {
	      mass[g] = nop->mass[g];
	      summass += nop->mass[g];
	      dx[g] = nop->s[0][g] - pos_x;
	      dy[g] = nop->s[1][g] - pos_y;
	      dz[g] = nop->s[2][g] - pos_z;
	      if(PER)
		{
		  dx[g] = NEAREST(dx[g]);
		  dy[g] = NEAREST(dy[g]);
		  dz[g] = NEAREST(dz[g]);
		}
	      r2[g] = dx[g] * dx[g] + dy[g] * dy[g] + dz[g] * dz[g];
	      if(r2[g] < r2min)
		r2min = r2[g];
	      if(r2[g] > r2max)
		r2max = r2[g];
	    }
	  sG = -1;
	  if(SR && r2min > rcut2)
	    {			/* forcetree.c:1828-1862 */
	      double eff_dist = rcut + 0.5 * nop->len, dist;
	      int k, cull = 0;
	      const double pp[3] = { pos_x, pos_y, pos_z };
	      for(k = 0; k < 3 && !cull; k++)
		{
		  dist = nop->center[k] - pp[k];
		  if(PER)
		    dist = NEAREST(dist);
		  if(dist < -eff_dist || dist > eff_dist)
		    cull = 1;
		}
	      if(cull)
		{
		  no = nop->sibling;
		  continue;
		}
	    }
	  if(o->theta)
	    {
	      if(nop->len * nop->len > r2min * o->theta * o->theta)
		{
		  no = nop->nextnode;
		  continue;
		}
	    }
	  else
	    {
	      if(summass * nop->len * nop->len > r2min * r2min * aold)
		{
		  no = nop->nextnode;
		  continue;
		}
	      if(fabs(nop->center[0] - pos_x) < 0.60 * nop->len && fabs(nop->center[1] - pos_y) < 0.60 * nop->len
		 && fabs(nop->center[2] - pos_z) < 0.60 * nop->len)
		{
		  no = nop->nextnode;
		  continue;
		}
	    }
	  if(o->unequal)
	    {			/* forcetree.c:1475-1501 */
	      int maxsofttype = (nop->bitflags >> 2) & 7;
	      h = o->fsoft[ptype];
	      if(maxsofttype == 7)
		{
		  no = nop->nextnode;
		  continue;
		}
	      if(h < o->fsoft[maxsofttype])
		{
		  h = o->fsoft[maxsofttype];
		  if(r2max < h * h && ((nop->bitflags >> 5) & 1))
		    {
		      no = nop->nextnode;
		      continue;
		    }
		}
	    }
	  no = nop->sibling;
	}

      if(!SR)
	{			/* forcetree.c:1538-1585 */
	  for(g = (sG > -1 ? sG : 0); g < (sG > -1 ? sG + 1 : D); g++)
	    {
	      if(sG < 0 && mass[g] == 0.0)
		continue;
	      r = sqrt(r2[g]);
	      if(r >= h)
		fac = accel_fxn(o, pg, g, pmass, mass[g], r2[g], r, 1) / r;
	      else
		fac = accel_spline(o, pg, g, pmass, mass[g], h, r, 1);
	      acc_x += dx[g] * fac;
	      acc_y += dy[g] * fac;
	      acc_z += dz[g] * fac;
	    }
	  ninteractions++;
	}
      else
	{			/* forcetree.c:1955-2032 */
	  int nintflag = 0;
	  for(g = (sG > -1 ? sG : 0); g < (sG > -1 ? sG + 1 : D); g++)
	    {
	      int tabindex;
	      if(sG < 0 && mass[g] == 0.0)
		continue;
	      r = sqrt(r2[g]);
	      tabindex = (int) (asmthfac * r);
	      if(tabindex < o->ntab)
		{
		  if(r >= h)
		    {
		      fac = accel_fxn(o, pg, g, pmass, mass[g], r2[g], r, 1);
		      fac -= mass[g] * utor2wpi * o->srtab[((size_t) pg * D + g) * o->ntab + tabindex];
		      fac /= r;
		    }
		  else
		    fac = accel_spline(o, pg, g, pmass, mass[g], h, r, 1);
		  acc_x += dx[g] * fac;
		  acc_y += dy[g] * fac;
		  acc_z += dz[g] * fac;
		  nintflag = 1;
		}
	    }
	  if(nintflag)
	    ninteractions++;
	}
    }
  o->P[target].accd[0] = acc_x;
  o->P[target].accd[1] = acc_y;
  o->P[target].accd[2] = acc_z;
  o->P[target].acc[0] = acc_x;
  o->P[target].acc[1] = acc_y;
  o->P[target].acc[2] = acc_z;
  o->P[target].cost = ninteractions;
  return ninteractions;
}

/* ---- potential walks ----
 * PotentialFxns[tg][sg](pm, m, h, r, N) / PotentialSplines[tg][sg](pm, m, h, r, N): ngravs.c:368, 375, 459-490 */
static double plummer_pot_law(double m, double h, double r)	/* ngravs.c:459-471 */
{
  double hinv = 1 / h;
  r *= hinv;
  if(r < 0.5)
    return m * hinv * (-2.8 + r * r * (5.333333333333 + r * r * (6.4 * r - 9.6)));
  return m * hinv * (-3.2 + 0.066666666667 / r + r * r * (10.666666666667 + r * (-16.0 + r * (9.6 - 2.133333333333 * r))));
}

static double pot_fxn(const g2o * o, int tg, int sg, double m, double r)
{
  switch (o->pot_id[tg][sg])
    {
    case 33: return m / r;
    case 34: return -m / r;
    default: return 0.0;
    }
}

static double pot_spline(const g2o * o, int tg, int sg, double m, double h, double r)
{
  switch (o->potspline_id[tg][sg])
    {
    case 49: return plummer_pot_law(m, h, r);
    case 50: return -plummer_pot_law(m, h, r);
    default: return 0.0;
    }
}

double g2o_potfxn(g2o * o, int tg, int sg, double m, double h, double r) { (void) h; return pot_fxn(o, tg, sg, m, r); }
double g2o_potspline(g2o * o, int tg, int sg, double m, double h, double r) { return pot_spline(o, tg, sg, m, h, r); }

void g2o_set_potlaws(g2o * o, const int *pot_id, const int *potspline_id, int node_table_term)
{
  int i, j;
  for(i = 0; i < o->D; i++)
    for(j = 0; j < o->D; j++)
      {
	o->pot_id[i][j] = pot_id[i * o->D + j];
	o->potspline_id[i][j] = potspline_id[i * o->D + j];
      }
  o->node_table_term = node_table_term;
}

void g2o_set_srpot(g2o * o, const double *tab)
{
  size_t n = (size_t) o->D * o->D * o->ntab;
  free(o->srpot);
  o->srpot = malloc(sizeof(double) * n);
  memcpy(o->srpot, tab, sizeof(double) * n);
}

/* force_treeevaluate_potential_shortrange (forcetree.c:2789-3163) when o->shortrange, else force_treeevaluate_potential
 * (forcetree.c:2467-2776; the latter does not compile in the reference -- this branch restates what its text says, minus the
 * lattice correction of PERIODIC builds -- so it is NOT pinned against a reference run). */
static double lattice_pot_corr_port(const g2o * o, double dx, double dy, double dz, int target, int source);

static void tree_potential(g2o * o, int target)
{
  const int D = o->D, MP = o->maxpart, SR = o->shortrange, PER = o->periodic;
  const particle *tp = &o->P[target];
  double r2[MAXG], mass[MAXG], r, h, d[3], dsp[MAXG][3];
  double pot = 0;
  const double pos[3] = { tp->pos[0], tp->pos[1], tp->pos[2] };
  const int ptype = tp->type, pg = o->t2g[ptype];
  const double aold = o->errtol * tp->oldacc;
  const double boxsize = o->boxsize, boxhalf = 0.5 * o->boxsize;
  const double rcut = o->rcut, asmth = o->asmth;
  const double asmthfac = SR ? 0.5 / asmth * (o->ntab / 3.0) : 0, utorwpi = SR ? 1.0 / (2 * M_PI * asmth) : 0;
  int no = MP, sG, g, k;
  node *nop = 0;

  h = o->fsoft[ptype];
  while(no >= 0)
    {
      if(no < MP)
	{
	  const particle *sp = &o->P[no];
	  sG = o->t2g[sp->type];
	  mass[sG] = sp->mass;
	  for(k = 0; k < 3; k++)
	    {
	      d[k] = sp->pos[k] - pos[k];
	      if(PER)
		d[k] = NEAREST(d[k]);
	    }
	  r2[sG] = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
	  dsp[sG][0] = d[0]; dsp[sG][1] = d[1]; dsp[sG][2] = d[2];
	  if(o->unequal)
	    {
	      h = o->fsoft[ptype];
	      if(h < o->fsoft[sp->type])
		h = o->fsoft[sp->type];
	    }
	  no = o->nextnode[no];
	}
      else
	{
	  double r2min = INFINITY, r2max = -INFINITY, summass = 0;
	  nop = NODE_OF(o, no);
	  for(g = 0; g < D; g++)
	    {
	      mass[g] = nop->mass[g];
	      summass += nop->mass[g];
	      for(k = 0; k < 3; k++)
		{
		  d[k] = nop->s[k][g] - pos[k];
		  if(PER)
		    d[k] = NEAREST(d[k]);
		}
	      r2[g] = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
	      dsp[g][0] = d[0]; dsp[g][1] = d[1]; dsp[g][2] = d[2];
	      if(r2[g] < r2min)
		r2min = r2[g];
	      if(r2[g] > r2max)
		r2max = r2[g];
	    }
	  sG = -1;
	  if(SR)
	    {			/* forcetree.c:2988-3015: the box test alone */
	      double eff_dist = rcut + 0.5 * nop->len, dist;
	      int cull = 0;
	      for(k = 0; k < 3 && !cull; k++)
		{
		  dist = nop->center[k] - pos[k];
		  if(PER)
		    dist = NEAREST(dist);
		  if(dist < -eff_dist || dist > eff_dist)
		    cull = 1;
		}
	      if(cull)
		{
		  no = nop->sibling;
		  continue;
		}
	    }
	  if(o->theta)
	    {
	      if(nop->len * nop->len > r2min * o->theta * o->theta)
		{
		  no = nop->nextnode;
		  continue;
		}
	    }
	  else
	    {
	      if(summass * nop->len * nop->len > r2min * r2min * aold)
		{
		  no = nop->nextnode;
		  continue;
		}
	      if(fabs(nop->center[0] - pos[0]) < 0.60 * nop->len && fabs(nop->center[1] - pos[1]) < 0.60 * nop->len
		 && fabs(nop->center[2] - pos[2]) < 0.60 * nop->len)
		{
		  no = nop->nextnode;
		  continue;
		}
	    }
	  if(o->unequal)
	    {			/* forcetree.c:3049-3077 */
	      int maxsofttype = (nop->bitflags >> 2) & 7;
	      h = o->fsoft[ptype];
	      if(maxsofttype == 7)
		{
		  no = nop->nextnode;
		  continue;
		}
	      if(h < o->fsoft[maxsofttype])
		{
		  h = o->fsoft[maxsofttype];
		  if(r2max < h * h && ((nop->bitflags >> 5) & 1))
		    {
		      no = nop->nextnode;
		      continue;
		    }
		}
	    }
	  no = nop->sibling;
	}

      for(g = (sG > -1 ? sG : 0); g < (sG > -1 ? sG + 1 : D); g++)
	{			/* forcetree.c:3104-3155 (2724-2768 without PM) */
	  int tabindex = 0;
	  if(sG < 0 && mass[g] == 0.0)
	    continue;
	  r = sqrt(r2[g]);
	  if(SR)
	    {
	      tabindex = (int) (r * asmthfac);
	      if(tabindex >= o->ntab)
		continue;
	    }
	  if(r >= h)
	    {
	      double p = pot_fxn(o, pg, g, mass[g], r);
	      if(SR && (sG > -1 || o->node_table_term))	/* forcetree.c:3115-3116 (no mass factor) vs 3139 */
		p -= utorwpi * o->srpot[((size_t) pg * D + g) * o->ntab + tabindex];
	      pot -= p;
	    }
	  else
	    pot += pot_spline(o, pg, g, mass[g], h, r);
	  if(PER && !SR)	/* forcetree.c:2736-2738, 2765-2767 */
	    pot += mass[g] * lattice_pot_corr_port(o, dsp[g][0], dsp[g][1], dsp[g][2], pg, g);
	}
    }
  o->pot[target] = pot;		/* FLOAT store, forcetree.c:3158 */
}

struct pslice { g2o *o; int lo, hi; };
static void *pot_slice(void *arg)
{
  struct pslice *s = arg;
  int i;
  for(i = s->lo; i < s->hi; i++)
    tree_potential(s->o, i);
  return NULL;
}

/* the per-particle loop of compute_potential (potential.c:86-97) over ALL particles; out[i] = P[i].Potential after the walk */
int g2o_potential(g2o * o, int nthreads, float *out)
{
  pthread_t *th;
  struct pslice *sl;
  int t;
  if(o->shortrange && !o->srpot)
    return -1;
  if(!o->shortrange && o->periodic && !o->potcorr)
    return -2;			/* a periodic box without PM needs the potcorr tables (g2o_set_lattice_pot_tables) */
  if(nthreads < 1)
    nthreads = 1;
  free(o->pot);
  o->pot = malloc(sizeof(float) * (o->n > 0 ? o->n : 1));
  th = malloc(sizeof(pthread_t) * nthreads);
  sl = malloc(sizeof(struct pslice) * nthreads);
  for(t = 0; t < nthreads; t++)
    {
      sl[t].o = o;
      sl[t].lo = (int) ((long long) o->n * t / nthreads);
      sl[t].hi = (int) ((long long) o->n * (t + 1) / nthreads);
      pthread_create(&th[t], NULL, pot_slice, &sl[t]);
    }
  for(t = 0; t < nthreads; t++)
    pthread_join(th[t], NULL);
  free(sl);
  free(th);
  memcpy(out, o->pot, sizeof(float) * o->n);
  return 0;
}

/* the same walk for a list of targets only (full-size spot checks) */
int g2o_potential_targets(g2o * o, int ntargets, const int *targets, float *out)
{
  int t;
  if(o->shortrange && !o->srpot)
    return -1;
  if(!o->shortrange && o->periodic && !o->potcorr)
    return -2;
  free(o->pot);
  o->pot = malloc(sizeof(float) * (o->n > 0 ? o->n : 1));
  for(t = 0; t < ntargets; t++)
    {
      if(targets[t] < 0 || targets[t] >= o->n)
	return -3;
      tree_potential(o, targets[t]);
      out[t] = o->pot[targets[t]];
    }
  return 0;
}

/* ---- lattice-sum correction walk: force_treeevaluate_lattice_correction, forcetree.c:2077-2455 ---- */
void g2o_set_lattice_tables(g2o * o, int en, const double *fcorr)
{
  size_t n = (size_t) 3 * o->D * o->D * (en + 1) * (en + 1) * (en + 1);
  free(o->fcorr);
  o->fcorr = NULL;
  if(!fcorr)
    return;
  o->fcorr = malloc(sizeof(double) * n);
  memcpy(o->fcorr, fcorr, sizeof(double) * n);
  o->lattice_en = en;
}

/* ewald_force (ngravs.c:1170-1236) on the grid of lattice_init (forcetree.c:3700-3710): out[c * (en+1)^3 + (i*(en+1)+j)*(en+1)+k],
 * x = 0.5 (i,j,k)/en, dimensionless (lattice_init divides by BoxSize^2 afterwards, forcetree.c:3757-3761).  Same sums in the same order. */
struct ewslice { int en, lo, hi; double *out; double zero; };
static void *ewald_slice(void *arg)
{
  struct ewslice *s = arg;
  const int en = s->en, n1 = en + 1;
  const size_t n3 = (size_t) n1 * n1 * n1;
  const double alpha = 2.0;
  int q, a, n[3], h[3];
  for(q = s->lo; q < s->hi; q++)
    {
      const int i = q / (n1 * n1), j = (q / n1) % n1, k = q % n1;
      double x[3] = { 0.5 * ((double) i) / en, 0.5 * ((double) j) / en, 0.5 * ((double) k) / en };
      double force[3] = { 0, 0, 0 }, r2, r, val, hdotx, dx[3];
      if(q != 0)
	{
	  r2 = x[0] * x[0] + x[1] * x[1] + x[2] * x[2];
	  for(a = 0; a < 3; a++)
	    force[a] += x[a] / (r2 * sqrt(r2));
	  for(n[0] = -4; n[0] <= 4; n[0]++)
	    for(n[1] = -4; n[1] <= 4; n[1]++)
	      for(n[2] = -4; n[2] <= 4; n[2]++)
		{
		  for(a = 0; a < 3; a++)
		    dx[a] = x[a] - n[a];
		  r = sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]);
		  val = erfc(alpha * r) + 2 * alpha * r / sqrt(M_PI) * exp(-alpha * alpha * r * r);
		  for(a = 0; a < 3; a++)
		    force[a] -= dx[a] / (r * r * r) * val;
		}
	  for(h[0] = -4; h[0] <= 4; h[0]++)
	    for(h[1] = -4; h[1] <= 4; h[1]++)
	      for(h[2] = -4; h[2] <= 4; h[2]++)
		{
		  int h2 = h[0] * h[0] + h[1] * h[1] + h[2] * h[2];
		  hdotx = x[0] * h[0] + x[1] * h[1] + x[2] * h[2];
		  if(h2 > 0)
		    {
		      val = 2.0 / ((double) h2) * exp(-M_PI * M_PI * h2 / (alpha * alpha)) * sin(2 * M_PI * hdotx);
		      for(a = 0; a < 3; a++)
			force[a] -= h[a] * val;
		    }
		}
	}
      for(a = 0; a < 3; a++)
	s->out[a * n3 + q] = force[a];
    }
  return NULL;
}

int g2o_make_ewald_table(int en, int nthreads, double *out)
{
  const int n3 = (en + 1) * (en + 1) * (en + 1);
  pthread_t *th;
  struct ewslice *sl;
  int t;
  if(nthreads < 1)
    nthreads = 1;
  th = malloc(sizeof(pthread_t) * nthreads);
  sl = malloc(sizeof(struct ewslice) * nthreads);
  for(t = 0; t < nthreads; t++)
    {
      sl[t].en = en;
      sl[t].out = out;
      sl[t].lo = (int) ((long long) n3 * t / nthreads);
      sl[t].hi = (int) ((long long) n3 * (t + 1) / nthreads);
      pthread_create(&th[t], NULL, ewald_slice, &sl[t]);
    }
  for(t = 0; t < nthreads; t++)
    pthread_join(th[t], NULL);
  free(sl);
  free(th);
  return 0;
}

/* ---- the potential of a periodic box without PM: lattice_pot_corr (forcetree.c:3895-3941) over potcorr[tgt][src] of lattice_init
 *      (forcetree.c:3697-3702, 3759), and ewald_psi (ngravs.c:761-816), the function lattice_init tabulates for the stock wiring ---- */
void g2o_set_lattice_pot_tables(g2o * o, int en, const double *potcorr)
{
  const size_t n = (size_t) o->D * o->D * (en + 1) * (en + 1) * (en + 1);
  free(o->potcorr);
  o->potcorr = NULL;
  if(!potcorr)
    return;
  o->potcorr = malloc(sizeof(double) * n);
  memcpy(o->potcorr, potcorr, sizeof(double) * n);
  o->pot_en = en;
}

static double lattice_pot_corr_port(const g2o * o, double dx, double dy, double dz, int target, int source)
{
  const int en = o->pot_en, n1 = en + 1;
  const double fac_intp = 2 * en / o->boxsize;	/* forcetree.c:3750 */
  const double *t = o->potcorr + ((size_t) target * o->D + source) * n1 * n1 * n1;
  int i, j, k;
  double u, v, w, f1, f2, f3, f4, f5, f6, f7, f8;
  if(dx < 0) dx = -dx;
  if(dy < 0) dy = -dy;
  if(dz < 0) dz = -dz;
  u = dx * fac_intp; i = (int) u; if(i >= en) i = en - 1; u -= i;
  v = dy * fac_intp; j = (int) v; if(j >= en) j = en - 1; v -= j;
  w = dz * fac_intp; k = (int) w; if(k >= en) k = en - 1; w -= k;
  f1 = (1 - u) * (1 - v) * (1 - w);
  f2 = (1 - u) * (1 - v) * (w);
  f3 = (1 - u) * (v) * (1 - w);
  f4 = (1 - u) * (v) * (w);
  f5 = (u) * (1 - v) * (1 - w);
  f6 = (u) * (1 - v) * (w);
  f7 = (u) * (v) * (1 - w);
  f8 = (u) * (v) * (w);
#define PC(a, b, c) t[((size_t) (a) * n1 + (b)) * n1 + (c)]
  return PC(i, j, k) * f1 + PC(i, j, k + 1) * f2 + PC(i, j + 1, k) * f3 + PC(i, j + 1, k + 1) * f4 +
    PC(i + 1, j, k) * f5 + PC(i + 1, j, k + 1) * f6 + PC(i + 1, j + 1, k) * f7 + PC(i + 1, j + 1, k + 1) * f8;
#undef PC
}

double g2o_lattice_pot_corr(g2o * o, double dx, double dy, double dz, int target, int source)
{
  return o->potcorr ? lattice_pot_corr_port(o, dx, dy, dz, target, source) : 0.0 / 0.0;
}

static double ewald_psi_port(const double x[3])
{
  const double alpha = 2.0;
  double r, sum1 = 0, sum2 = 0, hdotx, dx[3];
  int i, n[3], h[3], h2;
  for(n[0] = -4; n[0] <= 4; n[0]++)
    for(n[1] = -4; n[1] <= 4; n[1]++)
      for(n[2] = -4; n[2] <= 4; n[2]++)
	{
	  for(i = 0; i < 3; i++)
	    dx[i] = x[i] - n[i];
	  r = sqrt(dx[0] * dx[0] + dx[1] * dx[1] + dx[2] * dx[2]);
	  sum1 += erfc(alpha * r) / r;
	}
  for(h[0] = -4; h[0] <= 4; h[0]++)
    for(h[1] = -4; h[1] <= 4; h[1]++)
      for(h[2] = -4; h[2] <= 4; h[2]++)
	{
	  hdotx = x[0] * h[0] + x[1] * h[1] + x[2] * h[2];
	  h2 = h[0] * h[0] + h[1] * h[1] + h[2] * h[2];
	  if(h2 > 0)
	    sum2 += 1 / (M_PI * h2) * exp(-M_PI * M_PI * h2 / (alpha * alpha)) * cos(2 * M_PI * hdotx);
	}
  r = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
  return M_PI / (alpha * alpha) - sum1 - sum2 + 1 / r;
}

static void *ewald_pot_slice(void *arg)
{
  struct ewslice *s = arg;
  const int en = s->en, n1 = en + 1;
  int q;
  for(q = s->lo; q < s->hi; q++)
    {
      const int i = q / (n1 * n1), j = (q / n1) % n1, k = q % n1;
      double x[3] = { 0.5 * ((double) i) / en, 0.5 * ((double) j) / en, 0.5 * ((double) k) / en };
      s->out[q] = q == 0 ? s->zero : ewald_psi_port(x);	/* LatticeZero at the origin (forcetree.c:3699-3702) */
    }
  return NULL;
}

/* potcorr of the stock wiring BEFORE the division by BoxSize (forcetree.c:3759): out[(en+1)^3]; zero = LatticeZero[l][m] (a FLOAT:
 * 2.8372975 rounded to float in a FLOAT = float build, ngravs.c:133) */
int g2o_make_ewald_pot_table(int en, int nthreads, double zero, double *out)
{
  const int n3 = (en + 1) * (en + 1) * (en + 1);
  pthread_t *th;
  struct ewslice *sl;
  int t;
  if(nthreads < 1)
    nthreads = 1;
  th = malloc(sizeof(pthread_t) * nthreads);
  sl = malloc(sizeof(struct ewslice) * nthreads);
  for(t = 0; t < nthreads; t++)
    {
      sl[t].en = en;
      sl[t].out = out;
      sl[t].zero = zero;
      sl[t].lo = (int) ((long long) n3 * t / nthreads);
      sl[t].hi = (int) ((long long) n3 * (t + 1) / nthreads);
      pthread_create(&th[t], NULL, ewald_pot_slice, &sl[t]);
    }
  for(t = 0; t < nthreads; t++)
    pthread_join(th[t], NULL);
  free(sl);
  free(th);
  return 0;
}

static int lattice_correction(g2o * o, int target, double pos_x, double pos_y, double pos_z, double aold)
{
  const int D = o->D, MP = o->maxpart, EN = o->lattice_en, n1 = EN + 1;
  const size_t n3 = (size_t) n1 * n1 * n1;
  const double boxsize = o->boxsize, boxhalf = 0.5 * o->boxsize, fac_intp = 2 * EN / o->boxsize;	/* forcetree.c:3749 */
  const int pg = o->t2g[o->P[target].type];
  double dx[MAXG], dy[MAXG], dz[MAXG], mass[MAXG], r2[MAXG];
  double acc[3] = { 0, 0, 0 };
  int no = MP, cost = 0, sG, g, n;
  node *nop = 0;

  while(no >= 0)
    {
      if(no < MP)
	{
	  const particle *sp = &o->P[no];
	  sG = o->t2g[sp->type];
	  mass[sG] = sp->mass;
	  dx[sG] = NEAREST(sp->pos[0] - pos_x);
	  dy[sG] = NEAREST(sp->pos[1] - pos_y);
	  dz[sG] = NEAREST(sp->pos[2] - pos_z);
	  no = o->nextnode[no];
	}
      else
	{
	  double r2min = INFINITY, summass = 0, u;
	  int openflag = 0, k;
	  const double pp[3] = { pos_x, pos_y, pos_z };
	  nop = NODE_OF(o, no);
	  for(g = 0; g < D; g++)
	    {
	      mass[g] = nop->mass[g];
	      summass += nop->mass[g];
	      dx[g] = NEAREST(nop->s[0][g] - pos_x);
	      dy[g] = NEAREST(nop->s[1][g] - pos_y);
	      dz[g] = NEAREST(nop->s[2][g] - pos_z);
	      r2[g] = dx[g] * dx[g] + dy[g] * dy[g] + dz[g] * dz[g];
	      if(r2[g] < r2min)
		r2min = r2[g];
	    }
	  sG = -1;
	  if(o->theta)
	    {
	      if(nop->len * nop->len > r2min * o->theta * o->theta)
		openflag = 1;
	    }
	  else
	    {
	      if(summass * nop->len * nop->len > r2min * r2min * aold)
		openflag = 1;
	      else if(fabs(nop->center[0] - pos_x) < 0.60 * nop->len && fabs(nop->center[1] - pos_y) < 0.60 * nop->len
		      && fabs(nop->center[2] - pos_z) < 0.60 * nop->len)
		openflag = 1;
	    }
	  if(openflag)
	    {			/* forcetree.c:2211-2256 */
	      int must = 0;
	      for(k = 0; k < 3 && !must; k++)
		{
		  u = nop->center[k] - pp[k];
		  if(u > boxhalf)
		    u -= boxsize;
		  if(u < -boxhalf)
		    u += boxsize;
		  if(fabs(u) > 0.5 * (boxsize - nop->len))
		    must = 1;
		}
	      if(must || nop->len > 0.20 * boxsize)
		{
		  no = nop->nextnode;
		  continue;
		}
	    }
	  no = nop->sibling;
	}

      for(n = (sG > -1 ? sG : 0); n < (sG > -1 ? sG + 1 : D); n++)
	{			/* forcetree.c:2262-2428 */
	  double sg[3], d[3] = { dx[n], dy[n], dz[n] }, f[3], w[8];
	  int ix[3], c, k;
	  const double *tab;
	  if(sG < 0 && mass[n] == 0.0)
	    continue;
	  for(k = 0; k < 3; k++)
	    {
	      if(d[k] < 0)
		{
		  d[k] = -d[k];
		  sg[k] = +1;
		}
	      else
		sg[k] = -1;
	      f[k] = d[k] * fac_intp;
	      ix[k] = (int) f[k];
	      if(ix[k] >= EN)
		ix[k] = EN - 1;
	      f[k] -= ix[k];
	    }
	  w[0] = (1 - f[0]) * (1 - f[1]) * (1 - f[2]);
	  w[1] = (1 - f[0]) * (1 - f[1]) * (f[2]);
	  w[2] = (1 - f[0]) * (f[1]) * (1 - f[2]);
	  w[3] = (1 - f[0]) * (f[1]) * (f[2]);
	  w[4] = (f[0]) * (1 - f[1]) * (1 - f[2]);
	  w[5] = (f[0]) * (1 - f[1]) * (f[2]);
	  w[6] = (f[0]) * (f[1]) * (1 - f[2]);
	  w[7] = (f[0]) * (f[1]) * (f[2]);
	  for(c = 0; c < 3; c++)
	    {
	      tab = o->fcorr + ((size_t) (c * D + pg) * D + n) * n3 + ((size_t) ix[0] * n1 + ix[1]) * n1 + ix[2];
	      acc[c] += mass[n] * sg[c] * (tab[0] * w[0] + tab[1] * w[1] + tab[n1] * w[2] + tab[n1 + 1] * w[3] +
					   tab[(size_t) n1 * n1] * w[4] + tab[(size_t) n1 * n1 + 1] * w[5] +
					   tab[(size_t) n1 * n1 + n1] * w[6] + tab[(size_t) n1 * n1 + n1 + 1] * w[7]);
	    }
	}
      cost++;
    }
  /* forcetree.c:2435-2438: added to the FLOAT results of the tree walk */
  o->P[target].acc[0] += acc[0];
  o->P[target].acc[1] += acc[1];
  o->P[target].acc[2] += acc[2];
  o->P[target].cost += cost;
  return cost;
}

static int tree_evaluate_with_lattice(g2o * o, int target)
{
  int n = tree_evaluate(o, target);
  if(o->periodic && !o->shortrange && o->fcorr)
    lattice_correction(o, target, o->P[target].pos[0], o->P[target].pos[1], o->P[target].pos[2], o->errtol * o->P[target].oldacc);
  return n;
}

struct slice { g2o *o; int lo, hi; double cost; };
static void *walk_slice(void *arg)
{
  struct slice *s = arg;
  int i;
  s->cost = 0;
  for(i = s->lo; i < s->hi; i++)
    if(s->o->P[i].active)
      s->cost += tree_evaluate_with_lattice(s->o, i);
  return NULL;
}

/* the walk loop of gravity_tree (gravtree.c:112-123) over targets [lo,hi), optionally on several threads
 * (targets are independent: forcetree.c:1590-1596) */
double g2o_walk_range(g2o * o, int lo, int hi, int nthreads, double *cost)
{
  pthread_t *th;
  struct slice *sl;
  int t, n = hi - lo;
  double t0 = wall();
  if(nthreads < 1)
    nthreads = 1;
  th = malloc(sizeof(pthread_t) * nthreads);
  sl = malloc(sizeof(struct slice) * nthreads);
  for(t = 0; t < nthreads; t++)
    {
      sl[t].o = o;
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
  free(sl);
  free(th);
  return wall() - t0;
}

/* gravity_tree epilogue, gravtree.c:318-341: OldAcc from the pre-G acceleration, then the G scaling */
void g2o_epilogue(g2o * o, int use_gravpm)
{
  int i, j;
  for(i = 0; i < o->n; i++)
    if(o->P[i].active)
      {
	particle *p = &o->P[i];
	double a[3];
	for(j = 0; j < 3; j++)
	  a[j] = use_gravpm ? p->acc[j] + p->gravpm[j] / o->G : p->acc[j];
	p->oldacc = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
	for(j = 0; j < 3; j++)
	  p->acc[j] *= o->G;
      }
}

/* gravity_tree (gravtree.c:27-460) for one rank: build, walk all active particles, epilogue */
double g2o_gravity_tree(g2o * o, int nthreads)
{
  double cost, t;
  g2o_treebuild(o);
  t = g2o_walk_range(o, 0, o->n, nthreads, &cost);
  o->t_walk = t;
  g2o_epilogue(o, 0);
  return cost;
}

/* ---- force_treeevaluate_direct, forcetree.c:3428-3548 (non-periodic form) ---- */
void g2o_direct(g2o * o, int ntargets, const int *targets, double *acc)
{
  int t, i;
  for(t = 0; t < ntargets; t++)
    {
      const particle *tp = &o->P[targets[t]];
      const int pg = o->t2g[tp->type];
      double ax = 0, ay = 0, az = 0;
      for(i = 0; i < o->n; i++)
	{
	  const particle *sp = &o->P[i];
	  double h = o->fsoft[sp->type] > o->fsoft[tp->type] ? o->fsoft[sp->type] : o->fsoft[tp->type];
	  double dx = sp->pos[0] - (double) tp->pos[0], dy = sp->pos[1] - (double) tp->pos[1], dz = sp->pos[2] - (double) tp->pos[2];
	  double r2 = dx * dx + dy * dy + dz * dz, r = sqrt(r2), fac;
	  if(r * (1 / h) >= 1)
	    fac = accel_fxn(o, pg, o->t2g[sp->type], tp->mass, sp->mass, r2, r, 1) / r;
	  else
	    fac = accel_spline(o, pg, o->t2g[sp->type], tp->mass, sp->mass, h, r, 1);
	  ax += dx * fac;
	  ay += dy * fac;
	  az += dz * fac;
	}
      acc[3 * t] = ax;
      acc[3 * t + 1] = ay;
      acc[3 * t + 2] = az;
    }
}

/* ---- short-range table: performConvolution (ngravs_core.c:72-159) + fill loop (forcetree.c:3274-3354) for the
 *      Newtonian normalised Green's function normed_pgdelta = 1 (ngravs.c:400) or the normalised Yukawa one
 *      (ngravs.c:880-885, ym in table units).  n = 12*NTAB*ol*len - 6*ol*len + 2 samples, Z = 0.5.
 *      The backward DFT of the reference (FFTW-2, not vendored) is restated with a Bluestein transform. ---- */
typedef struct { double re, im; } cpx;
static void fft2(cpx * a, int n, int dir)
{
  int i, j, len;
  for(i = 1, j = 0; i < n; i++)
    {
      int bit = n >> 1;
      for(; j & bit; bit >>= 1)
	j ^= bit;
      j ^= bit;
      if(i < j)
	{
	  cpx t = a[i];
	  a[i] = a[j];
	  a[j] = t;
	}
    }
  for(len = 2; len <= n; len <<= 1)
    {
      int half = len / 2, k;
      for(k = 0; k < half; k++)
	{
	  double ang = dir * 2.0 * M_PI * k / len, wr = cos(ang), wi = sin(ang);
	  for(i = k; i < n; i += len)
	    {
	      cpx u = a[i], v = a[i + half], t;
	      t.re = v.re * wr - v.im * wi;
	      t.im = v.re * wi + v.im * wr;
	      a[i].re = u.re + t.re;
	      a[i].im = u.im + t.im;
	      a[i + half].re = u.re - t.re;
	      a[i + half].im = u.im - t.im;
	    }
	}
    }
}

static void dft_any(const cpx * in, cpx * out, int n, int dir)
{
  int m = 1, k;
  cpx *w, *a, *b;
  while(m < 2 * n - 1)
    m <<= 1;
  w = malloc(sizeof(cpx) * n);
  a = calloc(m, sizeof(cpx));
  b = calloc(m, sizeof(cpx));
  for(k = 0; k < n; k++)
    {
      long long k2 = ((long long) k * k) % (2LL * n);
      double ang = dir * M_PI * (double) k2 / n;
      w[k].re = cos(ang);
      w[k].im = sin(ang);
      a[k].re = in[k].re * w[k].re - in[k].im * w[k].im;
      a[k].im = in[k].re * w[k].im + in[k].im * w[k].re;
    }
  b[0].re = w[0].re;
  b[0].im = -w[0].im;
  for(k = 1; k < n; k++)
    {
      b[k].re = b[m - k].re = w[k].re;
      b[k].im = b[m - k].im = -w[k].im;
    }
  fft2(a, m, -1);
  fft2(b, m, -1);
  for(k = 0; k < m; k++)
    {
      cpx t;
      t.re = a[k].re * b[k].re - a[k].im * b[k].im;
      t.im = a[k].re * b[k].im + a[k].im * b[k].re;
      a[k] = t;
    }
  fft2(a, m, +1);
  for(k = 0; k < n; k++)
    {
      double re = a[k].re / m, im = a[k].im / m;
      out[k].re = re * w[k].re - im * w[k].im;
      out[k].im = re * w[k].im + im * w[k].re;
    }
  free(b);
  free(a);
  free(w);
}

/* kind 0: Newtonian (normed Green's function 1); kind 1: Yukawa with ym (already in table units) */
int g2o_make_srtables(int ntab, int kind, double ym, double *force_tab, double *pot_tab);
int g2o_make_srtable(int ntab, int kind, double ym, double *force_tab)
{
  return g2o_make_srtables(ntab, kind, ym, force_tab, NULL);
}

/* both tables of one pair law: shortrange_fourier_force and shortrange_fourier_pot (forcetree.c:3335-3353) */
int g2o_make_srtables(int ntab, int kind, double ym, double *force_tab, double *pot_tab)
{
  const int len = 3, ol = 8;
  const int n = 12 * ntab * ol * len - 6 * ol * len + 2;	/* ngravs_core.c:179 */
  const double Z = 0.5;
  const double dk = 2.0 * M_PI * ntab * 6.0 * ol / (3.0 * n);	/* jTok, ngravs_core.c:45-48; also `norm` at :114 */
  cpx *in = calloc(n, sizeof(cpx)), *out = calloc(n, sizeof(cpx));
  double *cum = calloc(n / 3 + 2, sizeof(double));
  double sum = 0;
  int j, m, i;
  if(!in || !out || !cum)
    return 1;
  for(j = 0; j < n / 2; j++)
    {
      double k = dk * j, k2 = k * k, gk = 1.0;
      if(kind == 1)
	gk = k2 / (k2 + ym * ym) * exp(-ym * ym * 0.25);
      in[j].re = gk * exp(-k2 * Z * Z);		/* fourierIntegrand, ngravs_core.c:65-70 */
      if(j > 0)
	in[n - j].re = in[j].re;
    }
  dft_any(in, out, n, +1);
  /* cumulative Newton-Cotes 3/8 integral of the transform (ngravs_core.c:131-141); x spacing = mTox(1) */
  for(m = 0; m < n - 3; m += 3)
    {
      double dx = 3.0 * (m + 3) / (6.0 * ntab * ol) - 3.0 * m / (6.0 * ntab * ol);
      sum += dx * 0.125 * dk * (out[m].re + 3.0 * out[m + 1].re + 3.0 * out[m + 2].re + out[m + 3].re);
      cum[m / 3 + 1] = sum;
    }
  for(i = 0; i < ntab; i++)
    {
      int f = ol * (6 * i + 3);			/* gadgetToFourier, ngravs_core.c:55-58 */
      double temp = out[f].re * dk, tempI = cum[f / 3];
      double u = 3.0 / ntab * (i + 0.5);		/* forcetree.c:3335 */
      if(force_tab)
	force_tab[i] = tempI / (u * u) - temp / u;	/* forcetree.c:3345-3353 */
      if(pot_tab)
	pot_tab[i] = temp / u;			/* forcetree.c:3342-3346 */
    }
  free(cum);
  free(out);
  free(in);
  return 0;
}

/* ---- read-back ---- */
void g2o_get_domain(g2o * o, double *out)
{
  int k;
  for(k = 0; k < 3; k++)
    {
      out[k] = o->corner[k];
      out[3 + k] = o->center[k];
    }
  out[6] = o->len;
  out[7] = o->fac;
}

int g2o_counts(g2o * o, int *out)
{
  out[0] = o->n; out[1] = o->ntop; out[2] = o->ntopleaves; out[3] = o->numnodes;
  return 0;
}

void g2o_get_particles(g2o * o, float *pos, float *mass, int *type, int *id, long long *key, float *acc, float *cost, float *oldacc, double *accd)
{
  int i, k;
  for(i = 0; i < o->n; i++)
    {
      const particle *p = &o->P[i];
      for(k = 0; k < 3; k++)
	{
	  if(pos) pos[3 * i + k] = p->pos[k];
	  if(acc) acc[3 * i + k] = p->acc[k];
	  if(accd) accd[3 * i + k] = p->accd[k];
	}
      if(mass) mass[i] = p->mass;
      if(type) type[i] = p->type;
      if(id) id[i] = p->id;
      if(key) key[i] = p->key;
      if(cost) cost[i] = p->cost;
      if(oldacc) oldacc[i] = p->oldacc;
    }
}

void g2o_get_topnodes(g2o * o, long long *tn /* 5 per node */, int *dni)
{
  int i;
  for(i = 0; i < o->ntop; i++)
    {
      tn[5 * i] = o->top[i].daughter;
      tn[5 * i + 1] = o->top[i].leaf;
      tn[5 * i + 2] = o->top[i].size;
      tn[5 * i + 3] = o->top[i].startkey;
      tn[5 * i + 4] = o->top[i].count;
    }
  for(i = 0; i < o->ntopleaves; i++)
    dni[i] = o->maxpart + o->dni[i];
}

void g2o_get_tree(g2o * o, float *len, float *center, float *s, float *mass, int *link, int *nextnode, int *father)
{
  int k, j, g, D = o->D;
  for(k = 0; k < o->numnodes; k++)
    {
      const node *nd = &o->nodes[k];
      len[k] = nd->len;
      for(j = 0; j < 3; j++)
	{
	  center[3 * k + j] = nd->center[j];
	  for(g = 0; g < D; g++)
	    s[(3 * k + j) * D + g] = nd->s[j][g];
	}
      for(g = 0; g < D; g++)
	mass[k * D + g] = nd->mass[g];
      link[4 * k] = nd->bitflags;
      link[4 * k + 1] = nd->sibling;
      link[4 * k + 2] = nd->nextnode;
      link[4 * k + 3] = nd->father;
    }
  for(k = 0; k < o->n; k++)
    {
      nextnode[k] = o->nextnode[k];
      father[k] = o->father[k];
    }
}

void g2o_timings(g2o * o, double *out)
{
  out[0] = o->t_domain;
  out[1] = o->t_build;
  out[2] = o->t_walk;
}

/* ---- design aid (not a reference restatement): statistics of a GROUP walk, in which `gsize` consecutive targets (in
 *      depth-first tree order) share one cursor and a node is descended into if ANY member opens it -- the traversal the
 *      GPU walk kernel performs with gsize = 32.  Returns per group: node visits and the sum over visits of members that
 *      were awake (not skipping an accepted/culled ancestor).  Decisions are those of tree_evaluate(). ---- */
static int node_decision(g2o * o, const particle * tp, const node * nop)	/* 0 cull, 1 accept, 2 open */
{
  const int D = o->D, SR = o->shortrange, PER = o->periodic;
  const double boxsize = o->boxsize, boxhalf = 0.5 * o->boxsize;
  double r2min = INFINITY, r2max = -INFINITY, summass = 0, h;
  const double pp[3] = { tp->pos[0], tp->pos[1], tp->pos[2] };
  const double aold = o->errtol * tp->oldacc;
  int g, k;
  for(g = 0; g < D; g++)
    {
      double d[3], r2 = 0;
      for(k = 0; k < 3; k++)
	{
	  d[k] = nop->s[k][g] - pp[k];
	  if(PER)
	    d[k] = NEAREST(d[k]);
	  r2 += d[k] * d[k];
	}
      summass += nop->mass[g];
      if(r2 < r2min) r2min = r2;
      if(r2 > r2max) r2max = r2;
    }
  if(SR && r2min > o->rcut * o->rcut)
    {
      double eff = o->rcut + 0.5 * nop->len;
      for(k = 0; k < 3; k++)
	{
	  double dist = nop->center[k] - pp[k];
	  if(PER)
	    dist = NEAREST(dist);
	  if(dist < -eff || dist > eff)
	    return 0;
	}
    }
  if(o->theta)
    {
      if(nop->len * nop->len > r2min * o->theta * o->theta)
	return 2;
    }
  else
    {
      if(summass * nop->len * nop->len > r2min * r2min * aold)
	return 2;
      if(fabs(nop->center[0] - pp[0]) < 0.60 * nop->len && fabs(nop->center[1] - pp[1]) < 0.60 * nop->len && fabs(nop->center[2] - pp[2]) < 0.60 * nop->len)
	return 2;
    }
  if(o->unequal)
    {
      int mt = (nop->bitflags >> 2) & 7;
      h = o->fsoft[tp->type];
      if(mt == 7)
	return 2;
      if(h < o->fsoft[mt])
	{
	  h = o->fsoft[mt];
	  if(r2max < h * h && ((nop->bitflags >> 5) & 1))
	    return 2;
	}
    }
  return 1;
}

/* order[] = targets in depth-first tree order (filled by following nextnode from the root) */
/* design aid: deferred dense evaluation of the accepted terms.  Every accepted species term / particle term is staged in a ring of
 * g2o_sim_ring term sources shared by the group and appended to the accepting member's list; when the ring is full all lists are
 * evaluated (one term per member per pass).  out[14] = passes per group, out[15] = member terms per group. */
int g2o_sim_ring = 128;
int g2o_sim_order = 0; int *g2o_sim_custom_order = 0;	/* 0 = depth-first tree order, 1 = Peano-Hilbert key order of all species */
double g2o_group_visits[4096];
static g2o *cmp_o_;
static int cmp_phkey_(const void *a, const void *b) { peanokey ka = cmp_o_->P[*(const int *) a].key, kb = cmp_o_->P[*(const int *) b].key; return ka < kb ? -1 : ka > kb; }
double g2o_hist_awake[65], g2o_hist_open[65], g2o_hist_awake_w[65];
void g2o_groupwalk_stats(g2o * o, int gsize, int ngroups_max, double *out)
{
  double sim_passes = 0, sim_terms = 0; int sim_staged = 0; int *sim_n = calloc(gsize, sizeof(int));
#define SIM_FLUSH() do { int mx_ = 0, m_; for(m_ = 0; m_ < gsize; m_++) { if(sim_n[m_] > mx_) mx_ = sim_n[m_]; sim_terms += sim_n[m_]; sim_n[m_] = 0; } sim_passes += mx_; sim_staged = 0; } while(0)
  const int MP = o->maxpart, n = o->n;
  int *order = malloc(sizeof(int) * n), *skip = malloc(sizeof(int) * gsize);
  int no = MP, cnt = 0, gi, ngroups = 0;
  double visits = 0, awake = 0, accepts = 0, pvis = 0, pawake = 0;
  double v_noaccept = 0, v_chain = 0, v_allcull = 0, v_allaccept = 0, opens = 0, culls = 0, v_leafonly = 0, passes = 0, ppasses = 0;
  /* DFS order of particles: open everything */
  while(no >= 0)
    {
      if(no < MP)
	{
	  order[cnt++] = no;
	  no = o->nextnode[no];
	}
      else
	no = NODE_OF(o, no)->nextnode;
    }
  if(g2o_sim_order == 1) { cmp_o_ = o; qsort(order, n, sizeof(int), cmp_phkey_); }
  if(g2o_sim_order == 2 && g2o_sim_custom_order) memcpy(order, g2o_sim_custom_order, sizeof(int) * n);
  /* pre-order rank of every node/particle to implement "sleep until the cursor leaves the subtree" */
  {
    int stride = n / (gsize * ngroups_max);
    if(stride < 1)
      stride = 1;
    for(gi = 0; gi + gsize <= n && ngroups < ngroups_max; gi += gsize * stride, ngroups++)
      {
	int m, cur = MP; double v0_ = visits;
	/* skip[m] = node index at which member m wakes up again (-2 = awake) */
	for(m = 0; m < gsize; m++)
	  skip[m] = -2;
	    SIM_FLUSH();
	while(cur >= 0)
	  {
	    if(cur < MP)
	      {			/* a particle: members that are awake interact */
		int na = 0;
		for(m = 0; m < gsize; m++)
		  {
		    if(skip[m] == cur)
		      skip[m] = -2;
		    if(skip[m] == -2)
		      na++;
		  }
		pvis += 1;
		pawake += na;
		    if(na) { for(m = 0; m < gsize; m++) if(skip[m] == -2) sim_n[m]++; if(++sim_staged >= g2o_sim_ring) SIM_FLUSH(); }
		ppasses += (na + 31) / 32;
		cur = o->nextnode[cur];
		continue;
	      }
	    {
	      node *nop = NODE_OF(o, cur);
	      int anyopen = 0, na = 0, nacc = 0, ncull = 0, nopen = 0;
	      for(m = 0; m < gsize; m++)
		{
		  if(skip[m] == cur)
		    skip[m] = -2;
		  if(skip[m] == -2)
		    {
		      int dec = node_decision(o, &o->P[order[gi + m]], nop);
		      na++;
		      if(dec == 2)
			{
			  anyopen = 1;
			  nopen++;
			}
		      else
			{
			  skip[m] = nop->sibling;	/* wake up at the sibling (-1: never) */
			  if(dec == 1)
			    {
			      accepts += 1;
			      nacc++;
				  sim_n[m] += o->D;
			    }
			  else
			    ncull++;
			}
		    }
		}
	      visits += 1;
	      awake += na;
		  if(gsize <= 64) { g2o_hist_awake[na] += 1; g2o_hist_open[nopen] += 1; }
		  if(nacc) { sim_staged += o->D; if(sim_staged >= g2o_sim_ring) SIM_FLUSH(); }
	      passes += (na + 31) / 32;
	      opens += nopen;
	      culls += ncull;
	      if(nacc == 0)
		v_noaccept += 1;
	      if(ncull == na)
		v_allcull += 1;
	      if(nacc == na)
		v_allaccept += 1;
	      {
		/* chain node: exactly one child, which is a node (the next in the walk is a node whose sibling equals ours) */
		int nx = nop->nextnode;
		if(nx >= MP && NODE_OF(o, nx)->sibling == nop->sibling)
		  v_chain += 1;
	      }
	      if(anyopen)
		cur = nop->nextnode;
	      else
		cur = nop->sibling;
	      /* members sleeping until a node that is skipped over by a sibling jump wake at that jump target:
	         their wake node is an ancestor's sibling, which the cursor reaches later or equals cur */
	      for(m = 0; m < gsize; m++)
		if(skip[m] == -1 && cur == -1)
		  skip[m] = -2;
	    }
	  }
	if(ngroups < 4096)
	  g2o_group_visits[ngroups] = visits - v0_;
      }
  }
  out[0] = ngroups;
  out[1] = visits / ngroups;
  out[2] = awake / ngroups;
  out[3] = accepts / ngroups;
  out[4] = pvis / ngroups;
  out[5] = pawake / ngroups;
  out[6] = v_noaccept / ngroups;
  out[7] = v_chain / ngroups;
  out[8] = v_allcull / ngroups;
  out[9] = v_allaccept / ngroups;
  out[10] = opens / ngroups;
  out[11] = culls / ngroups;
  out[12] = passes / ngroups;	/* 32-lane passes if awake members are packed onto lanes */
  out[13] = ppasses / ngroups;
  SIM_FLUSH();
  out[14] = sim_passes / ngroups;
  out[15] = sim_terms / ngroups;
  free(sim_n);
  free(skip);
  free(order);
}
