/* oracle/ref/ref_forcetree_unit.c — TEST INFRASTRUCTURE ONLY.
 * Compiles the reference's forcetree.c UNMODIFIED (textual include from /root/reference via -I) and
 * adds one accessor for its file-scope static short-range table (forcetree.c:33), which has no
 * external linkage. */
#include "forcetree.c"

int g2ref_srtable_copy(double *out)
{
#ifdef PMGRID
  memcpy(out, shortrange_fourier_force, sizeof(double) * N_GRAVS * N_GRAVS * NTAB);
  return N_GRAVS * N_GRAVS * NTAB;
#else
  (void) out;
  return 0;
#endif
}

/* the potential table shortrange_fourier_pot[tgt][src][NTAB] (forcetree.c:34, filled at 3346) */
int g2ref_srpot_copy(double *out)
{
#ifdef PMGRID
  memcpy(out, shortrange_fourier_pot, sizeof(double) * N_GRAVS * N_GRAVS * NTAB);
  return N_GRAVS * N_GRAVS * NTAB;
#else
  (void) out;
  return 0;
#endif
}

/* the lattice-sum force tables fcorrx/y/z (forcetree.c:49-53), file-scope statics like the tables above */
int g2ref_lattice_copy(double *out)
{
#if defined(PERIODIC) && !defined(PMGRID)
  const size_t n3 = (size_t) (NGRAVS_EN + 1) * (NGRAVS_EN + 1) * (NGRAVS_EN + 1);
  int l, m;
  size_t q;
  for(l = 0; l < N_GRAVS; l++)
    for(m = 0; m < N_GRAVS; m++)
      for(q = 0; q < n3; q++)
	{
	  out[((size_t) (0 * N_GRAVS + l) * N_GRAVS + m) * n3 + q] = (&fcorrx[l][m][0][0][0])[q];
	  out[((size_t) (1 * N_GRAVS + l) * N_GRAVS + m) * n3 + q] = (&fcorry[l][m][0][0][0])[q];
	  out[((size_t) (2 * N_GRAVS + l) * N_GRAVS + m) * n3 + q] = (&fcorrz[l][m][0][0][0])[q];
	}
  return NGRAVS_EN + 1;
#else
  (void) out;
  return 0;
#endif
}

/* potcorr[tgt][src] (forcetree.c:52) after lattice_init, i.e. divided by BoxSize (forcetree.c:3759) */
int g2ref_potcorr_copy(double *out)
{
#if defined(PERIODIC) && !defined(PMGRID)
  memcpy(out, potcorr, sizeof(potcorr));
  return NGRAVS_EN + 1;
#else
  (void) out;
  return 0;
#endif
}
