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
