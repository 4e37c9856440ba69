// g2_laws.cuh — device versions of the ngravs pair force laws (ngravs.c:344-861) in FP32.
//
// Calling convention of the reference (allvars.h:134, forcetree.c:1542-1544):
//   r >= h :  fac = AccelFxns[tgt][src](pmass, m, r2, r, N) / r
//   r <  h :  fac = AccelSplines[tgt][src](pmass, m, h, r, N)           (already contains the 1/r)
//   acc   += d * fac
// accel_over_r() returns the first form, spline() the second.  `N` is the particle count of that species in
// the node under NGRAVS_ACCUMULATOR and 1 otherwise (forcetree.c:1555-1569); we pass 1.
#pragma once
#include "g2_common.cuh"

#define G2_PI_F 3.14159265358979323846f

// GADGET-2 spline-softened kernel, ngravs.c:420-434 (h = 2.8 eps; u = r/h)
__device__ __forceinline__ float law_plummer(float m, float h, float r)
{
  float hinv = __frcp_rn(h);
  float u = r * hinv;
  float h3 = hinv * hinv * hinv;
  float f;
  if(u < 0.5f)
    f = 10.666666666667f + u * u * (32.0f * u - 38.4f);
  else
    f = 21.333333333333f - 48.0f * u + 38.4f * u * u - 10.666666666667f * u * u * u - 0.066666666667f / (u * u * u);
  return m * h3 * f;
}

// BAM family, ngravs.c:495-670.  eta, rho depend on which side is the BAM halo.  The closed form
// atan(x)/x^2 - 1/(x(1+x^2)) cancels badly just above the reference's Taylor switch (x = 0.1), so this (rarely wired)
// law is evaluated in FP64 and rounded once.
__device__ __forceinline__ float bam_core_over_r(float rhof, float etaf, float rf)	// law(r)/r for r >= h
{
  const double rho = rhof, eta = etaf, r = rf;
  const double reta = r * eta, reta2 = reta * reta, eta3 = eta * eta * eta;
  if(reta < 0.1)
    return (float) (rho * eta3 * (2.0 / 3.0 - 4.0 * reta2 / 5.0 + 6.0 * reta2 * reta2 / 7.0));
  return (float) (rho * eta3 * (atan(reta) / (reta2 * eta) - 1.0 / (reta * eta * (1.0 + reta2))) / r);
}
__device__ __forceinline__ float bam_core_spline(float rhof, float etaf, float rf)
{
  const double rho = rhof, eta = etaf, r = rf;
  const double reta = r * eta, reta2 = reta * reta, eta3 = eta * eta * eta;
  if(reta < 0.1)
    return (float) (rho * eta3 * (2.0 / 3.0 - 4.0 * reta2 / 5.0 + 6.0 * reta2 * reta2 / 7.0));
  return (float) (rho * eta3 * (atan(reta) / (reta2 * reta) - 1.0 / (reta2 * (1.0 + reta2))));
}

// AccelFxns[..](pm, m, r2, r, N) / r      (par[0] = Yukawa inverse range ym, par[1] = BAM_EPSILON)
__device__ __forceinline__ float accel_over_r(int id, const float *par, float pm, float m, float r2, float r, float rinv, float nn)
{
  switch (id)
    {
    case G2GPU_LAW_NEWTONIAN:
      return m * rinv * rinv * rinv;	// ngravs.c:351  source/r2, then /r
    case G2GPU_LAW_NEG_NEWTONIAN:
      return -m * rinv * rinv * rinv;	// ngravs.c:359
    case G2GPU_LAW_YUKAWA:
      {				// ngravs.c:856-861: source*exp(-r*ym)*(ym/r + 1/r2)
	float ym = par[0];
	return m * expf(-r * ym) * (ym * rinv + rinv * rinv) * rinv;
      }
    case G2GPU_LAW_COLOYUK:
      {				// ngravs.c:826
	float ym = par[0];
	return (m * expf(-r * ym) * (ym * rinv + rinv * rinv) + m * rinv * rinv) * rinv;
      }
    case G2GPU_LAW_BAMBAM:
      {				// ngravs.c:495-529
	float eta = 4.0f * G2_PI_F * par[1] / (pm + m / nn);
	return bam_core_over_r(2.0f * pm * m / G2_PI_F, eta, r);
      }
    case G2GPU_LAW_SOURCEBAMBARYON:
      {				// ngravs.c:590-614
	float eta = 4.0f * G2_PI_F * par[1] * nn / m;
	return bam_core_over_r(2.0f * pm * m / G2_PI_F, eta, r);
      }
    case G2GPU_LAW_SOURCEBARYONBAM:
      {				// ngravs.c:646-670
	float eta = 4.0f * G2_PI_F * par[1] / pm;
	return bam_core_over_r(2.0f * pm * m / G2_PI_F, eta, r);
      }
    default:			// G2GPU_LAW_NONE, ngravs.c:344
      return 0.0f;
    }
}

// AccelSplines[..](pm, m, h, r, N)
__device__ __forceinline__ float accel_spline(int id, const float *par, float pm, float m, float h, float r, float nn)
{
  switch (id)
    {
    case G2GPU_SPLINE_PLUMMER:
      return law_plummer(m, h, r);
    case G2GPU_SPLINE_NEG_PLUMMER:
      return -law_plummer(m, h, r);
    case G2GPU_SPLINE_BAMBAM:
      {				// ngravs.c:531-560
	float eta = 4.0f * G2_PI_F * par[1] / (pm + m / nn);
	return bam_core_spline(2.0f * pm * m / G2_PI_F, eta, r);
      }
    case G2GPU_SPLINE_SOURCEBAMBARYON:
      {				// ngravs.c:562-588
	float eta = 4.0f * G2_PI_F * par[1] * nn / m;
	return bam_core_spline(2.0f * pm * m / G2_PI_F, eta, r);
      }
    case G2GPU_SPLINE_SOURCEBARYONBAM:
      {				// ngravs.c:616-644
	float eta = 4.0f * G2_PI_F * par[1] / pm;
	return bam_core_spline(2.0f * pm * m / G2_PI_F, eta, r);
      }
    default:
      return 0.0f;
    }
}
