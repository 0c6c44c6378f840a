/*
 * eid_detmath.h — bit-reproducible replacements for the GLSL built-ins / libm calls on the
 * hot path (sin, cos, exp, pow, acos, asin, atan2).
 *
 * Why: the parity contract is "reservoir picks and G-buffer ids
 * bit-exact, radiance within 1e-3" between the CUDA kernels and the CPU oracle.  IEEE-754
 * add/mul/div/sqrt round identically on x86 SSE and on sm_100a *if no FMA contraction happens*,
 * but vendor transcendental functions do not.  Every function below is built only from
 * + - * / and bit operations in a fixed order, so host and device agree to the last bit.
 *
 * Rules for users: compile host code with -ffp-contract=off and device code with --fmad=false
 * (the EID_MUL/EID_ADD wrappers additionally pin the rounding on the device).
 *
 * This header plays the role libm / the GLSL built-in library plays for the reference; it is
 * not part of the reference's own algorithm.
 */
#ifndef EIDOLA_DETMATH_H
#define EIDOLA_DETMATH_H

#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define EID_HD __host__ __device__ __forceinline__
#else
#define EID_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define EID_MUL(a, b) __fmul_rn((a), (b))
#define EID_ADD(a, b) __fadd_rn((a), (b))
#define EID_SUB(a, b) __fsub_rn((a), (b))
#define EID_DIV(a, b) __fdiv_rn((a), (b))
#else
#define EID_MUL(a, b) ((a) * (b))
#define EID_ADD(a, b) ((a) + (b))
#define EID_SUB(a, b) ((a) - (b))
#define EID_DIV(a, b) ((a) / (b))
#endif

EID_HD float eid_u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
  return __uint_as_float(u);
#else
  float f; memcpy(&f, &u, 4); return f;
#endif
}
EID_HD uint32_t eid_f2u(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}

/* floor for |x| < 2^31 without calling libm (exact) */
EID_HD float eid_floorf(float x) {
  if (!(x > -2147483000.f && x < 2147483000.f)) return x;
  int32_t i = (int32_t)x;            /* trunc toward zero */
  float f = (float)i;
  return (f > x) ? EID_SUB(f, 1.0f) : f;
}

/* ---- sin / cos -------------------------------------------------------------------------
 * Cody–Waite reduction by pi/2 (three-term split, exact for |k| < 2^15), then the classic
 * single-precision minimax polynomials on [-pi/4, pi/4].  Max error ~1 ulp near 0, <2e-7 abs. */
EID_HD void eid_sincosf(float x, float* s, float* c) {
  const float TWO_OVER_PI = 0.636619772367581343f;
  const float P1 = 1.5703125f;                 /* 0x3FC90000 */
  const float P2 = 4.837512969970703125e-4f;   /* 0x39FDAA00 */
  const float P3 = 7.54978995489188e-8f;
  float kf = eid_floorf(EID_ADD(EID_MUL(x, TWO_OVER_PI), 0.5f));
  int32_t k = (int32_t)kf;
  float r = EID_SUB(x, EID_MUL(kf, P1));
  r = EID_SUB(r, EID_MUL(kf, P2));
  r = EID_SUB(r, EID_MUL(kf, P3));
  float z = EID_MUL(r, r);
  /* sin(r) = r + r*z*(S0 + z*(S1 + z*S2)) */
  float ps = EID_ADD(EID_MUL(-1.9515295891e-4f, z), 8.3321608736e-3f);
  ps = EID_ADD(EID_MUL(ps, z), -1.6666654611e-1f);
  float sr = EID_ADD(EID_MUL(EID_MUL(ps, z), r), r);
  /* cos(r) = 1 - z/2 + z*z*(C0 + z*(C1 + z*C2)) */
  float pc = EID_ADD(EID_MUL(2.443315711809948e-5f, z), -1.388731625493765e-3f);
  pc = EID_ADD(EID_MUL(pc, z), 4.166664568298827e-2f);
  float cr = EID_ADD(EID_SUB(1.0f, EID_MUL(0.5f, z)), EID_MUL(EID_MUL(pc, z), z));
  switch (k & 3) {
    case 0: *s = sr;  *c = cr;  break;
    case 1: *s = cr;  *c = -sr; break;
    case 2: *s = -sr; *c = -cr; break;
    default: *s = -cr; *c = sr; break;
  }
}
EID_HD float eid_sinf(float x) { float s, c; eid_sincosf(x, &s, &c); return s; }
EID_HD float eid_cosf(float x) { float s, c; eid_sincosf(x, &s, &c); return c; }

/* ---- exp --------------------------------------------------------------------------------
 * x = k*ln2 + r, |r| <= ln2/2 ; e^r by a degree-6 Taylor-like minimax ; result scaled by 2^k
 * through the exponent field.  Underflows to 0 below -87.3, overflows to +inf above 88.7. */
EID_HD float eid_expf(float x) {
  if (x != x) return x;
  if (x > 88.72f) return eid_u2f(0x7f800000u);
  if (x < -87.33f) return 0.0f;
  const float LOG2E = 1.44269504088896341f;
  const float LN2_HI = 0.693359375f;           /* 0x3F318000 */
  const float LN2_LO = -2.12194440e-4f;
  float kf = eid_floorf(EID_ADD(EID_MUL(x, LOG2E), 0.5f));
  int32_t k = (int32_t)kf;
  float r = EID_SUB(x, EID_MUL(kf, LN2_HI));
  r = EID_SUB(r, EID_MUL(kf, LN2_LO));
  float z = EID_MUL(r, r);
  float p = 1.9875691500e-4f;
  p = EID_ADD(EID_MUL(p, r), 1.3981999507e-3f);
  p = EID_ADD(EID_MUL(p, r), 8.3334519073e-3f);
  p = EID_ADD(EID_MUL(p, r), 4.1665795894e-2f);
  p = EID_ADD(EID_MUL(p, r), 1.6666665459e-1f);
  p = EID_ADD(EID_MUL(p, r), 5.0000001201e-1f);
  float e = EID_ADD(EID_ADD(EID_MUL(p, z), r), 1.0f);
  /* scale by 2^k in two steps so k in [-126-23, 128] stays representable */
  int32_t k1 = k / 2, k2 = k - k1;
  e = EID_MUL(e, eid_u2f((uint32_t)(k1 + 127) << 23));
  e = EID_MUL(e, eid_u2f((uint32_t)(k2 + 127) << 23));
  return e;
}

/* ---- log (natural) — used only by eid_powf ------------------------------------------------ */
EID_HD float eid_logf(float x) {
  if (x != x || x < 0.0f) return eid_u2f(0x7fc00000u);
  if (x == 0.0f) return eid_u2f(0xff800000u);
  if (x == eid_u2f(0x7f800000u)) return x;
  int32_t e = 0;
  uint32_t ux = eid_f2u(x);
  if (ux < 0x00800000u) { x = EID_MUL(x, 8388608.0f); ux = eid_f2u(x); e = -23; }
  e += (int32_t)(ux >> 23) - 127;
  float m = eid_u2f((ux & 0x007fffffu) | 0x3f800000u);   /* [1,2) */
  if (m > 1.41421356237f) { m = EID_MUL(m, 0.5f); e += 1; }
  float f = EID_SUB(m, 1.0f);
  float z = EID_MUL(f, f);
  float p = 7.0376836292e-2f;
  p = EID_ADD(EID_MUL(p, f), -1.1514610310e-1f);
  p = EID_ADD(EID_MUL(p, f), 1.1676998740e-1f);
  p = EID_ADD(EID_MUL(p, f), -1.2420140846e-1f);
  p = EID_ADD(EID_MUL(p, f), 1.4249322787e-1f);
  p = EID_ADD(EID_MUL(p, f), -1.6668057665e-1f);
  p = EID_ADD(EID_MUL(p, f), 2.0000714765e-1f);
  p = EID_ADD(EID_MUL(p, f), -2.4999993993e-1f);
  p = EID_ADD(EID_MUL(p, f), 3.3333331174e-1f);
  float y = EID_MUL(EID_MUL(f, z), p);
  float ef = (float)e;
  y = EID_ADD(y, EID_MUL(ef, -2.12194440e-4f));
  y = EID_SUB(y, EID_MUL(0.5f, z));
  float r = EID_ADD(f, y);
  r = EID_ADD(r, EID_MUL(ef, 0.693359375f));
  return r;
}

/* pow for x >= 0 (GLSL pow is undefined for x < 0): exp(y * log x). ~1e-6 relative. */
EID_HD float eid_powf(float x, float y) {
  if (x == 0.0f) return (y > 0.0f) ? 0.0f : ((y == 0.0f) ? 1.0f : eid_u2f(0x7f800000u));
  return eid_expf(EID_MUL(y, eid_logf(x)));
}

/* sqrt is IEEE on both sides; declared here so call sites read uniformly */
#if defined(__CUDA_ARCH__)
#define eid_sqrtf(x) __fsqrt_rn(x)
#else
#include <math.h>
#define eid_sqrtf(x) sqrtf(x)
#endif

/* ---- atan / atan2 / asin / acos (environment lookups) -------------------------------------- */
EID_HD float eid_atanf_pos(float x) { /* x >= 0 */
  const float PIO2 = 1.57079632679489661923f, PIO4 = 0.785398163397448309616f;
  float y0, t;
  if (x > 2.414213562373095f) { y0 = PIO2; t = EID_DIV(-1.0f, x); }
  else if (x > 0.4142135623730950f) { y0 = PIO4; t = EID_DIV(EID_SUB(x, 1.0f), EID_ADD(x, 1.0f)); }
  else { y0 = 0.0f; t = x; }
  float z = EID_MUL(t, t);
  float p = 8.05374449538e-2f;
  p = EID_ADD(EID_MUL(p, z), -1.38776856032e-1f);
  p = EID_ADD(EID_MUL(p, z), 1.99777106478e-1f);
  p = EID_ADD(EID_MUL(p, z), -3.33329491539e-1f);
  float r = EID_ADD(EID_MUL(EID_MUL(p, z), t), t);
  return EID_ADD(y0, r);
}
EID_HD float eid_atanf(float x) { return (x < 0.0f) ? -eid_atanf_pos(-x) : eid_atanf_pos(x); }
EID_HD float eid_atan2f(float y, float x) {
  const float PI = 3.14159265358979323846f, PIO2 = 1.57079632679489661923f;
  if (x != x || y != y) return eid_u2f(0x7fc00000u);
  if (x == 0.0f) { if (y > 0.0f) return PIO2; if (y < 0.0f) return -PIO2; return 0.0f; }
  float a = eid_atanf(EID_DIV(y, x));
  if (x > 0.0f) return a;
  return (y >= 0.0f) ? EID_ADD(a, PI) : EID_SUB(a, PI);
}
EID_HD float eid_asinf(float x) {
  const float PIO2 = 1.57079632679489661923f;
  float ax = (x < 0.0f) ? -x : x;
  if (ax > 1.0f) return eid_u2f(0x7fc00000u);
  float r;
  if (ax > 0.5f) {
    float z = EID_MUL(0.5f, EID_SUB(1.0f, ax));
    float sq = eid_sqrtf(z);
    float p = 4.2163199048e-2f;
    p = EID_ADD(EID_MUL(p, z), 2.4181311049e-2f);
    p = EID_ADD(EID_MUL(p, z), 4.5470025998e-2f);
    p = EID_ADD(EID_MUL(p, z), 7.4953002686e-2f);
    p = EID_ADD(EID_MUL(p, z), 1.6666752422e-1f);
    float a = EID_ADD(EID_MUL(EID_MUL(p, z), sq), sq);
    r = EID_SUB(PIO2, EID_ADD(a, a));
  } else {
    float z = EID_MUL(ax, ax);
    float p = 4.2163199048e-2f;
    p = EID_ADD(EID_MUL(p, z), 2.4181311049e-2f);
    p = EID_ADD(EID_MUL(p, z), 4.5470025998e-2f);
    p = EID_ADD(EID_MUL(p, z), 7.4953002686e-2f);
    p = EID_ADD(EID_MUL(p, z), 1.6666752422e-1f);
    r = EID_ADD(EID_MUL(EID_MUL(p, z), ax), ax);
  }
  return (x < 0.0f) ? -r : r;
}
EID_HD float eid_acosf(float x) {
  const float PIO2 = 1.57079632679489661923f;
  return EID_SUB(PIO2, eid_asinf(x));
}

#endif /* EIDOLA_DETMATH_H */
