// urf_math.cuh — IEEE-exact scalar helpers and bit-exact emulation of the three libm functions the reference's hot path
// calls per point: acosf / asinf (lidar_segmentation.cpp:162,165,256-268; x_zero_method.cpp:58; z_zero_method.cpp:63)
// and atan2f (star_shaped_search.cpp:166).
//
// Why: labels depend on these results through threshold compares and integer-degree bins, and CUDA's own
// acosf/asinf/atan2f are not bit-identical to the glibc 2.39 (x86-64) routines the reference links against. glibc 2.39's
// float versions are the classic fdlibm-derived single-precision algorithms (sysdeps/ieee754/flt-32/{e_asinf,e_acosf,
// e_atan2f,s_atanf}.c): only float +,-,*,/ and sqrtf, no FMA, no double. Restated here operation for operation with
// round-to-nearest intrinsics so the compiler can neither contract nor reassociate them.
//
// The same header compiles for the host (plain C++ with -ffp-contract=off) so tests can sweep it exhaustively against
// the container's real libm (tests/test_math_kat.py); the device build is checked against host libm on the GPU.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define URF_HD __host__ __device__ __forceinline__
#define URF_HDM __host__ __device__ __forceinline__   /* member functions */
#else
#define URF_HD static inline
#define URF_HDM inline
#endif

#if defined(__CUDA_ARCH__)
#define URF_FADD(a, b) __fadd_rn((a), (b))
#define URF_FSUB(a, b) __fsub_rn((a), (b))
#define URF_FMUL(a, b) __fmul_rn((a), (b))
#define URF_FDIV(a, b) __fdiv_rn((a), (b))
#define URF_FSQRT(a) __fsqrt_rn((a))
#define URF_DADD(a, b) __dadd_rn((a), (b))
#define URF_DSUB(a, b) __dsub_rn((a), (b))
#define URF_DMUL(a, b) __dmul_rn((a), (b))
#define URF_DDIV(a, b) __ddiv_rn((a), (b))
#define URF_DSQRT(a) __dsqrt_rn((a))
#define URF_DFMA(a, b, c) __fma_rn((a), (b), (c))
#define URF_FFMA(a, b, c) __fmaf_rn((a), (b), (c))
#define URF_F2I(x) __float_as_int((x))
#define URF_I2F(x) __int_as_float((x))
#define URF_FABS(x) fabsf((x))
#else
#include <math.h>
// Host build: compile with -ffp-contract=off (x86-64 SSE2 float/double arithmetic is IEEE round-to-nearest).
#define URF_FADD(a, b) ((float)(a) + (float)(b))
#define URF_FSUB(a, b) ((float)(a) - (float)(b))
#define URF_FMUL(a, b) ((float)(a) * (float)(b))
#define URF_FDIV(a, b) ((float)(a) / (float)(b))
#define URF_FSQRT(a) sqrtf((a))
#define URF_DADD(a, b) ((double)(a) + (double)(b))
#define URF_DSUB(a, b) ((double)(a) - (double)(b))
#define URF_DMUL(a, b) ((double)(a) * (double)(b))
#define URF_DDIV(a, b) ((double)(a) / (double)(b))
#define URF_DSQRT(a) sqrt((a))
#define URF_DFMA(a, b, c) fma((double)(a), (double)(b), (double)(c))   /* correctly rounded with or without hardware FMA */
#define URF_FFMA(a, b, c) fmaf((float)(a), (float)(b), (float)(c))
static inline int32_t urf_f2i_(float x) { int32_t i; memcpy(&i, &x, 4); return i; }
static inline float urf_i2f_(int32_t i) { float x; memcpy(&x, &i, 4); return x; }
#define URF_F2I(x) urf_f2i_((x))
#define URF_I2F(x) urf_i2f_((x))
#define URF_FABS(x) fabsf((x))
#endif

#define URF_PI_D 3.14159265358979323846 /* M_PI */

namespace urfm {

// v / M_PI in double, correctly rounded, without the divide: q = RN(v * RN(1/pi)), one Markstein correction step with the
// exact remainder. Equal to the IEEE quotient for EVERY double that is a finite float (all 2^32 patterns are swept by
// tests/kat/math_sweep.cpp); the path only ever divides float radians * 180.0f (converted to double) by M_PI.
// the same for v >= +0 (never -0): what every call site of the path passes — float radians of acosf / asinf of a
// non-negative argument, or acosf of anything, times 180.0f — so the sign-of-zero test can go (+0 comes out as +0)
URF_HD double div_pi_nonneg(double v) {
  const double inv = 0.31830988618379069122;                // RN(1 / M_PI) = 0x1.45f306dc9c883p-2
  const double q = URF_DMUL(v, inv);
  const double r = URF_DFMA(-URF_PI_D, q, v);
  return URF_DFMA(r, inv, q);
}

URF_HD double div_pi(double v) {
  if (v == 0.0) return v;                                   // keeps the sign of zero
  const double inv = 0.31830988618379069122;                // RN(1 / M_PI) = 0x1.45f306dc9c883p-2
  const double q = URF_DMUL(v, inv);
  const double r = URF_DFMA(-URF_PI_D, q, v);
  return URF_DFMA(r, inv, q);
}

// ---- asinf: glibc 2.39 sysdeps/ieee754/flt-32/e_asinf.c (Moshier single-precision polynomial) -------------------
URF_HD float asinf_glibc(float x) {
  const float pio2_hi = 1.57079637050628662109375f, pio2_lo = -4.37113900018624283e-8f,
              pio4_hi = 0.785398185253143310546875f;
  const float p0 = 1.666675248e-1f, p1 = 7.495297643e-2f, p2 = 4.547037598e-2f, p3 = 2.417951451e-2f,
              p4 = 4.216630880e-2f;
  int32_t hx = URF_F2I(x);
  int32_t ix = hx & 0x7fffffff;
  if (ix == 0x3f800000) return URF_FADD(URF_FMUL(x, pio2_hi), URF_FMUL(x, pio2_lo));
  if (ix > 0x3f800000) return URF_FDIV(URF_FSUB(x, x), URF_FSUB(x, x));   // NaN
  if (ix < 0x3f000000) {                                                     // |x| < 0.5
    if (ix < 0x32000000) return x;                                           // |x| < 2^-27
    float t = URF_FMUL(x, x);
    float w = URF_FMUL(t, URF_FADD(p0, URF_FMUL(t, URF_FADD(p1, URF_FMUL(t, URF_FADD(p2, URF_FMUL(t, URF_FADD(p3, URF_FMUL(t, p4)))))))));
    return URF_FADD(x, URF_FMUL(x, w));
  }
  float w = URF_FSUB(1.0f, URF_FABS(x));
  float t = URF_FMUL(w, 0.5f);
  float p = URF_FMUL(t, URF_FADD(p0, URF_FMUL(t, URF_FADD(p1, URF_FMUL(t, URF_FADD(p2, URF_FMUL(t, URF_FADD(p3, URF_FMUL(t, p4)))))))));
  float s = URF_FSQRT(t);
  if (ix >= 0x3F79999A) {                                                    // |x| > 0.975
    t = URF_FSUB(pio2_hi, URF_FSUB(URF_FMUL(2.0f, URF_FADD(s, URF_FMUL(s, p))), pio2_lo));
  } else {
    w = URF_I2F(URF_F2I(s) & 0xfffff000);
    float c = URF_FDIV(URF_FSUB(t, URF_FMUL(w, w)), URF_FADD(s, w));
    float r = p;
    p = URF_FSUB(URF_FMUL(URF_FMUL(2.0f, s), r), URF_FSUB(pio2_lo, URF_FMUL(2.0f, c)));
    float q = URF_FSUB(pio4_hi, URF_FMUL(2.0f, w));
    t = URF_FSUB(pio4_hi, URF_FSUB(p, q));
  }
  return hx > 0 ? t : -t;
}

// ---- acosf: glibc 2.39 sysdeps/ieee754/flt-32/e_acosf.c ----------------------------------------------------------
URF_HD float acosf_glibc(float x) {
  const float pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
  const float pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f,
              pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f;
  const float qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f, qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
  int32_t hx = URF_F2I(x);
  int32_t ix = hx & 0x7fffffff;
  if (ix == 0x3f800000) {
    if (hx > 0) return 0.0f;
    return URF_FADD(pi, URF_FMUL(2.0f, pio2_lo));
  }
  if (ix > 0x3f800000) return URF_FDIV(URF_FSUB(x, x), URF_FSUB(x, x));   // NaN
  if (ix < 0x3f000000) {                                                     // |x| < 0.5
    if (ix <= 0x23000000) return URF_FADD(pio2_hi, pio2_lo);
    float z = URF_FMUL(x, x);
    float p = URF_FMUL(z, URF_FADD(pS0, URF_FMUL(z, URF_FADD(pS1, URF_FMUL(z, URF_FADD(pS2, URF_FMUL(z, URF_FADD(pS3, URF_FMUL(z, URF_FADD(pS4, URF_FMUL(z, pS5)))))))))));
    float q = URF_FADD(1.0f, URF_FMUL(z, URF_FADD(qS1, URF_FMUL(z, URF_FADD(qS2, URF_FMUL(z, URF_FADD(qS3, URF_FMUL(z, qS4))))))));
    float r = URF_FDIV(p, q);
    return URF_FSUB(pio2_hi, URF_FSUB(x, URF_FSUB(pio2_lo, URF_FMUL(x, r))));
  }
  if (hx < 0) {                                                              // x < -0.5
    float z = URF_FMUL(URF_FADD(1.0f, x), 0.5f);
    float p = URF_FMUL(z, URF_FADD(pS0, URF_FMUL(z, URF_FADD(pS1, URF_FMUL(z, URF_FADD(pS2, URF_FMUL(z, URF_FADD(pS3, URF_FMUL(z, URF_FADD(pS4, URF_FMUL(z, pS5)))))))))));
    float q = URF_FADD(1.0f, URF_FMUL(z, URF_FADD(qS1, URF_FMUL(z, URF_FADD(qS2, URF_FMUL(z, URF_FADD(qS3, URF_FMUL(z, qS4))))))));
    float s = URF_FSQRT(z);
    float r = URF_FDIV(p, q);
    float w = URF_FSUB(URF_FMUL(r, s), pio2_lo);
    return URF_FSUB(pi, URF_FMUL(2.0f, URF_FADD(s, w)));
  }
  {                                                                          // x > 0.5
    float z = URF_FMUL(URF_FSUB(1.0f, x), 0.5f);
    float s = URF_FSQRT(z);
    float df = URF_I2F(URF_F2I(s) & 0xfffff000);
    float c = URF_FDIV(URF_FSUB(z, URF_FMUL(df, df)), URF_FADD(s, df));
    float p = URF_FMUL(z, URF_FADD(pS0, URF_FMUL(z, URF_FADD(pS1, URF_FMUL(z, URF_FADD(pS2, URF_FMUL(z, URF_FADD(pS3, URF_FMUL(z, URF_FADD(pS4, URF_FMUL(z, pS5)))))))))));
    float q = URF_FADD(1.0f, URF_FMUL(z, URF_FADD(qS1, URF_FMUL(z, URF_FADD(qS2, URF_FMUL(z, URF_FADD(qS3, URF_FMUL(z, qS4))))))));
    float r = URF_FDIV(p, q);
    float w = URF_FADD(URF_FMUL(r, s), c);
    return URF_FMUL(2.0f, URF_FADD(df, w));
  }
}

// ---- atanf: glibc 2.39 sysdeps/ieee754/flt-32/s_atanf.c ----------------------------------------------------------
URF_HD float atanf_glibc(float x) {
  const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  const float aT0 = 3.3333334327e-01f, aT1 = -2.0000000298e-01f, aT2 = 1.4285714924e-01f, aT3 = -1.1111110449e-01f,
              aT4 = 9.0908870101e-02f, aT5 = -7.6918758452e-02f, aT6 = 6.6610731184e-02f, aT7 = -5.8335702866e-02f,
              aT8 = 4.9768779427e-02f, aT9 = -3.6531571299e-02f, aT10 = 1.6285819933e-02f;
  int32_t hx = URF_F2I(x);
  int32_t ix = hx & 0x7fffffff;
  int id;
  if (ix >= 0x4c000000) {                                                    // |x| >= 2^25
    if (ix > 0x7f800000) return URF_FADD(x, x);
    float r = URF_FADD(atanhi[3], atanlo[3]);
    return hx > 0 ? r : -r;
  }
  if (ix < 0x3ee00000) {                                                     // |x| < 0.4375
    if (ix < 0x31000000) return x;                                           // |x| < 2^-29
    id = -1;
  } else {
    x = URF_FABS(x);
    if (ix < 0x3f980000) {                                                   // |x| < 1.1875
      if (ix < 0x3f300000) { id = 0; x = URF_FDIV(URF_FSUB(URF_FMUL(2.0f, x), 1.0f), URF_FADD(2.0f, x)); }
      else                 { id = 1; x = URF_FDIV(URF_FSUB(x, 1.0f), URF_FADD(x, 1.0f)); }
    } else {
      if (ix < 0x401c0000) { id = 2; x = URF_FDIV(URF_FSUB(x, 1.5f), URF_FADD(1.0f, URF_FMUL(1.5f, x))); }
      else                 { id = 3; x = URF_FDIV(-1.0f, x); }
    }
  }
  float z = URF_FMUL(x, x);
  float w = URF_FMUL(z, z);
  float s1 = URF_FMUL(z, URF_FADD(aT0, URF_FMUL(w, URF_FADD(aT2, URF_FMUL(w, URF_FADD(aT4, URF_FMUL(w, URF_FADD(aT6, URF_FMUL(w, URF_FADD(aT8, URF_FMUL(w, aT10)))))))))));
  float s2 = URF_FMUL(w, URF_FADD(aT1, URF_FMUL(w, URF_FADD(aT3, URF_FMUL(w, URF_FADD(aT5, URF_FMUL(w, URF_FADD(aT7, URF_FMUL(w, aT9)))))))));
  if (id < 0) return URF_FSUB(x, URF_FMUL(x, URF_FADD(s1, s2)));
  float hi = id == 0 ? atanhi[0] : id == 1 ? atanhi[1] : id == 2 ? atanhi[2] : atanhi[3];
  float lo = id == 0 ? atanlo[0] : id == 1 ? atanlo[1] : id == 2 ? atanlo[2] : atanlo[3];
  z = URF_FSUB(hi, URF_FSUB(URF_FSUB(URF_FMUL(x, URF_FADD(s1, s2)), lo), x));
  return hx < 0 ? -z : z;
}

// ---- atan2f: glibc 2.39 sysdeps/ieee754/flt-32/e_atan2f.c --------------------------------------------------------
URF_HD float atan2f_glibc(float y, float x) {
  const float tiny = 1.0e-30f, pi_o_4 = 7.8539818525e-01f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f,
              pi_lo = -8.7422776573e-08f;
  int32_t hx = URF_F2I(x), hy = URF_F2I(y);
  int32_t ix = hx & 0x7fffffff, iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return URF_FADD(x, y);
  if (hx == 0x3f800000) return atanf_glibc(y);
  int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    switch (m) {
      case 0: case 1: return y;
      case 2: return URF_FADD(pi, tiny);
      default: return URF_FSUB(-pi, tiny);
    }
  }
  if (ix == 0) return hy < 0 ? URF_FSUB(-pi_o_2, tiny) : URF_FADD(pi_o_2, tiny);
  if (ix == 0x7f800000) {
    if (iy == 0x7f800000) {
      switch (m) {
        case 0: return URF_FADD(pi_o_4, tiny);
        case 1: return URF_FSUB(-pi_o_4, tiny);
        case 2: return URF_FADD(URF_FMUL(3.0f, pi_o_4), tiny);
        default: return URF_FSUB(URF_FMUL(-3.0f, pi_o_4), tiny);
      }
    } else {
      switch (m) {
        case 0: return 0.0f;
        case 1: return -0.0f;
        case 2: return URF_FADD(pi, tiny);
        default: return URF_FSUB(-pi, tiny);
      }
    }
  }
  if (iy == 0x7f800000) return hy < 0 ? URF_FSUB(-pi_o_2, tiny) : URF_FADD(pi_o_2, tiny);
  int k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = URF_FADD(pi_o_2, URF_FMUL(0.5f, pi_lo));
  else if (hx < 0 && k < -60) z = 0.0f;
  else z = atanf_glibc(URF_FABS(URF_FDIV(y, x)));
  switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return URF_FSUB(pi, URF_FSUB(z, pi_lo));
    default: return URF_FSUB(URF_FSUB(z, pi_lo), pi);
  }
}

}  // namespace urfm
