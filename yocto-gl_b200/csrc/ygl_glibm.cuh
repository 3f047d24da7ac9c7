// ygl_glibm.cuh — device restatement of the float libm routines the reference calls (glibc 2.39,
// x86-64): sinf, cosf (sysdeps/ieee754/flt-32/s_sinf.c, s_cosf.c, sincosf.h), expf (e_expf.c),
// logf (e_logf.c) — the double-precision table/polynomial algorithms glibc adopted from ARM's
// optimized routines — and atanf (s_atanf.c, the fdlibm float kernel). Arguments outside the
// ranges the path tracer produces fall back to "fp64 evaluation rounded once" (ygl_math.cuh).
//
// The reference calls these through libm at run time; matching their last bit removes the only
// known source of per-pixel differences against the unmodified reference.
// Two contraction variants exist because x86-64 glibc selects FMA builds of sinf/cosf/expf/logf at
// run time (ifunc) on FMA-capable CPUs: FMA = true fuses every `a + b*c` exactly like GCC does
// for those builds. Which variant matches the GPU box's host libm is decided by measurement
// (tools/libm_check.cu, exhaustive over all float inputs) and recorded in ygl_math.cuh.
#pragma once

#include "ygl_glibm_tables.h"

#ifdef __CUDACC__
#define YGL_GLIBM_FN __device__ __forceinline__
#else
#define YGL_GLIBM_FN inline  // host instantiation: tools/libm_check_host.cpp (exhaustive check against glibc)
#endif

namespace ygl {
namespace glibm {

template <bool FMA>
YGL_GLIBM_FN double mad(double a, double b, double c) {  // a*b + c
#ifdef __CUDA_ARCH__
  return FMA ? fma(a, b, c) : __dadd_rn(__dmul_rn(a, b), c);
#else
  volatile double prod = a * b;  // keep the product rounded even if the host compiler contracts
  return FMA ? fma(a, b, c) : prod + c;
#endif
}
template <bool FMA>
YGL_GLIBM_FN float madf(float a, float b, float c) {
#ifdef __CUDA_ARCH__
  return FMA ? fmaf(a, b, c) : __fadd_rn(__fmul_rn(a, b), c);
#else
  volatile float prod = a * b;
  return FMA ? fmaf(a, b, c) : prod + c;
#endif
}
YGL_GLIBM_FN unsigned asuint(float f) {
#ifdef __CUDA_ARCH__
  return __float_as_uint(f);
#else
  unsigned u;
  memcpy(&u, &f, 4);
  return u;
#endif
}
YGL_GLIBM_FN float asfloat(unsigned u) {
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  float f;
  memcpy(&f, &u, 4);
  return f;
#endif
}
YGL_GLIBM_FN unsigned long long asuint64(double f) {
#ifdef __CUDA_ARCH__
  return (unsigned long long)__double_as_longlong(f);
#else
  unsigned long long u;
  memcpy(&u, &f, 8);
  return u;
#endif
}
YGL_GLIBM_FN double asdouble(unsigned long long u) {
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)u);
#else
  double f;
  memcpy(&f, &u, 8);
  return f;
#endif
}
YGL_GLIBM_FN unsigned abstop12(float x) { return (asuint(x) >> 20) & 0x7ff; }

// sincosf.h: sinf_poly. tab = {sign[4], hpi_inv, hpi, c0, c1, s1, c2, s2, c3, s3, c4}
template <bool FMA>
YGL_GLIBM_FN double sincos_poly(double x, double x2, const double* p, int n) {
  if ((n & 1) == 0) {
    double x3 = x * x2;
    double s1 = mad<FMA>(x2, p[12], p[10]);  // s2 + x2*s3
    double x7 = x3 * x2;
    double s  = mad<FMA>(x3, p[8], x);  // x + x3*s1
    return mad<FMA>(x7, s1, s);         // s + x7*s1
  } else {
    double x4 = x2 * x2;
    double c2 = mad<FMA>(x2, p[13], p[11]);  // c3 + x2*c4
    double c1 = mad<FMA>(x2, p[7], p[6]);    // c0 + x2*c1
    double x6 = x4 * x2;
    double c  = mad<FMA>(x4, p[9], c1);  // c1 + x4*c2
    return mad<FMA>(x6, c2, c);          // c + x6*c2
  }
}
// reduce_fast: x - n*hpi with n = round(x * 2/pi)
template <bool FMA>
YGL_GLIBM_FN double reduce_fast(double x, int* np) {
  double r = x * kSinCosTab0[4];
  int    n = ((int)r + 0x800000) >> 24;
  *np      = n;
  return mad<FMA>(-(double)n, kSinCosTab0[5], x);
}
// returns false when |y| >= 120 (or non-finite): caller falls back
template <bool FMA>
YGL_GLIBM_FN bool sinf_(float y, float* out) {
  double x = y;
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
    double s = x * x;
    if (abstop12(y) < abstop12(0x1p-12f)) {
      *out = y;
      return true;
    }
    *out = (float)sincos_poly<FMA>(x, s, kSinCosTab0, 0);
    return true;
  } else if (abstop12(y) < abstop12(120.0f)) {
    int n;
    x               = reduce_fast<FMA>(x, &n);
    double        s = kSinCosTab0[n & 3];
    const double* p = (n & 2) ? kSinCosTab1 : kSinCosTab0;
    *out            = (float)sincos_poly<FMA>(x * s, x * x, p, n);
    return true;
  }
  return false;
}
template <bool FMA>
YGL_GLIBM_FN bool cosf_(float y, float* out) {
  double x = y;
  if (abstop12(y) < abstop12(0x1.921FB6p-1f)) {
    double x2 = x * x;
    if (abstop12(y) < abstop12(0x1p-12f)) {
      *out = 1.0f;
      return true;
    }
    *out = (float)sincos_poly<FMA>(x, x2, kSinCosTab0, 1);
    return true;
  } else if (abstop12(y) < abstop12(120.0f)) {
    int n;
    x               = reduce_fast<FMA>(x, &n);
    double        s = kSinCosTab0[n & 3];
    const double* p = (n & 2) ? kSinCosTab1 : kSinCosTab0;
    *out            = (float)sincos_poly<FMA>(x * s, x * x, p, n ^ 1);
    return true;
  }
  return false;
}

// e_expf.c. kExp2fRest = {shift_scaled, poly[3], shift, invln2_scaled, poly_scaled[3]}
template <bool FMA>
YGL_GLIBM_FN bool expf_(float x, float* out) {
  if (abstop12(x) >= abstop12(88.0f)) {
    // |x| >= 88 or nan: glibc's special cases (__math_oflowf / __math_uflowf / __math_may_uflowf results)
    if (asuint(x) == 0xff800000u) {
      *out = 0.0f;
      return true;
    }
    if (abstop12(x) >= abstop12(asfloat(0x7f800000u))) return false;  // +inf, nan: fallback
    if (x > 0x1.62e42ep6f) {
      *out = asfloat(0x7f800000u);
      return true;
    }
    if (x < -0x1.9fe368p6f) {
      *out = 0.0f;
      return true;
    }
    if (x < -0x1.9d1d9ep6f) {
      *out = asfloat(1u);  // 0x1.4p-75f * 0x1.4p-75f rounds to the smallest subnormal
      return true;
    }
  }
  double             xd = x;
  double             z  = kExp2fRest[5] * xd;
  double             kd = z + kExp2fRest[4];
  unsigned long long ki = asuint64(kd);
  kd -= kExp2fRest[4];
  // GCC's FMA build fuses r = z - kd with z = InvLn2N * x (seen on 2 of 2^32 inputs)
  double             r = FMA ? fma(kExp2fRest[5], xd, -kd) : z - kd;
  unsigned long long t = kExp2fTab[ki % 32];
  t += ki << (52 - 5);
  double s  = asdouble(t);
  z         = mad<FMA>(kExp2fRest[6], r, kExp2fRest[7]);
  double r2 = r * r;
  double y  = mad<FMA>(kExp2fRest[8], r, 1.0);
  y         = mad<FMA>(z, r2, y);
  y         = y * s;
  *out      = (float)y;
  return true;
}

// e_logf.c. kLogfRest = {ln2, poly[3]}
template <bool FMA>
YGL_GLIBM_FN bool logf_(float x, float* out) {
  unsigned ix = asuint(x);
  if (ix == 0x3f800000) {
    *out = 0;
    return true;
  }
  if (ix - 0x00800000 >= 0x7f800000 - 0x00800000) {
    if (ix == 0 || ix >= 0x7f800000) return false;  // zero, negative, inf, nan: fallback (exact in fp64)
    ix = asuint(x * 0x1p23f);                         // subnormal x: normalize
    ix -= 23u << 23;
  }
  unsigned tmp = ix - 0x3f330000;
  int      i   = (tmp >> (23 - 4)) % 16;
  int      k   = (int)tmp >> 23;
  unsigned iz  = ix - (tmp & 0x1ffu << 23);
  double invc = kLogfTab[2 * i], logc = kLogfTab[2 * i + 1];
  double z  = (double)asfloat(iz);
  double r  = mad<FMA>(z, invc, -1.0);
  double y0 = mad<FMA>((double)k, kLogfRest[0], logc);
  double r2 = r * r;
  double y  = mad<FMA>(kLogfRest[2], r, kLogfRest[3]);
  y         = mad<FMA>(kLogfRest[1], r2, y);
  y         = mad<FMA>(y, r2, y0 + r);
  *out      = (float)y;
  return true;
}

// s_atanf.c (fdlibm float kernel)
template <bool FMA>
YGL_GLIBM_FN bool atanf_(float x, float* out) {
  const float atanhi[4] = {4.6364760399e-01f, 7.8539812565e-01f, 9.8279368877e-01f, 1.5707962513e+00f};
  const float atanlo[4] = {5.0121582440e-09f, 3.7748947079e-08f, 3.4473217170e-08f, 7.5497894159e-08f};
  const float aT[11]    = {3.3333334327e-01f, -2.0000000298e-01f, 1.4285714924e-01f, -1.1111110449e-01f,
         9.0908870101e-02f, -7.6918758452e-02f, 6.6610731184e-02f, -5.8335702866e-02f, 4.9768779427e-02f,
         -3.6531571299e-02f, 1.6285819933e-02f};
  int hx = (int)asuint(x), ix = hx & 0x7fffffff, id;
  if (ix >= 0x4c000000) return false;  // |x| >= 2^25, inf, nan: fallback
  if (ix < 0x3ee00000) {               // |x| < 0.4375
    if (ix < 0x31000000) {             // |x| < 2^-29
      *out = x;
      return true;
    }
    id = -1;
  } else {
    x = fabsf(x);
    if (ix < 0x3f980000) {    // |x| < 1.1875
      if (ix < 0x3f300000) {  // 7/16 <= |x| < 11/16
        id = 0;
        x  = madf<FMA>(2.0f, x, -1.0f) / (2.0f + x);
      } else {  // 11/16 <= |x| < 19/16
        id = 1;
        x  = (x - 1.0f) / (x + 1.0f);
      }
    } else {
      if (ix < 0x401c0000) {  // |x| < 2.4375
        id = 2;
        x  = (x - 1.5f) / madf<FMA>(1.5f, x, 1.0f);
      } else {  // 2.4375 <= |x| < 2^25
        id = 3;
        x  = -1.0f / x;
      }
    }
  }
  float z = x * x, w = z * z;
  float t1 = madf<FMA>(w, aT[10], aT[8]);
  t1       = madf<FMA>(w, t1, aT[6]);
  t1       = madf<FMA>(w, t1, aT[4]);
  t1       = madf<FMA>(w, t1, aT[2]);
  t1       = madf<FMA>(w, t1, aT[0]);
  float s1 = z * t1;
  float t2 = madf<FMA>(w, aT[9], aT[7]);
  t2       = madf<FMA>(w, t2, aT[5]);
  t2       = madf<FMA>(w, t2, aT[3]);
  t2       = madf<FMA>(w, t2, aT[1]);
  float s2 = w * t2;
  if (id < 0) {
    *out = madf<FMA>(-x, s1 + s2, x);  // x - x*(s1+s2)
    return true;
  }
  float zz = atanhi[id] - ((madf<FMA>(x, s1 + s2, -atanlo[id])) - x);
  *out     = (hx < 0) ? -zz : zz;
  return true;
}

// e_powf.c (log2_inline + exp2_inline) for normal x (negative x only with an integer y: the result
// takes the sign (-1)^y, as glibc's sign_bias does) and finite non-zero y with |y*log2|x|| < 126;
// everything else (zero/subnormal/inf/nan operands, negative x with fractional y, overflow and
// underflow handling) falls back. kExp2fRest[0] = shift_scaled, [1..3] = poly.
template <bool FMA>
YGL_GLIBM_FN bool powf_(float x, float y, float* out) {
  unsigned ix = asuint(x), iy = asuint(y);
  if (2 * iy - 1 >= 2u * 0x7f800000 - 1) return false;  // zeroinfnan(iy)
  if (ix & 0x80000000u) {
    // checkint(iy): 0 = not an integer, 1 = odd, 2 = even
    int e = (int)(iy >> 23 & 0xff), yint;
    if (e < 0x7f) yint = 0;
    else if (e > 0x7f + 23) yint = 2;
    else if (iy & ((1u << (0x7f + 23 - e)) - 1)) yint = 0;
    else yint = (iy & (1u << (0x7f + 23 - e))) ? 1 : 2;
    if (yint == 0) return false;
    float r;
    if (!powf_<FMA>(asfloat(ix & 0x7fffffffu), y, &r)) return false;
    *out = yint == 1 ? -r : r;
    return true;
  }
  if (ix - 0x00800000 >= 0x7f800000 - 0x00800000) {
    if (ix == 0 || ix >= 0x7f800000) return false;  // zero, inf, nan: fallback (exact in fp64)
    ix = asuint(x * 0x1p23f);                         // subnormal x: normalize
    ix &= 0x7fffffff;
    ix -= 23u << 23;
  }
  unsigned tmp = ix - 0x3f330000;
  int      i   = (tmp >> (23 - 4)) % 16;
  unsigned top = tmp & 0xff800000;
  unsigned iz  = ix - top;
  int      k   = (int)top >> 23;
  double invc = kPowfLog2Tab[2 * i], logc = kPowfLog2Tab[2 * i + 1];
  double z  = (double)asfloat(iz);
  double r  = mad<FMA>(z, invc, -1.0);
  double y0 = logc + (double)k;
  double r2 = r * r;
  double yy = mad<FMA>(kPowfLog2Poly[0], r, kPowfLog2Poly[1]);
  double p  = mad<FMA>(kPowfLog2Poly[2], r, kPowfLog2Poly[3]);
  double r4 = r2 * r2;
  double q  = mad<FMA>(kPowfLog2Poly[4], r, y0);
  q         = mad<FMA>(p, r2, q);
  double logx  = mad<FMA>(yy, r4, q);
  double ylogx = (double)y * logx;
  if ((asuint64(ylogx) >> 47 & 0xffff) >= asuint64(126.0) >> 47) {
    // |y*log2(x)| >= 126 (or nan): glibc's overflow / underflow results
    if (!(ylogx == ylogx)) return false;
    if (ylogx > 0x1.fffffffd1d571p+6) {
      *out = asfloat(0x7f800000u);
      return true;
    }
    if (ylogx <= -150.0) {
      *out = 0.0f;
      return true;
    }
    if (ylogx < -149.0) {
      *out = asfloat(1u);
      return true;
    }
  }
  double             kd = ylogx + kExp2fRest[0];
  unsigned long long ki = asuint64(kd);
  kd -= kExp2fRest[0];
  double             rr = ylogx - kd;
  unsigned long long t  = kExp2fTab[ki % 32];
  t += ki << (52 - 5);
  double s   = asdouble(t);
  double zz  = mad<FMA>(kExp2fRest[1], rr, kExp2fRest[2]);
  double rr2 = rr * rr;
  double res = mad<FMA>(kExp2fRest[3], rr, 1.0);
  res        = mad<FMA>(zz, rr2, res);
  res        = res * s;
  *out       = (float)res;
  return true;
}

// e_acosf.c (fdlibm float kernel, the 6/4 rational form). VARIANT 1 = msun's shorter 3/1 form.
template <bool FMA>
YGL_GLIBM_FN bool acosf_(float x, float* out) {
  const float one = 1.0f, pi = 3.1415925026e+00f, pio2_hi = 1.5707962513e+00f, pio2_lo = 7.5497894159e-08f;
  const float pS0 = 1.6666667163e-01f, pS1 = -3.2556581497e-01f, pS2 = 2.0121252537e-01f, pS3 = -4.0055535734e-02f,
              pS4 = 7.9153501429e-04f, pS5 = 3.4793309169e-05f, qS1 = -2.4033949375e+00f, qS2 = 2.0209457874e+00f,
              qS3 = -6.8828397989e-01f, qS4 = 7.7038154006e-02f;
  int hx = (int)asuint(x), ix = hx & 0x7fffffff;
  if (ix == 0x3f800000) {
    *out = hx > 0 ? 0.0f : pi + 2.0f * pio2_lo;
    return true;
  }
  if (ix > 0x3f800000) return false;  // |x| > 1 or nan: fallback
  auto P = [&](float z) {
    float t = madf<FMA>(z, pS5, pS4);
    t       = madf<FMA>(z, t, pS3);
    t       = madf<FMA>(z, t, pS2);
    t       = madf<FMA>(z, t, pS1);
    t       = madf<FMA>(z, t, pS0);
    return z * t;
  };
  auto Q = [&](float z) {
    float t = madf<FMA>(z, qS4, qS3);
    t       = madf<FMA>(z, t, qS2);
    t       = madf<FMA>(z, t, qS1);
    return madf<FMA>(z, t, one);
  };
  if (ix < 0x3f000000) {  // |x| < 0.5
    if (ix <= 0x23000000) {
      *out = pio2_hi + pio2_lo;
      return true;
    }
    float z = x * x, r = P(z) / Q(z);
    *out = pio2_hi - (x - madf<FMA>(-x, r, pio2_lo));  // pio2_hi - (x - (pio2_lo - x*r))
    return true;
  } else if (hx < 0) {  // x < -0.5
    float z = (one + x) * 0.5f, p = P(z), q = Q(z), s = sqrtf(z), r = p / q;
    float w = madf<FMA>(r, s, -pio2_lo);
    *out    = pi - 2.0f * (s + w);
    return true;
  } else {  // x > 0.5
    float z = (one - x) * 0.5f, s = sqrtf(z);
    float df = asfloat(asuint(s) & 0xfffff000u);
    float c  = madf<FMA>(-df, df, z) / (s + df);  // (z - df*df)/(s + df)
    float p = P(z), q = Q(z), r = p / q;
    float w = madf<FMA>(r, s, c);
    *out    = 2.0f * (df + w);
    return true;
  }
}

// e_atan2f.c (fdlibm float kernel) on top of atanf_
template <bool FMA>
YGL_GLIBM_FN bool atan2f_(float y, float x, float* out) {
  const float tiny = 1.0e-30f, pi_o_2 = 1.5707963705e+00f, pi = 3.1415927410e+00f, pi_lo = -8.7422776573e-08f;
  int hx = (int)asuint(x), ix = hx & 0x7fffffff, hy = (int)asuint(y), iy = hy & 0x7fffffff;
  if (ix > 0x7f800000 || iy > 0x7f800000) return false;  // nan
  if (hx == 0x3f800000) return atanf_<FMA>(y, out);      // x = 1
  int m = ((hy >> 31) & 1) | ((hx >> 30) & 2);
  if (iy == 0) {
    *out = (m == 0 || m == 1) ? y : (m == 2 ? pi + tiny : -pi - tiny);
    return true;
  }
  if (ix == 0) {
    *out = (hy < 0) ? -pi_o_2 - tiny : pi_o_2 + tiny;
    return true;
  }
  if (ix == 0x7f800000 || iy == 0x7f800000) return false;  // infinities: fallback
  int   k = (iy - ix) >> 23;
  float z;
  if (k > 60) z = pi_o_2 + 0.5f * pi_lo;
  else if (hx < 0 && k < -60) z = 0.0f;
  else if (!atanf_<FMA>(fabsf(y / x), &z)) return false;
  switch (m) {
    case 0: *out = z; break;
    case 1: *out = asfloat(asuint(z) ^ 0x80000000u); break;
    case 2: *out = pi - (z - pi_lo); break;
    default: *out = (z - pi_lo) - pi; break;
  }
  return true;
}

}  // namespace glibm
}  // namespace ygl
