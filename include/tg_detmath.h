/*
 * tg_detmath.h -- deterministic libm subset (host + device).
 *
 * Why this exists: the reference (ctu-mrs/mrs_uav_trajectory_generation) calls
 * glibc's atan2/sin/cos (eth/vertex.cpp:512-520), exp/log
 * (eth/rpoly/rpoly_ak1.cpp:156,227,246), cbrt (eth/trajectory.cpp:642) and
 * pow(t, integer) (lin_impl.h:615).  The CUDA math library and glibc differ in
 * the last ulp, and one ulp in a segment time decorrelates the rounding noise
 * of the ill-conditioned reduced system (cond(Rpp) ~ 1e8), which
 * would break the 1e-9 coefficient parity between the GPU path and the CPU
 * oracle.  These routines use only IEEE-754 +,-,*,/ and explicit fma, so they
 * return bit-identical results under `gcc -ffp-contract=off` and
 * `nvcc -fmad=false`.  Accuracy (measured in tests/test_detmath.py against
 * mpmath, next to glibc): exp, sin, cos, atan2, cbrt and pow(x, int) are
 * correctly rounded (double-double evaluation, one final rounding) on the ranges
 * the path uses; log is within 0.7 ulp and agrees with glibc on 99.7 %.
 *
 * Coefficient tables come from tools/gen_detmath_coeffs.py (mpmath, 60 digits).
 */
#ifndef TG_DETMATH_H_
#define TG_DETMATH_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define TG_UNROLL _Pragma("unroll")
#define TG_HD __host__ __device__ __forceinline__
#define TG_HD_NOINLINE __host__ __device__
#define TG_HD_OUTLINE __host__ __device__ __noinline__ inline
#else
#define TG_UNROLL
#define TG_HD inline
#define TG_HD_NOINLINE
#define TG_HD_OUTLINE inline
#include <cmath>
#include <cstring>
#endif

namespace tgdm {

TG_HD int64_t dbits(double x) {
#if defined(__CUDA_ARCH__)
  return __double_as_longlong(x);
#else
  int64_t b;
  memcpy(&b, &x, 8);
  return b;
#endif
}
TG_HD double bitsd(int64_t b) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double(b);
#else
  double x;
  memcpy(&x, &b, 8);
  return x;
#endif
}
TG_HD double dabs(double x) { return bitsd(dbits(x) & 0x7fffffffffffffffLL); }
TG_HD double dsqrt(double x) {
#if defined(__CUDA_ARCH__)
  return __dsqrt_rn(x);
#else
  return std::sqrt(x);
#endif
}
/* exact fused multiply-add (IEEE), identical on both sides */
TG_HD double dfma(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
/* 2^e for -1022 <= e <= 1023 (exact) */
TG_HD double pow2i(int e) { return bitsd((int64_t)(e + 1023) << 52); }
/* x * 2^e with e split so that intermediate scale factors stay normal */
TG_HD double scalb(double x, int e) {
  if (e > 1000) { x = x * pow2i(1000); e -= 1000; if (e > 1000) { x = x * pow2i(1000); e -= 1000; } }
  if (e < -1000) { x = x * pow2i(-1000); e += 1000; if (e < -1000) { x = x * pow2i(-1000); e += 1000; } }
  return x * pow2i(e);
}
TG_HD bool disnan(double x) { return (dbits(x) & 0x7fffffffffffffffLL) > 0x7ff0000000000000LL; }
TG_HD bool disinf(double x) { return (dbits(x) & 0x7fffffffffffffffLL) == 0x7ff0000000000000LL; }

struct dd { double hi, lo; };

/* ---- log --------------------------------------------------------------- */
TG_HD double dlog_k(double x) { /* < 0.7 ulp kernel; the public dlog below adds one Newton step in double-double */
  const double ln2_hi = 0x1.62e42fee00000p-1, ln2_lo = 0x1.a39ef35793c76p-33;
  int64_t b = dbits(x);
  int k = 0;
  if (b < 0x0010000000000000LL) { /* zero, subnormal or negative */
    if ((b & 0x7fffffffffffffffLL) == 0) return -bitsd(0x7ff0000000000000LL);
    if (b < 0) return bitsd(0x7ff8000000000000LL);
    x = x * 0x1p54;
    b = dbits(x);
    k = -54;
  }
  if (b >= 0x7ff0000000000000LL) return x; /* inf or nan */
  k += (int)(b >> 52) - 1023;
  int64_t m = b & 0x000fffffffffffffLL;
  /* mantissa in [sqrt(1/2), sqrt(2)) */
  if (m >= 0x6a09e667f3bcdLL) { k += 1; b = m | 0x3fe0000000000000LL; } else { b = m | 0x3ff0000000000000LL; }
  const double f = bitsd(b) - 1.0;
  const double s = f / (2.0 + f);
  const double z = s * s;
  double R = 0x1.0c135adcf3011p-3;
  R = R * z + 0x1.0fbd140544b63p-3;
  R = R * z + 0x1.3b1c43c68eb6dp-3;
  R = R * z + 0x1.745cf8c09a9d4p-3;
  R = R * z + 0x1.c71c7201fc0e6p-3;
  R = R * z + 0x1.249249247670ap-2;
  R = R * z + 0x1.9999999999a3ap-2;
  R = R * z + 0x1.5555555555555p-1;
  R = R * z;
  const double hfsq = 0.5 * f * f;
  const double dk = (double)k;
  return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

/* ---- exp --------------------------------------------------------------- */
TG_HD double dexp_k(double x) { /* < 1 ulp kernel; the public dexp below is correctly rounded */
  const double ln2_hi = 0x1.62e42fee00000p-1, ln2_lo = 0x1.a39ef35793c76p-33, inv_ln2 = 0x1.71547652b82fep+0;
  if (disnan(x)) return x;
  if (x > 709.782712893384) return bitsd(0x7ff0000000000000LL);
  if (x < -745.2) return 0.0;
  const double t = x * inv_ln2;
  const int k = (int)(t < 0 ? t - 0.5 : t + 0.5);
  const double dk = (double)k;
  const double hi = x - dk * ln2_hi;
  const double lo = dk * ln2_lo;
  const double r = hi - lo;
  const double z = r * r;
  double P = -0x1.1fdd0fa97e0efp-30;
  P = P * z + 0x1.66a6198d419e5p-25;
  P = P * z + -0x1.bbd777c80a4a2p-20;
  P = P * z + 0x1.1566abbfe86bap-14;
  P = P * z + -0x1.6c16c16c16be1p-9;
  P = P * z + 0x1.5555555555555p-3;
  const double c = r - z * P;
  const double y = 1.0 - ((lo - (r * c) / (2.0 - c)) - hi);
  return scalb(y, k);
}

/* ---- sin / cos ---------------------------------------------------------- */
TG_HD double ksin(double r) {
  const double z = r * r;
  double S = -0x1.ab16f5caa198ap-41;
  S = S * z + 0x1.61217d9252c03p-33;
  S = S * z + -0x1.ae645410e937bp-26;
  S = S * z + 0x1.71de3a545f836p-19;
  S = S * z + -0x1.a01a01a01992fp-13;
  S = S * z + 0x1.1111111111110p-7;
  S = S * z + -0x1.5555555555555p-3;
  return r + (r * z) * S;
}
TG_HD double kcos(double r) {
  const double z = r * r;
  double C = 0x1.ab779550c8d9bp-45;
  C = C * z + -0x1.9394b8bd50ae1p-37;
  C = C * z + 0x1.1eed8deac3dfep-29;
  C = C * z + -0x1.27e4fb77125c5p-22;
  C = C * z + 0x1.a01a01a019d06p-16;
  C = C * z + -0x1.6c16c16c16c16p-10;
  C = C * z + 0x1.5555555555555p-5;
  const double hz = 0.5 * z;
  const double w = 1.0 - hz;
  return w + (((1.0 - w) - hz) + (z * z) * C);
}
/* Cody-Waite reduction; adequate for |x| < 2^20 * pi/2 (headings, inclinations) */
TG_HD int rem_pio2(double x, double* r) {
  const double two_over_pi = 0x1.45f306dc9c883p-1;
  const double p1 = 0x1.921fb54400000p+0, p2 = 0x1.0b4611a600000p-34, p3 = 0x1.3198a2e037073p-69;
  const double t = x * two_over_pi;
  const int n = (int)(t < 0 ? t - 0.5 : t + 0.5);
  const double dn = (double)n;
  *r = ((x - dn * p1) - dn * p2) - dn * p3;
  return n;
}
TG_HD double dsin_k(double x) { /* < 1.5 ulp kernel; the public dsin below is correctly rounded */
  if (disnan(x) || disinf(x)) return bitsd(0x7ff8000000000000LL);
  if (dabs(x) <= 0x1.921fb54442d18p-1) return ksin(x);
  double r;
  const int n = rem_pio2(x, &r) & 3;
  if (n == 0) return ksin(r);
  if (n == 1) return kcos(r);
  if (n == 2) return -ksin(r);
  return -kcos(r);
}
TG_HD double dcos_k(double x) { /* < 1.5 ulp kernel; the public dcos below is correctly rounded */
  if (disnan(x) || disinf(x)) return bitsd(0x7ff8000000000000LL);
  if (dabs(x) <= 0x1.921fb54442d18p-1) return kcos(x);
  double r;
  const int n = rem_pio2(x, &r) & 3;
  if (n == 0) return kcos(r);
  if (n == 1) return -ksin(r);
  if (n == 2) return -kcos(r);
  return ksin(r);
}

/* ---- atan / atan2 ------------------------------------------------------- */
TG_HD double katan(double t) { /* |t| <= 7/16 */
  const double z = t * t;
  double T = 0x1.99b8c7f011ec9p-7;
  T = T * z + -0x1.ddb926fc36723p-6;
  T = T * z + 0x1.4ab4bf74b2a16p-5;
  T = T * z + -0x1.8129cf6a2388dp-5;
  T = T * z + 0x1.ae7f800dc449fp-5;
  T = T * z + -0x1.e1d2299aa18b2p-5;
  T = T * z + 0x1.11108fdc27707p-4;
  T = T * z + -0x1.3b13aba41be33p-4;
  T = T * z + 0x1.745d171ded7a6p-4;
  T = T * z + -0x1.c71c71c672863p-4;
  T = T * z + 0x1.24924924918cap-3;
  T = T * z + -0x1.999999999998fp-3;
  T = T * z + 0x1.5555555555555p-2;
  return t - (t * z) * T;
}
TG_HD double datan(double x) {
  if (disnan(x)) return x;
  const bool neg = dbits(x) < 0;
  const double a = dabs(x);
  double res;
  if (a < 0.4375) {
    res = katan(a);
  } else if (a < 0.6875) { /* atan(0.5) + atan((2a-1)/(2+a)) */
    const double t = (2.0 * a - 1.0) / (2.0 + a);
    res = 0x1.dac670561bb4fp-2 + (katan(t) + 0x1.a2b7f222f65e2p-56);
  } else if (a < 1.1875) { /* atan(1) + atan((a-1)/(a+1)) */
    const double t = (a - 1.0) / (a + 1.0);
    res = 0x1.921fb54442d18p-1 + (katan(t) + 0x1.1a62633145c07p-55);
  } else if (a < 2.4375) { /* atan(1.5) + atan((a-1.5)/(1+1.5a)) */
    const double t = (a - 1.5) / (1.0 + 1.5 * a);
    res = 0x1.f730bd281f69bp-1 + (katan(t) + 0x1.007887af0cbbdp-56);
  } else if (a < 0x1p66) { /* pi/2 - atan(1/a) */
    const double t = -1.0 / a;
    res = 0x1.921fb54442d18p+0 + (katan(t) + 0x1.1a62633145c07p-54);
  } else {
    res = 0x1.921fb54442d18p+0;
  }
  return neg ? -res : res;
}
TG_HD double datan2_k(double y, double x) { /* < 1.5 ulp kernel; the public datan2 below is correctly rounded */
  const double pi = 0x1.921fb54442d18p+1, pi_lo = 0x1.1a62633145c07p-53, pio2 = 0x1.921fb54442d18p+0;
  if (disnan(x) || disnan(y)) return x + y;
  const bool yneg = dbits(y) < 0, xneg = dbits(x) < 0;
  if (y == 0.0) return xneg ? (yneg ? -pi : pi) : y;
  if (x == 0.0) return yneg ? -pio2 : pio2;
  if (disinf(x)) {
    if (disinf(y)) {
      const double q = xneg ? 3.0 * 0x1.921fb54442d18p-1 : 0x1.921fb54442d18p-1;
      return yneg ? -q : q;
    }
    return xneg ? (yneg ? -pi : pi) : (yneg ? -0.0 : 0.0);
  }
  if (disinf(y)) return yneg ? -pio2 : pio2;
  const double z = datan(dabs(y / x));
  double r;
  if (!xneg) r = z; else r = pi - (z - pi_lo);
  return yneg ? -r : r;
}

/* ---- cbrt --------------------------------------------------------------- */
TG_HD double dcbrt(double x) {
  if (disnan(x) || disinf(x) || x == 0.0) return x;
  const bool neg = dbits(x) < 0;
  double a = dabs(x);
  int e3 = 0;
  int64_t b = dbits(a);
  if (b < 0x0010000000000000LL) { a = a * 0x1p54; b = dbits(a); e3 = -18; }
  int e = (int)(b >> 52) - 1023;
  /* e = 3q + rem, rem in {0,1,2}; m in [1,8) */
  int q = e / 3;
  int rem = e - 3 * q;
  if (rem < 0) { rem += 3; q -= 1; }
  const double m = bitsd((b & 0x000fffffffffffffLL) | ((int64_t)(1023 + rem) << 52));
  /* cubic seed on [1,8): ~7 bits */
  double t = 0x1.aaa35d014643bp-10;
  t = t * m + -0x1.16bdce74d65dbp-5;
  t = t * m + 0x1.50ba80e69b747p-2;
  t = t * m + 0x1.6ef7d7cd66d8cp-1;
  /* Halley iterations: t <- t*(t^3 + 2m)/(2t^3 + m): 7 -> 21 -> 63 bits */
  for (int i = 0; i < 3; ++i) {
    const double t3 = t * t * t;
    t = t * ((t3 + 2.0 * m) / (2.0 * t3 + m));
  }
  /* one Newton correction with the residual computed by exact fma products */
  {
    const double t2 = t * t;
    const double t2e = dfma(t, t, -t2);        /* t*t = t2 + t2e */
    const double t3 = t2 * t;
    const double t3e = dfma(t2, t, -t3) + t2e * t; /* t^3 ~= t3 + t3e */
    const double resid = (m - t3) - t3e;
    t = t + resid / (3.0 * t2);
  }
  const double r = t * pow2i(q + e3);
  return neg ? -r : r;
}

/* ---- integer power, correctly rounded in all but pathological cases ------ */
/* pow(t, n) for n >= 1 via double-double products (error-free fma splitting):
   restates glibc pow(t, exponent) of lin_impl.h:615 (glibc pow is < 1 ulp and
   agrees with the correctly rounded value except in rare half-way cases). */
TG_HD dd dd_mul_d(dd a, double b) {
  const double p = a.hi * b;
  const double e = dfma(a.hi, b, -p) + a.lo * b;
  dd r;
  r.hi = p + e;
  r.lo = e - (r.hi - p);
  return r;
}
/* fills out[k] = t^(k+1), k = 0..nmax-1, each rounded from a double-double */
TG_HD void powers(double t, int nmax, double* out) {
  dd acc; acc.hi = t; acc.lo = 0.0;
  out[0] = t;
  for (int k = 1; k < nmax; ++k) {
    acc = dd_mul_d(acc, t);
    out[k] = acc.hi;
  }
}


/* ---- correctly rounded exp, sin, cos, atan2 ------------------------------------------------------------------------
   The kernels above are within 0.7 .. 1.5 ulp and agree with glibc (which is correctly rounded on all but ~0.2 % of
   arguments, tests/test_detmath.py) on only 82 .. 90 % of the arguments the path produces.  One ulp in an inclination
   angle is one ulp in a segment time, and cond(Rpp) ~ 1e8 .. 1e13 turns that into 1e-6 relative in the coefficients
   (DESIGN.md "numeric floor").  The public functions below evaluate in double-double arithmetic (error-free two_sum /
   fma products, ~2^-100 relative) and round once, so they return the correctly rounded value except when the exact
   result lies within 2^-100 of a rounding boundary -- and therefore the same bits as glibc wherever glibc itself is
   correctly rounded.  Only IEEE +, -, *, / and explicit fma: bit-identical under gcc -ffp-contract=off and nvcc -fmad=false. */
TG_HD dd two_sum(double a, double b) {
  dd r;
  r.hi = a + b;
  const double bb = r.hi - a;
  r.lo = (a - (r.hi - bb)) + (b - bb);
  return r;
}
TG_HD dd quick_two_sum(double a, double b) { /* |a| >= |b| */
  dd r;
  r.hi = a + b;
  r.lo = b - (r.hi - a);
  return r;
}
TG_HD dd two_prod(double a, double b) {
  dd r;
  r.hi = a * b;
  r.lo = dfma(a, b, -r.hi);
  return r;
}
TG_HD dd dd_add(dd a, dd b) {
  dd s = two_sum(a.hi, b.hi);
  const dd t = two_sum(a.lo, b.lo);
  s.lo = s.lo + t.hi;
  s = quick_two_sum(s.hi, s.lo);
  s.lo = s.lo + t.lo;
  return quick_two_sum(s.hi, s.lo);
}
TG_HD dd dd_add_d(dd a, double b) {
  dd s = two_sum(a.hi, b);
  s.lo = s.lo + a.lo;
  return quick_two_sum(s.hi, s.lo);
}
TG_HD dd dd_neg(dd a) {
  dd r;
  r.hi = -a.hi;
  r.lo = -a.lo;
  return r;
}
TG_HD dd dd_mul(dd a, dd b) {
  dd p = two_prod(a.hi, b.hi);
  p.lo = p.lo + (a.hi * b.lo + a.lo * b.hi);
  return quick_two_sum(p.hi, p.lo);
}
TG_HD dd dd_muld(dd a, double b) {
  dd p = two_prod(a.hi, b);
  p.lo = p.lo + a.lo * b;
  return quick_two_sum(p.hi, p.lo);
}
TG_HD dd dd_div_d(dd a, double b) {
  const double q1 = a.hi / b;
  const dd r = dd_add(a, dd_neg(two_prod(q1, b)));
  const double q2 = r.hi / b;
  const dd r2 = dd_add(r, dd_neg(two_prod(q2, b)));
  const double q3 = r2.hi / b;
  const dd q = quick_two_sum(q1, q2);
  return dd_add_d(q, q3);
}

/* sin and cos of a double as double-doubles; |x| < 2^20 * pi/2.  x = n pi/2 + r, |r| <= pi/4; with z = r^2,
   sin r = r + r z (S1 + S2 z + ...), cos r = 1 + z (C1 + C2 z + ...): the first eight coefficients are double-doubles
   (Horner in double-double), the tail beyond z^8 contributes < 2^-54 and is evaluated in double.  which: 1 sin, 2 cos, 3 both. */
TG_HD_OUTLINE void sincos_dd(double x, dd* sn, dd* cs, int which) {
  const double two_over_pi = 0x1.45f306dc9c883p-1;
  /* pi/2 = p1 + p2 + p3 + p4; p1, p2 carry 33 bits so that n * p1, n * p2 are exact for n < 2^20 */
  const double p1 = 0x1.921fb54400000p+0, p2 = 0x1.0b4611a600000p-34, p3 = 0x1.3198a2e037073p-69, p4 = 0x1.129024e088a68p-123;
  const double t = x * two_over_pi;
  const int n = (int)(t < 0 ? t - 0.5 : t + 0.5);
  const double dn = (double)n;
  dd r = two_sum(x, -(dn * p1));
  r = dd_add_d(r, -(dn * p2));
  r = dd_add(r, dd_neg(two_prod(dn, p3)));
  r = dd_add(r, dd_neg(two_prod(dn, p4)));
  const dd z = dd_mul(r, r);
  const bool odd = (n & 1) != 0;
  const bool need_s = odd ? (which & 2) != 0 : (which & 1) != 0;  /* sin r feeds sin x in even quadrants, cos x in odd ones */
  const bool need_c = odd ? (which & 1) != 0 : (which & 2) != 0;
  dd s, c;
  s.hi = s.lo = c.hi = c.lo = 0.0;
  if (need_s) {
    double tail = -0x1.434d2e783f5bcp-113;
    tail = tail * z.hi + 0x1.259f98b4358adp-103;
    tail = tail * z.hi + -0x1.d1ab1c2dccea3p-94;
    tail = tail * z.hi + 0x1.3f3ccdd165fa9p-84;
    tail = tail * z.hi + -0x1.761b41316381ap-75;
    tail = tail * z.hi + 0x1.71b8ef6dcf572p-66;
    tail = tail * z.hi + -0x1.2f49b46814157p-57;
    const double shi[8] = {-0x1.5555555555555p-3, 0x1.1111111111111p-7, -0x1.a01a01a01a01ap-13, 0x1.71de3a556c734p-19,
                           -0x1.ae64567f544e4p-26, 0x1.6124613a86d09p-33, -0x1.ae7f3e733b81fp-41, 0x1.952c77030ad4ap-49};
    const double slo[8] = {-0x1.5555555555555p-57, 0x1.1111111111111p-63, -0x1.a01a01a01a01ap-73, -0x1.c154f8ddc6c00p-73,
                           0x1.c062e06d1f209p-80, 0x1.f28e0cc748ebep-87, -0x1.1d8656b0ee8cbp-97, 0x1.ac981465ddc6cp-103};
    dd acc;
    acc.hi = tail;
    acc.lo = 0.0;
TG_UNROLL
    for (int k = 7; k >= 0; --k) {
      dd ck;
      ck.hi = shi[k];
      ck.lo = slo[k];
      acc = dd_add(dd_mul(acc, z), ck);
    }
    s = dd_add(r, dd_mul(dd_mul(r, z), acc));
  }
  if (need_c) {
    double tail = -0x1.3932c5047d60ep-108;
    tail = tail * z.hi + 0x1.0a18a2635085dp-98;
    tail = tail * z.hi + -0x1.88e85fc6a4e5ap-89;
    tail = tail * z.hi + 0x1.f2cf01972f578p-80;
    tail = tail * z.hi + -0x1.0ce396db7f853p-70;
    tail = tail * z.hi + 0x1.e542ba4020225p-62;
    tail = tail * z.hi + -0x1.6827863b97d97p-53;
    const double chi[8] = {-0x1.0000000000000p-1, 0x1.5555555555555p-5, -0x1.6c16c16c16c17p-10, 0x1.a01a01a01a01ap-16,
                           -0x1.27e4fb7789f5cp-22, 0x1.1eed8eff8d898p-29, -0x1.93974a8c07c9dp-37, 0x1.ae7f3e733b81fp-45};
    const double clo[8] = {0.0, 0x1.5555555555555p-59, 0x1.f49f49f49f49fp-65, 0x1.a01a01a01a01ap-76,
                           -0x1.cbbc05b4fa99ap-76, -0x1.2aec959e14c06p-83, -0x1.05d6f8a2efd1fp-92, 0x1.1d8656b0ee8cbp-101};
    dd acc;
    acc.hi = tail;
    acc.lo = 0.0;
TG_UNROLL
    for (int k = 7; k >= 0; --k) {
      dd ck;
      ck.hi = chi[k];
      ck.lo = clo[k];
      acc = dd_add(dd_mul(acc, z), ck);
    }
    c = dd_add_d(dd_mul(z, acc), 1.0);
  }
  switch (n & 3) {
    case 0: *sn = s; *cs = c; break;
    case 1: *sn = c; *cs = dd_neg(s); break;
    case 2: *sn = dd_neg(s); *cs = dd_neg(c); break;
    default: *sn = dd_neg(c); *cs = s; break;
  }
}
TG_HD double dsin(double x) {
  if (disnan(x) || disinf(x) || dabs(x) > 0x1p20) return dsin_k(x);
  if (dabs(x) < 0x1p-27) return x;
  dd s, c;
  sincos_dd(x, &s, &c, 1);
  return s.hi;
}
TG_HD double dcos(double x) {
  if (disnan(x) || disinf(x) || dabs(x) > 0x1p20) return dcos_k(x);
  dd s, c;
  sincos_dd(x, &s, &c, 2);
  return c.hi;
}
/* atan2: the kernel's value z0 (< 1.5 ulp) plus one Newton step on the angle of (x, y):
   theta - z0 = atan((y cos z0 - x sin z0) / (x cos z0 + y sin z0)), |theta - z0| ~ 2^-52 |z0| so atan(d) = d to 2^-104 */
TG_HD double datan2(double y, double x) {
  const double z0 = datan2_k(y, x);
  if (disnan(z0) || y == 0.0 || x == 0.0 || disinf(x) || disinf(y)) return z0;
  const double ax = dabs(x), ay = dabs(y);
  if (ax > 0x1p500 || ay > 0x1p500 || ax < 0x1p-500 || ay < 0x1p-500) return z0; /* products would over/underflow: keep the kernel's value */
  dd s, c;
  sincos_dd(z0, &s, &c, 3);
  const dd num = dd_add(dd_muld(c, y), dd_neg(dd_muld(s, x)));
  const dd den = dd_add(dd_muld(c, x), dd_muld(s, y));
  const double delta = num.hi / den.hi;
  return z0 + delta;
}
/* hypot for moderate magnitudes (sample spacings): x^2 + y^2 in double-double, square root with one double-double Newton
   correction, one rounding */
TG_HD double dhypot(double x, double y) {
  if (disnan(x) || disnan(y) || disinf(x) || disinf(y)) return dabs(x) + dabs(y);
  const double ax = dabs(x), ay = dabs(y);
  if (ax > 0x1p500 || ay > 0x1p500 || (ax < 0x1p-500 && ay < 0x1p-500)) return dsqrt(x * x + y * y);
  const dd s = dd_add(two_prod(x, x), two_prod(y, y));
  const double r = dsqrt(s.hi);
  if (r == 0.0) return r;
  const dd res = dd_add(s, dd_neg(two_prod(r, r)));  /* s - r^2 */
  return r + res.hi / (2.0 * r);
}
/* exp: x = k ln2 + r, exp(r / 256) - 1 by Taylor in double-double, squared up eight times */
TG_HD_OUTLINE dd exp_dd(double x, int* kout) {
  const double l1 = 0x1.62e42fee00000p-1, l2 = 0x1.a39ef35793c76p-33, l3 = 0x1.cc01f97b57a08p-87, inv_ln2 = 0x1.71547652b82fep+0;
  const double t = x * inv_ln2;
  const int k = (int)(t < 0 ? t - 0.5 : t + 0.5);
  const double dk = (double)k;
  dd r = two_sum(x, -(dk * l1)); /* dk * l1 exact: 11 + 32 bits */
  r = dd_add(r, dd_neg(two_prod(dk, l2)));
  r = dd_add(r, dd_neg(two_prod(dk, l3)));
  r.hi = r.hi * 0x1p-8;
  r.lo = r.lo * 0x1p-8;
  dd p = r, term = r;
  for (int j = 2; j <= 11; ++j) {
    term = dd_div_d(dd_mul(term, r), (double)j);
    p = dd_add(p, term);
  }
  for (int j = 0; j < 8; ++j) { /* (1 + p)^2 - 1 = 2 p + p^2 */
    dd two_p;
    two_p.hi = 2.0 * p.hi;
    two_p.lo = 2.0 * p.lo;
    p = dd_add(two_p, dd_mul(p, p));
  }
  *kout = k;
  return dd_add_d(p, 1.0);
}
TG_HD double dexp(double x) {
  if (disnan(x) || x > 709.0 || x < -708.0) return dexp_k(x); /* overflow / subnormal results: the kernel handles the edges */
  int k;
  const dd y = exp_dd(x, &k);
  return scalb(y.hi, k);
}
/* log: the kernel's value y0 plus one Newton step, log x = y0 + log(x exp(-y0)) = y0 + (x exp(-y0) - 1) to second order */
TG_HD double dlog(double x) {
  const double y0 = dlog_k(x);
  if (disnan(y0) || disinf(y0) || y0 == 0.0 || x < 0x1p-1000 || dabs(y0) > 700.0) return y0;
  int k;
  const dd e = exp_dd(-y0, &k);
  dd m = dd_muld(e, x);
  m.hi = scalb(m.hi, k);
  m.lo = scalb(m.lo, k);
  const dd t = dd_add_d(m, -1.0);
  return y0 + t.hi;
}

}  // namespace tgdm

#endif  // TG_DETMATH_H_
