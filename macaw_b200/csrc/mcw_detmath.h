// mcw_detmath.h -- deterministic FP64 elementary functions (log, sin, cos) for host and device.
//
// Why this exists: the reference calls glibc's std::log / std::sin / std::cos (e.g. MaCaw
// CollisionKernel.C:150, openmc/src/random_dist.cpp:15-34, openmc/src/math_functions.cpp:644-672).
// CUDA's libdevice versions of those functions round differently from glibc in the last ulp for a
// fraction of arguments, which would make a GPU history diverge from a CPU history after a few
// hundred random numbers.  Both the CUDA kernels and the CPU oracle therefore evaluate these three
// functions with the SAME sequence of IEEE-754 operations (+,-,*,/ and integer bit moves only,
// compiled with FMA contraction disabled on both sides: nvcc -fmad=false, gcc -ffp-contract=off),
// which makes every history bit-reproducible across CPU and GPU.  sqrt, division and int<->double
// conversions are correctly rounded on both sides already and are used directly.
//
// Algorithms: the classic argument-reduction + minimax-polynomial schemes (log: x = 2^k(1+f),
// s = f/(2+f), 7-term even polynomial; sin/cos: Cody-Waite reduction by pi/2 in three pieces and
// the standard degree-13/14 kernels).  Accuracy < 1 ulp on the argument ranges the transport path
// uses (log: positive normal numbers; sin/cos: |x| < ~1e5).  tests/test_detmath.py checks the ulp
// distance to glibc.  The oracle can also be built against glibc (-DMCW_ORACLE_LIBM) to show that
// swapping the math library only moves results by ulps.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define MCW_HD __host__ __device__ __forceinline__
#else
#define MCW_HD static inline
#endif

MCW_HD uint64_t mcw_d2bits(double x)
{
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u;
  memcpy(&u, &x, 8);
  return u;
#endif
}

MCW_HD double mcw_bits2d(uint64_t u)
{
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x;
  memcpy(&x, &u, 8);
  return x;
#endif
}

// Natural logarithm, x > 0 and normal (x == 0 returns -inf, like the distance sampling needs when
// prn() returns exactly 0).
MCW_HD double mcw_log(double x)
{
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
               Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
               Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  uint64_t ix = mcw_d2bits(x);
  if ((ix << 1) == 0)
    return mcw_bits2d(0xFFF0000000000000ull); // -inf
  int k = 0;
  if ((ix >> 52) == 0)
  { // subnormal: scale up by 2^54
    x = x * 18014398509481984.0;
    ix = mcw_d2bits(x);
    k = -54;
  }
  int hx = (int)(ix >> 32);
  k += (hx >> 20) - 1023;
  hx &= 0x000fffff;
  int i = (hx + 0x95f64) & 0x100000; // 1 when mantissa > sqrt(2)
  // normalise x to [sqrt(2)/2, sqrt(2))
  uint64_t nx = ((uint64_t)(uint32_t)(hx | (i ^ 0x3ff00000)) << 32) | (ix & 0xffffffffull);
  k += (i >> 20);
  double f = mcw_bits2d(nx) - 1.0;
  double dk = (double)k;
  double s = f / (2.0 + f);
  double z = s * s;
  double w = z * z;
  double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  double R = t2 + t1;
  double hfsq = 0.5 * f * f;
  return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}

// sin/cos kernels on |r| <= pi/4 with a tail term rt (r + rt is the reduced argument).
MCW_HD double mcw_ksin(double r, double rt)
{
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
               S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
               S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = r * r;
  double v = z * r;
  double p = S2 + z * (S3 + z * (S4 + z * (S5 + z * S6)));
  return r - ((z * (0.5 * rt - v * p) - rt) - v * S1);
}

MCW_HD double mcw_kcos(double r, double rt)
{
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
               C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
               C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double z = r * r;
  double p = z * (C1 + z * (C2 + z * (C3 + z * (C4 + z * (C5 + z * C6)))));
  double hz = 0.5 * z;
  double w = 1.0 - hz;
  return w + (((1.0 - w) - hz) + (z * p - r * rt));
}

// Reduce x to r + rt with |r| <= ~pi/4, return quadrant n (x ~= n*pi/2 + r).
MCW_HD int mcw_rem_pio2(double x, double* r, double* rt)
{
  const double invpio2 = 6.36619772367581382433e-01;
  const double pio2_1 = 1.57079632673412561417e+00;  // first 33 bits of pi/2
  const double pio2_2 = 6.07710050630396597660e-11;  // second 33 bits
  const double pio2_2t = 2.02226624879595063154e-21; // pi/2 - (pio2_1 + pio2_2)
  double fn = x * invpio2;
  // round to nearest integer without libm
  fn = (fn >= 0.0) ? (double)(long long)(fn + 0.5) : (double)(long long)(fn - 0.5);
  int n = (int)fn;
  // two-stage subtraction (always taken, so the operation sequence has no data-dependent branch)
  double t = x - fn * pio2_1;
  double w = fn * pio2_2;
  double rr = t - w;
  w = fn * pio2_2t - ((t - rr) - w);
  double y0 = rr - w;
  *r = y0;
  *rt = (rr - y0) - w;
  return n;
}

MCW_HD double mcw_sin(double x)
{
  double r, rt;
  int n = mcw_rem_pio2(x, &r, &rt);
  switch (n & 3)
  {
    case 0:
      return mcw_ksin(r, rt);
    case 1:
      return mcw_kcos(r, rt);
    case 2:
      return -mcw_ksin(r, rt);
    default:
      return -mcw_kcos(r, rt);
  }
}

MCW_HD double mcw_cos(double x)
{
  double r, rt;
  int n = mcw_rem_pio2(x, &r, &rt);
  switch (n & 3)
  {
    case 0:
      return mcw_kcos(r, rt);
    case 1:
      return -mcw_ksin(r, rt);
    case 2:
      return -mcw_kcos(r, rt);
    default:
      return mcw_ksin(r, rt);
  }
}

MCW_HD void mcw_sincos(double x, double* s, double* c)
{
  double r, rt;
  int n = mcw_rem_pio2(x, &r, &rt);
  double ks = mcw_ksin(r, rt), kc = mcw_kcos(r, rt);
  switch (n & 3)
  {
    case 0:
      *s = ks;
      *c = kc;
      break;
    case 1:
      *s = kc;
      *c = -ks;
      break;
    case 2:
      *s = -ks;
      *c = -kc;
      break;
    default:
      *s = -kc;
      *c = ks;
      break;
  }
}

// exp(x) for |x| < ~700 (no overflow / underflow handling: the Kalbach-Mann slope parameters it is used for are O(1)).
// Reduction x = k ln2 + r, |r| <= ln2 / 2, then the degree-5 rational form of the classic scheme; same IEEE operations on
// host and device.  < 1 ulp.
MCW_HD double mcw_exp(double x)
{
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10, inv_ln2 = 1.44269504088896338700e+00;
  const double P1 = 1.66666666666666019037e-01, P2 = -2.77777777770155933842e-03, P3 = 6.61375632143793436117e-05,
               P4 = -1.65339022054652515390e-06, P5 = 4.13813679705723846039e-08;
  double ax = x < 0 ? -x : x;
  int k = 0;
  double hi = x, lo = 0.0;
  if (ax > 0.5 * 6.93147180559945286227e-01)
  {
    k = (int)(inv_ln2 * x + (x < 0 ? -0.5 : 0.5));
    hi = x - (double)k * ln2_hi;
    lo = (double)k * ln2_lo;
    x = hi - lo;
  }
  const double t = x * x;
  const double c = x - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))));
  double y;
  if (k == 0)
    return 1.0 - ((x * c) / (c - 2.0) - x);
  y = 1.0 - ((lo - (x * c) / (2.0 - c)) - hi);
  return mcw_bits2d(mcw_d2bits(y) + ((uint64_t)(int64_t)k << 52)); // y * 2^k (y is in [0.5, 2): no exponent overflow here)
}

// sinh through exp; fine for the O(1) arguments of the Kalbach-Mann law (cancellation only matters for |x| << 1)
MCW_HD double mcw_sinh(double x) { return 0.5 * (mcw_exp(x) - mcw_exp(-x)); }
