// rj_pow(x, y): x^y for finite positive normal x with ~0.50x ulp error.
//
// Why: the convex limiter of an (almost) isentropic flow sits on its entropy
// bound with zero slack, so the last bit of every pow() in specific_entropy /
// Limiter::limit decides limiter coefficients ("pow").  The
// reference uses Agner Fog's VCL pow (source/simd.template.h:196-231, < 1 ulp);
// CUDA's pow() is specified to 2 ulp and is the single largest source of
// GPU-vs-reference deviation.  This routine is both closer to correctly rounded
// and cheaper than the generic CUDA pow (no handling of negative bases, NaN,
// subnormals on the fast path):
//
//   log(x)  = k ln2 + log(c) + log1p(r),  z = x / 2^k in [0.7055, 1.411),
//             r = fma(z, invc, -1) exact (invc has 8 significant bits),
//             (log c) and k ln2 and the first two series terms carried as
//             double-double, tail as a degree-8 Taylor polynomial;
//   x^y     = exp(y log x) with a 128-entry table of 2^(j/128) and a degree-6
//             polynomial on |r| <= ln2/256.
//
// Same structure as the pow of Szabolcs Nagy (ARM optimized routines / glibc
// >= 2.28), re-derived with own tables (gen_pow_tables.py).  Anything outside
// the fast path (x <= 0, inf, nan, subnormal, |y log x| > 700) goes to the
// platform pow.
#pragma once

#include "pow_tables.h"

#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDA_ARCH__)
#define RJ_POW_FN __device__ __forceinline__
#define RJ_POW_FMA(a, b, c) __fma_rn((a), (b), (c))
#define RJ_POW_ADD(a, b) __dadd_rn((a), (b))
#define RJ_POW_MUL(a, b) __dmul_rn((a), (b))
#elif defined(__CUDACC__)
#define RJ_POW_FN __host__ __device__ inline
#define RJ_POW_FMA(a, b, c) std::fma((a), (b), (c))
#define RJ_POW_ADD(a, b) ((a) + (b))
#define RJ_POW_MUL(a, b) ((a) * (b))
#else
#define RJ_POW_FN inline
#define RJ_POW_FMA(a, b, c) std::fma((a), (b), (c))
#define RJ_POW_ADD(a, b) ((a) + (b))
#define RJ_POW_MUL(a, b) ((a) * (b))
#endif

namespace rjb
{
  struct PowLogEntry {
    double invc, logc, logctail;
  };
  struct PowExpEntry {
    double hi, tail;
  };

#if defined(__CUDACC__)
  /* read through the read-only path; 5 KB, resident in L1 */
  __device__ const PowLogEntry rj_pow_log_table[1 << RJ_POW_LOG_TABLE_BITS] = {RJ_POW_LOG_TABLE};
  __device__ const PowExpEntry rj_pow_exp_table[1 << RJ_POW_EXP_TABLE_BITS] = {RJ_POW_EXP_TABLE};
#endif
#if !defined(__CUDA_ARCH__)
  static const PowLogEntry rj_pow_log_table_host[1 << RJ_POW_LOG_TABLE_BITS] = {RJ_POW_LOG_TABLE};
  static const PowExpEntry rj_pow_exp_table_host[1 << RJ_POW_EXP_TABLE_BITS] = {RJ_POW_EXP_TABLE};
#endif

  RJ_POW_FN uint64_t rj_pow_bits(double x)
  {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u;
    std::memcpy(&u, &x, sizeof(u));
    return u;
#endif
  }

  RJ_POW_FN double rj_pow_from_bits(uint64_t u)
  {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x;
    std::memcpy(&x, &u, sizeof(x));
    return x;
#endif
  }

  /* s + e = a + b exactly (Knuth) */
  RJ_POW_FN void rj_two_sum(double a, double b, double &s, double &e)
  {
    s = RJ_POW_ADD(a, b);
    const double bb = RJ_POW_ADD(s, -a);
    e = RJ_POW_ADD(RJ_POW_ADD(a, -RJ_POW_ADD(s, -bb)), RJ_POW_ADD(b, -bb));
  }

  RJ_POW_FN double rj_pow(const double x, const double y)
  {
    const uint64_t ix = rj_pow_bits(x);
    /* positive normal finite x only: biased exponent in [1, 0x7fe] and sign 0 */
    if (ix - 0x0010000000000000ULL >= 0x7ff0000000000000ULL - 0x0010000000000000ULL)
      return pow(x, y);

    /* ---- log(x) = hi + lo ---- */
    const uint64_t tmp = ix - RJ_POW_OFF;
    const int i = (int)((tmp >> (52 - RJ_POW_LOG_TABLE_BITS)) & ((1 << RJ_POW_LOG_TABLE_BITS) - 1));
    const int64_t k = (int64_t)tmp >> 52;
    const uint64_t iz = ix - (tmp & (0xfffULL << 52));
    const double z = rj_pow_from_bits(iz);
    const double kd = (double)k;

#if defined(__CUDA_ARCH__)
    const double invc = __ldg(&rj_pow_log_table[i].invc);
    const double logc = __ldg(&rj_pow_log_table[i].logc);
    const double logctail = __ldg(&rj_pow_log_table[i].logctail);
#else
    const double invc = rj_pow_log_table_host[i].invc;
    const double logc = rj_pow_log_table_host[i].logc;
    const double logctail = rj_pow_log_table_host[i].logctail;
#endif

    const double r = RJ_POW_FMA(z, invc, -1.0); /* exact */

    /* t1 + e1 = k ln2_hi + logc exactly (k ln2_hi is exact) */
    double t1, e1;
    rj_two_sum(RJ_POW_MUL(kd, RJ_POW_LN2_HI), logc, t1, e1);
    /* t2 + e2 = t1 + r */
    double t2, e2;
    rj_two_sum(t1, r, t2, e2);
    /* -r^2/2 as double-double */
    const double r2 = RJ_POW_MUL(r, r);
    const double a_hi = RJ_POW_MUL(-0.5, r2);
    const double a_lo = RJ_POW_MUL(-0.5, RJ_POW_FMA(r, r, -r2));
    double t3, e3;
    rj_two_sum(t2, a_hi, t3, e3);
    /* r^3 (1/3 - r/4 + r^2/5 - r^3/6 + r^4/7 - r^5/8) */
    double p = RJ_POW_FMA(r, -0.125, 0x1.2492492492492p-3);
    p = RJ_POW_FMA(r, p, -0x1.5555555555555p-3);
    p = RJ_POW_FMA(r, p, 0x1.999999999999ap-3);
    p = RJ_POW_FMA(r, p, -0.25);
    p = RJ_POW_FMA(r, p, 0x1.5555555555555p-2);
    p = RJ_POW_MUL(RJ_POW_MUL(r, r2), p);
    double lo = RJ_POW_FMA(kd, RJ_POW_LN2_LO, logctail);
    lo = RJ_POW_ADD(lo, e1);
    lo = RJ_POW_ADD(lo, e2);
    lo = RJ_POW_ADD(lo, e3);
    lo = RJ_POW_ADD(lo, a_lo);
    lo = RJ_POW_ADD(lo, p);
    const double hi = RJ_POW_ADD(t3, lo);
    lo = RJ_POW_ADD(RJ_POW_ADD(t3, -hi), lo);

    /* ---- y log(x) = ehi + elo ---- */
    const double ehi = RJ_POW_MUL(y, hi);
    const double elo = RJ_POW_FMA(y, lo, RJ_POW_FMA(y, hi, -ehi));
    if (!(fabs(ehi) < 700.0))
      return pow(x, y);

    /* ---- exp(ehi + elo) ---- */
    const double shift = 0x1.8p52;
    const double zk = RJ_POW_MUL(ehi, RJ_POW_INVLN2N);
    const double kdd = RJ_POW_ADD(zk, shift);
    const uint64_t ki = rj_pow_bits(kdd);
    const double kde = RJ_POW_ADD(kdd, -shift);
    double rr = RJ_POW_FMA(kde, -RJ_POW_LN2N_HI, ehi);
    rr = RJ_POW_FMA(kde, -RJ_POW_LN2N_LO, rr);
    rr = RJ_POW_ADD(rr, elo);

    const int j = (int)(ki & ((1 << RJ_POW_EXP_TABLE_BITS) - 1));
#if defined(__CUDA_ARCH__)
    const double thi = __ldg(&rj_pow_exp_table[j].hi);
    const double ttail = __ldg(&rj_pow_exp_table[j].tail);
#else
    const double thi = rj_pow_exp_table_host[j].hi;
    const double ttail = rj_pow_exp_table_host[j].tail;
#endif
    /* 2^(k/N): add the integer part of k/N to the exponent field */
    const int64_t top = ((int64_t)ki >> RJ_POW_EXP_TABLE_BITS) << 52;
    const double scale = rj_pow_from_bits(rj_pow_bits(thi) + (uint64_t)top);

    /* exp(rr) - 1 = rr + rr^2/2 + ... + rr^6/720 */
    const double rr2 = RJ_POW_MUL(rr, rr);
    double q = RJ_POW_FMA(rr, 0x1.6c16c16c16c17p-10, 0x1.1111111111111p-7);
    q = RJ_POW_FMA(rr, q, 0x1.5555555555555p-5);
    q = RJ_POW_FMA(rr, q, 0x1.5555555555555p-3);
    q = RJ_POW_FMA(rr, q, 0.5);
    const double t = RJ_POW_ADD(RJ_POW_ADD(ttail, rr), RJ_POW_MUL(rr2, q));
    return RJ_POW_FMA(scale, t, scale);
  }
} // namespace rjb
