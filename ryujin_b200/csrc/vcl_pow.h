// vcl_pow<fma>(x, y): scalar restatement of the pow() the reference calls on this path.
//
// ryujin::pow(double, double) is `vcl::pow(vcl::Vec2d(x), b).extract(0)`
// (source/simd.template.h:207-213), i.e. Agner Fog's vector class library as vendored under
// simd-math/: pow_template_d (simd-math/vectormath_exp.h:1508-1688) on top of
//   fraction_2 / exponent_f           simd-math/vectormath_exp.h:611-615, 694-712
//   polynomial_6 / _6n / _13m         simd-math/vectormath_common.h:226-243, 311-320
//   round / roundi / if_add           simd-math/vectorf128.h:2096-2123, 2219-2225, 1856-1862
//   mul_add / nmul_add / mul_sub_x    simd-math/vectorf128.h:2279-2332
// Every step is a deterministic sequence of IEEE double operations, so a lane-wise restatement
// with the same operations in the same order gives the same bits as the reference's SIMD lanes.
// VCL has two arithmetic forms, selected by the instruction set it is compiled for:
//   fma = false  (no __FMA__: SSE2 ... AVX): mul_add(a,b,c) = a*b + c with two roundings and
//                mul_sub_x = Dekker product with 27-bit masks.  This is the form
//                oracle/_ref/liboracle_vcl.so is built in (-march=x86-64-v2) and the default of
//                the product (RJ_POW_IMPL = 1).
//   fma = true   (__FMA__: AVX2 / AVX-512 hosts): the three helpers are single fused
//                operations (RJ_POW_IMPL = 2; oracle/_ref/liboracle_vcl_fma.so).
// The special-case tail (zero / negative / non-finite arguments) follows the reference's select
// chain; it is never taken on admissible states but keeps NaN propagation identical.
#pragma once

#include <cmath>
#include <cstdint>
#include <cstring>

#if defined(__CUDA_ARCH__)
#define RJ_VCL_FN __device__ __forceinline__
#define RJ_VCL_FMA(a, b, c) __fma_rn((a), (b), (c))
#define RJ_VCL_ADD(a, b) __dadd_rn((a), (b))
#define RJ_VCL_SUB(a, b) __dsub_rn((a), (b))
#define RJ_VCL_MUL(a, b) __dmul_rn((a), (b))
#define RJ_VCL_DIV(a, b) __ddiv_rn((a), (b))
#define RJ_VCL_RINT(a) rint(a)
#elif defined(__CUDACC__)
#define RJ_VCL_FN __host__ __device__ inline
#else
#define RJ_VCL_FN inline
#endif
#if !defined(__CUDA_ARCH__)
/* host: the translation units that include this header are built with -ffp-contract=off */
#define RJ_VCL_FMA(a, b, c) std::fma((a), (b), (c))
#define RJ_VCL_ADD(a, b) ((a) + (b))
#define RJ_VCL_SUB(a, b) ((a) - (b))
#define RJ_VCL_MUL(a, b) ((a) * (b))
#define RJ_VCL_DIV(a, b) ((a) / (b))
#define RJ_VCL_RINT(a) std::nearbyint(a)
#endif

namespace rjb
{
  RJ_VCL_FN uint64_t vcl_bits(double x)
  {
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double_as_longlong(x);
#else
    uint64_t u;
    std::memcpy(&u, &x, sizeof(u));
    return u;
#endif
  }

  RJ_VCL_FN double vcl_from_bits(uint64_t u)
  {
#if defined(__CUDA_ARCH__)
    return __longlong_as_double((long long)u);
#else
    double x;
    std::memcpy(&x, &u, sizeof(x));
    return x;
#endif
  }

  /* a * b + c, vectorf128.h:2279-2287 */
  template <bool fma>
  RJ_VCL_FN double vcl_mul_add(double a, double b, double c)
  {
    if (fma)
      return RJ_VCL_FMA(a, b, c);
    return RJ_VCL_ADD(RJ_VCL_MUL(a, b), c);
  }

  /* c - a * b, vectorf128.h:2301-2309 */
  template <bool fma>
  RJ_VCL_FN double vcl_nmul_add(double a, double b, double c)
  {
    if (fma)
      return RJ_VCL_FMA(-a, b, c);
    return RJ_VCL_SUB(c, RJ_VCL_MUL(a, b));
  }

  /* a * b - c with extra precision, vectorf128.h:2316-2332 */
  template <bool fma>
  RJ_VCL_FN double vcl_mul_sub_x(double a, double b, double c)
  {
    if (fma)
      return RJ_VCL_FMA(a, b, -c);
    const uint64_t upper_mask = 0xFFFFFFFFF8000000ull; /* -(1 << 27) */
    const double a_high = vcl_from_bits(vcl_bits(a) & upper_mask);
    const double b_high = vcl_from_bits(vcl_bits(b) & upper_mask);
    const double a_low = RJ_VCL_SUB(a, a_high);
    const double b_low = RJ_VCL_SUB(b, b_high);
    const double r1 = RJ_VCL_MUL(a_high, b_high);
    const double r2 = RJ_VCL_SUB(r1, c);
    const double mixed = RJ_VCL_ADD(RJ_VCL_MUL(a_high, b_low), RJ_VCL_MUL(b_high, a_low));
    return RJ_VCL_ADD(RJ_VCL_ADD(r2, mixed), RJ_VCL_MUL(a_low, b_low));
  }

  /* the select chain after the fast path, vectormath_exp.h:1637-1687 (scalar lane) */
  RJ_VCL_FN double vcl_pow_special(double x0, double y, double z, double ee)
  {
    const uint64_t bx = vcl_bits(x0), by = vcl_bits(y);
    const uint64_t exp_mask = 0x7FF0000000000000ull;
    const bool xfinite = (bx & exp_mask) != exp_mask;
    const bool yfinite = (by & exp_mask) != exp_mask;
    const bool efinite = (vcl_bits(ee) & exp_mask) != exp_mask;
    const bool xzero = (bx & exp_mask) == 0; /* zero or subnormal */
    const bool xsign = (bx >> 63) != 0;
    const double inf = vcl_from_bits(exp_mask);
    const double x1 = fabs(x0);

    /* wm_pow_case_x0, :1478-1484 */
    if (xzero)
      z = y < 0. ? inf : (y == 0. ? 1. : 0.);

    double yodd = 0.;
    if (xsign) {
      const bool yinteger = y == RJ_VCL_RINT(y);
      /* roundi(y) << 63: truncatei(round(y)); out of range conversions give INT64_MIN on x86 */
      const double yr = RJ_VCL_RINT(y);
      uint64_t yi;
      if (yr >= -9223372036854775808. && yr < 9223372036854775808.)
        yi = (uint64_t)(int64_t)yr;
      else
        yi = 0x8000000000000000ull;
      yodd = vcl_from_bits(yi << 63);
      const double nan_pow = vcl_from_bits(0x7FF8000000000000ull | 0x101ull); /* NAN_POW payload */
      const double z1 =
          yinteger ? vcl_from_bits(vcl_bits(z) | vcl_bits(yodd)) : (x0 == 0. ? z : nan_pow);
      yodd = yinteger ? yodd : 0.;
      z = z1;
    }

    if (xfinite && yfinite && (efinite || xzero))
      return z;

    double z1 = (yfinite && efinite) ? z : (x1 == 1. ? 1. : (((x1 > 1.) != ((by >> 63) != 0)) ? inf : 0.));
    if (!xfinite)
      z1 = y == 0. ? 1.
                   : (y < 0. ? vcl_from_bits(vcl_bits(yodd) & vcl_bits(z))
                             : vcl_from_bits(vcl_bits(x1) | (bx & vcl_bits(yodd))));
    if (x0 != x0 || y != y)
      z1 = x0 + y;
    return z1;
  }

  template <bool fma>
  RJ_VCL_FN double vcl_pow(const double x0, const double y)
  {
    const double ln2d_hi = 0.693145751953125;
    const double ln2d_lo = 1.42860682030941723212E-6;
    const double log2e = 1.44269504088896340736; /* VM_LOG2E */
    const double ln2 = 0.693147180559945309417;  /* VM_LN2 */
    const double sqrt2_half = 1.41421356237309504880 * 0.5;

    const double P0logl = 2.0039553499201281259648E1;
    const double P1logl = 5.7112963590585538103336E1;
    const double P2logl = 6.0949667980987787057556E1;
    const double P3logl = 2.9911919328553073277375E1;
    const double P4logl = 6.5787325942061044846969E0;
    const double P5logl = 4.9854102823193375972212E-1;
    const double P6logl = 4.5270000862445199635215E-5;
    const double Q0logl = 6.0118660497603843919306E1;
    const double Q1logl = 2.1642788614495947685003E2;
    const double Q2logl = 3.0909872225312059774938E2;
    const double Q3logl = 2.2176239823732856465394E2;
    const double Q4logl = 8.3047565967967209469434E1;
    const double Q5logl = 1.5062909083469192043167E1;

    const double p2 = 1. / 2.;
    const double p3 = 1. / 6.;
    const double p4 = 1. / 24.;
    const double p5 = 1. / 120.;
    const double p6 = 1. / 720.;
    const double p7 = 1. / 5040.;
    const double p8 = 1. / 40320.;
    const double p9 = 1. / 362880.;
    const double p10 = 1. / 3628800.;
    const double p11 = 1. / 39916800.;
    const double p12 = 1. / 479001600.;
    const double p13 = 1. / 6227020800.;

    const double x1 = fabs(x0);
    const uint64_t bx1 = vcl_bits(x1);

    /* fraction_2: mantissa * 0.5 */
    double x = vcl_from_bits((bx1 & 0x000FFFFFFFFFFFFFull) | 0x3FE0000000000000ull);
    const bool blend = x > sqrt2_half;
    x = RJ_VCL_ADD(x, blend ? 0. : x); /* if_add(!blend, x, x) */

    x = RJ_VCL_SUB(x, 1.0);
    const double x2 = RJ_VCL_MUL(x, x);
    const double x4 = RJ_VCL_MUL(x2, x2);

    /* polynomial_6(x, P0..P6) */
    double px = vcl_mul_add<fma>(
        vcl_mul_add<fma>(P6logl, x2, vcl_mul_add<fma>(P5logl, x, P4logl)),
        x4,
        vcl_mul_add<fma>(vcl_mul_add<fma>(P3logl, x, P2logl), x2, vcl_mul_add<fma>(P1logl, x, P0logl)));
    px = RJ_VCL_MUL(px, RJ_VCL_MUL(x, x2));
    /* polynomial_6n(x, Q0..Q5) */
    const double qx = vcl_mul_add<fma>(
        vcl_mul_add<fma>(Q5logl, x, RJ_VCL_ADD(Q4logl, x2)),
        x4,
        vcl_mul_add<fma>(vcl_mul_add<fma>(Q3logl, x, Q2logl), x2, vcl_mul_add<fma>(Q1logl, x, Q0logl)));
    const double lg1 = RJ_VCL_DIV(px, qx);

    /* exponent_f: (2^52 + biased exponent) - (2^52 + 1023), exact */
    double ef = RJ_VCL_SUB(vcl_from_bits((bx1 >> 52) | 0x4330000000000000ull), 4503599627370496.0 + 1023.0);
    ef = RJ_VCL_ADD(ef, blend ? 1. : 0.);

    const double e1 = RJ_VCL_RINT(RJ_VCL_MUL(ef, y));
    const double yr = vcl_mul_sub_x<fma>(ef, y, e1);

    const double lg = RJ_VCL_ADD(vcl_nmul_add<fma>(0.5, x2, x), lg1);
    const double x2err = vcl_mul_sub_x<fma>(RJ_VCL_MUL(0.5, x), x, RJ_VCL_MUL(0.5, x2));
    const double lgerr = RJ_VCL_SUB(vcl_mul_add<fma>(0.5, x2, RJ_VCL_SUB(lg, x)), lg1);

    const double e2 = RJ_VCL_RINT(RJ_VCL_MUL(RJ_VCL_MUL(lg, y), log2e));
    double v = vcl_mul_sub_x<fma>(lg, y, RJ_VCL_MUL(e2, ln2d_hi));
    v = vcl_nmul_add<fma>(e2, ln2d_lo, v);
    v = vcl_mul_add<fma>(yr, ln2, v);
    v = vcl_nmul_add<fma>(RJ_VCL_ADD(lgerr, x2err), y, v);

    /* exp */
    x = v;
    const double e3 = RJ_VCL_RINT(RJ_VCL_MUL(x, log2e));
    x = vcl_nmul_add<fma>(e3, ln2, x);

    /* polynomial_13m(x, p2..p13) */
    const double t2 = RJ_VCL_MUL(x, x);
    const double t4 = RJ_VCL_MUL(t2, t2);
    const double t8 = RJ_VCL_MUL(t4, t4);
    double z = vcl_mul_add<fma>(
        vcl_mul_add<fma>(vcl_mul_add<fma>(p13, x, p12),
                         t4,
                         vcl_mul_add<fma>(vcl_mul_add<fma>(p11, x, p10), t2, vcl_mul_add<fma>(p9, x, p8))),
        t8,
        vcl_mul_add<fma>(vcl_mul_add<fma>(vcl_mul_add<fma>(p7, x, p6), t2, vcl_mul_add<fma>(p5, x, p4)),
                         t4,
                         vcl_mul_add<fma>(vcl_mul_add<fma>(p3, x, p2), t2, x)));
    z = RJ_VCL_ADD(z, 1.0);

    const double ee = RJ_VCL_ADD(RJ_VCL_ADD(e1, e2), e3);
    /* roundi(ee): only consumed when |ee| <= 3000 */
    const bool ee_in_range = ee >= -4000. && ee <= 4000.;
    const int64_t ei = ee_in_range ? (int64_t)RJ_VCL_RINT(ee) : 0;
    const int64_t ej = ei + (int64_t)(vcl_bits(z) >> 52);
    const bool overflow = !ee_in_range ? ee > 3000. : (ej >= 0x07FF || ee > 3000.);
    const bool underflow = !ee_in_range ? ee < -3000. : (ej <= 0x0000 || ee < -3000.);

    z = vcl_from_bits(vcl_bits(z) + ((uint64_t)ei << 52));
    if (underflow)
      z = 0.;
    if (overflow)
      z = vcl_from_bits(0x7FF0000000000000ull);

    /* fast return: positive normal finite x, finite y, finite ee */
    const uint64_t exp_mask = 0x7FF0000000000000ull;
    const uint64_t bx = vcl_bits(x0);
    const bool plain = (bx >> 63) == 0 && (bx & exp_mask) != 0 && (bx & exp_mask) != exp_mask &&
                       (vcl_bits(y) & exp_mask) != exp_mask && (vcl_bits(ee) & exp_mask) != exp_mask;
    if (plain)
      return z;
    return vcl_pow_special(x0, y, z, ee);
  }
} // namespace rjb
