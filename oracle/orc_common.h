// ORACLE -- TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of ryujin's explicit hyperbolic update path.  Nothing under
// oracle/ is part of the product: only tests/, __graft_entry__.smoke() and the
// cpu_baseline / --impl reference legs of bench.py may load it.
//
// Parity pinning: the functions in this directory are checked against the
// reference's own golden outputs (tests/golden/*.json, extracted from
// /root/reference/tests/{euler,common,shallow_water}/*.output by
// tests/golden/extract_reference_goldens.py).
//
// This header: scalar helpers.
//   positive_part / negative_part   <- source/simd.h:112-127
//   quadratic_newton_step           <- source/newton.h:39-101
//   pow                             <- source/simd.template.h:196-231 (VCL pow
//                                      on x86; libm pow in the "port" build)
#pragma once

#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

#ifdef ORC_USE_RJ_POW
// liboracle_rjpow.so only: the product's pow (ryujin_b200/csrc/accurate_pow.h, host path).  With
// it the CPU restatement and the CUDA path execute literally the same operations, so a parity
// test can demand BIT-IDENTICAL states over hundreds of updates; the libm build stays the
// independent check of that pow (0.504 ulp, equal to libm in 99.97 % of all calls).
#include "../ryujin_b200/csrc/accurate_pow.h"
#endif

#ifdef ORC_USE_VCL
// oracle/_ref build only: Agner Fog's vector class library as vendored by the
// reference (simd-math/, vectorclass/version2 @08959eb). Compiled from where it
// lies under /root/reference; never copied into this repository.
#include <vectorclass.h>
#include <vectormath_exp.h>
#endif

namespace orc
{
  constexpr double eps_double = std::numeric_limits<double>::epsilon();
  constexpr double min_double = std::numeric_limits<double>::min();
  constexpr double max_double = std::numeric_limits<double>::max();

  inline double pow(double x, double b)
  {
#if defined(ORC_USE_VCL)
    return ::pow(Vec2d(x), b).extract(0);
#elif defined(ORC_USE_RJ_POW)
    return rjb::rj_pow(x, b);
#else
    return std::pow(x, b);
#endif
  }

  inline double positive_part(double x)
  {
    return 0.5 * (std::abs(x) + x);
  }

  inline double negative_part(double x)
  {
    return 0.5 * (std::abs(x) - x);
  }

  /* One step of the quadratic Newton bracketing iteration on [p_1, p_2]. */
  inline void quadratic_newton_step(double &p_1,
                                    double &p_2,
                                    const double phi_p_1,
                                    const double phi_p_2,
                                    const double dphi_p_1,
                                    const double dphi_p_2,
                                    const double sign = 1.0)
  {
    const double scaling = 1. / (p_2 - p_1 + eps_double);

    const double dd_11 = dphi_p_1;
    const double dd_12 = (phi_p_2 - phi_p_1) * scaling;
    const double dd_22 = dphi_p_2;

    const double dd_112 = (dd_12 - dd_11) * scaling;
    const double dd_122 = (dd_22 - dd_12) * scaling;

    const double disc_1 = std::abs(dphi_p_1 * dphi_p_1 - 4. * phi_p_1 * dd_112);
    const double disc_2 = std::abs(dphi_p_2 * dphi_p_2 - 4. * phi_p_2 * dd_122);

    const double den_1 = dphi_p_1 + sign * std::sqrt(disc_1);
    const double den_2 = dphi_p_2 + sign * std::sqrt(disc_2);

    double t_1 = p_1 - (std::abs(den_1) < eps_double ? 0. : 2. * phi_p_1 / den_1);
    double t_2 = p_2 - (std::abs(den_2) < eps_double ? 0. : 2. * phi_p_2 / den_2);

    t_1 = std::max(p_1, t_1);
    t_1 = std::min(p_2, t_1);
    t_2 = std::max(p_1, t_2);
    t_2 = std::min(p_2, t_2);

    p_1 = std::min(t_1, t_2);
    p_2 = std::max(t_1, t_2);
  }

  /* Boundary ids, source/discretization.h Boundary enum. */
  enum Boundary : int {
    do_nothing = 0,
    periodic = 1,
    slip = 2,
    no_slip = 3,
    dirichlet = 4,
    dynamic = 5,
    dirichlet_momentum = 6
  };

  template <int dim>
  using Vec = std::array<double, dim>;

  template <int dim>
  inline double dot(const Vec<dim> &a, const Vec<dim> &b)
  {
    double s = a[0] * b[0];
    for (int d = 1; d < dim; ++d)
      s += a[d] * b[d];
    return s;
  }

  template <int dim>
  inline double norm_square(const Vec<dim> &a)
  {
    return dot<dim>(a, a);
  }

  template <int dim>
  inline double norm(const Vec<dim> &a)
  {
    return std::sqrt(norm_square<dim>(a));
  }
} // namespace orc
