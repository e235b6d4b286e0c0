// Polytropic-gas Euler equation module as __device__ code (sm_100a).
//
// Device restatement of the reference's equation plug-in ("Description",
// source/euler/description.h:31-50):
//   HyperbolicSystemView  source/euler/hyperbolic_system.h:705-1216
//   RiemannSolver         source/euler/riemann_solver.template.h:23-597
//   Indicator             source/euler/indicator.h:187-258
//   Limiter               source/euler/limiter.h:255-377, limiter.template.h:15-327
//
// One thread owns one stencil row, so every per-row accumulation runs in the
// reference's column order (gotcha 1).  Fluxes are never
// materialised as k x dim tensors: the sweep only ever needs f(U).c, which is
// formed directly from (U, p, 1/rho) -- this keeps the per-thread register
// footprint at k + 2 doubles per state instead of k*dim.
#pragma once

#include "device_common.cuh"

namespace rjb
{
  struct EulerParams {
    double gamma, gamma_inverse, gamma_plus_one_inverse, gamma_minus_one_inverse;
    double reference_density, vacuum_relaxation_small, vacuum_relaxation_large;
    double evc_factor;
    double limiter_newton_tolerance, limiter_relaxation_factor;
    unsigned int limiter_newton_max_iterations;
    double riemann_newton_tolerance;
    unsigned int riemann_newton_max_iterations;
  };


  template <int dim_>
  struct Euler {
    static constexpr int dim = dim_;
    static constexpr int k = dim + 2;      /* problem_dimension */
    static constexpr int n_prec = 2;       /* (specific entropy, Harten entropy) */
    static constexpr int n_bounds = 3;     /* rho_min, rho_max, s_min */
    static constexpr bool shallow_water = false;
    static constexpr bool aeos = false;
    static constexpr bool have_source_terms = false;
    static constexpr bool have_high_order_flux = false;
    static constexpr bool edge_cache = true; /* kernels.cuh: RJ_EDGE_CACHE */
    using Params = EulerParams;

    /* ---------------- HyperbolicSystemView ---------------- */

    RJ_DEVICE static double internal_energy(const double (&U)[k])
    {
      const double rho_inverse = 1. / U[0];
      double m2 = U[1] * U[1];
#pragma unroll
      for (int d = 1; d < dim; ++d)
        m2 += U[1 + d] * U[1 + d];
      return U[dim + 1] - 0.5 * m2 * rho_inverse;
    }

    RJ_DEVICE static double pressure(const Params &p, const double (&U)[k])
    {
      return (p.gamma - 1.) * internal_energy(U);
    }

    RJ_DEVICE static double specific_entropy(const Params &p, const double (&U)[k])
    {
      const double rho_inverse = 1. / U[0];
      return internal_energy(U) * path_pow(rho_inverse, p.gamma);
    }

    RJ_DEVICE static double harten_entropy(const Params &p, const double (&U)[k])
    {
      double m2 = U[1] * U[1];
#pragma unroll
      for (int d = 1; d < dim; ++d)
        m2 += U[1 + d] * U[1 + d];
      const double rho_rho_e = U[0] * U[dim + 1] - 0.5 * m2;
      return path_pow(rho_rho_e, p.gamma_plus_one_inverse);
    }

    /* precomputation_loop body, hyperbolic_system.h:732-735 */
    RJ_DEVICE static void precompute(const Params &p, const double (&U)[k], double (&prec)[n_prec])
    {
      prec[0] = specific_entropy(p, U);
      prec[1] = harten_entropy(p, U);
    }

    /* Per-node part of flux_contribution (1/rho, p), stored once per prepared
     * state vector next to the precomputed values instead of being recomputed
     * for every stencil entry (the reference recomputes f(U_j) per entry,
     * indicator.h:228, hyperbolic_module.template.h:760-761).  Same operations,
     * same rounding: values are bit-identical to flux_contribution(). */
    static constexpr int n_aux = 2;

    RJ_DEVICE static void
    harten_entropy_derivative(const Params &p, const double (&U)[k], double (&result)[k])
    {
      double m2 = U[1] * U[1];
#pragma unroll
      for (int d = 1; d < dim; ++d)
        m2 += U[1 + d] * U[1 + d];
      const double rho_rho_e = U[0] * U[dim + 1] - 0.5 * m2;
      const double factor =
          p.gamma_plus_one_inverse * path_pow(rho_rho_e, -p.gamma * p.gamma_plus_one_inverse);
      result[0] = factor * U[dim + 1];
#pragma unroll
      for (int d = 0; d < dim; ++d)
        result[1 + d] = -factor * U[1 + d];
      result[dim + 1] = factor * U[0];
    }

    /* What the sweep keeps per state instead of the flux tensor f(U):
     * flux_contribution_type of the reference is f(U) itself
     * (hyperbolic_system.h:1186-1205). */
    struct Flux {
      double rho_inverse, p;
    };

    RJ_DEVICE static void store_aux(const Flux &f, double (&aux)[n_aux])
    {
      aux[0] = f.rho_inverse;
      aux[1] = f.p;
    }

    RJ_DEVICE static Flux load_aux(const double (&aux)[n_aux])
    {
      Flux f;
      f.rho_inverse = aux[0];
      f.p = aux[1];
      return f;
    }


    RJ_DEVICE static Flux
    flux_contribution(const Params &prm, const double (&/*prec*/)[n_prec], const double (&U)[k])
    {
      Flux f;
      f.rho_inverse = 1. / U[0];
      double m2 = U[1] * U[1];
#pragma unroll
      for (int d = 1; d < dim; ++d)
        m2 += U[1 + d] * U[1 + d];
      f.p = (prm.gamma - 1.) * (U[dim + 1] - 0.5 * m2 * f.rho_inverse);
      return f;
    }

    /* Element (comp, d) of the flux tensor f(U), hyperbolic_system.h:1164-1181,
     * rebuilt from (U, 1/rho, p) with the reference's operation order:
     *   f[0] = m;  f[1+i] = m * (m[i] * rho_inverse), f[1+i][i] += p;
     *   f[dim+1] = m * (rho_inverse * (E + p)).
     * comp and d are compile time constants after unrolling. */
    RJ_DEVICE static double f_element(const double (&U)[k], const Flux &f, const int comp, const int d)
    {
      if (comp == 0)
        return U[1 + d];
      if (comp == dim + 1)
        return U[1 + d] * (f.rho_inverse * (U[dim + 1] + f.p));
      const double v = U[1 + d] * (U[comp] * f.rho_inverse);
      return (d == comp - 1) ? v + f.p : v;
    }

    /* flux_divergence(flux_i, flux_j, c_ij) = -contract(add(flux_i, flux_j), c_ij),
     * hyperbolic_system.h:1210-1216, convenience_macros.h:75-96: per component
     * the sum over d runs in ascending order on the added tensors. */
    RJ_DEVICE static void flux_divergence(const double (&U_i)[k],
                                          const Flux &f_i,
                                          const double (&U_j)[k],
                                          const Flux &f_j,
                                          const double (&c)[dim],
                                          double (&out)[k])
    {
#pragma unroll
      for (int q = 0; q < k; ++q) {
        double s = (f_element(U_i, f_i, q, 0) + f_element(U_j, f_j, q, 0)) * c[0];
#pragma unroll
        for (int d = 1; d < dim; ++d)
          s += (f_element(U_i, f_i, q, d) + f_element(U_j, f_j, q, d)) * c[d];
        out[q] = -s;
      }
    }

    RJ_DEVICE static double filter_vacuum_density(const Params &p, const double rho)
    {
      const double cutoff = p.reference_density * p.vacuum_relaxation_large * DBL_EPSILON;
      return fabs(rho) < cutoff ? 0. : rho;
    }

    /* ---------------- boundary conditions ---------------- */

    RJ_DEVICE static double speed_of_sound(const Params &p, const double (&U)[k])
    {
      const double rho_inverse = 1. / U[0];
      return sqrt(p.gamma * pressure(p, U) * rho_inverse);
    }

    /* hyperbolic_system.h:1040-1093 */
    template <int component>
    RJ_DEVICE static void prescribe_riemann_characteristic(const Params &p,
                                                           const double (&U)[k],
                                                           const double (&U_bar)[k],
                                                           const double (&normal)[dim],
                                                           double (&U_new)[k])
    {
      const double gamma = p.gamma;
      const double rho = U[0];
      const double a = speed_of_sound(p, U);
      double mn = 0., mn_bar = 0.;
#pragma unroll
      for (int d = 0; d < dim; ++d) {
        mn += U[1 + d] * normal[d];
        mn_bar += U_bar[1 + d] * normal[d];
      }
      const double vn = mn / rho;
      const double rho_bar = U_bar[0];
      const double a_bar = speed_of_sound(p, U_bar);
      const double vn_bar = mn_bar / rho_bar;

      const double R_1 = component == 1 ? vn_bar - 2. * a_bar / (gamma - 1.) : vn - 2. * a / (gamma - 1.);
      const double R_2 = component == 2 ? vn_bar + 2. * a_bar / (gamma - 1.) : vn + 2. * a / (gamma - 1.);

      const double pr = pressure(p, U);
      const double s = pr / path_pow(rho, gamma);

      const double rho_inverse = 1. / rho; /* m / rho: Tensor / scalar multiplies by the inverse */
      double vperp[dim];
      double vperp2 = 0.;
#pragma unroll
      for (int d = 0; d < dim; ++d) {
        vperp[d] = U[1 + d] * rho_inverse - vn * normal[d];
        vperp2 += vperp[d] * vperp[d];
      }
      const double vn_new = 0.5 * (R_1 + R_2);
      const double tmp = ((gamma - 1.) / 4.) * (R_2 - R_1);
      double rho_new = 1. / (gamma * s) * (tmp * tmp);
      rho_new = path_pow(rho_new, 1. / (gamma - 1.));
      /* sic: the reference calls std::pow here, not ryujin::pow (hyperbolic_system.h:1082); on its
       * hosts that is glibc's pow, which rj_pow (accurate_pow.h, 0.50x ulp) reproduces in 99.97 % of
       * all arguments -- the closest a device routine gets to the libm the oracle links */
      const double p_new = s * rj_pow(rho_new, gamma);

      U_new[0] = rho_new;
#pragma unroll
      for (int d = 0; d < dim; ++d)
        U_new[1 + d] = rho_new * (vn_new * normal[d] + vperp[d]);
      U_new[1 + dim] = p_new / (gamma - 1.) + 0.5 * rho_new * (vn_new * vn_new + vperp2);
    }

    /* apply_boundary_conditions, hyperbolic_system.h:1099-1159; D() returns
     * the Dirichlet state lazily */
    template <typename DirichletData>
    RJ_DEVICE static void apply_boundary_conditions(const Params &p,
                                                    const int id,
                                                    double (&U)[k],
                                                    const double (&normal)[dim],
                                                    const DirichletData &D)
    {
      if (id == RJ_BOUNDARY_DIRICHLET) {
        D(U);
      } else if (id == RJ_BOUNDARY_SLIP) {
        double mn = 0.;
#pragma unroll
        for (int d = 0; d < dim; ++d)
          mn += U[1 + d] * normal[d];
#pragma unroll
        for (int d = 0; d < dim; ++d)
          U[1 + d] -= 1. * mn * normal[d];
      } else if (id == RJ_BOUNDARY_NO_SLIP) {
#pragma unroll
        for (int d = 0; d < dim; ++d)
          U[1 + d] = 0.;
      } else if (id == RJ_BOUNDARY_DYNAMIC) {
        double mn = 0.;
#pragma unroll
        for (int d = 0; d < dim; ++d)
          mn += U[1 + d] * normal[d];
        const double a = speed_of_sound(p, U);
        const double vn = mn / U[0];
        double result[k];
#pragma unroll
        for (int q = 0; q < k; ++q)
          result[q] = U[q];
        if (vn < -a)
          D(result);
        if (vn >= -a && vn <= 0.) {
          double U_dirichlet[k];
          D(U_dirichlet);
          prescribe_riemann_characteristic<2>(p, U_dirichlet, U, normal, result);
        }
        if (vn > 0. && vn <= a) {
          double U_dirichlet[k];
          D(U_dirichlet);
          prescribe_riemann_characteristic<1>(p, U, U_dirichlet, normal, result);
        }
#pragma unroll
        for (int q = 0; q < k; ++q)
          U[q] = result[q];
      }
    }

    /* ---------------- RiemannSolver ---------------- */

    struct RiemannData {
      double rho, u, p, a;
    };

    /* riemann_solver.template.h:379-403 */
    RJ_DEVICE static RiemannData
    riemann_data_from_state(const Params &p, const double (&U)[k], const double (&n)[dim])
    {
      const double rho = U[0];
      const double rho_inverse = 1.0 / rho;
      double proj_m = n[0] * U[1];
#pragma unroll
      for (int d = 1; d < dim; ++d)
        proj_m += n[d] * U[1 + d];
      double perp2 = 0.;
#pragma unroll
      for (int d = 0; d < dim; ++d) {
        const double perp = U[1 + d] - proj_m * n[d];
        perp2 += perp * perp;
      }
      const double E = U[dim + 1] - 0.5 * perp2 * rho_inverse;
      const double rho_e = E - 0.5 * (proj_m * proj_m) * rho_inverse;
      RiemannData rd;
      rd.rho = rho;
      rd.u = proj_m * rho_inverse;
      rd.p = (p.gamma - 1.) * rho_e;
      rd.a = sqrt(p.gamma * rd.p * rho_inverse);
      return rd;
    }

    RJ_DEVICE static double riemann_f(const Params &p, const RiemannData &rd, const double p_star)
    {
      const double gamma = p.gamma;
      const double Az = 2. / (rd.rho * (gamma + 1.));
      const double Bz = (gamma - 1.) / (gamma + 1.) * rd.p;
      const double radicand = Az / (p_star + Bz);
      const double true_value = (p_star - rd.p) * sqrt(radicand);
      const double exponent = 0.5 * (gamma - 1.) / gamma;
      const double factor = path_pow(p_star / rd.p, exponent) - 1.;
      const double false_value = 2. * rd.a * factor / (gamma - 1.);
      return p_star >= rd.p ? true_value : false_value;
    }

    RJ_DEVICE static double riemann_df(const Params &p, const RiemannData &rd, const double p_star)
    {
      const double gamma = p.gamma;
      const double radicand_inverse = 0.5 * rd.rho * ((gamma + 1.) * p_star + (gamma - 1.) * rd.p);
      const double denominator = (p_star + (gamma - 1.) * p.gamma_plus_one_inverse * rd.p);
      const double true_value =
          (denominator - 0.5 * (p_star - rd.p)) / (denominator * sqrt(radicand_inverse));
      const double exponent = (-1. - gamma) * 0.5 * p.gamma_inverse;
      const double factor = (gamma - 1.) * 0.5 * p.gamma_inverse * path_pow(p_star / rd.p, exponent) / rd.p;
      const double false_value = factor * 2. * rd.a * p.gamma_minus_one_inverse;
      return p_star >= rd.p ? true_value : false_value;
    }

    RJ_DEVICE static double
    lambda1_minus(const Params &p, const RiemannData &rd, const double p_star)
    {
      const double factor = (p.gamma + 1.0) * 0.5 * p.gamma_inverse;
      const double inv_p = 1.0 / rd.p;
      const double tmp = positive_part((p_star - rd.p) * inv_p);
      return rd.u - rd.a * sqrt(1.0 + factor * tmp);
    }

    RJ_DEVICE static double
    lambda3_plus(const Params &p, const RiemannData &rd, const double p_star)
    {
      const double factor = (p.gamma + 1.0) * 0.5 * p.gamma_inverse;
      const double inv_p = 1.0 / rd.p;
      const double tmp = positive_part((p_star - rd.p) * inv_p);
      return rd.u + rd.a * sqrt(1.0 + factor * tmp);
    }

    /* riemann_solver.template.h:407-582 */
    RJ_DEVICE static double
    riemann_compute(const Params &p, const RiemannData &rd_i, const RiemannData &rd_j)
    {
      const double gamma = p.gamma;
      const double p_max = fmax(rd_i.p, rd_j.p);

      /* p_star_two_rarefaction, :276-319 */
      double rarefaction;
      {
        const double inv_p_j = 1. / rd_j.p;
        const double factor = (gamma - 1.) * 0.5;
        const double numerator = positive_part(rd_i.a + rd_j.a - factor * (rd_j.u - rd_i.u));
        const double denominator =
            rd_i.a * path_pow(rd_i.p * inv_p_j, -factor * p.gamma_inverse) + rd_j.a;
        const double exponent = 2.0 * gamma * p.gamma_minus_one_inverse;
        rarefaction = rd_j.p * path_pow(numerator / denominator, exponent);
      }

      /* p_star_failsafe, :332-374; radicand_inverse terms shared with
       * phi_of_p_max, :124-149 */
      const double ri = rd_i.rho * ((gamma + 1.) * p_max + (gamma - 1.) * rd_i.p);
      const double rj = rd_j.rho * ((gamma + 1.) * p_max + (gamma - 1.) * rd_j.p);
      double failsafe;
      {
        const double x_i = sqrt(2. * p_max / ri);
        const double x_j = sqrt(2. * p_max / rj);
        const double a = x_i + x_j;
        const double b = rd_j.u - rd_i.u;
        const double c = -rd_i.p * x_i - rd_j.p * x_j;
        const double base = (-b + sqrt(b * b - 4. * a * c)) / (2. * a);
        failsafe = base * base;
      }
      const double p_star_tilde = fmin(rarefaction, failsafe);

      const double phi_p_max = (p_max - rd_i.p) / sqrt(0.5 * ri) +
                               (p_max - rd_j.p) / sqrt(0.5 * rj) + rd_j.u - rd_i.u;

      double p_2 = phi_p_max < 0. ? p_star_tilde : fmin(p_max, p_star_tilde);

      if (p.riemann_newton_max_iterations == 0)
        return fmax(positive_part(lambda3_plus(p, rd_j, p_2)),
                    negative_part(lambda1_minus(p, rd_i, p_2)));

      const double p_min = fmin(rd_i.p, rd_j.p);
      double p_1 = phi_p_max < 0. ? p_max : p_min;
      p_1 = p_1 <= p_2 ? p_1 : p_2;

      double gap, lambda_max;
      auto compute_gap = [&]() {
        const double nu_11 = lambda1_minus(p, rd_i, p_2);
        const double nu_12 = lambda1_minus(p, rd_i, p_1);
        const double nu_31 = lambda3_plus(p, rd_j, p_1);
        const double nu_32 = lambda3_plus(p, rd_j, p_2);
        lambda_max = fmax(positive_part(nu_32), negative_part(nu_11));
        gap = fmax(fabs(nu_32 - nu_31), fabs(nu_12 - nu_11));
      };
      compute_gap();

      for (unsigned int it = 0; it < p.riemann_newton_max_iterations; ++it) {
        if (fmax(0., gap - p.riemann_newton_tolerance) == 0.)
          break;
        const double phi_p_1 = riemann_f(p, rd_i, p_1) + riemann_f(p, rd_j, p_1) + rd_j.u - rd_i.u;
        const double phi_p_2 = riemann_f(p, rd_i, p_2) + riemann_f(p, rd_j, p_2) + rd_j.u - rd_i.u;
        const double dphi_p_1 = riemann_df(p, rd_i, p_1) + riemann_df(p, rd_j, p_1);
        const double dphi_p_2 = riemann_df(p, rd_i, p_2) + riemann_df(p, rd_j, p_2);
        quadratic_newton_step(p_1, p_2, phi_p_1, phi_p_2, dphi_p_1, dphi_p_2, 1.0);
        compute_gap();
      }
      return lambda_max;
    }

    /* RiemannSolver::compute(U_i, U_j, i, js, n_ij), :586-597.  prec is
     * unused by the polytropic solver. */
    RJ_DEVICE static double riemann_lambda_max(const Params &p,
                                               const double (&U_i)[k],
                                               const double (&/*prec_i*/)[n_prec],
                                               const double (&U_j)[k],
                                               const double (&/*prec_j*/)[n_prec],
                                               const double (&n_ij)[dim])
    {
      const RiemannData rd_i = riemann_data_from_state(p, U_i, n_ij);
      const RiemannData rd_j = riemann_data_from_state(p, U_j, n_ij);
      return riemann_compute(p, rd_i, rd_j);
    }

    /* ---------------- Indicator ---------------- */

    struct Indicator {
      double rho_i_inverse, eta_i, p_i;
      double d_eta_i[k];
      double left;
      double right[k];

      RJ_DEVICE void reset(const Params &p, const double (&prec_i)[n_prec], const double (&U_i)[k])
      {
        rho_i_inverse = 1. / U_i[0];
        eta_i = prec_i[1];
        harten_entropy_derivative(p, U_i, d_eta_i);
        d_eta_i[0] -= eta_i * rho_i_inverse;
        p_i = pressure(p, U_i);
        left = 0.;
#pragma unroll
        for (int q = 0; q < k; ++q)
          right[q] = 0.;
      }

      /* indicator.h:212-238 with (f_j - f_i).c = f_j.c - f_i.c */
      RJ_DEVICE void accumulate(const Params &p,
                                const double (&U_i)[k],
                                const double (&/*prec_i*/)[n_prec],
                                const double (&U_j)[k],
                                const double (&prec_j)[n_prec],
                                const Flux &fj,
                                const double (&c_ij)[dim])
      {
        const double eta_j = prec_j[1];
        Flux fi;
        fi.rho_inverse = rho_i_inverse;
        fi.p = p_i;
        double mjc = U_j[1] * c_ij[0];
#pragma unroll
        for (int d = 1; d < dim; ++d)
          mjc += U_j[1 + d] * c_ij[d];
        left += (eta_j * fj.rho_inverse - eta_i * rho_i_inverse) * mjc;
#pragma unroll
        for (int q = 0; q < k; ++q) {
          double s = (f_element(U_j, fj, q, 0) - f_element(U_i, fi, q, 0)) * c_ij[0];
#pragma unroll
          for (int d = 1; d < dim; ++d)
            s += (f_element(U_j, fj, q, d) - f_element(U_i, fi, q, d)) * c_ij[d];
          right[q] += s;
        }
      }

      RJ_DEVICE double alpha(const Params &p, const double hd_i) const
      {
        double numerator = left;
        double denominator = fabs(left);
#pragma unroll
        for (int q = 0; q < k; ++q) {
          numerator -= d_eta_i[q] * right[q];
          denominator += fabs(d_eta_i[q] * right[q]);
        }
        const double quotient = fabs(numerator) / (denominator + hd_i * fabs(eta_i));
        return fmin(1., p.evc_factor * quotient);
      }
    };

    /* ---------------- Limiter ---------------- */

    /* The one quantity of Limiter::accumulate that is symmetric in (i, j): the specific entropy of the
     * midpoint state, limiter.h:314-319.  (U_i + U_j) * .5 == (U_j + U_i) * .5 bit for bit; on the diagonal
     * (U_i + U_i) * .5 == U_i exactly, whose specific entropy is the precomputed s_i. */
    RJ_DEVICE static double edge_value(const Params &p, const double (&U_i)[k], const double (&U_j)[k])
    {
      double U_mid[k];
#pragma unroll
      for (int q = 0; q < k; ++q)
        U_mid[q] = (U_i[q] + U_j[q]) * .5;
      return specific_entropy(p, U_mid);
    }
    RJ_DEVICE static double edge_value_diagonal(const double (&prec_i)[n_prec]) { return prec_i[0]; }

    struct LimiterAccumulator {
      double rho_min, rho_max, s_min;
      double rho_relaxation_numerator, rho_relaxation_denominator, s_interp_max;

      RJ_DEVICE void reset()
      {
        rho_min = DBL_MAX;
        rho_max = 0.;
        s_min = DBL_MAX;
        rho_relaxation_numerator = 0.;
        rho_relaxation_denominator = 0.;
        s_interp_max = 0.;
      }

      /* limiter.h:280-327 */
      RJ_DEVICE void accumulate(const Params &p,
                                const double (&U_i)[k],
                                const double (&U_j)[k],
                                const double (&prec_j)[n_prec],
                                const double (&scaled_c_ij)[dim])
      {
        const double s_j = prec_j[0];
        double dmc = (U_i[1] - U_j[1]) * scaled_c_ij[0];
#pragma unroll
        for (int d = 1; d < dim; ++d)
          dmc += (U_i[1 + d] - U_j[1 + d]) * scaled_c_ij[d];
        const double rho_ij_bar = 0.5 * (U_i[0] + U_j[0] + dmc);
        rho_min = fmin(rho_min, rho_ij_bar);
        rho_max = fmax(rho_max, rho_ij_bar);
        s_min = fmin(s_min, s_j);
        rho_relaxation_numerator += (U_i[0] + U_j[0]);
        rho_relaxation_denominator += 1.;
        double U_mid[k];
#pragma unroll
        for (int q = 0; q < k; ++q)
          U_mid[q] = (U_i[q] + U_j[q]) * .5;
        const double s_interp = specific_entropy(p, U_mid);
        s_interp_max = fmax(s_interp_max, s_interp);
      }

      /* the same with s_interp = edge_value(U_i, U_j) taken from the edge cache */
      RJ_DEVICE void accumulate_cached(const double (&U_i)[k],
                                       const double (&U_j)[k],
                                       const double (&prec_j)[n_prec],
                                       const double (&scaled_c_ij)[dim],
                                       const double s_interp)
      {
        const double s_j = prec_j[0];
        double dmc = (U_i[1] - U_j[1]) * scaled_c_ij[0];
#pragma unroll
        for (int d = 1; d < dim; ++d)
          dmc += (U_i[1 + d] - U_j[1 + d]) * scaled_c_ij[d];
        const double rho_ij_bar = 0.5 * (U_i[0] + U_j[0] + dmc);
        rho_min = fmin(rho_min, rho_ij_bar);
        rho_max = fmax(rho_max, rho_ij_bar);
        s_min = fmin(s_min, s_j);
        rho_relaxation_numerator += (U_i[0] + U_j[0]);
        rho_relaxation_denominator += 1.;
        s_interp_max = fmax(s_interp_max, s_interp);
      }

      /* limiter.h:332-363 */
      RJ_DEVICE void bounds(const Params &p, const double hd_i, double (&b)[n_bounds]) const
      {
        double r_i = sqrt(hd_i);
        if (dim == 2) {
          const double q = sqrt(r_i);
          r_i = q * q * q;
        } else if (dim == 1) {
          r_i = r_i * r_i * r_i;
        }
        r_i *= p.limiter_relaxation_factor;
        const double rho_relaxation =
            fabs(rho_relaxation_numerator) / (fabs(rho_relaxation_denominator) + DBL_EPSILON);
        const double relaxation = (2. * p.limiter_relaxation_factor) * rho_relaxation;
        b[0] = fmax((1. - r_i) * rho_min, rho_min - relaxation);
        b[1] = fmin((1. + r_i) * rho_max, rho_max + relaxation);
        const double entropy_relaxation = p.limiter_relaxation_factor * (s_interp_max - s_min);
        b[2] = fmax((1. - r_i) * s_min, s_min - entropy_relaxation);
      }
    };

    /* Limiter::limit, limiter.template.h:15-327 (EXPENSIVE_BOUNDS_CHECK off).
     * Returns t_l; *success is cleared when the low-order state is out of
     * bounds. */
    RJ_DEVICE static double limit(const Params &p,
                                  const double (&bounds)[n_bounds],
                                  const double (&U)[k],
                                  const double (&P)[k],
                                  bool &success)
    {
      const double t_min = 0., t_max = 1.;
      success = true;
      double t_r = t_max;

      const double relax_small = 1. + p.vacuum_relaxation_small * DBL_EPSILON;
      const double relax = 1. + p.vacuum_relaxation_large * DBL_EPSILON;

      {
        const double rho_U = U[0];
        const double rho_P = P[0];
        const double rho_min = bounds[0];
        const double rho_max = bounds[1];

        const double test_min = filter_vacuum_density(p, fmax(0., rho_U - relax * rho_max));
        const double test_max = filter_vacuum_density(p, fmax(0., rho_min - relax * rho_U));
        if (!(test_min == 0. && test_max == 0.))
          success = false;

        const double denominator = 1. / (fabs(rho_P) + DBL_EPSILON * rho_max);
        t_r = rho_max < rho_U + t_r * rho_P ? (rho_max - rho_U) * denominator : t_r;
        t_r = rho_U + t_r * rho_P < rho_min ? (rho_U - rho_min) * denominator : t_r;
        t_r = fmin(t_r, t_max);
        t_r = fmax(t_r, t_min);
      }

      double t_l = t_min;
      const double gamma = p.gamma;
      const double gp1 = gamma + 1.;
      const double s_min = bounds[2];

      for (unsigned int n = 0; n < p.limiter_newton_max_iterations; ++n) {
        double U_r[k];
#pragma unroll
        for (int q = 0; q < k; ++q)
          U_r[q] = U[q] + t_r * P[q];
        const double rho_r = U_r[0];
        const double rho_r_gamma = path_pow(rho_r, gamma);
        const double rho_e_r = internal_energy(U_r);
        const double psi_r = relax_small * rho_r * rho_e_r - s_min * rho_r * rho_r_gamma;

        t_l = psi_r > 0. ? t_r : t_l;
        if (t_l == t_r)
          break;

        double U_l[k];
#pragma unroll
        for (int q = 0; q < k; ++q)
          U_l[q] = U[q] + t_l * P[q];
        const double rho_l = U_l[0];
        const double rho_l_gamma = path_pow(rho_l, gamma);
        const double rho_e_l = internal_energy(U_l);
        const double psi_l = relax_small * rho_l * rho_e_l - s_min * rho_l * rho_l_gamma;

        const double lower_bound = (1. - relax) * s_min * rho_l * rho_l_gamma;
        if (n == 0 && !(fmin(0., psi_l - lower_bound) == 0.))
          success = false;

        if (fmax(0., t_r - t_l - p.limiter_newton_tolerance) == 0.)
          break;

        /* internal_energy_derivative(U) . P, hyperbolic_system.h:797-819 */
        auto de_dot_P = [&](const double(&V)[k]) {
          const double rho_inverse = 1. / V[0];
          double u2 = 0.;
          double acc;
          {
            double u[dim];
#pragma unroll
            for (int d = 0; d < dim; ++d) {
              u[d] = V[1 + d] * rho_inverse;
              u2 += u[d] * u[d];
            }
            acc = (0.5 * u2) * P[0];
#pragma unroll
            for (int d = 0; d < dim; ++d)
              acc += (-u[d]) * P[1 + d];
          }
          acc += P[dim + 1];
          return acc;
        };
        const double drho = P[0];
        const double drho_e_l = de_dot_P(U_l);
        const double drho_e_r = de_dot_P(U_r);
        const double dpsi_l = rho_l * drho_e_l + (rho_e_l - gp1 * s_min * rho_l_gamma) * drho;
        const double dpsi_r = rho_r * drho_e_r + (rho_e_r - gp1 * s_min * rho_r_gamma) * drho;

        quadratic_newton_step(t_l, t_r, psi_l, psi_r, dpsi_l, dpsi_r, -1.);
      }
      return t_l;
    }

    /* ---------------- initial states (Dirichlet data on device) -------- */

    RJ_DEVICE static void
    from_initial_state(const Params &p, const double (&prim)[3], double (&U)[k])
    {
      /* hyperbolic_system.h:1245-1272 */
      const double rho = prim[0], u = prim[1], pr = prim[2];
#pragma unroll
      for (int q = 0; q < k; ++q)
        U[q] = 0.;
      U[0] = rho;
      U[1] = rho * u;
      U[dim + 1] = pr / (p.gamma - 1.) + 0.5 * rho * (u * u);
    }

    /* InitialState::compute(point, t) of the configured state; IV is
     * InitialValuesDevice (initial_values.cuh) */
    template <class IV>
    RJ_DEVICE static void initial_state_compute(
        const Params &p, const IV &iv, const double (&x)[dim], const double t, double (&U)[k])
    {
      if (iv.kind == RJ_INITIAL_CONTRAST) {
        /* euler/initial_state_contrast.h:62-65 */
        if (x[0] > 0.)
          from_initial_state(p, iv.primitive_right, U);
        else
          from_initial_state(p, iv.primitive, U);
      } else if (iv.kind == RJ_INITIAL_ISENTROPIC_VORTEX) {
        /* euler/initial_state_isentropic_vortex.h:53-92 */
        const double gamma = p.gamma;
        const double pb0 = x[0] - iv.mach_number * t;
        const double pb1 = dim >= 2 ? x[dim >= 2 ? 1 : 0] : 0.;
        const double r_square = pb0 * pb0 + pb1 * pb1;
        const double factor = iv.beta / (2. * 3.14159265358979323846) * exp(0.5 - 0.5 * r_square);
        const double T = 1. - (gamma - 1.) / (2. * gamma) * factor * factor;
        const double u = iv.mach_number - factor * pb1;
        const double v = factor * pb0;
        const double rho = path_pow(T, 1. / (gamma - 1.));
        const double pr = path_pow(rho, gamma);
        const double E = pr / (gamma - 1.) + 0.5 * rho * (u * u + v * v);
#pragma unroll
        for (int q = 0; q < k; ++q)
          U[q] = 0.;
        U[0] = rho;
        U[1] = rho * u;
        if (dim >= 2)
          U[dim >= 2 ? 2 : 1] = rho * v;
        U[dim + 1] = E;
      } else {
        /* euler/initial_state_uniform.h:49-53 */
        from_initial_state(p, iv.primitive, U);
      }
    }
  };
} // namespace rjb
