// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_common.h).
//
// Polytropic-gas Euler equation module, scalar double semantics.
//   System / View    <- source/euler/hyperbolic_system.h:665-1272
//   RiemannSolver    <- source/euler/riemann_solver.template.h:23-597
//   Indicator        <- source/euler/indicator.h:187-258
//   Limiter          <- source/euler/limiter.h:255-377,
//                       source/euler/limiter.template.h:15-327
#pragma once

#include "orc_common.h"

namespace orc
{
  namespace euler
  {
    /* Runtime parameters: hyperbolic_system.h:668-684, riemann_solver.h:26-38,
     * indicator.h:28-31, limiter.h:26-49. */
    struct Parameters {
      double gamma = 7. / 5.;
      double reference_density = 1.;
      double vacuum_state_relaxation_small = 1.e2;
      double vacuum_state_relaxation_large = 1.e4;

      double riemann_newton_tolerance = 1.e-10;
      unsigned int riemann_newton_max_iterations = 0;

      double indicator_evc_factor = 1.;

      unsigned int limiter_iterations = 2;
      double limiter_newton_tolerance = 1.e-10;
      unsigned int limiter_newton_max_iterations = 2;
      double limiter_relaxation_factor = 1.;
      bool expensive_bounds_check = false; /* compile time option upstream */
    };


    template <int dim>
    struct System {
      static constexpr int k = dim + 2;           /* problem_dimension */
      static constexpr int n_precomputed = 2;     /* (s, eta) */
      static constexpr int n_bounds = 3;          /* rho_min, rho_max, s_min */
      static constexpr int n_initial_precomputed = 0;
      static constexpr bool shallow_water = false;
      static constexpr bool aeos = false;
      static constexpr bool have_source_terms = false;
      static constexpr bool have_high_order_flux = false;
      using state_type = std::array<double, k>;
      using flux_type = std::array<Vec<dim>, k>;
      using precomputed_type = std::array<double, n_precomputed>;
      using bounds_type = std::array<double, n_bounds>;

      Parameters prm;
      double gamma, gamma_inverse, gamma_plus_one_inverse,
          gamma_minus_one_inverse, gamma_minus_one_over_gamma_plus_one;

      explicit System(const Parameters &p = Parameters())
          : prm(p)
      {
        gamma = prm.gamma;
        gamma_inverse = 1. / gamma;
        gamma_plus_one_inverse = 1. / (gamma + 1.);
        gamma_minus_one_inverse = 1. / (gamma - 1.);
        gamma_minus_one_over_gamma_plus_one = (gamma - 1.) / (gamma + 1.);
      }

      static double density(const state_type &U) { return U[0]; }

      static Vec<dim> momentum(const state_type &U)
      {
        Vec<dim> m;
        for (int d = 0; d < dim; ++d)
          m[d] = U[1 + d];
        return m;
      }

      static double total_energy(const state_type &U) { return U[1 + dim]; }

      double filter_vacuum_density(const double rho) const
      {
        const double cutoff = prm.reference_density *
                              prm.vacuum_state_relaxation_large * eps_double;
        return std::abs(rho) < cutoff ? 0. : rho;
      }

      static double internal_energy(const state_type &U)
      {
        const double rho_inverse = 1. / density(U);
        const auto m = momentum(U);
        return total_energy(U) - 0.5 * norm_square<dim>(m) * rho_inverse;
      }

      static state_type internal_energy_derivative(const state_type &U)
      {
        const double rho_inverse = 1. / density(U);
        Vec<dim> u = momentum(U);
        for (int d = 0; d < dim; ++d)
          u[d] *= rho_inverse;
        state_type result;
        result[0] = 0.5 * norm_square<dim>(u);
        for (int d = 0; d < dim; ++d)
          result[1 + d] = -u[d];
        result[dim + 1] = 1.;
        return result;
      }

      double pressure(const state_type &U) const
      {
        return (gamma - 1.) * internal_energy(U);
      }

      double speed_of_sound(const state_type &U) const
      {
        const double rho_inverse = 1. / density(U);
        const double p = pressure(U);
        return std::sqrt(gamma * p * rho_inverse);
      }

      double specific_entropy(const state_type &U) const
      {
        const double rho_inverse = 1. / density(U);
        return internal_energy(U) * orc::pow(rho_inverse, gamma);
      }

      double harten_entropy(const state_type &U) const
      {
        const double rho = density(U);
        const auto m = momentum(U);
        const double E = total_energy(U);
        const double rho_rho_e = rho * E - 0.5 * norm_square<dim>(m);
        return orc::pow(rho_rho_e, gamma_plus_one_inverse);
      }

      state_type harten_entropy_derivative(const state_type &U) const
      {
        const double rho = density(U);
        const auto m = momentum(U);
        const double E = total_energy(U);
        const double rho_rho_e = rho * E - 0.5 * norm_square<dim>(m);
        const double factor =
            gamma_plus_one_inverse *
            orc::pow(rho_rho_e, -gamma * gamma_plus_one_inverse);
        state_type result;
        result[0] = factor * E;
        for (int d = 0; d < dim; ++d)
          result[1 + d] = -factor * m[d];
        result[dim + 1] = factor * rho;
        return result;
      }

      double mathematical_entropy(const state_type &U) const
      {
        return orc::pow(pressure(U), gamma_inverse);
      }

      state_type mathematical_entropy_derivative(const state_type &U) const
      {
        const double rho_inverse = 1. / density(U);
        Vec<dim> u = momentum(U);
        for (int d = 0; d < dim; ++d)
          u[d] *= rho_inverse;
        const double p = pressure(U);
        const double factor =
            (gamma - 1.0) * gamma_inverse * orc::pow(p, gamma_inverse - 1.);
        state_type result;
        result[0] = factor * 0.5 * norm_square<dim>(u);
        result[dim + 1] = factor;
        for (int d = 0; d < dim; ++d)
          result[1 + d] = -factor * u[d];
        return result;
      }

      bool is_admissible(const state_type &U) const
      {
        return density(U) > 0. && internal_energy(U) > 0. &&
               specific_entropy(U) > 0.;
      }

      /* flux_contribution(pv, ipv, i, U_i) = f(U_i), hyperbolic_system.h:1186-1205 */
      using flux_contribution_type = std::array<Vec<dim>, k>;
      flux_contribution_type
      flux_contribution(const std::array<double, n_precomputed> &, const double, const state_type &U) const
      {
        return f(U);
      }

      precomputed_type precompute(const state_type &U) const
      {
        return {{specific_entropy(U), harten_entropy(U)}};
      }

      flux_type f(const state_type &U) const
      {
        const double rho_inverse = 1. / density(U);
        const auto m = momentum(U);
        const double p = pressure(U);
        const double E = total_energy(U);

        flux_type result;
        result[0] = m;
        for (int i = 0; i < dim; ++i) {
          for (int d = 0; d < dim; ++d)
            result[1 + i][d] = m[d] * (m[i] * rho_inverse);
          result[1 + i][i] += p;
        }
        for (int d = 0; d < dim; ++d)
          result[dim + 1][d] = m[d] * (rho_inverse * (E + p));
        return result;
      }

      /* -contract(add(flux_i, flux_j), c_ij) */
      static state_type flux_divergence(const flux_type &flux_i,
                                        const flux_type &flux_j,
                                        const Vec<dim> &c_ij)
      {
        state_type result;
        for (int c = 0; c < k; ++c) {
          Vec<dim> sum;
          for (int d = 0; d < dim; ++d)
            sum[d] = flux_i[c][d] + flux_j[c][d];
          result[c] = -dot<dim>(sum, c_ij);
        }
        return result;
      }

      /* 1D primitive (rho, u, p) embedded into dim dimensions -> conserved */
      state_type from_initial_state(const std::array<double, 3> &prim) const
      {
        state_type primitive{};
        primitive[0] = prim[0];
        primitive[1] = prim[1];
        primitive[dim + 1] = prim[2];
        return from_primitive_state(primitive);
      }

      state_type from_primitive_state(const state_type &primitive) const
      {
        const double rho = primitive[0];
        const auto u = momentum(primitive);
        const double p = primitive[dim + 1];
        state_type state = primitive;
        for (int d = 0; d < dim; ++d)
          state[1 + d] *= rho;
        state[dim + 1] =
            p / (gamma - 1.) + 0.5 * rho * norm_square<dim>(u);
        return state;
      }

      template <int component>
      state_type prescribe_riemann_characteristic(const state_type &U,
                                                  const state_type &U_bar,
                                                  const Vec<dim> &normal) const
      {
        const auto m = momentum(U);
        const double rho = density(U);
        const double a = speed_of_sound(U);
        const double vn = dot<dim>(m, normal) / rho;

        const auto m_bar = momentum(U_bar);
        const double rho_bar = density(U_bar);
        const double a_bar = speed_of_sound(U_bar);
        const double vn_bar = dot<dim>(m_bar, normal) / rho_bar;

        const double R_1 = component == 1
                               ? vn_bar - 2. * a_bar / (gamma - 1.)
                               : vn - 2. * a / (gamma - 1.);
        const double R_2 = component == 2
                               ? vn_bar + 2. * a_bar / (gamma - 1.)
                               : vn + 2. * a / (gamma - 1.);

        const double p = pressure(U);
        const double s = p / orc::pow(rho, gamma);

        const double rho_inverse = 1. / rho; /* m / rho: Tensor / scalar */
        Vec<dim> vperp;
        for (int d = 0; d < dim; ++d)
          vperp[d] = m[d] * rho_inverse - vn * normal[d];

        const double vn_new = 0.5 * (R_1 + R_2);

        const double tmp = ((gamma - 1.) / 4.) * (R_2 - R_1);
        double rho_new = 1. / (gamma * s) * (tmp * tmp);
        rho_new = orc::pow(rho_new, 1. / (gamma - 1.));

        const double p_new = s * std::pow(rho_new, gamma);

        state_type U_new;
        U_new[0] = rho_new;
        for (int d = 0; d < dim; ++d)
          U_new[1 + d] = rho_new * (vn_new * normal[d] + vperp[d]);
        U_new[1 + dim] =
            p_new / (gamma - 1.) +
            0.5 * rho_new * (vn_new * vn_new + norm_square<dim>(vperp));
        return U_new;
      }

      template <typename Lambda>
      state_type apply_boundary_conditions(const int id,
                                           const state_type &U,
                                           const Vec<dim> &normal,
                                           const Lambda &get_dirichlet_data) const
      {
        state_type result = U;

        if (id == Boundary::dirichlet) {
          result = get_dirichlet_data();

        } else if (id == Boundary::slip) {
          auto m = momentum(U);
          const double mn = dot<dim>(m, normal);
          for (int d = 0; d < dim; ++d)
            m[d] -= 1. * mn * normal[d];
          for (int d = 0; d < dim; ++d)
            result[1 + d] = m[d];

        } else if (id == Boundary::no_slip) {
          for (int d = 0; d < dim; ++d)
            result[1 + d] = 0.;

        } else if (id == Boundary::dynamic) {
          const auto m = momentum(U);
          const double rho = density(U);
          const double a = speed_of_sound(U);
          const double vn = dot<dim>(m, normal) / rho;

          if (vn < -a)
            result = get_dirichlet_data();

          if (vn >= -a && vn <= 0.) {
            const auto U_dirichlet = get_dirichlet_data();
            result = prescribe_riemann_characteristic<2>(U_dirichlet, U, normal);
          }

          if (vn > 0. && vn <= a) {
            const auto U_dirichlet = get_dirichlet_data();
            result = prescribe_riemann_characteristic<1>(U, U_dirichlet, normal);
          }
        }
        return result;
      }
    };


    /* Intermediate values printed by the reference's DEBUG_RIEMANN_SOLVER
     * build (tests/euler/riemann_solver.cc) so that goldens pin every stage. */
    struct RiemannTrace {
      double p_star_two_rarefaction = 0., p_star_failsafe = 0., p_2 = 0.,
             phi_p_2 = 0., lambda_max = 0.;
    };


    template <int dim>
    struct RiemannSolver {
      using Sys = System<dim>;
      using state_type = typename Sys::state_type;
      using primitive_type = std::array<double, 4>; /* rho, u, p, a */

      const Sys &sys;
      explicit RiemannSolver(const Sys &s)
          : sys(s)
      {
      }

      double f(const primitive_type &rd, const double p_star) const
      {
        const double gamma = sys.gamma;
        const double rho = rd[0], p = rd[2], a = rd[3];

        const double Az = 2. / (rho * (gamma + 1.));
        const double Bz = (gamma - 1.) / (gamma + 1.) * p;
        const double radicand = Az / (p_star + Bz);
        const double true_value = (p_star - p) * std::sqrt(radicand);

        const double exponent = 0.5 * (gamma - 1.) / gamma;
        const double factor = orc::pow(p_star / p, exponent) - 1.;
        const double false_value = 2. * a * factor / (gamma - 1.);

        return p_star >= p ? true_value : false_value;
      }

      double df(const primitive_type &rd, const double p_star) const
      {
        const double gamma = sys.gamma;
        const double rho = rd[0], p = rd[2], a = rd[3];

        const double radicand_inverse =
            0.5 * rho * ((gamma + 1.) * p_star + (gamma - 1.) * p);
        const double denominator =
            (p_star + (gamma - 1.) * sys.gamma_plus_one_inverse * p);
        const double true_value = (denominator - 0.5 * (p_star - p)) /
                                  (denominator * std::sqrt(radicand_inverse));

        const double exponent = (-1. - gamma) * 0.5 * sys.gamma_inverse;
        const double factor = (gamma - 1.) * 0.5 * sys.gamma_inverse *
                              orc::pow(p_star / p, exponent) / p;
        const double false_value =
            factor * 2. * a * sys.gamma_minus_one_inverse;

        return p_star >= p ? true_value : false_value;
      }

      double phi(const primitive_type &rd_i,
                 const primitive_type &rd_j,
                 const double p_in) const
      {
        return f(rd_i, p_in) + f(rd_j, p_in) + rd_j[1] - rd_i[1];
      }

      double dphi(const primitive_type &rd_i,
                  const primitive_type &rd_j,
                  const double p) const
      {
        return df(rd_i, p) + df(rd_j, p);
      }

      double phi_of_p_max(const primitive_type &rd_i,
                          const primitive_type &rd_j) const
      {
        const double gamma = sys.gamma;
        const double rho_i = rd_i[0], u_i = rd_i[1], p_i = rd_i[2];
        const double rho_j = rd_j[0], u_j = rd_j[1], p_j = rd_j[2];

        const double p_max = std::max(p_i, p_j);

        const double radicand_inverse_i =
            0.5 * rho_i * ((gamma + 1.) * p_max + (gamma - 1.) * p_i);
        const double value_i = (p_max - p_i) / std::sqrt(radicand_inverse_i);

        const double radicand_inverse_j =
            0.5 * rho_j * ((gamma + 1.) * p_max + (gamma - 1.) * p_j);
        const double value_j = (p_max - p_j) / std::sqrt(radicand_inverse_j);

        return value_i + value_j + u_j - u_i;
      }

      double lambda1_minus(const primitive_type &rd, const double p_star) const
      {
        const double factor = (sys.gamma + 1.0) * 0.5 * sys.gamma_inverse;
        const double u = rd[1], p = rd[2], a = rd[3];
        const double inv_p = 1.0 / p;
        const double tmp = positive_part((p_star - p) * inv_p);
        return u - a * std::sqrt(1.0 + factor * tmp);
      }

      double lambda3_plus(const primitive_type &rd, const double p_star) const
      {
        const double factor = (sys.gamma + 1.0) * 0.5 * sys.gamma_inverse;
        const double u = rd[1], p = rd[2], a = rd[3];
        const double inv_p = 1.0 / p;
        const double tmp = positive_part((p_star - p) * inv_p);
        return u + a * std::sqrt(1.0 + factor * tmp);
      }

      std::array<double, 2> compute_gap(const primitive_type &rd_i,
                                        const primitive_type &rd_j,
                                        const double p_1,
                                        const double p_2) const
      {
        const double nu_11 = lambda1_minus(rd_i, p_2);
        const double nu_12 = lambda1_minus(rd_i, p_1);
        const double nu_31 = lambda3_plus(rd_j, p_1);
        const double nu_32 = lambda3_plus(rd_j, p_2);

        const double lambda_max =
            std::max(positive_part(nu_32), negative_part(nu_11));
        const double gap =
            std::max(std::abs(nu_32 - nu_31), std::abs(nu_12 - nu_11));
        return {{gap, lambda_max}};
      }

      double compute_lambda(const primitive_type &rd_i,
                            const primitive_type &rd_j,
                            const double p_star) const
      {
        const double nu_11 = lambda1_minus(rd_i, p_star);
        const double nu_32 = lambda3_plus(rd_j, p_star);
        return std::max(positive_part(nu_32), negative_part(nu_11));
      }

      double p_star_two_rarefaction(const primitive_type &rd_i,
                                    const primitive_type &rd_j) const
      {
        const double gamma = sys.gamma;
        const double u_i = rd_i[1], p_i = rd_i[2], a_i = rd_i[3];
        const double u_j = rd_j[1], p_j = rd_j[2], a_j = rd_j[3];
        const double inv_p_j = 1. / p_j;

        const double factor = (gamma - 1.) * 0.5;

        const double numerator =
            positive_part(a_i + a_j - factor * (u_j - u_i));
        const double denominator =
            a_i * orc::pow(p_i * inv_p_j, -factor * sys.gamma_inverse) + a_j;

        const double exponent = 2.0 * gamma * sys.gamma_minus_one_inverse;

        return p_j * orc::pow(numerator / denominator, exponent);
      }

      double p_star_failsafe(const primitive_type &rd_i,
                             const primitive_type &rd_j) const
      {
        const double gamma = sys.gamma;
        const double rho_i = rd_i[0], u_i = rd_i[1], p_i = rd_i[2];
        const double rho_j = rd_j[0], u_j = rd_j[1], p_j = rd_j[2];

        const double p_max = std::max(p_i, p_j);

        double radicand_i = 2. * p_max;
        radicand_i /= rho_i * ((gamma + 1.) * p_max + (gamma - 1.) * p_i);
        const double x_i = std::sqrt(radicand_i);

        double radicand_j = 2. * p_max;
        radicand_j /= rho_j * ((gamma + 1.) * p_max + (gamma - 1.) * p_j);
        const double x_j = std::sqrt(radicand_j);

        const double a = x_i + x_j;
        const double b = u_j - u_i;
        const double c = -p_i * x_i - p_j * x_j;

        const double base = (-b + std::sqrt(b * b - 4. * a * c)) / (2. * a);
        return base * base;
      }

      primitive_type riemann_data_from_state(const state_type &U,
                                             const Vec<dim> &n_ij) const
      {
        const double rho = Sys::density(U);
        const double rho_inverse = 1.0 / rho;

        const auto m = Sys::momentum(U);
        const double proj_m = dot<dim>(n_ij, m);
        Vec<dim> perp;
        for (int d = 0; d < dim; ++d)
          perp[d] = m[d] - proj_m * n_ij[d];

        const double E =
            Sys::total_energy(U) - 0.5 * norm_square<dim>(perp) * rho_inverse;

        /* 1D pressure and speed of sound of (rho, proj_m, E): */
        const double rho_e_1d = E - 0.5 * (proj_m * proj_m) * (1. / rho);
        const double p = (sys.gamma - 1.) * rho_e_1d;
        const double a = std::sqrt(sys.gamma * p * (1. / rho));
        return {{rho, proj_m * rho_inverse, p, a}};
      }

      double compute(const primitive_type &rd_i,
                     const primitive_type &rd_j,
                     RiemannTrace *trace = nullptr) const
      {
        const double p_i = rd_i[2], p_j = rd_j[2];
        const double p_max = std::max(p_i, p_j);

        const double rarefaction = p_star_two_rarefaction(rd_i, rd_j);
        const double failsafe = p_star_failsafe(rd_i, rd_j);
        const double p_star_tilde = std::min(rarefaction, failsafe);

        const double phi_p_max = phi_of_p_max(rd_i, rd_j);

        double p_2 =
            phi_p_max < 0. ? p_star_tilde : std::min(p_max, p_star_tilde);

        if (trace) {
          trace->p_star_two_rarefaction = rarefaction;
          trace->p_star_failsafe = failsafe;
          trace->p_2 = p_2;
          trace->phi_p_2 = phi(rd_i, rd_j, p_2);
        }

        if (sys.prm.riemann_newton_max_iterations == 0) {
          const double lambda_max = compute_lambda(rd_i, rd_j, p_2);
          if (trace)
            trace->lambda_max = lambda_max;
          return lambda_max;
        }

        const double p_min = std::min(p_i, p_j);
        double p_1 = phi_p_max < 0. ? p_max : p_min;
        p_1 = p_1 <= p_2 ? p_1 : p_2;

        auto gl = compute_gap(rd_i, rd_j, p_1, p_2);
        double gap = gl[0], lambda_max = gl[1];

        for (unsigned int i = 0; i < sys.prm.riemann_newton_max_iterations;
             ++i) {
          const double tolerance = sys.prm.riemann_newton_tolerance;
          if (std::max(0., gap - tolerance) == 0.)
            break;

          const double phi_p_1 = phi(rd_i, rd_j, p_1);
          const double phi_p_2 = phi(rd_i, rd_j, p_2);
          const double dphi_p_1 = dphi(rd_i, rd_j, p_1);
          const double dphi_p_2 = dphi(rd_i, rd_j, p_2);

          quadratic_newton_step(p_1, p_2, phi_p_1, phi_p_2, dphi_p_1, dphi_p_2);

          gl = compute_gap(rd_i, rd_j, p_1, p_2);
          gap = gl[0];
          lambda_max = gl[1];
        }

        if (trace)
          trace->lambda_max = lambda_max;
        return lambda_max;
      }

      double compute(const state_type &U_i,
                     const state_type &U_j,
                     const Vec<dim> &n_ij) const
      {
        const auto rd_i = riemann_data_from_state(U_i, n_ij);
        const auto rd_j = riemann_data_from_state(U_j, n_ij);
        return compute(rd_i, rd_j);
      }
    };


    template <int dim>
    struct Indicator {
      using Sys = System<dim>;
      using state_type = typename Sys::state_type;
      using flux_type = typename Sys::flux_type;
      using precomputed_type = typename Sys::precomputed_type;

      const Sys &sys;
      double rho_i_inverse = 0., eta_i = 0.;
      flux_type f_i;
      state_type d_eta_i;
      double left = 0.;
      state_type right;

      explicit Indicator(const Sys &s)
          : sys(s)
      {
      }

      void reset(const precomputed_type &prec_i, const state_type &U_i)
      {
        const double rho_i = Sys::density(U_i);
        rho_i_inverse = 1. / rho_i;
        eta_i = prec_i[1];

        d_eta_i = sys.harten_entropy_derivative(U_i);
        d_eta_i[0] -= eta_i * rho_i_inverse;
        f_i = sys.f(U_i);

        left = 0.;
        right.fill(0.);
      }

      void accumulate(const precomputed_type &prec_j,
                      const state_type &U_j,
                      const Vec<dim> &c_ij)
      {
        const double eta_j = prec_j[1];
        const double rho_j = Sys::density(U_j);
        const double rho_j_inverse = 1. / rho_j;

        const auto m_j = Sys::momentum(U_j);
        const auto f_j = sys.f(U_j);

        const double entropy_flux =
            (eta_j * rho_j_inverse - eta_i * rho_i_inverse) *
            dot<dim>(m_j, c_ij);

        left += entropy_flux;
        for (int c = 0; c < Sys::k; ++c) {
          Vec<dim> diff;
          for (int d = 0; d < dim; ++d)
            diff[d] = f_j[c][d] - f_i[c][d];
          right[c] += dot<dim>(diff, c_ij);
        }
      }

      double alpha(const double hd_i) const
      {
        double numerator = left;
        double denominator = std::abs(left);
        for (int c = 0; c < Sys::k; ++c) {
          numerator -= d_eta_i[c] * right[c];
          denominator += std::abs(d_eta_i[c] * right[c]);
        }
        const double quotient =
            std::abs(numerator) / (denominator + hd_i * std::abs(eta_i));
        return std::min(1., sys.prm.indicator_evc_factor * quotient);
      }
    };


    /* Per Newton round trace, as printed with DEBUG_OUTPUT_LIMITER. */
    struct LimiterTrace {
      int n_rounds = 0;
      double t_l[8], t_r[8], psi_l[8], psi_r[8];
      double t_l_start = 0., t_r_start = 0.;
    };


    template <int dim>
    struct Limiter {
      using Sys = System<dim>;
      using state_type = typename Sys::state_type;
      using precomputed_type = typename Sys::precomputed_type;
      using Bounds = typename Sys::bounds_type;

      const Sys &sys;
      state_type U_i;
      Bounds bounds_;
      double rho_relaxation_numerator = 0., rho_relaxation_denominator = 0.,
             s_interp_max = 0.;

      explicit Limiter(const Sys &s)
          : sys(s)
      {
      }

      void reset(const state_type &new_U_i)
      {
        U_i = new_U_i;
        bounds_[0] = max_double;
        bounds_[1] = 0.;
        bounds_[2] = max_double;
        rho_relaxation_numerator = 0.;
        rho_relaxation_denominator = 0.;
        s_interp_max = 0.;
      }

      void accumulate(const precomputed_type &prec_j,
                      const state_type &U_j,
                      const Vec<dim> &scaled_c_ij)
      {
        double &rho_min = bounds_[0], &rho_max = bounds_[1], &s_min = bounds_[2];

        const double rho_i = Sys::density(U_i);
        const auto m_i = Sys::momentum(U_i);
        const double rho_j = Sys::density(U_j);
        const auto m_j = Sys::momentum(U_j);
        const double rho_affine_shift = 0.;

        Vec<dim> dm;
        for (int d = 0; d < dim; ++d)
          dm[d] = m_i[d] - m_j[d];

        const double rho_ij_bar =
            0.5 * (rho_i + rho_j + dot<dim>(dm, scaled_c_ij)) +
            rho_affine_shift;

        rho_min = std::min(rho_min, rho_ij_bar);
        rho_max = std::max(rho_max, rho_ij_bar);

        const double s_j = prec_j[0];
        s_min = std::min(s_min, s_j);

        const double beta_ij = 1.;
        rho_relaxation_numerator += beta_ij * (rho_i + rho_j);
        rho_relaxation_denominator += std::abs(beta_ij);

        state_type U_mid;
        for (int c = 0; c < Sys::k; ++c)
          U_mid[c] = (U_i[c] + U_j[c]) * .5;
        const double s_interp = sys.specific_entropy(U_mid);
        s_interp_max = std::max(s_interp_max, s_interp);
      }

      Bounds bounds(const double hd_i) const
      {
        Bounds relaxed = bounds_;
        double &rho_min = relaxed[0], &rho_max = relaxed[1], &s_min = relaxed[2];

        double r_i = std::sqrt(hd_i); /* 3D: ^ 3/6 */
        if constexpr (dim == 2) {
          const double q = std::sqrt(r_i); /* 2D: ^ 3/4 */
          r_i = q * q * q;
        } else if constexpr (dim == 1) {
          r_i = r_i * r_i * r_i; /* 1D: ^ 3/2 */
        }
        r_i *= sys.prm.limiter_relaxation_factor;

        const double rho_relaxation =
            std::abs(rho_relaxation_numerator) /
            (std::abs(rho_relaxation_denominator) + eps_double);

        const double relaxation =
            (2. * sys.prm.limiter_relaxation_factor) * rho_relaxation;

        rho_min = std::max((1. - r_i) * rho_min, rho_min - relaxation);
        rho_max = std::min((1. + r_i) * rho_max, rho_max + relaxation);

        const double entropy_relaxation =
            sys.prm.limiter_relaxation_factor * (s_interp_max - s_min);

        s_min = std::max((1. - r_i) * s_min, s_min - entropy_relaxation);
        return relaxed;
      }

      static Bounds combine_bounds(const Bounds &l, const Bounds &r)
      {
        return {{std::min(l[0], r[0]), std::max(l[1], r[1]),
                 std::min(l[2], r[2])}};
      }

      /* returns t_l; success reports whether low-order state was in bounds */
      double limit(const Bounds &bounds,
                   const state_type &U,
                   const state_type &P,
                   bool &success,
                   const double t_min = 0.,
                   const double t_max = 1.,
                   LimiterTrace *trace = nullptr) const
      {
        const bool expensive = sys.prm.expensive_bounds_check;

        success = true;
        double t_r = t_max;

        const double small = sys.prm.vacuum_state_relaxation_small;
        const double large = sys.prm.vacuum_state_relaxation_large;
        const double relax_small = 1. + small * eps_double;
        const double relax = 1. + large * eps_double;

        auto add = [](const state_type &A, double t, const state_type &B) {
          state_type R;
          for (int c = 0; c < Sys::k; ++c)
            R[c] = A[c] + t * B[c];
          return R;
        };

        {
          const double rho_U = Sys::density(U);
          const double rho_P = Sys::density(P);
          const double rho_min = bounds[0];
          const double rho_max = bounds[1];

          const double test_min = sys.filter_vacuum_density(
              std::max(0., rho_U - relax * rho_max));
          const double test_max = sys.filter_vacuum_density(
              std::max(0., rho_min - relax * rho_U));
          if (!(test_min == 0. && test_max == 0.))
            success = false;

          const double denominator =
              1. / (std::abs(rho_P) + eps_double * rho_max);

          t_r = rho_max < rho_U + t_r * rho_P
                    ? (rho_max - rho_U) * denominator
                    : t_r;

          t_r = rho_U + t_r * rho_P < rho_min
                    ? (rho_U - rho_min) * denominator
                    : t_r;

          t_r = std::min(t_r, t_max);
          t_r = std::max(t_r, t_min);

          if (expensive) {
            const double rho_new = Sys::density(add(U, t_r, P));
            const double test_new_min = sys.filter_vacuum_density(
                std::max(0., rho_new - relax * rho_max));
            const double test_new_max = sys.filter_vacuum_density(
                std::max(0., rho_min - relax * rho_new));
            if (!(test_new_min == 0. && test_new_max == 0.))
              success = false;
          }
        }

        double t_l = t_min;

        const double gamma = sys.gamma;
        const double gp1 = gamma + 1.;

        {
          const double s_min = bounds[2];

          if (trace) {
            trace->t_l_start = t_l;
            trace->t_r_start = t_r;
            trace->n_rounds = 0;
          }

          for (unsigned int n = 0; n < sys.prm.limiter_newton_max_iterations;
               ++n) {
            const auto U_r = add(U, t_r, P);
            const double rho_r = Sys::density(U_r);
            const double rho_r_gamma = orc::pow(rho_r, gamma);
            const double rho_e_r = Sys::internal_energy(U_r);

            const double psi_r =
                relax_small * rho_r * rho_e_r - s_min * rho_r * rho_r_gamma;

            if (!expensive) {
              t_l = psi_r > 0. ? t_r : t_l;
              if (t_l == t_r)
                break;
            }

            const auto U_l = add(U, t_l, P);
            const double rho_l = Sys::density(U_l);
            const double rho_l_gamma = orc::pow(rho_l, gamma);
            const double rho_e_l = Sys::internal_energy(U_l);

            const double psi_l =
                relax_small * rho_l * rho_e_l - s_min * rho_l * rho_l_gamma;

            const double lower_bound =
                (1. - relax) * s_min * rho_l * rho_l_gamma;
            if (n == 0 && !(std::min(0., psi_l - lower_bound) == 0.))
              success = false;

            if (expensive)
              t_l = psi_r > 0. ? t_r : t_l;

            auto record = [&]() {
              if (trace && trace->n_rounds < 8) {
                const int r = trace->n_rounds++;
                trace->t_l[r] = t_l;
                trace->t_r[r] = t_r;
                trace->psi_l[r] = psi_l;
                trace->psi_r[r] = psi_r;
              }
            };

            const double tolerance = sys.prm.limiter_newton_tolerance;
            if (std::max(0., t_r - t_l - tolerance) == 0.) {
              record();
              break;
            }

            const double drho = Sys::density(P);
            const auto de_l = Sys::internal_energy_derivative(U_l);
            const auto de_r = Sys::internal_energy_derivative(U_r);
            double drho_e_l = de_l[0] * P[0];
            double drho_e_r = de_r[0] * P[0];
            for (int c = 1; c < Sys::k; ++c) {
              drho_e_l += de_l[c] * P[c];
              drho_e_r += de_r[c] * P[c];
            }
            const double dpsi_l =
                rho_l * drho_e_l + (rho_e_l - gp1 * s_min * rho_l_gamma) * drho;
            const double dpsi_r =
                rho_r * drho_e_r + (rho_e_r - gp1 * s_min * rho_r_gamma) * drho;

            quadratic_newton_step(t_l, t_r, psi_l, psi_r, dpsi_l, dpsi_r, -1.);
            record();
          }

          if (expensive) {
            const auto U_new = add(U, t_l, P);
            const double rho_new = Sys::density(U_new);
            const double rho_new_gamma = orc::pow(rho_new, gamma);
            const double rho_e_new = Sys::internal_energy(U_new);
            const double psi_new = relax_small * rho_new * rho_e_new -
                                   s_min * rho_new * rho_new_gamma;
            const double lower_bound =
                (1. - relax) * s_min * rho_new * rho_new_gamma;
            const bool e_valid = std::min(0., rho_e_new) == 0.;
            const bool psi_valid = std::min(0., psi_new - lower_bound) == 0.;
            if (!e_valid || !psi_valid)
              success = false;
          }
        }

        return t_l;
      }
    };
  } // namespace euler
} // namespace orc
