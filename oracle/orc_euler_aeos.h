// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_common.h).
//
// Euler equations with an arbitrary / tabulated equation of state ("euler aeos"), scalar double
// semantics.  Restates
//   System / View    <- source/euler_aeos/hyperbolic_system.h:843-1528
//   EquationOfState  <- source/euler_aeos/equation_of_state_{polytropic_gas,noble_abel_stiffened_gas,
//                       van_der_waals}.h
//   RiemannSolver    <- source/euler_aeos/riemann_solver.template.h:23-676
//   Indicator        <- source/euler_aeos/indicator.h:165-263
//   Limiter          <- source/euler_aeos/limiter.h:244-445, limiter.template.h:15-361
// Differences to the polytropic Euler module that matter to the driver (orc_module.h):
// 4 precomputed values (p, gamma_min, s, eta) in TWO precomputation cycles -- cycle 0 per node,
// cycle 1 with a loop over the stencil --, the pressure in the flux comes from the precomputed
// vector, the Riemann solver reads p_i, p_j, the limiter keeps 4 bounds (rho_min, rho_max, s_min,
// gamma_min).
#pragma once

#include "orc_common.h"

#include <stdexcept>

namespace orc
{
  namespace euler_aeos
  {
    enum class Eos : int { polytropic_gas = 0, noble_abel_stiffened_gas = 1, van_der_waals = 2 };

    struct Parameters {
      /* equation of state */
      Eos eos = Eos::polytropic_gas;
      double gamma = 7. / 5.;
      double covolume_b = 0.;       /* NASG, vdW */
      double reference_pressure = 0.;                 /* NASG pinf */
      double reference_specific_internal_energy = 0.; /* NASG q */
      double vdw_a = 0.;
      /* the interpolatory parameters may also be set directly (the reference's unit tests select
       * "van der waals" only to set a covolume) */
      bool compute_strict_bounds = true;
      double reference_density = 1.;
      double vacuum_state_relaxation_small = 1.e2;
      double vacuum_state_relaxation_large = 1.e4;

      double indicator_evc_factor = 1.;

      unsigned int limiter_iterations = 2;
      double limiter_newton_tolerance = 1.e-10;
      unsigned int limiter_newton_max_iterations = 2;
      double limiter_relaxation_factor = 1.;
      bool expensive_bounds_check = false; /* compile time option upstream */
    };


    template <int dim>
    struct System {
      static constexpr int k = dim + 2;
      static constexpr int n_precomputed = 4; /* (p, gamma_min, s, eta) */
      static constexpr int n_bounds = 4;      /* rho_min, rho_max, s_min, gamma_min */
      static constexpr int n_initial_precomputed = 0;
      static constexpr int n_precomputation_cycles = 2;
      static constexpr bool shallow_water = false;
      static constexpr bool aeos = true;
      static constexpr bool have_source_terms = false;
      static constexpr bool have_high_order_flux = false;
      using state_type = std::array<double, k>;
      using flux_type = std::array<Vec<dim>, k>;
      using flux_contribution_type = flux_type;
      using precomputed_type = std::array<double, n_precomputed>;
      using bounds_type = std::array<double, n_bounds>;

      Parameters prm;
      double b = 0., pinf = 0., q = 0.; /* eos_interpolation_{b, pinfty, q} */
      double gamma = 7. / 5.;           /* read by the shared initial state library only */

      explicit System(const Parameters &p = Parameters())
          : prm(p)
      {
        gamma = prm.gamma;
        /* parse_parameters_call_back of the selected equation of state */
        if (prm.eos == Eos::noble_abel_stiffened_gas) {
          b = prm.covolume_b;
          pinf = prm.reference_pressure;
          q = prm.reference_specific_internal_energy;
        } else if (prm.eos == Eos::van_der_waals) {
          b = prm.covolume_b;
          if (prm.covolume_b > 0.)
            pinf = prm.vdw_a / (prm.covolume_b * prm.covolume_b);
        }
      }

      /* ---------------- equation of state ---------------- */

      double eos_pressure(const double rho, const double e) const
      {
        const double gamma = prm.gamma;
        switch (prm.eos) {
        case Eos::polytropic_gas: return (gamma - 1.) * rho * e;
        case Eos::noble_abel_stiffened_gas:
          return (gamma - 1.) * rho * (e - prm.reference_specific_internal_energy) / (1. - prm.covolume_b * rho) -
                 gamma * prm.reference_pressure;
        case Eos::van_der_waals: {
          const double intermolecular = prm.vdw_a * rho * rho;
          const double numerator = rho * e + intermolecular;
          const double covolume = 1. - prm.covolume_b * rho;
          return (gamma - 1.) * numerator / covolume - intermolecular;
        }
        }
        return 0.;
      }

      double eos_specific_internal_energy(const double rho, const double p) const
      {
        const double gamma = prm.gamma;
        switch (prm.eos) {
        case Eos::polytropic_gas: return p / (rho * (gamma - 1.));
        case Eos::noble_abel_stiffened_gas: {
          const double numerator = (p + gamma * prm.reference_pressure) * (1. - prm.covolume_b * rho);
          const double denominator = rho * (gamma - 1.);
          return prm.reference_specific_internal_energy + numerator / denominator;
        }
        case Eos::van_der_waals: {
          const double intermolecular = prm.vdw_a * rho * rho;
          const double covolume = 1. - prm.covolume_b * rho;
          const double numerator = (p + intermolecular) * covolume;
          const double denominator = rho * (gamma - 1.);
          return numerator / denominator - prm.vdw_a * rho;
        }
        }
        return 0.;
      }

      /* ---------------- state ---------------- */

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
        const double cutoff = prm.reference_density * prm.vacuum_state_relaxation_large * eps_double;
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

      double surrogate_specific_entropy(const state_type &U, const double gamma_min) const
      {
        const double rho = density(U);
        const double rho_inverse = 1. / rho;
        const double covolume = 1. - b * rho;
        const double shift = internal_energy(U) - rho * q - pinf * covolume;
        return shift * orc::pow(rho_inverse - b, gamma_min) / covolume;
      }

      double surrogate_harten_entropy(const state_type &U, const double gamma_min) const
      {
        const double rho = density(U);
        const auto m = momentum(U);
        const double E = total_energy(U);
        const double rho_rho_e_q = rho * E - 0.5 * norm_square<dim>(m) - rho * rho * q;
        const double exponent = 1. / (gamma_min + 1.);
        const double covolume = 1. - b * rho;
        const double covolume_term = orc::pow(covolume, gamma_min - 1.);
        const double rho_pinfcov = rho * pinf * covolume;
        return orc::pow((rho_rho_e_q - rho_pinfcov) * covolume_term, exponent);
      }

      state_type
      surrogate_harten_entropy_derivative(const state_type &U, const double eta, const double gamma_min) const
      {
        const double rho = density(U);
        const auto m = momentum(U);
        const double E = total_energy(U);
        const double covolume = 1. - b * rho;
        const double covolume_inverse = 1. / covolume;
        const double shift = rho * E - 0.5 * norm_square<dim>(m) - rho * rho * q - rho * pinf * covolume;
        const double factor =
            orc::pow(eta * covolume_inverse, -gamma_min) * (covolume_inverse * covolume_inverse) / (gamma_min + 1.);
        state_type result;
        const double first_term = E - 2. * rho * q - pinf * (1. - 2. * b * rho);
        const double second_term = -(gamma_min - 1.) * shift * b;
        result[0] = factor * (covolume * first_term + second_term);
        for (int d = 0; d < dim; ++d)
          result[1 + d] = -factor * covolume * m[d];
        result[dim + 1] = factor * covolume * rho;
        return result;
      }

      double surrogate_gamma(const state_type &U, const double p) const
      {
        const double rho = density(U);
        const double rho_e = internal_energy(U);
        const double covolume = 1. - b * rho;
        const double numerator = (p + pinf) * covolume;
        const double denominator = rho_e - rho * q - covolume * pinf;
        return 1. + numerator / denominator;
      }

      double surrogate_pressure(const state_type &U, const double gamma) const
      {
        const double rho = density(U);
        const double rho_e = internal_energy(U);
        const double covolume = 1. - b * rho;
        return (gamma - 1.) * (rho_e - rho * q) / covolume - gamma * pinf;
      }

      bool is_admissible(const state_type &U) const
      {
        const double rho = density(U);
        const double rho_e = internal_energy(U);
        const double covolume = 1. - b * rho;
        const double shift = rho_e - rho * q - pinf * covolume;
        return rho > 0. && shift > 0.;
      }

      /* ---------------- precomputation, hyperbolic_system.h:843-975 ---------------- */

      /* cycle 0: (p_i, gamma_i, 0, 0) */
      precomputed_type precompute_cycle_0(const state_type &U_i) const
      {
        const double rho_i = density(U_i);
        const double e_i = internal_energy(U_i) / rho_i;
        const double p_i = eos_pressure(rho_i, e_i);
        const double gamma_i = surrogate_gamma(U_i, p_i);
        return {{p_i, gamma_i, 0., 0.}};
      }

      /* cycle 1: gamma_min over the stencil (the caller loops), then s_i, eta_i */
      void precompute_cycle_1_finish(precomputed_type &prec_i, const state_type &U_i) const
      {
        prec_i[2] = surrogate_specific_entropy(U_i, prec_i[1]);
        prec_i[3] = surrogate_harten_entropy(U_i, prec_i[1]);
      }

      /* ---------------- flux ---------------- */

      flux_type f(const state_type &U, const double p) const
      {
        const double rho_inverse = 1. / density(U);
        const auto m = momentum(U);
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

      flux_contribution_type flux_contribution(const precomputed_type &prec, const double, const state_type &U) const
      {
        return f(U, prec[0]);
      }

      static state_type flux_divergence(const flux_type &flux_i, const flux_type &flux_j, const Vec<dim> &c_ij)
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

      /* 1D primitive (rho, u, p) -> conserved, through the equation of state */
      state_type from_initial_state(const std::array<double, 3> &prim) const
      {
        state_type primitive{};
        primitive[0] = prim[0];
        primitive[1] = prim[1];
        const double rho = prim[0];
        const double p = prim[2];
        primitive[dim + 1] = eos_specific_internal_energy(rho, p);
        return from_primitive_state(primitive);
      }

      /* primitive = (rho, u, e) */
      state_type from_primitive_state(const state_type &primitive) const
      {
        const double rho = primitive[0];
        const auto u = momentum(primitive);
        const double e = primitive[dim + 1];
        state_type state = primitive;
        for (int d = 0; d < dim; ++d)
          state[1 + d] *= rho;
        state[dim + 1] = rho * e + 0.5 * rho * norm_square<dim>(u);
        return state;
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
          throw std::runtime_error("euler aeos: dynamic boundary conditions trap in the reference");
        }
        return result;
      }
    };


    /* ------------------------------------------------------------------ */

    template <int dim>
    struct RiemannSolver {
      using Sys = System<dim>;
      using state_type = typename Sys::state_type;
      using precomputed_type = typename Sys::precomputed_type;
      using primitive_type = std::array<double, 5>; /* rho, u, p, gamma, a */

      const Sys &sys;
      explicit RiemannSolver(const Sys &s)
          : sys(s)
      {
      }

      static double c(const double gamma)
      {
        constexpr double slope = -0.34976871477801828189920753948709;
        const double first_radicand = (3. * gamma + 11.) / (6. * gamma + 6.);
        const double second_radicand = 5. / 6. + slope * (gamma - 3.);
        double radicand = std::min(first_radicand, second_radicand);
        radicand = std::min(1., radicand);
        radicand = std::max(1. / 2., radicand);
        return std::sqrt(radicand);
      }

      double alpha(const double rho, const double gamma, const double a) const
      {
        const double numerator = 2. * a * (1. - sys.b * rho);
        const double denominator = gamma - 1.;
        return numerator / denominator;
      }

      double p_star_RS_full(const primitive_type &ri, const primitive_type &rj) const
      {
        const double pinf = sys.pinf;
        const double rho_i = ri[0], u_i = ri[1], p_i = ri[2], gamma_i = ri[3], a_i = ri[4];
        const double rho_j = rj[0], u_j = rj[1], p_j = rj[2], gamma_j = rj[3], a_j = rj[4];
        const double alpha_i = alpha(rho_i, gamma_i, a_i);
        const double alpha_j = alpha(rho_j, gamma_j, a_j);

        const double p_min = std::min(p_i, p_j);
        const double p_max = std::max(p_i, p_j);
        const double gamma_min = p_i < p_j ? gamma_i : gamma_j;
        const double alpha_min = p_i < p_j ? alpha_i : alpha_j;
        const double alpha_hat_min = c(gamma_min) * alpha_min;
        const double alpha_max = p_i >= p_j ? alpha_i : alpha_j;
        const double gamma_m = std::min(gamma_i, gamma_j);
        const double gamma_M = std::max(gamma_i, gamma_j);

        const double numerator = positive_part(alpha_hat_min + alpha_max - (u_j - u_i));
        const double p_ratio = (p_min + pinf) / (p_max + pinf);
        const double r_exponent = (gamma_M - gamma_min) / (2. * gamma_min * gamma_M);

        const double first_exponent = (gamma_M - 1.) / (2. * gamma_M);
        const double first_exponent_inverse = 1. / first_exponent;
        const double first_denom = alpha_hat_min * orc::pow(p_ratio, r_exponent - first_exponent) + alpha_max;
        const double p_1_tilde = (p_max + pinf) * orc::pow(numerator / first_denom, first_exponent_inverse) - pinf;

        const double second_exponent = (gamma_m - 1.) / (2. * gamma_m);
        const double second_exponent_inverse = 1. / second_exponent;
        const double second_denom =
            alpha_hat_min * orc::pow(p_ratio, -second_exponent) + alpha_max * orc::pow(p_ratio, r_exponent);
        const double p_2_tilde = (p_max + pinf) * orc::pow(numerator / second_denom, second_exponent_inverse) - pinf;
        return std::min(p_1_tilde, p_2_tilde);
      }

      double p_star_failsafe(const primitive_type &ri, const primitive_type &rj) const
      {
        const double b = sys.b, pinf = sys.pinf;
        const double rho_i = ri[0], u_i = ri[1], p_i = ri[2], gamma_i = ri[3];
        const double rho_j = rj[0], u_j = rj[1], p_j = rj[2], gamma_j = rj[3];
        const double p_max = std::max(p_i, p_j) + pinf;

        double radicand_i = 2. * (1. - b * rho_i) * p_max;
        radicand_i /= rho_i * ((gamma_i + 1.) * p_max + (gamma_i - 1.) * (p_i + pinf));
        const double x_i = std::sqrt(radicand_i);

        double radicand_j = 2. * (1. - b * rho_j) * p_max;
        radicand_j /= rho_j * ((gamma_j + 1.) * p_max + (gamma_j - 1.) * (p_j + pinf));
        const double x_j = std::sqrt(radicand_j);

        const double a = x_i + x_j;
        const double bb = u_j - u_i;
        const double cc = -(p_i + pinf) * x_i - (p_j + pinf) * x_j;
        const double base = (-bb + std::sqrt(bb * bb - 4. * a * cc)) / (2. * a);
        return base * base - pinf;
      }

      double p_star_SS_full(const primitive_type &ri, const primitive_type &rj) const
      {
        const double pinf = sys.pinf;
        const double rho_i = ri[0], u_i = ri[1], p_i = ri[2], gamma_i = ri[3], a_i = ri[4];
        const double rho_j = rj[0], u_j = rj[1], p_j = rj[2], gamma_j = rj[3], a_j = rj[4];
        const double gamma_m = std::min(gamma_i, gamma_j);
        const double alpha_hat_i = c(gamma_i) * alpha(rho_i, gamma_i, a_i);
        const double alpha_hat_j = c(gamma_j) * alpha(rho_j, gamma_j, a_j);
        const double exponent = (gamma_m - 1.) / (2. * gamma_m);
        const double exponent_inverse = 1. / exponent;
        const double numerator = positive_part(alpha_hat_i + alpha_hat_j - (u_j - u_i));
        const double denominator = alpha_hat_i * orc::pow((p_i + pinf) / (p_j + pinf), -exponent) + alpha_hat_j;
        const double p_1_tilde = (p_j + pinf) * orc::pow(numerator / denominator, exponent_inverse) - pinf;
        const double p_2_tilde = p_star_failsafe(ri, rj);
        return std::min(p_1_tilde, p_2_tilde);
      }

      double p_star_interpolated(const primitive_type &ri, const primitive_type &rj) const
      {
        const double pinf = sys.pinf;
        const double rho_i = ri[0], u_i = ri[1], p_i = ri[2], gamma_i = ri[3], a_i = ri[4];
        const double rho_j = rj[0], u_j = rj[1], p_j = rj[2], gamma_j = rj[3], a_j = rj[4];
        const double alpha_i = alpha(rho_i, gamma_i, a_i);
        const double alpha_j = alpha(rho_j, gamma_j, a_j);
        const double p_min = std::min(p_i, p_j) + pinf;
        const double p_max = std::max(p_i, p_j) + pinf;
        const double gamma_min = p_i < p_j ? gamma_i : gamma_j;
        const double alpha_min = p_i < p_j ? alpha_i : alpha_j;
        const double alpha_hat_min = c(gamma_min) * alpha_min;
        const double gamma_max = p_i >= p_j ? gamma_i : gamma_j;
        const double alpha_max = p_i >= p_j ? alpha_i : alpha_j;
        const double alpha_hat_max = c(gamma_max) * alpha_max;
        const double gamma_m = std::min(gamma_i, gamma_j);
        const double gamma_M = std::max(gamma_i, gamma_j);
        const double p_ratio = p_min / p_max;
        const double r_exponent = (gamma_M - gamma_min) / (2. * gamma_min * gamma_M);
        const double exponent = (gamma_m - 1.) / (2. * gamma_m);
        const double exponent_inverse = 1. / exponent;
        const double numerator = positive_part(alpha_hat_min + /*SIC!*/ alpha_max - (u_j - u_i));
        const double denominator =
            alpha_hat_min * orc::pow(p_ratio, -exponent) + alpha_hat_max * orc::pow(p_ratio, r_exponent);
        return p_max * orc::pow(numerator / denominator, exponent_inverse) - pinf;
      }

      double f(const primitive_type &rd, const double p_star) const
      {
        const double b = sys.b, pinf = sys.pinf;
        const double rho = rd[0], p = rd[2], gamma = rd[3], a = rd[4];
        const double one_minus_b_rho = 1. - b * rho;
        const double gamma_minus_one = gamma - 1.;
        const double Az = 2. * one_minus_b_rho / (rho * (gamma + 1.));
        const double Bz = gamma_minus_one / (gamma + 1.) * (p + pinf);
        const double radicand = Az / (p_star + pinf + Bz);
        const double true_value = (p_star - p) * std::sqrt(radicand);
        const double exponent = 0.5 * gamma_minus_one / gamma;
        const double ratio = (p_star + pinf) / (p + pinf);
        const double factor = orc::pow(ratio, exponent) - 1.;
        const double false_value = 2. * a * one_minus_b_rho * factor / gamma_minus_one;
        return p_star >= p ? true_value : false_value;
      }

      double phi(const primitive_type &ri, const primitive_type &rj, const double p_in) const
      {
        return f(ri, p_in) + f(rj, p_in) + rj[1] - ri[1];
      }

      double phi_of_p_max(const primitive_type &ri, const primitive_type &rj) const
      {
        const double b = sys.b, pinf = sys.pinf;
        const double rho_i = ri[0], u_i = ri[1], p_i = ri[2], gamma_i = ri[3];
        const double rho_j = rj[0], u_j = rj[1], p_j = rj[2], gamma_j = rj[3];
        const double p_max = std::max(p_i, p_j) + pinf;
        const double radicand_inverse_i =
            0.5 * rho_i / (1. - b * rho_i) * ((gamma_i + 1.) * p_max + (gamma_i - 1.) * (p_i + pinf));
        const double value_i = (p_max - p_i) / std::sqrt(radicand_inverse_i);
        const double radicand_inverse_j =
            0.5 * rho_j / (1. - b * rho_j) * ((gamma_j + 1.) * p_max + (gamma_j - 1.) * (p_j + pinf));
        const double value_j = (p_max - p_j) / std::sqrt(radicand_inverse_j);
        return value_i + value_j + u_j - u_i;
      }

      double lambda1_minus(const primitive_type &rd, const double p_star) const
      {
        const double u = rd[1], p = rd[2], gamma = rd[3], a = rd[4];
        const double factor = 0.5 * (gamma + 1.) / gamma;
        const double tmp = positive_part((p_star - p) / (p + sys.pinf));
        return u - a * std::sqrt(1. + factor * tmp);
      }

      double lambda3_plus(const primitive_type &rd, const double p_star) const
      {
        const double u = rd[1], p = rd[2], gamma = rd[3], a = rd[4];
        const double factor = 0.5 * (gamma + 1.) / gamma;
        const double tmp = positive_part((p_star - p) / (p + sys.pinf));
        return u + a * std::sqrt(1. + factor * tmp);
      }

      double compute_lambda(const primitive_type &ri, const primitive_type &rj, const double p_star) const
      {
        const double nu_11 = lambda1_minus(ri, p_star);
        const double nu_32 = lambda3_plus(rj, p_star);
        return std::max(positive_part(nu_32), negative_part(nu_11));
      }

      primitive_type riemann_data_from_state(const state_type &U, const double p, const Vec<dim> &n_ij) const
      {
        const double rho = Sys::density(U);
        const double rho_inverse = 1.0 / rho;
        const auto m = Sys::momentum(U);
        const double proj_m = dot<dim>(n_ij, m);
        const double gamma = sys.surrogate_gamma(U, p);
        const double x = 1. - sys.b * rho;
        const double a = std::sqrt(gamma * (p + sys.pinf) / (rho * x));
        return {{rho, proj_m * rho_inverse, p, gamma, a}};
      }

      double compute(const primitive_type &ri, const primitive_type &rj, double *p_2_out = nullptr) const
      {
        const double pinf = sys.pinf;
        const double p_max = std::max(ri[2], rj[2]) + pinf;
        const double phi_p_max = phi_of_p_max(ri, rj);
        double p_2;
        if (!sys.prm.compute_strict_bounds) {
          const double p_star_tilde = p_star_interpolated(ri, rj);
          const double p_star_backup = p_star_failsafe(ri, rj);
          p_2 = phi_p_max < 0. ? std::min(p_star_tilde, p_star_backup) : std::min(p_max, p_star_tilde);
        } else {
          const double p_star_RS = p_star_RS_full(ri, rj);
          const double p_star_SS = p_star_SS_full(ri, rj);
          p_2 = phi_p_max < 0. ? p_star_SS : std::min(p_max, p_star_RS);
        }
        if (p_2_out)
          *p_2_out = p_2;
        return compute_lambda(ri, rj, p_2);
      }

      double compute(const state_type &U_i,
                     const precomputed_type &prec_i,
                     const state_type &U_j,
                     const precomputed_type &prec_j,
                     const Vec<dim> &n_ij) const
      {
        const auto ri = riemann_data_from_state(U_i, prec_i[0], n_ij);
        const auto rj = riemann_data_from_state(U_j, prec_j[0], n_ij);
        return compute(ri, rj);
      }
    };


    /* ------------------------------------------------------------------ */

    template <int dim>
    struct Indicator {
      using Sys = System<dim>;
      using state_type = typename Sys::state_type;
      using flux_type = typename Sys::flux_type;
      using precomputed_type = typename Sys::precomputed_type;

      const Sys &sys;
      double rho_i_inverse = 0., eta_i = 0., gamma_min = 0.;
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
        gamma_min = prec_i[1];
        const double rho_i = Sys::density(U_i);
        rho_i_inverse = 1. / rho_i;
        eta_i = prec_i[3];
        d_eta_i = sys.surrogate_harten_entropy_derivative(U_i, eta_i, gamma_min);
        d_eta_i[0] -= eta_i * rho_i_inverse;
        const double surrogate_p_i = sys.surrogate_pressure(U_i, gamma_min);
        f_i = sys.f(U_i, surrogate_p_i);
        left = 0.;
        right.fill(0.);
      }

      void accumulate(const precomputed_type &, const state_type &U_j, const Vec<dim> &c_ij)
      {
        const double eta_j = sys.surrogate_harten_entropy(U_j, gamma_min);
        const double rho_j = Sys::density(U_j);
        const double rho_j_inverse = 1. / rho_j;
        const auto m_j = Sys::momentum(U_j);
        const double surrogate_p_j = sys.surrogate_pressure(U_j, gamma_min);
        const auto f_j = sys.f(U_j, surrogate_p_j);
        const double entropy_flux = (eta_j * rho_j_inverse - eta_i * rho_i_inverse) * dot<dim>(m_j, c_ij);
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
        const double quotient = std::abs(numerator) / (denominator + hd_i * std::abs(eta_i));
        return std::min(1., sys.prm.indicator_evc_factor * quotient);
      }
    };


    /* ------------------------------------------------------------------ */

    template <int dim>
    struct Limiter {
      using Sys = System<dim>;
      using state_type = typename Sys::state_type;
      using flux_type = typename Sys::flux_type;
      using precomputed_type = typename Sys::precomputed_type;
      using Bounds = typename Sys::bounds_type;

      const Sys &sys;
      state_type U_i;
      flux_type flux_i;
      Bounds bounds_;
      double rho_relaxation_numerator = 0., rho_relaxation_denominator = 0., s_interp_max = 0.;

      explicit Limiter(const Sys &s)
          : sys(s)
      {
      }

      void reset(const precomputed_type &prec_i, const state_type &new_U_i, const flux_type &new_flux_i)
      {
        U_i = new_U_i;
        flux_i = new_flux_i;
        bounds_[0] = max_double;
        bounds_[1] = 0.;
        bounds_[2] = max_double;
        bounds_[3] = prec_i[1];
        rho_relaxation_numerator = 0.;
        rho_relaxation_denominator = 0.;
        s_interp_max = 0.;
      }

      void accumulate(const precomputed_type &prec_j,
                      const state_type &U_j,
                      const flux_type &flux_j,
                      const Vec<dim> &scaled_c_ij)
      {
        double &rho_min = bounds_[0], &rho_max = bounds_[1], &s_min = bounds_[2];
        const double gamma_min = bounds_[3];
        const double rho_i = Sys::density(U_i);
        const double rho_j = Sys::density(U_j);

        /* 0.5 (U_i + U_j) - 0.5 contract(add(flux_j, -flux_i), scaled_c_ij) + affine_shift (= 0) */
        state_type U_ij_bar;
        for (int c = 0; c < Sys::k; ++c) {
          Vec<dim> diff;
          for (int d = 0; d < dim; ++d)
            diff[d] = flux_j[c][d] + (-flux_i[c][d]);
          U_ij_bar[c] = 0.5 * (U_i[c] + U_j[c]) - 0.5 * dot<dim>(diff, scaled_c_ij) + 0.;
        }
        const double rho_ij_bar = Sys::density(U_ij_bar);
        rho_min = std::min(rho_min, rho_ij_bar);
        rho_max = std::max(rho_max, rho_ij_bar);

        const double beta_ij = 1.;
        rho_relaxation_numerator += beta_ij * (rho_i + rho_j);
        rho_relaxation_denominator += std::abs(beta_ij);

        if (sys.prm.compute_strict_bounds) {
          const double s_j = sys.surrogate_specific_entropy(U_j, gamma_min);
          const double s_ij_bar = sys.surrogate_specific_entropy(U_ij_bar, gamma_min);
          state_type U_mid;
          for (int c = 0; c < Sys::k; ++c)
            U_mid[c] = (U_i[c] + U_j[c]) * .5;
          const double s_interp = sys.surrogate_specific_entropy(U_mid, gamma_min);
          s_min = std::min(s_min, s_j);
          s_min = std::min(s_min, s_ij_bar);
          s_interp_max = std::max(s_interp_max, s_interp);
        } else {
          const double s_j = prec_j[2];
          const double s_ij_bar = sys.surrogate_specific_entropy(U_ij_bar, gamma_min);
          s_min = std::min(s_min, s_j);
          s_min = std::min(s_min, s_ij_bar);
          s_interp_max = std::max(s_interp_max, s_ij_bar);
        }
      }

      Bounds bounds(const double hd_i) const
      {
        Bounds relaxed = bounds_;
        double &rho_min = relaxed[0], &rho_max = relaxed[1], &s_min = relaxed[2];
        const double gamma_min = relaxed[3];

        double r_i = std::sqrt(hd_i);
        if constexpr (dim == 2) {
          const double qq = std::sqrt(r_i);
          r_i = qq * qq * qq;
        } else if constexpr (dim == 1) {
          r_i = r_i * r_i * r_i;
        }
        r_i *= sys.prm.limiter_relaxation_factor;

        const double rho_relaxation =
            std::abs(rho_relaxation_numerator) / (std::abs(rho_relaxation_denominator) + eps_double);
        const double relaxation = (2. * sys.prm.limiter_relaxation_factor) * rho_relaxation;

        rho_min = std::max((1. - r_i) * rho_min, rho_min - relaxation);
        rho_max = std::min((1. + r_i) * rho_max, rho_max + relaxation);

        const double entropy_relaxation = sys.prm.limiter_relaxation_factor * (s_interp_max - s_min);
        s_min = std::max((1. - r_i) * s_min, s_min - entropy_relaxation);

        const double numerator = (gamma_min + 1.) * rho_max;
        const double denominator = gamma_min - 1. + 2. * sys.b * rho_max;
        const double upper_bound = numerator / denominator;
        rho_max = std::min(upper_bound, rho_max);
        return relaxed;
      }

      /* returns t_l; limiter.template.h:15-361 */
      double limit(const Bounds &bounds,
                   const state_type &U,
                   const state_type &P,
                   bool &success,
                   const double t_min = 0.,
                   const double t_max = 1.) const
      {
        const bool expensive = sys.prm.expensive_bounds_check;
        success = true;
        double t_r = t_max;

        const double small = sys.prm.vacuum_state_relaxation_small;
        const double large = sys.prm.vacuum_state_relaxation_large;
        const double relax_small = 1. + small * eps_double;
        const double relax = 1. + large * eps_double;

        auto axpy = [](const state_type &a, double t, const state_type &p) {
          state_type r;
          for (int c = 0; c < Sys::k; ++c)
            r[c] = a[c] + t * p[c];
          return r;
        };
        auto sdot = [](const state_type &a, const state_type &p) {
          double s = a[0] * p[0];
          for (int c = 1; c < Sys::k; ++c)
            s += a[c] * p[c];
          return s;
        };

        {
          const double rho_U = Sys::density(U);
          const double rho_P = Sys::density(P);
          const double rho_min = bounds[0];
          const double rho_max = bounds[1];

          const double test_min = sys.filter_vacuum_density(std::max(0., rho_U - relax * rho_max));
          const double test_max = sys.filter_vacuum_density(std::max(0., rho_min - relax * rho_U));
          if (!(test_min == 0. && test_max == 0.))
            success = false;

          const double denominator = 1. / (std::abs(rho_P) + eps_double * rho_max);
          t_r = rho_max < rho_U + t_r * rho_P ? (rho_max - rho_U) * denominator : t_r;
          t_r = rho_U + t_r * rho_P < rho_min ? (rho_U - rho_min) * denominator : t_r;
          t_r = std::min(t_r, t_max);
          t_r = std::max(t_r, t_min);

          if (expensive) {
            const double rho_new = Sys::density(axpy(U, t_r, P));
            const double test_new_min = sys.filter_vacuum_density(std::max(0., rho_new - relax * rho_max));
            const double test_new_max = sys.filter_vacuum_density(std::max(0., rho_min - relax * rho_new));
            if (!(test_new_min == 0. && test_new_max == 0.))
              success = false;
          }
        }

        double t_l = t_min;
        const double gamma = bounds[3];
        const double gm1 = gamma - 1.;
        const double b = sys.b, pinf = sys.pinf, q = sys.q;
        {
          const double s_min = bounds[2];
          for (unsigned int n = 0; n < sys.prm.limiter_newton_max_iterations; ++n) {
            const auto U_r = axpy(U, t_r, P);
            const double rho_r = Sys::density(U_r);
            const double rho_r_gamma = orc::pow(rho_r, gamma);
            const double covolume_r = 1. - b * rho_r;
            const double rho_e_r = Sys::internal_energy(U_r);
            const double shift_r = rho_e_r - rho_r * q - pinf * covolume_r;
            const double psi_r =
                relax_small * rho_r * shift_r - s_min * rho_r * rho_r_gamma * orc::pow(covolume_r, -gm1);

            if (!expensive) {
              t_l = psi_r > 0. ? t_r : t_l;
              if (t_l == t_r)
                break;
            }

            const auto U_l = axpy(U, t_l, P);
            const double rho_l = Sys::density(U_l);
            const double rho_l_gamma = orc::pow(rho_l, gamma);
            const double covolume_l = 1. - b * rho_l;
            const double rho_e_l = Sys::internal_energy(U_l);
            const double shift_l = rho_e_l - rho_l * q - pinf * covolume_l;
            const double psi_l =
                relax_small * rho_l * shift_l - s_min * rho_l * rho_l_gamma * orc::pow(covolume_l, -gm1);

            const double lower_bound = (1. - relax) * s_min * rho_l * rho_l_gamma * orc::pow(covolume_l, -gm1);
            if (n == 0 && !(std::min(0., psi_l - lower_bound) == 0.))
              success = false;

            if (expensive)
              t_l = psi_r > 0. ? t_r : t_l;

            const double tolerance = sys.prm.limiter_newton_tolerance;
            if (std::max(0., t_r - t_l - tolerance) == 0.)
              break;

            const double drho = Sys::density(P);
            const double drho_e_l = sdot(Sys::internal_energy_derivative(U_l), P);
            const double drho_e_r = sdot(Sys::internal_energy_derivative(U_r), P);

            const double q_pinf_term_l = 2. * rho_l * q + pinf * (1. - 2. * b * rho_l);
            const double q_pinf_term_r = 2. * rho_r * q + pinf * (1. - 2. * b * rho_r);

            const double extra_term_l = s_min * orc::pow(rho_l / covolume_l, gamma) * (covolume_l + gamma - b * rho_l);
            const double extra_term_r = s_min * orc::pow(rho_r / covolume_r, gamma) * (covolume_r + gamma - b * rho_r);

            const double dpsi_l = rho_l * drho_e_l + (rho_e_l - q_pinf_term_l - extra_term_l) * drho;
            const double dpsi_r = rho_r * drho_e_r + (rho_e_r - q_pinf_term_r - extra_term_r) * drho;

            quadratic_newton_step(t_l, t_r, psi_l, psi_r, dpsi_l, dpsi_r, -1.);
          }

          if (expensive) {
            const auto U_new = axpy(U, t_l, P);
            const double rho_new = Sys::density(U_new);
            const double covolume_new = 1. - b * rho_new;
            const double rho_new_gamma = orc::pow(rho_new, gamma);
            const double rho_e_new = Sys::internal_energy(U_new);
            const double shift_new = rho_e_new - rho_new * q - pinf * covolume_new;
            const double psi_new =
                relax_small * rho_new * shift_new - s_min * rho_new * rho_new_gamma * orc::pow(covolume_new, -gm1);
            const double lower_bound =
                (1. - relax) * s_min * rho_new * rho_new_gamma * orc::pow(covolume_new, -gm1);
            const bool e_valid = std::min(0., shift_new) == 0.;
            const bool psi_valid = std::min(0., psi_new - lower_bound) == 0.;
            if (!e_valid || !psi_valid)
              success = false;
          }
        }
        return t_l;
      }
    };
  } // namespace euler_aeos
} // namespace orc
