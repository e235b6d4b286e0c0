// Euler equations with an arbitrary equation of state ("euler aeos") as a device equation module:
// the third instantiation of the Eq plug-in surface the sweep kernels are templated on
// (INTEGRATION.md section 7).  Scalar restatement, operation by operation, of
//   HyperbolicSystemView  source/euler_aeos/hyperbolic_system.h:843-1528
//   equations of state    source/euler_aeos/equation_of_state_{polytropic_gas,
//                         noble_abel_stiffened_gas,van_der_waals}.h
//   RiemannSolver         source/euler_aeos/riemann_solver.template.h:23-676
//   Indicator             source/euler_aeos/indicator.h:165-263
//   Limiter               source/euler_aeos/limiter.h:244-445, limiter.template.h:15-361
// What differs from the polytropic module for the driver: 4 precomputed values (p, gamma_min, s,
// eta) produced in TWO precomputation cycles (cycle 0 per node, cycle 1 takes the minimum of the
// surrogate gamma over the stencil: k_precompute_stencil), the pressure of the flux is the
// equation-of-state pressure kept in the precomputed vector (aux = (1 / rho, p_eos)), 4 limiter
// bounds (rho_min, rho_max, s_min, gamma_min).  `dynamic` boundaries trap in the reference.
#pragma once

#include "device_common.cuh"

namespace rjb
{
  struct EulerAeosParams {
    int eos; /* rj_equation_of_state */
    double gamma;
    double eos_b, eos_q, eos_pinf, vdw_a; /* parameters of the selected equation of state */
    double b, pinf, q;                    /* eos_interpolation_{b, pinfty, q} */
    int compute_strict_bounds;
    double reference_density, vacuum_relaxation_small, vacuum_relaxation_large;
    double evc_factor;
    double limiter_newton_tolerance, limiter_relaxation_factor;
    unsigned int limiter_newton_max_iterations;
  };


  template <int dim_>
  struct EulerAeos {
    static constexpr int dim = dim_;
    static constexpr int k = dim + 2;
    static constexpr int n_prec = 4;   /* (p, gamma_min, s, eta) */
    static constexpr int n_bounds = 4; /* rho_min, rho_max, s_min, gamma_min */
    static constexpr int n_aux = 2;    /* (1 / rho, p) */
    static constexpr bool shallow_water = false;
    static constexpr bool aeos = true;
    static constexpr bool have_source_terms = false;
    static constexpr bool have_high_order_flux = false;
    static constexpr bool edge_cache = false; /* its midpoint entropy depends on gamma_min of row i */
    using Params = EulerAeosParams;

    /* ---------------- equation of state ---------------- */

    RJ_DEVICE static double eos_pressure(const Params &p, const double rho, const double e)
    {
      const double gamma = p.gamma;
      if (p.eos == RJ_EOS_NOBLE_ABEL_STIFFENED_GAS)
        return (gamma - 1.) * rho * (e - p.eos_q) / (1. - p.eos_b * rho) - gamma * p.eos_pinf;
      if (p.eos == RJ_EOS_VAN_DER_WAALS) {
        const double intermolecular = p.vdw_a * rho * rho;
        const double numerator = rho * e + intermolecular;
        const double covolume = 1. - p.eos_b * rho;
        return (gamma - 1.) * numerator / covolume - intermolecular;
      }
      return (gamma - 1.) * rho * e;
    }

    RJ_DEVICE static double eos_specific_internal_energy(const Params &p, const double rho, const double pr)
    {
      const double gamma = p.gamma;
      if (p.eos == RJ_EOS_NOBLE_ABEL_STIFFENED_GAS) {
        const double numerator = (pr + gamma * p.eos_pinf) * (1. - p.eos_b * rho);
        const double denominator = rho * (gamma - 1.);
        return p.eos_q + numerator / denominator;
      }
      if (p.eos == RJ_EOS_VAN_DER_WAALS) {
        const double intermolecular = p.vdw_a * rho * rho;
        const double covolume = 1. - p.eos_b * rho;
        const double numerator = (pr + intermolecular) * covolume;
        const double denominator = rho * (gamma - 1.);
        return numerator / denominator - p.vdw_a * rho;
      }
      return pr / (rho * (gamma - 1.));
    }

    /* ---------------- HyperbolicSystemView ---------------- */

    RJ_DEVICE static double momentum_norm_square(const double (&U)[k])
    {
      double m2 = U[1] * U[1];
#pragma unroll
      for (int d = 1; d < dim; ++d)
        m2 += U[1 + d] * U[1 + d];
      return m2;
    }

    RJ_DEVICE static double internal_energy(const double (&U)[k])
    {
      const double rho_inverse = 1. / U[0];
      return U[dim + 1] - 0.5 * momentum_norm_square(U) * rho_inverse;
    }

    RJ_DEVICE static double filter_vacuum_density(const Params &p, const double rho)
    {
      const double cutoff = p.reference_density * p.vacuum_relaxation_large * DBL_EPSILON;
      return fabs(rho) < cutoff ? 0. : rho;
    }

    RJ_DEVICE static double surrogate_specific_entropy(const Params &p, const double (&U)[k], const double gamma_min)
    {
      const double rho = U[0];
      const double rho_inverse = 1. / rho;
      const double covolume = 1. - p.b * rho;
      const double shift = internal_energy(U) - rho * p.q - p.pinf * covolume;
      return shift * path_pow(rho_inverse - p.b, gamma_min) / covolume;
    }

    RJ_DEVICE static double surrogate_harten_entropy(const Params &p, const double (&U)[k], const double gamma_min)
    {
      const double rho = U[0];
      const double E = U[dim + 1];
      const double rho_rho_e_q = rho * E - 0.5 * momentum_norm_square(U) - rho * rho * p.q;
      const double exponent = 1. / (gamma_min + 1.);
      const double covolume = 1. - p.b * rho;
      const double covolume_term = path_pow(covolume, gamma_min - 1.);
      const double rho_pinfcov = rho * p.pinf * covolume;
      return path_pow((rho_rho_e_q - rho_pinfcov) * covolume_term, exponent);
    }

    RJ_DEVICE static void surrogate_harten_entropy_derivative(
        const Params &p, const double (&U)[k], const double eta, const double gamma_min, double (&result)[k])
    {
      const double rho = U[0];
      const double E = U[dim + 1];
      const double covolume = 1. - p.b * rho;
      const double covolume_inverse = 1. / covolume;
      const double shift = rho * E - 0.5 * momentum_norm_square(U) - rho * rho * p.q - rho * p.pinf * covolume;
      const double factor =
          path_pow(eta * covolume_inverse, -gamma_min) * (covolume_inverse * covolume_inverse) / (gamma_min + 1.);
      const double first_term = E - 2. * rho * p.q - p.pinf * (1. - 2. * p.b * rho);
      const double second_term = -(gamma_min - 1.) * shift * p.b;
      result[0] = factor * (covolume * first_term + second_term);
#pragma unroll
      for (int d = 0; d < dim; ++d)
        result[1 + d] = -factor * covolume * U[1 + d];
      result[dim + 1] = factor * covolume * rho;
    }

    RJ_DEVICE static double surrogate_gamma(const Params &p, const double (&U)[k], const double pr)
    {
      const double rho = U[0];
      const double rho_e = internal_energy(U);
      const double covolume = 1. - p.b * rho;
      const double numerator = (pr + p.pinf) * covolume;
      const double denominator = rho_e - rho * p.q - covolume * p.pinf;
      return 1. + numerator / denominator;
    }

    RJ_DEVICE static double surrogate_pressure(const Params &p, const double (&U)[k], const double gamma)
    {
      const double rho = U[0];
      const double rho_e = internal_energy(U);
      const double covolume = 1. - p.b * rho;
      return (gamma - 1.) * (rho_e - rho * p.q) / covolume - gamma * p.pinf;
    }

    /* precomputation cycle 0, hyperbolic_system.h:925-941: (p_i, gamma_i, 0, 0) */
    RJ_DEVICE static void precompute(const Params &p, const double (&U)[k], double (&prec)[n_prec])
    {
      const double rho_i = U[0];
      const double e_i = internal_energy(U) / rho_i;
      const double p_i = eos_pressure(p, rho_i, e_i);
      prec[0] = p_i;
      prec[1] = surrogate_gamma(p, U, p_i);
      prec[2] = 0.;
      prec[3] = 0.;
    }

    /* precomputation cycle 1 after the stencil loop, :969-972 */
    RJ_DEVICE static void precompute_finish(const Params &p, const double (&U)[k], double (&prec)[n_prec])
    {
      prec[2] = surrogate_specific_entropy(p, U, prec[1]);
      prec[3] = surrogate_harten_entropy(p, U, prec[1]);
    }

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

    /* flux_contribution = f(U_i, p_i) with the precomputed equation-of-state pressure, :1390-1417 */
    RJ_DEVICE static Flux flux_contribution(const Params &, const double (&prec)[n_prec], const double (&U)[k])
    {
      Flux f;
      f.rho_inverse = 1. / U[0];
      f.p = prec[0];
      return f;
    }

    /* element (comp, d) of f(U, p), :1370-1387 */
    RJ_DEVICE static double f_element(const double (&U)[k], const Flux &f, const int comp, const int d)
    {
      if (comp == 0)
        return U[1 + d];
      if (comp == dim + 1)
        return U[1 + d] * (f.rho_inverse * (U[dim + 1] + f.p));
      const double v = U[1 + d] * (U[comp] * f.rho_inverse);
      return (d == comp - 1) ? v + f.p : v;
    }

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

    template <typename DirichletData>
    RJ_DEVICE static void apply_boundary_conditions(
        const Params &, const int id, double (&U)[k], const double (&normal)[dim], const DirichletData &D)
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
      }
      /* dynamic: __builtin_trap() in the reference (hyperbolic_system.h:1315); rejected on the host */
    }

    /* ---------------- RiemannSolver ---------------- */

    struct RiemannData {
      double rho, u, p, gamma, a;
    };

    RJ_DEVICE static double riemann_c(const double gamma)
    {
      const double slope = -0.34976871477801828189920753948709;
      const double first_radicand = (3. * gamma + 11.) / (6. * gamma + 6.);
      const double second_radicand = 5. / 6. + slope * (gamma - 3.);
      double radicand = fmin(first_radicand, second_radicand);
      radicand = fmin(1., radicand);
      radicand = fmax(1. / 2., radicand);
      return sqrt(radicand);
    }

    RJ_DEVICE static double riemann_alpha(const Params &p, const double rho, const double gamma, const double a)
    {
      const double numerator = 2. * a * (1. - p.b * rho);
      const double denominator = gamma - 1.;
      return numerator / denominator;
    }

    RJ_DEVICE static RiemannData
    riemann_data_from_state(const Params &p, const double (&U)[k], const double pr, const double (&n)[dim])
    {
      const double rho = U[0];
      const double rho_inverse = 1.0 / rho;
      double proj_m = n[0] * U[1];
#pragma unroll
      for (int d = 1; d < dim; ++d)
        proj_m += n[d] * U[1 + d];
      const double gamma = surrogate_gamma(p, U, pr);
      const double x = 1. - p.b * rho;
      RiemannData rd;
      rd.rho = rho;
      rd.u = proj_m * rho_inverse;
      rd.p = pr;
      rd.gamma = gamma;
      rd.a = sqrt(gamma * (pr + p.pinf) / (rho * x));
      return rd;
    }

    RJ_DEVICE static double p_star_failsafe(const Params &p, const RiemannData &ri, const RiemannData &rj)
    {
      const double b = p.b, pinf = p.pinf;
      const double p_max = fmax(ri.p, rj.p) + pinf;
      double radicand_i = 2. * (1. - b * ri.rho) * p_max;
      radicand_i /= ri.rho * ((ri.gamma + 1.) * p_max + (ri.gamma - 1.) * (ri.p + pinf));
      const double x_i = sqrt(radicand_i);
      double radicand_j = 2. * (1. - b * rj.rho) * p_max;
      radicand_j /= rj.rho * ((rj.gamma + 1.) * p_max + (rj.gamma - 1.) * (rj.p + pinf));
      const double x_j = sqrt(radicand_j);
      const double a = x_i + x_j;
      const double bb = rj.u - ri.u;
      const double cc = -(ri.p + pinf) * x_i - (rj.p + pinf) * x_j;
      const double base = (-bb + sqrt(bb * bb - 4. * a * cc)) / (2. * a);
      return base * base - pinf;
    }

    RJ_DEVICE static double p_star_RS_full(const Params &p, const RiemannData &ri, const RiemannData &rj)
    {
      const double pinf = p.pinf;
      const double alpha_i = riemann_alpha(p, ri.rho, ri.gamma, ri.a);
      const double alpha_j = riemann_alpha(p, rj.rho, rj.gamma, rj.a);
      const double p_min = fmin(ri.p, rj.p);
      const double p_max = fmax(ri.p, rj.p);
      const double gamma_min = ri.p < rj.p ? ri.gamma : rj.gamma;
      const double alpha_min = ri.p < rj.p ? alpha_i : alpha_j;
      const double alpha_hat_min = riemann_c(gamma_min) * alpha_min;
      const double alpha_max = ri.p >= rj.p ? alpha_i : alpha_j;
      const double gamma_m = fmin(ri.gamma, rj.gamma);
      const double gamma_M = fmax(ri.gamma, rj.gamma);
      const double numerator = positive_part(alpha_hat_min + alpha_max - (rj.u - ri.u));
      const double p_ratio = (p_min + pinf) / (p_max + pinf);
      const double r_exponent = (gamma_M - gamma_min) / (2. * gamma_min * gamma_M);
      const double first_exponent = (gamma_M - 1.) / (2. * gamma_M);
      const double first_exponent_inverse = 1. / first_exponent;
      const double first_denom = alpha_hat_min * path_pow(p_ratio, r_exponent - first_exponent) + alpha_max;
      const double p_1_tilde = (p_max + pinf) * path_pow(numerator / first_denom, first_exponent_inverse) - pinf;
      const double second_exponent = (gamma_m - 1.) / (2. * gamma_m);
      const double second_exponent_inverse = 1. / second_exponent;
      const double second_denom =
          alpha_hat_min * path_pow(p_ratio, -second_exponent) + alpha_max * path_pow(p_ratio, r_exponent);
      const double p_2_tilde = (p_max + pinf) * path_pow(numerator / second_denom, second_exponent_inverse) - pinf;
      return fmin(p_1_tilde, p_2_tilde);
    }

    RJ_DEVICE static double p_star_SS_full(const Params &p, const RiemannData &ri, const RiemannData &rj)
    {
      const double pinf = p.pinf;
      const double gamma_m = fmin(ri.gamma, rj.gamma);
      const double alpha_hat_i = riemann_c(ri.gamma) * riemann_alpha(p, ri.rho, ri.gamma, ri.a);
      const double alpha_hat_j = riemann_c(rj.gamma) * riemann_alpha(p, rj.rho, rj.gamma, rj.a);
      const double exponent = (gamma_m - 1.) / (2. * gamma_m);
      const double exponent_inverse = 1. / exponent;
      const double numerator = positive_part(alpha_hat_i + alpha_hat_j - (rj.u - ri.u));
      const double denominator = alpha_hat_i * path_pow((ri.p + pinf) / (rj.p + pinf), -exponent) + alpha_hat_j;
      const double p_1_tilde = (rj.p + pinf) * path_pow(numerator / denominator, exponent_inverse) - pinf;
      const double p_2_tilde = p_star_failsafe(p, ri, rj);
      return fmin(p_1_tilde, p_2_tilde);
    }

    RJ_DEVICE static double p_star_interpolated(const Params &p, const RiemannData &ri, const RiemannData &rj)
    {
      const double pinf = p.pinf;
      const double alpha_i = riemann_alpha(p, ri.rho, ri.gamma, ri.a);
      const double alpha_j = riemann_alpha(p, rj.rho, rj.gamma, rj.a);
      const double p_min = fmin(ri.p, rj.p) + pinf;
      const double p_max = fmax(ri.p, rj.p) + pinf;
      const double gamma_min = ri.p < rj.p ? ri.gamma : rj.gamma;
      const double alpha_min = ri.p < rj.p ? alpha_i : alpha_j;
      const double alpha_hat_min = riemann_c(gamma_min) * alpha_min;
      const double gamma_max = ri.p >= rj.p ? ri.gamma : rj.gamma;
      const double alpha_max = ri.p >= rj.p ? alpha_i : alpha_j;
      const double alpha_hat_max = riemann_c(gamma_max) * alpha_max;
      const double gamma_m = fmin(ri.gamma, rj.gamma);
      const double gamma_M = fmax(ri.gamma, rj.gamma);
      const double p_ratio = p_min / p_max;
      const double r_exponent = (gamma_M - gamma_min) / (2. * gamma_min * gamma_M);
      const double exponent = (gamma_m - 1.) / (2. * gamma_m);
      const double exponent_inverse = 1. / exponent;
      const double numerator = positive_part(alpha_hat_min + /*SIC!*/ alpha_max - (rj.u - ri.u));
      const double denominator =
          alpha_hat_min * path_pow(p_ratio, -exponent) + alpha_hat_max * path_pow(p_ratio, r_exponent);
      return p_max * path_pow(numerator / denominator, exponent_inverse) - pinf;
    }

    RJ_DEVICE static double phi_of_p_max(const Params &p, const RiemannData &ri, const RiemannData &rj)
    {
      const double b = p.b, pinf = p.pinf;
      const double p_max = fmax(ri.p, rj.p) + pinf;
      const double radicand_inverse_i =
          0.5 * ri.rho / (1. - b * ri.rho) * ((ri.gamma + 1.) * p_max + (ri.gamma - 1.) * (ri.p + pinf));
      const double value_i = (p_max - ri.p) / sqrt(radicand_inverse_i);
      const double radicand_inverse_j =
          0.5 * rj.rho / (1. - b * rj.rho) * ((rj.gamma + 1.) * p_max + (rj.gamma - 1.) * (rj.p + pinf));
      const double value_j = (p_max - rj.p) / sqrt(radicand_inverse_j);
      return value_i + value_j + rj.u - ri.u;
    }

    RJ_DEVICE static double compute_lambda(const Params &p, const RiemannData &ri, const RiemannData &rj, const double p_star)
    {
      const double factor_i = 0.5 * (ri.gamma + 1.) / ri.gamma;
      const double tmp_i = positive_part((p_star - ri.p) / (ri.p + p.pinf));
      const double nu_11 = ri.u - ri.a * sqrt(1. + factor_i * tmp_i);
      const double factor_j = 0.5 * (rj.gamma + 1.) / rj.gamma;
      const double tmp_j = positive_part((p_star - rj.p) / (rj.p + p.pinf));
      const double nu_32 = rj.u + rj.a * sqrt(1. + factor_j * tmp_j);
      return fmax(positive_part(nu_32), negative_part(nu_11));
    }

    RJ_DEVICE static double riemann_compute(const Params &p, const RiemannData &ri, const RiemannData &rj)
    {
      const double p_max = fmax(ri.p, rj.p) + p.pinf;
      const double phi_p_max = phi_of_p_max(p, ri, rj);
      double p_2;
      if (!p.compute_strict_bounds) {
        const double p_star_tilde = p_star_interpolated(p, ri, rj);
        const double p_star_backup = p_star_failsafe(p, ri, rj);
        p_2 = phi_p_max < 0. ? fmin(p_star_tilde, p_star_backup) : fmin(p_max, p_star_tilde);
      } else {
        const double p_star_RS = p_star_RS_full(p, ri, rj);
        const double p_star_SS = p_star_SS_full(p, ri, rj);
        p_2 = phi_p_max < 0. ? p_star_SS : fmin(p_max, p_star_RS);
      }
      return compute_lambda(p, ri, rj, p_2);
    }

    RJ_DEVICE static double riemann_lambda_max(const Params &p,
                                               const double (&U_i)[k],
                                               const double (&prec_i)[n_prec],
                                               const double (&U_j)[k],
                                               const double (&prec_j)[n_prec],
                                               const double (&n_ij)[dim])
    {
      const RiemannData rd_i = riemann_data_from_state(p, U_i, prec_i[0], n_ij);
      const RiemannData rd_j = riemann_data_from_state(p, U_j, prec_j[0], n_ij);
      return riemann_compute(p, rd_i, rd_j);
    }

    /* ---------------- Indicator ---------------- */

    struct Indicator {
      double rho_i_inverse, eta_i, gamma_min, p_i;
      double d_eta_i[k];
      double left;
      double right[k];

      RJ_DEVICE void reset(const Params &p, const double (&prec_i)[n_prec], const double (&U_i)[k])
      {
        gamma_min = prec_i[1];
        rho_i_inverse = 1. / U_i[0];
        eta_i = prec_i[3];
        surrogate_harten_entropy_derivative(p, U_i, eta_i, gamma_min, d_eta_i);
        d_eta_i[0] -= eta_i * rho_i_inverse;
        p_i = surrogate_pressure(p, U_i, gamma_min);
        left = 0.;
#pragma unroll
        for (int q = 0; q < k; ++q)
          right[q] = 0.;
      }

      /* f_i, f_j with the SURROGATE pressures for gamma_min of row i (indicator.h:215-240): the
       * precomputed flux contribution of j is not used */
      RJ_DEVICE void accumulate(const Params &p,
                                const double (&U_i)[k],
                                const double (&/*prec_i*/)[n_prec],
                                const double (&U_j)[k],
                                const double (&/*prec_j*/)[n_prec],
                                const Flux & /*flux_j*/,
                                const double (&c_ij)[dim])
      {
        const double eta_j = surrogate_harten_entropy(p, U_j, gamma_min);
        Flux fi, fj;
        fi.rho_inverse = rho_i_inverse;
        fi.p = p_i;
        fj.rho_inverse = 1. / U_j[0];
        fj.p = surrogate_pressure(p, U_j, gamma_min);
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

    struct LimiterAccumulator {
      double rho_min, rho_max, s_min, gamma_min;
      double rho_relaxation_numerator, rho_relaxation_denominator, s_interp_max;

      RJ_DEVICE void reset(const double (&prec_i)[n_prec])
      {
        rho_min = DBL_MAX;
        rho_max = 0.;
        s_min = DBL_MAX;
        gamma_min = prec_i[1];
        rho_relaxation_numerator = 0.;
        rho_relaxation_denominator = 0.;
        s_interp_max = 0.;
      }

      /* limiter.h:277-345; flux_i, flux_j carry the equation-of-state pressures */
      RJ_DEVICE void accumulate(const Params &p,
                                const double (&U_i)[k],
                                const Flux &flux_i,
                                const double (&U_j)[k],
                                const Flux &flux_j,
                                const double (&prec_j)[n_prec],
                                const double (&scaled_c_ij)[dim])
      {
        double U_ij_bar[k];
#pragma unroll
        for (int q = 0; q < k; ++q) {
          double s = (f_element(U_j, flux_j, q, 0) + (-f_element(U_i, flux_i, q, 0))) * scaled_c_ij[0];
#pragma unroll
          for (int d = 1; d < dim; ++d)
            s += (f_element(U_j, flux_j, q, d) + (-f_element(U_i, flux_i, q, d))) * scaled_c_ij[d];
          U_ij_bar[q] = 0.5 * (U_i[q] + U_j[q]) - 0.5 * s + 0.;
        }
        const double rho_ij_bar = U_ij_bar[0];
        rho_min = fmin(rho_min, rho_ij_bar);
        rho_max = fmax(rho_max, rho_ij_bar);
        rho_relaxation_numerator += 1. * (U_i[0] + U_j[0]);
        rho_relaxation_denominator += 1.;
        if (p.compute_strict_bounds) {
          const double s_j = surrogate_specific_entropy(p, U_j, gamma_min);
          const double s_ij_bar = surrogate_specific_entropy(p, U_ij_bar, gamma_min);
          double U_mid[k];
#pragma unroll
          for (int q = 0; q < k; ++q)
            U_mid[q] = (U_i[q] + U_j[q]) * .5;
          const double s_interp = surrogate_specific_entropy(p, U_mid, gamma_min);
          s_min = fmin(s_min, s_j);
          s_min = fmin(s_min, s_ij_bar);
          s_interp_max = fmax(s_interp_max, s_interp);
        } else {
          const double s_j = prec_j[2];
          const double s_ij_bar = surrogate_specific_entropy(p, U_ij_bar, gamma_min);
          s_min = fmin(s_min, s_j);
          s_min = fmin(s_min, s_ij_bar);
          s_interp_max = fmax(s_interp_max, s_ij_bar);
        }
      }

      RJ_DEVICE void bounds(const Params &p, const double hd_i, double (&b)[n_bounds]) const
      {
        double r_i = sqrt(hd_i);
        if (dim == 2) {
          const double qq = sqrt(r_i);
          r_i = qq * qq * qq;
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
        b[3] = gamma_min;
        const double numerator = (gamma_min + 1.) * b[1];
        const double denominator = gamma_min - 1. + 2. * p.b * b[1];
        const double upper_bound = numerator / denominator;
        b[1] = fmin(upper_bound, b[1]);
      }
    };

    /* Limiter::limit, limiter.template.h:15-361 (EXPENSIVE_BOUNDS_CHECK off) */
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
      const double gamma = bounds[3];
      const double gm1 = gamma - 1.;
      const double b = p.b, pinf = p.pinf, q = p.q;
      const double s_min = bounds[2];

      for (unsigned int n = 0; n < p.limiter_newton_max_iterations; ++n) {
        double U_r[k];
#pragma unroll
        for (int c = 0; c < k; ++c)
          U_r[c] = U[c] + t_r * P[c];
        const double rho_r = U_r[0];
        const double rho_r_gamma = path_pow(rho_r, gamma);
        const double covolume_r = 1. - b * rho_r;
        const double rho_e_r = internal_energy(U_r);
        const double shift_r = rho_e_r - rho_r * q - pinf * covolume_r;
        const double psi_r = relax_small * rho_r * shift_r - s_min * rho_r * rho_r_gamma * path_pow(covolume_r, -gm1);

        t_l = psi_r > 0. ? t_r : t_l;
        if (t_l == t_r)
          break;

        double U_l[k];
#pragma unroll
        for (int c = 0; c < k; ++c)
          U_l[c] = U[c] + t_l * P[c];
        const double rho_l = U_l[0];
        const double rho_l_gamma = path_pow(rho_l, gamma);
        const double covolume_l = 1. - b * rho_l;
        const double rho_e_l = internal_energy(U_l);
        const double shift_l = rho_e_l - rho_l * q - pinf * covolume_l;
        const double psi_l = relax_small * rho_l * shift_l - s_min * rho_l * rho_l_gamma * path_pow(covolume_l, -gm1);

        const double lower_bound = (1. - relax) * s_min * rho_l * rho_l_gamma * path_pow(covolume_l, -gm1);
        if (n == 0 && !(fmin(0., psi_l - lower_bound) == 0.))
          success = false;

        if (fmax(0., t_r - t_l - p.limiter_newton_tolerance) == 0.)
          break;

        /* internal_energy_derivative(V) * P */
        auto de_dot_P = [&](const double(&V)[k]) {
          const double rho_inverse = 1. / V[0];
          double u[dim];
          double u2 = 0.;
#pragma unroll
          for (int d = 0; d < dim; ++d) {
            u[d] = V[1 + d] * rho_inverse;
            u2 += u[d] * u[d];
          }
          double acc = (0.5 * u2) * P[0];
#pragma unroll
          for (int d = 0; d < dim; ++d)
            acc += (-u[d]) * P[1 + d];
          acc += 1. * P[dim + 1];
          return acc;
        };
        const double drho = P[0];
        const double drho_e_l = de_dot_P(U_l);
        const double drho_e_r = de_dot_P(U_r);
        const double q_pinf_term_l = 2. * rho_l * q + pinf * (1. - 2. * b * rho_l);
        const double q_pinf_term_r = 2. * rho_r * q + pinf * (1. - 2. * b * rho_r);
        const double extra_term_l = s_min * path_pow(rho_l / covolume_l, gamma) * (covolume_l + gamma - b * rho_l);
        const double extra_term_r = s_min * path_pow(rho_r / covolume_r, gamma) * (covolume_r + gamma - b * rho_r);
        const double dpsi_l = rho_l * drho_e_l + (rho_e_l - q_pinf_term_l - extra_term_l) * drho;
        const double dpsi_r = rho_r * drho_e_r + (rho_e_r - q_pinf_term_r - extra_term_r) * drho;
        quadratic_newton_step(t_l, t_r, psi_l, psi_r, dpsi_l, dpsi_r, -1.);
      }
      return t_l;
    }

    /* ---------------- initial states (Dirichlet data on device) -------- */

    /* from_initial_state -> from_primitive_state through the equation of state, :1441-1479 */
    RJ_DEVICE static void from_initial_state(const Params &p, const double (&prim)[3], double (&U)[k])
    {
      const double rho = prim[0], u = prim[1], pr = prim[2];
      const double e = eos_specific_internal_energy(p, rho, pr);
#pragma unroll
      for (int c = 0; c < k; ++c)
        U[c] = 0.;
      U[0] = rho;
      U[1] = u * rho;
      /* rho * e + 0.5 * rho * u * u with u the velocity tensor: (0.5 * rho) * |u|^2 */
      U[dim + 1] = rho * e + 0.5 * rho * (u * u);
    }

    template <class IV>
    RJ_DEVICE static void initial_state_compute(
        const Params &p, const IV &iv, const double (&x)[dim], const double t, double (&U)[k])
    {
      if (iv.kind == RJ_INITIAL_CONTRAST) {
        if (x[0] > 0.)
          from_initial_state(p, iv.primitive_right, U);
        else
          from_initial_state(p, iv.primitive, U);
      } else if (iv.kind == RJ_INITIAL_ISENTROPIC_VORTEX) {
        /* euler/initial_state_isentropic_vortex.h:53-92 with its own gamma parameter */
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
        for (int c = 0; c < k; ++c)
          U[c] = 0.;
        U[0] = rho;
        U[1] = rho * u;
        if (dim >= 2)
          U[dim >= 2 ? 2 : 1] = rho * v;
        U[dim + 1] = E;
      } else {
        from_initial_state(p, iv.primitive, U);
      }
    }
  };
} // namespace rjb
