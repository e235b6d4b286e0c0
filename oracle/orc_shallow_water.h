// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_common.h).
//
// Shallow water equation module, scalar double semantics.
//   System / View   <- source/shallow_water/hyperbolic_system.h:643-1330
//   RiemannSolver   <- source/shallow_water/riemann_solver.template.h:24-252
//   Indicator       <- source/shallow_water/indicator.h:157-222
//   Limiter         <- source/shallow_water/limiter.h:249-395,
//                      source/shallow_water/limiter.template.h:16-449
#pragma once

#include "orc_common.h"

namespace orc
{
  namespace sw
  {
    struct Parameters {
      /* shallow_water/hyperbolic_system.h:643-674 */
      double gravity = 9.81;
      double manning_friction_coefficient = 0.;
      double reference_water_depth = 1.;
      double dry_state_relaxation_factor = 2.e-1;
      double dry_state_relaxation_small = 1.e2;
      double dry_state_relaxation_large = 1.e4;
      /* shallow_water/indicator.h:28-31 */
      double indicator_evc_factor = 1.;
      /* shallow_water/limiter.h:27-59 */
      unsigned int limiter_iterations = 2;
      double limiter_newton_tolerance = 1.e-10;
      unsigned int limiter_newton_max_iterations = 2;
      double limiter_relaxation_factor = 1.;
      bool limit_on_kinetic_energy = false;
      bool limit_on_square_velocity = true;
      bool expensive_bounds_check = false;
    };


    template <int dim>
    struct System {
      static constexpr int k = dim + 1;
      static constexpr int n_precomputed = 2;         /* (eta_m, h_star) */
      static constexpr int n_initial_precomputed = 1; /* bathymetry */
      static constexpr int n_bounds = 5;              /* h_min, h_max, h_small, kin_max, v2_max */
      static constexpr bool shallow_water = true;
      static constexpr bool aeos = false;
      static constexpr bool have_source_terms = true;
      static constexpr bool have_high_order_flux = true;
      using state_type = std::array<double, k>;
      using flux_type = std::array<Vec<dim>, k>;
      using precomputed_type = std::array<double, n_precomputed>;
      using bounds_type = std::array<double, n_bounds>;
      struct flux_contribution_type {
        state_type U;
        double Z;
      };

      Parameters prm;

      explicit System(const Parameters &p = Parameters())
          : prm(p)
      {
      }

      static double water_depth(const state_type &U) { return U[0]; }

      static Vec<dim> momentum(const state_type &U)
      {
        Vec<dim> m;
        for (int d = 0; d < dim; ++d)
          m[d] = U[1 + d];
        return m;
      }

      double inverse_water_depth_mollified(const state_type &U) const
      {
        const double h_cutoff_mollified =
            prm.reference_water_depth * prm.dry_state_relaxation_large * eps_double;
        const double h = water_depth(U);
        const double h_pos = positive_part(water_depth(U));
        const double h_max = std::max(h, h_cutoff_mollified);
        const double denom = h * h + h_max * h_max;
        return 2. * h_pos / denom;
      }

      double water_depth_sharp(const state_type &U) const
      {
        const double h_cutoff_small =
            prm.reference_water_depth * prm.dry_state_relaxation_small * eps_double;
        return std::max(water_depth(U), h_cutoff_small);
      }

      double inverse_water_depth_sharp(const state_type &U) const
      {
        return 1. / water_depth_sharp(U);
      }

      double filter_dry_water_depth(const double h) const
      {
        const double h_cutoff_large =
            prm.reference_water_depth * prm.dry_state_relaxation_large * eps_double;
        return std::abs(h) < h_cutoff_large ? 0. : h;
      }

      double kinetic_energy(const state_type &U) const
      {
        const double h = water_depth(U);
        const double hi = inverse_water_depth_sharp(U);
        Vec<dim> vel = momentum(U);
        for (int d = 0; d < dim; ++d)
          vel[d] *= hi;
        return 0.5 * h * norm_square<dim>(vel);
      }

      double pressure(const state_type &U) const
      {
        const double h_sqd = U[0] * U[0];
        return 0.5 * prm.gravity * h_sqd;
      }

      double speed_of_sound(const state_type &U) const
      {
        return std::sqrt(prm.gravity * U[0]);
      }

      double mathematical_entropy(const state_type &U) const
      {
        const double p = pressure(U);
        const double k_e = kinetic_energy(U);
        return p + k_e;
      }

      state_type mathematical_entropy_derivative(const state_type &U) const
      {
        state_type result;
        const double h = U[0];
        const double hi = inverse_water_depth_sharp(U);
        Vec<dim> vel = momentum(U);
        for (int d = 0; d < dim; ++d)
          vel[d] *= hi;
        result[0] = prm.gravity * h - 0.5 * norm_square<dim>(vel);
        for (int d = 0; d < dim; ++d)
          result[1 + d] = vel[d];
        return result;
      }

      bool is_admissible(const state_type &U) const
      {
        return filter_dry_water_depth(water_depth(U)) >= 0.;
      }

      /* precomputation_loop body, :701-707 */
      precomputed_type precompute(const state_type &U) const
      {
        const double eta_m = mathematical_entropy(U);
        const double h_sharp = water_depth_sharp(U);
        const double h_star = orc::pow(h_sharp, 4. / 3.);
        return {{eta_m, h_star}};
      }

      flux_type f(const state_type &U) const
      {
        const double h_inverse = inverse_water_depth_sharp(U);
        const auto m = momentum(U);
        const double p = pressure(U);
        flux_type result;
        for (int d = 0; d < dim; ++d)
          result[0][d] = (m[d] * h_inverse) * U[0];
        for (int i = 0; i < dim; ++i) {
          for (int d = 0; d < dim; ++d)
            result[1 + i][d] = (m[d] * h_inverse) * m[i];
          result[1 + i][i] += p;
        }
        return result;
      }

      flux_type g(const state_type &U) const
      {
        const double h_inverse = inverse_water_depth_sharp(U);
        const auto m = momentum(U);
        flux_type result;
        for (int d = 0; d < dim; ++d)
          result[0][d] = (m[d] * h_inverse) * U[0];
        for (int i = 0; i < dim; ++i)
          for (int d = 0; d < dim; ++d)
            result[1 + i][d] = (m[d] * h_inverse) * m[i];
        return result;
      }

      state_type star_state(const state_type &U, const double Z_left, const double Z_right) const
      {
        const double Z_max = std::max(Z_left, Z_right);
        const double h = water_depth(U);
        const double H_star = std::max(0., h + Z_left - Z_max);
        /* U * H_star * inverse_water_depth_mollified(U): (U * H_star) * inv */
        const double inv = inverse_water_depth_mollified(U);
        state_type r;
        for (int c = 0; c < k; ++c)
          r[c] = (U[c] * H_star) * inv;
        return r;
      }

      std::array<state_type, 2> equilibrated_states(const flux_contribution_type &flux_i,
                                                    const flux_contribution_type &flux_j) const
      {
        return {{star_state(flux_i.U, flux_i.Z, flux_j.Z), star_state(flux_j.U, flux_j.Z, flux_i.Z)}};
      }

      flux_contribution_type
      flux_contribution(const precomputed_type &, const double Z, const state_type &U) const
      {
        return {U, Z};
      }

      static state_type contract(const flux_type &t, const Vec<dim> &c)
      {
        state_type r;
        for (int q = 0; q < k; ++q)
          r[q] = dot<dim>(t[q], c);
        return r;
      }

      state_type flux_divergence(const flux_contribution_type &flux_i,
                                 const flux_contribution_type &flux_j,
                                 const Vec<dim> &c_ij) const
      {
        const auto &U_i = flux_i.U;
        const auto eq = equilibrated_states(flux_i, flux_j);
        const auto &U_star_ij = eq[0];
        const auto &U_star_ji = eq[1];

        const double H_i = water_depth(U_i);
        const double H_star_ij = water_depth(U_star_ij);
        const double H_star_ji = water_depth(U_star_ji);

        const auto g_i = g(U_star_ij);
        const auto g_j = g(U_star_ji);

        flux_type result;
        for (int q = 0; q < k; ++q)
          for (int d = 0; d < dim; ++d)
            result[q][d] = -(g_i[q][d] + g_j[q][d]);

        const double factor =
            (0.5 * (H_star_ji * H_star_ji - H_star_ij * H_star_ij) + H_i * H_i) * prm.gravity;

        for (int i = 0; i < dim; ++i)
          result[1 + i][i] -= factor;

        return contract(result, c_ij);
      }

      state_type high_order_flux_divergence(const flux_contribution_type &flux_i,
                                            const flux_contribution_type &flux_j,
                                            const Vec<dim> &c_ij) const
      {
        const auto &U_i = flux_i.U;
        const auto &U_j = flux_j.U;
        const double H_i = water_depth(U_i);
        const double H_j = water_depth(U_j);
        const auto g_i = g(U_i);
        const auto g_j = g(U_j);
        flux_type result;
        for (int q = 0; q < k; ++q)
          for (int d = 0; d < dim; ++d)
            result[q][d] = -(g_i[q][d] + g_j[q][d]);
        const double factor = prm.gravity * H_i * (H_j + flux_j.Z - flux_i.Z);
        for (int i = 0; i < dim; ++i)
          result[1 + i][i] -= factor;
        return contract(result, c_ij);
      }

      state_type affine_shift(const flux_contribution_type &flux_i,
                              const flux_contribution_type &flux_j,
                              const Vec<dim> &c_ij,
                              const double d_ij) const
      {
        const auto &U_i = flux_i.U;
        const auto U_star_ij = star_state(U_i, flux_i.Z, flux_j.Z);
        const double h_inverse = inverse_water_depth_sharp(U_i);
        const auto m = momentum(U_i);
        const double factor = 2. * (d_ij + h_inverse * dot<dim>(m, c_ij));
        state_type r;
        for (int c = 0; c < k; ++c)
          r[c] = -factor * (U_star_ij[c] - U_i[c]);
        return r;
      }

      state_type manning_friction(const state_type &U, const double h_star, const double tau) const
      {
        state_type result{};
        const double g = prm.gravity;
        const double n = prm.manning_friction_coefficient;
        const double h_inverse = inverse_water_depth_mollified(U);
        Vec<dim> v = momentum(U);
        for (int d = 0; d < dim; ++d)
          v[d] *= h_inverse;
        const double v_norm = norm<dim>(v);
        const double factor = 2. * g * n * n * v_norm;
        const double denominator = h_star + std::max(h_star, tau * factor);
        const double denominator_inverse = 1. / denominator;
        for (int d = 0; d < dim; ++d)
          result[d + 1] = -factor * denominator_inverse * U[1 + d];
        return result;
      }

      state_type nodal_source(const precomputed_type &prec, const state_type &U, const double tau) const
      {
        return manning_friction(U, prec[1], tau);
      }

      /* 1D primitive (h, u) -> conserved, hyperbolic_system.h:1251-1282 */
      state_type from_initial_state(const std::array<double, 2> &prim) const
      {
        state_type s{};
        s[0] = prim[0];
        s[1] = prim[1] * prim[0];
        return s;
      }

      template <int component>
      state_type prescribe_riemann_characteristic(const state_type &U,
                                                  const state_type &U_bar,
                                                  const Vec<dim> &normal) const
      {
        const auto m = momentum(U);
        const double a = speed_of_sound(U);
        const double hi = inverse_water_depth_sharp(U);
        const double vn = dot<dim>(m, normal) * hi;
        const auto m_bar = momentum(U_bar);
        const double a_bar = speed_of_sound(U_bar);
        const double vn_bar = dot<dim>(m_bar, normal) * inverse_water_depth_sharp(U_bar);
        const double R_1 = component == 1 ? vn_bar - 2. * a_bar : vn - 2. * a;
        const double R_2 = component == 2 ? vn_bar + 2. * a_bar : vn + 2. * a;
        Vec<dim> vperp;
        for (int d = 0; d < dim; ++d)
          vperp[d] = m[d] * hi - vn * normal[d];
        const double vn_new = 0.5 * (R_1 + R_2);
        const double q = (R_2 - R_1) / 4.;
        const double h_new = (q * q) / prm.gravity;
        state_type U_new;
        U_new[0] = h_new;
        for (int d = 0; d < dim; ++d)
          U_new[1 + d] = h_new * (vn_new * normal[d] + vperp[d]);
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
        } else if (id == Boundary::dirichlet_momentum) {
          const auto m_dir = momentum(get_dirichlet_data());
          for (int d = 0; d < dim; ++d)
            result[d + 1] = m_dir[d];
        } else if (id == Boundary::slip) {
          auto m = momentum(U);
          const double mn = dot<dim>(m, normal);
          for (int d = 0; d < dim; ++d)
            m[d] -= 1. * mn * normal[d];
          for (int d = 0; d < dim; ++d)
            result[d + 1] = m[d];
        } else if (id == Boundary::no_slip) {
          for (int d = 0; d < dim; ++d)
            result[d + 1] = 0.;
        } else if (id == Boundary::dynamic) {
          const auto m = momentum(U);
          const double h_inverse = inverse_water_depth_sharp(U);
          const double a = speed_of_sound(U);
          const double vn = dot<dim>(m, normal) * h_inverse;
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


    struct RiemannTrace {
      double h_star = 0., lambda_max = 0.;
    };

    template <int dim>
    struct RiemannSolver {
      using Sys = System<dim>;
      using state_type = typename Sys::state_type;
      using primitive_type = std::array<double, 3>; /* h, u, a */

      const Sys &sys;
      explicit RiemannSolver(const Sys &s)
          : sys(s)
      {
      }

      double f(const primitive_type &rd_Z, const double h) const
      {
        const double gravity = sys.prm.gravity;
        const double h_Z = rd_Z[0], a_Z = rd_Z[2];
        const double left_value = 2. * (std::sqrt(gravity * h) - a_Z);
        const double radicand = 0.5 * gravity * (h + h_Z) / (h * h_Z);
        const double right_value = (h - h_Z) * std::sqrt(radicand);
        return h <= h_Z ? left_value : right_value;
      }

      double phi(const primitive_type &rd_i, const primitive_type &rd_j, const double h) const
      {
        return f(rd_i, h) + f(rd_j, h) + rd_j[1] - rd_i[1];
      }

      double lambda1_minus(const primitive_type &rd, const double h_star) const
      {
        const double h = rd[0], u = rd[1], a = rd[2];
        const double factor = positive_part((h_star - h) / h);
        const double half_factor = 0.5 * factor;
        return u - a * std::sqrt((1. + half_factor) * (1. + factor));
      }

      double lambda3_plus(const primitive_type &rd, const double h_star) const
      {
        const double h = rd[0], u = rd[1], a = rd[2];
        const double factor = positive_part((h_star - h) / h);
        const double half_factor = 0.5 * factor;
        return u + a * std::sqrt((1. + half_factor) * (1. + factor));
      }

      double compute_lambda(const primitive_type &rd_i, const primitive_type &rd_j, const double h_star) const
      {
        const double lambda1 = lambda1_minus(rd_i, h_star);
        const double lambda3 = lambda3_plus(rd_j, h_star);
        return std::max(negative_part(lambda1), positive_part(lambda3));
      }

      double compute_h_star(const primitive_type &rd_i, const primitive_type &rd_j) const
      {
        const double gravity = sys.prm.gravity;
        const double gravity_inverse = 1. / gravity;
        const double h_i = rd_i[0], u_i = rd_i[1], a_i = rd_i[2];
        const double h_j = rd_j[0], u_j = rd_j[1], a_j = rd_j[2];

        const double h_min = std::min(h_i, h_j);
        const double h_max = std::max(h_i, h_j);
        const double a_min = std::sqrt(gravity * h_min);
        const double a_max = std::sqrt(gravity * h_max);
        const double sqrt_two = std::sqrt(2.);
        const double x0 = 9. - 4. * sqrt_two;

        const double phi_value_min = phi(rd_i, rd_j, x0 * h_min);
        const double phi_value_max = phi(rd_i, rd_j, x0 * h_max);

        double tmp;
        tmp = positive_part(u_i - u_j + 2. * (a_i + a_j));
        const double h_star_left = 0.0625 * gravity_inverse * tmp * tmp;

        tmp = 1. + sqrt_two * (u_i - u_j) / (a_min + a_max);
        const double h_star_middle = std::sqrt(h_min * h_max) * tmp;

        const double left_radicand = 3. * h_min + 2. * sqrt_two * std::sqrt(h_min * h_max);
        const double right_radicand = sqrt_two * std::sqrt(gravity_inverse * h_min) * (u_i - u_j);
        tmp = std::sqrt(positive_part(left_radicand + right_radicand));
        tmp -= sqrt_two * std::sqrt(h_min);
        const double h_star_right = tmp * tmp;

        double h_star = 0. <= phi_value_min ? h_star_left : h_star_right;
        h_star = phi_value_max < 0. ? h_star_middle : h_star_right;
        return h_star;
      }

      primitive_type riemann_data_from_state(const state_type &U, const Vec<dim> &n_ij) const
      {
        const double h = sys.water_depth_sharp(U);
        const double gravity = sys.prm.gravity;
        /* momentum / h: Tensor / scalar multiplies by the rounded inverse */
        const double h_inverse = 1. / h;
        Vec<dim> velocity = Sys::momentum(U);
        for (int d = 0; d < dim; ++d)
          velocity[d] *= h_inverse;
        const double projected_velocity = dot<dim>(n_ij, velocity);
        const double a = std::sqrt(h * gravity);
        return {{h, projected_velocity, a}};
      }

      double compute(const primitive_type &rd_i, const primitive_type &rd_j, RiemannTrace *trace = nullptr) const
      {
        const double h_star = compute_h_star(rd_i, rd_j);
        const double lambda_max = compute_lambda(rd_i, rd_j, h_star);
        if (trace) {
          trace->h_star = h_star;
          trace->lambda_max = lambda_max;
        }
        return lambda_max;
      }

      double compute(const state_type &U_i, const state_type &U_j, const Vec<dim> &n_ij) const
      {
        return compute(riemann_data_from_state(U_i, n_ij), riemann_data_from_state(U_j, n_ij));
      }
    };


    template <int dim>
    struct Indicator {
      using Sys = System<dim>;
      using state_type = typename Sys::state_type;
      using flux_type = typename Sys::flux_type;
      using precomputed_type = typename Sys::precomputed_type;

      const Sys &sys;
      double h_i = 0., eta_i = 0., pressure_i = 0.;
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
        h_i = Sys::water_depth(U_i);
        eta_i = prec_i[0];
        d_eta_i = sys.mathematical_entropy_derivative(U_i);
        f_i = sys.f(U_i);
        pressure_i = sys.pressure(U_i);
        left = 0.;
        right.fill(0.);
      }

      void accumulate(const precomputed_type &prec_j, const state_type &U_j, const Vec<dim> &c_ij)
      {
        const double eta_j = prec_j[0];
        const double hi = sys.inverse_water_depth_sharp(U_j);
        Vec<dim> velocity_j = Sys::momentum(U_j);
        for (int d = 0; d < dim; ++d)
          velocity_j[d] *= hi;
        const auto f_j = sys.f(U_j);
        const double pressure_j = sys.pressure(U_j);
        left += (eta_j + pressure_j) * dot<dim>(velocity_j, c_ij);
        for (int q = 0; q < Sys::k; ++q) {
          Vec<dim> diff;
          for (int d = 0; d < dim; ++d)
            diff[d] = f_j[q][d] - f_i[q][d];
          right[q] += dot<dim>(diff, c_ij);
        }
      }

      double alpha(const double hd_i) const
      {
        double my_sum = 0.;
        for (int q = 0; q < Sys::k; ++q)
          my_sum += d_eta_i[q] * right[q];
        const double numerator = std::abs(left - my_sum);
        const double denominator = std::abs(left) + std::abs(my_sum);
        const double regularization = 100. * min_double;
        const double quotient =
            std::abs(numerator) / (denominator + std::max(hd_i * std::abs(eta_i), regularization));
        return std::min(1., sys.prm.indicator_evc_factor * quotient);
      }
    };


    template <int dim>
    struct Limiter {
      using Sys = System<dim>;
      using state_type = typename Sys::state_type;
      using Bounds = typename Sys::bounds_type;

      const Sys &sys;
      state_type U_i;
      Bounds bounds_;
      double h_relaxation_numerator = 0., kin_relaxation_numerator = 0., v2_relaxation_numerator = 0.,
             relaxation_denominator = 0.;

      explicit Limiter(const Sys &s)
          : sys(s)
      {
      }

      void reset(const state_type &new_U_i)
      {
        U_i = new_U_i;
        bounds_[0] = max_double;
        bounds_[1] = 0.;
        bounds_[2] = 0.;
        bounds_[3] = 0.;
        bounds_[4] = 0.;
        h_relaxation_numerator = 0.;
        kin_relaxation_numerator = 0.;
        v2_relaxation_numerator = 0.;
        relaxation_denominator = 0.;
      }

      void accumulate(const state_type &U_j,
                      const state_type &U_star_ij,
                      const state_type &U_star_ji,
                      const Vec<dim> &scaled_c_ij,
                      const state_type &affine_shift)
      {
        const auto f_star_ij = sys.f(U_star_ij);
        const auto f_star_ji = sys.f(U_star_ji);

        state_type U_ij_bar;
        for (int q = 0; q < Sys::k; ++q) {
          Vec<dim> diff;
          for (int d = 0; d < dim; ++d)
            diff[d] = f_star_ij[q][d] + (-f_star_ji[q][d]);
          U_ij_bar[q] = 0.5 * (U_star_ij[q] + U_star_ji[q] + dot<dim>(diff, scaled_c_ij)) + affine_shift[q];
        }

        double &h_min = bounds_[0], &h_max = bounds_[1], &kin_max = bounds_[3], &v2_max = bounds_[4];

        const double h_bar_ij = Sys::water_depth(U_ij_bar);
        h_min = std::min(h_min, h_bar_ij);
        h_max = std::max(h_max, h_bar_ij);

        const double kin_bar_ij = sys.kinetic_energy(U_ij_bar);
        kin_max = std::max(kin_max, kin_bar_ij);

        const double inv_bar = sys.inverse_water_depth_mollified(U_ij_bar);
        Vec<dim> v_bar_ij = Sys::momentum(U_ij_bar);
        for (int d = 0; d < dim; ++d)
          v_bar_ij[d] *= inv_bar;
        const double v2_bar_ij = norm_square<dim>(v_bar_ij);
        v2_max = std::max(v2_max, v2_bar_ij);

        const double beta_ij = 1.;
        relaxation_denominator += std::abs(beta_ij);

        const double h_i = Sys::water_depth(U_i);
        const double h_j = Sys::water_depth(U_j);
        h_relaxation_numerator += beta_ij * (h_i + h_j);

        const double kin_i = sys.kinetic_energy(U_i);
        const double kin_j = sys.kinetic_energy(U_j);
        kin_relaxation_numerator += beta_ij * (kin_i + kin_j);

        const double inv_i = sys.inverse_water_depth_mollified(U_i);
        const double inv_j = sys.inverse_water_depth_mollified(U_j);
        Vec<dim> vel_i = Sys::momentum(U_i), vel_j = Sys::momentum(U_j);
        for (int d = 0; d < dim; ++d) {
          vel_i[d] *= inv_i;
          vel_j[d] *= inv_j;
        }
        v2_relaxation_numerator += beta_ij * (-norm_square<dim>(vel_i) + norm_square<dim>(vel_j));
      }

      Bounds bounds(const double hd_i) const
      {
        Bounds relaxed = bounds_;
        double &h_min = relaxed[0], &h_max = relaxed[1], &h_small = relaxed[2], &kin_max = relaxed[3],
               &v2_max = relaxed[4];

        double r_i = std::sqrt(hd_i);
        if constexpr (dim == 2) {
          const double q = std::sqrt(r_i);
          r_i = q * q * q;
        } else if constexpr (dim == 1) {
          r_i = r_i * r_i * r_i;
        }
        r_i *= sys.prm.limiter_relaxation_factor;

        const double h_relaxed = 2. * std::abs(h_relaxation_numerator) / (relaxation_denominator + eps_double);
        h_min = std::max((1. - r_i) * h_min, h_min - h_relaxed);
        h_max = std::min((1. + r_i) * h_max, h_max + h_relaxed);

        const double kin_relaxed =
            2. * std::abs(kin_relaxation_numerator) / (relaxation_denominator + eps_double);
        kin_max = std::min((1. + r_i) * kin_max, kin_max + kin_relaxed);

        const double v2_relaxed = 2. * std::abs(v2_relaxation_numerator) / (relaxation_denominator + eps_double);
        v2_max = std::min((1. + r_i) * v2_max, v2_max + v2_relaxed);

        r_i = hd_i;
        if constexpr (dim == 2)
          r_i = std::sqrt(hd_i);
        r_i *= sys.prm.dry_state_relaxation_factor;
        h_small = sys.prm.reference_water_depth * r_i;
        return relaxed;
      }

      double limit(const Bounds &bounds,
                   const state_type &U,
                   const state_type &P,
                   bool &success,
                   const double t_min = 0.,
                   const double t_max = 1.) const
      {
        success = true;
        double t_l = t_min;
        double t_r = t_max;

        const double h_min = bounds[0], h_max = bounds[1], h_small = bounds[2], kin_max = bounds[3],
                     v2_max = bounds[4];

        const double small = sys.prm.dry_state_relaxation_small;
        const double large = sys.prm.dry_state_relaxation_large;
        const double relax_small = 1. + small * eps_double;
        const double relax = 1. + large * eps_double;

        auto add = [](const state_type &A, double t, const state_type &B) {
          state_type R;
          for (int c = 0; c < Sys::k; ++c)
            R[c] = A[c] + t * B[c];
          return R;
        };

        {
          const double h_U = Sys::water_depth(U);
          const double h_P = Sys::water_depth(P);

          const double test_min = sys.filter_dry_water_depth(std::max(0., h_U - relax * h_max));
          const double test_max = sys.filter_dry_water_depth(std::max(0., h_min - relax * h_U));
          if (!(test_min == 0. && test_max == 0.))
            success = false;

          const double denominator = 1. / (std::abs(h_P) + eps_double * h_max + min_double);

          t_r = h_max < h_U + t_r * h_P ? (h_max - h_U) * denominator : t_r;

          const double h_min_tilde = std::max(h_small, h_min);
          t_r = h_U + t_r * h_P < h_min_tilde ? (h_U - h_min_tilde) * denominator : t_r;

          t_r = std::min(t_r, t_max);
          t_r = std::max(t_r, t_min);
        }

        if (!sys.prm.limit_on_square_velocity && !sys.prm.limit_on_kinetic_energy)
          return t_l; /* sic: limiter.template.h:144-146 */

        const double tolerance = sys.prm.limiter_newton_tolerance;

        if (sys.prm.limit_on_kinetic_energy) {
          const auto U_r = add(U, t_r, P);
          const double h_r = Sys::water_depth(U_r);
          const auto q_r = Sys::momentum(U_r);
          const double psi_r = relax_small * h_r * kin_max - 0.5 * norm_square<dim>(q_r);

          t_l = psi_r > 0. ? t_r : t_l;

          if (!sys.prm.limit_on_square_velocity && t_l == t_r)
            return t_l;

          const auto U_l = add(U, t_l, P);
          const double h_l = Sys::water_depth(U_l);
          const auto q_l = Sys::momentum(U_l);
          const double psi_l = relax_small * h_l * kin_max - 0.5 * norm_square<dim>(q_l);

          const double filtered_h_l = sys.filter_dry_water_depth(h_l);
          const double lower_bound = (1. - relax) * filtered_h_l * kin_max - eps_double;
          if (!(std::min(0., psi_l - lower_bound) == 0.))
            success = false;

          if (!(std::max(0., t_r - t_l - tolerance) == 0.)) {
            const double h_P = Sys::water_depth(P);
            const auto q_U = Sys::momentum(U);
            const auto q_P = Sys::momentum(P);
            const double dpsi_l = h_P * kin_max - dot<dim>(q_U, q_P) - dot<dim>(q_P, q_P) * t_l;
            const double dpsi_r = h_P * kin_max - dot<dim>(q_U, q_P) - dot<dim>(q_P, q_P) * t_r;
            quadratic_newton_step(t_l, t_r, psi_l, psi_r, dpsi_l, dpsi_r, -1.);
          }

          if (sys.prm.limit_on_square_velocity) {
            t_r = t_l;
            t_l = t_min;
          }
        }

        if (sys.prm.limit_on_square_velocity) {
          const auto U_r = add(U, t_r, P);
          const double h_r = Sys::water_depth(U_r);
          const auto q_r = Sys::momentum(U_r);
          const double psi_r = relax_small * h_r * h_r * v2_max - norm_square<dim>(q_r);

          t_l = psi_r > 0. ? t_r : t_l;
          if (t_l == t_r)
            return t_l;

          const auto U_l = add(U, t_l, P);
          const double h_l = Sys::water_depth(U_l);
          const auto q_l = Sys::momentum(U_l);
          const double psi_l = relax_small * h_l * h_l * v2_max - norm_square<dim>(q_l);

          const double filtered_h_l = sys.filter_dry_water_depth(h_l);
          const double lower_bound = (1. - relax) * filtered_h_l * filtered_h_l * v2_max - 100. * eps_double;
          if (!(std::min(0., psi_l - lower_bound) == 0.))
            success = false;

          if (!(std::max(0., t_r - t_l - tolerance) == 0.)) {
            const double h_U = Sys::water_depth(U);
            const double h_P = Sys::water_depth(P);
            const auto q_U = Sys::momentum(U);
            const auto q_P = Sys::momentum(P);
            const double dpsi_l =
                (h_U + t_l * h_P) * h_P * v2_max - 2. * (dot<dim>(q_U, q_P) - dot<dim>(q_P, q_P) * t_l);
            const double dpsi_r =
                (h_U + t_r * h_P) * h_P * v2_max - 2. * (dot<dim>(q_U, q_P) - dot<dim>(q_P, q_P) * t_r);
            quadratic_newton_step(t_l, t_r, psi_l, psi_r, dpsi_l, dpsi_r, -1.);
          }
        }

        return t_l;
      }
    };
  } // namespace sw
} // namespace orc
