// Shallow water equation module as __device__ code (second "Description" plug-in,
// SURVEY.md 8a row a18; BASELINE config 5).
//   HyperbolicSystemView  source/shallow_water/hyperbolic_system.h:677-1282
//   RiemannSolver         source/shallow_water/riemann_solver.template.h:24-252
//   Indicator             source/shallow_water/indicator.h:157-222
//   Limiter               source/shallow_water/limiter.h:249-381,
//                         source/shallow_water/limiter.template.h:16-449
// Same plug-in surface as euler.cuh; the kernels special-case the shallow water
// structure exactly where the reference does (`constexpr bool shallow_water`,
// hyperbolic_module.template.h:270-271,700-720,773-800,816-845).
#pragma once

#include "device_common.cuh"

namespace rjb
{
  struct ShallowWaterParams {
    double gravity, manning_friction_coefficient, reference_water_depth;
    double dry_state_relaxation_factor, dry_state_relaxation_small, dry_state_relaxation_large;
    double evc_factor;
    double limiter_newton_tolerance, limiter_relaxation_factor;
    unsigned int limiter_newton_max_iterations;
    int limit_on_kinetic_energy, limit_on_square_velocity;
  };


  template <int dim_>
  struct ShallowWater {
    static constexpr int dim = dim_;
    static constexpr int k = dim + 1;
    static constexpr int n_prec = 2;   /* (mathematical entropy, h_star) */
    static constexpr int n_bounds = 5; /* h_min, h_max, h_small, kin_max, v2_max */
    static constexpr int n_aux = 1;    /* bathymetry Z_i: the per-node part of flux_contribution = (U, Z) */
    static constexpr bool shallow_water = true;
    static constexpr bool aeos = false;
    static constexpr bool have_source_terms = true;
    static constexpr bool have_high_order_flux = true;
    static constexpr bool edge_cache = false;
    using Params = ShallowWaterParams;

    struct Flux {
      double Z;
    };

    RJ_DEVICE static void store_aux(const Flux &f, double (&aux)[n_aux]) { aux[0] = f.Z; }
    RJ_DEVICE static Flux load_aux(const double (&aux)[n_aux])
    {
      Flux f;
      f.Z = aux[0];
      return f;
    }
    RJ_DEVICE static Flux
    flux_contribution(const Params &, const double (&)[n_prec], const double (&)[k], const double Z)
    {
      Flux f;
      f.Z = Z;
      return f;
    }

    /* ---------------- HyperbolicSystemView ---------------- */

    RJ_DEVICE static double inverse_water_depth_mollified(const Params &p, const double (&U)[k])
    {
      const double h_cutoff_mollified = p.reference_water_depth * p.dry_state_relaxation_large * DBL_EPSILON;
      const double h = U[0];
      const double h_pos = positive_part(h);
      const double h_max = fmax(h, h_cutoff_mollified);
      const double denom = h * h + h_max * h_max;
      return 2. * h_pos / denom;
    }

    RJ_DEVICE static double water_depth_sharp(const Params &p, const double (&U)[k])
    {
      const double h_cutoff_small = p.reference_water_depth * p.dry_state_relaxation_small * DBL_EPSILON;
      return fmax(U[0], h_cutoff_small);
    }

    RJ_DEVICE static double inverse_water_depth_sharp(const Params &p, const double (&U)[k])
    {
      return 1. / water_depth_sharp(p, U);
    }

    RJ_DEVICE static double filter_dry_water_depth(const Params &p, const double h)
    {
      const double h_cutoff_large = p.reference_water_depth * p.dry_state_relaxation_large * DBL_EPSILON;
      return fabs(h) < h_cutoff_large ? 0. : h;
    }

    RJ_DEVICE static double kinetic_energy(const Params &p, const double (&U)[k])
    {
      const double hi = inverse_water_depth_sharp(p, U);
      double v2 = 0.;
#pragma unroll
      for (int d = 0; d < dim; ++d) {
        const double v = U[1 + d] * hi;
        v2 = d == 0 ? v * v : v2 + v * v;
      }
      return 0.5 * U[0] * v2;
    }

    RJ_DEVICE static double pressure(const Params &p, const double (&U)[k])
    {
      const double h_sqd = U[0] * U[0];
      return 0.5 * p.gravity * h_sqd;
    }

    RJ_DEVICE static double mathematical_entropy(const Params &p, const double (&U)[k])
    {
      return pressure(p, U) + kinetic_energy(p, U);
    }

    /* precomputation_loop body, hyperbolic_system.h:699-707 */
    RJ_DEVICE static void precompute(const Params &p, const double (&U)[k], double (&prec)[n_prec])
    {
      prec[0] = mathematical_entropy(p, U);
      prec[1] = path_pow(water_depth_sharp(p, U), 4. / 3.);
    }

    /* element (comp, d) of f(U) (with_pressure) or g(U), hyperbolic_system.h:1022-1056 */
    template <bool with_pressure>
    RJ_DEVICE static double
    fg_element(const Params &p, const double (&U)[k], const double h_inverse, const int comp, const int d)
    {
      const double vd = U[1 + d] * h_inverse;
      if (comp == 0)
        return vd * U[0];
      const double v = vd * U[comp];
      if (with_pressure && d == comp - 1)
        return v + pressure(p, U);
      return v;
    }

    RJ_DEVICE static void
    star_state(const Params &p, const double (&U)[k], const double Z_left, const double Z_right, double (&out)[k])
    {
      const double Z_max = fmax(Z_left, Z_right);
      const double H_star = fmax(0., U[0] + Z_left - Z_max);
      const double inv = inverse_water_depth_mollified(p, U);
#pragma unroll
      for (int q = 0; q < k; ++q)
        out[q] = (U[q] * H_star) * inv;
    }

    /* flux_divergence, hyperbolic_system.h:1117-1144 */
    RJ_DEVICE static void flux_divergence(const Params &p,
                                          const double (&U_i)[k],
                                          const Flux &f_i,
                                          const double (&U_j)[k],
                                          const Flux &f_j,
                                          const double (&c)[dim],
                                          double (&out)[k])
    {
      double U_star_ij[k], U_star_ji[k];
      star_state(p, U_i, f_i.Z, f_j.Z, U_star_ij);
      star_state(p, U_j, f_j.Z, f_i.Z, U_star_ji);
      const double H_i = U_i[0];
      const double H_star_ij = U_star_ij[0];
      const double H_star_ji = U_star_ji[0];
      const double hi_ij = inverse_water_depth_sharp(p, U_star_ij);
      const double hi_ji = inverse_water_depth_sharp(p, U_star_ji);
      const double factor = (0.5 * (H_star_ji * H_star_ji - H_star_ij * H_star_ij) + H_i * H_i) * p.gravity;
#pragma unroll
      for (int q = 0; q < k; ++q) {
        double s = 0.;
#pragma unroll
        for (int d = 0; d < dim; ++d) {
          double r = -(fg_element<false>(p, U_star_ij, hi_ij, q, d) + fg_element<false>(p, U_star_ji, hi_ji, q, d));
          if (q >= 1 && d == q - 1)
            r -= factor;
          s = d == 0 ? r * c[d] : s + r * c[d];
        }
        out[q] = s;
      }
    }

    /* high_order_flux_divergence, hyperbolic_system.h:1149-1171 */
    RJ_DEVICE static void high_order_flux_divergence(const Params &p,
                                                     const double (&U_i)[k],
                                                     const Flux &f_i,
                                                     const double (&U_j)[k],
                                                     const Flux &f_j,
                                                     const double (&c)[dim],
                                                     double (&out)[k])
    {
      const double hi_i = inverse_water_depth_sharp(p, U_i);
      const double hi_j = inverse_water_depth_sharp(p, U_j);
      const double factor = p.gravity * U_i[0] * (U_j[0] + f_j.Z - f_i.Z);
#pragma unroll
      for (int q = 0; q < k; ++q) {
        double s = 0.;
#pragma unroll
        for (int d = 0; d < dim; ++d) {
          double r = -(fg_element<false>(p, U_i, hi_i, q, d) + fg_element<false>(p, U_j, hi_j, q, d));
          if (q >= 1 && d == q - 1)
            r -= factor;
          s = d == 0 ? r * c[d] : s + r * c[d];
        }
        out[q] = s;
      }
    }

    /* affine_shift, hyperbolic_system.h:1176-1191 */
    RJ_DEVICE static void affine_shift(const Params &p,
                                       const double (&U_i)[k],
                                       const Flux &f_i,
                                       const Flux &f_j,
                                       const double (&c)[dim],
                                       const double d_ij,
                                       double (&out)[k])
    {
      double U_star_ij[k];
      star_state(p, U_i, f_i.Z, f_j.Z, U_star_ij);
      const double h_inverse = inverse_water_depth_sharp(p, U_i);
      double mc = U_i[1] * c[0];
#pragma unroll
      for (int d = 1; d < dim; ++d)
        mc += U_i[1 + d] * c[d];
      const double factor = 2. * (d_ij + h_inverse * mc);
#pragma unroll
      for (int q = 0; q < k; ++q)
        out[q] = -factor * (U_star_ij[q] - U_i[q]);
    }

    /* manning_friction / nodal_source, hyperbolic_system.h:1196-1248 */
    RJ_DEVICE static void
    nodal_source(const Params &p, const double (&prec)[n_prec], const double (&U)[k], const double tau, double (&out)[k])
    {
      const double h_star = prec[1];
      const double g = p.gravity;
      const double n = p.manning_friction_coefficient;
      const double h_inverse = inverse_water_depth_mollified(p, U);
      double v2 = 0.;
#pragma unroll
      for (int d = 0; d < dim; ++d) {
        const double v = U[1 + d] * h_inverse;
        v2 = d == 0 ? v * v : v2 + v * v;
      }
      const double v_norm = sqrt(v2);
      const double factor = 2. * g * n * n * v_norm;
      const double denominator = h_star + fmax(h_star, tau * factor);
      const double denominator_inverse = 1. / denominator;
      out[0] = 0.;
#pragma unroll
      for (int d = 0; d < dim; ++d)
        out[d + 1] = -factor * denominator_inverse * U[1 + d];
    }

    /* ---------------- boundary conditions, hyperbolic_system.h:935-1017 ---------------- */

    template <int component>
    RJ_DEVICE static void prescribe_riemann_characteristic(const Params &p,
                                                           const double (&U)[k],
                                                           const double (&U_bar)[k],
                                                           const double (&normal)[dim],
                                                           double (&U_new)[k])
    {
      const double a = sqrt(p.gravity * U[0]);
      const double hi = inverse_water_depth_sharp(p, U);
      double mn = 0., mn_bar = 0.;
#pragma unroll
      for (int d = 0; d < dim; ++d) {
        mn = d == 0 ? U[1] * normal[0] : mn + U[1 + d] * normal[d];
        mn_bar = d == 0 ? U_bar[1] * normal[0] : mn_bar + U_bar[1 + d] * normal[d];
      }
      const double vn = mn * hi;
      const double a_bar = sqrt(p.gravity * U_bar[0]);
      const double vn_bar = mn_bar * inverse_water_depth_sharp(p, U_bar);
      const double R_1 = component == 1 ? vn_bar - 2. * a_bar : vn - 2. * a;
      const double R_2 = component == 2 ? vn_bar + 2. * a_bar : vn + 2. * a;
      const double vn_new = 0.5 * (R_1 + R_2);
      const double q = (R_2 - R_1) / 4.;
      const double h_new = (q * q) / p.gravity;
      U_new[0] = h_new;
#pragma unroll
      for (int d = 0; d < dim; ++d) {
        const double vperp = U[1 + d] * hi - vn * normal[d];
        U_new[1 + d] = h_new * (vn_new * normal[d] + vperp);
      }
    }

    template <typename DirichletData>
    RJ_DEVICE static void apply_boundary_conditions(
        const Params &p, const int id, double (&U)[k], const double (&normal)[dim], const DirichletData &D)
    {
      if (id == RJ_BOUNDARY_DIRICHLET) {
        D(U);
      } else if (id == RJ_BOUNDARY_DIRICHLET_MOMENTUM) {
        double Ud[k];
        D(Ud);
#pragma unroll
        for (int d = 0; d < dim; ++d)
          U[1 + d] = Ud[1 + d];
      } else if (id == RJ_BOUNDARY_SLIP) {
        double mn = 0.;
#pragma unroll
        for (int d = 0; d < dim; ++d)
          mn = d == 0 ? U[1] * normal[0] : mn + U[1 + d] * normal[d];
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
          mn = d == 0 ? U[1] * normal[0] : mn + U[1 + d] * normal[d];
        const double h_inverse = inverse_water_depth_sharp(p, U);
        const double a = sqrt(p.gravity * U[0]);
        const double vn = mn * h_inverse;
        double result[k];
#pragma unroll
        for (int q = 0; q < k; ++q)
          result[q] = U[q];
        if (vn < -a)
          D(result);
        if (vn >= -a && vn <= 0.) {
          double Ud[k];
          D(Ud);
          prescribe_riemann_characteristic<2>(p, Ud, U, normal, result);
        }
        if (vn > 0. && vn <= a) {
          double Ud[k];
          D(Ud);
          prescribe_riemann_characteristic<1>(p, U, Ud, normal, result);
        }
#pragma unroll
        for (int q = 0; q < k; ++q)
          U[q] = result[q];
      }
    }

    /* ---------------- RiemannSolver ---------------- */

    struct RiemannData {
      double h, u, a;
    };

    RJ_DEVICE static RiemannData
    riemann_data_from_state(const Params &p, const double (&U)[k], const double (&n)[dim])
    {
      RiemannData rd;
      rd.h = water_depth_sharp(p, U);
      const double h_inverse = 1. / rd.h; /* momentum / h: Tensor / scalar */
      double pv = 0.;
#pragma unroll
      for (int d = 0; d < dim; ++d) {
        const double v = U[1 + d] * h_inverse;
        pv = d == 0 ? n[0] * v : pv + n[d] * v;
      }
      rd.u = pv;
      rd.a = sqrt(rd.h * p.gravity);
      return rd;
    }

    RJ_DEVICE static double riemann_f(const Params &p, const RiemannData &rd_Z, const double h)
    {
      const double left_value = 2. * (sqrt(p.gravity * h) - rd_Z.a);
      const double radicand = 0.5 * p.gravity * (h + rd_Z.h) / (h * rd_Z.h);
      const double right_value = (h - rd_Z.h) * sqrt(radicand);
      return h <= rd_Z.h ? left_value : right_value;
    }

    RJ_DEVICE static double
    riemann_phi(const Params &p, const RiemannData &rd_i, const RiemannData &rd_j, const double h)
    {
      return riemann_f(p, rd_i, h) + riemann_f(p, rd_j, h) + rd_j.u - rd_i.u;
    }

    RJ_DEVICE static double riemann_lambda_max(const Params &p,
                                               const double (&U_i)[k],
                                               const double (&)[n_prec],
                                               const double (&U_j)[k],
                                               const double (&)[n_prec],
                                               const double (&n_ij)[dim])
    {
      const RiemannData rd_i = riemann_data_from_state(p, U_i, n_ij);
      const RiemannData rd_j = riemann_data_from_state(p, U_j, n_ij);

      /* compute_h_star, riemann_solver.template.h:110-204 */
      const double gravity = p.gravity;
      const double gravity_inverse = 1. / gravity;
      const double h_min = fmin(rd_i.h, rd_j.h);
      const double h_max = fmax(rd_i.h, rd_j.h);
      const double a_min = sqrt(gravity * h_min);
      const double a_max = sqrt(gravity * h_max);
      const double sqrt_two = sqrt(2.);
      const double x0 = 9. - 4. * sqrt_two;
      const double phi_value_min = riemann_phi(p, rd_i, rd_j, x0 * h_min);
      const double phi_value_max = riemann_phi(p, rd_i, rd_j, x0 * h_max);
      (void)phi_value_min; /* the reference overwrites the first selection, :193-201 */

      double tmp = 1. + sqrt_two * (rd_i.u - rd_j.u) / (a_min + a_max);
      const double h_star_middle = sqrt(h_min * h_max) * tmp;

      const double left_radicand = 3. * h_min + 2. * sqrt_two * sqrt(h_min * h_max);
      const double right_radicand = sqrt_two * sqrt(gravity_inverse * h_min) * (rd_i.u - rd_j.u);
      tmp = sqrt(positive_part(left_radicand + right_radicand));
      tmp -= sqrt_two * sqrt(h_min);
      const double h_star_right = tmp * tmp;

      const double h_star = phi_value_max < 0. ? h_star_middle : h_star_right;

      /* compute_lambda, :69-104 */
      auto lam = [&](const RiemannData &rd, const double sign) {
        const double factor = positive_part((h_star - rd.h) / rd.h);
        const double half_factor = 0.5 * factor;
        return rd.u + sign * (rd.a * sqrt((1. + half_factor) * (1. + factor)));
      };
      const double lambda1 = lam(rd_i, -1.);
      const double lambda3 = lam(rd_j, 1.);
      return fmax(negative_part(lambda1), positive_part(lambda3));
    }

    /* ---------------- Indicator ---------------- */

    struct Indicator {
      double eta_i, pressure_i, hi_i;
      double d_eta_i[k];
      double left;
      double right[k];

      RJ_DEVICE void reset(const Params &p, const double (&prec_i)[n_prec], const double (&U_i)[k])
      {
        eta_i = prec_i[0];
        hi_i = inverse_water_depth_sharp(p, U_i);
        /* mathematical_entropy_derivative, hyperbolic_system.h:846-875 */
        double v2 = 0.;
#pragma unroll
        for (int d = 0; d < dim; ++d) {
          const double v = U_i[1 + d] * hi_i;
          d_eta_i[1 + d] = v;
          v2 = d == 0 ? v * v : v2 + v * v;
        }
        d_eta_i[0] = p.gravity * U_i[0] - 0.5 * v2;
        pressure_i = pressure(p, U_i);
        left = 0.;
#pragma unroll
        for (int q = 0; q < k; ++q)
          right[q] = 0.;
      }

      RJ_DEVICE void accumulate(const Params &p,
                                const double (&U_i)[k],
                                const double (&)[n_prec],
                                const double (&U_j)[k],
                                const double (&prec_j)[n_prec],
                                const Flux &,
                                const double (&c_ij)[dim])
      {
        const double eta_j = prec_j[0];
        const double hi_j = inverse_water_depth_sharp(p, U_j);
        const double pressure_j = pressure(p, U_j);
        double vc = 0.;
#pragma unroll
        for (int d = 0; d < dim; ++d) {
          const double v = U_j[1 + d] * hi_j;
          vc = d == 0 ? v * c_ij[0] : vc + v * c_ij[d];
        }
        left += (eta_j + pressure_j) * vc;
#pragma unroll
        for (int q = 0; q < k; ++q) {
          double s = 0.;
#pragma unroll
          for (int d = 0; d < dim; ++d) {
            const double diff = fg_element<true>(p, U_j, hi_j, q, d) - fg_element<true>(p, U_i, hi_i, q, d);
            s = d == 0 ? diff * c_ij[0] : s + diff * c_ij[d];
          }
          right[q] += s;
        }
      }

      RJ_DEVICE double alpha(const Params &p, const double hd_i) const
      {
        double my_sum = 0.;
#pragma unroll
        for (int q = 0; q < k; ++q)
          my_sum += d_eta_i[q] * right[q];
        const double numerator = fabs(left - my_sum);
        const double denominator = fabs(left) + fabs(my_sum);
        const double regularization = 100. * DBL_MIN;
        const double quotient = fabs(numerator) / (denominator + fmax(hd_i * fabs(eta_i), regularization));
        return fmin(1., p.evc_factor * quotient);
      }
    };

    /* ---------------- Limiter ---------------- */

    struct LimiterAccumulator {
      double h_min, h_max, kin_max, v2_max;
      double h_relaxation_numerator, kin_relaxation_numerator, v2_relaxation_numerator, relaxation_denominator;

      RJ_DEVICE void reset()
      {
        h_min = DBL_MAX;
        h_max = 0.;
        kin_max = 0.;
        v2_max = 0.;
        h_relaxation_numerator = 0.;
        kin_relaxation_numerator = 0.;
        v2_relaxation_numerator = 0.;
        relaxation_denominator = 0.;
      }

      /* limiter.h:271-330 */
      RJ_DEVICE void accumulate_sw(const Params &p,
                                   const double (&U_i)[k],
                                   const double (&U_j)[k],
                                   const double (&U_star_ij)[k],
                                   const double (&U_star_ji)[k],
                                   const double (&scaled_c_ij)[dim],
                                   const double (&affine_shift)[k])
      {
        const double hi_ij = inverse_water_depth_sharp(p, U_star_ij);
        const double hi_ji = inverse_water_depth_sharp(p, U_star_ji);
        double U_ij_bar[k];
#pragma unroll
        for (int q = 0; q < k; ++q) {
          double s = 0.;
#pragma unroll
          for (int d = 0; d < dim; ++d) {
            const double diff =
                fg_element<true>(p, U_star_ij, hi_ij, q, d) + (-fg_element<true>(p, U_star_ji, hi_ji, q, d));
            s = d == 0 ? diff * scaled_c_ij[0] : s + diff * scaled_c_ij[d];
          }
          U_ij_bar[q] = 0.5 * (U_star_ij[q] + U_star_ji[q] + s) + affine_shift[q];
        }

        const double h_bar_ij = U_ij_bar[0];
        h_min = fmin(h_min, h_bar_ij);
        h_max = fmax(h_max, h_bar_ij);
        kin_max = fmax(kin_max, kinetic_energy(p, U_ij_bar));

        auto v2_mollified = [&](const double(&V)[k]) {
          const double inv = inverse_water_depth_mollified(p, V);
          double v2 = 0.;
#pragma unroll
          for (int d = 0; d < dim; ++d) {
            const double v = V[1 + d] * inv;
            v2 = d == 0 ? v * v : v2 + v * v;
          }
          return v2;
        };
        v2_max = fmax(v2_max, v2_mollified(U_ij_bar));

        relaxation_denominator += 1.;
        h_relaxation_numerator += (U_i[0] + U_j[0]);
        kin_relaxation_numerator += (kinetic_energy(p, U_i) + kinetic_energy(p, U_j));
        v2_relaxation_numerator += (-v2_mollified(U_i) + v2_mollified(U_j));
      }

      /* limiter.h:333-381 */
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

        const double h_relaxed = 2. * fabs(h_relaxation_numerator) / (relaxation_denominator + DBL_EPSILON);
        b[0] = fmax((1. - r_i) * h_min, h_min - h_relaxed);
        b[1] = fmin((1. + r_i) * h_max, h_max + h_relaxed);
        const double kin_relaxed = 2. * fabs(kin_relaxation_numerator) / (relaxation_denominator + DBL_EPSILON);
        b[3] = fmin((1. + r_i) * kin_max, kin_max + kin_relaxed);
        const double v2_relaxed = 2. * fabs(v2_relaxation_numerator) / (relaxation_denominator + DBL_EPSILON);
        b[4] = fmin((1. + r_i) * v2_max, v2_max + v2_relaxed);

        double r2 = hd_i;
        if (dim == 2)
          r2 = sqrt(hd_i);
        r2 *= p.dry_state_relaxation_factor;
        b[2] = p.reference_water_depth * r2;
      }
    };

    /* Limiter::limit, limiter.template.h:16-449 (EXPENSIVE_BOUNDS_CHECK off) */
    RJ_DEVICE static double
    limit(const Params &p, const double (&bounds)[n_bounds], const double (&U)[k], const double (&P)[k], bool &success)
    {
      const double t_min = 0., t_max = 1.;
      success = true;
      double t_l = t_min;
      double t_r = t_max;
      const double h_min = bounds[0], h_max = bounds[1], h_small = bounds[2], kin_max = bounds[3], v2_max = bounds[4];
      const double relax_small = 1. + p.dry_state_relaxation_small * DBL_EPSILON;
      const double relax = 1. + p.dry_state_relaxation_large * DBL_EPSILON;

      {
        const double h_U = U[0];
        const double h_P = P[0];
        const double test_min = filter_dry_water_depth(p, fmax(0., h_U - relax * h_max));
        const double test_max = filter_dry_water_depth(p, fmax(0., h_min - relax * h_U));
        if (!(test_min == 0. && test_max == 0.))
          success = false;
        const double denominator = 1. / (fabs(h_P) + DBL_EPSILON * h_max + DBL_MIN);
        t_r = h_max < h_U + t_r * h_P ? (h_max - h_U) * denominator : t_r;
        const double h_min_tilde = fmax(h_small, h_min);
        t_r = h_U + t_r * h_P < h_min_tilde ? (h_U - h_min_tilde) * denominator : t_r;
        t_r = fmin(t_r, t_max);
        t_r = fmax(t_r, t_min);
      }

      if (!p.limit_on_square_velocity && !p.limit_on_kinetic_energy)
        return t_l; /* sic: limiter.template.h:144-146 */

      auto q2 = [&](const double t) {
        double s = 0.;
#pragma unroll
        for (int d = 0; d < dim; ++d) {
          const double q = U[1 + d] + t * P[1 + d];
          s = d == 0 ? q * q : s + q * q;
        }
        return s;
      };
      double qUqP = 0., qPqP = 0.;
#pragma unroll
      for (int d = 0; d < dim; ++d) {
        qUqP = d == 0 ? U[1] * P[1] : qUqP + U[1 + d] * P[1 + d];
        qPqP = d == 0 ? P[1] * P[1] : qPqP + P[1 + d] * P[1 + d];
      }

      if (p.limit_on_kinetic_energy) {
        const double h_r = U[0] + t_r * P[0];
        const double psi_r = relax_small * h_r * kin_max - 0.5 * q2(t_r);
        t_l = psi_r > 0. ? t_r : t_l;
        if (!p.limit_on_square_velocity && t_l == t_r)
          return t_l;
        const double h_l = U[0] + t_l * P[0];
        const double psi_l = relax_small * h_l * kin_max - 0.5 * q2(t_l);
        const double filtered_h_l = filter_dry_water_depth(p, h_l);
        const double lower_bound = (1. - relax) * filtered_h_l * kin_max - DBL_EPSILON;
        if (!(fmin(0., psi_l - lower_bound) == 0.))
          success = false;
        if (!(fmax(0., t_r - t_l - p.limiter_newton_tolerance) == 0.)) {
          const double dpsi_l = P[0] * kin_max - qUqP - qPqP * t_l;
          const double dpsi_r = P[0] * kin_max - qUqP - qPqP * t_r;
          quadratic_newton_step(t_l, t_r, psi_l, psi_r, dpsi_l, dpsi_r, -1.);
        }
        if (p.limit_on_square_velocity) {
          t_r = t_l;
          t_l = t_min;
        }
      }

      if (p.limit_on_square_velocity) {
        const double h_r = U[0] + t_r * P[0];
        const double psi_r = relax_small * h_r * h_r * v2_max - q2(t_r);
        t_l = psi_r > 0. ? t_r : t_l;
        if (t_l == t_r)
          return t_l;
        const double h_l = U[0] + t_l * P[0];
        const double psi_l = relax_small * h_l * h_l * v2_max - q2(t_l);
        const double filtered_h_l = filter_dry_water_depth(p, h_l);
        const double lower_bound = (1. - relax) * filtered_h_l * filtered_h_l * v2_max - 100. * DBL_EPSILON;
        if (!(fmin(0., psi_l - lower_bound) == 0.))
          success = false;
        if (!(fmax(0., t_r - t_l - p.limiter_newton_tolerance) == 0.)) {
          const double dpsi_l = (U[0] + t_l * P[0]) * P[0] * v2_max - 2. * (qUqP - qPqP * t_l);
          const double dpsi_r = (U[0] + t_r * P[0]) * P[0] * v2_max - 2. * (qUqP - qPqP * t_r);
          quadratic_newton_step(t_l, t_r, psi_l, psi_r, dpsi_l, dpsi_r, -1.);
        }
      }
      return t_l;
    }

    /* ---------------- initial states ---------------- */

    template <class IV>
    RJ_DEVICE static void initial_state_compute(
        const Params &p, const IV &iv, const double (&x)[dim], const double t, double (&U)[k])
    {
#pragma unroll
      for (int q = 0; q < k; ++q)
        U[q] = 0.;
      if (iv.kind == RJ_INITIAL_SW_CIRCULAR_DAM_BREAK) {
        /* shallow_water/initial_state_circular_dam_break.h:45-51; sw = (still depth, radius, amplitude) */
        double r = 0.;
#pragma unroll
        for (int d = 0; d < dim; ++d)
          r = d == 0 ? x[0] * x[0] : r + x[d] * x[d];
        U[0] = (r <= iv.sw[1] ? iv.sw[2] : iv.sw[0]);
      } else if (iv.kind == RJ_INITIAL_SW_RITTER_DAM_BREAK) {
        /* shallow_water/initial_state_ritter_dam_break.h:59-80; sw = (time initial, left depth) */
        const double g = p.gravity;
        const double t_initial = iv.sw[0], left_depth = iv.sw[1];
        const double xx = x[0];
        const double aL = sqrt(g * left_depth);
        const double xA = -(t + t_initial) * aL;
        const double xB = 2. * (t + t_initial) * aL;
        const double tmp = aL - xx / (2. * (t + t_initial));
        const double h_expansion = 4. / (9. * g) * tmp * tmp;
        const double v_expansion = 2. / 3. * (xx / (t + t_initial) + aL);
        if (xx <= xA) {
          U[0] = left_depth;
        } else if (xx <= xB) {
          U[0] = h_expansion;
          U[1] = h_expansion * v_expansion;
        }
      } else if (iv.kind == RJ_INITIAL_SW_SMOOTH_VORTEX) {
        /* shallow_water/initial_state_smooth_vortex.h:62-88; sw = (reference depth, mach, beta) */
        const double depth = iv.sw[0], mach = iv.sw[1], beta = iv.sw[2];
        const double pb0 = x[0] - mach * t;
        const double pb1 = dim >= 2 ? x[dim >= 2 ? 1 : 0] : 0.;
        const double r_square = pb0 * pb0 + pb1 * pb1;
        const double factor = beta / (2. * 3.14159265358979323846) * exp(0.5 - 0.5 * r_square);
        const double h = depth - (1. / (2. * p.gravity)) * factor * factor;
        const double u = mach - factor * pb1;
        const double v = factor * pb0;
        U[0] = h;
        U[1] = h * u;
        if (dim >= 2)
          U[dim >= 2 ? 2 : 1] = h * v;
      } else {
        /* shallow_water/initial_state_uniform.h; sw = (h, u) */
        U[0] = iv.sw[0];
        U[1] = iv.sw[1] * iv.sw[0];
      }
    }
  };
} // namespace rjb
