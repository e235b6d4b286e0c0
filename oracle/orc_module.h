// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_common.h).
//
// Restatement of the explicit hyperbolic update:
//   prepare_state_vector  <- source/hyperbolic_module.template.h:97-193
//   step<stages>          <- source/hyperbolic_module.template.h:236-1211
// in the reference's *scalar* (non-SIMD) row semantics on a plain CSR stencil
// with "diagonal first, then ascending column" row ordering
// (tests/common/sparse_matrix_simd.output).
#pragma once

#include "orc_common.h"

#include <atomic>
#include <functional>
#include <stdexcept>

namespace orc
{
  /* Thrown on an invariant-domain violation when the strategy is
   * raise_exception; hyperbolic_module.h:37-61. */
  struct Restart {
  };

  enum class IDViolationStrategy { warn, raise_exception };

  template <int dim>
  struct BoundaryDescription { /* offline_data.h:66-72 */
    unsigned int i;
    Vec<dim> normal;
    double normal_mass;
    double boundary_mass;
    int id;
    Vec<dim> position;
  };

  struct CouplingDescription { /* offline_data.h:78-80 */
    unsigned int i, col_idx, j;
  };

  /* Host view of what OfflineData hands to the module (offline_data.h:109-264),
   * single rank, local numbering == global numbering. */
  template <int dim>
  struct OfflineData {
    unsigned int n_owned = 0;
    std::vector<unsigned int> row_ptr;    /* n_owned + 1 */
    std::vector<unsigned int> cols;       /* nnz, diag first then ascending */
    std::vector<unsigned int> transposed; /* nnz: flat index of (j,i) */
    std::vector<double> mij;              /* nnz */
    std::vector<double> cij;              /* nnz * dim, entry-major */
    std::vector<double> mi, mi_inv;       /* n_owned */
    double measure_of_omega = 0.;
    std::vector<BoundaryDescription<dim>> boundary_map;
    std::vector<CouplingDescription> coupling_boundary_pairs;
    /* Optional: owner rank of every row.  Emulates an N-rank run of the reference inside one
     * address space (tests of the multi-GPU path): with MPI every rank computes d_ij for ALL
     * couplings to its ghost columns itself -- a ghost index is always larger than any owned
     * index, so "j < i: skip and mirror" (hyperbolic_module.template.h:395-400, :461-476,
     * :511-535) only ever applies to pairs owned by the same rank.  The caller supplies rows
     * whose column order is the owner's local order, so row sums run in the rank's order. */
    std::vector<int> partition;

    /* true if entry (i, j) is filled from (j, i) instead of being computed by row i */
    bool mirrored_from_upper(unsigned int i, unsigned int j) const
    {
      return j < i && (partition.empty() || partition[i] == partition[j]);
    }

    unsigned int row_length(unsigned int i) const
    {
      return row_ptr[i + 1] - row_ptr[i];
    }

    Vec<dim> c(std::size_t e) const
    {
      Vec<dim> r;
      for (int d = 0; d < dim; ++d)
        r[d] = cij[e * dim + d];
      return r;
    }

    void build_transposed()
    {
      transposed.assign(cols.size(), 0u);
      bool symmetric = true;
#pragma omp parallel for schedule(static) reduction(&& : symmetric)
      for (unsigned int i = 0; i < n_owned; ++i)
        for (unsigned int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
          const unsigned int j = cols[e];
          bool found = false;
          for (unsigned int f = row_ptr[j]; f < row_ptr[j + 1]; ++f)
            if (cols[f] == i) {
              transposed[e] = f;
              found = true;
              break;
            }
          if (!found)
            symmetric = false;
        }
      if (!symmetric)
        throw std::runtime_error("oracle: stencil is not symmetric");
    }
  };


  /* Description concept (source/euler/description.h:31-50), reduced to what
   * the hyperbolic update needs. */
  template <typename SystemT,
            typename RiemannSolverT,
            typename IndicatorT,
            typename LimiterT>
  struct Description {
    using System = SystemT;
    using RiemannSolver = RiemannSolverT;
    using Indicator = IndicatorT;
    using Limiter = LimiterT;
  };


  template <typename Desc, int dim>
  class HyperbolicModule
  {
  public:
    using System = typename Desc::System;
    using state_type = typename System::state_type;
    using precomputed_type = typename System::precomputed_type;
    using bounds_type = typename System::bounds_type;
    static constexpr int k = System::k;
    static constexpr int n_prec = System::n_precomputed;
    static constexpr int n_bounds = System::n_bounds;

    /* StateVector = (U, precomputed); AoS like MultiComponentVector
     * (multicomponent_vector.h:266): (U_i)_c at i*k + c. */
    struct StateVector {
      std::vector<double> U;
      std::vector<double> prec;
      void reinit(unsigned int n)
      {
        U.assign(std::size_t(n) * k, 0.);
        prec.assign(std::size_t(n) * n_prec, 0.);
      }
      void swap(StateVector &o)
      {
        U.swap(o.U);
        prec.swap(o.prec);
      }
    };

    using DirichletFunction =
        std::function<state_type(const Vec<dim> &, double)>;

    HyperbolicModule(const OfflineData<dim> &od,
                     const System &sys,
                     DirichletFunction initial_state,
                     std::vector<double> initial_precomputed = {})
        : id_violation_strategy_(IDViolationStrategy::warn)
        , od_(od)
        , sys_(sys)
        , initial_state_(std::move(initial_state))
        , initial_precomputed_(std::move(initial_precomputed))
        , cfl_(0.2)
        , n_restarts_(0)
        , n_warnings_(0)
    {
    }

    /* test hook: replace initial_precomputed_ (bathymetry) after construction */
    void set_initial_precomputed(std::vector<double> v) { initial_precomputed_ = std::move(v); }

    void prepare()
    {
      if (sys_.prm.limiter_iterations > 2)
        throw std::runtime_error(
            "The number of limiter iterations must be between [0,2]");
      const std::size_t n = od_.n_owned, nnz = od_.cols.size();
      alpha_.assign(n, 0.);
      bounds_.assign(n * n_bounds, 0.);
      r_.assign(n * k, 0.);
      dij_.assign(nnz, 0.);
      lij_.assign(nnz, 0.);
      lij_next_.assign(nnz, 0.);
      pij_.assign(nnz * k, 0.);
    }

    static state_type get(const std::vector<double> &v, std::size_t i)
    {
      state_type s;
      for (int c = 0; c < k; ++c)
        s[c] = v[i * k + c];
      return s;
    }

    static void put(std::vector<double> &v, std::size_t i, const state_type &s)
    {
      for (int c = 0; c < k; ++c)
        v[i * k + c] = s[c];
    }

    static precomputed_type getp(const std::vector<double> &v, std::size_t i)
    {
      precomputed_type s;
      for (int c = 0; c < n_prec; ++c)
        s[c] = v[i * n_prec + c];
      return s;
    }

    /* hyperbolic_module.template.h:97-193 */
    void prepare_state_vector(StateVector &sv, double t) const
    {
      for (const auto &entry : od_.boundary_map) {
        if (entry.id == Boundary::do_nothing)
          continue;
        auto U_i = get(sv.U, entry.i);
        auto get_dirichlet_data = [&]() {
          return initial_state_(entry.position, t);
        };
        U_i = sys_.apply_boundary_conditions(
            entry.id, U_i, entry.normal, get_dirichlet_data);
        put(sv.U, entry.i, U_i);
      }

      const unsigned int n = od_.n_owned;
      if constexpr (System::aeos) {
        /* two precomputation cycles (euler_aeos/hyperbolic_system.h:433,843-975): cycle 0 per
         * node, [ghost exchange of the precomputed vector], cycle 1 with a loop over the stencil */
#pragma omp parallel for schedule(static)
        for (unsigned int i = 0; i < n; ++i) {
          if (od_.row_length(i) == 1)
            continue;
          const auto prec_i = sys_.precompute_cycle_0(get(sv.U, i));
          for (int c = 0; c < n_prec; ++c)
            sv.prec[std::size_t(i) * n_prec + c] = prec_i[c];
        }
        std::vector<double> gamma_min(n, 0.);
#pragma omp parallel for schedule(static)
        for (unsigned int i = 0; i < n; ++i) {
          const unsigned int row_length = od_.row_length(i);
          if (row_length == 1)
            continue;
          double g = sv.prec[std::size_t(i) * n_prec + 1];
          for (unsigned int col_idx = 1; col_idx < row_length; ++col_idx) {
            const unsigned int j = od_.cols[od_.row_ptr[i] + col_idx];
            const double gamma_j = sys_.surrogate_gamma(get(sv.U, j), sv.prec[std::size_t(j) * n_prec + 0]);
            g = std::min(g, gamma_j);
          }
          gamma_min[i] = g;
        }
#pragma omp parallel for schedule(static)
        for (unsigned int i = 0; i < n; ++i) {
          if (od_.row_length(i) == 1)
            continue;
          auto prec_i = getp(sv.prec, i);
          prec_i[1] = gamma_min[i];
          sys_.precompute_cycle_1_finish(prec_i, get(sv.U, i));
          for (int c = 0; c < n_prec; ++c)
            sv.prec[std::size_t(i) * n_prec + c] = prec_i[c];
        }
      } else {
#pragma omp parallel for schedule(static)
        for (unsigned int i = 0; i < n; ++i) {
          if (od_.row_length(i) == 1)
            continue;
          const auto U_i = get(sv.U, i);
          const auto prec_i = sys_.precompute(U_i);
          for (int c = 0; c < n_prec; ++c)
            sv.prec[std::size_t(i) * n_prec + c] = prec_i[c];
        }
      }
    }

    /* hyperbolic_module.template.h:236-1211; stage vectors passed as
     * pointers, stages = stage_U.size(). */
    double step(const StateVector &old_sv,
                const std::vector<const StateVector *> &stage_sv,
                const std::vector<double> &stage_weights,
                StateVector &new_sv,
                double tau = 0.,
                double tau_max_in = max_double) const
    {
      using RiemannSolver = typename Desc::RiemannSolver;
      using Indicator = typename Desc::Indicator;
      using Limiter = typename Desc::Limiter;

      const int stages = int(stage_sv.size());
      const auto &old_U = old_sv.U;
      const auto &old_prec = old_sv.prec;
      auto &new_U = new_sv.U;

      const unsigned int n_owned = od_.n_owned;
      const double measure_of_omega_inverse = 1. / od_.measure_of_omega;

      bool restart_needed = false;

      /* Step 2: off-diagonal d_ij (upper triangle), alpha_i; :341-424 */
#pragma omp parallel
      {
        RiemannSolver riemann_solver(sys_);
        Indicator indicator(sys_);
#pragma omp for schedule(static)
        for (unsigned int i = 0; i < n_owned; ++i) {
          const unsigned int row_length = od_.row_length(i);
          if (row_length == 1)
            continue;
          const auto U_i = get(old_U, i);
          indicator.reset(getp(old_prec, i), U_i);

          const std::size_t e0 = od_.row_ptr[i];
          for (unsigned int col_idx = 0; col_idx < row_length; ++col_idx) {
            const std::size_t e = e0 + col_idx;
            const unsigned int j = od_.cols[e];
            const auto U_j = get(old_U, j);
            const auto c_ij = od_.c(e);

            indicator.accumulate(getp(old_prec, j), U_j, c_ij);

            if (col_idx == 0)
              continue;
            if (od_.mirrored_from_upper(i, j))
              continue;

            const double nrm = norm<dim>(c_ij);
            /* Tensor / scalar multiplies by the rounded inverse (dealii::Tensor::operator/=) */
            const double nrm_inverse = 1. / nrm;
            Vec<dim> n_ij;
            for (int d = 0; d < dim; ++d)
              n_ij[d] = c_ij[d] * nrm_inverse;
            double lambda_max;
            if constexpr (System::aeos)
              lambda_max = riemann_solver.compute(U_i, getp(old_prec, i), U_j, getp(old_prec, j), n_ij);
            else
              lambda_max = riemann_solver.compute(U_i, U_j, n_ij);
            dij_[e] = nrm * lambda_max;
          }

          const double hd_i = od_.mi[i] * measure_of_omega_inverse;
          alpha_[i] = indicator.alpha(hd_i);
        }
      }

      /* Step 3: boundary d_ij, symmetrize, diagonal, tau_max; :432-564 */
      double tau_max = tau_max_in;
      {
        RiemannSolver riemann_solver(sys_);
        for (const auto &cp : od_.coupling_boundary_pairs) {
          const unsigned int i = cp.i, col_idx = cp.col_idx, j = cp.j;
          if (od_.mirrored_from_upper(i, j))
            continue;
          const auto U_i = get(old_U, i);
          const auto U_j = get(old_U, j);
          const std::size_t e = od_.row_ptr[i] + col_idx;
          const auto c_ji = od_.c(od_.transposed[e]);
          const double norm_ji = norm<dim>(c_ji);
          const double norm_ji_inverse = 1. / norm_ji;
          Vec<dim> n_ji;
          for (int d = 0; d < dim; ++d)
            n_ji[d] = c_ji[d] * norm_ji_inverse;
          const double d_ij = dij_[e];
          double lambda_max;
          if constexpr (System::aeos)
            lambda_max = riemann_solver.compute(U_j, getp(old_prec, j), U_i, getp(old_prec, i), n_ji);
          else
            lambda_max = riemann_solver.compute(U_j, U_i, n_ji);
          const double d_ji = norm_ji * lambda_max;
          dij_[e] = std::max(d_ij, d_ji);
        }

        double local_tau_max = max_double;
        /* the lower-triangular fill reads upper-triangular entries only, so
         * rows are independent: */
#pragma omp parallel for schedule(static) reduction(min : local_tau_max)
        for (unsigned int i = 0; i < n_owned; ++i) {
          const unsigned int row_length = od_.row_length(i);
          if (row_length == 1)
            continue;
          double d_sum = 0.;
          const std::size_t e0 = od_.row_ptr[i];
          for (unsigned int col_idx = 1; col_idx < row_length; ++col_idx) {
            const std::size_t e = e0 + col_idx;
            const unsigned int j = od_.cols[e];
            if (od_.mirrored_from_upper(i, j))
              dij_[e] = dij_[od_.transposed[e]];
            d_sum -= dij_[e];
          }
          d_sum = std::min(d_sum, -1.e6 * min_double);
          dij_[e0] = d_sum;
          const double mass = od_.mi[i];
          const double tau_i = cfl_ * mass / (-2. * d_sum);
          local_tau_max = std::min(local_tau_max, tau_i);
        }
        tau_max = std::min(tau_max, local_tau_max);
      }

      if (std::isnan(tau_max) || std::isinf(tau_max) || !(tau_max > 0.))
        throw std::runtime_error(
            "I'm sorry, Dave. I'm afraid I can't do that.\nWe crashed.");

      tau = (tau == 0. ? tau_max : tau);
      last_tau_max_ = tau_max;

      /* Step 4: low-order update, bounds, r_i, p_ij; :597-884 */
      {
        double weight = 1.;
        {
          /* -std::accumulate(w.begin(), w.end(), -1.) */
          double acc = -1.;
          for (double w : stage_weights)
            acc += w;
          weight = -acc;
        }

#pragma omp parallel
        {
          Limiter limiter(sys_);
#pragma omp for schedule(static)
          for (unsigned int i = 0; i < n_owned; ++i) {
            const unsigned int row_length = od_.row_length(i);
            if (row_length == 1)
              continue;

            const auto U_i = get(old_U, i);
            auto U_i_new = U_i;

            const double alpha_i = alpha_[i];
            const double m_i = od_.mi[i];
            const double m_i_inv = od_.mi_inv[i];

            using flux_contribution_type = typename System::flux_contribution_type;
            const flux_contribution_type flux_i =
                sys_.flux_contribution(getp(old_prec, i), ipv(i), U_i);
            std::vector<flux_contribution_type> flux_iHs(stages);
            state_type S_iH;
            S_iH.fill(0.);
            for (int s = 0; s < stages; ++s) {
              const auto U_iHs = get(stage_sv[s]->U, i);
              flux_iHs[s] = sys_.flux_contribution(getp(stage_sv[s]->prec, i), ipv(i), U_iHs);
              if constexpr (System::have_source_terms) {
                const auto S = sys_.nodal_source(getp(stage_sv[s]->prec, i), U_iHs, tau);
                for (int c = 0; c < k; ++c)
                  S_iH[c] += stage_weights[s] * S[c];
              }
            }

            state_type S_i;
            S_i.fill(0.);
            state_type F_iH;
            F_iH.fill(0.);

            if constexpr (System::have_source_terms) {
              S_i = sys_.nodal_source(getp(old_prec, i), U_i, tau);
              for (int c = 0; c < k; ++c) {
                S_iH[c] += weight * S_i[c];
                U_i_new[c] += tau * S_i[c];
                F_iH[c] += m_i * S_iH[c];
              }
            }

            if constexpr (System::aeos)
              limiter.reset(getp(old_prec, i), U_i, flux_i);
            else
              limiter.reset(U_i);

            state_type affine_shift;
            affine_shift.fill(0.);

            const std::size_t e0 = od_.row_ptr[i];

            if constexpr (System::shallow_water) {
              for (unsigned int col_idx = 0; col_idx < row_length; ++col_idx) {
                const std::size_t e = e0 + col_idx;
                const unsigned int j = od_.cols[e];
                const auto U_j = get(old_U, j);
                const auto flux_j = sys_.flux_contribution(getp(old_prec, j), ipv(j), U_j);
                const double d_ij = dij_[e];
                const auto c_ij = od_.c(e);
                const auto B_ij = sys_.affine_shift(flux_i, flux_j, c_ij, d_ij);
                for (int c = 0; c < k; ++c)
                  affine_shift[c] += B_ij[c];
              }
              for (int c = 0; c < k; ++c)
                affine_shift[c] *= tau * m_i_inv;
            }

            if constexpr (System::have_source_terms) {
              for (int c = 0; c < k; ++c)
                affine_shift[c] += tau * S_i[c];
            }

            for (unsigned int col_idx = 0; col_idx < row_length; ++col_idx) {
              const std::size_t e = e0 + col_idx;
              const unsigned int j = od_.cols[e];

              const auto U_j = get(old_U, j);
              const double alpha_j = alpha_[j];

              const double d_ij = dij_[e];
              const double factor = (alpha_i + alpha_j) * .5;
              const double d_ijH = d_ij * factor;

              const auto c_ij = od_.c(e);
              const double regularization = 100. * min_double;
              Vec<dim> scaled_c_ij;
              {
                const double inverse = 1. / std::max(d_ij, regularization);
                for (int d = 0; d < dim; ++d)
                  scaled_c_ij[d] = c_ij[d] * inverse;
              }

              const auto flux_j = sys_.flux_contribution(getp(old_prec, j), ipv(j), U_j);
              const double m_ij = od_.mij[e];

              const auto flux_ij = sys_.flux_divergence(flux_i, flux_j, c_ij);
              state_type P_ij;
              for (int c = 0; c < k; ++c) {
                U_i_new[c] += tau * m_i_inv * flux_ij[c];
                P_ij[c] = -flux_ij[c];
              }

              if constexpr (System::shallow_water) {
                const auto eq = sys_.equilibrated_states(flux_i, flux_j);
                const auto &U_star_ij = eq[0];
                const auto &U_star_ji = eq[1];
                for (int c = 0; c < k; ++c) {
                  const double diff = U_star_ji[c] - U_star_ij[c];
                  U_i_new[c] += tau * m_i_inv * d_ij * diff;
                  F_iH[c] += d_ijH * diff;
                  P_ij[c] += (d_ijH - d_ij) * diff;
                }
                limiter.accumulate(U_j, U_star_ij, U_star_ji, scaled_c_ij, affine_shift);
              } else {
                for (int c = 0; c < k; ++c) {
                  const double diff = U_j[c] - U_i[c];
                  U_i_new[c] += tau * m_i_inv * d_ij * diff;
                  F_iH[c] += d_ijH * diff;
                  P_ij[c] += (d_ijH - d_ij) * diff;
                }
                if constexpr (System::aeos)
                  limiter.accumulate(getp(old_prec, j), U_j, flux_j, scaled_c_ij);
                else
                  limiter.accumulate(getp(old_prec, j), U_j, scaled_c_ij);
              }

              if constexpr (System::have_source_terms) {
                for (int c = 0; c < k; ++c) {
                  F_iH[c] -= m_ij * S_iH[c];
                  P_ij[c] -= m_ij * /*sic!*/ S_i[c];
                }
              }

              if constexpr (System::have_high_order_flux) {
                const auto high_order_flux_ij = sys_.high_order_flux_divergence(flux_i, flux_j, c_ij);
                for (int c = 0; c < k; ++c) {
                  F_iH[c] += weight * high_order_flux_ij[c];
                  P_ij[c] += weight * high_order_flux_ij[c];
                }
              } else {
                for (int c = 0; c < k; ++c) {
                  F_iH[c] += weight * flux_ij[c];
                  P_ij[c] += weight * flux_ij[c];
                }
              }

              if constexpr (System::have_source_terms) {
                const auto S_j = sys_.nodal_source(getp(old_prec, j), U_j, tau);
                for (int c = 0; c < k; ++c) {
                  F_iH[c] += weight * m_ij * S_j[c];
                  P_ij[c] += weight * m_ij * S_j[c];
                }
              }

              for (int s = 0; s < stages; ++s) {
                const auto U_jHs = get(stage_sv[s]->U, j);
                const auto flux_jHs = sys_.flux_contribution(getp(stage_sv[s]->prec, j), ipv(j), U_jHs);
                if constexpr (System::have_high_order_flux) {
                  const auto ho = sys_.high_order_flux_divergence(flux_iHs[s], flux_jHs, c_ij);
                  for (int c = 0; c < k; ++c) {
                    F_iH[c] += stage_weights[s] * ho[c];
                    P_ij[c] += stage_weights[s] * ho[c];
                  }
                } else {
                  const auto flux_ij_s = sys_.flux_divergence(flux_iHs[s], flux_jHs, c_ij);
                  for (int c = 0; c < k; ++c) {
                    F_iH[c] += stage_weights[s] * flux_ij_s[c];
                    P_ij[c] += stage_weights[s] * flux_ij_s[c];
                  }
                }
                if constexpr (System::have_source_terms) {
                  const auto S_js = sys_.nodal_source(getp(stage_sv[s]->prec, j), U_jHs, tau);
                  for (int c = 0; c < k; ++c) {
                    F_iH[c] += stage_weights[s] * m_ij * S_js[c];
                    P_ij[c] += stage_weights[s] * m_ij * S_js[c];
                  }
                }
              }

              for (int c = 0; c < k; ++c)
                pij_[e * k + c] = P_ij[c];
            }

            if (sys_.prm.expensive_bounds_check && !sys_.is_admissible(U_i_new)) {
#pragma omp atomic write
              restart_needed = true;
            }

            put(new_U, i, U_i_new);
            put(r_, i, F_iH);

            const double hd_i = m_i * measure_of_omega_inverse;
            const auto relaxed_bounds = limiter.bounds(hd_i);
            for (int b = 0; b < n_bounds; ++b)
              bounds_[std::size_t(i) * n_bounds + b] = relaxed_bounds[b];
          }
        }
      }

      /* Step 5: second part of p_ij, first l_ij; :892-1041 */
      if (sys_.prm.limiter_iterations != 0) {
#pragma omp parallel
        {
          Limiter limiter(sys_);
#pragma omp for schedule(static)
          for (unsigned int i = 0; i < n_owned; ++i) {
            const unsigned int row_length = od_.row_length(i);
            if (row_length == 1)
              continue;

            bounds_type bounds;
            for (int b = 0; b < n_bounds; ++b)
              bounds[b] = bounds_[std::size_t(i) * n_bounds + b];

            const double m_i_inv = od_.mi_inv[i];
            const auto U_i_new = get(new_U, i);
            const auto F_iH = get(r_, i);

            const double lambda_inv = double(row_length - 1);
            const double factor = tau * m_i_inv * lambda_inv;

            const std::size_t e0 = od_.row_ptr[i];
            for (unsigned int col_idx = 1; col_idx < row_length; ++col_idx) {
              const std::size_t e = e0 + col_idx;
              const unsigned int j = od_.cols[e];

              state_type P_ij;
              for (int c = 0; c < k; ++c)
                P_ij[c] = pij_[e * k + c];
              const auto F_jH = get(r_, j);

              const double kronecker_ij = col_idx == 0 ? 1. : 0.;
              const double m_j_inv = od_.mi_inv[j];
              const double m_ij = od_.mij[e];
              const double b_ij = kronecker_ij - m_ij * m_j_inv;
              const double b_ji = kronecker_ij - m_ij * m_i_inv;

              for (int c = 0; c < k; ++c) {
                P_ij[c] += b_ij * F_jH[c] - b_ji * F_iH[c];
                P_ij[c] *= factor;
                pij_[e * k + c] = P_ij[c];
              }

              bool success;
              const double l_ij = limiter.limit(bounds, U_i_new, P_ij, success);
              lij_[e] = l_ij;

              if (!success) {
#pragma omp atomic write
                restart_needed = true;
              }
            }
          }
        }
      }

      /* Steps 6, 7: symmetrize l_ij, high-order update, next l_ij;
       * :1053-1182 */
      const unsigned int n_iterations = sys_.prm.limiter_iterations;
      for (unsigned int pass = 0; pass < n_iterations; ++pass) {
        const bool last_round = (pass + 1 == n_iterations);

        if ((n_iterations == 2) && last_round)
          std::swap(lij_, lij_next_);

#pragma omp parallel
        {
          Limiter limiter(sys_);
          std::vector<double> lij_row;
#pragma omp for schedule(static)
          for (unsigned int i = 0; i < n_owned; ++i) {
            const unsigned int row_length = od_.row_length(i);
            if (row_length == 1)
              continue;

            auto U_i_new = get(new_U, i);
            const double lambda = 1. / double(row_length - 1);
            lij_row.resize(row_length);

            const std::size_t e0 = od_.row_ptr[i];
            for (unsigned int col_idx = 1; col_idx < row_length; ++col_idx) {
              const std::size_t e = e0 + col_idx;
              const double l_ij = std::min(lij_[e], lij_[od_.transposed[e]]);
              for (int c = 0; c < k; ++c)
                U_i_new[c] += l_ij * lambda * pij_[e * k + c];
              if (!last_round)
                lij_row[col_idx] = l_ij;
            }

            if (sys_.prm.expensive_bounds_check && !sys_.is_admissible(U_i_new)) {
#pragma omp atomic write
              restart_needed = true;
            }

            put(new_U, i, U_i_new);

            if (last_round)
              continue;

            bounds_type bounds;
            for (int b = 0; b < n_bounds; ++b)
              bounds[b] = bounds_[std::size_t(i) * n_bounds + b];

            for (unsigned int col_idx = 1; col_idx < row_length; ++col_idx) {
              const std::size_t e = e0 + col_idx;
              const double old_l_ij = lij_row[col_idx];
              state_type new_p_ij;
              for (int c = 0; c < k; ++c)
                new_p_ij[c] = (1. - old_l_ij) * pij_[e * k + c];

              bool success;
              const double new_l_ij =
                  limiter.limit(bounds, U_i_new, new_p_ij, success);
              if (sys_.prm.expensive_bounds_check && !success) {
#pragma omp atomic write
                restart_needed = true;
              }
              lij_next_[e] = (1. - old_l_ij) * new_l_ij;
            }
          }
        }
      }

      /* :1190-1210 */
      if (restart_needed) {
        switch (id_violation_strategy_) {
        case IDViolationStrategy::warn:
          n_warnings_++;
          break;
        case IDViolationStrategy::raise_exception:
          n_restarts_++;
          throw Restart();
        }
      }

      return tau;
    }

    void cfl(double new_cfl) const { cfl_ = new_cfl; }
    double cfl() const { return cfl_; }
    const std::vector<double> &alpha() const { return alpha_; }
    const std::vector<double> &dij() const { return dij_; }
    const std::vector<double> &lij() const { return lij_; }
    const std::vector<double> &lij_next() const { return lij_next_; }
    const std::vector<double> &pij() const { return pij_; }
    const std::vector<double> &r() const { return r_; }
    const std::vector<double> &bounds() const { return bounds_; }
    unsigned int n_restarts() const { return n_restarts_; }
    unsigned int n_warnings() const { return n_warnings_; }
    double last_tau_max() const { return last_tau_max_; }

    mutable IDViolationStrategy id_violation_strategy_;

    const OfflineData<dim> &offline_data() const { return od_; }
    const System &system() const { return sys_; }
    const DirichletFunction &initial_state() const { return initial_state_; }

  private:
    const OfflineData<dim> &od_;
    const System &sys_;
    DirichletFunction initial_state_;
    /* initial_precomputed_ (hyperbolic_module.h:311): one value per node (bathymetry) or empty */
    std::vector<double> initial_precomputed_;

    double ipv(std::size_t i) const { return initial_precomputed_.empty() ? 0. : initial_precomputed_[i]; }

    mutable double cfl_;
    mutable unsigned int n_restarts_, n_warnings_;
    mutable double last_tau_max_ = 0.;

    mutable std::vector<double> alpha_, bounds_, r_;
    mutable std::vector<double> dij_, lij_, lij_next_, pij_;
  };
} // namespace orc
