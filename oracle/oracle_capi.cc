// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_common.h).
//
// extern "C" surface of the CPU oracle, bound from Python with ctypes
// (tests/oracle_binding.py). Everything here is a thin adapter around the
// restatements in orc_*.h.

#include "orc_euler.h"
#include "orc_euler_aeos.h"
#include "orc_integrator.h"
#include "orc_mesh.h"
#include "orc_mesh_quads.h"
#include "orc_module.h"
#include "orc_sparsity.h"

#include <cstring>
#include <memory>
#include <string>

#ifdef _OPENMP
#include <omp.h>
#endif

using namespace orc;

namespace
{
  thread_local std::string last_error;

  template <typename F>
  int guarded(const F &f)
  {
    try {
      f();
      return 0;
    } catch (const Restart &) {
      return 1;
    } catch (const std::exception &e) {
      last_error = e.what();
      return -1;
    }
  }

  struct OfflineBase {
    virtual ~OfflineBase() = default;
    int dimension = 0;
  };

  template <int dim>
  struct OfflineHolder : OfflineBase {
    OfflineHolder() { this->dimension = dim; }
    CartesianMesh<dim> mesh; /* only valid if has_mesh */
    bool has_mesh = false;
    OfflineData<dim> od;
  };

  /* layout: oracle/orc.py AEOS_PARAM_FIELDS */
  euler_aeos::Parameters make_aeos_params(const double *p)
  {
    euler_aeos::Parameters prm;
    prm.eos = euler_aeos::Eos(int(p[0]));
    prm.gamma = p[1];
    prm.covolume_b = p[2];
    prm.reference_pressure = p[3];
    prm.reference_specific_internal_energy = p[4];
    prm.vdw_a = p[5];
    prm.compute_strict_bounds = p[6] != 0.;
    prm.reference_density = p[7];
    prm.vacuum_state_relaxation_small = p[8];
    prm.vacuum_state_relaxation_large = p[9];
    prm.indicator_evc_factor = p[10];
    prm.limiter_iterations = (unsigned int)p[11];
    prm.limiter_newton_tolerance = p[12];
    prm.limiter_newton_max_iterations = (unsigned int)p[13];
    prm.limiter_relaxation_factor = p[14];
    prm.expensive_bounds_check = p[15] != 0.;
    return prm;
  }

  euler::Parameters make_params(const double *p)
  {
    /* layout documented in tests/oracle_binding.py: EULER_PARAM_FIELDS */
    euler::Parameters prm;
    prm.gamma = p[0];
    prm.reference_density = p[1];
    prm.vacuum_state_relaxation_small = p[2];
    prm.vacuum_state_relaxation_large = p[3];
    prm.riemann_newton_tolerance = p[4];
    prm.riemann_newton_max_iterations = (unsigned int)p[5];
    prm.indicator_evc_factor = p[6];
    prm.limiter_iterations = (unsigned int)p[7];
    prm.limiter_newton_tolerance = p[8];
    prm.limiter_newton_max_iterations = (unsigned int)p[9];
    prm.limiter_relaxation_factor = p[10];
    prm.expensive_bounds_check = p[11] != 0.;
    return prm;
  }

  sw::Parameters make_sw_params(const double *p)
  {
    /* layout documented in oracle/orc.py: SW_PARAM_FIELDS */
    sw::Parameters prm;
    prm.gravity = p[0];
    prm.manning_friction_coefficient = p[1];
    prm.reference_water_depth = p[2];
    prm.dry_state_relaxation_factor = p[3];
    prm.dry_state_relaxation_small = p[4];
    prm.dry_state_relaxation_large = p[5];
    prm.indicator_evc_factor = p[6];
    prm.limiter_iterations = (unsigned int)p[7];
    prm.limiter_newton_tolerance = p[8];
    prm.limiter_newton_max_iterations = (unsigned int)p[9];
    prm.limiter_relaxation_factor = p[10];
    prm.limit_on_kinetic_energy = p[11] != 0.;
    prm.limit_on_square_velocity = p[12] != 0.;
    prm.expensive_bounds_check = p[13] != 0.;
    return prm;
  }

  struct SimBase {
    virtual ~SimBase() = default;
    virtual int k() const = 0;
    virtual int n_bounds() const = 0;
    virtual int n_precomputed() const = 0;
    virtual unsigned int n() const = 0;
    virtual std::size_t nnz() const = 0;
    virtual void interpolate(double t, int slot) = 0;
    virtual void set_U(int slot, const double *U) = 0;
    virtual void get_U(int slot, double *U) const = 0;
    virtual void get_prec(int slot, double *p) const = 0;
    virtual void prepare_state(int slot, double t) = 0;
    virtual double hyp_step(int old_slot,
                            int stages,
                            const int *stage_slots,
                            const double *weights,
                            int new_slot,
                            double tau,
                            double tau_max) = 0;
    virtual double ti_step(double t, double t_final) = 0;
    virtual void set_cfl(double cfl) = 0;
    virtual void set_strategy(int raise) = 0;
    virtual void configure_ti(int scheme, double cfl_min, double cfl_max, int recovery) = 0;
    virtual void get_scratch(int what, double *out) const = 0;
    virtual void counters(unsigned int *restarts, unsigned int *warnings) const = 0;
    virtual double last_tau_max() const = 0;
    virtual void errors(double t, double *out3) = 0;
    virtual void set_initial_precomputed(const double *) { throw std::runtime_error("oracle: no initial precomputed vector"); }
    unsigned int error_mask = ~0u; /* bit c: include component c in the error norms */
  };

  template <int dim>
  struct EulerTraits {
    using Sys = euler::System<dim>;
    using Prm = euler::Parameters;
    using Desc = Description<Sys, euler::RiemannSolver<dim>, euler::Indicator<dim>, euler::Limiter<dim>>;
  };
  template <int dim>
  struct AeosTraits {
    using Sys = euler_aeos::System<dim>;
    using Prm = euler_aeos::Parameters;
    using Desc = Description<Sys, euler_aeos::RiemannSolver<dim>, euler_aeos::Indicator<dim>, euler_aeos::Limiter<dim>>;
  };

  template <int dim, class Traits>
  struct GasSim : SimBase {
    using Sys = typename Traits::Sys;
    using Desc = typename Traits::Desc;
    using Module = HyperbolicModule<Desc, dim>;
    using SV = typename Module::StateVector;

    OfflineHolder<dim> *off;
    Sys sys;
    euler::InitialValues<dim, Sys> iv;
    std::unique_ptr<Module> module;
    std::unique_ptr<TimeIntegrator<Desc, dim>> ti;
    std::vector<SV> slots; /* slot 0 = the state vector of the time loop */

    GasSim(OfflineHolder<dim> *o, const typename Traits::Prm &prm)
        : off(o)
        , sys(prm)
        , iv(sys)
    {
    }

    void finalize()
    {
      iv.transform.normalize();
      module.reset(new Module(off->od, sys, [this](const Vec<dim> &x, double t) {
        return iv.initial_state(x, t);
      }));
      module->prepare();
      ti.reset(new TimeIntegrator<Desc, dim>(*module));
      ti->prepare();
      slots.resize(6);
      for (auto &s : slots)
        s.reinit(off->od.n_owned);
    }

    int k() const override { return Sys::k; }
    int n_bounds() const override { return Sys::n_bounds; }
    int n_precomputed() const override { return Sys::n_precomputed; }
    unsigned int n() const override { return off->od.n_owned; }
    std::size_t nnz() const override { return off->od.cols.size(); }

    void positions_required() const
    {
      if (!off->has_mesh)
        throw std::runtime_error("oracle: offline data has no mesh positions");
    }

    void interpolate(double t, int slot) override
    {
      positions_required();
      auto &U = slots[slot].U;
      for (unsigned int i = 0; i < off->od.n_owned; ++i) {
        const auto s = iv.initial_state(off->mesh.position(i), t);
        for (int c = 0; c < Sys::k; ++c)
          U[std::size_t(i) * Sys::k + c] = s[c];
      }
    }

    void set_U(int slot, const double *U) override
    {
      std::memcpy(slots[slot].U.data(), U, slots[slot].U.size() * sizeof(double));
    }
    void get_U(int slot, double *U) const override
    {
      std::memcpy(U, slots[slot].U.data(), slots[slot].U.size() * sizeof(double));
    }
    void get_prec(int slot, double *p) const override
    {
      std::memcpy(p, slots[slot].prec.data(), slots[slot].prec.size() * sizeof(double));
    }
    void prepare_state(int slot, double t) override
    {
      module->prepare_state_vector(slots[slot], t);
    }
    double hyp_step(int old_slot,
                    int stages,
                    const int *stage_slots,
                    const double *weights,
                    int new_slot,
                    double tau,
                    double tau_max) override
    {
      std::vector<const SV *> st;
      std::vector<double> w;
      for (int s = 0; s < stages; ++s) {
        st.push_back(&slots[stage_slots[s]]);
        w.push_back(weights[s]);
      }
      return module->step(slots[old_slot], st, w, slots[new_slot], tau, tau_max);
    }
    double ti_step(double t, double t_final) override
    {
      return ti->step(slots[0], t, t_final);
    }
    void set_cfl(double cfl) override { module->cfl(cfl); }
    void set_strategy(int raise) override
    {
      module->id_violation_strategy_ = raise ? IDViolationStrategy::raise_exception
                                             : IDViolationStrategy::warn;
    }
    void configure_ti(int scheme, double cfl_min, double cfl_max, int recovery) override
    {
      ti->time_stepping_scheme_ = TimeSteppingScheme(scheme);
      ti->cfl_min_ = cfl_min;
      ti->cfl_max_ = cfl_max;
      ti->cfl_recovery_strategy_ = CFLRecoveryStrategy(recovery);
      ti->prepare();
    }
    void get_scratch(int what, double *out) const override
    {
      const std::vector<double> *v = nullptr;
      switch (what) {
      case 0: v = &module->alpha(); break;
      case 1: v = &module->dij(); break;
      case 2: v = &module->lij(); break;
      case 3: v = &module->lij_next(); break;
      case 4: v = &module->pij(); break;
      case 5: v = &module->r(); break;
      case 6: v = &module->bounds(); break;
      default: throw std::runtime_error("oracle: unknown scratch id");
      }
      std::memcpy(out, v->data(), v->size() * sizeof(double));
    }
    void counters(unsigned int *restarts, unsigned int *warnings) const override
    {
      *restarts = module->n_restarts();
      *warnings = module->n_warnings();
    }
    double last_tau_max() const override { return module->last_tau_max(); }

    /* time_loop.template.h:690-832 with all components, normalized */
    void errors(double t, double *out3) override
    {
      positions_required();
      module->prepare_state_vector(slots[0], t);
      const unsigned int n = off->od.n_owned;
      std::vector<double> analytic(std::size_t(n) * Sys::k);
      for (unsigned int i = 0; i < n; ++i) {
        const auto s = iv.initial_state(off->mesh.position(i), t);
        for (int c = 0; c < Sys::k; ++c)
          analytic[std::size_t(i) * Sys::k + c] = s[c];
      }
      double linf = 0., l1 = 0., l2 = 0.;
      std::vector<double> a(n), e(n);
      for (int c = 0; c < Sys::k; ++c) {
        if (!((error_mask >> c) & 1u))
          continue;
        for (unsigned int i = 0; i < n; ++i) {
          a[i] = analytic[std::size_t(i) * Sys::k + c];
          e[i] = slots[0].U[std::size_t(i) * Sys::k + c] - a[i];
        }
        const auto na = q1_norms<dim>(off->mesh, a);
        const auto ne = q1_norms<dim>(off->mesh, e);
        linf += ne[0] / na[0];
        l1 += ne[1] / na[1];
        l2 += ne[2] / na[2];
      }
      out3[0] = linf;
      out3[1] = l1;
      out3[2] = l2;
    }
  };
  template <int dim>
  using EulerSim = GasSim<dim, EulerTraits<dim>>;
  template <int dim>
  using AeosSim = GasSim<dim, AeosTraits<dim>>;

  template <int dim>
  struct SwSim : SimBase {
    using Sys = sw::System<dim>;
    using Desc = Description<Sys,
                             sw::RiemannSolver<dim>,
                             sw::Indicator<dim>,
                             sw::Limiter<dim>>;
    using Module = HyperbolicModule<Desc, dim>;
    using SV = typename Module::StateVector;

    OfflineHolder<dim> *off;
    Sys sys;
    sw::InitialValues<dim> iv;
    std::unique_ptr<Module> module;
    std::unique_ptr<TimeIntegrator<Desc, dim>> ti;
    std::vector<SV> slots; /* slot 0 = the state vector of the time loop */

    SwSim(OfflineHolder<dim> *o, const sw::Parameters &prm)
        : off(o)
        , sys(prm)
        , iv(sys)
    {
    }

    void finalize()
    {
      iv.transform.normalize();
      /* InitialValues::interpolate_initial_precomputed_vector: bathymetry of the configuration
       * at the nodes (flat when the offline data came without a mesh) */
      std::vector<double> z(off->od.n_owned, 0.);
      if (off->has_mesh)
        for (unsigned int i = 0; i < off->od.n_owned; ++i)
          z[i] = iv.bathymetry(off->mesh.position(i));
      module.reset(new Module(
          off->od,
          sys,
          [this](const Vec<dim> &x, double t) { return iv.initial_state(x, t); },
          std::move(z)));
      module->prepare();
      ti.reset(new TimeIntegrator<Desc, dim>(*module));
      ti->prepare();
      slots.resize(6);
      for (auto &s : slots)
        s.reinit(off->od.n_owned);
    }

    int k() const override { return Sys::k; }
    int n_bounds() const override { return Sys::n_bounds; }
    int n_precomputed() const override { return Sys::n_precomputed; }
    unsigned int n() const override { return off->od.n_owned; }
    std::size_t nnz() const override { return off->od.cols.size(); }

    void set_initial_precomputed(const double *z) override
    {
      module->set_initial_precomputed(std::vector<double>(z, z + off->od.n_owned));
    }

    void positions_required() const
    {
      if (!off->has_mesh)
        throw std::runtime_error("oracle: offline data has no mesh positions");
    }

    void interpolate(double t, int slot) override
    {
      positions_required();
      auto &U = slots[slot].U;
      for (unsigned int i = 0; i < off->od.n_owned; ++i) {
        const auto s = iv.initial_state(off->mesh.position(i), t);
        for (int c = 0; c < Sys::k; ++c)
          U[std::size_t(i) * Sys::k + c] = s[c];
      }
    }

    void set_U(int slot, const double *U) override
    {
      std::memcpy(slots[slot].U.data(), U, slots[slot].U.size() * sizeof(double));
    }
    void get_U(int slot, double *U) const override
    {
      std::memcpy(U, slots[slot].U.data(), slots[slot].U.size() * sizeof(double));
    }
    void get_prec(int slot, double *p) const override
    {
      std::memcpy(p, slots[slot].prec.data(), slots[slot].prec.size() * sizeof(double));
    }
    void prepare_state(int slot, double t) override
    {
      module->prepare_state_vector(slots[slot], t);
    }
    double hyp_step(int old_slot,
                    int stages,
                    const int *stage_slots,
                    const double *weights,
                    int new_slot,
                    double tau,
                    double tau_max) override
    {
      std::vector<const SV *> st;
      std::vector<double> w;
      for (int s = 0; s < stages; ++s) {
        st.push_back(&slots[stage_slots[s]]);
        w.push_back(weights[s]);
      }
      return module->step(slots[old_slot], st, w, slots[new_slot], tau, tau_max);
    }
    double ti_step(double t, double t_final) override
    {
      return ti->step(slots[0], t, t_final);
    }
    void set_cfl(double cfl) override { module->cfl(cfl); }
    void set_strategy(int raise) override
    {
      module->id_violation_strategy_ = raise ? IDViolationStrategy::raise_exception
                                             : IDViolationStrategy::warn;
    }
    void configure_ti(int scheme, double cfl_min, double cfl_max, int recovery) override
    {
      ti->time_stepping_scheme_ = TimeSteppingScheme(scheme);
      ti->cfl_min_ = cfl_min;
      ti->cfl_max_ = cfl_max;
      ti->cfl_recovery_strategy_ = CFLRecoveryStrategy(recovery);
      ti->prepare();
    }
    void get_scratch(int what, double *out) const override
    {
      const std::vector<double> *v = nullptr;
      switch (what) {
      case 0: v = &module->alpha(); break;
      case 1: v = &module->dij(); break;
      case 2: v = &module->lij(); break;
      case 3: v = &module->lij_next(); break;
      case 4: v = &module->pij(); break;
      case 5: v = &module->r(); break;
      case 6: v = &module->bounds(); break;
      default: throw std::runtime_error("oracle: unknown scratch id");
      }
      std::memcpy(out, v->data(), v->size() * sizeof(double));
    }
    void counters(unsigned int *restarts, unsigned int *warnings) const override
    {
      *restarts = module->n_restarts();
      *warnings = module->n_warnings();
    }
    double last_tau_max() const override { return module->last_tau_max(); }

    /* time_loop.template.h:690-832 with all components, normalized */
    void errors(double t, double *out3) override
    {
      positions_required();
      module->prepare_state_vector(slots[0], t);
      const unsigned int n = off->od.n_owned;
      std::vector<double> analytic(std::size_t(n) * Sys::k);
      for (unsigned int i = 0; i < n; ++i) {
        const auto s = iv.initial_state(off->mesh.position(i), t);
        for (int c = 0; c < Sys::k; ++c)
          analytic[std::size_t(i) * Sys::k + c] = s[c];
      }
      double linf = 0., l1 = 0., l2 = 0.;
      std::vector<double> a(n), e(n);
      for (int c = 0; c < Sys::k; ++c) {
        if (!((error_mask >> c) & 1u))
          continue;
        for (unsigned int i = 0; i < n; ++i) {
          a[i] = analytic[std::size_t(i) * Sys::k + c];
          e[i] = slots[0].U[std::size_t(i) * Sys::k + c] - a[i];
        }
        const auto na = q1_norms<dim>(off->mesh, a);
        const auto ne = q1_norms<dim>(off->mesh, e);
        linf += ne[0] / na[0];
        l1 += ne[1] / na[1];
        l2 += ne[2] / na[2];
      }
      out3[0] = linf;
      out3[1] = l1;
      out3[2] = l2;
    }
  };
} // namespace


namespace
{
  template <int D>
  void postprocess_impl(const OfflineData<D> &od, int k, const double *U, int mode, int first, double beta, double *out, double *bounds)
  {
      
  const unsigned int n = od.n_owned;
  for (unsigned int i = 0; i < n; ++i) {
    out[i] = 0.;
    const unsigned int row_length = od.row_length(i);
    if (row_length == 1)
      continue;
    double acc[3] = {0., 0., 0.};
    for (unsigned int col_idx = 0; col_idx < row_length; ++col_idx) {
      const std::size_t e = od.row_ptr[i] + col_idx;
      const unsigned int j = od.cols[e];
      const auto c_ij = od.c(e);
      if (mode == 0) {
        const double q_j = U[std::size_t(j) * k + first];
        for (int d = 0; d < D; ++d)
          acc[d] -= c_ij[d] * q_j;
      } else {
        double q_j[3] = {0., 0., 0.};
        for (int d = 0; d < D; ++d)
          q_j[d] = U[std::size_t(j) * k + first + d];
        if (D == 2)
          acc[0] -= c_ij[D - 1] * q_j[0] - c_ij[0] * q_j[D - 1];
        else if (D == 3) {
          acc[0] -= c_ij[1 % D] * q_j[2 % D] - c_ij[2 % D] * q_j[1 % D];
          acc[1] -= c_ij[2 % D] * q_j[0] - c_ij[0] * q_j[2 % D];
          acc[2] -= c_ij[0] * q_j[1 % D] - c_ij[1 % D] * q_j[0];
        }
      }
    }
    const double m_i = od.mi[i];
    if (mode == 1 && D == 2)
      out[i] = acc[0] / m_i;
    else {
      double n2 = acc[0] * acc[0];
      const int n_acc = mode == 0 ? D : 3;
      for (int d = 1; d < n_acc; ++d)
        n2 += acc[d] * acc[d];
      out[i] = std::sqrt(n2) / m_i;
    }
  }
  double q_max = 0., q_min = max_double;
  for (unsigned int i = 0; i < n; ++i) {
    q_max = std::max(q_max, std::abs(out[i]));
    q_min = std::min(q_min, std::abs(out[i]));
  }
  bounds[0] = q_max;
  bounds[1] = q_min;
  const double floor = std::max(1.0e-10, eps_double);
  for (unsigned int i = 0; i < n; ++i) {
    const double q = out[i];
    const double ratio = std::max(0., std::abs(q) - q_min - floor) / std::max(q_max - q_min, eps_double);
    const double magnitude = 1. - std::exp(-beta * ratio);
    out[i] = std::copysign(magnitude, q);
  }
  }
} // namespace

extern "C" {

const char *orc_last_error() { return last_error.c_str(); }

int orc_uses_vcl()
{
#ifdef ORC_USE_VCL
  return 1;
#else
  return 0;
#endif
}

int orc_max_threads()
{
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* FTZ / DAZ on every OpenMP thread, as the reference's main() does on x86-64
 * (source/main.cc:26-36, flush_denormals_to_zero).  Only bench.py's CPU arm calls this: the
 * parity tests keep IEEE gradual underflow, which is what the GPU's FP64 units implement. */
int orc_flush_denormals_to_zero(int on)
{
#if defined(__x86_64__)
  int n = 0;
#pragma omp parallel reduction(+ : n)
  {
    unsigned int mxcsr = __builtin_ia32_stmxcsr();
    if (on)
      mxcsr |= (1u << 6) | (1u << 15);
    else
      mxcsr &= ~((1u << 6) | (1u << 15));
    __builtin_ia32_ldmxcsr(mxcsr);
    n += 1;
  }
  return n;
#else
  (void)on;
  return 0;
#endif
}

void orc_set_threads(int n)
{
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

double orc_pow(double x, double b) { return orc::pow(x, b); }

/* ---- unit level: Euler ------------------------------------------------ */

/* rd = (rho, u, p); out = (p*_two_rarefaction, p*_failsafe, p_2, phi(p_2),
 * lambda_max, a_i, a_j); tests/euler/riemann_solver.cc */
int orc_euler_riemann(const double *params,
                      const double *state_i,
                      const double *state_j,
                      double *out)
{
  return guarded([&] {
    euler::System<1> sys(make_params(params));
    euler::RiemannSolver<1> rs(sys);
    auto rd = [&](const double *s) {
      std::array<double, 4> r{{s[0], s[1], s[2], std::sqrt(sys.gamma * s[2] / s[0])}};
      return r;
    };
    const auto rd_i = rd(state_i), rd_j = rd(state_j);
    euler::RiemannTrace tr;
    rs.compute(rd_i, rd_j, &tr);
    out[0] = tr.p_star_two_rarefaction;
    out[1] = tr.p_star_failsafe;
    out[2] = tr.p_2;
    out[3] = tr.phi_p_2;
    out[4] = tr.lambda_max;
    out[5] = rd_i[3];
    out[6] = rd_j[3];
  });
}

/* full state version: lambda_max(U_i, U_j, n_ij) in dim dimensions */
int orc_euler_riemann_states(int dim,
                             const double *params,
                             const double *U_i,
                             const double *U_j,
                             const double *n_ij,
                             double *lambda_max)
{
  return guarded([&] {
    auto run = [&](auto dimc) {
      constexpr int d = decltype(dimc)::value;
      euler::System<d> sys(make_params(params));
      euler::RiemannSolver<d> rs(sys);
      typename euler::System<d>::state_type a, b;
      Vec<d> n;
      for (int c = 0; c < d + 2; ++c) {
        a[c] = U_i[c];
        b[c] = U_j[c];
      }
      for (int c = 0; c < d; ++c)
        n[c] = n_ij[c];
      *lambda_max = rs.compute(a, b, n);
    };
    if (dim == 1)
      run(std::integral_constant<int, 1>());
    else if (dim == 2)
      run(std::integral_constant<int, 2>());
    else
      run(std::integral_constant<int, 3>());
  });
}

/* dim = 1 limiter; out: l, success, t_l_start, t_r_start, n_rounds,
 * then per round (t_l, t_r, psi_l, psi_r); tests/euler/limiter.cc */
int orc_euler_limiter(int dim,
                      const double *params,
                      const double *U,
                      const double *P,
                      const double *bounds,
                      double *out)
{
  return guarded([&] {
    auto run = [&](auto dimc) {
      constexpr int d = decltype(dimc)::value;
      euler::System<d> sys(make_params(params));
      euler::Limiter<d> lim(sys);
      typename euler::System<d>::state_type u, p;
      for (int c = 0; c < d + 2; ++c) {
        u[c] = U[c];
        p[c] = P[c];
      }
      typename euler::System<d>::bounds_type b{{bounds[0], bounds[1], bounds[2]}};
      bool success;
      euler::LimiterTrace tr;
      const double l = lim.limit(b, u, p, success, 0., 1., &tr);
      out[0] = l;
      out[1] = success ? 1. : 0.;
      out[2] = tr.t_l_start;
      out[3] = tr.t_r_start;
      out[4] = tr.n_rounds;
      for (int r = 0; r < tr.n_rounds; ++r) {
        out[5 + 4 * r + 0] = tr.t_l[r];
        out[5 + 4 * r + 1] = tr.t_r[r];
        out[5 + 4 * r + 2] = tr.psi_l[r];
        out[5 + 4 * r + 3] = tr.psi_r[r];
      }
      out[5 + 4 * 8] = sys.specific_entropy(u);
    };
    if (dim == 1)
      run(std::integral_constant<int, 1>());
    else if (dim == 2)
      run(std::integral_constant<int, 2>());
    else
      run(std::integral_constant<int, 3>());
  });
}

/* tests/euler/hyperbolic_system.cc: out is filled in print order:
 * momentum[dim], internal_energy, internal_energy_derivative[k], pressure,
 * specific_entropy, harten_entropy, harten_entropy_derivative[k],
 * mathematical_entropy, mathematical_entropy_derivative[k], f[k*dim] */
int orc_euler_system_quantities(int dim, const double *params, const double *U, double *out)
{
  return guarded([&] {
    auto run = [&](auto dimc) {
      constexpr int d = decltype(dimc)::value;
      constexpr int k = d + 2;
      euler::System<d> sys(make_params(params));
      typename euler::System<d>::state_type u;
      for (int c = 0; c < k; ++c)
        u[c] = U[c];
      int o = 0;
      const auto m = sys.momentum(u);
      for (int c = 0; c < d; ++c)
        out[o++] = m[c];
      out[o++] = sys.internal_energy(u);
      const auto ied = sys.internal_energy_derivative(u);
      for (int c = 0; c < k; ++c)
        out[o++] = ied[c];
      out[o++] = sys.pressure(u);
      out[o++] = sys.specific_entropy(u);
      out[o++] = sys.harten_entropy(u);
      const auto hed = sys.harten_entropy_derivative(u);
      for (int c = 0; c < k; ++c)
        out[o++] = hed[c];
      out[o++] = sys.mathematical_entropy(u);
      const auto med = sys.mathematical_entropy_derivative(u);
      for (int c = 0; c < k; ++c)
        out[o++] = med[c];
      const auto f = sys.f(u);
      for (int c = 0; c < k; ++c)
        for (int e = 0; e < d; ++e)
          out[o++] = f[c][e];
    };
    if (dim == 1)
      run(std::integral_constant<int, 1>());
    else if (dim == 2)
      run(std::integral_constant<int, 2>());
    else
      run(std::integral_constant<int, 3>());
  });
}

/* ---- unit level: SIMD sparsity layout ---------------------------------- */

/* outputs sized nnz; row_starts sized n_rows+1 (as size_t -> uint64) */
int orc_sparsity_simd(unsigned int n_rows,
                      const unsigned int *row_ptr,
                      const unsigned int *cols,
                      unsigned int n_internal,
                      unsigned int simd_length,
                      unsigned long long *row_starts,
                      unsigned int *column_indices,
                      unsigned int *indices_transposed)
{
  return guarded([&] {
    std::vector<unsigned int> rp(row_ptr, row_ptr + n_rows + 1);
    std::vector<unsigned int> cl(cols, cols + rp.back());
    SparsityPatternSIMD sp;
    sp.reinit(n_internal, simd_length, rp, cl);
    for (unsigned int i = 0; i <= n_rows; ++i)
      row_starts[i] = sp.row_starts[i];
    std::memcpy(column_indices, sp.column_indices.data(), cl.size() * sizeof(unsigned int));
    std::memcpy(indices_transposed, sp.indices_transposed.data(), cl.size() * sizeof(unsigned int));
  });
}

/* ---- offline data ------------------------------------------------------- */

/* Cartesian (optionally masked) Q1 mesh. box_ids[2*dim]: boundary id of the
 * outer box faces (x-,x+,y-,y+,z-,z+); masked_id: id of faces bordering an
 * inactive cell inside the box. */
void *orc_offline_assemble(int dim,
                           const unsigned int *n_cells,
                           const double *origin,
                           const double *h,
                           const unsigned char *mask,
                           const int *box_ids,
                           int masked_id)
{
  void *result = nullptr;
  const int rc = guarded([&] {
    auto run = [&](auto dimc) {
      constexpr int d = decltype(dimc)::value;
      auto *holder = new OfflineHolder<d>();
      auto &mesh = holder->mesh;
      std::size_t ncell = 1;
      for (int c = 0; c < d; ++c) {
        mesh.n_cells[c] = n_cells[c];
        mesh.origin[c] = origin[c];
        mesh.h[c] = h[c];
        ncell *= n_cells[c];
      }
      if (mask)
        mesh.cell_mask.assign(mask, mask + ncell);
      std::array<int, 2 * d> ids;
      for (int f = 0; f < 2 * d; ++f)
        ids[f] = box_ids[f];
      auto nc = mesh.n_cells;
      mesh.boundary_id = [ids, nc, masked_id](const std::array<unsigned int, d> &cell, int f) {
        const int fd = f / 2, side = f % 2;
        const bool outer = side ? (cell[fd] + 1 == nc[fd]) : (cell[fd] == 0);
        return outer ? ids[f] : masked_id;
      };
      holder->od = assemble_offline_data<d>(mesh);
      holder->has_mesh = true;
      result = static_cast<OfflineBase *>(holder);
    };
    if (dim == 1)
      run(std::integral_constant<int, 1>());
    else if (dim == 2)
      run(std::integral_constant<int, 2>());
    else
      run(std::integral_constant<int, 3>());
  });
  return rc == 0 ? result : nullptr;
}

/* Build from caller-provided arrays (e.g. the product's generator output in
 * its own numbering). boundary arrays are structure-of-arrays. positions may
 * be NULL. */
void *orc_offline_from_arrays(int dim,
                              unsigned int n,
                              const unsigned int *row_ptr,
                              const unsigned int *cols,
                              const double *mij,
                              const double *cij,
                              const double *mi,
                              double measure,
                              unsigned int n_boundary,
                              const unsigned int *b_index,
                              const double *b_normal,
                              const int *b_id,
                              const double *b_position,
                              unsigned int n_pairs,
                              const unsigned int *pairs /* (i,col,j) triples */)
{
  void *result = nullptr;
  const int rc = guarded([&] {
    auto run = [&](auto dimc) {
      constexpr int d = decltype(dimc)::value;
      auto *holder = new OfflineHolder<d>();
      auto &od = holder->od;
      od.n_owned = n;
      od.row_ptr.assign(row_ptr, row_ptr + n + 1);
      const std::size_t nnz = od.row_ptr.back();
      od.cols.assign(cols, cols + nnz);
      od.mij.assign(mij, mij + nnz);
      od.cij.assign(cij, cij + nnz * d);
      od.mi.assign(mi, mi + n);
      od.mi_inv.resize(n);
      for (unsigned int i = 0; i < n; ++i)
        od.mi_inv[i] = 1. / mi[i];
      od.measure_of_omega = measure;
      od.build_transposed();
      for (unsigned int b = 0; b < n_boundary; ++b) {
        BoundaryDescription<d> bd;
        bd.i = b_index[b];
        for (int c = 0; c < d; ++c) {
          bd.normal[c] = b_normal[b * d + c];
          bd.position[c] = b_position ? b_position[b * d + c] : 0.;
        }
        bd.normal_mass = 0.;
        bd.boundary_mass = 0.;
        bd.id = b_id[b];
        od.boundary_map.push_back(bd);
      }
      for (unsigned int p = 0; p < n_pairs; ++p)
        od.coupling_boundary_pairs.push_back({pairs[3 * p], pairs[3 * p + 1], pairs[3 * p + 2]});
      result = static_cast<OfflineBase *>(holder);
    };
    if (dim == 1)
      run(std::integral_constant<int, 1>());
    else if (dim == 2)
      run(std::integral_constant<int, 2>());
    else
      run(std::integral_constant<int, 3>());
  });
  return rc == 0 ? result : nullptr;
}

/* General mapped Q1 quadrilateral mesh in 2D (orc_mesh_quads.h): vertices (x, y), cells as 4 vertex
 * indices in deal.II's lexicographic local order, boundary faces (a, b, id, flip) in deal.II's
 * visiting order; outward normal of a -> b = (dy, -dx) / |ab|, reversed if flip. */
void *orc_offline_assemble_quads(unsigned int n_vertices,
                                 const double *xy,
                                 unsigned int n_cells,
                                 const unsigned int *cell_vertices,
                                 unsigned int n_boundary_faces,
                                 const unsigned int *face_vertices,
                                 const unsigned char *face_flip,
                                 const int *face_ids)
{
  void *result = nullptr;
  const int rc = guarded([&] {
    QuadMesh mesh;
    mesh.vertices.resize(n_vertices);
    for (unsigned int v = 0; v < n_vertices; ++v)
      mesh.vertices[v] = {xy[2 * v], xy[2 * v + 1]};
    mesh.cells.resize(n_cells);
    for (unsigned int c = 0; c < n_cells; ++c)
      for (int v = 0; v < 4; ++v) {
        if (cell_vertices[4 * c + v] >= n_vertices)
          throw std::runtime_error("orc_offline_assemble_quads: vertex index out of range");
        mesh.cells[c][v] = cell_vertices[4 * c + v];
      }
    mesh.boundary_faces.resize(n_boundary_faces);
    for (unsigned int f = 0; f < n_boundary_faces; ++f)
      mesh.boundary_faces[f] = {face_vertices[2 * f], face_vertices[2 * f + 1], face_ids[f], face_flip && face_flip[f] != 0};
    auto *holder = new OfflineHolder<2>();
    holder->od = assemble_offline_data_quads(mesh);
    holder->has_mesh = false;
    result = static_cast<OfflineBase *>(holder);
  });
  return rc == 0 ? result : nullptr;
}

void orc_offline_destroy(void *h) { delete static_cast<OfflineBase *>(h); }

#define WITH_OFFLINE(h, body)                                                  \
  do {                                                                         \
    OfflineBase *base__ = static_cast<OfflineBase *>(h);                       \
    if (base__->dimension == 1) {                                                    \
      auto &H = *static_cast<OfflineHolder<1> *>(base__);                      \
      constexpr int D = 1;                                                     \
      (void)D;                                                                 \
      body                                                                     \
    } else if (base__->dimension == 2) {                                             \
      auto &H = *static_cast<OfflineHolder<2> *>(base__);                      \
      constexpr int D = 2;                                                     \
      (void)D;                                                                 \
      body                                                                     \
    } else {                                                                   \
      auto &H = *static_cast<OfflineHolder<3> *>(base__);                      \
      constexpr int D = 3;                                                     \
      (void)D;                                                                 \
      body                                                                     \
    }                                                                          \
  } while (0)


int orc_postprocess(void *h, int k, const double *U, int mode, int first, double beta, double *out, double *bounds)
{
  return guarded([&] {
    WITH_OFFLINE(h, { postprocess_impl<D>(H.od, k, U, mode, first, beta, out, bounds); });
    return 0;
  });
}

/* sizes: out = (n_dofs, nnz, n_boundary, n_pairs) */
/* Postprocessor::compute for one quantity (source/postprocessor.template.h:113-281): mode 0 schlieren
 * of component `first`, mode 1 vorticity of the vector starting at `first`; U is AoS with k components;
 * out: n values on the exponential scale, bounds: {q_max, q_min}. */
/* owner rank of every row: N-rank emulation, see OfflineData::partition (orc_module.h) */
void orc_offline_set_partition(void *h, const int *partition)
{
  WITH_OFFLINE(h, {
    if (partition)
      H.od.partition.assign(partition, partition + H.od.n_owned);
    else
      H.od.partition.clear();
  });
}

void orc_offline_sizes(void *h, unsigned long long *out)
{
  WITH_OFFLINE(h, {
    out[0] = H.od.n_owned;
    out[1] = H.od.cols.size();
    out[2] = H.od.boundary_map.size();
    out[3] = H.od.coupling_boundary_pairs.size();
  });
}

void orc_offline_get(void *h,
                     unsigned int *row_ptr,
                     unsigned int *cols,
                     unsigned int *transposed,
                     double *mij,
                     double *cij,
                     double *mi,
                     double *measure,
                     unsigned int *b_index,
                     double *b_normal,
                     double *b_masses,
                     int *b_id,
                     double *b_position,
                     unsigned int *pairs,
                     double *positions)
{
  WITH_OFFLINE(h, {
    const auto &od = H.od;
    std::memcpy(row_ptr, od.row_ptr.data(), od.row_ptr.size() * sizeof(unsigned int));
    std::memcpy(cols, od.cols.data(), od.cols.size() * sizeof(unsigned int));
    std::memcpy(transposed, od.transposed.data(), od.transposed.size() * sizeof(unsigned int));
    std::memcpy(mij, od.mij.data(), od.mij.size() * sizeof(double));
    std::memcpy(cij, od.cij.data(), od.cij.size() * sizeof(double));
    std::memcpy(mi, od.mi.data(), od.mi.size() * sizeof(double));
    *measure = od.measure_of_omega;
    for (std::size_t b = 0; b < od.boundary_map.size(); ++b) {
      const auto &bd = od.boundary_map[b];
      b_index[b] = bd.i;
      b_id[b] = bd.id;
      b_masses[2 * b] = bd.normal_mass;
      b_masses[2 * b + 1] = bd.boundary_mass;
      for (int c = 0; c < D; ++c) {
        b_normal[b * D + c] = bd.normal[c];
        b_position[b * D + c] = bd.position[c];
      }
    }
    for (std::size_t p = 0; p < od.coupling_boundary_pairs.size(); ++p) {
      pairs[3 * p] = od.coupling_boundary_pairs[p].i;
      pairs[3 * p + 1] = od.coupling_boundary_pairs[p].col_idx;
      pairs[3 * p + 2] = od.coupling_boundary_pairs[p].j;
    }
    if (positions && H.has_mesh)
      for (unsigned int i = 0; i < od.n_owned; ++i) {
        const auto x = H.mesh.position(i);
        for (int c = 0; c < D; ++c)
          positions[std::size_t(i) * D + c] = x[c];
      }
  });
}

/* ---- simulation handle (module + integrator) ---------------------------- */

/* iv = (kind, direction[3], position[3], prim/left[3], right[3], mach, beta) */
void *orc_sim_create_euler(void *offline, const double *params, const double *iv)
{
  void *result = nullptr;
  const int rc = guarded([&] {
    WITH_OFFLINE(offline, {
      auto *sim = new EulerSim<D>(&H, make_params(params));
      sim->iv.kind = euler::InitialStateKind(int(iv[0]));
      for (int c = 0; c < D; ++c) {
        sim->iv.transform.direction[c] = iv[1 + c];
        sim->iv.transform.position[c] = iv[4 + c];
      }
      for (int c = 0; c < 3; ++c) {
        sim->iv.primitive[c] = iv[7 + c];
        sim->iv.primitive_left[c] = iv[7 + c];
        sim->iv.primitive_right[c] = iv[10 + c];
      }
      sim->iv.mach_number = iv[13];
      sim->iv.beta = iv[14];
      sim->finalize();
      result = static_cast<SimBase *>(sim);
    });
  });
  return rc == 0 ? result : nullptr;
}

/* euler aeos: params = AEOS_PARAM_FIELDS, iv as for orc_sim_create_euler */
void *orc_sim_create_aeos(void *offline, const double *params, const double *iv)
{
  void *result = nullptr;
  const int rc = guarded([&] {
    WITH_OFFLINE(offline, {
      auto *sim = new AeosSim<D>(&H, make_aeos_params(params));
      sim->iv.kind = euler::InitialStateKind(int(iv[0]));
      for (int c = 0; c < D; ++c) {
        sim->iv.transform.direction[c] = iv[1 + c];
        sim->iv.transform.position[c] = iv[4 + c];
      }
      for (int c = 0; c < 3; ++c) {
        sim->iv.primitive[c] = iv[7 + c];
        sim->iv.primitive_left[c] = iv[7 + c];
        sim->iv.primitive_right[c] = iv[10 + c];
      }
      sim->iv.mach_number = iv[13];
      sim->iv.beta = iv[14];
      sim->finalize();
      result = static_cast<SimBase *>(sim);
    });
  });
  return rc == 0 ? result : nullptr;
}

/* tests/euler_aeos/riemann_solver{,-strict}.cc: 1D primitive states (rho, u, p, gamma);
 * out = (lambda_max, p_2, p_star_RS, p_star_SS, p_star_interpolated, p_star_failsafe, a_i, a_j) */
int orc_aeos_riemann(const double *params, const double *state_i, const double *state_j, double *out)
{
  return guarded([&] {
    euler_aeos::System<1> sys(make_aeos_params(params));
    euler_aeos::RiemannSolver<1> rs(sys);
    auto rd = [&](const double *s) {
      const double x = 1. - sys.b * s[0];
      std::array<double, 5> r{{s[0], s[1], s[2], s[3], std::sqrt(s[3] * s[2] / (s[0] * x))}};
      return r;
    };
    const auto ri = rd(state_i), rj = rd(state_j);
    double p_2 = 0.;
    out[0] = rs.compute(ri, rj, &p_2);
    out[1] = p_2;
    out[2] = rs.p_star_RS_full(ri, rj);
    out[3] = rs.p_star_SS_full(ri, rj);
    out[4] = rs.p_star_interpolated(ri, rj);
    out[5] = rs.p_star_failsafe(ri, rj);
    out[6] = ri[4];
    out[7] = rj[4];
  });
}

/* tests/euler_aeos/limiter{,-NASG}.cc (dim = 1): out = (l, success, surrogate specific entropy(U, gamma)) */
int orc_aeos_limiter(const double *params, const double *U, const double *P, const double *bounds, double *out)
{
  return guarded([&] {
    euler_aeos::System<1> sys(make_aeos_params(params));
    euler_aeos::Limiter<1> lim(sys);
    euler_aeos::System<1>::state_type u{{U[0], U[1], U[2]}}, pp{{P[0], P[1], P[2]}};
    euler_aeos::System<1>::bounds_type b{{bounds[0], bounds[1], bounds[2], bounds[3]}};
    bool success = true;
    out[0] = lim.limit(b, u, pp, success);
    out[1] = success ? 1. : 0.;
    out[2] = sys.surrogate_specific_entropy(u, bounds[3]);
  });
}

/* tests/euler_aeos/hyperbolic_system.cc: U (dim + 2), gamma -> internal energy, its derivative (k),
 * surrogate specific entropy, harten entropy, its derivative (k), surrogate pressure, surrogate
 * gamma, f (k * dim) */
int orc_aeos_system_quantities(int dim, const double *params, const double *U, double gamma, double *out)
{
  return guarded([&] {
    auto run = [&](auto dimc) {
      constexpr int d = decltype(dimc)::value;
      constexpr int k = d + 2;
      euler_aeos::System<d> sys(make_aeos_params(params));
      typename euler_aeos::System<d>::state_type u;
      for (int c = 0; c < k; ++c)
        u[c] = U[c];
      int o = 0;
      out[o++] = sys.internal_energy(u);
      const auto ied = sys.internal_energy_derivative(u);
      for (int c = 0; c < k; ++c)
        out[o++] = ied[c];
      out[o++] = sys.surrogate_specific_entropy(u, gamma);
      const double eta = sys.surrogate_harten_entropy(u, gamma);
      out[o++] = eta;
      const auto hd = sys.surrogate_harten_entropy_derivative(u, eta, gamma);
      for (int c = 0; c < k; ++c)
        out[o++] = hd[c];
      const double p = sys.surrogate_pressure(u, gamma);
      out[o++] = p;
      out[o++] = sys.surrogate_gamma(u, p);
      const auto f = sys.f(u, p);
      for (int c = 0; c < k; ++c)
        for (int e = 0; e < d; ++e)
          out[o++] = f[c][e];
    };
    if (dim == 1)
      run(std::integral_constant<int, 1>());
    else if (dim == 2)
      run(std::integral_constant<int, 2>());
    else
      run(std::integral_constant<int, 3>());
  });
}

/* iv = (kind, direction[3], position[3], p[4]) */
void *orc_sim_create_sw(void *offline, const double *params, const double *iv)
{
  void *result = nullptr;
  const int rc = guarded([&] {
    WITH_OFFLINE(offline, {
      auto *sim = new SwSim<D>(&H, make_sw_params(params));
      sim->iv.kind = sw::InitialStateKind(int(iv[0]));
      for (int c = 0; c < D; ++c) {
        sim->iv.transform.direction[c] = iv[1 + c];
        sim->iv.transform.position[c] = iv[4 + c];
      }
      for (int c = 0; c < 4; ++c)
        sim->iv.p[c] = iv[7 + c];
      sim->finalize();
      result = static_cast<SimBase *>(sim);
    });
  });
  return rc == 0 ? result : nullptr;
}

int orc_sim_n_bounds(void *s) { return static_cast<SimBase *>(s)->n_bounds(); }
int orc_sim_n_precomputed(void *s) { return static_cast<SimBase *>(s)->n_precomputed(); }

int orc_sim_set_initial_precomputed(void *s, const double *values)
{
  return guarded([&] { static_cast<SimBase *>(s)->set_initial_precomputed(values); });
}

void orc_sim_set_error_components(void *s, unsigned int mask) { static_cast<SimBase *>(s)->error_mask = mask; }

/* tests/shallow_water/riemann_solver.cc: conserved 1D states (h, m); out = (h_i, u_i, h_j, u_j,
 * h_star, lambda_max) */
int orc_sw_riemann(const double *params, const double *state_i, const double *state_j, double *out)
{
  return guarded([&] {
    sw::System<1> sys(make_sw_params(params));
    sw::RiemannSolver<1> rs(sys);
    auto rd = [&](const double *st) {
      const double h = sys.water_depth_sharp({{st[0], st[1]}});
      return std::array<double, 3>{{h, st[1] / h, std::sqrt(sys.prm.gravity * h)}};
    };
    const auto rd_i = rd(state_i), rd_j = rd(state_j);
    out[0] = rd_i[0];
    out[1] = rd_i[1];
    out[2] = rd_j[0];
    out[3] = rd_j[1];
    out[4] = rs.compute_h_star(rd_i, rd_j);
    out[5] = rs.compute(rd_i, rd_j);
  });
}

void orc_sim_destroy(void *s) { delete static_cast<SimBase *>(s); }

int orc_sim_k(void *s) { return static_cast<SimBase *>(s)->k(); }

int orc_sim_interpolate(void *s, double t, int slot)
{
  return guarded([&] { static_cast<SimBase *>(s)->interpolate(t, slot); });
}
int orc_sim_set_U(void *s, int slot, const double *U)
{
  return guarded([&] { static_cast<SimBase *>(s)->set_U(slot, U); });
}
int orc_sim_get_U(void *s, int slot, double *U)
{
  return guarded([&] { static_cast<SimBase *>(s)->get_U(slot, U); });
}
int orc_sim_get_prec(void *s, int slot, double *p)
{
  return guarded([&] { static_cast<SimBase *>(s)->get_prec(slot, p); });
}
int orc_sim_prepare_state(void *s, int slot, double t)
{
  return guarded([&] { static_cast<SimBase *>(s)->prepare_state(slot, t); });
}
int orc_sim_hyp_step(void *s,
                     int old_slot,
                     int stages,
                     const int *stage_slots,
                     const double *weights,
                     int new_slot,
                     double tau,
                     double tau_max,
                     double *tau_out)
{
  return guarded([&] {
    *tau_out = static_cast<SimBase *>(s)->hyp_step(
        old_slot, stages, stage_slots, weights, new_slot, tau, tau_max);
  });
}
int orc_sim_ti_step(void *s, double t, double t_final, double *tau_out)
{
  return guarded([&] { *tau_out = static_cast<SimBase *>(s)->ti_step(t, t_final); });
}
int orc_sim_set_cfl(void *s, double cfl)
{
  return guarded([&] { static_cast<SimBase *>(s)->set_cfl(cfl); });
}
int orc_sim_set_strategy(void *s, int raise)
{
  return guarded([&] { static_cast<SimBase *>(s)->set_strategy(raise); });
}
int orc_sim_configure_ti(void *s, int scheme, double cfl_min, double cfl_max, int recovery)
{
  return guarded([&] {
    static_cast<SimBase *>(s)->configure_ti(scheme, cfl_min, cfl_max, recovery);
  });
}
int orc_sim_get_scratch(void *s, int what, double *out)
{
  return guarded([&] { static_cast<SimBase *>(s)->get_scratch(what, out); });
}
int orc_sim_counters(void *s, unsigned int *restarts, unsigned int *warnings)
{
  return guarded([&] { static_cast<SimBase *>(s)->counters(restarts, warnings); });
}
double orc_sim_last_tau_max(void *s) { return static_cast<SimBase *>(s)->last_tau_max(); }
int orc_sim_errors(void *s, double t, double *out3)
{
  return guarded([&] { static_cast<SimBase *>(s)->errors(t, out3); });
}

} /* extern "C" */
