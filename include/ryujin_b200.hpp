// C++ shim on top of the C ABI (include/ryujin_b200.h): re-creates the reference's class
// surface for the explicit hyperbolic update so that host code written against
//   ryujin::HyperbolicModule<Description,dim,Number>  (source/hyperbolic_module.h:110-278)
//   ryujin::TimeIntegrator<Description,dim,Number>    (source/time_integrator.h)
// can call the B200 path with the same call sequence.  Header only; link with
// libryujin_b200.so.  See INTEGRATION.md.
#pragma once

#include "ryujin_b200.h"

#include <array>
#include <functional>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace ryujin_b200
{
  /** ryujin::Restart, source/hyperbolic_module.h:59-61 */
  class Restart final
  {
  };

  /** ryujin::IDViolationStrategy, source/hyperbolic_module.h:37-52 */
  enum class IDViolationStrategy : int { warn = RJ_STRATEGY_WARN, raise_exception = RJ_STRATEGY_RAISE_EXCEPTION };

  inline void check(int rc)
  {
    if (rc == RJ_OK)
      return;
    if (rc == RJ_RESTART)
      throw Restart();
    throw std::runtime_error(std::string("ryujin_b200: ") + rj_last_error());
  }

  /** Device resident OfflineData (source/offline_data.h). */
  class OfflineData
  {
  public:
    OfflineData(const rj_offline_desc &desc, int device = 0) { check(rj_offline_create(&handle_, &desc, device)); }
    ~OfflineData() { rj_offline_destroy(handle_); }
    OfflineData(const OfflineData &) = delete;
    OfflineData &operator=(const OfflineData &) = delete;
    rj_offline *handle() const { return handle_; }

  private:
    rj_offline *handle_ = nullptr;
  };

  class HyperbolicModule;

  /** StateVector = (U, precomputed) on the device (source/state_vector.h:47-51). */
  class StateVector
  {
  public:
    explicit StateVector(HyperbolicModule &module);
    ~StateVector() { rj_state_destroy(handle_); }
    StateVector(const StateVector &) = delete;
    StateVector &operator=(const StateVector &) = delete;

    /** host AoS (U_i)_c at i*k + c, as MultiComponentVector stores it */
    void upload(const double *host_U) { check(rj_state_upload(handle_, host_U)); }
    void download(double *host_U) const { check(rj_state_download(handle_, host_U)); }
    void interpolate(double t) { check(rj_state_interpolate(handle_, t)); }
    void swap(StateVector &other) { check(rj_state_swap(handle_, other.handle_)); }
    /** dst = s * dst + b * src, time_integrator.template.h:18-25 */
    void sadd(double s, double b, const StateVector &src) { check(rj_state_sadd(handle_, s, b, src.handle_)); }
    rj_state *handle() const { return handle_; }

  private:
    rj_state *handle_ = nullptr;
  };

  /** HyperbolicModule, source/hyperbolic_module.h:110-278 */
  class HyperbolicModule
  {
  public:
    HyperbolicModule(const OfflineData &offline_data,
                     const rj_equation_desc &hyperbolic_system,
                     const rj_initial_values_desc &initial_values,
                     const rj_module_params &parameters,
                     rj_comm *mpi_communicator = nullptr,
                     void *stream = nullptr)
    {
      check(rj_hyperbolic_module_create(
          &handle_, offline_data.handle(), &hyperbolic_system, &parameters, &initial_values, mpi_communicator, stream));
    }
    ~HyperbolicModule() { rj_hyperbolic_module_destroy(handle_); }
    HyperbolicModule(const HyperbolicModule &) = delete;
    HyperbolicModule &operator=(const HyperbolicModule &) = delete;

    /** prepare() is part of construction (scratch allocation), hyperbolic_module.template.h:53-86 */
    void prepare() const {}

    /** initial_precomputed() of the reference (hyperbolic_module.h:259,309): the bathymetry of the
     * shallow water equations, one value per owned row; ghost rows are exchanged inside */
    void set_initial_precomputed(const double *owned_values) const
    {
      check(rj_hyperbolic_module_set_initial_precomputed(handle_, owned_values));
    }

    /** hyperbolic_module.template.h:97-193 */
    void prepare_state_vector(StateVector &state_vector, double t) const
    {
      check(rj_hyperbolic_module_prepare_state_vector(handle_, state_vector.handle(), t));
    }

    /** template <int stages> Number step(old, stage_state_vectors, stage_weights, new, tau, tau_max),
     * hyperbolic_module.h:215-223; throws Restart like the reference */
    template <int stages>
    double step(const StateVector &old_state_vector,
                std::array<std::reference_wrapper<const StateVector>, stages> stage_state_vectors,
                const std::array<double, stages> stage_weights,
                StateVector &new_state_vector,
                double tau = 0.,
                double tau_max = std::numeric_limits<double>::max()) const
    {
      const rj_state *st[stages > 0 ? stages : 1] = {};
      double w[stages > 0 ? stages : 1] = {};
      for (int s = 0; s < stages; ++s) {
        st[s] = stage_state_vectors[s].get().handle();
        w[s] = stage_weights[s];
      }
      double tau_out = 0.;
      check(rj_hyperbolic_module_step(
          handle_, old_state_vector.handle(), stages, st, w, new_state_vector.handle(), tau, tau_max, &tau_out));
      return tau_out;
    }

    void cfl(double new_cfl) const { check(rj_hyperbolic_module_set_cfl(handle_, new_cfl)); }
    double cfl() const { return rj_hyperbolic_module_cfl(handle_); }

    unsigned int n_restarts() const
    {
      uint32_t r = 0, w = 0;
      check(rj_hyperbolic_module_counters(handle_, &r, &w));
      return r;
    }
    unsigned int n_warnings() const
    {
      uint32_t r = 0, w = 0;
      check(rj_hyperbolic_module_counters(handle_, &r, &w));
      return w;
    }

    /** the reference exposes a public mutable member id_violation_strategy_ */
    void id_violation_strategy(IDViolationStrategy s) const
    {
      check(rj_hyperbolic_module_set_id_violation_strategy(handle_, static_cast<int>(s)));
    }

    rj_hyperbolic_module *handle() const { return handle_; }

  private:
    rj_hyperbolic_module *handle_ = nullptr;
  };

  inline StateVector::StateVector(HyperbolicModule &module) { check(rj_state_create(&handle_, module.handle())); }

  /** TimeIntegrator (explicit schemes), source/time_integrator.template.h:207-443 */
  class TimeIntegrator
  {
  public:
    TimeIntegrator(HyperbolicModule &module,
                   rj_time_stepping_scheme scheme = RJ_ERK_33,
                   double cfl_min = 0.45,
                   double cfl_max = 0.90,
                   rj_cfl_recovery_strategy recovery = RJ_CFL_RECOVERY_BANG_BANG)
    {
      check(rj_time_integrator_create(&handle_, module.handle(), scheme, cfl_min, cfl_max, recovery));
    }
    ~TimeIntegrator() { rj_time_integrator_destroy(handle_); }
    TimeIntegrator(const TimeIntegrator &) = delete;
    TimeIntegrator &operator=(const TimeIntegrator &) = delete;

    /** Number step(StateVector &state_vector, Number t, Number t_final) */
    double step(StateVector &state_vector, double t, double t_final = std::numeric_limits<double>::max())
    {
      double tau = 0.;
      check(rj_time_integrator_step(handle_, state_vector.handle(), t, t_final, &tau));
      return tau;
    }
    double efficiency() const { return rj_time_integrator_efficiency(handle_); }

    /** how a cycle is driven: reference control flow, device resident, or CUDA graph replay
     * (bit-identical results; see rj_stepping) */
    void stepping(rj_stepping mode) { check(rj_time_integrator_set_stepping(handle_, static_cast<int>(mode))); }
    rj_stepping stepping() const { return static_cast<rj_stepping>(rj_time_integrator_stepping(handle_)); }

  private:
    rj_time_integrator *handle_ = nullptr;
  };
} // namespace ryujin_b200
