// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_common.h).
//
//   TimeIntegrator (ssprk 22/33, erk 11/22/33/43), CFL bang-bang restart
//       <- source/time_integrator.template.h:18-25,207-443
//   InitialValues affine / Galilei transform
//       <- source/initial_values.template.h:65-219
//   Euler initial states uniform / contrast / isentropic vortex
//       <- source/euler/initial_state_{uniform,contrast,isentropic_vortex}.h
//   error norms (Q1, QGauss(3) cell integration)
//       <- source/time_loop.template.h:690-832
#pragma once

#include "orc_mesh.h"
#include "orc_shallow_water.h"
#include "orc_module.h"

namespace orc
{
  enum class TimeSteppingScheme { ssprk_22, ssprk_33, erk_11, erk_22, erk_33, erk_43, erk_54 };
  enum class CFLRecoveryStrategy { none, bang_bang_control };

  template <typename Desc, int dim>
  class TimeIntegrator
  {
  public:
    using Module = HyperbolicModule<Desc, dim>;
    using StateVector = typename Module::StateVector;

    double cfl_min_ = 0.45, cfl_max_ = 0.90;
    CFLRecoveryStrategy cfl_recovery_strategy_ =
        CFLRecoveryStrategy::bang_bang_control;
    TimeSteppingScheme time_stepping_scheme_ = TimeSteppingScheme::erk_33;

    explicit TimeIntegrator(const Module &module)
        : module_(module)
    {
    }

    void prepare()
    {
      unsigned int n_temp = 0;
      switch (time_stepping_scheme_) {
      case TimeSteppingScheme::ssprk_22: n_temp = 2; efficiency_ = 1.; break;
      case TimeSteppingScheme::ssprk_33: n_temp = 2; efficiency_ = 1.; break;
      case TimeSteppingScheme::erk_11: n_temp = 1; efficiency_ = 1.; break;
      case TimeSteppingScheme::erk_22: n_temp = 2; efficiency_ = 2.; break;
      case TimeSteppingScheme::erk_33: n_temp = 3; efficiency_ = 3.; break;
      case TimeSteppingScheme::erk_43: n_temp = 4; efficiency_ = 4.; break;
      case TimeSteppingScheme::erk_54: n_temp = 5; efficiency_ = 5.; break;
      }
      temp_.resize(n_temp);
      for (auto &it : temp_)
        it.reinit(module_.offline_data().n_owned);
      if (!(cfl_min_ > 0.) || !(cfl_max_ >= cfl_min_))
        throw std::runtime_error("invalid cfl min / cfl max");
      module_.cfl(cfl_max_);
    }

    double efficiency() const { return efficiency_; }

    double step(StateVector &sv, double t, double t_final = max_double)
    {
      double tau_max = t_final - t;

      const auto single_step = [&]() {
        switch (time_stepping_scheme_) {
        case TimeSteppingScheme::ssprk_22: return step_ssprk_22(sv, t, tau_max);
        case TimeSteppingScheme::ssprk_33: return step_ssprk_33(sv, t, tau_max);
        case TimeSteppingScheme::erk_11: return step_erk_11(sv, t, tau_max);
        case TimeSteppingScheme::erk_22: return step_erk_22(sv, t, tau_max);
        case TimeSteppingScheme::erk_33: return step_erk_33(sv, t, tau_max);
        case TimeSteppingScheme::erk_43: return step_erk_43(sv, t, tau_max);
        case TimeSteppingScheme::erk_54: return step_erk_54(sv, t, tau_max);
        }
        return 0.;
      };

      if (cfl_recovery_strategy_ == CFLRecoveryStrategy::bang_bang_control) {
        module_.id_violation_strategy_ = IDViolationStrategy::raise_exception;
        module_.cfl(cfl_max_);
      }

      try {
        return single_step();
      } catch (Restart) {
        if (cfl_recovery_strategy_ == CFLRecoveryStrategy::none)
          throw std::logic_error("Restart with cfl recovery strategy none");
        module_.id_violation_strategy_ = IDViolationStrategy::warn;
        module_.cfl(cfl_min_);
        return single_step();
      }
    }

  private:
    static void sadd(StateVector &dst, double s, double b, const StateVector &src)
    {
      /* dst = s * dst + b * src on U only */
      for (std::size_t i = 0; i < dst.U.size(); ++i)
        dst.U[i] = s * dst.U[i] + b * src.U[i];
    }

    double step_ssprk_22(StateVector &sv, double t, double tau_max)
    {
      module_.prepare_state_vector(sv, t);
      const double tau = module_.step(sv, {}, {}, temp_[0], 0., tau_max);
      module_.prepare_state_vector(temp_[0], t + 1.0 * tau);
      module_.step(temp_[0], {}, {}, temp_[1], tau);
      sadd(temp_[1], 1.0 / 2.0, 1.0 / 2.0, sv);
      sv.swap(temp_[1]);
      return tau;
    }

    double step_ssprk_33(StateVector &sv, double t, double tau_max)
    {
      module_.prepare_state_vector(sv, t);
      const double tau = module_.step(sv, {}, {}, temp_[0], 0., tau_max);
      module_.prepare_state_vector(temp_[0], t + 1.0 * tau);
      module_.step(temp_[0], {}, {}, temp_[1], tau);
      sadd(temp_[1], 1.0 / 4.0, 3.0 / 4.0, sv);
      module_.prepare_state_vector(temp_[1], t + 0.5 * tau);
      module_.step(temp_[1], {}, {}, temp_[0], tau);
      sadd(temp_[0], 2.0 / 3.0, 1.0 / 3.0, sv);
      sv.swap(temp_[0]);
      return tau;
    }

    double step_erk_11(StateVector &sv, double t, double tau_max)
    {
      module_.prepare_state_vector(sv, t);
      const double tau = module_.step(sv, {}, {}, temp_[0], 0., tau_max);
      sv.swap(temp_[0]);
      return tau;
    }

    double step_erk_22(StateVector &sv, double t, double tau_max)
    {
      module_.prepare_state_vector(sv, t);
      const double tau = module_.step(sv, {}, {}, temp_[0], 0., tau_max / 2.);
      module_.prepare_state_vector(temp_[0], t + 1.0 * tau);
      module_.step(temp_[0], {&sv}, {-1.}, temp_[1], tau);
      sv.swap(temp_[1]);
      return 2. * tau;
    }

    double step_erk_33(StateVector &sv, double t, double tau_max)
    {
      module_.prepare_state_vector(sv, t);
      const double tau = module_.step(sv, {}, {}, temp_[0], 0., tau_max / 3.);
      module_.prepare_state_vector(temp_[0], t + 1.0 * tau);
      module_.step(temp_[0], {&sv}, {-1.}, temp_[1], tau);
      module_.prepare_state_vector(temp_[1], t + 2.0 * tau);
      module_.step(temp_[1], {&sv, &temp_[0]}, {0.75, -2.}, temp_[2], tau);
      sv.swap(temp_[2]);
      return 3. * tau;
    }

    double step_erk_43(StateVector &sv, double t, double tau_max)
    {
      module_.prepare_state_vector(sv, t);
      const double tau = module_.step(sv, {}, {}, temp_[0], 0., tau_max / 4.);
      module_.prepare_state_vector(temp_[0], t + 1.0 * tau);
      module_.step(temp_[0], {&sv}, {-1.}, temp_[1], tau);
      module_.prepare_state_vector(temp_[1], t + 2.0 * tau);
      module_.step(temp_[1], {&temp_[0]}, {-1.}, temp_[2], tau);
      module_.prepare_state_vector(temp_[2], t + 3.0 * tau);
      module_.step(temp_[2],
                   {&temp_[0], &temp_[1]},
                   {5. / 3., -10. / 3.},
                   temp_[3],
                   tau);
      sv.swap(temp_[3]);
      return 4. * tau;
    }

    /* time_integrator.template.h:446-510 */
    double step_erk_54(StateVector &sv, double t, double tau_max)
    {
      constexpr double c = 0.2;
      constexpr double a_21 = +0.2;
      constexpr double a_31 = +0.26075582269554909;
      constexpr double a_32 = +0.13924417730445096;
      constexpr double a_41 = -0.25856517872570289;
      constexpr double a_42 = +0.91136274166280729;
      constexpr double a_43 = -0.05279756293710430;
      constexpr double a_51 = +0.21623276431503774;
      constexpr double a_52 = +0.51534223099602405;
      constexpr double a_53 = -0.81662794199265554;
      constexpr double a_54 = +0.88505294668159373;
      constexpr double a_61 = -0.10511678454691901;
      constexpr double a_62 = +0.87880047152100838;
      constexpr double a_63 = -0.58903404061484477;
      constexpr double a_64 = +0.46213380485434047;

      module_.prepare_state_vector(sv, t);
      const double tau = module_.step(sv, {}, {}, temp_[0], 0., tau_max / 5.);
      module_.prepare_state_vector(temp_[0], t + 1.0 * tau);
      module_.step(temp_[0], {&sv}, {(a_31 - a_21) / c}, temp_[1], tau);
      module_.prepare_state_vector(temp_[1], t + 2.0 * tau);
      module_.step(temp_[1], {&sv, &temp_[0]}, {(a_41 - a_31) / c, (a_42 - a_32) / c}, temp_[2], tau);
      module_.prepare_state_vector(temp_[2], t + 3.0 * tau);
      module_.step(temp_[2],
                   {&sv, &temp_[0], &temp_[1]},
                   {(a_51 - a_41) / c, (a_52 - a_42) / c, (a_53 - a_43) / c},
                   temp_[3],
                   tau);
      module_.prepare_state_vector(temp_[3], t + 4.0 * tau);
      module_.step(temp_[3],
                   {&sv, &temp_[0], &temp_[1], &temp_[2]},
                   {(a_61 - a_51) / c, (a_62 - a_52) / c, (a_63 - a_53) / c, (a_64 - a_54) / c},
                   temp_[4],
                   tau);
      sv.swap(temp_[4]);
      return 5. * tau;
    }

    const Module &module_;
    std::vector<StateVector> temp_;
    double efficiency_ = 1.;
  };


  /* InitialValues: direction / position transform around an initial state. */
  template <int dim>
  struct InitialTransform {
    Vec<dim> direction{}; /* normalized in ctor */
    Vec<dim> position{};

    InitialTransform()
    {
      direction[0] = 1.;
      position[0] = 1.; /* initial_values.template.h:45-49 */
    }

    void normalize()
    {
      const double n = norm<dim>(direction);
      if (n == 0.)
        throw std::runtime_error("Initial direction is set to the zero vector.");
      for (int d = 0; d < dim; ++d)
        direction[d] /= n;
    }

    Vec<dim> point(const Vec<dim> &x) const
    {
      Vec<dim> dir;
      for (int d = 0; d < dim; ++d)
        dir[d] = x[d] - position[d];

      if constexpr (dim == 3) {
        double n_x = direction[0], n_z = direction[2];
        const double nrm = std::sqrt(n_x * n_x + n_z * n_z);
        n_x /= nrm;
        n_z /= nrm;
        auto nd = dir;
        if (nrm > 1.0e-14) {
          nd[0] = n_x * dir[0] + n_z * dir[2];
          nd[2] = -n_z * dir[0] + n_x * dir[2];
        }
        dir = nd;
      }
      if constexpr (dim >= 2) {
        double n_x = direction[0], n_y = direction[1];
        const double nrm = std::sqrt(n_x * n_x + n_y * n_y);
        n_x /= nrm;
        n_y /= nrm;
        auto nd = dir;
        if (nrm > 1.0e-14) {
          nd[0] = n_x * dir[0] + n_y * dir[1];
          nd[1] = -n_y * dir[0] + n_x * dir[1];
        }
        dir = nd;
      }
      return dir;
    }

    Vec<dim> vector(Vec<dim> v) const
    {
      if constexpr (dim >= 2) {
        double n_x = direction[0], n_y = direction[1];
        const double nrm = std::sqrt(n_x * n_x + n_y * n_y);
        n_x /= nrm;
        n_y /= nrm;
        auto nv = v;
        if (nrm > 1.0e-14) {
          nv[0] = n_x * v[0] - n_y * v[1];
          nv[1] = n_y * v[0] + n_x * v[1];
        }
        v = nv;
      }
      if constexpr (dim == 3) {
        double n_x = direction[0], n_z = direction[2];
        const double nrm = std::sqrt(n_x * n_x + n_z * n_z);
        n_x /= nrm;
        n_z /= nrm;
        auto nv = v;
        if (nrm > 1.0e-14) {
          nv[0] = n_x * v[0] - n_z * v[2];
          nv[2] = n_z * v[0] + n_x * v[2];
        }
        v = nv;
      }
      return v;
    }
  };


  namespace euler
  {
    /* 0..2 match enum rj_initial_state_kind; leblanc / rarefaction exist in the oracle only (1D
     * verification decks of the reference) */
    enum class InitialStateKind { uniform = 0, contrast = 1, isentropic_vortex = 2, leblanc = 3, rarefaction = 4 };

    /* SysT: euler::System<dim>, or euler_aeos::System<dim> (the reference shares the
     * EulerInitialStateLibrary between the two equations; only from_initial_state differs) */
    template <int dim, class SysT = System<dim>>
    struct InitialValues {
      using Sys = SysT;
      using state_type = typename Sys::state_type;

      const Sys &sys;
      InitialStateKind kind = InitialStateKind::uniform;
      InitialTransform<dim> transform;
      /* uniform: primitive (rho,u,p); contrast: left / right */
      std::array<double, 3> primitive{{1.4, 3., 1.}};
      std::array<double, 3> primitive_left{{1.4, 0., 1.}};
      std::array<double, 3> primitive_right{{1.4, 0., 1.}};
      /* vortex */
      double mach_number = 2.0, beta = 5.0;

      explicit InitialValues(const Sys &s)
          : sys(s)
      {
      }

      state_type compute(const Vec<dim> &point, double t) const
      {
        switch (kind) {
        case InitialStateKind::uniform:
          return sys.from_initial_state(primitive);
        case InitialStateKind::contrast:
          return point[0] > 0. ? sys.from_initial_state(primitive_right)
                               : sys.from_initial_state(primitive_left);
        case InitialStateKind::isentropic_vortex: {
          const double gamma = sys.gamma;
          double pb[2];
          pb[0] = point[0] - mach_number * t;
          pb[1] = dim >= 2 ? point[dim >= 2 ? 1 : 0] : 0.;
          const double r_square = pb[0] * pb[0] + pb[1] * pb[1];
          const double factor =
              beta / (2. * M_PI) * std::exp(0.5 - 0.5 * r_square);
          const double T =
              1. - (gamma - 1.) / (2. * gamma) * factor * factor;
          const double u = mach_number - factor * pb[1];
          const double v = factor * pb[0];
          const double rho = orc::pow(T, 1. / (gamma - 1.));
          const double p = orc::pow(rho, gamma);
          const double E =
              p / (gamma - 1.) + 0.5 * rho * (u * u + v * v);
          state_type s{};
          s[0] = rho;
          s[1] = rho * u;
          if constexpr (dim >= 2)
            s[2] = rho * v;
          s[dim + 1] = E;
          return s;
        }
        case InitialStateKind::leblanc: {
          /* euler/initial_state_leblanc.h:45-111 (gamma = 5/3 hard wired in the energy) */
          const double primitive_left[3] = {1., 0., 2. / 3. * 1.e-1};
          const double primitive_right[3] = {1.e-3, 0., 2. / 3. * 1.e-10};
          constexpr double rarefaction_speed = 0.49578489518897934;
          constexpr double contact_velocity = 0.62183867139173454;
          constexpr double right_shock_speed = 0.82911836253346982;
          constexpr double pre_contact_density = 5.4079335349316249e-02;
          constexpr double post_contact_density = 3.9999980604299963e-03;
          constexpr double contact_pressure = 0.51557792765096996e-03;
          double prim[3];
          const double x = point[0];
          if (x <= -1.0 / 3.0 * t) {
            for (int q = 0; q < 3; ++q)
              prim[q] = primitive_left[q];
          } else if (x < rarefaction_speed * t) {
            const double chi = x / t;
            prim[0] = std::pow(0.75 - 0.75 * chi, 3.0);
            prim[1] = 0.75 * (1.0 / 3.0 + chi);
            prim[2] = (1.0 / 15.0) * std::pow(0.75 - 0.75 * chi, 5.0);
          } else if (x < contact_velocity * t) {
            prim[0] = pre_contact_density;
            prim[1] = contact_velocity;
            prim[2] = contact_pressure;
          } else if (x < right_shock_speed * t) {
            prim[0] = post_contact_density;
            prim[1] = contact_velocity;
            prim[2] = contact_pressure;
          } else {
            for (int q = 0; q < 3; ++q)
              prim[q] = primitive_right[q];
          }
          state_type s{};
          s[0] = prim[0];
          s[1] = prim[0] * prim[1];
          s[dim + 1] = prim[2] / (5. / 3. - 1.) + 0.5 * prim[0] * prim[1] * prim[1];
          return s;
        }
        case InitialStateKind::rarefaction: {
          /* euler/initial_state_rarefaction.h:37-155 */
          const double gamma = sys.gamma;
          auto speed_of_sound = [&](double rho, double p) { return std::sqrt(gamma * p / rho); };
          const double rho_left = 3.0, p_left = 1.0;
          const double c_left = speed_of_sound(rho_left, p_left);
          const double u_left = c_left;
          const double rho_right = 0.5;
          const double p_right = std::pow(rho_right / rho_left, gamma) * p_left;
          const double c_right = speed_of_sound(rho_right, p_right);
          const double u_right = u_left + 2.0 * (c_left - c_right) / (gamma - 1.0);
          const double k1 = 2.0 / (gamma + 1.0);
          const double k2 = ((gamma - 1.0) / ((gamma + 1.0) * c_left));
          const double density_exponent = 2.0 / (gamma - 1.0);
          const double k3 = c_left + ((gamma - 1.0) / 2.0) * u_left;
          const double pressure_exponent = 2.0 * gamma / (gamma - 1.0);
          const double x = point[0];
          const double t_0 = 0.2 / (u_right - u_left);
          const double tt = t_0 + t;
          double prim[3];
          if (x <= tt * (u_left - c_left)) {
            prim[0] = rho_left;
            prim[1] = u_left;
            prim[2] = p_left;
          } else if (x <= tt * (u_right - c_right)) {
            const double chi = x / tt;
            prim[0] = rho_left * std::pow(k1 + k2 * (u_left - chi), density_exponent);
            prim[1] = k1 * (k3 + chi);
            prim[2] = p_left * std::pow(k1 + k2 * (u_left - chi), pressure_exponent);
          } else {
            prim[0] = rho_right;
            prim[1] = u_right;
            prim[2] = p_right;
          }
          state_type s{};
          s[0] = prim[0];
          s[1] = prim[0] * prim[1];
          s[dim + 1] = prim[2] / (gamma - 1.) + 0.5 * prim[0] * prim[1] * prim[1];
          return s;
        }
        }
        return state_type{};
      }

      state_type initial_state(const Vec<dim> &x, double t) const
      {
        const auto tp = transform.point(x);
        auto state = compute(tp, t);
        const auto M = transform.vector(Sys::momentum(state));
        for (int d = 0; d < dim; ++d)
          state[1 + d] = M[d];
        return state;
      }
    };
  } // namespace euler


  namespace sw
  {
    /* kinds match enum rj_initial_state_kind of include/ryujin_b200.h */
    enum class InitialStateKind { circular_dam_break = 16, uniform = 17, ritter_dam_break = 18, smooth_vortex = 19, sloping_friction = 20 };

    template <int dim>
    struct InitialValues {
      using Sys = System<dim>;
      using state_type = typename Sys::state_type;

      const Sys &sys;
      InitialStateKind kind = InitialStateKind::uniform;
      InitialTransform<dim> transform;
      /* circular dam break: (still water depth, radius, dam amplitude)
       * uniform: (h, u); ritter: (time initial, left water depth)
       * smooth vortex: (reference depth, mach number, beta)
       * sloping friction: (ramp slope, initial discharge) */
      std::array<double, 4> p{{0.5, 2.5, 2.5, 0.}};

      explicit InitialValues(const Sys &s)
          : sys(s)
      {
      }

      state_type compute(const Vec<dim> &point, double t) const
      {
        state_type s{};
        switch (kind) {
        case InitialStateKind::circular_dam_break: {
          /* initial_state_circular_dam_break.h:45-51 (sic: compares |x|^2 with the radius) */
          const double r = norm_square<dim>(point);
          s[0] = (r <= p[1] ? p[2] : p[0]);
          return s;
        }
        case InitialStateKind::uniform:
          return sys.from_initial_state({{p[0], p[1]}});
        case InitialStateKind::ritter_dam_break: {
          /* initial_state_ritter_dam_break.h:59-80 */
          const double g = sys.prm.gravity;
          const double t_initial = p[0], left_depth = p[1];
          const double x = point[0];
          const double aL = std::sqrt(g * left_depth);
          const double xA = -(t + t_initial) * aL;
          const double xB = 2. * (t + t_initial) * aL;
          const double tmp = aL - x / (2. * (t + t_initial));
          const double h_expansion = 4. / (9. * g) * tmp * tmp;
          const double v_expansion = 2. / 3. * (x / (t + t_initial) + aL);
          if (x <= xA) {
            s[0] = left_depth;
            s[1] = 0.;
          } else if (x <= xB) {
            s[0] = h_expansion;
            s[1] = h_expansion * v_expansion;
          }
          return s;
        }
        case InitialStateKind::smooth_vortex: {
          /* initial_state_smooth_vortex.h:62-88 (dim == 2 only upstream) */
          const double gravity = sys.prm.gravity;
          const double depth = p[0], mach = p[1], beta = p[2];
          double pb[2];
          pb[0] = point[0] - mach * t;
          pb[1] = dim >= 2 ? point[dim >= 2 ? 1 : 0] : 0.;
          const double r_square = pb[0] * pb[0] + pb[1] * pb[1];
          const double factor = beta / (2. * M_PI) * std::exp(0.5 - 0.5 * r_square);
          const double h = depth - (1. / (2. * gravity)) * factor * factor;
          const double u = mach - factor * pb[1];
          const double v = factor * pb[0];
          s[0] = h;
          s[1] = h * u;
          if constexpr (dim >= 2)
            s[2] = h * v;
          return s;
        }
        case InitialStateKind::sloping_friction: {
          /* initial_state_sloping_friction.h:54-69: steady flow down an inclined plane with
           * Manning friction */
          const double n = sys.prm.manning_friction_coefficient;
          const double slope = p[0], q_0 = p[1];
          const double exponent = 1. / (2. + 4. / 3.);
          const double profile = n * n * q_0 * q_0 / slope;
          s[0] = orc::pow(profile, exponent);
          s[1] = q_0;
          return s;
        }
        }
        return s;
      }

      /* initial_precomputations(point): the bathymetry of the configuration (0 if it has none) */
      double bathymetry(const Vec<dim> &x) const
      {
        const auto tp = transform.point(x);
        if (kind == InitialStateKind::sloping_friction)
          return -p[0] * tp[0]; /* initial_state_sloping_friction.h:83-87 */
        return 0.;
      }

      state_type initial_state(const Vec<dim> &x, double t) const
      {
        const auto tp = transform.point(x);
        auto state = compute(tp, t);
        const auto M = transform.vector(Sys::momentum(state));
        for (int d = 0; d < dim; ++d)
          state[1 + d] = M[d];
        return state;
      }
    };
  } // namespace sw


  /* Linf / L1 / L2 norms of the Q1 function with nodal values v over the
   * Cartesian mesh, QGauss(3) per direction. */
  template <int dim>
  inline std::array<double, 3> q1_norms(const CartesianMesh<dim> &mesh,
                                        const std::vector<double> &v)
  {
    constexpr int nv = 1 << dim;
    std::vector<double> qx, qw;
    gauss_01(3, qx, qw);
    int nq = 1;
    for (int d = 0; d < dim; ++d)
      nq *= 3;
    double cell_measure = 1.;
    for (int d = 0; d < dim; ++d)
      cell_measure *= mesh.h[d];

    double linf = 0.;
    for (double x : v)
      linf = std::max(linf, std::abs(x));

    double l1 = 0., l2sq = 0.;
    mesh.for_each_cell([&](const std::array<unsigned int, dim> &cell) {
      double nodal[nv];
      for (int a = 0; a < nv; ++a)
        nodal[a] = v[mesh.node_to_dof[mesh.node_index(mesh.vertex(cell, a))]];
      double c1 = 0., c2 = 0.;
      for (int q = 0; q < nq; ++q) {
        double JxW = cell_measure, xi[dim];
        int r = q;
        for (int d = 0; d < dim; ++d) {
          xi[d] = qx[r % 3];
          JxW *= qw[r % 3];
          r /= 3;
        }
        double val = 0.;
        for (int a = 0; a < nv; ++a) {
          double phi = 1.;
          for (int d = 0; d < dim; ++d)
            phi *= ((a >> d) & 1) ? xi[d] : 1. - xi[d];
          val += nodal[a] * phi;
        }
        c1 += std::abs(val) * JxW;
        c2 += val * val * JxW;
      }
      l1 += c1;
      l2sq += c2;
    });
    return {{linf, l1, std::sqrt(l2sq)}};
  }
} // namespace orc
