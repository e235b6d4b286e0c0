// Initial / Dirichlet states evaluated on the device.
//   affine + Galilei transform   <- source/initial_values.template.h:65-145,169-183
//   uniform / contrast / vortex  <- source/euler/initial_state_{uniform,contrast,
//                                    isentropic_vortex}.h
// The reference evaluates Dirichlet data in a serial host loop over the boundary
// map (hyperbolic_module.template.h:122-144); here it is one thread per
// boundary DoF inside k_apply_boundary_conditions.
#pragma once

#include "device_common.cuh"

namespace rjb
{
  struct InitialValuesDevice {
    int kind;
    double direction[3]; /* normalized */
    double position[3];
    double primitive[3], primitive_right[3];
    double mach_number, beta;
    double sw[4];

    template <int dim>
    RJ_DEVICE void transform_point(const double (&x)[dim], double (&out)[dim]) const
    {
      double dir[dim];
#pragma unroll
      for (int d = 0; d < dim; ++d)
        dir[d] = x[d] - position[d];
      if constexpr (dim == 3) {
        double n_x = direction[0], n_z = direction[2];
        const double nrm = sqrt(n_x * n_x + n_z * n_z);
        n_x /= nrm;
        n_z /= nrm;
        if (nrm > 1.0e-14) {
          const double a = n_x * dir[0] + n_z * dir[2];
          const double b = -n_z * dir[0] + n_x * dir[2];
          dir[0] = a;
          dir[2] = b;
        }
      }
      if constexpr (dim >= 2) {
        double n_x = direction[0], n_y = direction[1];
        const double nrm = sqrt(n_x * n_x + n_y * n_y);
        n_x /= nrm;
        n_y /= nrm;
        if (nrm > 1.0e-14) {
          const double a = n_x * dir[0] + n_y * dir[1];
          const double b = -n_y * dir[0] + n_x * dir[1];
          dir[0] = a;
          dir[1] = b;
        }
      }
#pragma unroll
      for (int d = 0; d < dim; ++d)
        out[d] = dir[d];
    }

    template <int dim>
    RJ_DEVICE void transform_vector(double (&v)[dim]) const
    {
      if constexpr (dim >= 2) {
        double n_x = direction[0], n_y = direction[1];
        const double nrm = sqrt(n_x * n_x + n_y * n_y);
        n_x /= nrm;
        n_y /= nrm;
        if (nrm > 1.0e-14) {
          const double a = n_x * v[0] - n_y * v[1];
          const double b = n_y * v[0] + n_x * v[1];
          v[0] = a;
          v[1] = b;
        }
      }
      if constexpr (dim == 3) {
        double n_x = direction[0], n_z = direction[2];
        const double nrm = sqrt(n_x * n_x + n_z * n_z);
        n_x /= nrm;
        n_z /= nrm;
        if (nrm > 1.0e-14) {
          const double a = n_x * v[0] - n_z * v[2];
          const double b = n_z * v[0] + n_x * v[2];
          v[0] = a;
          v[2] = b;
        }
      }
    }

    /* InitialValues::initial_state(point, t), initial_values.template.h:169-183 */
    template <class Eq>
    RJ_DEVICE void initial_state(const typename Eq::Params &prm,
                                 const double (&x)[Eq::dim],
                                 const double t,
                                 double (&U)[Eq::k]) const
    {
      constexpr int dim = Eq::dim;
      double xt[dim];
      transform_point<dim>(x, xt);
      Eq::initial_state_compute(prm, *this, xt, t, U);
      double m[dim];
#pragma unroll
      for (int d = 0; d < dim; ++d)
        m[d] = U[1 + d];
      transform_vector<dim>(m);
#pragma unroll
      for (int d = 0; d < dim; ++d)
        U[1 + d] = m[d];
    }
  };
} // namespace rjb
