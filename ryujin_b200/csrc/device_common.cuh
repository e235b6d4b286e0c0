// Shared device helpers.
//   positive_part / negative_part  <- source/simd.h:112-127
//   quadratic_newton_step          <- source/newton.h:39-101
#pragma once

#include "../../include/ryujin_b200.h"
#include "accurate_pow.h"
#include "vcl_pow.h"

#include <cfloat>
#include <cstdint>
#include <cuda_runtime.h>
#include <stdexcept>
#include <string>

#define RJ_DEVICE __device__ __forceinline__

/* The one place where kernels are launched (module.cu: launch()).  tests/host_emulation/ shadows
 * <cuda_runtime.h> and defines this macro to run the same kernel source on the host (CI without
 * a GPU); the product is only ever built by nvcc with the definition below. */
#ifndef RJ_LAUNCH_KERNEL
#define RJ_LAUNCH_KERNEL(kernel, grid, block, stream, ...) kernel<<<(grid), (block), 0, (stream)>>>(__VA_ARGS__)
#define RJ_LAUNCH_KERNEL_SMEM(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

/* Which pow() the equation modules call (every pow of the path goes through path_pow): the
 * reference's own pow -- Agner Fog's VCL as vendored under simd-math/, called by ryujin::pow,
 * source/simd.template.h:207-213 -- restated lane-wise in vcl_pow.h.  VCL has two arithmetic forms:
 *   2  vcl_pow<true>: as it compiles on an FMA host (AVX2 / AVX-512 builds of the reference, i.e.
 *      any current x86 server including the hosts of the GPU boxes); bit-identical to
 *      oracle/_ref/liboracle_vcl_fma.so.  DEFAULT (libryujin_b200.so): half the FP64 instructions
 *      of the non-FMA form (one fused operation instead of a 17-operation Dekker product).
 *   1  vcl_pow<false>: the non-FMA form (SSE2 ... AVX builds); bit-identical to
 *      oracle/_ref/liboracle_vcl.so.  Built as libryujin_b200_vcl.so; the parity suite runs both.
 * Either makes shock flows agree with the oracle to north_star's 1e-12 (in fact to the last bit).
 *   0  rj_pow (accurate_pow.h): 0.50x ulp table-driven pow, ~libm; fastest, but not the
 *      reference's rounding (build option: make RJ_POW=accurate). */
#ifndef RJ_POW_IMPL
#define RJ_POW_IMPL 2
#endif

namespace rjb
{
  constexpr int warp_size = 32; /* slice height C of the sliced-ELL layout */

  __device__ __forceinline__ double path_pow(const double x, const double y)
  {
#if RJ_POW_IMPL == 1
    return vcl_pow<false>(x, y);
#elif RJ_POW_IMPL == 2
    return vcl_pow<true>(x, y);
#else
    return rj_pow(x, y);
#endif
  }

  RJ_DEVICE double positive_part(const double x)
  {
    return 0.5 * (fabs(x) + x);
  }

  RJ_DEVICE double negative_part(const double x)
  {
    return 0.5 * (fabs(x) - x);
  }

  RJ_DEVICE void quadratic_newton_step(double &p_1,
                                       double &p_2,
                                       const double phi_p_1,
                                       const double phi_p_2,
                                       const double dphi_p_1,
                                       const double dphi_p_2,
                                       const double sign)
  {
    const double scaling = 1. / (p_2 - p_1 + DBL_EPSILON);

    const double dd_11 = dphi_p_1;
    const double dd_12 = (phi_p_2 - phi_p_1) * scaling;
    const double dd_22 = dphi_p_2;

    const double dd_112 = (dd_12 - dd_11) * scaling;
    const double dd_122 = (dd_22 - dd_12) * scaling;

    const double discriminant_1 = fabs(dphi_p_1 * dphi_p_1 - 4. * phi_p_1 * dd_112);
    const double discriminant_2 = fabs(dphi_p_2 * dphi_p_2 - 4. * phi_p_2 * dd_122);

    const double denominator_1 = dphi_p_1 + sign * sqrt(discriminant_1);
    const double denominator_2 = dphi_p_2 + sign * sqrt(discriminant_2);

    double t_1 = p_1 - (fabs(denominator_1) < DBL_EPSILON ? 0. : 2. * phi_p_1 / denominator_1);
    double t_2 = p_2 - (fabs(denominator_2) < DBL_EPSILON ? 0. : 2. * phi_p_2 / denominator_2);

    t_1 = fmax(p_1, t_1);
    t_1 = fmin(p_2, t_1);
    t_2 = fmax(p_1, t_2);
    t_2 = fmin(p_2, t_2);

    p_1 = fmin(t_1, t_2);
    p_2 = fmax(t_1, t_2);
  }

  /* host side error plumbing */
  struct CudaError : std::runtime_error {
    using std::runtime_error::runtime_error;
  };
  struct NcclError : std::runtime_error {
    using std::runtime_error::runtime_error;
  };
  struct InvalidArgument : std::runtime_error {
    using std::runtime_error::runtime_error;
  };
  struct NumericError : std::runtime_error {
    using std::runtime_error::runtime_error;
  };
  struct NoDevice : std::runtime_error {
    using std::runtime_error::runtime_error;
  };

  inline void cuda_check(cudaError_t err, const char *what, const char *file, int line)
  {
    if (err != cudaSuccess)
      throw CudaError(std::string(what) + " failed: " + cudaGetErrorString(err) + " (" + file +
                      ":" + std::to_string(line) + ")");
  }
} // namespace rjb

#define RJ_CUDA(expr) ::rjb::cuda_check((expr), #expr, __FILE__, __LINE__)
