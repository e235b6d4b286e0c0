// One persistent kernel per update for meshes that do not fill the GPU.
//
// prepare_state_vector() + step() of the reference (hyperbolic_module.template.h:97-193, 236-1211)
// are ~14 launches per update here (kernels.cuh); on a mesh of 10^4 .. 10^5 rows every one of them
// runs for 3-5 us and the update is bound by launch / dependency latency, not by the GPU (DESIGN.md
// section 5b: 2D Sod on 66k DoFs, 0.38 ms per ERK33 cycle under graph replay).  k_update_fused runs
// the SAME sweep bodies (the *_body functions of kernels.cuh, over virtual block indices) as phases of
// one cooperative kernel, separated by grid-wide barriers:
//
//   0  reset of the step scalars; boundary conditions            (prepare_state_vector, :122-144)
//   1  precomputed values and per-node flux data                 (:154-192)
//   1b second precomputation cycle over the stencil              (euler_aeos only)
//   2  d_ij for j > i (edge parallel) and alpha_i                (step 2, :341-424)
//   3  d_ij on coupling boundary pairs                           (step 3, :461-490)
//   4  d_ii and tau_max; every thread then forms tau itself      (:494-578)
//   5  low-order update, bounds, r_i, P_ij                       (step 4, :597-884)
//   6  P_ij correction, first l_ij                               (step 5, :892-1041)
//   7  symmetrized l_ij, high-order update, next l_ij            (step 6, :1053-1182)
//   8  final high-order update                                   (step 7)
//
// Same arithmetic in the same order as the separate kernels: results are bit-identical (parity suite
// with RYUJIN_B200_FUSED=0/1).  Single rank only (no exchange inside the kernel); the module uses it
// for device-resident cycles when the mesh is small (module.cu: fused_eligible()).
#pragma once

#include "kernels.cuh"

#ifndef RJ_FUSED_MAX_ROWS_DEFAULT
#define RJ_FUSED_MAX_ROWS_DEFAULT 262144u /* automatic mode: meshes up to this many rows (tools/small_mesh_stepping.py) */
#endif
#ifndef RJ_FUSED
#define RJ_FUSED 1 /* 0: the library is built without the fused kernels (the second pow form: shorter build) */
#endif
#ifndef RJ_FUSED_MAX_STAGES
#define RJ_FUSED_MAX_STAGES 2
#endif
#ifndef RJ_MB_FUSED
#define RJ_MB_FUSED 4 /* resident blocks of 128 threads per SM: 128 registers, as the heaviest phase (step 4) */
#endif

#if defined(__CUDA_ARCH__)
/* fence.acq_rel.gpu instead of __threadfence() (= fence.sc.gpu): the barrier needs release / acquire only */
#define RJ_FENCE_GPU() asm volatile("fence.acq_rel.gpu;" ::: "memory")
#else
#define RJ_FENCE_GPU() __threadfence()
#endif

namespace rjb
{
  /* Grid-wide barrier of a cooperative launch: one arrival counter that only grows during a launch
   * (block b waits until it reaches epoch * gridDim.x) and is reset by the last block that leaves the
   * kernel.  Thread 0 of every block: fence, arrive, spin on a volatile load, fence; __syncthreads() on
   * both sides extends the ordering to the block (the cooperative-groups recipe). */
  struct GridBarrier {
    unsigned int count;
    unsigned int finished;
  };

  RJ_DEVICE void grid_sync(GridBarrier *barrier, unsigned int &epoch)
  {
    __syncthreads();
    if (threadIdx.x == 0) {
      epoch += gridDim.x;
      RJ_FENCE_GPU(); /* release: this block's writes (ordered before by the __syncthreads above) */
      atomicAdd(&barrier->count, 1u);
      while (*(volatile unsigned int *)&barrier->count < epoch)
        RJ_SPIN_YIELD();
      RJ_FENCE_GPU(); /* acquire: invalidates this SM's L1 (CCTL.IVALL) */
    }
    __syncthreads();
  }

  template <class Eq, int stages>
  struct FusedUpdate {
    StencilView v;
    typename Eq::Params prm;
    InitialValuesDevice iv;
    /* boundary conditions */
    uint32_t n_bgroups;
    const uint32_t *bgroup_dof, *bgroup_start;
    const double *bnormal, *bposition;
    const int32_t *bid;
    /* coupling boundary pairs */
    uint32_t n_pairs;
    const uint32_t *pair_i, *pair_j, *pair_e, *pair_eT;
    const double *pair_cji;
    uint32_t n_relevant, n_slices;
    /* old state (prepared here if do_prepare), stage states, new state */
    double *U, *prec, *aux;
    const double *initial_precomputed;
    StagePointers<stages> st;
    double *U_new;
    /* module scratch */
    double *alpha, *dij, *edge_cache, *r, *bounds, *pij, *lij, *lij_next;
    StepScalars *scalars;
    TauSource tau_source;
    CycleScalars *cycle;
    int substep;
    double tau_factor; /* Dirichlet time t + tau_factor * tau of the prepare */
    double cfl;
    unsigned int n_iterations;
    int do_prepare;
    GridBarrier *barrier;
  };

  /* virtual blocks vb = blockIdx.x, blockIdx.x + gridDim.x, ... < n_vblocks */
#define RJ_FOR_VBLOCKS(vb, n_vblocks) for (uint32_t vb = blockIdx.x, vb##_end = (n_vblocks); vb < vb##_end; vb += gridDim.x)

  template <class Eq, int stages>
  __global__ void __launch_bounds__(128, RJ_MB_FUSED) k_update_fused(const FusedUpdate<Eq, stages> a)
  {
    constexpr uint32_t B = 128; /* blockDim.x: row kernels, and 32 rows x 4 warps of the d_ij phase */
    constexpr bool cache = Eq::edge_cache && (RJ_EDGE_CACHE != 0);
    unsigned int epoch = 0;
    const uint32_t n_owned = a.v.n_owned;
    const uint32_t row_blocks = (n_owned + B - 1) / B;

    /* phase 0 */
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      StepScalars *s = a.scalars;
      s->tau_max_bits = (unsigned long long)__double_as_longlong(DBL_MAX);
      s->tau_max = 0.;
      s->tau = 0.;
      s->restart_needed = 0;
      s->numeric_error = 0;
      s->blocks_done_tau = 0;
      s->blocks_done_step = 0;
    }
    if (a.do_prepare) {
      RJ_FOR_VBLOCKS(vb, (a.n_bgroups + B - 1) / B)
        apply_boundary_conditions_body<Eq, InitialValuesDevice>(vb, a.n_bgroups, a.bgroup_dof, a.bgroup_start, a.bnormal,
                                                                 a.bposition, a.bid, a.v.ld, a.prm, a.iv, 0., a.cycle,
                                                                 a.tau_factor, a.U);
      grid_sync(a.barrier, epoch);
      /* phase 1 */
      RJ_FOR_VBLOCKS(vb, (a.n_relevant + B - 1) / B)
        precompute_body<Eq>(vb, a.v, a.prm, 0u, a.n_relevant, a.U, a.initial_precomputed, a.prec, a.aux);
      grid_sync(a.barrier, epoch);
      if constexpr (Eq::aeos) {
        RJ_FOR_VBLOCKS(vb, row_blocks)
          precompute_stencil_body<Eq>(vb, a.v, a.prm, 0u, n_owned, a.U, a.prec);
        grid_sync(a.barrier, epoch);
      }
    }

    /* phase 2 */
    RJ_FOR_VBLOCKS(vb, a.n_slices)
      dij_body<Eq, 4>(vb, a.v, a.prm, 0u, a.n_slices, a.U, a.prec, a.dij);
    RJ_FOR_VBLOCKS(vb, row_blocks)
      alpha_body<Eq, cache>(vb, a.v, a.prm, 0u, n_owned, a.U, a.prec, a.aux, a.alpha, a.edge_cache);
    grid_sync(a.barrier, epoch);

    /* phase 3 */
    if (a.n_pairs > 0) {
      RJ_FOR_VBLOCKS(vb, (a.n_pairs + B - 1) / B)
        boundary_dij_body<Eq>(vb, a.n_pairs, a.pair_i, a.pair_j, a.pair_e, a.pair_eT, a.pair_cji, a.v.ld, a.v.E_owned, a.prm,
                              a.U, a.prec, a.dij);
      grid_sync(a.barrier, epoch);
    }

    /* phase 4 */
    RJ_FOR_VBLOCKS(vb, row_blocks)
    {
      dii_tau_body(vb, a.v, 0u, n_owned, a.cfl, a.dij, a.scalars);
      __syncthreads(); /* the block's reduction buffer is reused by the next virtual block */
    }
    grid_sync(a.barrier, epoch);

    /* k_tau_local + k_tau_finalize, evaluated by every thread (a pure function of values that are
     * final after the barrier); one thread stores the results */
    double tau;
    {
      const unsigned long long bits = *(volatile unsigned long long *)&a.scalars->tau_max_bits;
      const double local = __longlong_as_double((long long)bits);
      const TauSource &src = a.tau_source;
      double tau_max_in = src.tau_max;
      if (src.mode == 1) {
        const double cycle_tau_max = *(volatile double *)&a.cycle->tau_max;
        tau_max_in = cycle_tau_max == DBL_MAX ? DBL_MAX : __ddiv_rn(cycle_tau_max, src.divisor);
      } else if (src.mode == 2)
        tau_max_in = DBL_MAX;
      const double tau_max = fmin(local, tau_max_in);
      const bool numeric_error = isnan(tau_max) || isinf(tau_max) || !(tau_max > 0.);
      const double tau_in = src.mode == 0 ? src.tau : (src.mode == 1 ? 0. : *(volatile double *)&a.cycle->tau);
      tau = (tau_in == 0. ? tau_max : tau_in);
      if (blockIdx.x == 0 && threadIdx.x == 0) {
        a.scalars->tau_max = tau_max;
        a.scalars->tau = tau;
        if (numeric_error)
          a.scalars->numeric_error = 1;
        if (src.mode == 1)
          a.cycle->tau = tau;
      }
    }

    /* phase 5 */
    RJ_FOR_VBLOCKS(vb, row_blocks)
      low_order_body<Eq, stages>(vb, tau, a.v, a.prm, 0u, n_owned, a.U, a.prec, a.aux, a.st, a.alpha, a.dij, a.edge_cache,
                                 a.U_new, a.r, a.bounds, a.pij);

    if (a.n_iterations != 0) {
      grid_sync(a.barrier, epoch);
      /* phase 6 */
      RJ_FOR_VBLOCKS(vb, row_blocks)
        pij_lij_body<Eq>(vb, tau, a.v, a.prm, 0u, n_owned, a.U_new, a.r, a.bounds, a.scalars, a.pij, a.lij);
      grid_sync(a.barrier, epoch);
      /* phases 7, 8 */
      double *lij = a.lij, *lij_next = a.lij_next;
      for (unsigned int pass = 0; pass < a.n_iterations; ++pass) {
        const bool last_round = (pass + 1 == a.n_iterations);
        if (a.n_iterations == 2 && last_round) {
          double *tmp = lij;
          lij = lij_next;
          lij_next = tmp;
        }
        if (last_round) {
          RJ_FOR_VBLOCKS(vb, row_blocks)
            high_order_body<Eq, true>(vb, a.v, a.prm, 0u, n_owned, a.bounds, a.pij, lij, lij_next, a.U_new);
        } else {
          RJ_FOR_VBLOCKS(vb, row_blocks)
            high_order_body<Eq, false>(vb, a.v, a.prm, 0u, n_owned, a.bounds, a.pij, lij, lij_next, a.U_new);
          grid_sync(a.barrier, epoch);
        }
      }
    }

    /* the block that leaves last records the substep's flags (k_cycle_record) and re-arms the barrier */
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      const unsigned int before = atomicAdd(&a.barrier->finished, 1u);
      if (before == gridDim.x - 1) {
        __threadfence();
        if (a.cycle) {
          a.cycle->restart[a.substep] = *(volatile int *)&a.scalars->restart_needed;
          a.cycle->numeric_error[a.substep] = *(volatile int *)&a.scalars->numeric_error;
        }
        /* leave the scalars clean, as the folded scalar kernels of the separate path expect them */
        a.scalars->restart_needed = 0;
        a.scalars->numeric_error = 0;
        a.scalars->tau_max_bits = (unsigned long long)__double_as_longlong(DBL_MAX);
        a.barrier->count = 0;
        a.barrier->finished = 0;
      }
    }
  }
} // namespace rjb
