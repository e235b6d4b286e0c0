// The sweep kernels of the explicit hyperbolic update, one per loop of
// HyperbolicModule::step() (source/hyperbolic_module.template.h:236-1211) and
// prepare_state_vector() (:97-193).  Templated on the equation module Eq
// (euler.cuh, shallow_water.cuh): the same plug-in surface as the reference's
// Description (HyperbolicSystemView / RiemannSolver / Indicator / Limiter).
//
// Mapping: one thread per stencil row; lane l of a warp owns row 32 g + l of
// slice g, so all matrix loads (c_ij, m_ij, d_ij, l_ij, P_ij, column indices)
// are coalesced 128/256 byte lines, and neighbour gathers U_j are contiguous
// whenever the DoF ordering keeps neighbours of consecutive rows consecutive.
// Per-row accumulations stay in one thread in column order, exactly as the
// reference's scalar loop; warp shuffles are only used for the CFL minimum.
// No tensor cores: the per-edge work is k = dim+2 component FP64 algebra.
#pragma once

#include "device_common.cuh"

/* minimum resident blocks of 128 threads per SM requested from ptxas for the four
 * FP64-heavy kernels (caps registers at 65536 / (128 * MB)); tuned in profiles/ */
#ifndef RJ_MB_DIJ
#define RJ_MB_DIJ 12 /* blocks of 32 * RJ_DIJ_W threads: 42 registers; r2d: 8 -> 5.29 ms, 10 -> 4.93, 12 -> 4.80 */
#endif
#define RJ_PRAGMA_(x) _Pragma(#x)
#define RJ_PRAGMA(x) RJ_PRAGMA_(x)
#define RJ_UNROLL(n) RJ_PRAGMA(unroll n)
/* unroll factors of the column loops (software pipelining of the index -> gather chain) */
#ifndef RJ_UNROLL_ALPHA
#define RJ_UNROLL_ALPHA 1
#endif
#ifndef RJ_UNROLL_DII
#define RJ_UNROLL_DII 4
#endif
#ifndef RJ_UNROLL_LOW
#define RJ_UNROLL_LOW 3
#endif
#ifndef RJ_UNROLL_PIJ
#define RJ_UNROLL_PIJ 1
#endif
#ifndef RJ_UNROLL_HO1
#define RJ_UNROLL_HO1 4
#endif
#ifndef RJ_UNROLL_HO2
#define RJ_UNROLL_HO2 3
#endif
#ifndef RJ_HO2_REVERSE
#define RJ_HO2_REVERSE 1
#endif
#if RJ_HO2_REVERSE
#define RJ_HO2_FIRST(len) ((len)-1)
#define RJ_HO2_COND(col, len) ((col) >= 1)
#define RJ_HO2_NEXT(col) (--(col))
#else
#define RJ_HO2_FIRST(len) 1
#define RJ_HO2_COND(col, len) ((col) < (len))
#define RJ_HO2_NEXT(col) (++(col))
#endif
#ifndef RJ_PREFETCH_ALPHA
#define RJ_PREFETCH_ALPHA 0
#endif
#ifndef RJ_PREFETCH_LOW
#define RJ_PREFETCH_LOW 1
#endif
#ifndef RJ_PREFETCH_PIJ
#define RJ_PREFETCH_PIJ 1
#endif
#ifndef RJ_PREFETCH_HO
#define RJ_PREFETCH_HO 0
#endif
#ifndef RJ_TILE_PREFETCH_MAXCOL
#define RJ_TILE_PREFETCH_MAXCOL 27 /* columns whose l_ji the tiled step 6 keeps in registers (Q1 in 3D) */
#endif
#ifndef RJ_DIJ_W
#define RJ_DIJ_W 4 /* warps per slice of the edge-parallel d_ij kernel */
#endif
/* Edge cache: a quantity of step 4 that is symmetric in (i, j) is evaluated once per edge instead of twice.
 * For the polytropic Euler limiter that is the specific entropy of the midpoint state (one pow + one division
 * per stencil entry, limiter.h:314-319): (U_i + U_j) * .5 is the same state in row i and in row j bit for
 * bit, so are the cached values.  k_alpha evaluates it for the entries j > i (it has U_j gathered and is bound
 * by gather latency, not by FP64 issue) and mirrors it into (j, i) as k_dij mirrors d_ij; step 4 reads it.
 * Measured on the B200 (profiles/r2h, r2i): step 4 -1.4 ms, k_alpha +1.1 ms per update of 8M DoFs; doing it
 * in k_dij instead (edge parallel, FP64-issue bound already) gave +1.2 ms there: removed.
 *   0: off (step 4 evaluates the entropy per entry)   1: on */
#ifndef RJ_EDGE_CACHE
#define RJ_EDGE_CACHE 1
#endif
#ifndef RJ_MB_ALPHA
#define RJ_MB_ALPHA 8
#endif
#ifndef RJ_MB_DIJ_ALPHA
#define RJ_MB_DIJ_ALPHA 8 /* the merged step-2 launch (k_dij_alpha) */
#endif
#ifndef RJ_MB_ALPHA_CACHE
#define RJ_MB_ALPHA_CACHE 6 /* with the edge cache (one pow per entry j > i): 8 -> 3.56 ms, 6 -> 3.69, 5 -> 4.69 next to step 4 at 6.42 / 6.29 / 6.19 */
#endif
#ifndef RJ_MB_LOW
#define RJ_MB_LOW 4
#endif
#ifndef RJ_MB_LOW_SW
#define RJ_MB_LOW_SW 2 /* shallow water: star states + sources need ~200 registers */
#endif
#ifndef RJ_MB_PIJ
#define RJ_MB_PIJ 5
#endif
#ifndef RJ_MB_HO
#define RJ_MB_HO 6
#endif

/* L2 prefetch of the streamed matrix side (cols, c_ij, d_ij, edge cache) of the row-parallel sweeps
 * RJ_L2_PREFETCH columns ahead: the entries of a slice are one contiguous run per array, a column further is
 * 256 bytes further.  No registers, no shared memory; 0 = off */
#ifndef RJ_L2_PREFETCH
#define RJ_L2_PREFETCH 2 /* B200, 8M DoFs: step 4 6.17 -> 5.77 / 5.46 / 5.60 / 5.66 / 5.70 ms for 1 .. 5 columns, k_alpha 3.58 -> 3.39 */
#endif
#ifndef RJ_L2_PREFETCH_LOW
#define RJ_L2_PREFETCH_LOW RJ_L2_PREFETCH
#endif
#ifndef RJ_L2_PREFETCH_ALPHA
#define RJ_L2_PREFETCH_ALPHA RJ_L2_PREFETCH
#endif
#if defined(__CUDA_ARCH__)
#define RJ_PREFETCH_L2(pointer) asm volatile("prefetch.global.L2 [%0];" ::"l"(pointer))
#else
#define RJ_PREFETCH_L2(pointer) ((void)(pointer))
#endif

namespace rjb
{
  struct StencilView {
    uint32_t n_owned, ld, E_owned;
    const uint32_t *__restrict__ slice_offsets;
    const uint8_t *__restrict__ row_length;
    const uint32_t *__restrict__ cols;
    const uint32_t *__restrict__ transposed;
    const double *__restrict__ mij;
    const double *__restrict__ cij;
    const double *__restrict__ mi;
    const double *__restrict__ mi_inv;
    double measure_of_omega_inverse;
  };

  /* device scalars of one step() */
  struct StepScalars {
    unsigned long long tau_max_bits; /* min over rows, ordered-bits of a positive double */
    double tau_max;                  /* after min with caller's tau_max (and all-reduce) */
    double tau;                      /* time step used by the update kernels */
    int restart_needed;
    int numeric_error;
    /* last-block-done counters of the kernels that carry the scalar work of a substep (k_dii_tau_final,
     * k_high_order_record); zero between launches */
    unsigned int blocks_done_tau, blocks_done_step;
  };

  /* device scalars of one TimeIntegrator cycle (device-resident stepping: the substeps of a
   * cycle are enqueued without host round trips, time_integrator.template.h:302-443) */
  constexpr int max_substeps = 8;
  struct CycleScalars {
    double t;       /* time at the start of the cycle            */
    double tau_max; /* t_final - t (DBL_MAX if unbounded)        */
    double tau;     /* step size fixed by the first substep      */
    int restart[max_substeps];
    int numeric_error[max_substeps];
  };

  template <int n>
  struct StagePointers {
    const double *U[n > 0 ? n : 1];
    const double *prec[n > 0 ? n : 1];
    const double *aux[n > 0 ? n : 1];
    double weight[n > 0 ? n : 1];
  };

  template <int k>
  RJ_DEVICE void load_state(const double *__restrict__ v, uint32_t ld, uint32_t i, double (&out)[k])
  {
#pragma unroll
    for (int q = 0; q < k; ++q)
      out[q] = v[size_t(q) * ld + i];
  }

  template <int k>
  RJ_DEVICE void store_state(double *__restrict__ v, uint32_t ld, uint32_t i, const double (&in)[k])
  {
#pragma unroll
    for (int q = 0; q < k; ++q)
      v[size_t(q) * ld + i] = in[q];
  }

  /* Software pipelining of the index -> gather chain of the row-parallel sweeps: the column (or
   * transposed) index of the NEXT column is loaded before the gathers of the current one are
   * consumed, so that the two dependent memory latencies of an iteration overlap.  Same loads, same
   * arithmetic, different issue order.  Per kernel (measured on 8.0M DoFs, profiles/r2a_variants.txt:
   * k_low_order 8.32 -> 8.08 ms, k_pij_lij 5.35 -> 4.75 ms, k_alpha 2.47 -> 2.69 ms, k_high_order
   * unchanged): on for step 4 and step 5 only. */
  template <bool prefetch>
  RJ_DEVICE uint32_t
  pipelined_index(const uint32_t *__restrict__ idx, uint32_t e, uint32_t col, uint32_t len, uint32_t &next)
  {
    if (prefetch) {
      const uint32_t current = next;
      if (col + 1 < len)
        next = idx[e + 32];
      return current;
    }
    return idx[e];
  }
#define RJ_INDEX_INIT(pf, name, idx, e0) uint32_t name = (pf) ? (idx)[e0] : 0u;
#define RJ_INDEX(pf, name, idx, e, col, len) pipelined_index<(pf) != 0>((idx), (e), (col), (len), name)

  /* common row prologue: returns false if the thread has no row to work on */
  struct RowInfo {
    uint32_t i, len, width, base;
  };

  RJ_DEVICE bool row_info(const uint32_t vblock, const StencilView &v, uint32_t row_begin, uint32_t row_end, RowInfo &r)
  {
    r.i = row_begin + vblock * blockDim.x + threadIdx.x;
    if (r.i >= row_end)
      return false;
    const uint32_t g = r.i >> 5;
    const uint32_t s0 = v.slice_offsets[g];
    r.width = (v.slice_offsets[g + 1] - s0) >> 5;
    r.base = s0 + (r.i & 31u);
    r.len = v.row_length[r.i];
    return true;
  }


  /* ---------------------------------------------------------------------- *
   * Step 1a: boundary conditions, hyperbolic_module.template.h:122-144.
   * One thread per boundary DoF; entries of a DoF are applied in stored order.
   * ---------------------------------------------------------------------- */
  template <class Eq, class IV>
  RJ_DEVICE void apply_boundary_conditions_body(const uint32_t vblock,
                                                uint32_t n_groups,
                                                const uint32_t *__restrict__ group_dof,
                                                const uint32_t *__restrict__ group_start,
                                                const double *__restrict__ bnormal,
                                                const double *__restrict__ bposition,
                                                const int32_t *__restrict__ bid,
                                                uint32_t ld,
                                                const typename Eq::Params &prm,
                                                const IV &iv,
                                                double t_host,
                                                const CycleScalars *__restrict__ cycle,
                                                double tau_factor,
                                                double *__restrict__ U)
  {
    constexpr int k = Eq::k, dim = Eq::dim;
    const uint32_t gidx = vblock * blockDim.x + threadIdx.x;
    if (gidx >= n_groups)
      return;
    /* t + c * tau exactly as the host computes it (time_integrator.template.h:380,395) */
    const double t =
        cycle ? (tau_factor == 0. ? cycle->t : __dadd_rn(cycle->t, __dmul_rn(tau_factor, cycle->tau))) : t_host;
    const uint32_t i = group_dof[gidx];
    double U_i[k];
    load_state<k>(U, ld, i, U_i);
    for (uint32_t b = group_start[gidx]; b < group_start[gidx + 1]; ++b) {
      const int id = bid[b];
      if (id == RJ_BOUNDARY_DO_NOTHING)
        continue;
      double normal[dim], position[dim];
#pragma unroll
      for (int d = 0; d < dim; ++d) {
        normal[d] = bnormal[size_t(b) * dim + d];
        position[d] = bposition[size_t(b) * dim + d];
      }
      Eq::apply_boundary_conditions(
          prm, id, U_i, normal, [&](double(&D)[k]) { iv.template initial_state<Eq>(prm, position, t, D); });
    }
    store_state<k>(U, ld, i, U_i);
  }

  template <class Eq, class IV>
  __global__ void k_apply_boundary_conditions(uint32_t n_groups, const uint32_t *__restrict__ group_dof, const uint32_t *__restrict__ group_start, const double *__restrict__ bnormal, const double *__restrict__ bposition, const int32_t *__restrict__ bid, uint32_t ld, typename Eq::Params prm, IV iv, double t_host, const CycleScalars *__restrict__ cycle, double tau_factor, double *__restrict__ U)
  {
    apply_boundary_conditions_body<Eq, IV>(blockIdx.x, n_groups, group_dof, group_start, bnormal, bposition, bid, ld, prm, iv, t_host, cycle, tau_factor, U);
  }


  /* U <- initial_state(x_i, t): InitialValues::interpolate_hyperbolic_vector */
  template <class Eq, class IV>
  __global__ void k_interpolate(uint32_t n_owned,
                                uint32_t ld,
                                const double *__restrict__ positions,
                                typename Eq::Params prm,
                                IV iv,
                                double t,
                                double *__restrict__ U)
  {
    constexpr int k = Eq::k, dim = Eq::dim;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_owned)
      return;
    double x[dim];
#pragma unroll
    for (int d = 0; d < dim; ++d)
      x[d] = positions[size_t(d) * n_owned + i];
    double U_i[k];
    iv.template initial_state<Eq>(prm, x, t, U_i);
    store_state<k>(U, ld, i, U_i);
  }


  /* ---------------------------------------------------------------------- *
   * Step 1b: precomputed values, hyperbolic_module.template.h:154-192 and
   * euler/hyperbolic_system.h:705-737.
   * ---------------------------------------------------------------------- */
  template <class Eq>
  RJ_DEVICE void precompute_body(const uint32_t vblock,
                                 const StencilView &v,
                                 const typename Eq::Params &prm,
                                 uint32_t row_begin,
                                 uint32_t row_end,
                                 const double *__restrict__ U,
                                 const double *__restrict__ initial_precomputed,
                                 double *__restrict__ prec,
                                 double *__restrict__ aux)
  {
    /* owned rows as in the reference; ghost rows are recomputed locally from the
     * exchanged ghost state (bit-identical to the owner's values), which replaces
     * the ghost exchange of the precomputed vector (C2, :157-160) */
    constexpr int k = Eq::k, np = Eq::n_prec, na = Eq::n_aux;
    const uint32_t i = row_begin + vblock * blockDim.x + threadIdx.x;
    if (i >= row_end)
      return;
    if (i < v.n_owned && v.row_length[i] <= 1)
      return;
    double U_i[k], prec_i[np], aux_i[na];
    load_state<k>(U, v.ld, i, U_i);
    Eq::precompute(prm, U_i, prec_i);
    if constexpr (Eq::shallow_water)
      /* flux_contribution = (U_i, Z_i): the bathymetry is the "initial precomputed"
       * vector of the reference, shallow_water/hyperbolic_system.h:710-728 */
      Eq::store_aux(Eq::flux_contribution(prm, prec_i, U_i, initial_precomputed[i]), aux_i);
    else
      Eq::store_aux(Eq::flux_contribution(prm, prec_i, U_i), aux_i);
    store_state<np>(prec, v.ld, i, prec_i);
    store_state<na>(aux, v.ld, i, aux_i);
  }

  template <class Eq>
  __global__ void k_precompute(StencilView v, typename Eq::Params prm, uint32_t row_begin, uint32_t row_end, const double *__restrict__ U, const double *__restrict__ initial_precomputed, double *__restrict__ prec, double *__restrict__ aux)
  {
    precompute_body<Eq>(blockIdx.x, v, prm, row_begin, row_end, U, initial_precomputed, prec, aux);
  }


  /* Second precomputation cycle of equations that have one (euler_aeos/hyperbolic_system.h:
   * 943-973): gamma_min_i = min over the stencil of the surrogate gamma of (U_j, p_j), then the
   * surrogate entropies of row i.  Owned rows only (a ghost row's stencil is not local): the
   * precomputed vector is exchanged afterwards, as the reference does between its cycles.  Reads
   * component 0 of other rows, writes components 1.. of its own row: no hazard within the launch. */
  template <class Eq>
  RJ_DEVICE void precompute_stencil_body(const uint32_t vblock,
                                         const StencilView &v,
                                         const typename Eq::Params &prm,
                                         uint32_t row_begin,
                                         uint32_t row_end,
                                         const double *__restrict__ U,
                                         double *prec)
  {
    constexpr int k = Eq::k, np = Eq::n_prec;
    RowInfo r;
    if (!row_info(vblock, v, row_begin, row_end, r) || r.len <= 1)
      return;
    const uint32_t i = r.i;
    double U_i[k], prec_i[np];
    load_state<k>(U, v.ld, i, U_i);
    load_state<np>(prec, v.ld, i, prec_i);
    double gamma_min = prec_i[1];
    for (uint32_t col = 1; col < r.len; ++col) {
      const uint32_t j = v.cols[r.base + col * 32];
      double U_j[k];
      load_state<k>(U, v.ld, j, U_j);
      const double p_j = prec[j];
      gamma_min = fmin(gamma_min, Eq::surrogate_gamma(prm, U_j, p_j));
    }
    prec_i[1] = gamma_min;
    Eq::precompute_finish(prm, U_i, prec_i);
#pragma unroll
    for (int q = 1; q < np; ++q)
      prec[size_t(q) * v.ld + i] = prec_i[q];
  }

  template <class Eq>
  __global__ void __launch_bounds__(128) k_precompute_stencil(StencilView v, typename Eq::Params prm, uint32_t row_begin, uint32_t row_end, const double *__restrict__ U, double *prec)
  {
    precompute_stencil_body<Eq>(blockIdx.x, v, prm, row_begin, row_end, U, prec);
  }


  /* ---------------------------------------------------------------------- *
   * Step 2: off-diagonal d_ij for j > i and alpha_i,
   * hyperbolic_module.template.h:341-424, as two kernels:
   *
   *  k_dij   EDGE parallel: block = one slice of 32 rows x W warps; warp w takes
   *          columns w, w+W, ... of the slice, so one thread runs one or two
   *          Riemann solves instead of ~13 in sequence.  The per-edge solve is a
   *          long dependent FP64 chain (2 pow, 7 sqrt, 9 div); spreading it over
   *          8x more, lighter threads is what fills the FP64 pipe (profiles/).
   *          Also mirrors d_ij into the transposed slot (j, i), which replaces
   *          the gather of step 3 (:511-535).
   *  k_alpha ROW parallel: the entropy-viscosity commutator, accumulated in
   *          column order in one thread as the reference does.
   * ---------------------------------------------------------------------- */
  template <class Eq, int W>
  RJ_DEVICE void dij_body(const uint32_t vblock,
                          const StencilView &v,
                          const typename Eq::Params &prm,
                          uint32_t slice_begin,
                          uint32_t slice_end,
                          const double *__restrict__ U,
                          const double *__restrict__ prec,
                          double *__restrict__ dij)
  {
    constexpr int k = Eq::k, dim = Eq::dim, np = Eq::n_prec;
    const uint32_t g = slice_begin + vblock;
    if (g >= slice_end)
      return;
    const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
    const uint32_t i = (g << 5) + lane;
    if (i >= v.n_owned)
      return;
    const uint32_t len = v.row_length[i];
    if (len <= 1)
      return;
    const uint32_t base = v.slice_offsets[g] + lane;

    double U_i[k], prec_i[np];
    load_state<k>(U, v.ld, i, U_i);
    load_state<np>(prec, v.ld, i, prec_i);

    for (uint32_t col = w; col < len; col += W) {
      if (col == 0)
        continue;
      const uint32_t e = base + col * 32;
      const uint32_t j = v.cols[e];
      if (j < i)
        continue;
      double U_j[k], prec_j[np], c_ij[dim];
      load_state<k>(U, v.ld, j, U_j);
      load_state<np>(prec, v.ld, j, prec_j);
#pragma unroll
      for (int d = 0; d < dim; ++d)
        c_ij[d] = v.cij[size_t(d) * v.E_owned + e];

      double nrm2 = c_ij[0] * c_ij[0];
#pragma unroll
      for (int d = 1; d < dim; ++d)
        nrm2 += c_ij[d] * c_ij[d];
      const double nrm = sqrt(nrm2);
      const double inv = 1. / nrm;
      double n_ij[dim];
#pragma unroll
      for (int d = 0; d < dim; ++d)
        n_ij[d] = c_ij[d] * inv;
      const double lambda_max = Eq::riemann_lambda_max(prm, U_i, prec_i, U_j, prec_j, n_ij);
      const double d_ij = nrm * lambda_max;
      dij[e] = d_ij;
      const uint32_t eT = v.transposed[e];
      if (eT < v.E_owned)
        dij[eT] = d_ij;
    }
  }

  template <class Eq, int W>
  __global__ void __launch_bounds__(32 * W, RJ_MB_DIJ) k_dij(StencilView v, typename Eq::Params prm, uint32_t slice_begin, uint32_t slice_end, const double *__restrict__ U, const double *__restrict__ prec, double *__restrict__ dij)
  {
    dij_body<Eq, W>(blockIdx.x, v, prm, slice_begin, slice_end, U, prec, dij);
  }

  template <class Eq, bool do_cache>
  RJ_DEVICE void alpha_body(const uint32_t vblock,
                            const StencilView &v,
                            const typename Eq::Params &prm,
                            uint32_t row_begin,
                            uint32_t row_end,
                            const double *__restrict__ U,
                            const double *__restrict__ prec,
                            const double *__restrict__ aux,
                            double *__restrict__ alpha,
                            double *__restrict__ edge_cache)
  {
    constexpr int k = Eq::k, dim = Eq::dim, np = Eq::n_prec, na = Eq::n_aux;
    RowInfo r;
    if (!row_info(vblock, v, row_begin, row_end, r) || r.len <= 1)
      return;
    const uint32_t i = r.i;

    double U_i[k], prec_i[np];
    load_state<k>(U, v.ld, i, U_i);
    load_state<np>(prec, v.ld, i, prec_i);

    typename Eq::Indicator indicator;
    indicator.reset(prm, prec_i, U_i);

    RJ_INDEX_INIT(RJ_PREFETCH_ALPHA, j_next, v.cols, r.base)
    RJ_UNROLL(RJ_UNROLL_ALPHA)
    for (uint32_t col = 0; col < r.len; ++col) {
      const uint32_t e = r.base + col * 32;
      const uint32_t j = RJ_INDEX(RJ_PREFETCH_ALPHA, j_next, v.cols, e, col, r.len);
      if (RJ_L2_PREFETCH_ALPHA > 0 && col + RJ_L2_PREFETCH_ALPHA < r.len) {
        const uint32_t ahead = e + 32 * RJ_L2_PREFETCH_ALPHA;
        RJ_PREFETCH_L2(&v.cols[ahead]);
#pragma unroll
        for (int d = 0; d < dim; ++d)
          RJ_PREFETCH_L2(&v.cij[size_t(d) * v.E_owned + ahead]);
        if constexpr (do_cache)
          RJ_PREFETCH_L2(&v.transposed[ahead]);
      }
      double U_j[k], prec_j[np], aux_j[na], c_ij[dim];
      load_state<k>(U, v.ld, j, U_j);
      load_state<np>(prec, v.ld, j, prec_j);
      load_state<na>(aux, v.ld, j, aux_j);
#pragma unroll
      for (int d = 0; d < dim; ++d)
        c_ij[d] = v.cij[size_t(d) * v.E_owned + e];
      indicator.accumulate(prm, U_i, prec_i, U_j, prec_j, Eq::load_aux(aux_j), c_ij);
      if constexpr (do_cache) {
        /* the symmetric edge quantity of step 4, for j > i, mirrored (as k_dij mirrors d_ij) */
        if (col == 0)
          edge_cache[e] = Eq::edge_value_diagonal(prec_i);
        else if (j > i) {
          const double value = Eq::edge_value(prm, U_i, U_j);
          edge_cache[e] = value;
          const uint32_t eT = v.transposed[e];
          if (eT < v.E_owned)
            edge_cache[eT] = value;
        }
      }
    }

    const double hd_i = v.mi[i] * v.measure_of_omega_inverse;
    alpha[i] = indicator.alpha(prm, hd_i);
  }

  template <class Eq, bool do_cache = false>
  __global__ void __launch_bounds__(128, do_cache ? RJ_MB_ALPHA_CACHE : RJ_MB_ALPHA) k_alpha(StencilView v, typename Eq::Params prm, uint32_t row_begin, uint32_t row_end, const double *__restrict__ U, const double *__restrict__ prec, const double *__restrict__ aux, double *__restrict__ alpha, double *__restrict__ edge_cache)
  {
    alpha_body<Eq, do_cache>(blockIdx.x, v, prm, row_begin, row_end, U, prec, aux, alpha, edge_cache);
  }


  /* Small single-rank meshes: d_ij and alpha_i (both read the prepared state only) as ONE launch of
   * heterogeneous blocks -- blocks [0, n_slices) run the edge-parallel d_ij body, the rest the row-parallel
   * indicator -- one dependency latency less per update where launches, not bandwidth, bound the update. */
  template <class Eq, bool alpha_cache>
  __global__ void __launch_bounds__(128, RJ_MB_DIJ_ALPHA) k_dij_alpha(StencilView v, typename Eq::Params prm, uint32_t n_slices, const double *__restrict__ U, const double *__restrict__ prec, const double *__restrict__ aux, double *__restrict__ dij, double *__restrict__ alpha, double *__restrict__ edge_cache)
  {
    /* groups of five blocks: four slices of the d_ij body and the row block (128 rows = the same four
     * slices) of the indicator body, so that the FP64-issue bound solves and the latency bound gathers of
     * step 2 are resident on an SM at the same time; grid = 5 * ceil(n_owned / 128) */
    const uint32_t group = blockIdx.x / 5, role = blockIdx.x % 5;
    if (role < 4)
      dij_body<Eq, 4>(4 * group + role, v, prm, 0u, n_slices, U, prec, dij);
    else
      alpha_body<Eq, alpha_cache>(group, v, prm, 0u, v.n_owned, U, prec, aux, alpha, edge_cache);
  }

  /* Step 3a: complete d_ij on coupling boundary pairs (i < j),
   * hyperbolic_module.template.h:461-490. */
  template <class Eq>
  RJ_DEVICE void boundary_dij_body(const uint32_t vblock,
                                   uint32_t n_pairs,
                                   const uint32_t *__restrict__ pair_i,
                                   const uint32_t *__restrict__ pair_j,
                                   const uint32_t *__restrict__ pair_e,
                                   const uint32_t *__restrict__ pair_eT,
                                   const double *__restrict__ pair_cji,
                                   uint32_t ld,
                                   uint32_t E_owned,
                                   const typename Eq::Params &prm,
                                   const double *__restrict__ U,
                                   const double *__restrict__ prec,
                                   double *__restrict__ dij)
  {
    constexpr int k = Eq::k, dim = Eq::dim, np = Eq::n_prec;
    const uint32_t p = vblock * blockDim.x + threadIdx.x;
    if (p >= n_pairs)
      return;
    const uint32_t i = pair_i[p], j = pair_j[p], e = pair_e[p], eT = pair_eT[p];
    double U_i[k], U_j[k], prec_i[np], prec_j[np];
    load_state<k>(U, ld, i, U_i);
    load_state<k>(U, ld, j, U_j);
    load_state<np>(prec, ld, i, prec_i);
    load_state<np>(prec, ld, j, prec_j);
    double c_ji[dim];
    double nrm2 = 0.;
#pragma unroll
    for (int d = 0; d < dim; ++d) {
      c_ji[d] = pair_cji[size_t(p) * dim + d];
      nrm2 += c_ji[d] * c_ji[d];
    }
    const double norm_ji = sqrt(nrm2);
    const double inv = 1. / norm_ji;
    double n_ji[dim];
#pragma unroll
    for (int d = 0; d < dim; ++d)
      n_ji[d] = c_ji[d] * inv;
    const double d_ij = dij[e];
    const double lambda_max = Eq::riemann_lambda_max(prm, U_j, prec_j, U_i, prec_i, n_ji);
    const double d_ji = norm_ji * lambda_max;
    const double d = fmax(d_ij, d_ji);
    dij[e] = d;
    if (eT < E_owned)
      dij[eT] = d;
  }

  template <class Eq>
  __global__ void k_boundary_dij(uint32_t n_pairs, const uint32_t *__restrict__ pair_i, const uint32_t *__restrict__ pair_j, const uint32_t *__restrict__ pair_e, const uint32_t *__restrict__ pair_eT, const double *__restrict__ pair_cji, uint32_t ld, uint32_t E_owned, typename Eq::Params prm, const double *__restrict__ U, const double *__restrict__ prec, double *__restrict__ dij)
  {
    boundary_dij_body<Eq>(blockIdx.x, n_pairs, pair_i, pair_j, pair_e, pair_eT, pair_cji, ld, E_owned, prm, U, prec, dij);
  }


  /* Step 3b: diagonal d_ii and tau_max, hyperbolic_module.template.h:494-560.
   * The row sum runs in column order in one thread; only the minimum over
   * rows (order independent) uses shuffles + one atomicMin per block. */
  RJ_DEVICE void dii_tau_body(const uint32_t vblock,
                              const StencilView &v,
                              uint32_t row_begin,
                              uint32_t row_end,
                              double cfl,
                              double *__restrict__ dij,
                              StepScalars *__restrict__ scalars)
  {
    RowInfo r;
    double tau_i = DBL_MAX;
    if (row_info(vblock, v, row_begin, row_end, r) && r.len > 1) {
      double d_sum = 0.;
    RJ_UNROLL(RJ_UNROLL_DII)
      for (uint32_t col = 1; col < r.len; ++col)
        d_sum -= dij[r.base + col * 32];
      d_sum = fmin(d_sum, -1.e6 * DBL_MIN);
      dij[r.base] = d_sum;
      tau_i = cfl * v.mi[r.i] / (-2. * d_sum);
    }
    /* NaN must survive the reduction (AssertThrow at :573-576) */
    unsigned long long bits =
        (tau_i != tau_i) ? 0ull : (unsigned long long)__double_as_longlong(fmax(tau_i, 0.));
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, off);
      bits = other < bits ? other : bits;
    }
    __shared__ unsigned long long smin[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0)
      smin[warp] = bits;
    __syncthreads();
    if (warp == 0) {
      const int n_warps = (blockDim.x + 31) >> 5;
      bits = lane < n_warps ? smin[lane] : ~0ull;
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, bits, off);
        bits = other < bits ? other : bits;
      }
      if (lane == 0)
        atomicMin(&scalars->tau_max_bits, bits);
    }
  }

  __global__ void k_dii_tau(StencilView v, uint32_t row_begin, uint32_t row_end, double cfl, double *__restrict__ dij, StepScalars *__restrict__ scalars)
  {
    dii_tau_body(blockIdx.x, v, row_begin, row_end, cfl, dij, scalars);
  }

  __global__ void k_reset_scalars(StepScalars *s)
  {
    s->tau_max_bits = (unsigned long long)__double_as_longlong(DBL_MAX);
    s->tau_max = 0.;
    s->tau = 0.;
    s->restart_needed = 0;
    s->numeric_error = 0;
    s->blocks_done_tau = 0;
    s->blocks_done_step = 0;
  }

  /* Where a substep takes tau and tau_max from: host values (the reference API), or the
   * device scalars of the running cycle. */
  struct TauSource {
    int mode;       /* 0: host values; 1: first substep of a cycle (tau = 0, tau_max = cycle->tau_max /
                       divisor); 2: later substep (tau = cycle->tau, tau_max unbounded) */
    double tau;     /* mode 0 */
    double tau_max; /* mode 0 */
    double divisor; /* mode 1 */
  };

  /* local minimum -> tau_max (before the all-reduce over ranks) */
  __global__ void k_tau_local(StepScalars *s, TauSource src, const CycleScalars *cycle)
  {
    const double local = __longlong_as_double((long long)s->tau_max_bits);
    /* tau_max_bits == 0 encodes NaN or a non-positive value */
    double tau_max_in = src.tau_max;
    if (src.mode == 1)
      tau_max_in = cycle->tau_max == DBL_MAX ? DBL_MAX : __ddiv_rn(cycle->tau_max, src.divisor);
    else if (src.mode == 2)
      tau_max_in = DBL_MAX;
    s->tau_max = fmin(local, tau_max_in);
  }

  /* :571-578; in cycle mode also records the substep's tau */
  __global__ void k_tau_finalize(StepScalars *s, TauSource src, CycleScalars *cycle)
  {
    const double tau_max = s->tau_max;
    if (isnan(tau_max) || isinf(tau_max) || !(tau_max > 0.))
      s->numeric_error = 1;
    const double tau_in = src.mode == 0 ? src.tau : (src.mode == 1 ? 0. : cycle->tau);
    s->tau = (tau_in == 0. ? tau_max : tau_in);
    if (src.mode == 1)
      cycle->tau = s->tau;
  }

  /* end of a substep in cycle mode: keep its flags for the host's post-mortem */
  __global__ void k_cycle_record(const StepScalars *s, CycleScalars *cycle, int substep)
  {
    cycle->restart[substep] = s->restart_needed;
    cycle->numeric_error[substep] = s->numeric_error;
  }

  /* Single-rank device-resident cycles: the scalar kernels of a substep (k_reset_scalars, k_tau_local,
   * k_tau_finalize, k_cycle_record: 4 of 14 launches, each a few microseconds of dependency latency) are
   * folded into the sweeps around them.  The block that finishes LAST (fence + atomic counter) does the
   * scalar work: it sees what every other block wrote before its own fence.  The scalars are left clean
   * for the next substep, which is what k_reset_scalars would have done. */
  RJ_DEVICE bool last_block_done(unsigned int *counter)
  {
    /* one thread per block calls this after the block's results are written (and fenced) */
    __threadfence();
    const bool last = atomicAdd(counter, 1u) == gridDim.x - 1;
    if (last) {
      *counter = 0;
      __threadfence();
    }
    return last;
  }

  /* k_dii_tau, then (last block) k_tau_local + k_tau_finalize; re-arms tau_max_bits */
  __global__ void k_dii_tau_final(StencilView v, uint32_t row_begin, uint32_t row_end, double cfl, double *__restrict__ dij, StepScalars *scalars, TauSource src, CycleScalars *cycle)
  {
    dii_tau_body(blockIdx.x, v, row_begin, row_end, cfl, dij, scalars);
    /* the block's atomicMin was issued by thread 0 (lane 0 of warp 0) */
    if (threadIdx.x == 0 && last_block_done(&scalars->blocks_done_tau)) {
      StepScalars *s = scalars;
      const double local = __longlong_as_double((long long)*(volatile unsigned long long *)&s->tau_max_bits);
      double tau_max_in = src.tau_max;
      if (src.mode == 1)
        tau_max_in = cycle->tau_max == DBL_MAX ? DBL_MAX : __ddiv_rn(cycle->tau_max, src.divisor);
      else if (src.mode == 2)
        tau_max_in = DBL_MAX;
      const double tau_max = fmin(local, tau_max_in);
      s->tau_max = tau_max;
      if (isnan(tau_max) || isinf(tau_max) || !(tau_max > 0.))
        s->numeric_error = 1;
      const double tau_in = src.mode == 0 ? src.tau : (src.mode == 1 ? 0. : cycle->tau);
      s->tau = (tau_in == 0. ? tau_max : tau_in);
      if (src.mode == 1)
        cycle->tau = s->tau;
      s->tau_max_bits = (unsigned long long)__double_as_longlong(DBL_MAX);
    }
  }


  /* ---------------------------------------------------------------------- *
   * Step 4: low-order update, limiter bounds, r_i and P_ij,
   * hyperbolic_module.template.h:597-884.
   * ---------------------------------------------------------------------- */
  /* (Running this sweep as two lighter kernels -- update + bounds, r_i + P_ij -- was measured in
   * round 1 and rejected: 9.6-11.0 ms instead of 7.0 ms, both halves re-read c_ij, d_ij, U_j.) */
  template <class Eq, int stages>
  RJ_DEVICE void low_order_body(const uint32_t vblock,
                                const double tau,
                                const StencilView &v,
                                const typename Eq::Params &prm,
                                uint32_t row_begin,
                                uint32_t row_end,
                                const double *__restrict__ U,
                                const double *__restrict__ prec,
                                const double *__restrict__ aux,
                                const StagePointers<stages> &st,
                                const double *__restrict__ alpha,
                                const double *__restrict__ dij,
                                const double *__restrict__ edge_cache,
                                double *__restrict__ U_new,
                                double *__restrict__ r_vec,
                                double *__restrict__ bounds,
                                double *__restrict__ pij)
  {
    constexpr int k = Eq::k, dim = Eq::dim, np = Eq::n_prec, nb = Eq::n_bounds, na = Eq::n_aux;
    constexpr bool cached = Eq::edge_cache && (RJ_EDGE_CACHE != 0);
    RowInfo r;
    if (!row_info(vblock, v, row_begin, row_end, r) || r.len <= 1)
      return;
    const uint32_t i = r.i;
    

    double weight = -1.;
#pragma unroll
    for (int s = 0; s < stages; ++s)
      weight += st.weight[s];
    weight = -weight;

    double U_i[k], prec_i[np];
    load_state<k>(U, v.ld, i, U_i);
    load_state<np>(prec, v.ld, i, prec_i);
    const double alpha_i = alpha[i];
    const double m_i = v.mi[i];
    const double m_i_inv = v.mi_inv[i];

    double aux_i[na];
    load_state<na>(aux, v.ld, i, aux_i);
    const typename Eq::Flux flux_i = Eq::load_aux(aux_i);

    double U_is[stages > 0 ? stages : 1][k];
    typename Eq::Flux flux_is[stages > 0 ? stages : 1];
    /* nodal sources, :640-672 (only equations with have_source_terms) */
    double S_i[Eq::have_source_terms ? k : 1], S_iH[Eq::have_source_terms ? k : 1];
    if constexpr (Eq::have_source_terms) {
#pragma unroll
      for (int q = 0; q < k; ++q)
        S_iH[q] = 0.;
    }
#pragma unroll
    for (int s = 0; s < stages; ++s) {
      double aux_s[na];
      load_state<k>(st.U[s], v.ld, i, U_is[s]);
      load_state<na>(st.aux[s], v.ld, i, aux_s);
      flux_is[s] = Eq::load_aux(aux_s);
      if constexpr (Eq::have_source_terms) {
        double prec_s[np], S[k];
        load_state<np>(st.prec[s], v.ld, i, prec_s);
        Eq::nodal_source(prm, prec_s, U_is[s], tau, S);
#pragma unroll
        for (int q = 0; q < k; ++q)
          S_iH[q] += st.weight[s] * S[q];
      }
    }

    double U_i_new[k], F_iH[k];
#pragma unroll
    for (int q = 0; q < k; ++q) {
      U_i_new[q] = U_i[q];
      F_iH[q] = 0.;
    }

    if constexpr (Eq::have_source_terms) {
      Eq::nodal_source(prm, prec_i, U_i, tau, S_i);
#pragma unroll
      for (int q = 0; q < k; ++q) {
        S_iH[q] += weight * S_i[q];
        U_i_new[q] += tau * S_i[q];
        F_iH[q] += m_i * S_iH[q];
      }
    }

    typename Eq::LimiterAccumulator limiter;
    if constexpr (Eq::aeos)
      limiter.reset(prec_i);
    else
      limiter.reset();

    const double tau_m_i_inv = tau * m_i_inv;

    /* affine shift of the bar states, :700-720 (shallow water only) */
    double affine_shift[Eq::shallow_water ? k : 1];
    if constexpr (Eq::shallow_water) {
#pragma unroll
      for (int q = 0; q < k; ++q)
        affine_shift[q] = 0.;
      for (uint32_t col = 0; col < r.len; ++col) {
        const uint32_t e = r.base + col * 32;
        const uint32_t j = v.cols[e];
        double aux_j[na], c_ij[dim], B_ij[k];
        load_state<na>(aux, v.ld, j, aux_j);
#pragma unroll
        for (int d = 0; d < dim; ++d)
          c_ij[d] = v.cij[size_t(d) * v.E_owned + e];
        Eq::affine_shift(prm, U_i, flux_i, Eq::load_aux(aux_j), c_ij, dij[e], B_ij);
#pragma unroll
        for (int q = 0; q < k; ++q)
          affine_shift[q] += B_ij[q];
      }
#pragma unroll
      for (int q = 0; q < k; ++q) {
        affine_shift[q] *= tau * m_i_inv;
        affine_shift[q] += tau * S_i[q];
      }
    }

    RJ_INDEX_INIT(RJ_PREFETCH_LOW, j_next, v.cols, r.base)
    RJ_UNROLL(RJ_UNROLL_LOW)
    for (uint32_t col = 0; col < r.len; ++col) {
      const uint32_t e = r.base + col * 32;
      const uint32_t j = RJ_INDEX(RJ_PREFETCH_LOW, j_next, v.cols, e, col, r.len);
      if (RJ_L2_PREFETCH_LOW > 0 && col + RJ_L2_PREFETCH_LOW < r.len) {
        const uint32_t ahead = e + 32 * RJ_L2_PREFETCH_LOW;
        RJ_PREFETCH_L2(&v.cols[ahead]);
#pragma unroll
        for (int d = 0; d < dim; ++d)
          RJ_PREFETCH_L2(&v.cij[size_t(d) * v.E_owned + ahead]);
        RJ_PREFETCH_L2(&dij[ahead]);
        if constexpr (cached)
          RJ_PREFETCH_L2(&edge_cache[ahead]);
      }

      double U_j[k], prec_j[np], c_ij[dim];
      load_state<k>(U, v.ld, j, U_j);
      load_state<np>(prec, v.ld, j, prec_j);
#pragma unroll
      for (int d = 0; d < dim; ++d)
        c_ij[d] = v.cij[size_t(d) * v.E_owned + e];

      const double alpha_j = alpha[j];
      const double d_ij = dij[e];
      const double factor = (alpha_i + alpha_j) * .5;
      const double d_ijH = d_ij * factor;

      const double scale = 1. / fmax(d_ij, 100. * DBL_MIN);
      double scaled_c_ij[dim];
#pragma unroll
      for (int d = 0; d < dim; ++d)
        scaled_c_ij[d] = c_ij[d] * scale;

      double aux_j[na];
      load_state<na>(aux, v.ld, j, aux_j);
      const typename Eq::Flux flux_j = Eq::load_aux(aux_j);

      double flux_ij[k], P_ij[k];
      if constexpr (Eq::shallow_water)
        Eq::flux_divergence(prm, U_i, flux_i, U_j, flux_j, c_ij, flux_ij);
      else
        Eq::flux_divergence(U_i, flux_i, U_j, flux_j, c_ij, flux_ij);
#pragma unroll
      for (int q = 0; q < k; ++q) {
        U_i_new[q] += tau_m_i_inv * flux_ij[q];
        P_ij[q] = -flux_ij[q];
      }

      if constexpr (Eq::shallow_water) {
        /* equilibrated (hydrostatically reconstructed) states, :773-800 */
        double U_star_ij[k], U_star_ji[k];
        Eq::star_state(prm, U_i, flux_i.Z, flux_j.Z, U_star_ij);
        Eq::star_state(prm, U_j, flux_j.Z, flux_i.Z, U_star_ji);
#pragma unroll
        for (int q = 0; q < k; ++q) {
          const double diff = U_star_ji[q] - U_star_ij[q];
          U_i_new[q] += tau_m_i_inv * d_ij * diff;
          F_iH[q] += d_ijH * diff;
          P_ij[q] += (d_ijH - d_ij) * diff;
        }
        limiter.accumulate_sw(prm, U_i, U_j, U_star_ij, U_star_ji, scaled_c_ij, affine_shift);
      } else {
#pragma unroll
        for (int q = 0; q < k; ++q) {
          const double diff = U_j[q] - U_i[q];
          U_i_new[q] += tau_m_i_inv * d_ij * diff;
          F_iH[q] += d_ijH * diff;
          P_ij[q] += (d_ijH - d_ij) * diff;
        }
        if constexpr (Eq::aeos)
          limiter.accumulate(prm, U_i, flux_i, U_j, flux_j, prec_j, scaled_c_ij);
        else if constexpr (cached)
          limiter.accumulate_cached(U_i, U_j, prec_j, scaled_c_ij, edge_cache[e]);
        else
          limiter.accumulate(prm, U_i, U_j, prec_j, scaled_c_ij);
      }

      double m_ij = 0.;
      if constexpr (Eq::have_source_terms) {
        m_ij = v.mij[e];
#pragma unroll
        for (int q = 0; q < k; ++q) {
          F_iH[q] -= m_ij * S_iH[q];
          P_ij[q] -= m_ij * /*sic, :812*/ S_i[q];
        }
      }

      if constexpr (Eq::have_high_order_flux) {
        double ho[k];
        Eq::high_order_flux_divergence(prm, U_i, flux_i, U_j, flux_j, c_ij, ho);
#pragma unroll
        for (int q = 0; q < k; ++q) {
          F_iH[q] += weight * ho[q];
          P_ij[q] += weight * ho[q];
        }
      } else {
#pragma unroll
        for (int q = 0; q < k; ++q) {
          F_iH[q] += weight * flux_ij[q];
          P_ij[q] += weight * flux_ij[q];
        }
      }

      if constexpr (Eq::have_source_terms) {
        double S_j[k];
        Eq::nodal_source(prm, prec_j, U_j, tau, S_j);
#pragma unroll
        for (int q = 0; q < k; ++q) {
          F_iH[q] += weight * m_ij * S_j[q];
          P_ij[q] += weight * m_ij * S_j[q];
        }
      }

#pragma unroll
      for (int s = 0; s < stages; ++s) {
        double U_js[k], aux_js[na], flux_ij_s[k];
        load_state<k>(st.U[s], v.ld, j, U_js);
        load_state<na>(st.aux[s], v.ld, j, aux_js);
        const typename Eq::Flux flux_js = Eq::load_aux(aux_js);
        if constexpr (Eq::have_high_order_flux)
          Eq::high_order_flux_divergence(prm, U_is[s], flux_is[s], U_js, flux_js, c_ij, flux_ij_s);
        else
          Eq::flux_divergence(U_is[s], flux_is[s], U_js, flux_js, c_ij, flux_ij_s);
#pragma unroll
        for (int q = 0; q < k; ++q) {
          F_iH[q] += st.weight[s] * flux_ij_s[q];
          P_ij[q] += st.weight[s] * flux_ij_s[q];
        }
        if constexpr (Eq::have_source_terms) {
          double prec_js[np], S_js[k];
          load_state<np>(st.prec[s], v.ld, j, prec_js);
          Eq::nodal_source(prm, prec_js, U_js, tau, S_js);
#pragma unroll
          for (int q = 0; q < k; ++q) {
            F_iH[q] += st.weight[s] * m_ij * S_js[q];
            P_ij[q] += st.weight[s] * m_ij * S_js[q];
          }
        }
      }

#pragma unroll
      for (int q = 0; q < k; ++q)
        pij[size_t(q) * v.E_owned + e] = P_ij[q];
    }

    store_state<k>(r_vec, v.ld, i, F_iH);
    store_state<k>(U_new, v.ld, i, U_i_new);
    const double hd_i = m_i * v.measure_of_omega_inverse;
    double b[nb];
    limiter.bounds(prm, hd_i, b);
    store_state<nb>(bounds, v.ld, i, b);
  }

  template <class Eq, int stages>
  __global__ void __launch_bounds__(128, Eq::shallow_water ? RJ_MB_LOW_SW : RJ_MB_LOW) k_low_order(StencilView v, typename Eq::Params prm, uint32_t row_begin, uint32_t row_end, const double *__restrict__ U, const double *__restrict__ prec, const double *__restrict__ aux, StagePointers<stages> st, const double *__restrict__ alpha, const double *__restrict__ dij, const double *__restrict__ edge_cache, const StepScalars *__restrict__ scalars, double *__restrict__ U_new, double *__restrict__ r_vec, double *__restrict__ bounds, double *__restrict__ pij)
  {
    low_order_body<Eq, stages>(blockIdx.x, scalars->tau, v, prm, row_begin, row_end, U, prec, aux, st, alpha, dij, edge_cache, U_new, r_vec, bounds, pij);
  }


  /* ---------------------------------------------------------------------- *
   * Step 5: mass-matrix correction of P_ij and first l_ij,
   * hyperbolic_module.template.h:892-1041.
   * ---------------------------------------------------------------------- */
  template <class Eq>
  RJ_DEVICE void pij_lij_body(const uint32_t vblock,
                              const double tau,
                              const StencilView &v,
                              const typename Eq::Params &prm,
                              uint32_t row_begin,
                              uint32_t row_end,
                              const double *__restrict__ U_new,
                              const double *__restrict__ r_vec,
                              const double *__restrict__ bounds,
                              StepScalars *__restrict__ scalars,
                              double *__restrict__ pij,
                              double *__restrict__ lij)
  {
    constexpr int k = Eq::k, nb = Eq::n_bounds;
    RowInfo r;
    if (!row_info(vblock, v, row_begin, row_end, r) || r.len <= 1)
      return;
    const uint32_t i = r.i;
    

    double b[nb], U_i_new[k], F_iH[k];
    load_state<nb>(bounds, v.ld, i, b);
    load_state<k>(U_new, v.ld, i, U_i_new);
    load_state<k>(r_vec, v.ld, i, F_iH);
    const double m_i_inv = v.mi_inv[i];

    const double lambda_inv = double(r.len - 1);
    const double factor = tau * m_i_inv * lambda_inv;

    bool all_ok = true;
    RJ_INDEX_INIT(RJ_PREFETCH_PIJ, j_next, v.cols, r.base + 32) /* r.len > 1 */
    RJ_UNROLL(RJ_UNROLL_PIJ)
    for (uint32_t col = 1; col < r.len; ++col) {
      const uint32_t e = r.base + col * 32;
      const uint32_t j = RJ_INDEX(RJ_PREFETCH_PIJ, j_next, v.cols, e, col, r.len);
      double P_ij[k], F_jH[k];
#pragma unroll
      for (int q = 0; q < k; ++q)
        P_ij[q] = pij[size_t(q) * v.E_owned + e];
      load_state<k>(r_vec, v.ld, j, F_jH);

      const double m_j_inv = v.mi_inv[j];
      const double m_ij = v.mij[e];
      const double b_ij = 0. - m_ij * m_j_inv;
      const double b_ji = 0. - m_ij * m_i_inv;

#pragma unroll
      for (int q = 0; q < k; ++q) {
        P_ij[q] += b_ij * F_jH[q] - b_ji * F_iH[q];
        P_ij[q] *= factor;
        pij[size_t(q) * v.E_owned + e] = P_ij[q];
      }

      bool success;
      const double l_ij = Eq::limit(prm, b, U_i_new, P_ij, success);
      lij[e] = l_ij;
      all_ok = all_ok && success;
    }
    if (!all_ok)
      atomicOr(&scalars->restart_needed, 1);
  }

  template <class Eq>
  __global__ void __launch_bounds__(128, RJ_MB_PIJ) k_pij_lij(StencilView v, typename Eq::Params prm, uint32_t row_begin, uint32_t row_end, const double *__restrict__ U_new, const double *__restrict__ r_vec, const double *__restrict__ bounds, StepScalars *__restrict__ scalars, double *__restrict__ pij, double *__restrict__ lij)
  {
    pij_lij_body<Eq>(blockIdx.x, scalars->tau, v, prm, row_begin, row_end, U_new, r_vec, bounds, scalars, pij, lij);
  }


  /* ---------------------------------------------------------------------- *
   * Steps 6, 7: symmetrize l_ij, high-order update, next l_ij,
   * hyperbolic_module.template.h:1053-1182.
   * ---------------------------------------------------------------------- */
  template <class Eq, bool last_round>
  RJ_DEVICE void high_order_body(const uint32_t vblock,
                                 const StencilView &v,
                                 const typename Eq::Params &prm,
                                 uint32_t row_begin,
                                 uint32_t row_end,
                                 const double *__restrict__ bounds,
                                 const double *__restrict__ pij,
                                 const double *__restrict__ lij,
                                 double *__restrict__ lij_next,
                                 double *__restrict__ U_new)
  {
    constexpr int k = Eq::k, nb = Eq::n_bounds;
    RowInfo r;
    if (!row_info(vblock, v, row_begin, row_end, r) || r.len <= 1)
      return;
    const uint32_t i = r.i;

    double U_i_new[k];
    load_state<k>(U_new, v.ld, i, U_i_new);
    const double lambda = 1. / double(r.len - 1);

    RJ_INDEX_INIT(RJ_PREFETCH_HO, eT_next, v.transposed, r.base + 32) /* r.len > 1 */
    RJ_UNROLL(RJ_UNROLL_HO1)
    for (uint32_t col = 1; col < r.len; ++col) {
      const uint32_t e = r.base + col * 32;
      const double l_ij = fmin(lij[e], lij[RJ_INDEX(RJ_PREFETCH_HO, eT_next, v.transposed, e, col, r.len)]);
      const double ll = l_ij * lambda;
#pragma unroll
      for (int q = 0; q < k; ++q)
        U_i_new[q] += ll * pij[size_t(q) * v.E_owned + e];
    }
    store_state<k>(U_new, v.ld, i, U_i_new);

    if (last_round)
      return;

    double b[nb];
    load_state<nb>(bounds, v.ld, i, b);
    /* second sweep over the row's P_ij and l_ij in REVERSE column order: the entries read
     * last by the first sweep are re-read first, while they are still in L2 (the per-entry
     * results are independent, so the order is free) */
    RJ_UNROLL(RJ_UNROLL_HO2)
    for (uint32_t col = RJ_HO2_FIRST(r.len); RJ_HO2_COND(col, r.len); RJ_HO2_NEXT(col)) {
      const uint32_t e = r.base + col * 32;
      const double old_l_ij = fmin(lij[e], lij[v.transposed[e]]);
      double new_p_ij[k];
#pragma unroll
      for (int q = 0; q < k; ++q)
        new_p_ij[q] = (1. - old_l_ij) * pij[size_t(q) * v.E_owned + e];
      bool success;
      const double new_l_ij = Eq::limit(prm, b, U_i_new, new_p_ij, success);
      lij_next[e] = (1. - old_l_ij) * new_l_ij;
    }
  }

  template <class Eq, bool last_round>
  __global__ void __launch_bounds__(128, RJ_MB_HO) k_high_order(StencilView v, typename Eq::Params prm, uint32_t row_begin, uint32_t row_end, const double *__restrict__ bounds, const double *__restrict__ pij, const double *__restrict__ lij, double *__restrict__ lij_next, double *__restrict__ U_new)
  {
    high_order_body<Eq, last_round>(blockIdx.x, v, prm, row_begin, row_end, bounds, pij, lij, lij_next, U_new);
  }

  /* the final high-order update, then (last block) k_cycle_record; clears the substep's flags */
  template <class Eq>
  __global__ void __launch_bounds__(128, RJ_MB_HO) k_high_order_record(StencilView v, typename Eq::Params prm, uint32_t row_begin, uint32_t row_end, const double *__restrict__ bounds, const double *__restrict__ pij, const double *__restrict__ lij, double *__restrict__ lij_next, double *__restrict__ U_new, StepScalars *scalars, CycleScalars *cycle, int substep)
  {
    high_order_body<Eq, true>(blockIdx.x, v, prm, row_begin, row_end, bounds, pij, lij, lij_next, U_new);
    __syncthreads();
    if (threadIdx.x == 0 && last_block_done(&scalars->blocks_done_step)) {
      cycle->restart[substep] = *(volatile int *)&scalars->restart_needed;
      cycle->numeric_error[substep] = *(volatile int *)&scalars->numeric_error;
      scalars->restart_needed = 0;
      scalars->numeric_error = 0;
    }
  }


  /* ---------------------------------------------------------------------- *
   * Postprocessor (source/postprocessor.template.h:113-281): schlieren |sum_j c_ij q_j| / m_i of a
   * scalar quantity and the vorticity (curl) of a vector quantity, on the exponential scale
   * 1 - exp(-beta ratio).  Row sums in column order like every other per-row accumulation.
   * ---------------------------------------------------------------------- */
  struct PostprocessorScalars {
    unsigned long long max_bits, min_bits; /* ordered bit patterns of |q| >= 0 */
    double q_max, q_min;
  };

  __global__ void k_postprocess_reset(PostprocessorScalars *s)
  {
    s->max_bits = 0ull;
    s->min_bits = (unsigned long long)__double_as_longlong(DBL_MAX);
    s->q_max = 0.;
    s->q_min = DBL_MAX;
  }

  /* mode 0: schlieren of component `first`; mode 1: vorticity of components first .. first+dim-1 */
  template <int dim>
  __global__ void k_postprocess_quantity(StencilView v,
                                         uint32_t row_begin,
                                         uint32_t row_end,
                                         int mode,
                                         uint32_t first,
                                         const double *__restrict__ U,
                                         double *__restrict__ quantity,
                                         PostprocessorScalars *__restrict__ scalars)
  {
    RowInfo r;
    double value = 0.;
    bool active = false;
    if (row_info(blockIdx.x, v, row_begin, row_end, r) && r.len > 1) {
      active = true;
      double acc[3] = {0., 0., 0.};
      for (uint32_t col = 0; col < r.len; ++col) {
        const uint32_t e = r.base + col * 32;
        const uint32_t j = v.cols[e];
        double c_ij[dim];
#pragma unroll
        for (int d = 0; d < dim; ++d)
          c_ij[d] = v.cij[size_t(d) * v.E_owned + e];
        if (mode == 0) {
          const double q_j = U[size_t(first) * v.ld + j];
#pragma unroll
          for (int d = 0; d < dim; ++d)
            acc[d] -= c_ij[d] * q_j;
        } else {
          double q_j[dim];
#pragma unroll
          for (int d = 0; d < dim; ++d)
            q_j[d] = U[size_t(first + d) * v.ld + j];
          if (dim == 2) {
            /* cross_product_2d(c_ij) * q_j = (c_y, -c_x) . q */
            acc[0] -= c_ij[dim - 1] * q_j[0] - c_ij[0] * q_j[dim - 1];
          } else if (dim == 3) {
            acc[0] -= c_ij[1 % dim] * q_j[2 % dim] - c_ij[2 % dim] * q_j[1 % dim];
            acc[1] -= c_ij[2 % dim] * q_j[0] - c_ij[0] * q_j[2 % dim];
            acc[2] -= c_ij[0] * q_j[1 % dim] - c_ij[1 % dim] * q_j[0];
          }
        }
      }
      const double m_i = v.mi[r.i];
      if (mode == 1 && dim == 2)
        value = acc[0] / m_i;
      else {
        double n2 = acc[0] * acc[0];
        const int n_acc = mode == 0 ? dim : 3;
        for (int d = 1; d < n_acc; ++d)
          n2 += acc[d] * acc[d];
        value = sqrt(n2) / m_i;
      }
      quantity[r.i] = value;
    } else if (r.i < row_end)
      quantity[r.i] = 0.;
    /* bounds over all owned rows (constrained rows hold 0) */
    const unsigned long long bits = (unsigned long long)__double_as_longlong(fabs(active ? value : 0.));
    if (r.i < row_end) {
      atomicMax(&scalars->max_bits, bits);
      atomicMin(&scalars->min_bits, bits);
    }
  }

  __global__ void k_postprocess_bounds(PostprocessorScalars *s)
  {
    s->q_max = __longlong_as_double((long long)s->max_bits);
    s->q_min = __longlong_as_double((long long)s->min_bits);
  }

  __global__ void k_postprocess_normalize(uint32_t n, double beta, const PostprocessorScalars *__restrict__ s, double *__restrict__ quantity)
  {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
      return;
    const double floor = fmax(1.0e-10, DBL_EPSILON);
    const double q = quantity[i];
    const double ratio = fmax(0., fabs(q) - s->q_min - floor) / fmax(s->q_max - s->q_min, DBL_EPSILON);
    const double magnitude = 1. - exp(-beta * ratio);
    quantity[i] = copysign(magnitude, q);
  }


  /* ---------------------------------------------------------------------- *
   * Plumbing kernels
   * ---------------------------------------------------------------------- */

  /* host AoS (i*k + c) <-> device component-major (c*ld + i) */
  __global__ void k_aos_to_soa(uint32_t n, uint32_t k, uint32_t ld, const double *__restrict__ aos, double *__restrict__ soa)
  {
    const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= size_t(n) * k)
      return;
    const uint32_t i = uint32_t(t / k), c = uint32_t(t % k);
    soa[size_t(c) * ld + i] = aos[t];
  }

  __global__ void k_soa_to_aos(uint32_t n, uint32_t k, uint32_t ld, const double *__restrict__ soa, double *__restrict__ aos)
  {
    const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= size_t(n) * k)
      return;
    const uint32_t i = uint32_t(t / k), c = uint32_t(t % k);
    aos[t] = soa[size_t(c) * ld + i];
  }

  /* dst = s*dst + b*src, time_integrator.template.h:18-25 */
  __global__ void k_sadd(size_t n, double s, double b, double *__restrict__ dst, const double *__restrict__ src)
  {
    const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t < n)
      dst[t] = s * dst[t] + b * src[t];
  }

  /* gather rows of a component-major vector into a contiguous send buffer
   * laid out [component][row] per neighbour block */
  __global__ void k_pack_rows(uint32_t n_rows, uint32_t n_comp, uint32_t ld, const uint32_t *__restrict__ rows, const double *__restrict__ v, double *__restrict__ buf)
  {
    const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t >= size_t(n_rows) * n_comp)
      return;
    const uint32_t c = uint32_t(t / n_rows), s = uint32_t(t % n_rows);
    buf[t] = v[size_t(c) * ld + rows[s]];
  }

  /* All neighbours of one vector exchange in ONE launch (instead of one launch, n_comp sends and
   * n_comp receives per neighbour): the send buffer holds one contiguous block per neighbour,
   * [component][row] inside the block, the receive buffer likewise with the neighbour's ghost
   * rows; one ncclSend / ncclRecv per neighbour moves a whole block. */
  constexpr int max_neighbors = 26;
  struct ExchangePlan {
    uint32_t n_neighbors;
    uint32_t row_offset[max_neighbors + 1]; /* prefix sums of the rows per neighbour */
    uint32_t ghost_start[max_neighbors];    /* unpack only: first ghost row of the neighbour */
  };

  RJ_DEVICE uint32_t plan_neighbor(const ExchangePlan &plan, uint32_t row)
  {
    uint32_t p = 0;
    while (p + 1 < plan.n_neighbors && row >= plan.row_offset[p + 1])
      ++p;
    return p;
  }

  /* t enumerates (neighbour block, component, row) in buffer order */
  __global__ void k_pack_blocks(ExchangePlan plan, uint32_t n_comp, uint32_t ld, const uint32_t *__restrict__ rows, const double *__restrict__ v, double *__restrict__ buf)
  {
    const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint32_t n_rows = plan.row_offset[plan.n_neighbors];
    if (t >= size_t(n_rows) * n_comp)
      return;
    /* block p starts at row_offset[p] * n_comp */
    uint32_t p = 0;
    while (p + 1 < plan.n_neighbors && t >= size_t(plan.row_offset[p + 1]) * n_comp)
      ++p;
    const uint32_t count = plan.row_offset[p + 1] - plan.row_offset[p];
    const size_t local = t - size_t(plan.row_offset[p]) * n_comp;
    const uint32_t c = uint32_t(local / count), s = uint32_t(local % count);
    buf[t] = v[size_t(c) * ld + rows[plan.row_offset[p] + s]];
  }

  __global__ void k_unpack_blocks(ExchangePlan plan, uint32_t n_comp, uint32_t ld, const double *__restrict__ buf, double *__restrict__ v)
  {
    const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    const uint32_t n_rows = plan.row_offset[plan.n_neighbors];
    if (t >= size_t(n_rows) * n_comp)
      return;
    uint32_t p = 0;
    while (p + 1 < plan.n_neighbors && t >= size_t(plan.row_offset[p + 1]) * n_comp)
      ++p;
    const uint32_t count = plan.row_offset[p + 1] - plan.row_offset[p];
    const size_t local = t - size_t(plan.row_offset[p]) * n_comp;
    const uint32_t c = uint32_t(local / count), s = uint32_t(local % count);
    v[size_t(c) * ld + plan.ghost_start[p] + s] = buf[t];
  }

  __global__ void k_pack_entries(uint32_t n, const uint32_t *__restrict__ entries, const double *__restrict__ m, double *__restrict__ buf)
  {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < n)
      buf[t] = m[entries[t]];
  }

  /* CSR-ordered copy of a sliced matrix (introspection) */
  __global__ void k_gather_entries(size_t n, const uint32_t *__restrict__ entries, const double *__restrict__ m, double *__restrict__ out)
  {
    const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
    if (t < n)
      out[t] = m[entries[t]];
  }
} // namespace rjb
