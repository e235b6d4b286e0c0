// Slice-tiled variants of the limiter sweeps: a block of W warps owns ONE slice of 32 rows, the
// slice's matrix tiles are staged in shared memory by 1D bulk copies (tile_copy.cuh: the TMA
// engine, cp.async.bulk + mbarrier) and the W warps share the slice's columns (warp w takes
// columns w, w + W, ...), so that W columns of every row are in flight at once.  Per-row sums
// that the reference accumulates in column order are formed by ONE warp in column order from the
// staged tile -- the same additions, bit for bit.
//
//   k_pij_lij_tile     step 5, hyperbolic_module.template.h:892-1041: P_ij tile in by bulk load,
//                      corrected in place in shared memory, out by bulk store; no reduction.
//   k_high_order_tile  step 6, :1053-1182: P_ij and l_ij tiles read from HBM ONCE (the
//                      row-parallel kernel reads P_ij twice: 2937 instead of 1724 B per DoF).
#pragma once

#include "kernels.cuh"
#include "tile_copy.cuh"

#if defined(RJ_HOST_EMULATION)
unsigned char *rj_emu_dynamic_smem(); /* tests/host_emulation/host_emulation.cc */
#define RJ_DYNAMIC_SMEM(name) unsigned char *name = rj_emu_dynamic_smem()
#else
#define RJ_DYNAMIC_SMEM(name) extern __shared__ __align__(128) unsigned char name[]
#endif

#ifndef RJ_TILE_DEFAULT_MASK
#define RJ_TILE_DEFAULT_MASK 3 /* bit 0: step 5 tiled, bit 1: step 6 tiled */
#endif
#ifndef RJ_OVERLAP_MODE_DEFAULT
#define RJ_OVERLAP_MODE_DEFAULT 1
#endif
#ifndef RJ_TILE_W
#define RJ_TILE_W 4 /* warps per slice */
#endif
#ifndef RJ_MB_PIJ_TILE
#define RJ_MB_PIJ_TILE 5
#endif
#ifndef RJ_MB_HO_TILE
#define RJ_MB_HO_TILE 5
#endif

namespace rjb
{
  constexpr uint32_t tile_header_bytes = 128; /* mbarrier + padding: tiles stay 128-byte aligned */

  /* bytes of dynamic shared memory of the two kernels for a stencil whose widest slice has
   * `width` columns */
  template <class Eq>
  constexpr size_t pij_tile_smem(uint32_t width)
  {
    return tile_header_bytes + size_t(width) * 32 * (size_t(Eq::k + 1) * 8 + 4);
  }
  template <class Eq>
  constexpr size_t high_order_tile_smem(uint32_t width)
  {
    return tile_header_bytes + size_t(width) * 32 * size_t(Eq::k + 1) * 8 + size_t(Eq::k) * 32 * 8;
  }

  struct SliceInfo {
    uint32_t g, lane, w, i, len, s0, width, base;
    bool has_row;
  };

  RJ_DEVICE SliceInfo slice_info(const StencilView &v, uint32_t g)
  {
    SliceInfo s;
    s.g = g;
    s.lane = threadIdx.x & 31u;
    s.w = threadIdx.x >> 5;
    s.i = (g << 5) + s.lane;
    s.s0 = v.slice_offsets[g];
    s.width = (v.slice_offsets[g + 1] - s.s0) >> 5;
    s.base = s.s0 + s.lane;
    s.len = s.i < v.n_owned ? v.row_length[s.i] : 0u;
    s.has_row = s.len > 1;
    if (!s.has_row)
      s.len = 0;
    return s;
  }


  /* ---------------------------------------------------------------------- *
   * Step 5 on a slice tile.
   * ---------------------------------------------------------------------- */
  template <class Eq, int W>
  __global__ void __launch_bounds__(32 * W, RJ_MB_PIJ_TILE) k_pij_lij_tile(StencilView v,
                                                                         typename Eq::Params prm,
                                                                         uint32_t slice_begin,
                                                                         uint32_t slice_end,
                                                                         uint32_t max_width,
                                                                         const double *__restrict__ U_new,
                                                                         const double *__restrict__ r_vec,
                                                                         const double *__restrict__ bounds,
                                                                         StepScalars *__restrict__ scalars,
                                                                         double *__restrict__ pij,
                                                                         double *__restrict__ lij)
  {
    constexpr int k = Eq::k, nb = Eq::n_bounds;
    RJ_DYNAMIC_SMEM(smem);
    const uint32_t g = slice_begin + blockIdx.x;
    if (g >= slice_end)
      return; /* the whole block */
    tile::Barrier *bar = reinterpret_cast<tile::Barrier *>(smem);
    const uint32_t tile_entries = max_width * 32;
    double *P = reinterpret_cast<double *>(smem + tile_header_bytes); /* [k][width][32] */
    double *M = P + size_t(k) * tile_entries;                         /* m_ij */
    uint32_t *C = reinterpret_cast<uint32_t *>(M + tile_entries);     /* column indices */

    const SliceInfo s = slice_info(v, g);
    const uint32_t bytes = s.width * 256;

    if (threadIdx.x == 0) {
      tile::init(bar, 1);
      tile::fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      tile::arrive_expect(bar, (k + 1) * bytes + bytes / 2);
#pragma unroll
      for (int q = 0; q < k; ++q)
        tile::load(P + size_t(q) * tile_entries, pij + size_t(q) * v.E_owned + s.s0, bytes, bar);
      tile::load(M, v.mij + s.s0, bytes, bar);
      tile::load(C, v.cols + s.s0, bytes / 2, bar);
    }

    /* per-row data while the tiles are in flight */
    double b[nb], U_i_new[k], F_iH[k];
    double m_i_inv = 0., factor = 0.;
    if (s.has_row) {
      const double tau = scalars->tau;
      load_state<nb>(bounds, v.ld, s.i, b);
      load_state<k>(U_new, v.ld, s.i, U_i_new);
      load_state<k>(r_vec, v.ld, s.i, F_iH);
      m_i_inv = v.mi_inv[s.i];
      const double lambda_inv = double(s.len - 1);
      factor = tau * m_i_inv * lambda_inv;
    }

    tile::wait(bar, 0);

    bool all_ok = true;
    for (uint32_t col = 1 + s.w; col < s.len; col += W) {
      const uint32_t t = col * 32 + s.lane;
      const uint32_t j = C[t];
      double P_ij[k], F_jH[k];
#pragma unroll
      for (int q = 0; q < k; ++q)
        P_ij[q] = P[size_t(q) * tile_entries + t];
      load_state<k>(r_vec, v.ld, j, F_jH);

      const double m_j_inv = v.mi_inv[j];
      const double m_ij = M[t];
      const double b_ij = 0. - m_ij * m_j_inv;
      const double b_ji = 0. - m_ij * m_i_inv;

#pragma unroll
      for (int q = 0; q < k; ++q) {
        P_ij[q] += b_ij * F_jH[q] - b_ji * F_iH[q];
        P_ij[q] *= factor;
        P[size_t(q) * tile_entries + t] = P_ij[q];
      }

      bool success;
      const double l_ij = Eq::limit(prm, b, U_i_new, P_ij, success);
      lij[s.base + col * 32] = l_ij;
      all_ok = all_ok && success;
    }
    if (!all_ok)
      atomicOr(&scalars->restart_needed, 1);

    /* corrected tile back to global memory: one bulk store per component */
    tile::fence_async();
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int q = 0; q < k; ++q)
        tile::store(pij + size_t(q) * v.E_owned + s.s0, P + size_t(q) * tile_entries, bytes);
      tile::store_commit();
      tile::store_wait_read();
    }
  }


  /* ---------------------------------------------------------------------- *
   * Step 6 on a slice tile (not the last limiter round).
   * ---------------------------------------------------------------------- */
  template <class Eq, int W>
  __global__ void __launch_bounds__(32 * W, RJ_MB_HO_TILE) k_high_order_tile(StencilView v,
                                                                           typename Eq::Params prm,
                                                                           uint32_t slice_begin,
                                                                           uint32_t slice_end,
                                                                           uint32_t max_width,
                                                                           const double *__restrict__ bounds,
                                                                           const double *__restrict__ pij,
                                                                           const double *__restrict__ lij,
                                                                           double *__restrict__ lij_next,
                                                                           double *__restrict__ U_new)
  {
    constexpr int k = Eq::k, nb = Eq::n_bounds;
    RJ_DYNAMIC_SMEM(smem);
    const uint32_t g = slice_begin + blockIdx.x;
    if (g >= slice_end)
      return; /* the whole block */
    tile::Barrier *bar = reinterpret_cast<tile::Barrier *>(smem);
    const uint32_t tile_entries = max_width * 32;
    double *P = reinterpret_cast<double *>(smem + tile_header_bytes); /* [k][width][32] */
    double *L = P + size_t(k) * tile_entries;                         /* l_ij, then min(l_ij, l_ji) */
    double *updated = L + tile_entries;                               /* U_i^{n+1} of the slice's rows */

    const SliceInfo s = slice_info(v, g);
    const uint32_t bytes = s.width * 256;

    if (threadIdx.x == 0) {
      tile::init(bar, 1);
      tile::fence_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      tile::arrive_expect(bar, (k + 1) * bytes);
#pragma unroll
      for (int q = 0; q < k; ++q)
        tile::load(P + size_t(q) * tile_entries, pij + size_t(q) * v.E_owned + s.s0, bytes, bar);
      tile::load(L, lij + s.s0, bytes, bar);
    }

    /* the transposed entries l_ji are scattered: the threads gather them while the tiles are in
     * flight (fully unrolled: the values stay in registers); rows longer than the unrolled part
     * gather the rest after the wait */
    constexpr int n_prefetch = (RJ_TILE_PREFETCH_MAXCOL - 1 + W - 1) / W;
    double l_ji[n_prefetch];
#pragma unroll
    for (int n = 0; n < n_prefetch; ++n) {
      const uint32_t col = 1 + s.w + uint32_t(n) * W;
      l_ji[n] = col < s.len ? lij[v.transposed[s.base + col * 32]] : 0.;
    }

    tile::wait(bar, 0);
#pragma unroll
    for (int n = 0; n < n_prefetch; ++n) {
      const uint32_t col = 1 + s.w + uint32_t(n) * W;
      if (col < s.len) {
        const uint32_t t = col * 32 + s.lane;
        L[t] = fmin(L[t], l_ji[n]);
      }
    }
    for (uint32_t col = 1 + s.w + uint32_t(n_prefetch) * W; col < s.len; col += W) {
      const uint32_t t = col * 32 + s.lane;
      L[t] = fmin(L[t], lij[v.transposed[s.base + col * 32]]);
    }
    __syncthreads();

    if (s.w == 0 && s.has_row) {
      double U_i_new[k];
      load_state<k>(U_new, v.ld, s.i, U_i_new);
      const double lambda = 1. / double(s.len - 1);
      for (uint32_t col = 1; col < s.len; ++col) {
        const uint32_t t = col * 32 + s.lane;
        const double ll = L[t] * lambda;
#pragma unroll
        for (int q = 0; q < k; ++q)
          U_i_new[q] += ll * P[size_t(q) * tile_entries + t];
      }
      store_state<k>(U_new, v.ld, s.i, U_i_new);
#pragma unroll
      for (int q = 0; q < k; ++q)
        updated[q * 32 + s.lane] = U_i_new[q];
    }
    __syncthreads();

    if (!s.has_row)
      return;
    double U_i_new[k], b[nb];
#pragma unroll
    for (int q = 0; q < k; ++q)
      U_i_new[q] = updated[q * 32 + s.lane];
    load_state<nb>(bounds, v.ld, s.i, b);
    for (uint32_t col = 1 + s.w; col < s.len; col += W) {
      const uint32_t t = col * 32 + s.lane;
      const double old_l_ij = L[t];
      double new_p_ij[k];
#pragma unroll
      for (int q = 0; q < k; ++q)
        new_p_ij[q] = (1. - old_l_ij) * P[size_t(q) * tile_entries + t];
      bool success;
      const double new_l_ij = Eq::limit(prm, b, U_i_new, new_p_ij, success);
      lij_next[s.base + col * 32] = (1. - old_l_ij) * new_l_ij;
    }
  }
} // namespace rjb
