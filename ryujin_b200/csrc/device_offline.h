// Device resident stencil data in the warp-sliced ELL layout (C = 32).
//
// This is the GPU form of the reference's SparsityPatternSIMD / SparseMatrixSIMD
// (source/sparse_matrix_simd.h:40-61): there, rows are grouped in slices of
// simd_length rows and a slice is stored column by column so that one SIMD load
// fetches the same column of all rows of the slice.  Here simd_length is the
// warp size: lane l of warp g owns row 32 g + l, entry (row, col) lives at
//     e = slice_offsets[g] + col * 32 + l
// and every matrix array is stored component-major over e, so a warp-wide load
// of one component of one column is a single coalesced 256 byte line.
// Slices are padded to the longest row of the slice (row_length[] has the true
// lengths); the reference instead demands equal row lengths inside a SIMD
// group and keeps a scalar CSR tail (offline_data.template.h:255-272).
//
// Vectors are component-major as well ("SoA"): (U_i)_c at U[c*ld + i], ghost
// rows [n_owned, n_relevant) in the tail exactly like the ghost range of
// dealii::LinearAlgebra::distributed::Vector.
#pragma once

#include "device_common.cuh"

#include <vector>

struct rj_offline {
  int device = 0;
  int dim = 0;
  uint32_t n_export = 0, n_owned = 0, n_relevant = 0;
  uint32_t ld = 0;       /* vector stride, multiple of 128 */
  uint32_t n_slices = 0; /* ceil(n_owned / 32) */
  uint32_t E_owned = 0;  /* padded entries of owned rows */
  uint32_t E_total = 0;  /* + compressed ghost rows (transposed targets only) */
  double measure_of_omega = 0.;

  /* sliced ELL pattern */
  uint32_t *slice_offsets = nullptr; /* n_slices + 1 */
  uint8_t *row_length = nullptr;     /* n_slices * 32 (0 beyond n_owned) */
  uint32_t *cols = nullptr;          /* E_owned */
  uint32_t *transposed = nullptr;    /* E_owned: entry index of (j, i) */
  double *mij = nullptr;             /* E_owned */
  double *cij = nullptr;             /* dim * E_owned, component-major */
  double *mi = nullptr, *mi_inv = nullptr; /* ld */

  /* boundary map grouped by dof */
  uint32_t n_bgroups = 0, n_bentries = 0;
  uint32_t *bgroup_dof = nullptr, *bgroup_start = nullptr; /* n_bgroups (+1) */
  double *bnormal = nullptr, *bposition = nullptr;         /* n_bentries * dim */
  int32_t *bid = nullptr;

  /* coupling boundary pairs with i < j */
  uint32_t n_pairs = 0;
  uint32_t *pair_i = nullptr, *pair_j = nullptr, *pair_e = nullptr, *pair_eT = nullptr;
  double *pair_cji = nullptr; /* n_pairs * dim */

  /* halo exchange */
  struct Neighbor {
    int rank;
    uint32_t ghost_start, ghost_count;        /* rows */
    uint32_t send_offset, send_count;         /* into send_rows */
    uint32_t mat_send_offset, mat_send_count; /* into mat_send_entries */
    uint32_t mat_recv_offset, mat_recv_count; /* entry range in [E_owned, E_total) */
  };
  std::vector<Neighbor> neighbors;
  uint32_t n_send_rows = 0, n_mat_send = 0;
  uint32_t *send_rows = nullptr;
  uint32_t *mat_send_entries = nullptr;

  double *positions = nullptr; /* dim * n_owned, component-major; may be null */

  /* host mirrors for introspection / CSR <-> sliced index maps */
  std::vector<uint32_t> h_slice_offsets;
  std::vector<uint32_t> h_row_length;
  std::vector<int32_t> h_boundary_ids; /* host copy: equations reject ids they do not implement */
  std::vector<uint32_t> h_cols, h_transposed; /* host-only layouts */
  std::vector<uint32_t> h_send_rows, h_mat_send_entries;
  bool host_only = false;
  /* all rows with ghost columns and all send rows lie in [0, n_export): precondition of the
   * export-rows-first overlap of the module (violated by e.g. an imported stencil that does not
   * declare its export range: the module then exchanges after the full sweep) */
  bool export_rows_first = true;

  uint32_t entry(uint32_t row, uint32_t col) const
  {
    return h_slice_offsets[row / 32] + col * 32 + row % 32;
  }

  ~rj_offline();
};

namespace rjb
{
  rj_offline *build_device_offline(const rj_offline_desc &desc, int device, bool host_only = false);
}
