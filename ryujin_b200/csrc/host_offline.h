// Host side stencil data ("OfflineData") and the Cartesian Q1 mesh generator.
//
// Replaces, for uniform (optionally masked) Cartesian meshes, the reference's
// OfflineData::prepare() (source/offline_data.template.h:58-1105):
//   - DoF numbering with export rows first   (offline_data.template.h:210-249,
//                                             local_index_handling.h:101-371)
//   - m_ij, c_ij = int phi_i grad phi_j       (offline_data.template.h:560-578)
//   - lumped mass, |Omega|                    (offline_data.template.h:756-806)
//   - boundary map                            (offline_data.template.h:1173-1364)
//   - coupling boundary pairs                 (offline_data.template.h:1369-1462)
//   - ghost ranges / export lists             (multicomponent_vector.cc:11-41)
// deal.II is not used: Q1 cell integrals on affine boxes are closed form.
#pragma once

#include "../../include/ryujin_b200.h"

#include <cstdint>
#include <string>
#include <vector>

namespace rjb
{
  struct HostOffline {
    int dim = 0;
    uint32_t n_export = 0, n_internal = 0, n_owned = 0, n_relevant = 0;

    /* CSR over owned rows, diagonal first then ascending local column */
    std::vector<uint64_t> row_starts;
    std::vector<uint32_t> cols;
    std::vector<double> mij;
    std::vector<double> cij; /* nnz * dim, entry major */
    std::vector<double> mi;  /* n_relevant */
    double measure_of_omega = 0.;

    std::vector<uint32_t> boundary_dof;
    std::vector<double> boundary_normal;
    std::vector<double> boundary_normal_mass, boundary_mass;
    std::vector<int32_t> boundary_id;
    std::vector<double> boundary_position;

    std::vector<uint32_t> coupling_pairs;
    std::vector<double> coupling_cji;

    std::vector<int32_t> neighbor_rank;
    std::vector<uint32_t> ghost_count;
    std::vector<uint32_t> send_offsets;
    std::vector<uint32_t> send_indices;

    std::vector<double> positions;       /* n_owned * dim */
    std::vector<uint64_t> global_index;  /* n_relevant */
    uint64_t n_global_dofs = 0;

    void fill_desc(rj_offline_desc &d) const;
  };

  /* throws std::runtime_error */
  void generate_cartesian_q1(HostOffline &out, const rj_mesh_desc &mesh, int rank, int n_ranks);

  /* SparsityPatternSIMD host layout -> CSR (sparse_matrix_simd.template.h:94-146) */
  void decode_simd_pattern(uint32_t n_rows,
                           uint32_t n_internal,
                           uint32_t simd_length,
                           const uint64_t *row_starts,
                           const uint32_t *column_indices,
                           std::vector<uint64_t> &csr_row_ptr,
                           std::vector<uint32_t> &csr_cols);

  /* flat data index of (row, col_idx, component) in the host layout the
   * descriptor uses (sparse_matrix_simd.h:389-448) */
  inline uint64_t host_entry_index(const rj_offline_desc &d,
                                   uint32_t row,
                                   uint32_t col_idx,
                                   uint32_t n_comp,
                                   uint32_t comp)
  {
    if (d.simd_length > 1 && row < d.n_internal) {
      const uint32_t L = d.simd_length;
      const uint32_t g = row / L;
      return (d.row_starts[g] + uint64_t(col_idx) * L) * n_comp + uint64_t(comp) * L + row % L;
    }
    return (d.row_starts[row] + col_idx) * n_comp + comp;
  }

  inline uint32_t host_row_length(const rj_offline_desc &d, uint32_t row)
  {
    if (d.simd_length > 1 && row < d.n_internal) {
      const uint32_t g = row / d.simd_length;
      return uint32_t((d.row_starts[g + 1] - d.row_starts[g]) / d.simd_length);
    }
    return uint32_t(d.row_starts[row + 1] - d.row_starts[row]);
  }

  inline uint32_t host_column(const rj_offline_desc &d, uint32_t row, uint32_t col_idx)
  {
    return d.column_indices[host_entry_index(d, row, col_idx, 1, 0)];
  }
} // namespace rjb
