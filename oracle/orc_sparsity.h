// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_common.h).
//
// Restatement of SparsityPatternSIMD / SparseMatrixSIMD indexing for a single
// rank (no ghost rows):
//   reinit (row_starts, column_indices, indices_transposed)
//       <- source/sparse_matrix_simd.template.h:39-147
//   entry addressing (SIMD part / CSR tail)
//       <- source/sparse_matrix_simd.h:389-448,585-646
// Input: a CSR pattern with "diagonal first, then ascending" rows (what
// dealii::SparsityPattern::copy_from produces for square patterns).
#pragma once

#include "orc_common.h"

#include <stdexcept>

namespace orc
{
  struct SparsityPatternSIMD {
    unsigned int simd_length = 1;
    unsigned int n_internal = 0;
    unsigned int n_rows = 0;
    /* size n_rows + 1; entries [0, n_internal/simd_length] index SIMD groups,
     * entries [n_internal, n_rows] index scalar rows (the reference reuses one
     * array for both). */
    std::vector<std::size_t> row_starts;
    std::vector<unsigned int> column_indices;
    std::vector<unsigned int> indices_transposed;

    /* position of (row, col) in the CSR input */
    static std::size_t find(const std::vector<unsigned int> &row_ptr,
                            const std::vector<unsigned int> &cols,
                            unsigned int row,
                            unsigned int col)
    {
      for (unsigned int e = row_ptr[row]; e < row_ptr[row + 1]; ++e)
        if (cols[e] == col)
          return e;
      throw std::runtime_error("SparsityPatternSIMD: transposed entry missing");
    }

    void reinit(unsigned int n_internal_dofs,
                unsigned int simd_len,
                const std::vector<unsigned int> &row_ptr,
                const std::vector<unsigned int> &cols)
    {
      simd_length = simd_len;
      n_internal = n_internal_dofs;
      n_rows = row_ptr.size() - 1;
      if (n_internal % simd_length != 0 || n_internal > n_rows)
        throw std::runtime_error("SparsityPatternSIMD: invalid n_internal");

      row_starts.assign(n_rows + 1, 0);
      column_indices.resize(cols.size());
      indices_transposed.resize(cols.size());

      auto row_len = [&](unsigned int r) { return row_ptr[r + 1] - row_ptr[r]; };

      auto transposed_of = [&](unsigned int row, unsigned int column) {
        const std::size_t position = find(row_ptr, cols, column, row);
        if (column < n_internal) {
          const unsigned int my_row_length = row_len(column);
          const std::size_t position_diag = row_ptr[column];
          const std::size_t pos_within_row = position - position_diag;
          const unsigned int simd_offset = column % simd_length;
          return (unsigned int)(position - simd_offset * my_row_length -
                                pos_within_row + simd_offset +
                                pos_within_row * simd_length);
        }
        return (unsigned int)position;
      };

      std::size_t ptr = 0;
      for (unsigned int i = 0; i < n_internal; i += simd_length) {
        const unsigned int len = row_len(i);
        for (unsigned int k = 1; k < simd_length; ++k)
          if (row_len(i + k) != len)
            throw std::runtime_error(
                "SparsityPatternSIMD: unequal row lengths in a SIMD group");
        for (unsigned int c = 0; c < len; ++c)
          for (unsigned int k = 0; k < simd_length; ++k) {
            const unsigned int column = cols[row_ptr[i + k] + c];
            column_indices[ptr] = column;
            indices_transposed[ptr] = transposed_of(i + k, column);
            ++ptr;
          }
        row_starts[i / simd_length + 1] = ptr;
      }

      row_starts[n_internal] = row_starts[n_internal / simd_length];

      for (unsigned int i = n_internal; i < n_rows; ++i) {
        for (unsigned int e = row_ptr[i]; e < row_ptr[i + 1]; ++e) {
          const unsigned int column = cols[e];
          column_indices[ptr] = column;
          indices_transposed[ptr] = transposed_of(i, column);
          ++ptr;
        }
        row_starts[i + 1] = ptr;
      }
    }

    unsigned int row_length(unsigned int row) const
    {
      if (row < n_internal) {
        const unsigned int g = row / simd_length;
        return (row_starts[g + 1] - row_starts[g]) / simd_length;
      }
      return row_starts[row + 1] - row_starts[row];
    }

    /* flat data index of entry (row, position_within_column), n_components=1 */
    std::size_t entry_index(unsigned int row, unsigned int pos) const
    {
      if (row < n_internal) {
        const unsigned int g = row / simd_length;
        return row_starts[g] + std::size_t(pos) * simd_length +
               row % simd_length;
      }
      return row_starts[row] + pos;
    }
  };
} // namespace orc
