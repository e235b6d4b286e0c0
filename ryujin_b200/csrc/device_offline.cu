// Host -> device re-layout of the stencil data (see device_offline.h).
//
// Index tables restated for slice height 32:
//   column_indices / indices_transposed  <- sparse_matrix_simd.template.h:94-147
//   ghost-row exchange lists             <- sparse_matrix_simd.template.h:152-267

#include "device_offline.h"
#include "host_offline.h"

#include <algorithm>
#include <cstring>

rj_offline::~rj_offline()
{
  if (host_only)
    return; /* no device allocations were made */
  auto f = [](void *p) {
    if (p)
      cudaFree(p);
  };
  cudaSetDevice(device);
  f(slice_offsets);
  f(row_length);
  f(cols);
  f(transposed);
  f(mij);
  f(cij);
  f(mi);
  f(mi_inv);
  f(bgroup_dof);
  f(bgroup_start);
  f(bnormal);
  f(bposition);
  f(bid);
  f(pair_i);
  f(pair_j);
  f(pair_e);
  f(pair_eT);
  f(pair_cji);
  f(send_rows);
  f(mat_send_entries);
  f(positions);
}

namespace rjb
{
  namespace
  {
    /* host_only: keep host mirrors, touch no CUDA API (layout tests on CPU boxes) */
    bool g_host_only = false;

    template <typename T>
    T *upload(const std::vector<T> &h)
    {
      T *d = nullptr;
      if (h.empty() || g_host_only)
        return d;
      RJ_CUDA(cudaMalloc(&d, h.size() * sizeof(T)));
      RJ_CUDA(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
      return d;
    }
  } // namespace


  rj_offline *build_device_offline(const rj_offline_desc &d, int device, bool host_only)
  {
    g_host_only = host_only;
    if (d.dim < 1 || d.dim > 3)
      throw InvalidArgument("rj_offline_create: dim must be 1, 2 or 3");
    if (!(d.n_export <= d.n_owned && d.n_owned <= d.n_relevant))
      throw InvalidArgument("rj_offline_create: need n_export <= n_owned <= n_relevant");
    if (d.simd_length == 0)
      throw InvalidArgument("rj_offline_create: simd_length must be >= 1");
    if (d.simd_length > 1 && (d.n_internal % d.simd_length != 0 || d.n_internal > d.n_owned))
      throw InvalidArgument("rj_offline_create: invalid n_internal");
    if (!d.row_starts || !d.column_indices || !d.mass_matrix || !d.cij_matrix ||
        !d.lumped_mass_matrix)
      throw InvalidArgument("rj_offline_create: null stencil arrays");

    if (!host_only) {
      int n_dev = 0;
      if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
        throw NoDevice("rj_offline_create: no CUDA device (there is no CPU fallback)");
      RJ_CUDA(cudaSetDevice(device));
    }

    auto *od = new rj_offline();
    try {
      od->device = device;
      od->host_only = host_only;
      od->dim = d.dim;
      od->n_export = d.n_export;
      od->n_owned = d.n_owned;
      od->n_relevant = d.n_relevant;
      od->ld = ((d.n_relevant + 127u) / 128u) * 128u;
      if (od->ld == 0)
        od->ld = 128;
      od->n_slices = (d.n_owned + 31u) / 32u;
      od->measure_of_omega = d.measure_of_omega;
      const int dim = d.dim;
      const uint32_t n_owned = d.n_owned;

      /* row lengths and slice offsets */
      std::vector<uint32_t> &row_len = od->h_row_length;
      row_len.assign(size_t(od->n_slices) * 32, 0);
#pragma omp parallel for schedule(static)
      for (uint32_t i = 0; i < n_owned; ++i)
        row_len[i] = host_row_length(d, i);
      for (uint32_t i = 0; i < n_owned; ++i)
        if (row_len[i] > 255 || row_len[i] == 0)
          throw InvalidArgument("rj_offline_create: row length must be in [1, 255]");

      std::vector<uint32_t> &slice_off = od->h_slice_offsets;
      slice_off.assign(size_t(od->n_slices) + 1, 0);
      {
        uint64_t acc = 0;
        for (uint32_t g = 0; g < od->n_slices; ++g) {
          uint32_t w = 0;
          for (int l = 0; l < 32; ++l)
            w = std::max(w, row_len[size_t(g) * 32 + l]);
          acc += uint64_t(w) * 32;
          if (acc >= (uint64_t(1) << 32))
            throw InvalidArgument("rj_offline_create: more than 2^32 matrix entries per rank");
          slice_off[g + 1] = uint32_t(acc);
        }
      }
      const uint32_t E_owned = slice_off[od->n_slices];
      od->E_owned = E_owned;

      /* ghost rows: for each ghost j the owned neighbours i (ascending) */
      const uint32_t n_ghost = d.n_relevant - d.n_owned;
      std::vector<uint32_t> ghost_row_start(size_t(n_ghost) + 1, 0);
      if (n_ghost > 0) {
        for (uint32_t i = 0; i < n_owned; ++i)
          for (uint32_t c = 1; c < row_len[i]; ++c) {
            const uint32_t j = host_column(d, i, c);
            if (j >= d.n_relevant)
              throw InvalidArgument("rj_offline_create: column index out of range");
            if (j >= n_owned) {
              ghost_row_start[j - n_owned + 1]++;
              /* the two-stream overlap launches rows [0, n_export) before the exchange: every row
               * that reads a ghost column has to be one of them (offline_data.h:136-169) */
              if (i >= d.n_export)
                od->export_rows_first = false;
            }
          }
        for (uint32_t q = 0; q < n_ghost; ++q)
          ghost_row_start[q + 1] += ghost_row_start[q];
      }
      if (uint64_t(E_owned) + ghost_row_start[n_ghost] >= (uint64_t(1) << 32))
        throw InvalidArgument("rj_offline_create: more than 2^32 matrix entries per rank");
      od->E_total = E_owned + ghost_row_start[n_ghost];

      /* pattern in sliced layout */
      std::vector<uint32_t> cols(E_owned), idxT(E_owned);
#pragma omp parallel for schedule(static)
      for (uint32_t g = 0; g < od->n_slices; ++g)
        for (uint32_t e = slice_off[g]; e < slice_off[g + 1]; ++e) {
          const uint32_t row = g * 32 + (e - slice_off[g]) % 32;
          cols[e] = row < n_owned ? row : 0; /* padding points at the diagonal */
          idxT[e] = row < n_owned ? slice_off[g] + row % 32 : 0;
        }

      /* position of column `i` in owned row j: rows are diag first, then ascending */
      auto position_in_row = [&](uint32_t j, uint32_t i) -> uint32_t {
        if (i == j)
          return 0;
        uint32_t lo = 1, hi = row_len[j];
        while (lo < hi) {
          const uint32_t mid = (lo + hi) / 2;
          if (host_column(d, j, mid) < i)
            lo = mid + 1;
          else
            hi = mid;
        }
        if (lo >= row_len[j] || host_column(d, j, lo) != i)
          throw InvalidArgument("rj_offline_create: stencil is not structurally symmetric "
                                "(or rows are not sorted diag-first/ascending)");
        return lo;
      };

      {
        std::string error;
#pragma omp parallel for schedule(static)
        for (uint32_t i = 0; i < n_owned; ++i) {
          try {
            if (host_column(d, i, 0) != i)
              throw InvalidArgument("rj_offline_create: rows must list the diagonal first");
            for (uint32_t c = 0; c < row_len[i]; ++c) {
              const uint32_t j = host_column(d, i, c);
              const uint32_t e = od->entry(i, c);
              cols[e] = j;
              if (j < n_owned)
                idxT[e] = od->entry(j, position_in_row(j, i));
            }
          } catch (const std::exception &ex) {
#pragma omp critical
            error = ex.what();
          }
        }
        if (!error.empty())
          throw InvalidArgument(error);
      }
      if (n_ghost > 0) {
        /* serial: assigns slots of ghost row j in ascending i */
        std::vector<uint32_t> fill(ghost_row_start.begin(), ghost_row_start.end() - 1);
        for (uint32_t i = 0; i < n_owned; ++i)
          for (uint32_t c = 1; c < row_len[i]; ++c) {
            const uint32_t j = host_column(d, i, c);
            if (j >= n_owned)
              idxT[od->entry(i, c)] = E_owned + fill[j - n_owned]++;
          }
      }

      {
        std::vector<uint8_t> rl8(row_len.size());
        for (size_t i = 0; i < row_len.size(); ++i)
          rl8[i] = uint8_t(row_len[i]);
        od->row_length = upload(rl8);
      }
      od->slice_offsets = upload(slice_off);
      od->cols = upload(cols);
      od->transposed = upload(idxT);

      /* matrices */
      {
        std::vector<double> buf(E_owned, 0.);
#pragma omp parallel for schedule(static)
        for (uint32_t i = 0; i < n_owned; ++i)
          for (uint32_t c = 0; c < row_len[i]; ++c)
            buf[od->entry(i, c)] = d.mass_matrix[host_entry_index(d, i, c, 1, 0)];
        od->mij = upload(buf);

        if (!host_only)
          RJ_CUDA(cudaMalloc(&od->cij, size_t(dim) * E_owned * sizeof(double)));
        for (int q = 0; q < dim && !host_only; ++q) {
          std::fill(buf.begin(), buf.end(), 0.);
#pragma omp parallel for schedule(static)
          for (uint32_t i = 0; i < n_owned; ++i)
            for (uint32_t c = 0; c < row_len[i]; ++c)
              buf[od->entry(i, c)] = d.cij_matrix[host_entry_index(d, i, c, dim, q)];
          RJ_CUDA(cudaMemcpy(od->cij + size_t(q) * E_owned,
                             buf.data(),
                             size_t(E_owned) * sizeof(double),
                             cudaMemcpyHostToDevice));
        }
      }

      /* lumped mass */
      {
        std::vector<double> m(od->ld, 1.), minv(od->ld, 1.);
        for (uint32_t i = 0; i < d.n_relevant; ++i) {
          m[i] = d.lumped_mass_matrix[i];
          minv[i] = 1. / m[i];
        }
        od->mi = upload(m);
        od->mi_inv = upload(minv);
      }

      /* boundary map: group entries by dof (input sorted by dof) */
      {
        std::vector<uint32_t> gdof, gstart;
        for (uint32_t b = 0; b < d.n_boundary; ++b) {
          if (d.boundary_dof[b] >= n_owned)
            throw InvalidArgument("rj_offline_create: boundary dof out of range");
          if (b == 0 || d.boundary_dof[b] != d.boundary_dof[b - 1]) {
            if (b > 0 && d.boundary_dof[b] < d.boundary_dof[b - 1])
              throw InvalidArgument("rj_offline_create: boundary map must be sorted by dof");
            gdof.push_back(d.boundary_dof[b]);
            gstart.push_back(b);
          }
        }
        gstart.push_back(d.n_boundary);
        od->n_bgroups = uint32_t(gdof.size());
        od->n_bentries = d.n_boundary;
        od->bgroup_dof = upload(gdof);
        od->bgroup_start = upload(gstart);
        if (d.n_boundary > 0) {
          std::vector<double> nrm(d.boundary_normal, d.boundary_normal + size_t(d.n_boundary) * dim);
          std::vector<double> pos(size_t(d.n_boundary) * dim, 0.);
          if (d.boundary_position)
            pos.assign(d.boundary_position, d.boundary_position + size_t(d.n_boundary) * dim);
          std::vector<int32_t> ids(d.boundary_id, d.boundary_id + d.n_boundary);
          od->bnormal = upload(nrm);
          od->bposition = upload(pos);
          od->bid = upload(ids);
          od->h_boundary_ids = ids;
        }
      }

      /* coupling boundary pairs: keep i < j only (hyperbolic_module.template.h:473-474) */
      {
        std::vector<uint32_t> pi, pj, pe, peT;
        std::vector<double> cji;
        for (uint32_t p = 0; p < d.n_coupling_pairs; ++p) {
          const uint32_t i = d.coupling_pairs[3 * p], c = d.coupling_pairs[3 * p + 1],
                         j = d.coupling_pairs[3 * p + 2];
          if (i >= n_owned || c >= row_len[i] || host_column(d, i, c) != j)
            throw InvalidArgument("rj_offline_create: inconsistent coupling boundary pair");
          if (j < i)
            continue;
          const uint32_t e = od->entry(i, c);
          pi.push_back(i);
          pj.push_back(j);
          pe.push_back(e);
          peT.push_back(idxT[e]);
          for (int q = 0; q < dim; ++q) {
            if (d.coupling_cji)
              cji.push_back(d.coupling_cji[size_t(p) * dim + q]);
            else {
              if (j >= n_owned)
                throw InvalidArgument("rj_offline_create: coupling_cji required for ghost rows");
              cji.push_back(
                  d.cij_matrix[host_entry_index(d, j, position_in_row(j, i), dim, q)]);
            }
          }
        }
        od->n_pairs = uint32_t(pi.size());
        od->pair_i = upload(pi);
        od->pair_j = upload(pj);
        od->pair_e = upload(pe);
        od->pair_eT = upload(peT);
        od->pair_cji = upload(cji);
      }

      /* exchange lists */
      {
        std::vector<uint32_t> send_rows, mat_send;
        uint32_t ghost_start = n_owned;
        for (uint32_t p = 0; p < d.n_neighbors; ++p) {
          rj_offline::Neighbor nb;
          nb.rank = d.neighbor_rank[p];
          nb.ghost_start = ghost_start;
          nb.ghost_count = d.ghost_count[p];
          ghost_start += nb.ghost_count;
          nb.send_offset = uint32_t(send_rows.size());
          nb.mat_send_offset = uint32_t(mat_send.size());
          for (uint32_t s = d.send_offsets[p]; s < d.send_offsets[p + 1]; ++s) {
            const uint32_t row = d.send_indices[s];
            if (row >= d.n_export)
              od->export_rows_first = false; /* ... and so has every row that is sent */
            if (row >= n_owned)
              throw InvalidArgument("rj_offline_create: send index out of range");
            send_rows.push_back(row);
            /* entries of `row` whose column lies in the ghost range of p */
            for (uint32_t c = 1; c < row_len[row]; ++c) {
              const uint32_t j = host_column(d, row, c);
              if (j >= nb.ghost_start && j < nb.ghost_start + nb.ghost_count)
                mat_send.push_back(od->entry(row, c));
            }
          }
          nb.send_count = uint32_t(send_rows.size()) - nb.send_offset;
          nb.mat_send_count = uint32_t(mat_send.size()) - nb.mat_send_offset;
          nb.mat_recv_offset = E_owned + ghost_row_start[nb.ghost_start - n_owned];
          nb.mat_recv_count = ghost_row_start[nb.ghost_start + nb.ghost_count - n_owned] -
                              ghost_row_start[nb.ghost_start - n_owned];
          od->neighbors.push_back(nb);
        }
        if (ghost_start != d.n_relevant)
          throw InvalidArgument("rj_offline_create: ghost counts do not add up to n_relevant");
        od->n_send_rows = uint32_t(send_rows.size());
        od->n_mat_send = uint32_t(mat_send.size());
        od->send_rows = upload(send_rows);
        od->mat_send_entries = upload(mat_send);
        od->h_send_rows = send_rows;
        od->h_mat_send_entries = mat_send;
      }

      if (d.positions) {
        std::vector<double> pos(size_t(dim) * n_owned);
        for (uint32_t i = 0; i < n_owned; ++i)
          for (int q = 0; q < dim; ++q)
            pos[size_t(q) * n_owned + i] = d.positions[size_t(i) * dim + q];
        od->positions = upload(pos);
      }

      if (host_only) {
        od->h_cols.swap(cols);
        od->h_transposed.swap(idxT);
      }
    } catch (...) {
      delete od;
      throw;
    }
    return od;
  }
} // namespace rjb
