// See host_offline.h.  Pure host C++ (no CUDA).

#include "host_offline.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <limits>
#include <stdexcept>

namespace rjb
{
  void HostOffline::fill_desc(rj_offline_desc &d) const
  {
    d = rj_offline_desc{};
    d.dim = dim;
    d.n_export = n_export;
    d.n_internal = 0;
    d.n_owned = n_owned;
    d.n_relevant = n_relevant;
    d.simd_length = 1;
    d.row_starts = row_starts.data();
    d.column_indices = cols.data();
    d.mass_matrix = mij.data();
    d.cij_matrix = cij.data();
    d.lumped_mass_matrix = mi.data();
    d.measure_of_omega = measure_of_omega;
    d.n_boundary = uint32_t(boundary_dof.size());
    d.boundary_dof = boundary_dof.data();
    d.boundary_normal = boundary_normal.data();
    d.boundary_id = boundary_id.data();
    d.boundary_position = boundary_position.data();
    d.n_coupling_pairs = uint32_t(coupling_pairs.size() / 3);
    d.coupling_pairs = coupling_pairs.data();
    d.coupling_cji = coupling_cji.data();
    d.n_neighbors = uint32_t(neighbor_rank.size());
    d.neighbor_rank = neighbor_rank.data();
    d.ghost_count = ghost_count.data();
    d.send_offsets = send_offsets.data();
    d.send_indices = send_indices.data();
    d.positions = positions.data();
  }


  namespace
  {
    constexpr uint32_t invalid = std::numeric_limits<uint32_t>::max();

    struct Grid {
      int dim;
      int64_t nc[3]; /* cells */
      int64_t nn[3]; /* nodes */
      double h[3], origin[3];
      const uint8_t *mask;
      int32_t box_ids[6];
      int32_t masked_id;

      bool cell_active(const int64_t c[3]) const
      {
        for (int d = 0; d < dim; ++d)
          if (c[d] < 0 || c[d] >= nc[d])
            return false;
        if (!mask)
          return true;
        int64_t idx = 0;
        for (int d = dim - 1; d >= 0; --d)
          idx = idx * nc[d] + c[d];
        return mask[idx] != 0;
      }

      /* cell adjacent to node idx with corner bits `code` (bit d = 1: cell on
       * the + side of the node in direction d); lexicographic cell order */
      void adjacent_cell(const int64_t idx[3], int code, int64_t c[3]) const
      {
        for (int d = 0; d < dim; ++d)
          c[d] = idx[d] - 1 + ((code >> d) & 1);
        for (int d = dim; d < 3; ++d)
          c[d] = 0;
      }

      bool node_active(const int64_t idx[3]) const
      {
        for (int code = 0; code < (1 << dim); ++code) {
          int64_t c[3];
          adjacent_cell(idx, code, c);
          if (cell_active(c))
            return true;
        }
        return false;
      }

      /* on a (non periodic) boundary face: some adjacent cell slot inactive */
      bool node_on_boundary(const int64_t idx[3]) const
      {
        bool any_inactive = false, any_active = false;
        for (int code = 0; code < (1 << dim); ++code) {
          int64_t c[3];
          adjacent_cell(idx, code, c);
          if (cell_active(c))
            any_active = true;
          else
            any_inactive = true;
        }
        return any_active && any_inactive;
      }

      int64_t lex_node(const int64_t idx[3]) const
      {
        int64_t r = 0;
        for (int d = dim - 1; d >= 0; --d)
          r = r * nn[d] + idx[d];
        return r;
      }

      void node_from_lex(int64_t n, int64_t idx[3]) const
      {
        for (int d = 0; d < dim; ++d) {
          idx[d] = n % nn[d];
          n /= nn[d];
        }
        for (int d = dim; d < 3; ++d)
          idx[d] = 0;
      }
    };

    /* Closed-form Q1 cell matrices on a box cell: vertices coded x fastest. */
    struct CellMatrices {
      double mass[8][8];
      double c[3][8][8];
      double lumped[8]; /* row sums in vertex order */
      double volume;

      CellMatrices(int dim, const double h[3])
      {
        const double M1[2][2] = {{1. / 3., 1. / 6.}, {1. / 6., 1. / 3.}};
        const int nv = 1 << dim;
        volume = 1.;
        for (int d = 0; d < dim; ++d)
          volume *= h[d];
        for (int a = 0; a < nv; ++a) {
          for (int b = 0; b < nv; ++b) {
            double m = 1.;
            for (int d = 0; d < dim; ++d)
              m *= h[d] * M1[(a >> d) & 1][(b >> d) & 1];
            mass[a][b] = m;
            for (int g = 0; g < dim; ++g) {
              double v = ((b >> g) & 1) ? 0.5 : -0.5;
              for (int d = 0; d < dim; ++d)
                if (d != g)
                  v *= h[d] * M1[(a >> d) & 1][(b >> d) & 1];
              c[g][a][b] = v;
            }
          }
        }
        for (int a = 0; a < nv; ++a) {
          double s = 0.;
          for (int b = 0; b < nv; ++b)
            s += mass[a][b];
          lumped[a] = s;
        }
      }
    };

    inline int offset_code(int dim, const int o[3])
    {
      int code = 0, mul = 1;
      for (int d = 0; d < dim; ++d) {
        code += (o[d] + 1) * mul;
        mul *= 3;
      }
      return code;
    }

    /* Row of the stencil at node idx, accumulated cell by cell. */
    struct Row {
      bool present[27];
      double m[27];
      double c[27][3];
    };

    void assemble_row(const Grid &g, const CellMatrices &cm, const int64_t idx[3], Row &row)
    {
      const int dim = g.dim, nv = 1 << dim;
      for (int q = 0; q < 27; ++q) {
        row.present[q] = false;
        row.m[q] = 0.;
        row.c[q][0] = row.c[q][1] = row.c[q][2] = 0.;
      }
      for (int code = 0; code < nv; ++code) {
        int64_t c[3];
        g.adjacent_cell(idx, code, c);
        if (!g.cell_active(c))
          continue;
        /* local vertex of idx inside the cell: a_d = idx_d - c_d */
        int a = 0;
        for (int d = 0; d < dim; ++d)
          a |= int(idx[d] - c[d]) << d;
        for (int b = 0; b < nv; ++b) {
          int o[3] = {0, 0, 0};
          for (int d = 0; d < dim; ++d)
            o[d] = ((b >> d) & 1) - ((a >> d) & 1);
          const int q = offset_code(dim, o);
          row.present[q] = true;
          row.m[q] += cm.mass[a][b];
          for (int e = 0; e < dim; ++e)
            row.c[q][e] += cm.c[e][a][b];
        }
      }
    }

    /* numbering independent lumped mass: sum over offsets in offset order */
    double lumped_mass(const Grid &g, const CellMatrices &cm, const int64_t idx[3])
    {
      Row row;
      assemble_row(g, cm, idx, row);
      double s = 0.;
      int n_off = 1;
      for (int d = 0; d < g.dim; ++d)
        n_off *= 3;
      for (int q = 0; q < n_off; ++q)
        if (row.present[q])
          s += row.m[q];
      return s;
    }
  } // namespace


  void generate_cartesian_q1(HostOffline &out, const rj_mesh_desc &mesh, int rank, int n_ranks)
  {
    const int dim = mesh.dim;
    if (dim < 1 || dim > 3)
      throw std::runtime_error("rj_mesh_generate: dim must be 1, 2 or 3");
    if (n_ranks < 1 || rank < 0 || rank >= n_ranks)
      throw std::runtime_error("rj_mesh_generate: invalid rank / n_ranks");

    Grid g;
    g.dim = dim;
    for (int d = 0; d < 3; ++d) {
      g.nc[d] = d < dim ? mesh.n_cells[d] : 0;
      g.nn[d] = d < dim ? int64_t(mesh.n_cells[d]) + 1 : 1;
      g.h[d] = d < dim ? mesh.h[d] : 1.;
      g.origin[d] = d < dim ? mesh.origin[d] : 0.;
      if (d < dim && (mesh.n_cells[d] == 0 || !(mesh.h[d] > 0.)))
        throw std::runtime_error("rj_mesh_generate: empty mesh direction");
    }
    g.mask = mesh.cell_mask;
    for (int f = 0; f < 6; ++f)
      g.box_ids[f] = mesh.box_boundary_ids[f];
    g.masked_id = mesh.masked_boundary_id;

    const CellMatrices cm(dim, g.h);
    const int nv = 1 << dim;

    /* Partition: blocks of the node grid, P[0] x P[1] x P[2] ranks (x fastest), as p4est /
     * deal.II hand contiguous blocks of the structured mesh to the ranks
     * (offline_data.template.h:353-360).  Default: slabs along the last axis. */
    int64_t P[3] = {1, 1, 1};
    if (mesh.partition[0] == 0 && mesh.partition[1] == 0 && mesh.partition[2] == 0)
      P[dim - 1] = n_ranks;
    else
      for (int d = 0; d < 3; ++d)
        P[d] = d < dim ? std::max<int64_t>(mesh.partition[d], 1) : 1;
    if (P[0] * P[1] * P[2] != n_ranks)
      throw std::runtime_error("rj_mesh_generate: partition does not multiply to n_ranks");
    for (int d = 0; d < dim; ++d)
      if (g.nn[d] < P[d])
        throw std::runtime_error("rj_mesh_generate: more ranks than node planes");
    const int64_t my_block[3] = {rank % P[0], (rank / P[0]) % P[1], rank / (P[0] * P[1])};
    int64_t lo[3], hi[3], wlo[3], whi[3];
    for (int d = 0; d < 3; ++d) {
      lo[d] = g.nn[d] * my_block[d] / P[d];
      hi[d] = g.nn[d] * (my_block[d] + 1) / P[d];
      wlo[d] = std::max<int64_t>(lo[d] - 1, 0);
      whi[d] = std::min<int64_t>(hi[d] + 1, g.nn[d]);
    }
    auto block_of = [&](int d, int64_t x) -> int64_t {
      int64_t bq = x * P[d] / g.nn[d];
      while (bq + 1 < P[d] && x >= g.nn[d] * (bq + 1) / P[d])
        ++bq;
      while (bq > 0 && x < g.nn[d] * bq / P[d])
        --bq;
      return bq;
    };
    auto owned = [&](const int64_t idx[3]) {
      for (int d = 0; d < dim; ++d)
        if (idx[d] < lo[d] || idx[d] >= hi[d])
          return false;
      return true;
    };
    auto owner_rank = [&](const int64_t idx[3]) -> int {
      int64_t bq[3] = {0, 0, 0};
      for (int d = 0; d < dim; ++d)
        bq[d] = block_of(d, idx[d]);
      return int(bq[0] + P[0] * (bq[1] + P[1] * bq[2]));
    };

    /* global dof numbering (lexicographic over active nodes) */
    std::vector<uint32_t> global_of_node; /* only for masked meshes */
    uint64_t n_global = 0;
    int64_t n_nodes_total = 1;
    for (int d = 0; d < dim; ++d)
      n_nodes_total *= g.nn[d];
    if (g.mask) {
      if (n_nodes_total >= int64_t(invalid))
        throw std::runtime_error("rj_mesh_generate: masked mesh too large");
      global_of_node.assign(n_nodes_total, invalid);
      for (int64_t n = 0; n < n_nodes_total; ++n) {
        int64_t idx[3];
        g.node_from_lex(n, idx);
        if (g.node_active(idx))
          global_of_node[n] = uint32_t(n_global++);
      }
    } else {
      n_global = uint64_t(n_nodes_total);
    }
    auto global_dof = [&](int64_t lex) -> uint64_t {
      return g.mask ? uint64_t(global_of_node[lex]) : uint64_t(lex);
    };

    /* local window: the owned box plus one layer of nodes, clipped */
    int64_t wn[3] = {1, 1, 1};
    for (int d = 0; d < dim; ++d)
      wn[d] = whi[d] - wlo[d];
    const int64_t window = wn[0] * wn[1] * wn[2];
    if (window >= int64_t(invalid))
      throw std::runtime_error("rj_mesh_generate: more than 2^32 rows per rank");
    std::vector<uint32_t> local_of(window, invalid);
    auto win_of = [&](const int64_t idx[3]) -> int64_t {
      int64_t r = 0;
      for (int d = dim - 1; d >= 0; --d)
        r = r * wn[d] + (idx[d] - wlo[d]);
      return r;
    };
    auto win = [&](int64_t lex) -> int64_t {
      int64_t idx[3];
      g.node_from_lex(lex, idx);
      return win_of(idx);
    };
    /* window nodes in ascending global lexicographic order */
    auto window_lex = [&](int64_t w, int64_t idx[3]) -> int64_t {
      idx[0] = idx[1] = idx[2] = 0;
      for (int d = 0; d < dim; ++d) {
        idx[d] = wlo[d] + w % wn[d];
        w /= wn[d];
      }
      return g.lex_node(idx);
    };

    /* pass A: classify the window's nodes.
     * sendmask (owned nodes): bit o = the node shares an active cell with a node of the block at
     * relative offset o (offset_code over {-1,0,1}^dim); non-zero = export row.
     * ghost (other nodes): shares an active cell with an owned node. */
    std::vector<uint32_t> sendmask(window, 0);
    std::vector<uint8_t> kind(window, 0); /* 1 owned & active, 2 ghost */
#pragma omp parallel for schedule(static)
    for (int64_t w = 0; w < window; ++w) {
      int64_t idx[3];
      window_lex(w, idx);
      const bool mine = owned(idx);
      bool active = false, touches_owned = false;
      uint32_t mask_bits = 0;
      for (int code = 0; code < nv; ++code) {
        int64_t c[3];
        g.adjacent_cell(idx, code, c);
        if (!g.cell_active(c))
          continue;
        active = true;
        for (int v = 0; v < nv; ++v) {
          int64_t vtx[3] = {0, 0, 0};
          for (int d = 0; d < dim; ++d)
            vtx[d] = c[d] + ((v >> d) & 1);
          if (owned(vtx)) {
            touches_owned = true;
            continue;
          }
          if (mine) {
            int o[3] = {0, 0, 0};
            for (int d = 0; d < dim; ++d)
              o[d] = int(block_of(d, vtx[d]) - my_block[d]);
            mask_bits |= 1u << offset_code(dim, o);
          }
        }
      }
      if (mine && active) {
        kind[w] = 1;
        sendmask[w] = mask_bits;
      } else if (!mine && active && touches_owned)
        kind[w] = 2;
    }

    /* local numbering: export rows, other owned rows (both in global lexicographic order), then the
     * ghosts grouped by owner rank (ascending), each group in global lexicographic order -- the
     * order in which the owner sends them */
    std::vector<int64_t> node_of_local;
    std::vector<std::pair<int, int64_t>> ghosts; /* (owner rank, lex) */
    {
      uint32_t n = 0;
      out.n_export = 0;
      for (int pass = 0; pass < 2; ++pass)
        for (int64_t w = 0; w < window; ++w) {
          if (kind[w] != 1)
            continue;
          const bool is_export = sendmask[w] != 0;
          if ((pass == 0) == is_export) {
            int64_t idx[3];
            local_of[w] = n++;
            node_of_local.push_back(window_lex(w, idx));
            if (is_export)
              out.n_export++;
          }
        }
      out.n_owned = n;
      for (int64_t w = 0; w < window; ++w)
        if (kind[w] == 2) {
          int64_t idx[3];
          const int64_t lex = window_lex(w, idx);
          ghosts.emplace_back(owner_rank(idx), lex);
        }
      std::sort(ghosts.begin(), ghosts.end());
      for (const auto &gh : ghosts) {
        local_of[win(gh.second)] = n++;
        node_of_local.push_back(gh.second);
      }
      out.n_relevant = n;
    }
    out.dim = dim;
    out.n_internal = 0;
    out.n_global_dofs = n_global;

    /* exchange lists: one neighbour per rank that owns ghosts of this rank (ascending rank); the
     * rows sent to it are this rank's export rows that share a cell with one of its nodes */
    out.neighbor_rank.clear();
    out.ghost_count.clear();
    out.send_offsets.assign(1, 0);
    out.send_indices.clear();
    for (size_t q = 0; q < ghosts.size();) {
      const int nb_rank = ghosts[q].first;
      size_t q1 = q;
      while (q1 < ghosts.size() && ghosts[q1].first == nb_rank)
        ++q1;
      out.neighbor_rank.push_back(nb_rank);
      out.ghost_count.push_back(uint32_t(q1 - q));
      const int64_t nb_block[3] = {nb_rank % P[0], (nb_rank / P[0]) % P[1], nb_rank / (P[0] * P[1])};
      int o[3] = {0, 0, 0};
      for (int d = 0; d < dim; ++d)
        o[d] = int(nb_block[d] - my_block[d]);
      const uint32_t bit = 1u << offset_code(dim, o);
      for (uint32_t i = 0; i < out.n_export; ++i)
        if (sendmask[win(node_of_local[i])] & bit)
          out.send_indices.push_back(i);
      out.send_offsets.push_back(uint32_t(out.send_indices.size()));
      q = q1;
    }

    /* rows: count, then fill */
    const uint32_t n_owned = out.n_owned;
    int n_off = 1;
    for (int d = 0; d < dim; ++d)
      n_off *= 3;
    const int center = (n_off - 1) / 2;

    out.row_starts.assign(size_t(n_owned) + 1, 0);
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < n_owned; ++i) {
      int64_t idx[3];
      g.node_from_lex(node_of_local[i], idx);
      Row row;
      assemble_row(g, cm, idx, row);
      uint32_t len = 0;
      for (int q = 0; q < n_off; ++q)
        len += row.present[q];
      out.row_starts[i + 1] = len;
    }
    for (uint32_t i = 0; i < n_owned; ++i)
      out.row_starts[i + 1] += out.row_starts[i];
    const uint64_t nnz = out.row_starts[n_owned];
    out.cols.resize(nnz);
    out.mij.resize(nnz);
    out.cij.resize(nnz * dim);

    int64_t off_stride[3] = {1, g.nn[0], g.nn[0] * g.nn[1]};

#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < n_owned; ++i) {
      const int64_t lex = node_of_local[i];
      int64_t idx[3];
      g.node_from_lex(lex, idx);
      Row row;
      assemble_row(g, cm, idx, row);

      struct Ent {
        uint32_t col;
        int q;
      };
      Ent ents[27];
      int ne = 0;
      for (int q = 0; q < n_off; ++q) {
        if (!row.present[q] || q == center)
          continue;
        int64_t nlex = lex;
        int qq = q;
        for (int d = 0; d < dim; ++d) {
          nlex += (qq % 3 - 1) * off_stride[d];
          qq /= 3;
        }
        ents[ne++] = {local_of[win(nlex)], q};
      }
      std::sort(ents, ents + ne, [](const Ent &a, const Ent &b) { return a.col < b.col; });

      uint64_t e = out.row_starts[i];
      auto emit = [&](uint32_t col, int q) {
        out.cols[e] = col;
        out.mij[e] = row.m[q];
        for (int d = 0; d < dim; ++d)
          out.cij[e * dim + d] = row.c[q][d];
        ++e;
      };
      emit(i, center);
      for (int n = 0; n < ne; ++n)
        emit(ents[n].col, ents[n].q);
    }

    /* lumped mass for all relevant rows, positions, global indices */
    out.mi.resize(out.n_relevant);
    out.global_index.resize(out.n_relevant);
    out.positions.resize(size_t(n_owned) * dim);
#pragma omp parallel for schedule(static)
    for (uint32_t i = 0; i < out.n_relevant; ++i) {
      int64_t idx[3];
      g.node_from_lex(node_of_local[i], idx);
      out.mi[i] = lumped_mass(g, cm, idx);
      out.global_index[i] = global_dof(node_of_local[i]);
      if (i < n_owned)
        for (int d = 0; d < dim; ++d)
          out.positions[size_t(i) * dim + d] = g.origin[d] + double(idx[d]) * g.h[d];
    }

    /* |Omega|: global, numbering independent */
    {
      uint64_t n_active = 0;
      if (g.mask) {
        int64_t total = 1;
        for (int d = 0; d < dim; ++d)
          total *= g.nc[d];
        for (int64_t c = 0; c < total; ++c)
          n_active += g.mask[c] != 0;
      } else {
        n_active = 1;
        for (int d = 0; d < dim; ++d)
          n_active *= uint64_t(g.nc[d]);
      }
      out.measure_of_omega = double(n_active) * cm.volume;
    }

    /* boundary map (sorted by local dof; entries of a dof in insertion order:
     * cells lexicographic, faces x-,x+,y-,y+,z-,z+) */
    out.boundary_dof.clear();
    out.boundary_normal.clear();
    out.boundary_normal_mass.clear();
    out.boundary_mass.clear();
    out.boundary_id.clear();
    out.boundary_position.clear();
    out.coupling_pairs.clear();
    out.coupling_cji.clear();

    struct BEntry {
      double normal[3];
      double boundary_mass;
      int32_t id;
    };

    for (uint32_t i = 0; i < n_owned; ++i) {
      int64_t idx[3];
      g.node_from_lex(node_of_local[i], idx);
      if (!g.node_on_boundary(idx))
        continue;

      std::vector<BEntry> filtered;
      for (int code = 0; code < nv; ++code) {
        int64_t c[3];
        g.adjacent_cell(idx, code, c);
        if (!g.cell_active(c))
          continue;
        for (int f = 0; f < 2 * dim; ++f) {
          const int fd = f / 2, side = f % 2;
          /* node must lie on the face */
          if (int(idx[fd] - c[fd]) != side)
            continue;
          int64_t nb[3] = {c[0], c[1], c[2]};
          nb[fd] += side ? 1 : -1;
          if (g.cell_active(nb))
            continue;
          const bool outer = side ? (c[fd] + 1 == g.nc[fd]) : (c[fd] == 0);
          const int32_t id = outer ? g.box_ids[f] : g.masked_id;
          if (id == RJ_BOUNDARY_PERIODIC)
            continue;

          double face_weight = 1.; /* int_face phi_v */
          for (int d = 0; d < dim; ++d)
            if (d != fd)
              face_weight *= g.h[d] * 0.5;

          BEntry nw{};
          nw.normal[fd] = side ? face_weight : -face_weight;
          nw.boundary_mass = face_weight;
          nw.id = id;

          bool inserted = false;
          for (auto &old : filtered) {
            if (old.id != nw.id)
              continue;
            double dotp = 0., n1 = 0., n2 = 0.;
            for (int d = 0; d < dim; ++d) {
              dotp += old.normal[d] * nw.normal[d];
              n1 += old.normal[d] * old.normal[d];
              n2 += nw.normal[d] * nw.normal[d];
            }
            if (dotp / std::sqrt(n1) / std::sqrt(n2) > 0.50) {
              for (int d = 0; d < dim; ++d)
                old.normal[d] += nw.normal[d];
              old.boundary_mass += nw.boundary_mass;
              inserted = true;
              continue;
            } else if (dim == 2) {
              if (nw.id == RJ_BOUNDARY_SLIP) {
                nw.id = RJ_BOUNDARY_NO_SLIP;
                old.id = RJ_BOUNDARY_NO_SLIP;
              }
            }
          }
          if (!inserted)
            filtered.push_back(nw);
        }
      }

      for (const auto &b : filtered) {
        double nrm = 0.;
        for (int d = 0; d < dim; ++d)
          nrm += b.normal[d] * b.normal[d];
        const double normal_mass = std::sqrt(nrm) + std::numeric_limits<double>::epsilon();
        out.boundary_dof.push_back(i);
        for (int d = 0; d < dim; ++d) {
          out.boundary_normal.push_back(b.normal[d] / normal_mass);
          out.boundary_position.push_back(g.origin[d] + double(idx[d]) * g.h[d]);
        }
        out.boundary_normal_mass.push_back(normal_mass);
        out.boundary_mass.push_back(b.boundary_mass);
        out.boundary_id.push_back(b.id);
      }

      /* coupling boundary pairs of row i */
      const uint64_t e0 = out.row_starts[i];
      const uint32_t len = uint32_t(out.row_starts[i + 1] - e0);
      for (uint32_t col_idx = 1; col_idx < len; ++col_idx) {
        const uint32_t j = out.cols[e0 + col_idx];
        int64_t jdx[3];
        g.node_from_lex(node_of_local[j], jdx);
        if (g.node_on_boundary(jdx)) {
          out.coupling_pairs.push_back(i);
          out.coupling_pairs.push_back(col_idx);
          out.coupling_pairs.push_back(j);
          /* c_ji from row j (which may be a ghost row) */
          Row rj;
          assemble_row(g, cm, jdx, rj);
          int o[3] = {0, 0, 0};
          for (int d = 0; d < dim; ++d)
            o[d] = int(idx[d] - jdx[d]);
          const int q = offset_code(dim, o);
          for (int d = 0; d < dim; ++d)
            out.coupling_cji.push_back(rj.c[q][d]);
        }
      }
    }
  }


  void decode_simd_pattern(uint32_t n_rows,
                           uint32_t n_internal,
                           uint32_t simd_length,
                           const uint64_t *row_starts,
                           const uint32_t *column_indices,
                           std::vector<uint64_t> &csr_row_ptr,
                           std::vector<uint32_t> &csr_cols)
  {
    if (simd_length == 0 || n_internal % simd_length != 0 || n_internal > n_rows)
      throw std::runtime_error("decode_simd_pattern: invalid n_internal / simd_length");
    rj_offline_desc d{};
    d.simd_length = simd_length;
    d.n_internal = n_internal;
    d.row_starts = row_starts;
    d.column_indices = column_indices;
    csr_row_ptr.assign(size_t(n_rows) + 1, 0);
    for (uint32_t i = 0; i < n_rows; ++i)
      csr_row_ptr[i + 1] = csr_row_ptr[i] + host_row_length(d, i);
    csr_cols.resize(csr_row_ptr[n_rows]);
    for (uint32_t i = 0; i < n_rows; ++i) {
      const uint32_t len = host_row_length(d, i);
      for (uint32_t c = 0; c < len; ++c)
        csr_cols[csr_row_ptr[i] + c] = host_column(d, i, c);
    }
  }
} // namespace rjb
