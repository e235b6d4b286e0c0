// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_common.h).
//
// Independent (quadrature based) restatement of the offline data assembly for
// continuous Q1 elements on a Cartesian cell grid with an optional cell mask
// (rectangular domain; forward facing step as a masked grid):
//   m_ij, c_ij cell integrals with QGauss(2)  <- offline_data.template.h:560-578
//                                                 discretization.template.h:161-165
//   lumped mass, |Omega|                        <- offline_data.template.h:756-806
//   boundary map (merge / 2D no_slip fix)       <- offline_data.template.h:1173-1364
//   coupling boundary pairs                     <- offline_data.template.h:1369-1462
// DoFs are numbered lexicographically over the active nodes; the product's
// closed-form generator is validated against this assembly in tests/.
#pragma once

#include "orc_module.h"

#include <map>
#include <set>

namespace orc
{
  template <int dim>
  struct CartesianMesh {
    std::array<unsigned int, dim> n_cells; /* cells per direction */
    Vec<dim> origin, h;
    std::vector<unsigned char> cell_mask;  /* empty: all active */
    /* boundary id of a face: (cell multi index, face number 0..2*dim-1) */
    std::function<int(const std::array<unsigned int, dim> &, int)> boundary_id;

    /* filled by enumerate(): */
    std::vector<unsigned int> node_to_dof; /* lexicographic node -> dof or ~0u */
    std::vector<std::array<unsigned int, dim>> dof_to_node;
    unsigned int n_dofs = 0;

    std::size_t n_nodes_total() const
    {
      std::size_t n = 1;
      for (int d = 0; d < dim; ++d)
        n *= n_cells[d] + 1;
      return n;
    }

    std::size_t node_index(const std::array<unsigned int, dim> &idx) const
    {
      std::size_t r = 0;
      for (int d = dim - 1; d >= 0; --d)
        r = r * (n_cells[d] + 1) + idx[d];
      return r;
    }

    std::size_t cell_index(const std::array<unsigned int, dim> &idx) const
    {
      std::size_t r = 0;
      for (int d = dim - 1; d >= 0; --d)
        r = r * n_cells[d] + idx[d];
      return r;
    }

    bool cell_active(const std::array<int, dim> &idx) const
    {
      std::array<unsigned int, dim> u;
      for (int d = 0; d < dim; ++d) {
        if (idx[d] < 0 || idx[d] >= int(n_cells[d]))
          return false;
        u[d] = idx[d];
      }
      return cell_mask.empty() ? true : cell_mask[cell_index(u)] != 0;
    }

    template <typename F>
    void for_each_cell(const F &f) const
    {
      std::array<unsigned int, dim> idx{};
      std::size_t total = 1;
      for (int d = 0; d < dim; ++d)
        total *= n_cells[d];
      for (std::size_t c = 0; c < total; ++c) {
        std::size_t r = c;
        for (int d = 0; d < dim; ++d) {
          idx[d] = r % n_cells[d];
          r /= n_cells[d];
        }
        std::array<int, dim> s;
        for (int d = 0; d < dim; ++d)
          s[d] = idx[d];
        if (cell_active(s))
          f(idx);
      }
    }

    /* local vertex v of the cell, lexicographic (x fastest) */
    std::array<unsigned int, dim>
    vertex(const std::array<unsigned int, dim> &cell, int v) const
    {
      std::array<unsigned int, dim> n;
      for (int d = 0; d < dim; ++d)
        n[d] = cell[d] + ((v >> d) & 1);
      return n;
    }

    void enumerate()
    {
      node_to_dof.assign(n_nodes_total(), ~0u);
      std::vector<unsigned char> touched(n_nodes_total(), 0);
      for_each_cell([&](const std::array<unsigned int, dim> &cell) {
        for (int v = 0; v < (1 << dim); ++v)
          touched[node_index(vertex(cell, v))] = 1;
      });
      n_dofs = 0;
      dof_to_node.clear();
      for (std::size_t n = 0; n < touched.size(); ++n)
        if (touched[n]) {
          node_to_dof[n] = n_dofs++;
          std::array<unsigned int, dim> idx;
          std::size_t r = n;
          for (int d = 0; d < dim; ++d) {
            idx[d] = r % (n_cells[d] + 1);
            r /= (n_cells[d] + 1);
          }
          dof_to_node.push_back(idx);
        }
    }

    Vec<dim> position(unsigned int dof) const
    {
      Vec<dim> x;
      for (int d = 0; d < dim; ++d)
        x[d] = origin[d] + dof_to_node[dof][d] * h[d];
      return x;
    }
  };


  /* 1D Gauss-Legendre points/weights on [0,1] */
  inline void gauss_01(int n, std::vector<double> &x, std::vector<double> &w)
  {
    x.resize(n);
    w.resize(n);
    if (n == 2) {
      const double a = 0.5 / std::sqrt(3.);
      x = {0.5 - a, 0.5 + a};
      w = {0.5, 0.5};
    } else if (n == 3) {
      const double a = 0.5 * std::sqrt(3. / 5.);
      x = {0.5 - a, 0.5, 0.5 + a};
      w = {5. / 18., 8. / 18., 5. / 18.};
    } else
      throw std::runtime_error("gauss_01: unsupported order");
  }


  template <int dim>
  OfflineData<dim> assemble_offline_data(CartesianMesh<dim> &mesh)
  {
    mesh.enumerate();
    constexpr int nv = 1 << dim;

    OfflineData<dim> od;
    od.n_owned = mesh.n_dofs;

    /* sparsity: row i couples to the vertices of the active cells around node i.  Rows are
     * independent, so they are built in parallel (the 8M-DoF meshes of bench.py's CPU arm);
     * diagonal first, then ascending -- the order the serial set-based construction produced */
    auto adjacent_dofs = [&](unsigned int i, unsigned int *out) -> unsigned int {
      const auto &node = mesh.dof_to_node[i];
      unsigned int n = 0;
      for (int o = 0; o < nv; ++o) {
        std::array<int, dim> cell_s;
        for (int d = 0; d < dim; ++d)
          cell_s[d] = int(node[d]) - ((o >> d) & 1);
        if (!mesh.cell_active(cell_s))
          continue;
        std::array<unsigned int, dim> cell;
        for (int d = 0; d < dim; ++d)
          cell[d] = cell_s[d];
        for (int v = 0; v < nv; ++v)
          out[n++] = mesh.node_to_dof[mesh.node_index(mesh.vertex(cell, v))];
      }
      std::sort(out, out + n);
      return (unsigned int)(std::unique(out, out + n) - out);
    };

    od.row_ptr.assign(mesh.n_dofs + 1, 0);
#pragma omp parallel for schedule(static)
    for (unsigned int i = 0; i < mesh.n_dofs; ++i) {
      unsigned int buffer[nv * nv];
      od.row_ptr[i + 1] = adjacent_dofs(i, buffer);
    }
    for (unsigned int i = 0; i < mesh.n_dofs; ++i)
      od.row_ptr[i + 1] += od.row_ptr[i];
    od.cols.resize(od.row_ptr.back());
#pragma omp parallel for schedule(static)
    for (unsigned int i = 0; i < mesh.n_dofs; ++i) {
      unsigned int buffer[nv * nv];
      const unsigned int n = adjacent_dofs(i, buffer);
      unsigned int e = od.row_ptr[i];
      od.cols[e++] = i;
      for (unsigned int q = 0; q < n; ++q)
        if (buffer[q] != i)
          od.cols[e++] = buffer[q];
    }
    od.build_transposed();

    const std::size_t nnz = od.cols.size();
    od.mij.assign(nnz, 0.);
    od.cij.assign(nnz * dim, 0.);

    auto find = [&](unsigned int i, unsigned int j) -> std::size_t {
      for (unsigned int e = od.row_ptr[i]; e < od.row_ptr[i + 1]; ++e)
        if (od.cols[e] == j)
          return e;
      throw std::runtime_error("entry not found");
    };

    /* reference cell integrals with QGauss(2) */
    std::vector<double> qx, qw;
    gauss_01(2, qx, qw);
    const int nq1 = 2;
    int nq = 1;
    for (int d = 0; d < dim; ++d)
      nq *= nq1;

    double cell_measure = 1.;
    for (int d = 0; d < dim; ++d)
      cell_measure *= mesh.h[d];

    double cell_mass[nv][nv];
    double cell_c[dim][nv][nv];
    for (int a = 0; a < nv; ++a)
      for (int b = 0; b < nv; ++b) {
        cell_mass[a][b] = 0.;
        for (int d = 0; d < dim; ++d)
          cell_c[d][a][b] = 0.;
      }
    double measure_per_cell = 0.;

    for (int q = 0; q < nq; ++q) {
      double xi[dim], JxW = cell_measure;
      {
        int r = q;
        for (int d = 0; d < dim; ++d) {
          xi[d] = qx[r % nq1];
          JxW *= qw[r % nq1];
          r /= nq1;
        }
      }
      measure_per_cell += JxW;
      double val[nv], grad[nv][dim];
      for (int v = 0; v < nv; ++v) {
        val[v] = 1.;
        for (int d = 0; d < dim; ++d)
          val[v] *= ((v >> d) & 1) ? xi[d] : 1. - xi[d];
        for (int g = 0; g < dim; ++g) {
          double t = 1.;
          for (int d = 0; d < dim; ++d) {
            if (d == g)
              t *= (((v >> d) & 1) ? 1. : -1.) / mesh.h[d];
            else
              t *= ((v >> d) & 1) ? xi[d] : 1. - xi[d];
          }
          grad[v][g] = t;
        }
      }
      for (int j = 0; j < nv; ++j) {
        const double value_JxW = val[j] * JxW;
        for (int i = 0; i < nv; ++i) {
          cell_mass[i][j] += val[i] * value_JxW;
          for (int d = 0; d < dim; ++d)
            cell_c[d][i][j] += val[i] * (grad[j][d] * JxW);
        }
      }
    }

    od.measure_of_omega = 0.;
    mesh.for_each_cell([&](const std::array<unsigned int, dim> &) { od.measure_of_omega += measure_per_cell; });
    /* entry (i, j) = sum over the cells that contain both nodes, added in lexicographic cell order
     * (what a serial loop over the cells does); row i only receives contributions from the cells
     * around node i, so the rows are assembled independently and in parallel */
#pragma omp parallel for schedule(static)
    for (unsigned int i = 0; i < mesh.n_dofs; ++i) {
      const auto &node = mesh.dof_to_node[i];
      for (int o = nv - 1; o >= 0; --o) { /* offset bits descending = cell index ascending */
        std::array<int, dim> cell_s;
        int a = 0;
        for (int d = 0; d < dim; ++d) {
          cell_s[d] = int(node[d]) - ((o >> d) & 1);
          a |= ((o >> d) & 1) << d; /* local vertex of node i in that cell */
        }
        if (!mesh.cell_active(cell_s))
          continue;
        std::array<unsigned int, dim> cell;
        for (int d = 0; d < dim; ++d)
          cell[d] = cell_s[d];
        for (int b = 0; b < nv; ++b) {
          const std::size_t e = find(i, mesh.node_to_dof[mesh.node_index(mesh.vertex(cell, b))]);
          od.mij[e] += cell_mass[a][b];
          for (int d = 0; d < dim; ++d)
            od.cij[e * dim + d] += cell_c[d][a][b];
        }
      }
    }

    od.mi.assign(mesh.n_dofs, 0.);
    od.mi_inv.assign(mesh.n_dofs, 0.);
#pragma omp parallel for schedule(static)
    for (unsigned int i = 0; i < mesh.n_dofs; ++i) {
      double s = 0.;
      for (unsigned int e = od.row_ptr[i]; e < od.row_ptr[i + 1]; ++e)
        s += od.mij[e];
      od.mi[i] = s;
      od.mi_inv[i] = 1. / s;
    }

    /* boundary map */
    struct BData {
      Vec<dim> normal;
      double normal_mass, boundary_mass;
      int id;
      Vec<dim> position;
    };
    std::multimap<unsigned int, BData> preliminary_map;
    std::set<unsigned int> boundary_indices;

    std::vector<double> fx, fw;
    gauss_01(3, fx, fw);

    mesh.for_each_cell([&](const std::array<unsigned int, dim> &cell) {
      for (int f = 0; f < 2 * dim; ++f) {
        const int fd = f / 2, side = f % 2;
        std::array<int, dim> nb;
        for (int d = 0; d < dim; ++d)
          nb[d] = cell[d];
        nb[fd] += side ? 1 : -1;
        if (mesh.cell_active(nb))
          continue;

        const int id = mesh.boundary_id(cell, f);
        if (id == Boundary::periodic)
          continue;

        double face_measure = 1.;
        for (int d = 0; d < dim; ++d)
          if (d != fd)
            face_measure *= mesh.h[d];

        for (int v = 0; v < nv; ++v) {
          if (((v >> fd) & 1) != side)
            continue; /* no support on face */

          double boundary_mass = 0.;
          Vec<dim> normal{};
          int nqf = 1;
          for (int d = 0; d < dim - 1; ++d)
            nqf *= 3;
          for (int q = 0; q < nqf; ++q) {
            double JxW = face_measure, phi = 1.;
            int r = q;
            for (int d = 0; d < dim; ++d) {
              if (d == fd)
                continue;
              const double x = fx[r % 3];
              JxW *= fw[r % 3];
              r /= 3;
              phi *= ((v >> d) & 1) ? x : 1. - x;
            }
            boundary_mass += phi * JxW;
            normal[fd] += phi * (side ? 1. : -1.) * JxW;
          }

          const unsigned int index =
              mesh.node_to_dof[mesh.node_index(mesh.vertex(cell, v))];
          boundary_indices.insert(index);
          if (od.row_length(index) == 1)
            continue;
          preliminary_map.insert(
              {index,
               {normal, boundary_mass, boundary_mass, id,
                mesh.position(index)}});
        }
      }
    });

    std::multimap<unsigned int, BData> filtered_map;
    for (auto entry : preliminary_map) {
      bool inserted = false;
      const auto range = filtered_map.equal_range(entry.first);
      for (auto it = range.first; it != range.second; ++it) {
        auto &nw = entry.second;
        auto &old = it->second;
        if (old.id != nw.id)
          continue;
        if (dot<dim>(old.normal, nw.normal) / norm<dim>(old.normal) /
                norm<dim>(nw.normal) >
            0.50) {
          for (int d = 0; d < dim; ++d)
            old.normal[d] += nw.normal[d];
          old.boundary_mass += nw.boundary_mass;
          inserted = true;
          continue;
        } else if constexpr (dim == 2) {
          if (nw.id == Boundary::slip) {
            nw.id = Boundary::no_slip;
            old.id = Boundary::no_slip;
          }
        }
      }
      if (!inserted)
        filtered_map.insert(entry);
    }

    for (const auto &it : filtered_map) {
      const auto &b = it.second;
      BoundaryDescription<dim> bd;
      bd.i = it.first;
      bd.normal_mass = norm<dim>(b.normal) + eps_double;
      for (int d = 0; d < dim; ++d)
        bd.normal[d] = b.normal[d] / bd.normal_mass;
      bd.boundary_mass = b.boundary_mass;
      bd.id = b.id;
      bd.position = b.position;
      od.boundary_map.push_back(bd);
    }

    /* coupling boundary pairs */
    for (const unsigned int i : boundary_indices) {
      const unsigned int row_length = od.row_length(i);
      if (row_length == 1)
        continue;
      for (unsigned int col_idx = 1; col_idx < row_length; ++col_idx) {
        const unsigned int j = od.cols[od.row_ptr[i] + col_idx];
        if (boundary_indices.count(j) != 0)
          od.coupling_boundary_pairs.push_back({i, col_idx, j});
      }
    }

    return od;
  }
} // namespace orc
