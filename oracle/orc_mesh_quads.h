// ORACLE -- TEST INFRASTRUCTURE ONLY (see orc_common.h).
//
// Offline data assembly for continuous Q1 elements on a general (unstructured, MAPPED) quadrilateral
// mesh in 2D: vertices with arbitrary coordinates, cells as 4 vertex indices in lexicographic
// local order (v0 = (0,0), v1 = (1,0), v2 = (0,1), v3 = (1,1) of the reference cell, deal.II's
// ordering), boundary faces as (vertex a, vertex b, boundary id).  The cells are images of the unit
// square under the bilinear (MappingQ1) map, so the Jacobian varies over a cell:
//   m_ij  = sum_K sum_q  phi_i phi_j |det J| w_q                 offline_data.template.h:560-578
//   c_ij  = sum_K sum_q  phi_i (J^{-T} grad phi_j) |det J| w_q   (QGauss(2), discretization.template.h:161-165)
//   boundary map: per boundary face QGauss(3), normal integrals, merge rule, 2D no_slip corner fix
//                                                                offline_data.template.h:1173-1364
//   coupling boundary pairs                                      offline_data.template.h:1369-1462
// This is what the reference assembles for its non-Cartesian geometries (the rounded corner of
// geometry_step.h:36-128, the cylinder meshes); DoF i = vertex i.  The product consumes the
// resulting arrays through its stencil importer (ryujin_b200/stencil_io.py).
#pragma once

#include "orc_mesh.h"

namespace orc
{
  struct QuadMesh {
    std::vector<Vec<2>> vertices;
    std::vector<std::array<unsigned int, 4>> cells;
    struct BoundaryFace {
      unsigned int a, b;
      int id;
      bool flip; /* outward normal of a -> b is (dy, -dx) / |ab|; reversed if set */
    };
    std::vector<BoundaryFace> boundary_faces;
  };

  inline OfflineData<2> assemble_offline_data_quads(const QuadMesh &mesh)
  {
    constexpr int dim = 2, nv = 4;
    const unsigned int n = (unsigned int)mesh.vertices.size();
    OfflineData<dim> od;
    od.n_owned = n;

    /* vertex -> cells (in cell order: contributions are added in the order of a serial cell loop) */
    std::vector<unsigned int> vc_ptr(n + 1, 0);
    for (const auto &c : mesh.cells)
      for (int v = 0; v < nv; ++v)
        vc_ptr[c[v] + 1]++;
    for (unsigned int i = 0; i < n; ++i)
      vc_ptr[i + 1] += vc_ptr[i];
    std::vector<unsigned int> vc(vc_ptr[n]);
    {
      std::vector<unsigned int> fill(vc_ptr.begin(), vc_ptr.end() - 1);
      for (unsigned int k = 0; k < mesh.cells.size(); ++k)
        for (int v = 0; v < nv; ++v)
          vc[fill[mesh.cells[k][v]]++] = k;
    }

    /* sparsity: diagonal first, then ascending */
    od.row_ptr.assign(n + 1, 0);
    std::vector<std::vector<unsigned int>> rows(n);
#pragma omp parallel for schedule(dynamic, 1024)
    for (unsigned int i = 0; i < n; ++i) {
      std::vector<unsigned int> nb;
      for (unsigned int q = vc_ptr[i]; q < vc_ptr[i + 1]; ++q)
        for (int v = 0; v < nv; ++v)
          nb.push_back(mesh.cells[vc[q]][v]);
      std::sort(nb.begin(), nb.end());
      nb.erase(std::unique(nb.begin(), nb.end()), nb.end());
      rows[i].push_back(i);
      for (unsigned int j : nb)
        if (j != i)
          rows[i].push_back(j);
    }
    for (unsigned int i = 0; i < n; ++i)
      od.row_ptr[i + 1] = od.row_ptr[i] + (unsigned int)rows[i].size();
    od.cols.resize(od.row_ptr[n]);
    for (unsigned int i = 0; i < n; ++i)
      std::copy(rows[i].begin(), rows[i].end(), od.cols.begin() + od.row_ptr[i]);
    rows.clear();
    od.build_transposed();

    const std::size_t nnz = od.cols.size();
    od.mij.assign(nnz, 0.);
    od.cij.assign(nnz * dim, 0.);

    /* QGauss(2) on the unit square, shape values and reference gradients */
    std::vector<double> qx, qw;
    gauss_01(2, qx, qw);
    double val[4][nv], gref[4][nv][dim], wq[4];
    for (int q = 0; q < 4; ++q) {
      const double x = qx[q % 2], y = qx[q / 2];
      wq[q] = qw[q % 2] * qw[q / 2];
      for (int v = 0; v < nv; ++v) {
        const double sx = (v & 1) ? x : 1. - x, sy = (v & 2) ? y : 1. - y;
        val[q][v] = sx * sy;
        gref[q][v][0] = ((v & 1) ? 1. : -1.) * sy;
        gref[q][v][1] = sx * ((v & 2) ? 1. : -1.);
      }
    }

    auto cell_matrices = [&](const std::array<unsigned int, 4> &c, double (&mass)[nv][nv], double (&cc)[dim][nv][nv],
                             double &measure) {
      for (int a = 0; a < nv; ++a)
        for (int b = 0; b < nv; ++b) {
          mass[a][b] = 0.;
          cc[0][a][b] = cc[1][a][b] = 0.;
        }
      measure = 0.;
      for (int q = 0; q < 4; ++q) {
        /* J = d x / d xi */
        double J[2][2] = {{0., 0.}, {0., 0.}};
        for (int v = 0; v < nv; ++v)
          for (int r = 0; r < 2; ++r)
            for (int s = 0; s < 2; ++s)
              J[r][s] += mesh.vertices[c[v]][r] * gref[q][v][s];
        const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
        if (!(det > 0.))
          throw std::runtime_error("assemble_offline_data_quads: inverted or degenerate cell");
        const double JxW = det * wq[q];
        measure += JxW;
        const double det_inverse = 1. / det;
        /* J^{-T} */
        const double Ji[2][2] = {{J[1][1] * det_inverse, -J[1][0] * det_inverse},
                                 {-J[0][1] * det_inverse, J[0][0] * det_inverse}};
        double grad[nv][dim];
        for (int v = 0; v < nv; ++v)
          for (int r = 0; r < 2; ++r)
            grad[v][r] = Ji[r][0] * gref[q][v][0] + Ji[r][1] * gref[q][v][1];
        for (int j = 0; j < nv; ++j) {
          const double value_JxW = val[q][j] * JxW;
          for (int i = 0; i < nv; ++i) {
            mass[i][j] += val[q][i] * value_JxW;
            for (int d = 0; d < dim; ++d)
              cc[d][i][j] += val[q][i] * (grad[j][d] * JxW);
          }
        }
      }
    };

    auto find = [&](unsigned int i, unsigned int j) -> std::size_t {
      for (unsigned int e = od.row_ptr[i]; e < od.row_ptr[i + 1]; ++e)
        if (od.cols[e] == j)
          return e;
      throw std::runtime_error("entry not found");
    };

    od.measure_of_omega = 0.;
    for (const auto &c : mesh.cells) {
      double mass[nv][nv], cc[dim][nv][nv], measure;
      cell_matrices(c, mass, cc, measure);
      od.measure_of_omega += measure;
    }
#pragma omp parallel for schedule(dynamic, 1024)
    for (unsigned int i = 0; i < n; ++i)
      for (unsigned int q = vc_ptr[i]; q < vc_ptr[i + 1]; ++q) {
        const auto &c = mesh.cells[vc[q]];
        double mass[nv][nv], cc[dim][nv][nv], measure;
        cell_matrices(c, mass, cc, measure);
        int a = 0;
        while (c[a] != i)
          ++a;
        for (int b = 0; b < nv; ++b) {
          const std::size_t e = find(i, c[b]);
          od.mij[e] += mass[a][b];
          for (int d = 0; d < dim; ++d)
            od.cij[e * dim + d] += cc[d][a][b];
        }
      }

    od.mi.assign(n, 0.);
    od.mi_inv.assign(n, 0.);
#pragma omp parallel for schedule(static)
    for (unsigned int i = 0; i < n; ++i) {
      double s = 0.;
      for (unsigned int e = od.row_ptr[i]; e < od.row_ptr[i + 1]; ++e)
        s += od.mij[e];
      od.mi[i] = s;
      od.mi_inv[i] = 1. / s;
    }

    /* boundary map.  A face is the straight segment a -> b (MappingQ1); faces arrive in deal.II's
     * visiting order (cells ascending, faces x-, x+, y-, y+, vertices ascending): the order of a
     * DoF's entries is the order in which its boundary conditions are applied. */
    struct BData {
      Vec<dim> normal;
      double boundary_mass;
      int id;
      Vec<dim> position;
    };
    std::multimap<unsigned int, BData> preliminary_map;
    std::set<unsigned int> boundary_indices;
    std::vector<double> fx, fw;
    gauss_01(3, fx, fw);
    for (const auto &f : mesh.boundary_faces) {
      if (f.id == Boundary::periodic)
        continue;
      const Vec<dim> &A = mesh.vertices[f.a], &B = mesh.vertices[f.b];
      const double tx = B[0] - A[0], ty = B[1] - A[1];
      const double length = std::sqrt(tx * tx + ty * ty);
      const double sign = f.flip ? -1. : 1.;
      const Vec<dim> n_out = {sign * ty / length, -sign * tx / length};
      for (int end = 0; end < 2; ++end) {
        double boundary_mass = 0.;
        Vec<dim> normal{};
        for (int q = 0; q < 3; ++q) {
          const double phi = end ? fx[q] : 1. - fx[q];
          const double JxW = length * fw[q];
          boundary_mass += phi * JxW;
          for (int d = 0; d < dim; ++d)
            normal[d] += phi * n_out[d] * JxW;
        }
        const unsigned int index = end ? f.b : f.a;
        boundary_indices.insert(index);
        if (od.row_length(index) == 1)
          continue;
        preliminary_map.insert({index, {normal, boundary_mass, f.id, mesh.vertices[index]}});
      }
    }

    std::multimap<unsigned int, BData> filtered_map;
    for (auto entry : preliminary_map) {
      bool inserted = false;
      const auto range = filtered_map.equal_range(entry.first);
      for (auto it = range.first; it != range.second; ++it) {
        auto &nw = entry.second;
        auto &old = it->second;
        if (old.id != nw.id)
          continue;
        if (dot<dim>(old.normal, nw.normal) / norm<dim>(old.normal) / norm<dim>(nw.normal) > 0.50) {
          for (int d = 0; d < dim; ++d)
            old.normal[d] += nw.normal[d];
          old.boundary_mass += nw.boundary_mass;
          inserted = true;
          continue;
        } else {
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
