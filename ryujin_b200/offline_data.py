"""OfflineData: stencil data of one rank (mirror of ryujin's OfflineData accessors,
source/offline_data.h:109-264) plus the Cartesian Q1 mesh generator that stands
in for Discretization + OfflineData::prepare() on uniform meshes.
"""
import ctypes as C
import enum

import numpy as np

from . import _lib
from ._lib import check, lib


class Boundary(enum.IntEnum):
    """source/discretization.h:48-123"""

    do_nothing = 0
    periodic = 1
    slip = 2
    no_slip = 3
    dirichlet = 4
    dynamic = 5
    dirichlet_momentum = 6


class CartesianMesh:
    """Uniform continuous-Q1 mesh on a box of cells with an optional cell mask.

    ``boundary_ids`` lists the ids of the box faces in deal.II face order
    (x-, x+, y-, y+, z-, z+) -- "left, right, bottom, top, back, front" of
    source/geometry_rectangular_domain.h.  ``masked_boundary_id`` is used for
    faces that border an inactive cell (e.g. the walls of the forward facing
    step).
    """

    def __init__(self, n_cells, origin, extent, boundary_ids, cell_mask=None, masked_boundary_id=Boundary.slip,
                 partition=None):
        """partition: ranks per direction, e.g. (2, 2, 2) -- blocks of the node grid as the reference's
        p4est partition hands them out (offline_data.template.h:353-360); None = slabs along the last axis"""
        self.dim = len(n_cells)
        self.partition = [0, 0, 0] if partition is None else [int(p) for p in partition] + [1] * (3 - len(partition))
        assert 1 <= self.dim <= 3
        assert len(origin) == self.dim and len(extent) == self.dim
        assert len(boundary_ids) == 2 * self.dim
        self.n_cells = [int(n) for n in n_cells]
        self.origin = [float(x) for x in origin]
        self.h = [float(e) / n for e, n in zip(extent, self.n_cells)]
        self.boundary_ids = [int(b) for b in boundary_ids]
        self.masked_boundary_id = int(masked_boundary_id)
        self.cell_mask = None
        if cell_mask is not None:
            m = np.ascontiguousarray(cell_mask, dtype=np.uint8).ravel()
            assert m.size == int(np.prod(self.n_cells))
            self.cell_mask = m

    @classmethod
    def rectangular_domain(cls, bottom_left, top_right, n_cells, boundary_ids, partition=None):
        """geometry "rectangular domain" with 2^refinement x ... subdivisions folded into n_cells"""
        extent = [b - a for a, b in zip(bottom_left, top_right)]
        return cls(n_cells, bottom_left, extent, boundary_ids, partition=partition)

    @classmethod
    def forward_facing_step(cls, refinement, length=3.0, height=1.0, step_position=0.6, step_height=0.2,
                            left=Boundary.dirichlet, right=Boundary.do_nothing, walls=Boundary.slip):
        """Topology of source/geometry_step.h as a masked uniform grid: channel
        [0,length] x [0,height] minus the step [step_position,length] x [0,step_height]."""
        base = 5 * 2**refinement  # cells per unit length: 0.2 and 0.6 are multiples of 1/5
        nx, ny = int(round(length * base)), int(round(height * base))
        sx, sy = int(round(step_position * base)), int(round(step_height * base))
        mask = np.ones((ny, nx), dtype=np.uint8)
        mask[:sy, sx:] = 0
        return cls([nx, ny], [0.0, 0.0], [length, height], [left, right, walls, walls], cell_mask=mask,
                   masked_boundary_id=walls)

    def desc(self):
        d = _lib.MeshDesc()
        d.dim = self.dim
        for q in range(self.dim):
            d.n_cells[q] = self.n_cells[q]
            d.origin[q] = self.origin[q]
            d.h[q] = self.h[q]
        for f, b in enumerate(self.boundary_ids):
            d.box_boundary_ids[f] = b
        d.masked_boundary_id = self.masked_boundary_id
        for q in range(3):
            d.partition[q] = self.partition[q]
        if self.cell_mask is not None:
            d.cell_mask = self.cell_mask.ctypes.data_as(_lib.u8p)
        return d

    @property
    def n_nodes(self):
        return [n + 1 for n in self.n_cells]


def _view(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(n,)).view(dtype)


class HostOfflineData:
    """Host arrays of one rank in local numbering (rj_host_offline)."""

    def __init__(self, mesh, rank=0, n_ranks=1):
        self.mesh = mesh
        self.rank, self.n_ranks = rank, n_ranks
        self._mesh_desc = mesh.desc()
        h = C.c_void_p()
        check(lib.rj_mesh_generate(C.byref(h), C.byref(self._mesh_desc), rank, n_ranks))
        self._h = h
        self.desc = _lib.OfflineDesc()
        check(lib.rj_host_offline_desc(self._h, C.byref(self.desc)))
        d = self.desc
        self.dim = d.dim
        self.n_export, self.n_owned, self.n_relevant = d.n_export, d.n_owned, d.n_relevant
        self.row_starts = _view(d.row_starts, d.n_owned + 1, np.uint64)
        nnz = int(self.row_starts[-1])
        self.nnz = nnz
        self.cols = _view(d.column_indices, nnz, np.uint32)
        self.mij = _view(d.mass_matrix, nnz, np.float64)
        self.cij = _view(d.cij_matrix, nnz * d.dim, np.float64).reshape(nnz, d.dim)
        self.mi = _view(d.lumped_mass_matrix, d.n_relevant, np.float64)
        self.measure_of_omega = d.measure_of_omega
        nb = d.n_boundary
        self.boundary_dof = _view(d.boundary_dof, nb, np.uint32)
        self.boundary_normal = _view(d.boundary_normal, nb * d.dim, np.float64).reshape(nb, d.dim)
        self.boundary_id = _view(d.boundary_id, nb, np.int32)
        self.boundary_position = _view(d.boundary_position, nb * d.dim, np.float64).reshape(nb, d.dim)
        npairs = d.n_coupling_pairs
        self.coupling_pairs = _view(d.coupling_pairs, 3 * npairs, np.uint32).reshape(npairs, 3)
        self.coupling_cji = _view(d.coupling_cji, npairs * d.dim, np.float64).reshape(npairs, d.dim)
        nn = d.n_neighbors
        self.neighbor_rank = _view(d.neighbor_rank, nn, np.int32)
        self.ghost_count = _view(d.ghost_count, nn, np.uint32)
        self.send_offsets = _view(d.send_offsets, nn + 1, np.uint32) if nn else np.zeros(1, dtype=np.uint32)
        self.send_indices = _view(d.send_indices, int(self.send_offsets[-1]), np.uint32)
        self.positions = _view(d.positions, d.n_owned * d.dim, np.float64).reshape(d.n_owned, d.dim)
        self.n_global_dofs = int(lib.rj_host_offline_n_global_dofs(self._h))
        g = np.zeros(d.n_relevant, dtype=np.uint64)
        check(lib.rj_host_offline_global_indices(self._h, g.ctypes.data_as(_lib.u64p)))
        self.global_index = g

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.rj_host_offline_destroy(h)
            self._h = None

    def row_length(self, i):
        return int(self.row_starts[i + 1] - self.row_starts[i])


class OfflineData:
    """Device resident stencil data (rj_offline): sliced-ELL matrices, lumped
    mass, boundary map, coupling pairs and halo lists on one GPU."""

    def __init__(self, host, device=0, host_layout_only=False):
        self.host = host
        self.device = device
        self.host_layout_only = host_layout_only
        h = C.c_void_p()
        if host_layout_only:
            check(lib.rj_offline_create_host_layout(C.byref(h), C.byref(host.desc)))
        else:
            check(lib.rj_offline_create(C.byref(h), C.byref(host.desc), device))
        self._h = h
        sizes = (C.c_uint64 * 4)()
        check(lib.rj_offline_layout_sizes(self._h, sizes))
        self.n_slices, self.n_entries_owned, self.n_entries_total, self.ld = (int(x) for x in sizes)
        self.dim = host.dim

    @classmethod
    def from_mesh(cls, mesh, rank=0, n_ranks=1, device=0):
        return cls(HostOfflineData(mesh, rank, n_ranks), device)

    @classmethod
    def from_stencil(cls, stencil, device=0, host_layout_only=False):
        """stencil: a stencil_io.StencilData or the path of an RJSTENCIL1 dump (arrays of a mesh
        that was assembled elsewhere, e.g. by a ryujin build with deal.II)"""
        from .stencil_io import StencilData

        if isinstance(stencil, str):
            stencil = StencilData.load(stencil)
        return cls(stencil, device, host_layout_only)

    # accessors in the reference's vocabulary (offline_data.h:136-169)
    @property
    def n_export_indices(self):
        return self.host.n_export

    @property
    def n_locally_owned(self):
        return self.host.n_owned

    @property
    def n_locally_relevant(self):
        return self.host.n_relevant

    @property
    def measure_of_omega(self):
        return self.host.measure_of_omega

    def layout(self):
        """(slice_offsets, row_length, cols, transposed) of the device layout"""
        so = np.zeros(self.n_slices + 1, dtype=np.uint64)
        rl = np.zeros(self.host.n_owned, dtype=np.uint32)
        cols = np.zeros(self.n_entries_owned, dtype=np.uint32)
        tr = np.zeros(self.n_entries_owned, dtype=np.uint32)
        check(lib.rj_offline_layout_get(self._h, so.ctypes.data_as(_lib.u64p), rl.ctypes.data_as(_lib.u32p),
                                        cols.ctypes.data_as(_lib.u32p), tr.ctypes.data_as(_lib.u32p)))
        return so, rl, cols, tr

    def exchange_lists(self):
        """(neighbors, send_rows, mat_send_entries); neighbors: list of dicts"""
        sizes = (C.c_uint64 * 3)()
        check(lib.rj_offline_exchange_sizes(self._h, sizes))
        nn, ns, nm = (int(x) for x in sizes)
        info = np.zeros(9 * max(nn, 1), dtype=np.uint32)
        rows = np.zeros(max(ns, 1), dtype=np.uint32)
        ents = np.zeros(max(nm, 1), dtype=np.uint32)
        check(lib.rj_offline_exchange_get(self._h, info.ctypes.data_as(_lib.u32p), rows.ctypes.data_as(_lib.u32p),
                                          ents.ctypes.data_as(_lib.u32p)))
        keys = ("rank", "ghost_start", "ghost_count", "send_offset", "send_count", "mat_send_offset",
                "mat_send_count", "mat_recv_offset", "mat_recv_count")
        neighbors = [dict(zip(keys, (int(v) for v in info[9 * p : 9 * p + 9]))) for p in range(nn)]
        return neighbors, rows[:ns], ents[:nm]

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.rj_offline_destroy(h)
            self._h = None
