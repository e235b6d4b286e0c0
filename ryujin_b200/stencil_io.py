"""Stencil importer: OfflineData from arrays that were not produced by the
built-in uniform-grid generator -- in memory (`StencilData.from_arrays`) or from a dump file that a
ryujin build with deal.II writes once per mesh (`StencilData.load`; INTEGRATION.md shows the C++
writer).  The arrays are the reference's own, in LOCAL indexing: the read-only accessors of
OfflineData (source/offline_data.h:109-264) in either plain CSR (`simd_length = 1`) or the
reference's SparsityPatternSIMD / SparseMatrixSIMD host layout (sparse_matrix_simd.h:40-61).

File format ("RJSTENCIL1"), little endian:
    line 1   b"RJSTENCIL1\\n"
    line 2   one JSON object + "\\n": {"scalars": {...}, "arrays": [[name, dtype, [shape...]], ...]}
    then     the arrays in that order, each padded to a multiple of 8 bytes
dtypes: "u4", "u8", "i4", "f8".
"""
import ctypes as C
import json

import numpy as np

from . import _lib

MAGIC = b"RJSTENCIL1\n"

# name -> (dtype, required)
ARRAYS = {
    "row_starts": ("u8", True), "column_indices": ("u4", True), "mass_matrix": ("f8", True), "cij_matrix": ("f8", True),
    "lumped_mass_matrix": ("f8", True), "boundary_dof": ("u4", False), "boundary_normal": ("f8", False),
    "boundary_id": ("i4", False), "boundary_position": ("f8", False), "coupling_pairs": ("u4", False),
    "coupling_cji": ("f8", False), "neighbor_rank": ("i4", False), "ghost_count": ("u4", False),
    "send_offsets": ("u4", False), "send_indices": ("u4", False), "positions": ("f8", False),
    "global_index": ("u8", False),
}
SCALARS = ("dim", "n_export", "n_internal", "n_owned", "n_relevant", "simd_length", "measure_of_omega", "n_global_dofs")


class StencilData:
    """Same attributes as HostOfflineData (what OfflineData, HyperbolicModule and the tests read),
    backed by caller supplied arrays."""

    def __init__(self, scalars, arrays):
        self.scalars = dict(scalars)
        self.arrays = {}
        for name, (dtype, required) in ARRAYS.items():
            a = arrays.get(name)
            if a is None:
                if required:
                    raise ValueError("StencilData: array »%s« is required" % name)
                continue
            self.arrays[name] = np.ascontiguousarray(a, dtype=np.dtype("<" + dtype)).ravel()
        s = self.scalars
        for key in ("dim", "n_owned", "n_relevant", "measure_of_omega"):
            if key not in s:
                raise ValueError("StencilData: scalar »%s« is required" % key)
        self.dim = int(s["dim"])
        self.n_owned = int(s["n_owned"])
        self.n_relevant = int(s["n_relevant"])
        self.n_export = int(s.get("n_export", 0))
        self.n_internal = int(s.get("n_internal", 0))
        self.simd_length = int(s.get("simd_length", 1))
        self.measure_of_omega = float(s["measure_of_omega"])
        self.n_global_dofs = int(s.get("n_global_dofs", self.n_owned))
        A = self.arrays
        if len(A["lumped_mass_matrix"]) != self.n_relevant:
            raise ValueError("StencilData: lumped_mass_matrix needs n_relevant entries (owned + ghost rows)")

        # CSR view of the pattern (decoded from the SIMD layout if necessary)
        if self.simd_length == 1:
            if len(A["row_starts"]) != self.n_owned + 1:
                raise ValueError("StencilData: row_starts needs n_owned + 1 entries for simd_length = 1")
            self.row_starts = A["row_starts"]
            self.cols = A["column_indices"]
            self.nnz = int(self.row_starts[-1])
            self.mij = A["mass_matrix"]
            self.cij = A["cij_matrix"].reshape(-1, self.dim)
        else:
            self.row_starts, self.cols = self.decode_pattern()
            self.nnz = int(self.row_starts[-1])
            self.mij = self.cij = None  # kept in the reference's layout; decoded by the library
        if len(A["column_indices"]) < self.nnz or len(A["mass_matrix"]) < self.nnz or len(A["cij_matrix"]) < self.nnz * self.dim:
            raise ValueError("StencilData: matrix arrays are shorter than the pattern")
        self.mi = A["lumped_mass_matrix"]

        nb = len(A.get("boundary_dof", ()))
        self.boundary_dof = A.get("boundary_dof", np.zeros(0, dtype=np.uint32))
        self.boundary_normal = A.get("boundary_normal", np.zeros(0)).reshape(nb, self.dim)
        self.boundary_id = A.get("boundary_id", np.zeros(0, dtype=np.int32))
        self.boundary_position = A.get("boundary_position", np.zeros(nb * self.dim)).reshape(nb, self.dim)
        if len(self.boundary_id) != nb or self.boundary_normal.shape != (nb, self.dim):
            raise ValueError("StencilData: inconsistent boundary map arrays")
        self.coupling_pairs = A.get("coupling_pairs", np.zeros(0, dtype=np.uint32)).reshape(-1, 3)
        self.coupling_cji = A["coupling_cji"].reshape(-1, self.dim) if "coupling_cji" in A else None
        self.neighbor_rank = A.get("neighbor_rank", np.zeros(0, dtype=np.int32))
        self.ghost_count = A.get("ghost_count", np.zeros(0, dtype=np.uint32))
        self.send_offsets = A.get("send_offsets", np.zeros(1, dtype=np.uint32))
        self.send_indices = A.get("send_indices", np.zeros(0, dtype=np.uint32))
        if len(self.neighbor_rank) != len(self.ghost_count) or len(self.send_offsets) != len(self.neighbor_rank) + 1:
            raise ValueError("StencilData: inconsistent exchange lists")
        if int(np.sum(self.ghost_count)) != self.n_relevant - self.n_owned:
            raise ValueError("StencilData: ghost_count does not add up to n_relevant - n_owned")
        self.positions = A["positions"].reshape(self.n_owned, self.dim) if "positions" in A else None
        self.global_index = A.get("global_index")

        self.desc = self._make_desc()

    # ---------------------------------------------------------------------------------
    def decode_pattern(self):
        """CSR (row_ptr, cols) of a SparsityPatternSIMD host layout, decoded by the library"""
        A = self.arrays
        n = self.n_owned
        # upper bound of the entries: the column array holds exactly nnz entries
        rp = np.zeros(n + 1, dtype=np.uint64)
        cl = np.zeros(len(A["column_indices"]), dtype=np.uint32)
        _lib.check(_lib.lib.rj_decode_simd_pattern(
            n, self.n_internal, self.simd_length, A["row_starts"].ctypes.data_as(_lib.u64p),
            A["column_indices"].ctypes.data_as(_lib.u32p), rp.ctypes.data_as(_lib.u64p), cl.ctypes.data_as(_lib.u32p)))
        return rp, cl[: int(rp[-1])]

    def _ptr(self, name, ctype):
        a = self.arrays.get(name)
        if a is None or a.size == 0:
            return None
        return a.ctypes.data_as(ctype)

    def _make_desc(self):
        d = _lib.OfflineDesc()
        d.dim = self.dim
        d.n_export, d.n_internal, d.n_owned, d.n_relevant = self.n_export, self.n_internal, self.n_owned, self.n_relevant
        d.simd_length = self.simd_length
        d.row_starts = self._ptr("row_starts", _lib.u64p)
        d.column_indices = self._ptr("column_indices", _lib.u32p)
        d.mass_matrix = self._ptr("mass_matrix", _lib.f64p)
        d.cij_matrix = self._ptr("cij_matrix", _lib.f64p)
        d.lumped_mass_matrix = self._ptr("lumped_mass_matrix", _lib.f64p)
        d.measure_of_omega = self.measure_of_omega
        d.n_boundary = len(self.boundary_dof)
        d.boundary_dof = self._ptr("boundary_dof", _lib.u32p)
        d.boundary_normal = self._ptr("boundary_normal", _lib.f64p)
        d.boundary_id = self._ptr("boundary_id", _lib.i32p)
        d.boundary_position = self._ptr("boundary_position", _lib.f64p)
        d.n_coupling_pairs = len(self.coupling_pairs)
        d.coupling_pairs = self._ptr("coupling_pairs", _lib.u32p)
        d.coupling_cji = self._ptr("coupling_cji", _lib.f64p)
        d.n_neighbors = len(self.neighbor_rank)
        d.neighbor_rank = self._ptr("neighbor_rank", _lib.i32p)
        d.ghost_count = self._ptr("ghost_count", _lib.u32p)
        d.send_offsets = self._ptr("send_offsets", _lib.u32p)
        d.send_indices = self._ptr("send_indices", _lib.u32p)
        d.positions = self._ptr("positions", _lib.f64p)
        return d

    def row_length(self, i):
        return int(self.row_starts[i + 1] - self.row_starts[i])

    # ---------------------------------------------------------------------------------
    @classmethod
    def from_arrays(cls, dim, row_starts, column_indices, mass_matrix, cij_matrix, lumped_mass_matrix, measure_of_omega,
                    n_owned=None, n_relevant=None, n_export=0, n_internal=0, simd_length=1, n_global_dofs=None, **optional):
        lumped = np.asarray(lumped_mass_matrix)
        if n_owned is None:
            if simd_length != 1:
                raise ValueError("n_owned is required for the SIMD host layout")
            n_owned = len(row_starts) - 1
        if n_relevant is None:
            n_relevant = len(lumped)
        scalars = dict(dim=dim, n_export=n_export, n_internal=n_internal, n_owned=n_owned, n_relevant=n_relevant,
                       simd_length=simd_length, measure_of_omega=measure_of_omega,
                       n_global_dofs=n_owned if n_global_dofs is None else n_global_dofs)
        arrays = dict(row_starts=row_starts, column_indices=column_indices, mass_matrix=mass_matrix, cij_matrix=cij_matrix,
                      lumped_mass_matrix=lumped)
        for key, value in optional.items():
            if key not in ARRAYS:
                raise KeyError(key)
            if value is not None:
                arrays[key] = value
        return cls(scalars, arrays)

    @classmethod
    def from_host(cls, host):
        """copy of a generated HostOfflineData (round trips, tests)"""
        return cls.from_arrays(
            host.dim, host.row_starts, host.cols, host.mij, host.cij, host.mi, host.measure_of_omega,
            n_owned=host.n_owned, n_relevant=host.n_relevant, n_export=host.n_export, n_global_dofs=host.n_global_dofs,
            boundary_dof=host.boundary_dof, boundary_normal=host.boundary_normal, boundary_id=host.boundary_id,
            boundary_position=host.boundary_position, coupling_pairs=host.coupling_pairs, coupling_cji=host.coupling_cji,
            neighbor_rank=host.neighbor_rank, ghost_count=host.ghost_count,
            send_offsets=host.send_offsets if len(host.neighbor_rank) else None,
            send_indices=host.send_indices if len(host.neighbor_rank) else None,
            positions=host.positions, global_index=host.global_index)

    # ---------------------------------------------------------------------------------
    def save(self, path):
        header = dict(scalars={k: self.scalars[k] for k in SCALARS if k in self.scalars}, arrays=[])
        header["scalars"].update(dim=self.dim, n_owned=self.n_owned, n_relevant=self.n_relevant, n_export=self.n_export,
                                 n_internal=self.n_internal, simd_length=self.simd_length,
                                 measure_of_omega=self.measure_of_omega, n_global_dofs=self.n_global_dofs)
        order = [n for n in ARRAYS if n in self.arrays]
        for name in order:
            a = self.arrays[name]
            header["arrays"].append([name, ARRAYS[name][0], [int(a.size)]])
        with open(path, "wb") as f:
            f.write(MAGIC)
            f.write((json.dumps(header) + "\n").encode())
            for name in order:
                raw = self.arrays[name].tobytes()
                f.write(raw)
                f.write(b"\0" * (-len(raw) % 8))

    @classmethod
    def load(cls, path):
        with open(path, "rb") as f:
            if f.readline() != MAGIC:
                raise ValueError("%s is not an RJSTENCIL1 file" % path)
            header = json.loads(f.readline().decode())
            arrays = {}
            for name, dtype, shape in header["arrays"]:
                if name not in ARRAYS or ARRAYS[name][0] != dtype:
                    raise ValueError("%s: unexpected array »%s« of type %s" % (path, name, dtype))
                count = int(np.prod(shape))
                nbytes = count * np.dtype(dtype).itemsize
                raw = f.read(nbytes + (-nbytes % 8))
                if len(raw) < nbytes:
                    raise ValueError("%s: truncated array »%s«" % (path, name))
                arrays[name] = np.frombuffer(raw[:nbytes], dtype=np.dtype("<" + dtype)).copy()
        return cls(header["scalars"], arrays)


def encode_simd_layout(row_ptr, cols, mass, cij, n_internal, simd_length):
    """CSR -> the reference's SparsityPatternSIMD / SparseMatrixSIMD host layout
    (sparse_matrix_simd.template.h:94-146 and sparse_matrix_simd.h:389-448): rows [0, n_internal)
    in groups of simd_length rows of equal length, entry (group row g0 + lane, col) at
    (row_starts[g0] + col * simd_length) * n_comp + c * simd_length + lane; CSR tail after that.
    Used to test the importer against the layout a ryujin build would hand over."""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    n = len(row_ptr) - 1
    dim = cij.shape[1]
    L = simd_length
    assert n_internal % L == 0 and n_internal <= n
    row_starts = np.zeros(n + 1, dtype=np.uint64)
    out_cols = np.zeros(len(cols), dtype=np.uint32)
    out_m = np.zeros(len(cols))
    out_c = np.zeros(len(cols) * dim)
    pos = 0
    # one array indexes both parts (sparse_matrix_simd.template.h:94-146): entries
    # [0, n_internal / L] are group offsets, entries [n_internal, n] offsets of the scalar rows
    for g0 in range(0, n_internal, L):
        length = row_ptr[g0 + 1] - row_ptr[g0]
        for lane in range(L):
            assert row_ptr[g0 + lane + 1] - row_ptr[g0 + lane] == length, "rows of a SIMD group must have equal length"
        for col in range(length):
            for lane in range(L):
                e = row_ptr[g0 + lane] + col
                out_cols[pos + col * L + lane] = cols[e]
                out_m[pos + col * L + lane] = mass[e]
                for c in range(dim):
                    out_c[(pos + col * L) * dim + c * L + lane] = cij[e, c]
        pos += length * L
        row_starts[g0 // L + 1] = pos
    row_starts[n_internal] = row_starts[n_internal // L]
    for i in range(n_internal, n):
        for e in range(row_ptr[i], row_ptr[i + 1]):
            out_cols[pos] = cols[e]
            out_m[pos] = mass[e]
            out_c[pos * dim : (pos + 1) * dim] = cij[e]
            pos += 1
        row_starts[i + 1] = pos
    return row_starts, out_cols, out_m, out_c
