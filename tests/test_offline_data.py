"""Host logic of the product (no GPU): the closed-form Cartesian Q1 generator against the
oracle's independent quadrature assembly, the reference's DEBUG invariants
(source/offline_data.template.h:1002-1104), the slab partition + exchange lists, the
SIMD-layout decoder, and the exported C ABI."""
import ctypes
import re

import numpy as np
import pytest

import ryujin_b200 as rj
from helpers import make_mesh
from oracle import orc

B = rj.Boundary


def rows(host):
    return [(int(host.row_starts[i]), int(host.row_starts[i + 1])) for i in range(host.n_owned)]


def compare_with_oracle(host, off):
    """host (product, local numbering) vs off (oracle assembly, lexicographic == global)."""
    a = off.arrays()
    g = host.global_index.astype(np.int64)
    assert host.n_global_dofs == off.n
    assert abs(host.measure_of_omega - a["measure"]) <= 1e-13 * a["measure"]
    for i, (r0, r1) in enumerate(rows(host)):
        gi = g[i]
        o0, o1 = int(a["row_ptr"][gi]), int(a["row_ptr"][gi + 1])
        assert r1 - r0 == o1 - o0
        assert host.cols[r0] == i  # diagonal first
        assert np.all(np.diff(host.cols[r0 + 1 : r1].astype(np.int64)) > 0)  # ascending
        mine = {int(g[host.cols[e]]): e for e in range(r0, r1)}
        for e in range(o0, o1):
            f = mine[int(a["cols"][e])]
            assert abs(host.mij[f] - a["mij"][e]) <= 4e-16 * abs(a["mij"][e])
            assert np.allclose(host.cij[f], a["cij"][e], rtol=0, atol=4e-16 * np.max(np.abs(a["cij"])))
        assert abs(host.mi[i] - a["mi"][gi]) <= 4e-16 * a["mi"][gi]
        assert np.allclose(host.positions[i], a["positions"][gi], atol=1e-15)


def compare_boundary_single_rank(host, off):
    a = off.arrays()
    # single rank: local numbering == lexicographic, so the maps must agree entry by entry
    assert np.array_equal(host.global_index, np.arange(host.n_owned))
    assert np.array_equal(host.boundary_dof, a["b_index"])
    assert np.array_equal(host.boundary_id, a["b_id"])
    assert np.allclose(host.boundary_normal, a["b_normal"], atol=1e-15)
    assert np.allclose(host.boundary_position, a["b_position"], atol=1e-15)
    assert np.array_equal(host.coupling_pairs, a["pairs"])


@pytest.mark.parametrize(
    "n_cells,origin,extent,ids",
    [
        ([8, 6], [0.0, 0.0], [1.0, 1.0], [B.dirichlet, B.dirichlet, B.slip, B.slip]),
        ([5, 7], [-5.0, -5.0], [10.0, 10.0], [B.dirichlet] * 4),
        ([4, 3, 5], [0.0, 0.0, 0.0], [1.0, 0.5, 0.25], [B.dirichlet, B.do_nothing, B.slip, B.slip, B.slip, B.dirichlet]),
        ([3, 3, 3], [-1.0, -1.0, -1.0], [2.0, 2.0, 2.0], [B.slip] * 6),
        ([16], [0.0], [1.0], [B.dirichlet, B.dirichlet]),
    ],
)
def test_generator_matches_quadrature_assembly(oracle, n_cells, origin, extent, ids):
    mesh = rj.CartesianMesh(n_cells, origin, extent, ids)
    host = rj.HostOfflineData(mesh)
    off = oracle.assemble(n_cells, origin, mesh.h, [int(b) for b in ids])
    compare_with_oracle(host, off)
    compare_boundary_single_rank(host, off)


def test_forward_facing_step_mesh(oracle):
    """config 3 topology: masked grid with one re-entrant corner (row length 8)."""
    mesh = rj.CartesianMesh.forward_facing_step(refinement=0)
    host = rj.HostOfflineData(mesh)
    mask = mesh.cell_mask
    off = oracle.assemble(mesh.n_cells, mesh.origin, mesh.h, mesh.boundary_ids, mask=mask, masked_id=int(B.slip))
    compare_with_oracle(host, off)
    compare_boundary_single_rank(host, off)
    lengths = np.diff(host.row_starts.astype(np.int64))
    assert sorted(set(lengths.tolist())) == [4, 6, 8, 9]
    assert np.sum(lengths == 8) == 1  # the re-entrant corner
    # slip/slip corners became no_slip (offline_data.template.h:1323-1336)
    ids = host.boundary_id
    assert np.sum(ids == B.no_slip) > 0


@pytest.mark.parametrize("n_cells", [[6, 5], [4, 3, 3]])
def test_offline_data_invariants(n_cells):
    dim = len(n_cells)
    mesh = rj.CartesianMesh(n_cells, [0.0] * dim, [1.0] * dim, [B.slip] * (2 * dim))
    host = rj.HostOfflineData(mesh)
    pairs = {(int(i), int(j)) for i, _, j in host.coupling_pairs}
    entry = {}
    for i, (r0, r1) in enumerate(rows(host)):
        for e in range(r0, r1):
            entry[(i, int(host.cols[e]))] = e
    for i, (r0, r1) in enumerate(rows(host)):
        assert abs(np.sum(host.mij[r0:r1]) - host.mi[i]) <= 1e-12 * host.mi[i]
        for e in range(r0, r1):
            j = int(host.cols[e])
            f = entry[(j, i)]
            assert abs(host.mij[e] - host.mij[f]) <= 1e-12 * abs(host.mij[e])
            if i != j and (i, j) not in pairs:
                assert np.allclose(host.cij[e], -host.cij[f], atol=1e-12)
    boundary = set(host.boundary_dof.tolist())
    for i, (r0, r1) in enumerate(rows(host)):
        if i not in boundary:
            assert np.allclose(np.sum(host.cij[r0:r1], axis=0), 0.0, atol=1e-12)
    assert abs(np.sum(host.mi) - host.measure_of_omega) <= 1e-12


@pytest.mark.parametrize("n_cells,n_ranks", [([6, 7], 2), ([6, 9], 3), ([3, 4, 8], 2), ([3, 3, 9], 4)])
def test_slab_partition_and_exchange_lists(n_cells, n_ranks):
    dim = len(n_cells)
    mesh = rj.CartesianMesh(n_cells, [0.0] * dim, [1.0] * dim, [B.dirichlet] * (2 * dim))
    single = rj.HostOfflineData(mesh)
    parts = [rj.HostOfflineData(mesh, r, n_ranks) for r in range(n_ranks)]
    owned = np.concatenate([p.global_index[: p.n_owned] for p in parts])
    assert sorted(owned.tolist()) == list(range(single.n_owned))

    # the stencil of every owned row equals the single-rank stencil (values bit-identical:
    # cell sums do not depend on the numbering)
    ref = {}
    for i, (r0, r1) in enumerate(rows(single)):
        for e in range(r0, r1):
            ref[(i, int(single.cols[e]))] = (single.mij[e], tuple(single.cij[e]))
    for p in parts:
        g = p.global_index.astype(np.int64)
        assert p.n_export <= p.n_owned <= p.n_relevant
        for i, (r0, r1) in enumerate(rows(p)):
            has_ghost = False
            for e in range(r0, r1):
                j = int(p.cols[e])
                has_ghost |= j >= p.n_owned
                m, c = ref[(int(g[i]), int(g[j]))]
                assert p.mij[e] == m and tuple(p.cij[e]) == c
            assert has_ghost == (i < p.n_export)  # export rows come first
        assert np.array_equal(p.mi, single.mi[g])

    # send list of p towards q == ghost rows of q owned by p, in the same order
    for p in parts:
        for a, q in enumerate(p.neighbor_rank):
            other = parts[int(q)]
            b = list(other.neighbor_rank).index(p.rank)
            sent = p.global_index[p.send_indices[p.send_offsets[a] : p.send_offsets[a + 1]]]
            start = other.n_owned + int(np.sum(other.ghost_count[:b]))
            recv = other.global_index[start : start + int(other.ghost_count[b])]
            assert np.array_equal(sent, recv)
    # boundary maps of the parts partition the single-rank map
    total = sum(len(p.boundary_dof) for p in parts)
    assert total == len(single.boundary_dof)


def test_decode_simd_pattern_roundtrip(oracle):
    """The ABI accepts the reference's SparsityPatternSIMD host layout: encode with the
    oracle's restatement of reinit(), decode with the product."""
    mesh = rj.CartesianMesh([9, 9], [0.0, 0.0], [1.0, 1.0], [B.slip] * 4)
    host = rj.HostOfflineData(mesh)
    lengths = np.diff(host.row_starts.astype(np.int64))
    # sort rows so that groups of 4 have equal length (as the reference's renumbering does)
    order = np.argsort(-lengths, kind="stable")
    inv = np.empty_like(order)
    inv[order] = np.arange(len(order))
    row_ptr = [0]
    cols = []
    for new_i, old_i in enumerate(order):
        r0, r1 = int(host.row_starts[old_i]), int(host.row_starts[old_i + 1])
        c = inv[host.cols[r0:r1]]
        c = [new_i] + sorted(int(x) for x in c if x != new_i)
        cols += c
        row_ptr.append(len(cols))
    row_ptr = np.array(row_ptr, dtype=np.uint32)
    cols = np.array(cols, dtype=np.uint32)
    new_len = np.diff(row_ptr.astype(np.int64))
    width = 4
    n_internal = 0
    while n_internal + width <= len(new_len) and len(set(new_len[n_internal : n_internal + width])) == 1:
        n_internal += width
    assert n_internal >= 48
    row_starts, ci, _ = oracle.sparsity_simd(row_ptr, cols, n_internal, width)
    out_ptr = np.zeros(len(row_ptr), dtype=np.uint64)
    out_cols = np.zeros(len(cols), dtype=np.uint32)
    rc = rj.lib.rj_decode_simd_pattern(
        len(row_ptr) - 1, n_internal, width, row_starts.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
        ci.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)), out_ptr.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
        out_cols.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32)),
    )
    assert rc == 0
    assert np.array_equal(out_ptr, row_ptr.astype(np.uint64))
    assert np.array_equal(out_cols, cols)


def test_c_abi_exports_every_declared_symbol():
    """Every function declared in include/ryujin_b200.h is exported by the library."""
    import os

    header = open(os.path.join(os.path.dirname(__file__), "..", "include", "ryujin_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(rj_[A-Za-z0-9_]+)\s*\(", header))
    assert len(declared) >= 45
    lib = ctypes.CDLL(rj.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), name
    assert declared == set(rj.EXPORTED_SYMBOLS)


def test_no_cpu_fallback():
    """Without a CUDA device every compute entry point refuses to run."""
    if rj.device_count() > 0:
        pytest.skip("a CUDA device is present")
    mesh = rj.CartesianMesh([4, 4], [0.0, 0.0], [1.0, 1.0], [B.slip] * 4)
    with pytest.raises(rj.NoDeviceError):
        rj.OfflineData.from_mesh(mesh)


@pytest.mark.parametrize("case,n_cells,ranks", [("sod2d", [17, 23], 2), ("sod3d", [6, 5, 13], 3), ("step", None, 3)])
def test_export_rows_are_exactly_the_rows_that_touch_ghosts(case, n_cells, ranks):
    """The two-stream overlap launches rows [0, n_export) before the halo
    exchange and the remaining rows concurrently with it.  That is only correct if (a) every row
    that reads a ghost column and (b) every row that is sent to a neighbour lies in [0, n_export)
    (offline_data.h:136-169: export indices come first)."""
    import ryujin_b200 as rj

    mesh = rj.CartesianMesh.forward_facing_step(1) if case == "step" else make_mesh(case, n_cells)
    for rank in range(ranks):
        host = rj.HostOfflineData(mesh, rank, ranks)
        rows = np.repeat(np.arange(host.n_owned), np.diff(host.row_starts.astype(np.int64)))
        touches_ghost = np.unique(rows[host.cols >= host.n_owned])
        assert host.n_export <= host.n_owned
        if len(touches_ghost):
            assert touches_ghost.max() < host.n_export
        if len(host.send_indices):
            assert host.send_indices.max() < host.n_export
        # and the range is tight: every export row is sent somewhere
        assert np.array_equal(np.unique(host.send_indices), np.arange(host.n_export))
