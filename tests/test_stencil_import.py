"""OfflineData from foreign arrays (in memory and from an RJSTENCIL1 dump),
in plain CSR and in the reference's SparsityPatternSIMD / SparseMatrixSIMD host layout.  CPU
part: the importer and the library's layout / exchange-list computation (host-only layout)."""
import numpy as np
import pytest

from helpers import make_mesh


def layouts_equal(a, b):
    la, lb = a.layout(), b.layout()
    return all(np.array_equal(x, y) for x, y in zip(la, lb))


@pytest.mark.parametrize("case,n_cells,ranks", [("sod2d", [19, 11], 1), ("sod3d", [7, 6, 9], 3)])
def test_round_trip_through_arrays_and_file(tmp_path, case, n_cells, ranks):
    import ryujin_b200 as rj
    from ryujin_b200.stencil_io import StencilData

    mesh = make_mesh(case, n_cells)
    for rank in range(ranks):
        host = rj.HostOfflineData(mesh, rank, ranks)
        ref = rj.OfflineData(host, host_layout_only=True)
        data = StencilData.from_host(host)
        path = str(tmp_path / ("rank%d.rjstencil" % rank))
        data.save(path)
        loaded = StencilData.load(path)
        for name, a in data.arrays.items():
            assert np.array_equal(a, loaded.arrays[name]), name
        assert (loaded.dim, loaded.n_export, loaded.n_owned, loaded.n_relevant, loaded.nnz) == (
            host.dim, host.n_export, host.n_owned, host.n_relevant, host.nnz)
        for src in (data, loaded):
            od = rj.OfflineData(src, host_layout_only=True)
            assert layouts_equal(od, ref)
            nb_a, rows_a, ents_a = od.exchange_lists()
            nb_b, rows_b, ents_b = ref.exchange_lists()
            assert nb_a == nb_b and np.array_equal(rows_a, rows_b) and np.array_equal(ents_a, ents_b)


def test_oracle_assembled_arrays_import(oracle):
    """arrays from an independent assembly (the oracle's quadrature based Q1 assembly) give the same
    device layout as the built-in closed-form generator"""
    import ryujin_b200 as rj
    from oracle import orc
    from ryujin_b200.stencil_io import StencilData

    n = [13, 9]
    mesh = rj.CartesianMesh.rectangular_domain([0.0, 0.0], [2.0, 1.0], n, [rj.Boundary.dirichlet, rj.Boundary.do_nothing,
                                                                            rj.Boundary.slip, rj.Boundary.slip])
    host = rj.HostOfflineData(mesh)
    off = oracle.assemble(n, [0.0, 0.0], [2.0 / n[0], 1.0 / n[1]], [orc.DIRICHLET, orc.DO_NOTHING, orc.SLIP, orc.SLIP])
    a = off.arrays()
    data = StencilData.from_arrays(
        2, a["row_ptr"].astype(np.uint64), a["cols"], a["mij"], a["cij"], a["mi"], a["measure"],
        boundary_dof=a["b_index"], boundary_normal=a["b_normal"], boundary_id=a["b_id"], boundary_position=a["b_position"],
        coupling_pairs=a["pairs"])
    od = rj.OfflineData(data, host_layout_only=True)
    assert layouts_equal(od, rj.OfflineData(host, host_layout_only=True))
    assert np.allclose(data.mij, host.mij, rtol=1e-14, atol=0) and np.allclose(data.cij, host.cij, rtol=0, atol=1e-15)


def renumbered_csr_and_simd(simd_length, n_cells=(12, 10)):
    """a 2D stencil renumbered so that the first n_internal rows have equal length (what the
    reference's renumbering guarantees), as CSR arrays and in the reference's SIMD host layout"""
    import ryujin_b200 as rj
    from ryujin_b200.stencil_io import StencilData, encode_simd_layout

    mesh = make_mesh("sod2d", list(n_cells))
    host = rj.HostOfflineData(mesh)
    lengths = np.diff(host.row_starts.astype(np.int64))
    interior = np.where(lengths == 9)[0]
    n_internal = (len(interior) // simd_length) * simd_length
    order = np.concatenate([interior[:n_internal], np.setdiff1d(np.arange(host.n_owned), interior[:n_internal])])
    new_of_old = np.empty(host.n_owned, dtype=np.int64)
    new_of_old[order] = np.arange(host.n_owned)
    rp = [0]
    cols, mass, cij = [], [], []
    for old in order:
        r0, r1 = int(host.row_starts[old]), int(host.row_starts[old + 1])
        c = new_of_old[host.cols[r0:r1]]
        # diagonal first, then ascending (what the library requires and the reference produces)
        idx = np.concatenate([[0], 1 + np.argsort(c[1:])])
        cols.extend(c[idx])
        mass.extend(host.mij[r0:r1][idx])
        cij.extend(host.cij[r0:r1][idx])
        rp.append(len(cols))
    rp, cols, mass, cij = np.array(rp), np.array(cols, dtype=np.uint32), np.array(mass), np.array(cij)
    mi = host.mi[: host.n_owned][order]
    bdof = new_of_old[host.boundary_dof]
    keep = np.argsort(bdof, kind="stable")
    common = dict(boundary_dof=bdof[keep], boundary_normal=host.boundary_normal[keep], boundary_id=host.boundary_id[keep],
                  boundary_position=host.boundary_position[keep], positions=host.positions[order])
    csr = StencilData.from_arrays(2, rp.astype(np.uint64), cols, mass, cij, mi, host.measure_of_omega, **common)
    rs, sc, sm, scij = encode_simd_layout(rp, cols, mass, cij, n_internal, simd_length)
    simd = StencilData.from_arrays(2, rs, sc, sm, scij, mi, host.measure_of_omega, n_owned=host.n_owned,
                                   n_internal=n_internal, simd_length=simd_length, **common)
    return csr, simd, (rp, cols, n_internal, rs, sc)


@pytest.mark.parametrize("simd_length", [2, 4, 8])
def test_reference_simd_host_layout_is_accepted(oracle, simd_length):
    """the arrays exactly as a ryujin build holds them (SIMD groups + CSR tail): the decoded
    pattern and the resulting device layout equal those of the CSR description, and the encoder
    used here agrees with the oracle's restatement of SparsityPatternSIMD::reinit"""
    import ryujin_b200 as rj

    csr, simd, (rp, cols, n_internal, rs, sc) = renumbered_csr_and_simd(simd_length)
    o_rs, o_ci, _ = oracle.sparsity_simd(rp, cols, n_internal, simd_length)
    assert np.array_equal(rs, o_rs) and np.array_equal(sc, o_ci)
    assert np.array_equal(simd.row_starts, rp) and np.array_equal(simd.cols, cols)
    assert layouts_equal(rj.OfflineData(simd, host_layout_only=True), rj.OfflineData(csr, host_layout_only=True))


def test_importer_rejects_inconsistent_input(tmp_path):
    from ryujin_b200.stencil_io import StencilData

    with pytest.raises(ValueError):
        StencilData.from_arrays(1, [0, 2, 4], [0, 1, 1, 0], [1.0] * 4, [[0.5]] * 4, [1.0], 2.0)  # m_i too short
    with pytest.raises(ValueError):
        StencilData.from_arrays(1, [0, 2, 4], [0, 1, 1, 0], [1.0] * 4, [[0.5]] * 4, [1.0, 1.0], 2.0, ghost_count=[3],
                                neighbor_rank=[1], send_offsets=[0, 0])
    bad = tmp_path / "x.rjstencil"
    bad.write_bytes(b"not a stencil\n")
    with pytest.raises(ValueError):
        StencilData.load(str(bad))


@pytest.mark.gpu
def test_imported_stencil_runs_like_the_generated_one(tmp_path):
    """the same mesh once through the generator and once through an RJSTENCIL1 dump: bit-identical
    states after three ERK33 cycles"""
    import ryujin_b200 as rj
    from helpers import CASES

    mesh = make_mesh("sod2d", [40, 24])
    host = rj.HostOfflineData(mesh)
    path = str(tmp_path / "sod.rjstencil")
    rj.StencilData.from_host(host).save(path)
    iv = dict(CASES["sod2d"]["iv"])
    kind = getattr(rj.InitialState, iv.pop("kind"))
    results = []
    for od in (rj.OfflineData(host), rj.OfflineData.from_stencil(path)):
        module = rj.HyperbolicModule(od, rj.HyperbolicSystem(), rj.InitialValues(kind, **iv))
        ti = rj.TimeIntegrator(module)
        state = module.state_vector()
        state.interpolate(0.0)
        t = 0.0
        for _ in range(3):
            t += ti.step(state, t)
        results.append((t, state.download()))
    assert results[0][0] == results[1][0]
    assert np.array_equal(results[0][1], results[1][1])


@pytest.mark.gpu
@pytest.mark.parametrize("simd_length", [4, 8])
def test_simd_host_layout_runs_like_the_csr_description(simd_length):
    """the matrices handed over in the reference's SparseMatrixSIMD layout (mass matrix with one,
    c_ij with dim components per entry) are decoded into the device layout exactly like the CSR
    description: bit-identical states after three ERK33 cycles on a renumbered stencil"""
    import ryujin_b200 as rj
    from helpers import CASES

    csr, simd, _ = renumbered_csr_and_simd(simd_length, n_cells=(20, 14))
    iv = dict(CASES["sod2d"]["iv"])
    kind = getattr(rj.InitialState, iv.pop("kind"))
    results = []
    for data in (csr, simd):
        od = rj.OfflineData(data)
        module = rj.HyperbolicModule(od, rj.HyperbolicSystem(), rj.InitialValues(kind, **iv))
        ti = rj.TimeIntegrator(module)
        state = module.state_vector()
        state.interpolate(0.0)
        t = 0.0
        for _ in range(3):
            t += ti.step(state, t)
        results.append((t, state.download()))
    assert results[0][0] == results[1][0] and np.array_equal(results[0][1], results[1][1])
    assert np.max(np.abs(results[0][1][:, 0] - 1.0)) > 1e-3  # the shock tube did evolve
