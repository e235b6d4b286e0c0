"""stencils that are NOT Cartesian.  The arrays of mapped (non-affine) Q1 meshes
-- the rounded corner of the reference's forward-facing step (source/geometry_step.h:36-128) and a
cylinder O-grid -- are assembled by the oracle's MappingQ1 quadrature assembly
(oracle/orc_mesh_quads.h; restates offline_data.template.h:420-1105), travel through the product's
stencil importer (StencilData.from_arrays / RJSTENCIL1 dump) and run on the device against the oracle
fed with the same arrays."""
import os
import time

import numpy as np
import pytest

import unstructured_meshes as um
from helpers import matching_oracle, rel_linf


@pytest.fixture(scope="module")
def oracle():
    return matching_oracle()[0]


def test_mapped_assembly_reduces_to_the_cartesian_one(oracle):
    """on affine cells the MappingQ1 assembly gives the Cartesian assembly's structure exactly (pattern,
    boundary map incl. the order of a DoF's entries, coupling pairs) and its values to round-off"""
    from oracle import orc

    nx, ny = 13, 9
    _, cells = um.structured_cells(nx, ny)
    X, Y = np.meshgrid(np.linspace(0, 2, nx + 1), np.linspace(0, 1, ny + 1))
    V = np.stack([X, Y], -1).reshape(-1, 2)
    faces, flip = um.boundary_faces(cells)
    mid = 0.5 * (V[faces[:, 0]] + V[faces[:, 1]])
    ids = np.full(len(faces), orc.SLIP, dtype=np.int32)
    ids[mid[:, 0] < 1e-9] = orc.DIRICHLET
    ids[mid[:, 0] > 2 - 1e-9] = orc.DO_NOTHING
    _, a, _ = um.assemble(oracle, V, cells, (faces, flip), ids, reorder=False)
    ref = oracle.assemble([nx, ny], [0.0, 0.0], [2.0 / nx, 1.0 / ny], [orc.DIRICHLET, orc.DO_NOTHING, orc.SLIP, orc.SLIP]).arrays()
    for key in ("row_ptr", "cols", "transposed", "b_index", "b_id", "pairs"):
        assert np.array_equal(a[key], ref[key]), key
    for key in ("mij", "cij", "mi", "b_normal", "b_masses"):
        assert np.max(np.abs(a[key] - ref[key])) < 1e-15, key
    assert abs(a["measure"] - ref["measure"]) < 1e-14


def test_rounded_step_geometry_and_importer(oracle, tmp_path):
    """the rounded corner: area = polygon of the arc, every boundary DoF has a unit normal, constants are in
    the kernel of c_ij (sum_j c_ij = boundary term only), and the arrays survive the RJSTENCIL1 dump"""
    import ryujin_b200 as rj

    v, c, f, ids = um.forward_facing_step_rounded(0)
    off, a, vertices = um.assemble(oracle, v, c, f, ids)
    exact = 3.0 - 2.4 * 0.2 + 0.0125**2 * (1.0 - np.pi / 4.0)  # sharp corner replaced by a quarter circle fillet
    assert abs(a["measure"] - exact) < 2e-5 and a["measure"] > 3.0 - 2.4 * 0.2
    assert np.allclose(np.linalg.norm(a["b_normal"], axis=1), 1.0, atol=1e-12)
    rows = np.repeat(np.arange(len(a["mi"])), np.diff(a["row_ptr"]))
    interior = np.ones(len(a["mi"]), dtype=bool)
    interior[a["b_index"]] = False
    csum = np.zeros((len(a["mi"]), 2))
    np.add.at(csum, rows, a["cij"])
    assert np.max(np.abs(csum[interior])) < 1e-15  # partition of unity: sum_j c_ij = 0 away from the boundary
    # some cells really are non-affine: c_ij != -c_ji is allowed only at the boundary, but m_ij rows near the
    # corner differ from the Cartesian value h^2 4/9
    assert np.ptp(a["mi"][interior]) > 1e-6

    data = um.stencil_data(a, vertices)
    path = str(tmp_path / "step.rjstencil")
    data.save(path)
    back = rj.StencilData.load(path)
    for name in ("cols", "mij", "cij", "mi", "boundary_normal", "boundary_id", "coupling_pairs"):
        assert np.array_equal(getattr(back, name), getattr(data, name)), name
    od = rj.OfflineData.from_stencil(path, host_layout_only=True)
    assert od.n_locally_owned == len(a["mi"])


def run_parity(oracle, off, arrays, vertices, primitive, cycles, via_file=None):
    import ryujin_b200 as rj
    from oracle import orc

    data = um.stencil_data(arrays, vertices)
    if via_file:
        data.save(via_file)
        od = rj.OfflineData.from_stencil(via_file)
    else:
        od = rj.OfflineData.from_stencil(data)
    iv = rj.InitialValues(rj.InitialState.uniform, direction=[1.0, 0.0], primitive=list(primitive))
    module = rj.HyperbolicModule(od, rj.HyperbolicSystem(), iv)
    sim = off.euler_sim(kind=orc.UNIFORM, primitive=tuple(primitive), direction=(1.0, 0.0))
    sim._keepalive = off
    ti = rj.TimeIntegrator(module)
    sim.configure_time_integrator("erk 33", 0.45, 0.9, "bang bang control")
    state = module.state_vector()
    state.interpolate(0.0)
    sim.set_U(state.download(), 0)
    t = t_ref = 0.0
    module.synchronize()
    t0 = time.perf_counter()
    taus = []
    for _ in range(cycles):
        tau = ti.step(state, t)
        taus.append(tau)
        t += tau
    module.synchronize()
    seconds = time.perf_counter() - t0
    taus_ref = []
    for _ in range(cycles):
        tau_ref = sim.time_step(t_ref)
        taus_ref.append(tau_ref)
        t_ref += tau_ref
    U, U_ref = state.download(), sim.get_U(0)
    err = rel_linf(U, U_ref)
    dtau = float(np.max(np.abs(np.array(taus) - np.array(taus_ref)) / np.array(taus_ref)))
    assert (module.n_restarts(), module.n_warnings()) == sim.counters()
    return err, dtau, seconds, len(arrays["mi"]), U


@pytest.mark.gpu
def test_rounded_step_parity(oracle):
    v, c, f, ids = um.forward_facing_step_rounded(0)
    off, a, vertices = um.assemble(oracle, v, c, f, ids)
    err, dtau, _, n, U = run_parity(oracle, off, a, vertices, (1.4, 3.0, 1.0), 60)
    print("rounded step %d DoFs, 60 cycles: rel Linf %.3e, tau %.2e" % (n, err, dtau))
    assert err <= 1e-12 and dtau <= 1e-13
    assert np.all(np.isfinite(U)) and U[:, 0].max() > 1.5  # the bow shock has formed


@pytest.mark.gpu
def test_cylinder_o_grid_parity(oracle):
    v, c, f, ids = um.cylinder_o_grid(32, 128)
    off, a, vertices = um.assemble(oracle, v, c, f, ids)
    err, dtau, _, n, U = run_parity(oracle, off, a, vertices, (1.4, 3.0, 1.0), 60)
    print("cylinder O-grid %d DoFs, 60 cycles: rel Linf %.3e, tau %.2e" % (n, err, dtau))
    assert err <= 1e-12 and dtau <= 1e-13


@pytest.mark.gpu
def test_rounded_step_benchmark_size(oracle, tmp_path):
    """prm/benchmarks/euler-mach3-forward-facing-step.prm with the reference's rounded corner at the
    benchmark's 1.04M gridpoints, imported through an RJSTENCIL1 dump; parity and 2D throughput"""
    v, c, f, ids = um.forward_facing_step_rounded(3)
    assert 1.0e6 < len(v) < 1.1e6
    off, a, vertices = um.assemble(oracle, v, c, f, ids)
    cycles = 20
    err, dtau, seconds, n, _ = run_parity(oracle, off, a, vertices, (1.4, 3.0, 1.0), cycles, via_file=str(tmp_path / "step7.rjstencil"))
    rate = n * 3 * cycles / seconds / 1e6
    b_alg = 3192.0  # BASELINE.md section 3, 2D Euler
    print("rounded step %d DoFs (RCM order), %d cycles: rel Linf %.3e, tau %.2e; %.1f MDoF-updates/s incl. host "
          "round trips = %.0f GB/s algorithmic (2D B_alg %d B)" % (n, cycles, err, dtau, rate, rate * 1e6 * b_alg / 1e9, b_alg))
    assert err <= 1e-12 and dtau <= 1e-13
