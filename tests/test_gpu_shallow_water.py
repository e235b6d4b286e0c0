"""Parity of the CUDA shallow water plug-in (ryujin_b200/csrc/shallow_water.cuh; SURVEY.md 8a
row a18, BASELINE config 5) against the CPU oracle (oracle/orc_shallow_water.h, pinned against
the reference's goldens in test_oracle_shallow_water.py).  Needs a B200: `-m gpu`.
"""
import numpy as np
import pytest

from helpers import (B_DYNAMIC, SW_CASES, make_sw_oracle, matching_oracle, make_sw_product, rel_linf, rel_linf_per_component, sw_bathymetry)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oracle():
    """the oracle build whose pow() is the one the product was built for (Manning friction:
    h^(4/3)); default: oracle/_ref/liboracle_vcl.so, the reference's own VCL pow"""
    return matching_oracle()[0]


def gpu_and_oracle(oracle, case, n_cells, bathymetry=False):
    od, module = make_sw_product(case, n_cells)
    off, sim = make_sw_oracle(oracle, case, od.host)
    if bathymetry:
        z = sw_bathymetry(od.host)
        module.set_initial_precomputed(z)
        sim.set_initial_precomputed(z)
    return od, module, off, sim


def numpy_initial_state(case, host, t):
    """InitialValues of the shallow water cases in numpy (direction along x only for dam breaks)."""
    c = SW_CASES[case]
    iv = c["iv"]
    dim = host.dim
    x = host.positions
    d = np.array(iv["direction"], dtype=float)
    d /= np.linalg.norm(d)
    y = x - np.array(iv["position"], dtype=float)
    if dim == 2:
        y = np.stack([d[0] * y[:, 0] + d[1] * y[:, 1], -d[1] * y[:, 0] + d[0] * y[:, 1]], axis=1)
    U = np.zeros((host.n_owned, dim + 1))
    p = iv["sw_params"]
    g = c["eq"]["gravity"]
    if iv["kind"] == "sw_circular_dam_break":
        U[:, 0] = np.where(np.sum(y * y, axis=1) <= p[1], p[2], p[0])
    elif iv["kind"] == "sw_ritter_dam_break":
        t0, hl = p[0], p[1]
        aL = np.sqrt(g * hl)
        xx = y[:, 0]
        xA, xB = -(t + t0) * aL, 2.0 * (t + t0) * aL
        tmp = aL - xx / (2.0 * (t + t0))
        h_exp = 4.0 / (9.0 * g) * tmp * tmp
        v_exp = 2.0 / 3.0 * (xx / (t + t0) + aL)
        U[:, 0] = np.where(xx <= xA, hl, np.where(xx <= xB, h_exp, 0.0))
        U[:, 1] = np.where(xx <= xA, 0.0, np.where(xx <= xB, h_exp * v_exp, 0.0))
    else:
        depth, mach, beta = p[0], p[1], p[2]
        pb0, pb1 = y[:, 0] - mach * t, y[:, 1]
        factor = beta / (2.0 * np.pi) * np.exp(0.5 - 0.5 * (pb0 * pb0 + pb1 * pb1))
        h = depth - (1.0 / (2.0 * g)) * factor * factor
        m0, m1 = h * (mach - factor * pb1), h * (factor * pb0)
        U[:, 0] = h
        U[:, 1] = d[0] * m0 - d[1] * m1
        U[:, 2] = d[1] * m0 + d[0] * m1
    return U


@pytest.mark.parametrize("case,n_cells", [("sw_vortex2d", [24, 24]), ("sw_dam2d", [20, 20]), ("sw_ritter1d", [200])])
def test_sw_interpolate_initial_values(case, n_cells):
    od, module = make_sw_product(case, n_cells)
    s = module.state_vector()
    for t in (0.0, 0.4):
        s.interpolate(t)
        ref = numpy_initial_state(case, od.host, t)
        assert rel_linf(s.download(), ref) < 5e-15, (case, t)


@pytest.mark.parametrize(
    "case,n_cells,bathymetry",
    [
        ("sw_vortex2d", [32, 32], False),
        ("sw_vortex2d", [24, 20], True),
        ("sw_dam2d", [40, 40], False),
        ("sw_dam2d_friction", [36, 28], True),
        ("sw_ritter1d", [400], False),
    ],
)
def test_sw_single_substep_parity(oracle, case, n_cells, bathymetry):
    """prepare_state_vector + step<0>: every intermediate of the sweep against the oracle."""
    od, module, off, sim = gpu_and_oracle(oracle, case, n_cells, bathymetry)
    assert (module.k, module.n_precomputed, module.n_bounds) == (od.host.dim + 1, 2, 5)
    old, new = module.state_vector(), module.state_vector()
    old.interpolate(0.0)
    U0 = old.download()
    if case == "sw_dam2d_friction":
        # give the friction source something to act on: a rotating velocity field
        x = od.host.positions
        U0[:, 1] = U0[:, 0] * (0.4 - 0.1 * x[:, 1])
        U0[:, 2] = U0[:, 0] * (0.1 * x[:, 0])
        old.upload(U0)
    sim.set_U(U0, 0)

    t = 0.05
    module.prepare_state_vector(old, t)
    sim.prepare_state_vector(0, t)
    assert rel_linf(old.download(), sim.get_U(0)) < 2e-15
    assert np.max(rel_linf_per_component(old.download_precomputed(), sim.get_prec(0))) < 4e-15

    module.cfl = 0.5
    sim.set_cfl(0.5)
    tau = module.step(old, [], [], new)
    tau_ref = sim.step(0, [], [], 1)
    assert abs(tau - tau_ref) <= 1e-14 * tau_ref

    k = module.k
    tol = 2e-13
    assert np.max(np.abs(module.scratch("alpha") - sim.scratch("alpha"))) < 1e-11
    d_gpu, d_ref = module.scratch("dij"), sim.scratch("dij")
    assert np.max(np.abs(d_gpu - d_ref)) <= 1e-14 * np.max(np.abs(d_ref))
    assert np.max(rel_linf_per_component(module.scratch("r").reshape(-1, k), sim.scratch("r").reshape(-1, k))) < 1e-11
    assert np.max(rel_linf_per_component(module.scratch("bounds").reshape(-1, 5), sim.scratch("bounds").reshape(-1, 5))) < tol
    p_gpu, p_ref = module.scratch("pij").reshape(-1, k), sim.scratch("pij").reshape(-1, k)
    diag = od.host.row_starts[:-1].astype(np.int64)
    mask = np.ones(len(p_ref), dtype=bool)
    mask[diag] = False
    scale = np.maximum(np.max(np.abs(p_ref[mask]), axis=0), 1e-300)
    assert np.max(np.abs(p_gpu[mask] - p_ref[mask]) / scale) < 1e-10
    for name in ("lij", "lij_next"):
        a, b = module.scratch(name)[mask], sim.scratch(name)[mask]
        assert np.max(np.abs(a - b)) < 1e-8, name
        assert np.all((a >= 0.0) & (a <= 1.0))
    assert rel_linf(new.download(), sim.get_U(1)) < 1e-13
    assert (module.n_restarts(), module.n_warnings()) == sim.counters()


@pytest.mark.parametrize("case,n_cells,bathymetry", [("sw_vortex2d", [24, 24], True), ("sw_dam2d_friction", [30, 30], True)])
def test_sw_stage_weights_parity(oracle, case, n_cells, bathymetry):
    """step<1>, step<2> with the ERK33 weights: high-order flux and source terms of the stages."""
    od, module, off, sim = gpu_and_oracle(oracle, case, n_cells, bathymetry)
    S = [module.state_vector() for _ in range(4)]
    S[0].interpolate(0.0)
    sim.set_U(S[0].download(), 0)
    module.cfl = 0.4
    sim.set_cfl(0.4)
    t = 0.0
    module.prepare_state_vector(S[0], t)
    sim.prepare_state_vector(0, t)
    tau = module.step(S[0], [], [], S[1])
    tau_ref = sim.step(0, [], [], 1)
    assert abs(tau - tau_ref) <= 1e-14 * tau_ref
    module.prepare_state_vector(S[1], t + tau)
    sim.prepare_state_vector(1, t + tau_ref)
    module.step(S[1], [S[0]], [-1.0], S[2], tau)
    sim.step(1, [0], [-1.0], 2, tau_ref)
    assert rel_linf(S[2].download(), sim.get_U(2)) < 1e-13
    module.prepare_state_vector(S[2], t + 2 * tau)
    sim.prepare_state_vector(2, t + 2 * tau_ref)
    module.step(S[2], [S[0], S[1]], [0.75, -2.0], S[3], tau)
    sim.step(2, [0, 1], [0.75, -2.0], 3, tau_ref)
    assert rel_linf(S[3].download(), sim.get_U(3)) < 1e-13


def run_cycles(module, sim, n_cycles, scheme="erk_33", cfl=0.5):
    import ryujin_b200 as rj

    ti = rj.TimeIntegrator(module, getattr(rj.TimeSteppingScheme, scheme), cfl, cfl, rj.CFLRecoveryStrategy.none)
    sim.configure_time_integrator(scheme.replace("_", " "), cfl, cfl, "none")
    state = module.state_vector()
    state.interpolate(0.0)
    sim.set_U(state.download(), 0)
    t = t_ref = 0.0
    taus = []
    for _ in range(n_cycles):
        tau = ti.step(state, t)
        tau_ref = sim.time_step(t_ref)
        taus.append((tau, tau_ref))
        t += tau
        t_ref += tau_ref
    return state.download(), sim.get_U(0), np.array(taus), t


def test_sw_smooth_vortex_100_cycles(oracle):
    """smooth flow: the 1e-12 bar of north_star applies outright."""
    od, module, off, sim = gpu_and_oracle(oracle, "sw_vortex2d", [48, 48])
    U, U_ref, taus, _ = run_cycles(module, sim, 100, cfl=0.25)
    assert np.max(np.abs(taus[:, 0] - taus[:, 1]) / taus[:, 1]) < 1e-13
    assert rel_linf(U, U_ref) < 1e-12


def test_sw_circular_dam_break_config5(oracle):
    """BASELINE config 5 (2D dam break) at reduced size: a shock + a limiter that is active
    everywhere.  The precomputed h_star = pow(h, 4/3) enters only through the friction term,
    which is off here, so no pow() noise: GPU and oracle agree to round-off."""
    od, module, off, sim = gpu_and_oracle(oracle, "sw_dam2d", [64, 64])
    U, U_ref, taus, t = run_cycles(module, sim, 60, cfl=0.5)
    assert np.max(np.abs(taus[:, 0] - taus[:, 1]) / taus[:, 1]) < 1e-12
    err = rel_linf(U, U_ref)
    print("sw dam break 60 cycles: t = %.4f, rel_linf = %.3e" % (t, err))
    assert err < 1e-10
    assert U[:, 0].min() > 0.1 and U[:, 0].max() < 2.6
    # symmetric problem, symmetric mesh: the solution keeps the x <-> -x symmetry
    x = od.host.positions
    order = np.lexsort((x[:, 0], x[:, 1]))
    h = U[order, 0].reshape(65, 65)
    assert np.max(np.abs(h - h[:, ::-1])) < 1e-8  # round-off, amplified at the shock


def test_sw_ritter_dam_break_dry_states(oracle):
    """1D dam break over a dry bed (kinetic-energy limiter, dry-state cut-offs)."""
    od, module, off, sim = gpu_and_oracle(oracle, "sw_ritter1d", [800])
    U, U_ref, taus, t = run_cycles(module, sim, 150, cfl=0.5)
    assert np.max(np.abs(taus[:, 0] - taus[:, 1]) / taus[:, 1]) < 1e-12
    assert rel_linf(U, U_ref) < 1e-10
    assert U[:, 0].min() >= 0.0


def test_sw_smooth_vortex_verification_golden_on_gpu(oracle):
    """tests/shallow_water/verification-smooth_vortex-erk33-l6.{prm,output} with the update
    on the GPU; error norms evaluated by the oracle's quadrature on the downloaded state."""
    import ryujin_b200 as rj

    od, module = make_sw_product("sw_vortex2d", [64, 64])
    assert od.host.n_owned == 4225
    n = 64
    from oracle import orc

    off = oracle.assemble([n, n], [-6.0, -6.0], [12.0 / n, 12.0 / n], [orc.DIRICHLET] * 4)
    c = SW_CASES["sw_vortex2d"]
    prm = dict(c["eq"])
    prm.update(c["mp"])
    sim = off.sw_sim(kind=orc.SW_SMOOTH_VORTEX, direction=(1.0, 1.0), position=(-1.0, -1.0), p=(2.0, 1.0, 2.0), **prm)
    ti = rj.TimeIntegrator(module, rj.TimeSteppingScheme.erk_33, 0.25, 0.25, rj.CFLRecoveryStrategy.none)
    state = module.state_vector()
    state.interpolate(0.0)
    t = 0.0
    while t < 2.0:
        t += ti.step(state, t)
    # product rows are in the generator's local numbering; the oracle mesh is lexicographic
    U = np.zeros((od.host.n_owned, 3))
    U[od.host.global_index[: od.host.n_owned].astype(np.int64)] = state.download()
    sim.set_U(U, 0)
    err = sim.errors(t)
    ref = dict(t=2.001005728507654, Linf=0.03571394823661699, L1=0.0006325612013505061, L2=0.003420776846038435)
    assert abs(t - ref["t"]) <= 1e-12 * ref["t"]
    for key in ("Linf", "L1", "L2"):
        assert abs(err[key] - ref[key]) <= 1e-8 * ref[key], (key, err[key], ref[key])


SW_CASES.update({
    # supersonic (Froude 3) uniform flow through `dynamic` boundaries, both velocity limiters off
    # (the reference then returns t_l = 0 after the depth limiter, limiter.template.h:144-146)
    "sw_supercritical": dict(
        dim=2, bottom_left=[0.0, 0.0], top_right=[2.0, 1.0], boundary=[B_DYNAMIC] * 4,
        iv=dict(kind="sw_uniform", direction=[1.0, 0.2], position=[0.0, 0.0], sw_params=[1.0, 9.5]),
        eq=dict(gravity=9.81, reference_water_depth=1.0),
        mp=dict(limit_on_kinetic_energy=0, limit_on_square_velocity=0),
    ),
})


def test_sw_uniform_supercritical_dynamic_boundaries_and_limiters_off(oracle):
    od, module, off, sim = gpu_and_oracle(oracle, "sw_supercritical", [24, 12], bathymetry=True)
    old, new = module.state_vector(), module.state_vector()
    old.interpolate(0.0)
    U0 = old.download()
    assert np.all(U0[:, 0] == 1.0)
    x = od.host.positions
    U0[:, 0] *= 1.0 + 0.1 * np.sin(5.0 * x[:, 0]) * np.cos(3.0 * x[:, 1])
    old.upload(U0)
    sim.set_U(U0, 0)
    module.prepare_state_vector(old, 0.0)
    sim.prepare_state_vector(0, 0.0)
    assert rel_linf(old.download(), sim.get_U(0)) < 4e-15
    module.cfl = 0.5
    sim.set_cfl(0.5)
    tau = module.step(old, [], [], new)
    tau_ref = sim.step(0, [], [], 1)
    assert abs(tau - tau_ref) <= 1e-14 * tau_ref
    mask = np.ones(od.host.nnz, dtype=bool)
    mask[od.host.row_starts[:-1].astype(np.int64)] = False
    assert np.all(module.scratch("lij")[mask] == 0.0) and np.all(sim.scratch("lij")[mask] == 0.0)  # the quirk
    assert rel_linf(new.download(), sim.get_U(1)) < 1e-13
    assert (module.n_restarts(), module.n_warnings()) == sim.counters()


def test_sw_limiter_failure_counts_like_the_oracle(oracle):
    """an over-aggressive CFL number drives the low-order update out of the bounds: the `success`
    flags of the depth / kinetic energy / velocity limiters (limiter.template.h:60-70,203-215,
    349-360) turn into warnings exactly as in the oracle"""
    od, module, off, sim = gpu_and_oracle(oracle, "sw_dam2d_friction", [20, 16], bathymetry=True)
    old, new = module.state_vector(), module.state_vector()
    old.interpolate(0.0)
    sim.set_U(old.download(), 0)
    module.prepare_state_vector(old, 0.0)
    sim.prepare_state_vector(0, 0.0)
    module.cfl = 6.0
    sim.set_cfl(6.0)
    module.step(old, [], [], new)
    sim.step(0, [], [], 1)
    assert (module.n_restarts(), module.n_warnings()) == sim.counters() == (0, 1)
    assert rel_linf(new.download(), sim.get_U(1)) < 1e-12
