"""The euler_aeos equation module on the CUDA path against the CPU oracle
(oracle/orc_euler_aeos.h, pinned to the reference's own test outputs by
tests/test_oracle_euler_aeos.py): the three analytic equations of state, both Riemann-solver /
limiter modes (compute strict bounds on / off), dim 1-3, every intermediate of one update and the
state after a few dozen SSPRK cycles.

Tolerances: the kernel source and the oracle execute the same IEEE operations in the same order with
the same pow (the reference's VCL, restated on the device): 1e-12 relative L-infinity on the state
as north_star asks; the intermediates are held to the tighter values written below.
"""
import numpy as np
import pytest

from helpers import AEOS_EOS as EOS, aeos_parameters, make_oracle, make_product, matching_oracle, rel_linf, rel_linf_per_component

pytestmark = pytest.mark.gpu

@pytest.fixture(scope="module")
def oracle():
    return matching_oracle()[0]


def pair(oracle, case, n_cells, eos, strict=True, **kw):
    eq, prm = aeos_parameters(eos, strict)
    od, module = make_product(case, n_cells, equation="euler_aeos", **eq, **kw)
    off, sim = make_oracle(oracle, case, od.host, equation="euler_aeos", **prm)
    return od, module, off, sim


def contrast_state(host, case, eos):
    """from_initial_state of the shock tube through the equation of state (numpy restatement of
    euler_aeos/hyperbolic_system.h:1441-1479 and the three specific_internal_energy functions)"""
    from helpers import CASES

    iv = CASES[case]["iv"]
    e = EOS[eos]
    gamma = e["gamma"]
    x = host.positions[:, 0] - iv["position"][0]
    left, right = np.array(iv["primitive"]), np.array(iv["primitive_right"])
    prim = np.where((x > 0.0)[:, None], right[None, :], left[None, :])
    rho, u, p = prim[:, 0], prim[:, 1], prim[:, 2]
    b = e.get("covolume_b", 0.0)
    if eos == "polytropic":
        sie = p / (rho * (gamma - 1.0))
    elif eos == "nasg":
        sie = e["reference_specific_internal_energy"] + (p + gamma * e["reference_pressure"]) * (1.0 - b * rho) / (rho * (gamma - 1.0))
    else:
        a = e["vdw_a"]
        sie = (p + a * rho * rho) * (1.0 - b * rho) / (rho * (gamma - 1.0)) - a * rho
    U = np.zeros((host.n_owned, host.dim + 2))
    U[:, 0] = rho
    U[:, 1] = rho * u
    U[:, -1] = rho * sie + 0.5 * rho * u * u
    return U


def test_sizes_and_rejected_boundary(oracle):
    import ryujin_b200 as rj

    od, module = make_product("sod2d", [8, 8], equation="euler_aeos")
    assert (module.k, module.n_precomputed, module.n_bounds) == (4, 4, 4)
    # `dynamic` boundaries trap in the reference (euler_aeos/hyperbolic_system.h:1315)
    with pytest.raises(rj.RyujinError, match="dynamic"):
        make_product("dynamic2d", [8, 8], equation="euler_aeos")


@pytest.mark.parametrize("strict", [True, False])
@pytest.mark.parametrize("eos", ["polytropic", "nasg", "vdw"])
@pytest.mark.parametrize("case,n_cells", [("sod1d", [96]), ("sod2d", [40, 24]), ("sod3d", [14, 8, 6])])
def test_single_update_intermediates(oracle, case, n_cells, eos, strict):
    """prepare_state_vector (both precomputation cycles) + step<0>: every array of the sweep"""
    od, module, off, sim = pair(oracle, case, n_cells, eos, strict)
    old, new = module.state_vector(), module.state_vector()
    old.interpolate(0.0)
    U0 = old.download()
    sim.set_U(U0, 0)
    assert rel_linf(U0, contrast_state(od.host, case, eos)) < 5e-16

    module.prepare_state_vector(old, 0.05)
    sim.prepare_state_vector(0, 0.05)
    assert rel_linf(old.download(), sim.get_U(0)) < 1e-15
    assert np.max(rel_linf_per_component(old.download_precomputed(), sim.get_prec(0))) < 1e-14

    module.cfl = 0.5
    sim.set_cfl(0.5)
    tau = module.step(old, [], [], new)
    tau_ref = sim.step(0, [], [], 1)
    assert abs(tau - tau_ref) <= 1e-14 * tau_ref

    k, nb = module.k, module.n_bounds
    assert np.max(np.abs(module.scratch("alpha") - sim.scratch("alpha"))) < 1e-12
    d_gpu, d_ref = module.scratch("dij"), sim.scratch("dij")
    assert np.max(np.abs(d_gpu - d_ref)) <= 1e-13 * np.max(np.abs(d_ref))
    assert np.max(rel_linf_per_component(module.scratch("r").reshape(-1, k), sim.scratch("r").reshape(-1, k))) < 1e-12
    assert np.max(rel_linf_per_component(module.scratch("bounds").reshape(-1, nb), sim.scratch("bounds").reshape(-1, nb))) < 1e-13
    diag = od.host.row_starts[:-1].astype(np.int64)
    mask = np.ones(od.host.nnz, dtype=bool)
    mask[diag] = False
    for name in ("lij", "lij_next"):
        a, b = module.scratch(name)[mask], sim.scratch(name)[mask]
        assert np.max(np.abs(a - b)) < 1e-9, name
        assert np.all((a >= 0.0) & (a <= 1.0))
    assert rel_linf(new.download(), sim.get_U(1)) < 1e-13
    assert (module.n_restarts(), module.n_warnings()) == sim.counters()


@pytest.mark.parametrize("eos,strict", [("polytropic", True), ("nasg", True), ("nasg", False), ("vdw", True), ("vdw", False)])
@pytest.mark.parametrize("case,n_cells", [("sod2d", [48, 24]), ("sod3d", [16, 8, 8])])
def test_cycles(oracle, case, n_cells, eos, strict):
    """30 SSPRK-33 cycles through the TimeIntegrator (graph replay) against the oracle's integrator"""
    import ryujin_b200 as rj

    od, module, off, sim = pair(oracle, case, n_cells, eos, strict)
    ti = rj.TimeIntegrator(module, rj.TimeSteppingScheme.erk_33, 0.45, 0.9)
    state = module.state_vector()
    state.interpolate(0.0)
    sim.set_U(state.download(), 0)
    sim.configure_time_integrator("erk 33", 0.45, 0.9)
    t = t_ref = 0.0
    for cycle in range(30):
        tau = ti.step(state, t)
        tau_ref = sim.time_step(t_ref)
        assert abs(tau - tau_ref) <= 1e-13 * tau_ref, cycle
        t += tau
        t_ref += tau_ref
    U, U_ref = state.download(), sim.get_U(0)
    assert np.all(np.isfinite(U))
    assert rel_linf(U, U_ref) < 1e-12
    assert (module.n_restarts(), module.n_warnings()) == sim.counters()


def test_polytropic_gas_agrees_with_the_euler_module_to_discretization_level(oracle):
    """sanity across modules: with the polytropic equation of state euler_aeos solves the same
    problem as euler (different Riemann solver estimate and limiter bounds, same scheme) -- the two
    shock tubes stay within a few percent after 20 cycles"""
    import ryujin_b200 as rj

    states = []
    for equation in ("euler", "euler_aeos"):
        od, module = make_product("sod2d", [64, 8], equation=equation)
        ti = rj.TimeIntegrator(module, rj.TimeSteppingScheme.erk_33, 0.45, 0.9)
        state = module.state_vector()
        state.interpolate(0.0)
        t = 0.0
        while t < 0.05:
            t += ti.step(state, t, t_final=0.05)
        states.append(state.download())
    assert np.max(np.abs(states[0][:, 0] - states[1][:, 0])) < 0.05
