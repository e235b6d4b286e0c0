"""End-to-end pin of the oracle: the reference's isentropic-vortex verification runs
(tests/euler/verification-isentropic_vortex-2d-{erk33,ssprk33}-l{5,6}.{prm,output}).

The mesh is reproducible (uniform Q1 on [-5,5]^2) but not deal.II's DoF numbering,
so summation order differs: agreement is expected at ~1e-10 relative (the reference's
own mpirun=4 vs mpirun=8 goldens differ at that level), not bitwise.
"""
import numpy as np
import pytest

from conftest import load_golden
from oracle import orc


def run_vortex(oracle, refinement, scheme):
    n = 2**refinement
    off = oracle.assemble([n, n], [-5.0, -5.0], [10.0 / n, 10.0 / n], [orc.DIRICHLET] * 4)
    sim = off.euler_sim(kind=orc.ISENTROPIC_VORTEX, direction=(1.0, 1.0), position=(-1.0, -1.0), mach=1.0, beta=5.0, gamma=1.4)
    sim.configure_time_integrator(scheme=scheme, cfl_min=0.2, cfl_max=0.2, recovery="none")
    sim.interpolate(0.0)
    t, t_final = 0.0, 2.0
    cycles = 0
    while t < t_final:  # time_loop.template.h:357 (final time not enforced)
        t += sim.time_step(t)
        cycles += 1
    return off, sim, t, cycles


@pytest.mark.parametrize("scheme,key", [("erk 33", "erk33"), ("ssprk 33", "ssprk33")])
@pytest.mark.parametrize("refinement", [5, 6, 7])
def test_isentropic_vortex_verification(oracle, scheme, key, refinement):
    g = load_golden("euler_verification.json")
    ref = [r for r in g["runs"] if r["scheme"] == key and r["refinement"] == refinement][0]
    off, sim, t, cycles = run_vortex(oracle, refinement, scheme)
    assert off.n == ref["dofs"]
    err = sim.errors(t)
    # final time depends on every tau_max, i.e. on every d_ij of every step
    assert abs(t - ref["t"]) <= 4e-15 * ref["t"], (t, ref["t"])
    for name in ("Linf", "L1", "L2"):
        assert abs(err[name] - ref[name]) <= (1e-10 if refinement == 7 else 1e-11) * ref[name], (name, err[name], ref[name])
    assert sim.counters() == (0, 0)


def test_erk54_is_consistent_with_erk33(oracle):
    """erk 54 (time_integrator.template.h:446-510) has no Euler golden in the reference; pin the
    restated coefficients through consistency: same vortex, same cfl per substep -> 5/3 fewer
    cycles to the same final time, and error norms within 1 % of erk 33's."""
    from oracle import orc

    res = {}
    for scheme in ("erk 33", "erk 54"):
        n = 32
        off = oracle.assemble([n, n], [-5.0, -5.0], [10.0 / n, 10.0 / n], [orc.DIRICHLET] * 4)
        sim = off.euler_sim(kind=orc.ISENTROPIC_VORTEX, direction=(1.0, 1.0), position=(-1.0, -1.0), mach=1.0, beta=5.0)
        sim.configure_time_integrator(scheme, 0.2, 0.2, "none")
        sim.interpolate(0.0)
        t, cycles = 0.0, 0
        while t < 1.0:
            t += sim.time_step(t)
            cycles += 1
        res[scheme] = (cycles, t, sim.errors(t))
    assert res["erk 33"][0] == 45 and res["erk 54"][0] == 27
    assert abs(res["erk 33"][1] - res["erk 54"][1]) < 1e-8
    for k in ("Linf", "L1", "L2"):
        assert abs(res["erk 54"][2][k] / res["erk 33"][2][k] - 1.0) < 1e-2


def space_averaged_primitive(U, mi, gamma=1.4):
    """Quantities::internal_accumulate (quantities.template.h:369-423) for the interior manifold
    with level set 0 (every node): lumped-mass weighted average of the primitive state and of
    its second moments"""
    rho = U[:, 0]
    v = U[:, 1:-1] / rho[:, None]
    p = (gamma - 1.0) * (U[:, -1] - 0.5 * rho * np.sum(v * v, axis=1))
    prim = np.column_stack([rho, v, p])
    m = mi / np.sum(mi)
    return np.concatenate([m @ prim, m @ (prim * prim)])


def test_mass_conservation_time_series(oracle):
    """tests/euler/check-mass-conservation_01: Mach 3 uniform state in a closed box with slip walls
    (shocks form at the right wall and reflect), 65^2 dofs, ssprk 33 at cfl 0.9.  The golden holds
    t and the space averaged primitive state at the start of each cycle (19 rows)."""
    from oracle import orc

    g = load_golden("euler_mass_conservation.json")
    n = 64
    off = oracle.assemble([n, n], [0.0, 0.0], [20.0 / n, 20.0 / n], [orc.SLIP] * 4)
    assert off.n == 4225
    sim = off.euler_sim(kind=orc.UNIFORM, direction=(1.0, 0.0), position=(0.0, 0.0), primitive=(1.4, 3.0, 1.0))
    sim.configure_time_integrator("ssprk 33", 0.9, 0.9, "none")
    sim.interpolate(0.0)
    mi = off.arrays()["mi"]
    t = 0.0
    rows = []
    while True:
        rows.append(np.concatenate([[t], space_averaged_primitive(sim.get_U(0).reshape(-1, 4), mi)]))
        if t >= 0.5:
            break
        t += sim.time_step(t)
    ref = np.array(g["rows"])
    got = np.array(rows)
    assert got.shape == ref.shape == (19, 9)
    assert np.max(np.abs(got[:, 0] - ref[:, 0])) <= 2e-14  # the tau sequence of all cycles
    # mass is conserved to round-off in a closed box
    assert np.max(np.abs(got[:, 1] - 1.4)) < 2e-14
    scale = np.maximum(np.abs(ref), 1e-3)  # <v_2> is round-off noise around zero
    err = np.abs(got - ref) / scale
    assert np.max(err[:, [1, 2, 4, 5, 6, 8]]) < 1e-12, np.max(err, axis=0)
    assert np.max(np.abs(got[:, 3])) < 1e-14 and np.max(err[:, 7]) < 1e-9


@pytest.mark.parametrize(
    "name,kind,position,gamma,t_final,ref",
    [
        ("leblanc", "LEBLANC", 0.326732673267, 1.66666666666667, 0.66666666666667,
         dict(t=0.6666710022008712, Linf=0.2233117555967211, L1=0.01167724325117752, L2=0.03259102542735855)),
        ("rarefaction", "RAREFACTION", 0.2, 1.4, 0.30558,
         dict(t=0.3056208093406663, Linf=0.001236294603917117, L1=2.234263733795138e-05, L2=8.123947460631401e-05)),
    ],
)
def test_euler_1d_verification_decks(oracle, name, kind, position, gamma, t_final, ref):
    """tests/euler/verification-{leblanc,rarefaction}-1d-erk33-l6.{prm,mpirun=4.output}: 1D Euler,
    1601 dofs, cfl 0.1, limiter relaxation factor 8, 8153 / 7137 cycles (Le Blanc: a 1e9 pressure
    ratio).  The restatement reproduces the number of cycles, the final time to 1.3e-8 / 5e-10
    and the error norms to 2e-6 / 2e-5 relative.  This residual is the same for the libm, VCL and
    FMA builds of the oracle and is not explained by mesh coordinates (a 1e-12 perturbation of the
    initial state moves t by 2e-9: an effective 3e-13 perturbation would account for it); the
    goldens come from a 4-rank run.  Weaker pins than the 2D decks, asserted as such."""
    from oracle import orc

    n = 25 * 2**6
    off = oracle.assemble([n], [0.0], [1.0 / n], [orc.DIRICHLET] * 2)
    assert off.n == 1601
    sim = off.euler_sim(kind=getattr(orc, kind), direction=(1.0,), position=(position,), gamma=gamma,
                        limiter_relaxation_factor=8.0)
    sim.configure_time_integrator("erk 33", 0.1, 0.1, "none")
    sim.interpolate(0.0)
    t, cycles = 0.0, 0
    while t < t_final:
        t += sim.time_step(t)
        cycles += 1
    err = sim.errors(t)
    print(name, cycles, t, err)
    assert abs(t - ref["t"]) <= 2e-8 * ref["t"]
    for k in ("Linf", "L1", "L2"):
        assert abs(err[k] - ref[k]) <= 3e-5 * ref[k], (k, err[k], ref[k])
    assert sim.counters() == (0, 0)
