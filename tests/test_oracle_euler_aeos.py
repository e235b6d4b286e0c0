"""The oracle's restatement of the `euler_aeos` equation module (oracle/orc_euler_aeos.h) against the
reference's own known-answer vectors (tests/golden/aeos_reference.json, extracted from
/root/reference/tests/euler_aeos/*.output by tests/golden/extract_reference_goldens_aeos.py)."""
import numpy as np
import pytest

from conftest import load_golden
from oracle import orc


@pytest.fixture(scope="module")
def golden():
    return load_golden("aeos_reference.json")


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def test_riemann_solver_goldens(oracle, golden):
    """tests/euler_aeos/riemann_solver{,-strict}.output: 44 cases, ideal gas and variable gamma, with covolume"""
    assert len(golden["riemann"]) == 44
    for case in golden["riemann"]:
        r = oracle.aeos_riemann(case["state_i"], case["state_j"], eos=orc.AEOS_VAN_DER_WAALS, covolume_b=case["covolume_b"],
                                compute_strict_bounds=int(case["strict"]))
        assert rel(r["a_left"], case["a_left"]) < 1e-15 and rel(r["a_right"], case["a_right"]) < 1e-15
        assert rel(r["lambda_max"], case["lambda_max"]) < 2e-15, case
        assert rel(r["p_2"], case["p_2"]) < 1e-10, case
        if "p_star_RS_1" in case:
            assert rel(r["p_star_RS"], min(case["p_star_RS_1"], case["p_star_RS_2"])) < 1e-10  # exponents up to 4e4 amplify pow rounding
            assert rel(r["p_star_SS"], min(case["p_star_SS_1"], case["p_star_SS_2"])) < 1e-10
        if "p_star_interpolated" in case:
            assert rel(r["p_star_interpolated"], case["p_star_interpolated"]) < 1e-10


def test_limiter_goldens(oracle, golden):
    """tests/euler_aeos/limiter.output (van der Waals covolumes) and limiter-NASG.output; both are built with
    EXPENSIVE_BOUNDS_CHECK"""
    assert len(golden["limiter"]) == 34
    for case in golden["limiter"]:
        if case["eos"] == "van der waals":
            prm = dict(eos=orc.AEOS_VAN_DER_WAALS, covolume_b=case["covolume_b"])
        else:
            nasg = case["covolume_b"] != 0.0  # the parameters are only set once set_covolume() ran
            prm = dict(eos=orc.AEOS_NOBLE_ABEL_STIFFENED_GAS, covolume_b=case["covolume_b"],
                       reference_pressure=0.1 if nasg else 0.0, reference_specific_internal_energy=0.1 if nasg else 0.0)
        r = oracle.aeos_limiter(case["U"], case["P"], case["bounds"], expensive_bounds_check=1, **prm)
        assert rel(r["specific_entropy"], case["specific_entropy"]) < 1e-15, case
        assert abs(r["l"] - case["l"]) <= 1e-13 * max(1.0, abs(case["l"])), (case, r)  # Newton steps close to maximal compressibility amplify pow rounding
        assert r["success"] == case["success"], (case, r)


def test_hyperbolic_system_goldens(oracle, golden):
    """tests/euler_aeos/hyperbolic_system.output (10 printed digits)"""
    assert len(golden["system"]) >= 9
    for case in golden["system"]:
        dim, gamma = case["dim"], case["density"]
        U = np.zeros(dim + 2)
        U[0] = case["density"]
        U[1:1 + dim] = case["momentum"]
        U[dim + 1] = case["total_energy"]
        q = oracle.aeos_system_quantities(dim, U, case["surrogate_gamma"] if False else gamma, eos=orc.AEOS_VAN_DER_WAALS,
                                          covolume_b=case["covolume_b"])
        tol = 2e-10
        assert rel(q["internal_energy"], case["internal_energy"]) < tol
        assert np.allclose(q["internal_energy_derivative"], case["internal_energy_derivative"], rtol=tol, atol=1e-12)
        assert rel(q["specific_entropy"], case["specific_entropy"]) < tol
        assert rel(q["harten_entropy"], case["harten_entropy"]) < tol
        assert np.allclose(q["harten_entropy_derivative"], case["harten_entropy_derivative"], rtol=tol, atol=1e-12)
        assert rel(q["surrogate_pressure"], case["surrogate_pressure"]) < tol
        assert rel(q["surrogate_gamma"], case["surrogate_gamma"]) < tol
        assert np.allclose(q["f"], case["f"], rtol=tol, atol=1e-12)


@pytest.mark.parametrize("scheme,key", [("erk 33", "erk33"), ("ssprk 33", "ssprk33")])
def test_verification_isentropic_vortex_pge(oracle, golden, scheme, key):
    """tests/euler_aeos/verification-isentropic_vortex-pge-2d-*-l5.output end to end through the oracle's own mesh,
    module (two precomputation cycles) and time integrator: polytropic gas equation of state, strict bounds"""
    ref = [r for r in golden["verification"] if r["scheme"] == key and r["refinement"] == 5][0]
    n = 32
    off = oracle.assemble([n, n], [-5.0, -5.0], [10.0 / n, 10.0 / n], [orc.DIRICHLET] * 4)
    assert off.n == ref["dofs"]
    sim = off.aeos_sim(kind=orc.ISENTROPIC_VORTEX, direction=(1.0, 1.0), position=(-1.0, -1.0), mach=1.0, beta=5.0)
    sim.configure_time_integrator(scheme, 0.2, 0.2, "none")
    sim.interpolate(0.0, 0)
    t = 0.0
    while t < 2.0:
        t += sim.time_step(t)
    assert abs(t - ref["t"]) <= 1e-13 * ref["t"], (t, ref["t"])
    err = sim.errors(t)
    for name in ("Linf", "L1", "L2"):
        assert abs(err[name] - ref[name]) <= 1e-10 * ref[name], (name, err[name], ref[name])
