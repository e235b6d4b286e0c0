"""Pin the shallow water part of the oracle against the reference's
goldens: tests/shallow_water/riemann_solver.output (3 two-state cases, "verified against
Mathematica") and, end to end, tests/shallow_water/verification-{smooth_vortex-erk33-l6,
ritter_dam_break-erk33-l7}.{prm,output}."""
import re

import numpy as np
import pytest

from conftest import load_golden
from oracle import orc


def test_sw_riemann_solver_goldens(oracle):
    text = load_golden("shallow_water_riemann_solver.json")["text"]
    ref = [(float(a), float(b)) for a, b in re.findall(r"lambda_max: (\S+)\nh_star: (\S+)", text)]
    relax = [tuple(float(x) for x in m) for m in re.findall(r"relaxation: (\S+) (\S+)", text)]
    cases = [((0.0, 0.0), (0.0, 0.0)), ((1.0, 1.0), (0.0, 0.0)), ((1.8, 0.0), (1.0, 0.0))]
    assert len(ref) == 3
    for n, (si, sj) in enumerate(cases):
        r = oracle.sw_riemann(si, sj)
        assert (r["h_i"], r["u_i"]) == relax[2 * n]
        assert (r["h_j"], r["u_j"]) == relax[2 * n + 1]
        assert abs(r["lambda_max"] - ref[n][0]) <= 2e-16 * abs(ref[n][0])
        assert abs(r["h_star"] - ref[n][1]) <= 2e-16 * abs(ref[n][1])


def test_sw_smooth_vortex_verification(oracle):
    """tests/shallow_water/verification-smooth_vortex-erk33-l6: 65^2 dofs on [-6,6]^2, dirichlet,
    square-velocity limiter, cfl 0.25, ERK33, final time 2."""
    n = 64
    off = oracle.assemble([n, n], [-6.0, -6.0], [12.0 / n, 12.0 / n], [orc.DIRICHLET] * 4)
    assert off.n == 4225
    sim = off.sw_sim(kind=orc.SW_SMOOTH_VORTEX, direction=(1.0, 1.0), position=(-1.0, -1.0), p=(2.0, 1.0, 2.0),
                     gravity=9.81, reference_water_depth=2.0, dry_state_relaxation_factor=0.0,
                     limit_on_kinetic_energy=0, limit_on_square_velocity=1)
    sim.configure_time_integrator("erk 33", 0.25, 0.25, "none")
    sim.interpolate(0.0)
    t = 0.0
    while t < 2.0:
        t += sim.time_step(t)
    err = sim.errors(t)
    ref = dict(t=2.001005728507654, Linf=0.03571394823661699, L1=0.0006325612013505061, L2=0.003420776846038435)
    print(t, err)
    assert abs(t - ref["t"]) <= 1e-12 * ref["t"]
    for k in ("Linf", "L1", "L2"):
        assert abs(err[k] - ref[k]) <= 1e-8 * ref[k], (k, err[k], ref[k])


def _ritter(o):
    n = 25 * 2**7
    off = o.assemble([n], [0.0], [10.0 / n], [orc.DIRICHLET] * 2)
    assert off.n == 3201
    sim = off.sw_sim(kind=orc.SW_RITTER_DAM_BREAK, direction=(1.0,), position=(5.0,), p=(1.0, 0.005),
                     gravity=9.81, reference_water_depth=0.005, dry_state_relaxation_factor=1.0e-3,
                     limit_on_kinetic_energy=1, limit_on_square_velocity=0)
    sim.configure_time_integrator("erk 33", 0.5, 0.5, "none")
    sim.set_error_components([0])
    sim.interpolate(0.0)
    t = 0.0
    while t < 6.0:
        t += sim.time_step(t)
    return t, sim.errors(t)


def test_sw_ritter_dam_break_verification(oracle):
    """tests/shallow_water/verification-ritter_dam_break-erk33-l7: 1D, 25 * 2^7 cells on [0,10],
    kinetic-energy limiter, dam break over a dry bed, error in h only, 1246 cycles.

    The final time (i.e. the CFL-limiting wave speeds of all cycles) is reproduced to 1e-13.  The
    error norms of this case are sensitive to round-off at the 1e-4 level (wet/dry front): the same
    restatement built with FP contraction (what GCC does for the reference on an AVX2/FMA host,
    oracle/liboracle_fma.so) moves them by 1e-4 and lands within 1e-5 of the golden, the
    contraction-free build within 1.1e-4.  The golden therefore pins the restatement to within the
    spread of two builds of the same sources, which is what is asserted."""
    ref = dict(t=6.002065094435738, Linf=0.001137142985470178, L1=2.034821089657167e-05, L2=6.232030880580931e-05)
    t, err = _ritter(oracle)
    t_fma, err_fma = _ritter(orc.Oracle("fma"))
    print(t, err, t_fma, err_fma)
    assert abs(t - ref["t"]) <= 1e-13 * ref["t"]
    assert abs(t_fma - ref["t"]) <= 1e-13 * ref["t"]
    for k in ("Linf", "L1", "L2"):
        spread = abs(err[k] - err_fma[k])
        assert abs(err_fma[k] - ref[k]) <= 2e-5 * ref[k], (k, err_fma[k], ref[k])
        assert abs(err[k] - ref[k]) <= 2.0 * spread + 1e-8 * ref[k], (k, err[k], err_fma[k], ref[k])
        assert abs(err[k] - ref[k]) <= 2e-4 * ref[k], (k, err[k], ref[k])


def test_sw_steady_incline_with_friction_verification(oracle):
    """tests/shallow_water/verification-steady_incline-erk33-l9.{prm,output}: steady flow down an
    inclined plane (bathymetry Z = -slope x) balanced by Manning friction, `dynamic` boundaries,
    513 dofs, 81 ERK33 cycles.  Pins what the other goldens do not touch: star states and affine
    shift over a non-flat bathymetry, the friction source with its stage weights, and the
    Riemann-characteristic boundary condition.  The golden errors are round-off (Linf 2.4e-14):
    the scheme is well balanced; the final time pins the tau sequence."""
    n = 2**9
    off = oracle.assemble([n], [0.0], [20.0 / n], [orc.DYNAMIC] * 2)
    assert off.n == 513
    sim = off.sw_sim(kind=orc.SW_SLOPING_FRICTION, direction=(1.0,), position=(0.0,), p=(1.0e-2, 1.0e-1),
                     gravity=9.81, manning_friction_coefficient=1.0e-2, reference_water_depth=1.0,
                     dry_state_relaxation_large=1.0e4, dry_state_relaxation_small=1.0e4,
                     limit_on_kinetic_energy=0, limit_on_square_velocity=1)
    sim.configure_time_integrator("erk 33", 0.5, 0.5, "none")
    sim.set_error_components([0, 1])
    sim.interpolate(0.0)
    U0 = sim.get_U(0).reshape(-1, 2)[7].copy()
    t, cycles = 0.0, 0
    while t < 1.0:
        t += sim.time_step(t)
        cycles += 1
    ref = dict(t=1.000593578808362, Linf=2.388278346583212e-14, L1=4.287451614926996e-15, L2=5.452329602107318e-15)
    err = sim.errors(t)
    print(cycles, t, err)
    assert abs(t - ref["t"]) <= 1e-14 * ref["t"]
    # round-off sized errors: same magnitude as the reference's, digits are noise
    for k in ("Linf", "L1", "L2"):
        assert err[k] <= 3.0 * ref[k], (k, err[k], ref[k])
    # and the steady state really carries a bathymetry gradient and friction (not a trivial test)
    h0 = (1.0e-2**2 * 1.0e-1**2 / 1.0e-2) ** (1.0 / (2.0 + 4.0 / 3.0))
    assert abs(U0[0] - h0) < 1e-15 and abs(U0[1] - 0.1) < 1e-17
