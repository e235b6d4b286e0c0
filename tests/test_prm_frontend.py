"""the reference's parameter files drive the B200 path unmodified and the
run prints the reference's output block.  CPU part: parser, error norms, mapping of the sections
up to the device boundary.  GPU part: the reference's end-to-end verification decks reproduce
their golden outputs."""
import io
import re

import numpy as np
import pytest

from conftest import load_golden


def test_parser_syntax():
    from ryujin_b200 import prm

    text = """
    # a comment
    subsection A - TimeLoop
      set basename   =  test run   # trailing comment
      set final   time = 2.5
      set enable compute error = true
      subsection  nested   section
        set list = 1, 2.5 , -3
        set empty =
      end
    end
    subsection B - Equation
      set dimension = 3
      set long = a \\
                 b
    end
    """
    p = prm.parse(text)
    A = p.subsection("A - TimeLoop")
    assert A.get_str("basename", "x") == "test run"
    assert A.get_float("final time", 0.0) == 2.5  # white space in names is collapsed
    assert A.get_bool("enable compute error", False) is True
    assert A.get_bool("not there", False) is False
    N = A.subsection("nested section")
    assert N.get_list("list", []) == [1.0, 2.5, -3.0]
    assert N.get_list("empty", [7.0]) == []
    B = p.subsection("B - Equation")
    assert B.get_int("dimension", 2) == 3
    assert B.get_str("long", "") == "a b"
    assert "B - Equation/long" not in p.unused() and p.unused() == []
    with pytest.raises(prm.ParameterError):
        prm.parse("subsection A\n set x = 1\n")
    with pytest.raises(prm.ParameterError):
        prm.parse("end\n")
    with pytest.raises(prm.ParameterError):
        prm.parse("subsection A\n what is this\nend\n")
    with pytest.raises(prm.ParameterError):
        prm.parse("subsection A\n set n = abc\nend\n").subsection("A").get_int("n", 1)


def test_reference_decks_parse_and_report_unused():
    from ryujin_b200 import prm

    for run in load_golden("prm_runs.json")["runs"]:
        p = prm.parse(run["prm"])
        assert "A - TimeLoop" in p.subsections and "B - Equation" in p.subsections, run["name"]
        assert p.subsection("A - TimeLoop").get_bool("enable compute error", False)
        assert set(p.unused()) >= {"A - TimeLoop/final time"}  # nothing else has been read yet


def test_q1_norms_exact_for_multilinear_functions():
    from ryujin_b200.time_loop import q1_norms

    # f = x y on [0,1]^2: the Q1 interpolant is exact, Gauss(3) integrates |f| and f^2 exactly
    n = [7, 5]
    x = np.linspace(0.0, 1.0, n[0] + 1)
    y = np.linspace(0.0, 1.0, n[1] + 1)
    f = (y[:, None] * x[None, :]).ravel()
    linf, l1, l2 = q1_norms(f, n, [1.0 / n[0], 1.0 / n[1]])
    assert linf == 1.0 and abs(l1 - 0.25) < 1e-15 and abs(l2 - 1.0 / 3.0) < 1e-15
    # 1D and 3D, sign changes inside cells are handled by the quadrature like deal.II does
    linf, l1, l2 = q1_norms(np.linspace(-1.0, 1.0, 11), [10], [0.2])
    assert abs(l1 - 1.0) < 1e-15 and abs(l2 - np.sqrt(2.0 / 3.0)) < 1e-15
    z = np.linspace(0.0, 2.0, 4)
    g = (z[:, None, None] * np.ones((1, 3, 5))).ravel()
    linf, l1, l2 = q1_norms(g, [4, 2, 3], [0.25, 0.5, 2.0 / 3.0])
    assert abs(l1 - 2.0) < 1e-14 and abs(l2 - np.sqrt(8.0 / 3.0)) < 1e-14


def test_q1_norms_match_the_oracle_quadrature(oracle):
    """same norms as the oracle's restatement of TimeLoop::compute_error (pinned against the
    reference's verification goldens in test_oracle_verification.py)"""
    from oracle import orc
    from ryujin_b200.time_loop import q1_norms

    n = 12
    off = oracle.assemble([n, n], [-5.0, -5.0], [10.0 / n, 10.0 / n], [orc.DIRICHLET] * 4)
    sim = off.euler_sim(kind=orc.ISENTROPIC_VORTEX, direction=(1.0, 1.0), position=(-1.0, -1.0), mach=1.0, beta=5.0)
    sim.interpolate(0.0, 0)
    sim.prepare_state_vector(0, 0.3)  # compute_error applies the boundary conditions at t first
    U = sim.get_U(0).reshape(-1, 4).copy()
    sim.interpolate(0.3, 1)
    A = sim.get_U(1).reshape(-1, 4)
    err = sim.errors(0.3)
    mine = np.zeros(3)
    for c in range(4):
        e = np.array(q1_norms(U[:, c] - A[:, c], [n, n], [10.0 / n] * 2))
        a = np.array(q1_norms(A[:, c], [n, n], [10.0 / n] * 2))
        mine += e / a
    for k, v in zip(("Linf", "L1", "L2"), mine):
        assert abs(v - err[k]) <= 1e-13 * err[k], (k, v, err[k])


def test_time_loop_maps_the_sections_and_rejects_what_is_off_the_path():
    import ryujin_b200 as rj
    from ryujin_b200 import prm

    deck = load_golden("prm_runs.json")["runs"][0]["prm"]
    # everything up to the device boundary is host logic: on a CPU box it must end in NoDeviceError
    if rj.device_count() == 0:
        with pytest.raises(rj.NoDeviceError):
            rj.TimeLoop(prm.parse(deck), out=io.StringIO())
    for bad, exc in (
        (deck.replace("set equation  = euler", "set equation  = navier stokes"), rj.NotOnThePath),
        (deck.replace("set geometry        = rectangular domain", "set geometry        = cylinder"), rj.NotOnThePath),
        (deck.replace("set time stepping scheme  = erk 33", "set time stepping scheme  = strang erk 33 cn"), rj.NotOnThePath),
        (deck.replace("set boundary condition top    = dirichlet", "set boundary condition top    = sticky"), prm.ParameterError),
        (deck.replace("set enable compute error      = true", "set enable checkpointing = true"), rj.NotOnThePath),
    ):
        assert bad != deck
        with pytest.raises(exc):
            rj.TimeLoop(prm.parse(bad), out=io.StringIO())


@pytest.mark.gpu
@pytest.mark.parametrize("index", range(10))
def test_reference_decks_reproduce_their_golden_outputs(index):
    """`python -m ryujin_b200 deck.prm` on the reference's verification decks: #dofs exact, final
    time to 1e-12, normalized Linf/L1/L2 errors to 1e-8 (Ritter dam break: inside the round-off
    spread of that case, see test_oracle_shallow_water.py)."""
    import ryujin_b200 as rj
    from ryujin_b200 import prm

    run = load_golden("prm_runs.json")["runs"][index]
    out = io.StringIO()
    loop = rj.TimeLoop(prm.parse(run["prm"]), out=out)
    result = loop.run()
    text = out.getvalue()
    exp = run["expected"]
    got = {k: float(re.search(r"^%s\s*=\s*(\S+)" % re.escape(k), text, re.M).group(1)) for k in ("#dofs", "t", "Linf", "L1", "L2")}
    assert got["#dofs"] == exp["dofs"] == result["n_dofs"]
    assert abs(got["t"] - exp["t"]) <= 1e-12 * exp["t"]
    tol = 2e-4 if "ritter" in run["name"] else 1e-8
    for k in ("Linf", "L1", "L2"):
        assert abs(got[k] - exp[k]) <= tol * exp[k], (run["name"], k, got[k], exp[k])
    # the informational lines of the golden output are reproduced verbatim
    for line in run["output"].splitlines():
        if line.startswith("[INFO]") or line.startswith("Normalized"):
            assert line.strip() in [l.strip() for l in text.splitlines()], line
    assert loop.unused_entries() == []


@pytest.mark.gpu
def test_mass_conservation_deck_reproduces_the_golden_time_series():
    """tests/euler/check-mass-conservation_01.{prm,output} unmodified: Mach 3 flow in a closed box
    with slip walls, ssprk 33 at cfl 0.9; the golden is the space averaged primitive state at the
    start of every cycle (K - Quantities, interior manifold with level set 0)."""
    import ryujin_b200 as rj
    from ryujin_b200 import prm

    g = load_golden("euler_mass_conservation.json")
    out = io.StringIO()
    loop = rj.TimeLoop(prm.parse(g["prm"]), out=out)
    result = loop.run()
    ref = np.array(g["rows"])
    got = np.array([[t] + list(v) for t, v in result["time_series"]])
    assert got.shape == ref.shape == (19, 9)
    assert np.max(np.abs(got[:, 0] - ref[:, 0])) <= 1e-12
    assert np.max(np.abs(got[:, 1] - 1.4)) < 2e-14  # mass conservation
    scale = np.maximum(np.abs(ref), 1e-3)
    err = np.abs(got - ref) / scale
    assert np.max(err[:, [1, 2, 4, 5, 6, 8]]) < 1e-10, np.max(err, axis=0)
    # and the printed dump parses back to the same numbers, in the reference's layout
    lines = [l for l in out.getvalue().splitlines() if l and not l.startswith("[") and not l.startswith("#")]
    assert out.getvalue().count("# time t\tprimitive state (rho, v_1, v_2, p)\t and 2nd moments") == 1
    assert len(lines) == 19 and all(len(l.split("\t")) == 3 for l in lines)
    assert abs(float(lines[5].split("\t")[0]) - ref[5, 0]) < 1e-13
