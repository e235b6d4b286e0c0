#!/usr/bin/env python
"""Extract golden vectors for the hot path from the reference's own test suite.

Run in the authoring container (the reference tree is not on the GPU box):

    python tests/golden/extract_reference_goldens.py [/root/reference]

Reads the reference's *output* fixtures (never its sources) and writes small
JSON files next to this script:

  euler_riemann_solver.json     <- tests/euler/riemann_solver{,-iterated-2,-iterated-10}.output
  euler_limiter.json            <- tests/euler/limiter.output (+ the (U,P,bounds)
                                    inputs of tests/euler/limiter.cc:61-139,
                                    transcribed as data below)
  euler_hyperbolic_system.json  <- tests/euler/hyperbolic_system.output (double)
  sparse_matrix_simd.json       <- tests/common/sparse_matrix_simd.output{,.sse2,.avx2,.avx512}
  euler_verification.json       <- tests/euler/verification-isentropic_vortex-2d-*.output
  shallow_water_riemann_solver.json <- tests/shallow_water/riemann_solver.output
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))

FLOAT = r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eE][-+]?\d+)?|[-+]?nan|[-+]?inf"


def floats(s):
    return [float(x) for x in re.findall(FLOAT, s)]


def read(path):
    with open(os.path.join(REF, path)) as f:
        return f.read()


def dump(name, obj):
    with open(os.path.join(OUT, name), "w") as f:
        json.dump(obj, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", name)


# --------------------------------------------------------------------------
def riemann(path):
    text = read(path)
    gamma = floats(text.split("\n", 1)[0])[0]
    blocks = re.split(r"\n\n\n", text.split("\n", 2)[2])
    cases = []
    for b in blocks:
        lines = [l for l in b.strip().split("\n") if l.strip()]
        if len(lines) < 4:
            continue
        case = {"state_i": floats(lines[0]), "state_j": floats(lines[1])}
        for l in lines[2:]:
            for key, tag in [
                ("a_left", "a_left:"),
                ("a_right", "a_right:"),
                ("p_star_two_rarefaction", "p_star_two_rarefaction ="),
                ("p_star_failsafe", "p_star_failsafe ="),
                ("p_star_tilde", "p^*_tilde  ="),
                ("phi_p_star_tilde", "phi(p_*_t) ="),
                ("lambda_max", "-> lambda_max ="),
            ]:
                if l.strip().startswith(tag):
                    case[key] = floats(l.split(tag)[1])[0]
        # std::fixed is sticky after the first debug print in iterated runs
        case["fixed_format"] = "e" not in lines[-1]
        cases.append(case)
    return {"gamma": gamma, "cases": cases}


dump(
    "euler_riemann_solver.json",
    {
        "source": "tests/euler/riemann_solver*.output",
        "newton_0": riemann("tests/euler/riemann_solver.output"),
        "newton_2": riemann("tests/euler/riemann_solver-iterated-2.output"),
        "newton_10": riemann("tests/euler/riemann_solver-iterated-10.output"),
    },
)

# --------------------------------------------------------------------------
# inputs of tests/euler/limiter.cc:61-139 (dim = 1; U, P, bounds)
LIMITER_INPUTS = [
    ("Minimum density violation:", [0.8, 1.4, 3.0], [-0.1, 0.1, 0.1], [0.9, 1.1, 2.0]),
    ("Minimum density violation (eps):", [0.9 - 1.0e-10, 1.4, 3.0], [-1.0e-20, 0.1, 0.1], [0.9, 1.1, 2.0]),
    ("Maximum density violation:", [1.2, 1.4, 3.0], [0.1, 0.1, 0.1], [0.9, 1.1, 2.0]),
    ("Maximum density violation (eps):", [1.1 + 1.0e-10, 1.4, 3.0], [1.0e-20, 0.1, 0.1], [0.9, 1.1, 2.0]),
    ("Minimum entropy violation:", [1.0, 1.4, 2.8], [0.1, 0.1, -0.1], [0.9, 1.1, 2.0]),
    ("Minimum entropy violation (eps):", [1.0, 1.4, 2.8], [0.1, 0.1, -1.0e-20], [0.9, 1.1, 1.82 + 1.0e-10]),
    ("Minimum density bound", [1.0, 1.4, 3.0], [-0.2, 0.1, 0.1], [0.9, 1.1, 2.0]),
    ("Minimum density bound (eps):", [0.9 + 1.0e-10, 1.4, 3.0], [-5.0e-10, 0.1, 0.1], [0.9, 1.1, 2.0]),
    ("Maximum density bound", [1.0, 1.4, 3.0], [0.2, 0.1, 0.1], [0.9, 1.1, 1.0]),
    ("Maximum density bound (eps):", [1.1 - 1.0e-10, 1.4, 3.0], [5.0e-10, 0.1, 0.1], [0.9, 1.1, 1.0]),
    ("Minimum entropy bound", [1.0, 1.4, 2.8], [0.1, 0.1, -0.3], [0.9, 1.1, 1.8]),
    ("Minimum entropy bound (eps):", [1.0, 1.4, 2.8], [0.1, 0.1, -4.0e-10], [0.9, 1.1, 1.82 - 1.0e-10]),
]


def limiter():
    text = read("tests/euler/limiter.output")
    cases = []
    for idx, (name, U, P, bounds) in enumerate(LIMITER_INPUTS):
        start = text.index("\n" + name + "\n") + 1
        end = len(text)
        if idx + 1 < len(LIMITER_INPUTS):
            end = text.index("\n" + LIMITER_INPUTS[idx + 1][0] + "\n", start)
        blk = text[start:end]
        case = {"name": name, "U": U, "P": P, "bounds": bounds, "rounds": []}
        case["state_printed"] = floats(re.search(r"State: (.*)", blk).group(1))
        case["specific_entropy"] = floats(re.search(r"Specific entropy: (.*)", blk).group(1))[0]
        case["bounds_printed"] = floats(re.search(r"Bounds: (.*)", blk).group(1))
        case["t_l_start"] = floats(re.search(r"t_l: \(start\) (.*)", blk).group(1))[0]
        case["t_r_start"] = floats(re.search(r"t_r: \(start\) (.*)", blk).group(1))[0]
        # per round: psi_l, psi_r [, dpsi_l, dpsi_r], t_l (n), t_r (n)
        for m in re.finditer(
            r"psi_l:\s+(\S+)\npsi_r:\s+(\S+)\n(?:dpsi_l:\s+(\S+)\ndpsi_r:\s+(\S+)\n)?"
            r"t_l: \(  (\d+)  \) (\S+)\nt_r: \(  \d+  \) (\S+)",
            blk,
        ):
            rd = {
                "psi_l": float(m.group(1)),
                "psi_r": float(m.group(2)),
                "n": int(m.group(5)),
                "t_l": float(m.group(6)),
                "t_r": float(m.group(7)),
            }
            if m.group(3):
                rd["dpsi_l"] = float(m.group(3))
                rd["dpsi_r"] = float(m.group(4))
            case["rounds"].append(rd)
        case["l"] = floats(re.search(r"\nl: (.*)", blk).group(1))[0]
        case["success"] = "Success!" in blk
        case["fixed_format"] = idx > 0
        cases.append(case)
    return {
        "source": "tests/euler/limiter.output (EXPENSIVE_BOUNDS_CHECK, dim=1)",
        "cases": cases,
    }


dump("euler_limiter.json", limiter())


# --------------------------------------------------------------------------
def hyperbolic_system():
    text = read("tests/euler/hyperbolic_system.output")
    dbl = text.split("double:")[1].split("float:")[0]
    out = {}
    for blk in re.split(r"(?=dim = \d)", dbl):
        blk = blk.strip()
        if not blk.startswith("dim ="):
            continue
        lines = blk.split("\n")
        dim = int(lines[0].split("=")[1])
        entry = {}
        for l in lines[1:]:
            key, val = l.split("=", 1)
            entry[key.strip().replace(" ", "_")] = floats(val)
        out[str(dim)] = entry
    return {
        "source": "tests/euler/hyperbolic_system.output (double, 10 digits)",
        "state_1d": "rho = gamma, u = 3, p = 1",
        "dims": out,
    }


dump("euler_hyperbolic_system.json", hyperbolic_system())


# --------------------------------------------------------------------------
def sparse_matrix():
    res = {"source": "tests/common/sparse_matrix_simd.output*"}
    for suffix, width in [("", 1), (".sse2", 2), (".avx2", 4), (".avx512", 8)]:
        text = read("tests/common/sparse_matrix_simd.output" + suffix)
        sec = re.split(
            r"Matrix entries row by row|Matrix entries by SIMD rows|"
            r"Matrix entries transposed row by row|Matrix entries transposed by SIMD row",
            text,
        )
        res["width_%d" % width] = {
            "row_by_row": [floats(l) for l in sec[1].strip().split("\n")],
            "by_simd_rows": floats(sec[2]),
            "transposed_row_by_row": [floats(l) for l in sec[3].strip().split("\n")],
            "transposed_by_simd_row": floats(sec[4]),
        }
    return res


dump("sparse_matrix_simd.json", sparse_matrix())


# --------------------------------------------------------------------------
def verification():
    res = {
        "source": "tests/euler/verification-isentropic_vortex-2d-*.{prm,output}",
        "setup": {
            "geometry": "rectangular domain [-5,5]^2, all dirichlet, mesh refinement L -> (2^L+1)^2 dofs",
            "initial": "isentropic vortex, direction (1,1), position (-1,-1), mach 1, beta 5, gamma 1.4",
            "time": "final time 2.0, cfl min = cfl max = 0.2, recovery none",
        },
        "runs": [],
    }
    d = os.path.join(REF, "tests/euler")
    for f in sorted(os.listdir(d)):
        m = re.match(r"verification-isentropic_vortex-2d-(erk33|ssprk33)-l(\d)(?:\.mpirun=(\d))?\.output", f)
        if not m:
            continue
        text = open(os.path.join(d, f)).read()
        res["runs"].append(
            {
                "file": f,
                "scheme": m.group(1),
                "refinement": int(m.group(2)),
                "mpirun": int(m.group(3) or 1),
                "dofs": int(re.search(r"#dofs = (\d+)", text).group(1)),
                "t": float(re.search(r"t     = (\S+)", text).group(1)),
                "Linf": float(re.search(r"Linf  = (\S+)", text).group(1)),
                "L1": float(re.search(r"L1    = (\S+)", text).group(1)),
                "L2": float(re.search(r"L2    = (\S+)", text).group(1)),
            }
        )
    return res


dump("euler_verification.json", verification())


# --------------------------------------------------------------------------
def sw_riemann():
    path = "tests/shallow_water/riemann_solver.output"
    if not os.path.exists(os.path.join(REF, path)):
        return None
    return {"source": path, "text": read(path)}


sw = sw_riemann()
if sw:
    dump("shallow_water_riemann_solver.json", sw)


# --------------------------------------------------------------------------
def prm_runs():
    """Parameter files of the reference's end-to-end verification tests together with the
    numbers of their golden outputs (input decks are test data, not source): what
    `python -m ryujin_b200 <file.prm>` must reproduce.  tests/test_prm_frontend.py"""
    runs = []
    for sub, stem in [
        ("euler", "verification-isentropic_vortex-2d-erk33-l5"),
        ("euler", "verification-isentropic_vortex-2d-erk33-l6"),
        ("euler", "verification-isentropic_vortex-2d-ssprk33-l5"),
        ("euler", "verification-isentropic_vortex-2d-ssprk33-l6"),
        ("shallow_water", "verification-smooth_vortex-erk33-l6"),
        ("shallow_water", "verification-ritter_dam_break-erk33-l7"),
        # appended (indices of the first six are referenced by tests): the euler_aeos module with the
        # polytropic gas equation of state
        ("euler_aeos", "verification-isentropic_vortex-pge-2d-erk33-l5"),
        ("euler_aeos", "verification-isentropic_vortex-pge-2d-erk33-l6"),
        ("euler_aeos", "verification-isentropic_vortex-pge-2d-ssprk33-l5"),
        ("euler_aeos", "verification-isentropic_vortex-pge-2d-ssprk33-l6"),
    ]:
        out = read("tests/%s/%s.output" % (sub, stem))
        vals = {}
        for key, tag in [("dofs", "#dofs"), ("t", "t "), ("Linf", "Linf"), ("L1", "L1"), ("L2", "L2")]:
            m = re.search(r"^%s\s*=\s*(\S+)" % re.escape(tag.strip()), out, re.M)
            vals[key] = float(m.group(1))
        runs.append({
            "name": "%s/%s" % (sub, stem),
            "prm": read("tests/%s/%s.prm" % (sub, stem)),
            "output": out,
            "expected": vals,
        })
    return {"source": "tests/{euler,shallow_water,euler_aeos}/verification-*.{prm,output}", "runs": runs}


dump("prm_runs.json", prm_runs())


# --------------------------------------------------------------------------
def mass_conservation():
    """tests/euler/check-mass-conservation_01.{prm,output}: uniform Mach 3 state in a closed slip
    box, ssprk 33 at cfl 0.9; the output is the space averaged primitive state (and its second
    moments) at the start of every cycle -- a per-cycle time series through reflecting shocks."""
    out = read("tests/euler/check-mass-conservation_01.output")
    rows = []
    for line in out.splitlines():
        if line.startswith("[") or line.startswith("#") or not line.strip():
            continue
        rows.append(floats(line))
    assert all(len(r) == 9 for r in rows)
    return {
        "source": "tests/euler/check-mass-conservation_01.{prm,output}",
        "prm": read("tests/euler/check-mass-conservation_01.prm"),
        "columns": ["t", "rho", "v_1", "v_2", "p", "rho^2", "v_1^2", "v_2^2", "p^2"],
        "rows": rows,
    }


dump("euler_mass_conservation.json", mass_conservation())
