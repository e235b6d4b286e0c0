#!/usr/bin/env python
"""Extracts the reference's own known-answer vectors for the `euler_aeos` equation module from
/root/reference/tests/euler_aeos/*.{cc,output} into tests/golden/aeos_*.json (inputs that the .output
files do not echo -- the update direction P and bounds of the limiter cases, the covolume of the
Riemann cases -- are read from the test drivers' argument lists).  Run in the authoring container:

    python tests/golden/extract_reference_goldens_aeos.py
"""
import json
import os
import re

REF = "/root/reference/tests/euler_aeos"
HERE = os.path.dirname(os.path.abspath(__file__))


def ev(expr):
    return float(eval(expr.replace("{", "(").replace("}", ")"), {"__builtins__": {}}))  # noqa: S307 -- arithmetic literals


def numbers(text):
    return [ev(x.strip()) for x in text.split(",") if x.strip()]


def riemann(cc, output, strict):
    src = open(os.path.join(REF, cc)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = src[src.index("std::cout << std::scientific"):]
    cases, b = [], 0.0
    for m in re.finditer(r"set_covolume\(([^)]*)\)|test\(\{([^}]*)\},\s*\{([^}]*)\}\)", body):
        if m.group(1) is not None:
            b = ev(m.group(1))
        else:
            cases.append(dict(state_i=numbers(m.group(2)), state_j=numbers(m.group(3)), covolume_b=b, strict=strict))
    out = open(os.path.join(REF, output)).read()
    blocks = [blk for blk in out.split("\n\n") if "lambda_max" in blk]
    assert len(blocks) == len(cases), (len(blocks), len(cases))
    for case, blk in zip(cases, blocks):
        lines = blk.strip().splitlines()
        echoed_i, echoed_j = [float(x) for x in lines[0].split()], [float(x) for x in lines[1].split()]
        assert echoed_i == case["state_i"] and echoed_j == case["state_j"], (echoed_i, case["state_i"])
        for key, pat in (("a_left", r"\ba_left: (\S+)"), ("a_right", r"\ba_right: (\S+)"), ("p_2", r"p\^\*_tilde  = (\S+)"),
                         ("lambda_max", r"-> lambda_max = (\S+)"), ("p_star_RS_1", r"RS p_1_tilde  = (\S+)"),
                         ("p_star_RS_2", r"RS p_2_tilde  = (\S+)"), ("p_star_SS_1", r"SS p_1_tilde  = (\S+)"),
                         ("p_star_SS_2", r"SS p_2_tilde  = (\S+)"), ("p_star_interpolated", r"IN p_\*_tilde  = (\S+)")):
            found = re.findall(pat, blk)
            if found:
                case[key] = float(found[0])
    return cases


def limiter(cc, output, eos):
    src = open(os.path.join(REF, cc)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    body = src[src.index("std::cout << std::scientific"):]
    cases, b = [], 0.0
    P = None
    for m in re.finditer(r"set_covolume\(([^)]*)\)|(?:auto )?P = state_type\{\{([^}]*)\}\}|test\(U, P, bounds\)", body):
        if m.group(1) is not None:
            b = ev(m.group(1))
        elif m.group(2) is not None:
            P = numbers(m.group(2))
        else:
            cases.append(dict(P=P, covolume_b=b, eos=eos))
    out = open(os.path.join(REF, output)).read()
    results = re.findall(r"State: (\S+) (\S+) (\S+)\nSpecific entropy: (\S+)\nBounds: (\S+) (\S+) (\S+)\n.*?\nl: (\S+)\n(Success|Failure)!", out, flags=re.S)
    assert len(results) == len(cases), (len(results), len(cases))
    for case, r in zip(cases, results):
        case["U"] = [float(x) for x in r[:3]]  # printed with 17 significant digits: exact
        case["bounds"] = [float(x) for x in r[4:7]] + [1.4]
        case["specific_entropy"] = float(r[3])
        case["l"] = float(r[7])
        case["success"] = r[8] == "Success"
    return cases


def system(output):
    """the `double:` sections (dim 1-3 each); the state is from_1d_state(gamma, 3, 1 / gamma / (gamma - 1)) with
    gamma = the printed density"""
    out = open(os.path.join(REF, output)).read()
    cases = []
    for section in re.split(r"\n(?=double:|float:)", out):
        if not section.startswith("double:"):
            continue
        for blk in re.split(r"(?=interpolatory covolume: )", section):
            if "dim =" not in blk:
                continue
            g = lambda name: [float(x) for x in re.search(name + r" = (.*)", blk).group(1).split()]  # noqa: E731
            cases.append(dict(covolume_b=float(re.search(r"interpolatory covolume: (\S+)", blk).group(1)), dim=int(g("dim")[0]),
                              density=g("density")[0], momentum=g("momentum"), total_energy=g("total_energy")[0],
                              internal_energy=g("internal_energy")[0], internal_energy_derivative=g("internal_energy_derivative"),
                              specific_entropy=g("specific_entropy")[0], harten_entropy=g("harten entropy")[0],
                              harten_entropy_derivative=g("harten_entropy_derivative"), surrogate_pressure=g("surrogate_pressure")[0],
                              surrogate_gamma=g("surrogate_gamma")[0], f=g("f")))
    return cases


def verification():
    runs = []
    for name in sorted(os.listdir(REF)):
        m = re.match(r"verification-isentropic_vortex-pge-2d-(erk33|ssprk33)-l(\d)\.output$", name)
        if not m:
            continue
        text = open(os.path.join(REF, name)).read()
        runs.append(dict(scheme=m.group(1), refinement=int(m.group(2)), dofs=int(re.search(r"#dofs = (\d+)", text).group(1)),
                         t=float(re.search(r"t     = (\S+)", text).group(1)), Linf=float(re.search(r"Linf  = (\S+)", text).group(1)),
                         L1=float(re.search(r"L1    = (\S+)", text).group(1)), L2=float(re.search(r"L2    = (\S+)", text).group(1))))
    return runs


if __name__ == "__main__":
    data = dict(
        source="/root/reference/tests/euler_aeos (ryujin v2.1.1)",
        riemann=riemann("riemann_solver.cc", "riemann_solver.output", False) + riemann("riemann_solver-strict.cc", "riemann_solver-strict.output", True),
        limiter=limiter("limiter.cc", "limiter.output", "van der waals") + limiter("limiter-NASG.cc", "limiter-NASG.output", "noble abel stiffened gas"),
        system=system("hyperbolic_system.output"),
        verification=verification(),
    )
    with open(os.path.join(HERE, "aeos_reference.json"), "w") as f:
        json.dump(data, f, indent=1)
    print({k: (len(v) if isinstance(v, list) else v) for k, v in data.items()})
