"""Pin the CPU oracle against the reference's own golden outputs.

Fixtures: tests/golden/*.json, extracted from /root/reference/tests/**/*.output
by tests/golden/extract_reference_goldens.py.
"""
import math

import numpy as np
import pytest

from conftest import load_golden


def close(a, b, fixed=False, rel=2e-15):
    """Goldens carry 17 significant digits (scientific) or 16 decimals (fixed)."""
    if fixed:
        return abs(a - b) <= 0.51e-16 + rel * abs(b) * 8
    return abs(a - b) <= rel * max(abs(a), abs(b)) * 8 or a == b


@pytest.mark.parametrize("which", ["oracle", "oracle_port"])
@pytest.mark.parametrize("iters", [0, 2, 10])
def test_euler_riemann_solver(request, which, iters):
    """tests/euler/riemann_solver{,-iterated-2,-iterated-10}.output: all 10 two-state cases."""
    orc = request.getfixturevalue(which)
    g = load_golden("euler_riemann_solver.json")["newton_%d" % iters]
    assert len(g["cases"]) == 10
    for case in g["cases"]:
        fixed = case["fixed_format"]
        r = orc.euler_riemann(case["state_i"], case["state_j"], gamma=g["gamma"], riemann_newton_max_iterations=iters)
        for key in ("a_left", "a_right", "p_star_two_rarefaction", "p_star_failsafe", "p_star_tilde", "phi_p_star_tilde", "lambda_max"):
            # the libm-pow "port" may differ from VCL pow in the last ulp of pow results
            rel = 2e-15 if which == "oracle" else 4e-15
            if key == "phi_p_star_tilde":
                # phi is a difference of O(1) terms: absolute floor
                ok = abs(r[key] - case[key]) <= 1e-14 * max(1.0, abs(case[key]))
            else:
                ok = close(r[key], case[key], fixed, rel)
            assert ok, (iters, case["state_i"], key, r[key], case[key])


@pytest.mark.parametrize("which", ["oracle", "oracle_port"])
def test_euler_limiter(request, which):
    """tests/euler/limiter.output: 12 cases with per-round t_l/t_r/psi traces."""
    orc = request.getfixturevalue(which)
    g = load_golden("euler_limiter.json")
    assert len(g["cases"]) == 12
    for case in g["cases"]:
        fixed = case["fixed_format"]
        r = orc.euler_limiter(1, case["U"], case["P"], case["bounds"], expensive_bounds_check=1)
        assert r["success"] == case["success"], case["name"]
        assert close(r["l"], case["l"], fixed), (case["name"], r["l"], case["l"])
        assert close(r["specific_entropy"], case["specific_entropy"], fixed)
        assert close(r["t_l_start"], case["t_l_start"], fixed)
        assert close(r["t_r_start"], case["t_r_start"], fixed)
        assert len(r["rounds"]) == len(case["rounds"]), case["name"]
        for mine, ref in zip(r["rounds"], case["rounds"]):
            for key in ("t_l", "t_r", "psi_l", "psi_r"):
                assert abs(mine[key] - ref[key]) <= 1e-15 + 1e-14 * abs(ref[key]), (case["name"], key, mine[key], ref[key])


def test_euler_hyperbolic_system(oracle):
    """tests/euler/hyperbolic_system.output (double, dim 1-3, 10 digits)."""
    g = load_golden("euler_hyperbolic_system.json")
    gamma = 7.0 / 5.0
    for dim in (1, 2, 3):
        k = dim + 2
        rho, u, p = gamma, 3.0, 1.0
        U = np.zeros(k)
        U[0] = rho
        U[1] = rho * u
        U[dim + 1] = p / (gamma - 1.0) + 0.5 * rho * u * u
        q = oracle.euler_system_quantities(dim, U)
        ref = g["dims"][str(dim)]
        mapping = {
            "momentum": "momentum",
            "internal_energy": "internal_energy",
            "internal_energy_derivative": "internal_energy_derivative",
            "pressure": "pressure",
            "specific_entropy": "specific_entropy",
            "harten_entropy": "harten_entropy",
            "harten_entropy_derivative": "harten_entropy_derivative",
            "mathematical_entropy": "mathematical_entropy",
            "mathematical_entropy_derivative": "mathematical_entropy_derivative",
            "f": "f",
        }
        for gkey, okey in mapping.items():
            refv = np.array(ref[gkey])
            mine = np.asarray(q[okey]).ravel()
            assert mine.shape == refv.shape, (dim, gkey)
            assert np.allclose(mine, refv, rtol=1e-10, atol=1e-10), (dim, gkey, mine, refv)


def _tridiagonal_wrap_pattern():
    """The 14-row pattern of tests/common/sparse_matrix_simd.cc:9-22, as CSR with
    diagonal first then ascending columns."""
    rows = {0: [0, 1, 13], 12: [12, 11], 13: [13, 0]}
    for i in range(1, 12):
        rows[i] = [i, i - 1, i + 1]
    row_ptr = [0]
    cols = []
    for i in range(14):
        cols += rows[i]
        row_ptr.append(len(cols))
    return np.array(row_ptr, dtype=np.uint32), np.array(cols, dtype=np.uint32)


@pytest.mark.parametrize("width", [1, 2, 4, 8])
def test_sparse_matrix_simd_layout(oracle, width):
    """tests/common/sparse_matrix_simd.output{,.sse2,.avx2,.avx512}: entry order by SIMD rows
    and the transposed index table."""
    g = load_golden("sparse_matrix_simd.json")["width_%d" % width]
    row_ptr, cols = _tridiagonal_wrap_pattern()
    n_internal = (12 // width) * width
    row_starts, ci, it = oracle.sparsity_simd(row_ptr, cols, n_internal, width)

    def entry_index(row, pos):
        if row < n_internal:
            grp = row // width
            return int(row_starts[grp]) + pos * width + row % width
        return int(row_starts[row]) + pos

    # the test writes value 3*i+j at (i, j) for i < 12, then 36..39
    data = np.full(len(cols), -1.0)
    for i in range(12):
        for j in range(3):
            data[entry_index(i, j)] = 3 * i + j
    data[entry_index(12, 0)] = 36.0
    data[entry_index(12, 1)] = 37.0
    data[entry_index(13, 0)] = 38.0
    data[entry_index(13, 1)] = 39.0

    row_len = np.diff(row_ptr)
    by_row = [[data[entry_index(i, j)] for j in range(row_len[i])] for i in range(14)]
    assert by_row == g["row_by_row"]
    transposed = [[data[it[entry_index(i, j)]] for j in range(row_len[i])] for i in range(14)]
    assert transposed == g["transposed_row_by_row"]

    # "by SIMD rows": raw storage order of the SIMD part, then rows 12.. of the tail
    simd = []
    for i in range(0, n_internal, width):
        for j in range(3):
            simd += [data[entry_index(i + kk, j)] for kk in range(width)]
    tail = []
    for i in range(n_internal, 14):
        tail += [data[entry_index(i, 0)], data[entry_index(i, 1)]]
    assert simd + tail == g["by_simd_rows"]

    simd_t = []
    for i in range(0, n_internal, width):
        for j in range(3):
            simd_t += [data[it[entry_index(i + kk, j)]] for kk in range(width)]
    tail_t = []
    for i in range(n_internal, 14):
        tail_t += [data[it[entry_index(i, 0)]], data[it[entry_index(i, 1)]]]
    assert simd_t + tail_t == g["transposed_by_simd_row"]
    # column indices decode to the graph
    for i in range(14):
        for j in range(row_len[i]):
            assert ci[entry_index(i, j)] == cols[row_ptr[i] + j]
