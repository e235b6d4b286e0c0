"""ORACLE -- TEST INFRASTRUCTURE ONLY.

ctypes binding of the CPU oracle (oracle/liboracle.so, or the VCL-pow variant
oracle/_ref/liboracle_vcl.so when it was built in the authoring container).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs may import this module; the product (ryujin_b200/) never
does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

EULER_PARAM_FIELDS = (
    "gamma",
    "reference_density",
    "vacuum_state_relaxation_small",
    "vacuum_state_relaxation_large",
    "riemann_newton_tolerance",
    "riemann_newton_max_iterations",
    "indicator_evc_factor",
    "limiter_iterations",
    "limiter_newton_tolerance",
    "limiter_newton_max_iterations",
    "limiter_relaxation_factor",
    "expensive_bounds_check",
)

EULER_PARAM_DEFAULTS = dict(
    gamma=7.0 / 5.0,
    reference_density=1.0,
    vacuum_state_relaxation_small=1.0e2,
    vacuum_state_relaxation_large=1.0e4,
    riemann_newton_tolerance=1.0e-10,
    riemann_newton_max_iterations=0,
    indicator_evc_factor=1.0,
    limiter_iterations=2,
    limiter_newton_tolerance=1.0e-10,
    limiter_newton_max_iterations=2,
    limiter_relaxation_factor=1.0,
    expensive_bounds_check=0,
)

AEOS_POLYTROPIC_GAS, AEOS_NOBLE_ABEL_STIFFENED_GAS, AEOS_VAN_DER_WAALS = 0, 1, 2
AEOS_PARAM_FIELDS = (
    "eos", "gamma", "covolume_b", "reference_pressure", "reference_specific_internal_energy", "vdw_a",
    "compute_strict_bounds", "reference_density", "vacuum_state_relaxation_small", "vacuum_state_relaxation_large",
    "indicator_evc_factor", "limiter_iterations", "limiter_newton_tolerance", "limiter_newton_max_iterations",
    "limiter_relaxation_factor", "expensive_bounds_check",
)
AEOS_PARAM_DEFAULTS = dict(
    eos=AEOS_POLYTROPIC_GAS, gamma=7.0 / 5.0, covolume_b=0.0, reference_pressure=0.0, reference_specific_internal_energy=0.0,
    vdw_a=0.0, compute_strict_bounds=1, reference_density=1.0, vacuum_state_relaxation_small=1.0e2,
    vacuum_state_relaxation_large=1.0e4, indicator_evc_factor=1.0, limiter_iterations=2, limiter_newton_tolerance=1.0e-10,
    limiter_newton_max_iterations=2, limiter_relaxation_factor=1.0, expensive_bounds_check=0,
)

SW_PARAM_FIELDS = (
    "gravity",
    "manning_friction_coefficient",
    "reference_water_depth",
    "dry_state_relaxation_factor",
    "dry_state_relaxation_small",
    "dry_state_relaxation_large",
    "indicator_evc_factor",
    "limiter_iterations",
    "limiter_newton_tolerance",
    "limiter_newton_max_iterations",
    "limiter_relaxation_factor",
    "limit_on_kinetic_energy",
    "limit_on_square_velocity",
    "expensive_bounds_check",
)

SW_PARAM_DEFAULTS = dict(
    gravity=9.81,
    manning_friction_coefficient=0.0,
    reference_water_depth=1.0,
    dry_state_relaxation_factor=2.0e-1,
    dry_state_relaxation_small=1.0e2,
    dry_state_relaxation_large=1.0e4,
    indicator_evc_factor=1.0,
    limiter_iterations=2,
    limiter_newton_tolerance=1.0e-10,
    limiter_newton_max_iterations=2,
    limiter_relaxation_factor=1.0,
    limit_on_kinetic_energy=0,
    limit_on_square_velocity=1,
    expensive_bounds_check=0,
)

# shallow water initial states (same numbers as enum rj_initial_state_kind)
SW_CIRCULAR_DAM_BREAK, SW_UNIFORM, SW_RITTER_DAM_BREAK, SW_SMOOTH_VORTEX, SW_SLOPING_FRICTION = 16, 17, 18, 19, 20

# Boundary ids, source/discretization.h:48-123
DO_NOTHING, PERIODIC, SLIP, NO_SLIP, DIRICHLET, DYNAMIC, DIRICHLET_MOMENTUM = range(7)

# initial state kinds
UNIFORM, CONTRAST, ISENTROPIC_VORTEX, LEBLANC, RAREFACTION = 0, 1, 2, 3, 4

SCHEMES = {"ssprk 22": 0, "ssprk 33": 1, "erk 11": 2, "erk 22": 3, "erk 33": 4, "erk 43": 5, "erk 54": 6}

SCRATCH = {"alpha": 0, "dij": 1, "lij": 2, "lij_next": 3, "pij": 4, "r": 5, "bounds": 6}


def build(ref=True):
    """(Re)build the oracle libraries with oracle/Makefile."""
    subprocess.check_call(["make", "-s", "-C", _HERE, "all"])
    if ref and os.path.isdir("/root/reference/simd-math"):
        subprocess.check_call(["make", "-s", "-C", _HERE, "ref"])


def euler_params(**kw):
    p = dict(EULER_PARAM_DEFAULTS)
    for k, v in kw.items():
        if k not in p:
            raise KeyError(k)
        p[k] = v
    return np.array([float(p[f]) for f in EULER_PARAM_FIELDS], dtype=np.float64)


def aeos_params(**kw):
    p = dict(AEOS_PARAM_DEFAULTS)
    for k, v in kw.items():
        if k not in p:
            raise KeyError(k)
        p[k] = v
    return np.array([float(p[f]) for f in AEOS_PARAM_FIELDS], dtype=np.float64)


def sw_params(**kw):
    p = dict(SW_PARAM_DEFAULTS)
    for k, v in kw.items():
        if k not in p:
            raise KeyError(k)
        p[k] = v
    return np.array([float(p[f]) for f in SW_PARAM_FIELDS], dtype=np.float64)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint))


def _ip(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


class OracleError(RuntimeError):
    pass


class Restart(Exception):
    """Mirror of ryujin::Restart (source/hyperbolic_module.h:59-61)."""


class Oracle:
    def __init__(self, variant="auto"):
        port = os.path.join(_HERE, "liboracle.so")
        vcl = os.path.join(_HERE, "_ref", "liboracle_vcl.so")
        if variant == "auto":
            path = vcl if os.path.exists(vcl) else port
        elif variant == "vcl":
            path = vcl
        elif variant == "vcl_fma":
            path = os.path.join(_HERE, "_ref", "liboracle_vcl_fma.so")
        elif variant == "native":  # timing build (-O3, contraction on): bench.py's CPU arm only
            path = os.path.join(_HERE, "_ref", "liboracle_native.so")
        elif variant == "port":
            path = port
        elif variant == "fma":
            path = os.path.join(_HERE, "liboracle_fma.so")
        elif variant == "rjpow":
            path = os.path.join(_HERE, "liboracle_rjpow.so")
        else:
            path = variant  # explicit path of another build of the same sources (experiments)
        if not os.path.exists(path):
            build(ref=(variant != "port"))
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.orc_last_error.restype = C.c_char_p
        L.orc_pow.restype = C.c_double
        L.orc_pow.argtypes = [C.c_double, C.c_double]
        L.orc_offline_assemble.restype = C.c_void_p
        L.orc_offline_from_arrays.restype = C.c_void_p
        L.orc_sim_create_euler.restype = C.c_void_p
        L.orc_sim_create_sw.restype = C.c_void_p
        L.orc_sim_create_aeos.restype = C.c_void_p
        L.orc_sim_last_tau_max.restype = C.c_double
        L.orc_sim_last_tau_max.argtypes = [C.c_void_p]
        self.kind = "reference-pow(VCL)" if L.orc_uses_vcl() else "port"

    # -- helpers -----------------------------------------------------------
    def _check(self, rc):
        if rc == 1:
            raise Restart()
        if rc != 0:
            raise OracleError(self.lib.orc_last_error().decode())

    def max_threads(self):
        return int(self.lib.orc_max_threads())

    def set_threads(self, n):
        self.lib.orc_set_threads(int(n))

    def flush_denormals_to_zero(self, on=True):
        """FTZ/DAZ on all OpenMP threads (source/main.cc:26-36); timing runs only"""
        return int(self.lib.orc_flush_denormals_to_zero(int(bool(on))))

    def pow(self, x, b):
        return float(self.lib.orc_pow(x, b))

    # -- unit level ----------------------------------------------------------
    def euler_riemann(self, state_i, state_j, **prm):
        out = np.zeros(7)
        si = np.asarray(state_i, dtype=np.float64)
        sj = np.asarray(state_j, dtype=np.float64)
        self._check(self.lib.orc_euler_riemann(_dp(euler_params(**prm)), _dp(si), _dp(sj), _dp(out)))
        keys = ("p_star_two_rarefaction", "p_star_failsafe", "p_star_tilde", "phi_p_star_tilde", "lambda_max", "a_left", "a_right")
        return dict(zip(keys, out.tolist()))

    def euler_riemann_states(self, dim, U_i, U_j, n_ij, **prm):
        lam = C.c_double()
        a = np.ascontiguousarray(U_i, dtype=np.float64)
        b = np.ascontiguousarray(U_j, dtype=np.float64)
        n = np.ascontiguousarray(n_ij, dtype=np.float64)
        self._check(self.lib.orc_euler_riemann_states(int(dim), _dp(euler_params(**prm)), _dp(a), _dp(b), _dp(n), C.byref(lam)))
        return lam.value

    # -- euler aeos unit level --------------------------------------------------
    def aeos_riemann(self, state_i, state_j, **prm):
        """tests/euler_aeos/riemann_solver*.cc: 1D primitive states (rho, u, p, gamma)"""
        out = np.zeros(8)
        si, sj = np.asarray(state_i, dtype=np.float64), np.asarray(state_j, dtype=np.float64)
        self._check(self.lib.orc_aeos_riemann(_dp(aeos_params(**prm)), _dp(si), _dp(sj), _dp(out)))
        keys = ("lambda_max", "p_2", "p_star_RS", "p_star_SS", "p_star_interpolated", "p_star_failsafe", "a_left", "a_right")
        return dict(zip(keys, out.tolist()))

    def aeos_limiter(self, U, P, bounds, **prm):
        out = np.zeros(3)
        u, p, b = (np.ascontiguousarray(x, dtype=np.float64) for x in (U, P, bounds))
        self._check(self.lib.orc_aeos_limiter(_dp(aeos_params(**prm)), _dp(u), _dp(p), _dp(b), _dp(out)))
        return dict(l=out[0], success=bool(out[1]), specific_entropy=out[2])

    def aeos_system_quantities(self, dim, U, gamma, **prm):
        k = dim + 2
        out = np.zeros(1 + k + 2 + k + 2 + k * dim)
        u = np.ascontiguousarray(U, dtype=np.float64)
        self._check(self.lib.orc_aeos_system_quantities(int(dim), _dp(aeos_params(**prm)), _dp(u), C.c_double(gamma), _dp(out)))
        o = 0
        res = {}
        for name, n in (("internal_energy", 1), ("internal_energy_derivative", k), ("specific_entropy", 1), ("harten_entropy", 1),
                        ("harten_entropy_derivative", k), ("surrogate_pressure", 1), ("surrogate_gamma", 1), ("f", k * dim)):
            res[name] = out[o] if n == 1 else out[o:o + n].copy()
            o += n
        return res

    def sw_riemann(self, state_i, state_j, **prm):
        """tests/shallow_water/riemann_solver.cc: conserved 1D states (h, m)"""
        out = np.zeros(6)
        si = np.asarray(state_i, dtype=np.float64)
        sj = np.asarray(state_j, dtype=np.float64)
        self._check(self.lib.orc_sw_riemann(_dp(sw_params(**prm)), _dp(si), _dp(sj), _dp(out)))
        return dict(zip(("h_i", "u_i", "h_j", "u_j", "h_star", "lambda_max"), out.tolist()))

    def euler_limiter(self, dim, U, P, bounds, **prm):
        out = np.zeros(5 + 4 * 8 + 1)
        u = np.ascontiguousarray(U, dtype=np.float64)
        p = np.ascontiguousarray(P, dtype=np.float64)
        b = np.ascontiguousarray(bounds, dtype=np.float64)
        self._check(self.lib.orc_euler_limiter(int(dim), _dp(euler_params(**prm)), _dp(u), _dp(p), _dp(b), _dp(out)))
        n = int(out[4])
        rounds = [dict(t_l=out[5 + 4 * r], t_r=out[6 + 4 * r], psi_l=out[7 + 4 * r], psi_r=out[8 + 4 * r]) for r in range(n)]
        return dict(l=out[0], success=bool(out[1]), t_l_start=out[2], t_r_start=out[3], rounds=rounds, specific_entropy=out[5 + 32])

    def euler_system_quantities(self, dim, U, **prm):
        k = dim + 2
        out = np.zeros(dim + 1 + k + 3 + k + 1 + k + k * dim)
        u = np.ascontiguousarray(U, dtype=np.float64)
        self._check(self.lib.orc_euler_system_quantities(int(dim), _dp(euler_params(**prm)), _dp(u), _dp(out)))
        o = 0
        res = {}
        for name, size in [
            ("momentum", dim),
            ("internal_energy", 1),
            ("internal_energy_derivative", k),
            ("pressure", 1),
            ("specific_entropy", 1),
            ("harten_entropy", 1),
            ("harten_entropy_derivative", k),
            ("mathematical_entropy", 1),
            ("mathematical_entropy_derivative", k),
            ("f", k * dim),
        ]:
            res[name] = out[o : o + size].copy()
            o += size
        return res

    def sparsity_simd(self, row_ptr, cols, n_internal, simd_length):
        rp = np.ascontiguousarray(row_ptr, dtype=np.uint32)
        cl = np.ascontiguousarray(cols, dtype=np.uint32)
        n = len(rp) - 1
        row_starts = np.zeros(n + 1, dtype=np.uint64)
        ci = np.zeros(len(cl), dtype=np.uint32)
        it = np.zeros(len(cl), dtype=np.uint32)
        self._check(
            self.lib.orc_sparsity_simd(
                C.c_uint(n), _up(rp), _up(cl), C.c_uint(n_internal), C.c_uint(simd_length),
                row_starts.ctypes.data_as(C.POINTER(C.c_ulonglong)), _up(ci), _up(it),
            )
        )
        return row_starts, ci, it

    # -- offline data -----------------------------------------------------------
    def assemble(self, n_cells, origin, h, box_ids, mask=None, masked_id=SLIP):
        dim = len(n_cells)
        nc = np.ascontiguousarray(n_cells, dtype=np.uint32)
        og = np.ascontiguousarray(origin, dtype=np.float64)
        hh = np.ascontiguousarray(h, dtype=np.float64)
        ids = np.ascontiguousarray(box_ids, dtype=np.int32)
        assert len(ids) == 2 * dim
        mk = None
        if mask is not None:
            mk = np.ascontiguousarray(mask, dtype=np.uint8)
        handle = self.lib.orc_offline_assemble(
            dim, _up(nc), _dp(og), _dp(hh),
            mk.ctypes.data_as(C.POINTER(C.c_ubyte)) if mk is not None else None,
            _ip(ids), int(masked_id),
        )
        if not handle:
            raise OracleError(self.lib.orc_last_error().decode())
        return Offline(self, handle, dim)

    def assemble_quads(self, vertices, cells, boundary_faces, boundary_flip, boundary_ids):
        """mapped Q1 quadrilateral mesh (orc_mesh_quads.h): vertices (n, 2); cells (m, 4) vertex indices in
        deal.II's lexicographic local order; boundary faces (f, 2) in deal.II's visiting order, outward
        normal of a -> b = (dy, -dx) / |ab|, reversed where boundary_flip is set"""
        xy = np.ascontiguousarray(vertices, dtype=np.float64)
        cl = np.ascontiguousarray(cells, dtype=np.uint32)
        bf = np.ascontiguousarray(boundary_faces, dtype=np.uint32)
        bi = np.ascontiguousarray(boundary_ids, dtype=np.int32)
        fl = np.ascontiguousarray(boundary_flip, dtype=np.uint8)
        self.lib.orc_offline_assemble_quads.restype = C.c_void_p
        handle = self.lib.orc_offline_assemble_quads(C.c_uint(len(xy)), _dp(xy), C.c_uint(len(cl)), _up(cl), C.c_uint(len(bf)),
                                                     _up(bf), fl.ctypes.data_as(C.POINTER(C.c_ubyte)), _ip(bi))
        if not handle:
            raise OracleError(self.lib.orc_last_error().decode())
        return Offline(self, handle, 2)

    def offline_from_arrays(self, dim, row_ptr, cols, mij, cij, mi, measure, b_index, b_normal, b_id, b_position, pairs):
        rp = np.ascontiguousarray(row_ptr, dtype=np.uint32)
        cl = np.ascontiguousarray(cols, dtype=np.uint32)
        mij = np.ascontiguousarray(mij, dtype=np.float64)
        cij = np.ascontiguousarray(cij, dtype=np.float64)
        mi = np.ascontiguousarray(mi, dtype=np.float64)
        bi = np.ascontiguousarray(b_index, dtype=np.uint32)
        bn = np.ascontiguousarray(b_normal, dtype=np.float64)
        bid = np.ascontiguousarray(b_id, dtype=np.int32)
        bp = np.ascontiguousarray(b_position, dtype=np.float64)
        pr = np.ascontiguousarray(pairs, dtype=np.uint32)
        handle = self.lib.orc_offline_from_arrays(
            int(dim), C.c_uint(len(rp) - 1), _up(rp), _up(cl), _dp(mij), _dp(cij), _dp(mi), C.c_double(measure),
            C.c_uint(len(bi)), _up(bi), _dp(bn), _ip(bid), _dp(bp), C.c_uint(len(pr) // 3), _up(pr),
        )
        if not handle:
            raise OracleError(self.lib.orc_last_error().decode())
        return Offline(self, handle, dim)


class Offline:
    def __init__(self, oracle, handle, dim):
        self.o = oracle
        self.h = C.c_void_p(handle)
        self.dim = dim
        sz = np.zeros(4, dtype=np.uint64)
        oracle.lib.orc_offline_sizes(self.h, sz.ctypes.data_as(C.POINTER(C.c_ulonglong)))
        self.n, self.nnz, self.n_boundary, self.n_pairs = (int(x) for x in sz)
        self._arrays = None

    def __del__(self):
        try:
            self.o.lib.orc_offline_destroy(self.h)
        except Exception:
            pass

    def postprocess(self, U, mode, first, beta=10.0):
        """Postprocessor::compute for one quantity -> (values on the exponential scale, (q_max, q_min))"""
        U = np.ascontiguousarray(U, dtype=np.float64)
        out = np.zeros(self.n)
        bounds = np.zeros(2)
        self.o._check(self.o.lib.orc_postprocess(self.h, int(U.shape[1]), _dp(U), int(mode), int(first), C.c_double(beta),
                                                 _dp(out), _dp(bounds)))
        return out, (bounds[0], bounds[1])

    def set_partition(self, partition):
        """owner rank of every row (N-rank emulation, orc_module.h OfflineData::partition); None
        restores the single-rank semantics"""
        if partition is None:
            self.o.lib.orc_offline_set_partition(self.h, None)
            return
        p = np.ascontiguousarray(partition, dtype=np.int32)
        assert len(p) == self.n
        self.o.lib.orc_offline_set_partition(self.h, _ip(p))

    def arrays(self):
        if self._arrays is not None:
            return self._arrays
        d = self.dim
        a = dict(
            row_ptr=np.zeros(self.n + 1, dtype=np.uint32),
            cols=np.zeros(self.nnz, dtype=np.uint32),
            transposed=np.zeros(self.nnz, dtype=np.uint32),
            mij=np.zeros(self.nnz),
            cij=np.zeros((self.nnz, d)),
            mi=np.zeros(self.n),
            b_index=np.zeros(self.n_boundary, dtype=np.uint32),
            b_normal=np.zeros((self.n_boundary, d)),
            b_masses=np.zeros((self.n_boundary, 2)),
            b_id=np.zeros(self.n_boundary, dtype=np.int32),
            b_position=np.zeros((self.n_boundary, d)),
            pairs=np.zeros((self.n_pairs, 3), dtype=np.uint32),
            positions=np.zeros((self.n, d)),
        )
        measure = C.c_double()
        self.o.lib.orc_offline_get(
            self.h, _up(a["row_ptr"]), _up(a["cols"]), _up(a["transposed"]), _dp(a["mij"]), _dp(a["cij"]), _dp(a["mi"]),
            C.byref(measure), _up(a["b_index"]), _dp(a["b_normal"]), _dp(a["b_masses"]), _ip(a["b_id"]),
            _dp(a["b_position"]), _up(a["pairs"]), _dp(a["positions"]),
        )
        a["measure"] = measure.value
        self._arrays = a
        return a

    def euler_sim(self, kind=UNIFORM, direction=(1.0, 0.0, 0.0), position=(1.0, 0.0, 0.0), primitive=(1.4, 3.0, 1.0),
                  primitive_right=(1.4, 0.0, 1.0), mach=2.0, beta=5.0, **prm):
        iv = np.zeros(15)
        iv[0] = kind
        dd = list(direction) + [0.0] * (3 - len(direction))
        pp = list(position) + [0.0] * (3 - len(position))
        iv[1:4] = dd
        iv[4:7] = pp
        iv[7:10] = primitive
        iv[10:13] = primitive_right
        iv[13] = mach
        iv[14] = beta
        handle = self.o.lib.orc_sim_create_euler(self.h, _dp(euler_params(**prm)), _dp(iv))
        if not handle:
            raise OracleError(self.o.lib.orc_last_error().decode())
        return Sim(self, handle)


def _aeos_sim(offline, kind=UNIFORM, direction=(1.0, 0.0, 0.0), position=(1.0, 0.0, 0.0), primitive=(1.4, 3.0, 1.0),
              primitive_right=(1.4, 0.0, 1.0), mach=2.0, beta=5.0, **prm):
    """euler aeos simulation on this offline data (orc_euler_aeos.h)"""
    iv = np.zeros(15)
    iv[0] = kind
    iv[1:4] = list(direction) + [0.0] * (3 - len(direction))
    iv[4:7] = list(position) + [0.0] * (3 - len(position))
    iv[7:10] = primitive
    iv[10:13] = primitive_right
    iv[13] = mach
    iv[14] = beta
    handle = offline.o.lib.orc_sim_create_aeos(offline.h, _dp(aeos_params(**prm)), _dp(iv))
    if not handle:
        raise OracleError(offline.o.lib.orc_last_error().decode())
    return Sim(offline, handle)


Offline.aeos_sim = _aeos_sim


def _sw_sim(offline, kind=SW_UNIFORM, direction=(1.0, 0.0, 0.0), position=(1.0, 0.0, 0.0), p=(1.0, 1.0, 0.0, 0.0), **prm):
    iv = np.zeros(11)
    iv[0] = kind
    iv[1:4] = list(direction) + [0.0] * (3 - len(direction))
    iv[4:7] = list(position) + [0.0] * (3 - len(position))
    iv[7:11] = list(p) + [0.0] * (4 - len(p))
    handle = offline.o.lib.orc_sim_create_sw(offline.h, _dp(sw_params(**prm)), _dp(iv))
    if not handle:
        raise OracleError(offline.o.lib.orc_last_error().decode())
    return Sim(offline, handle)


Offline.sw_sim = _sw_sim


class Sim:
    """Module + TimeIntegrator + six state-vector slots (slot 0 = the state)."""

    def set_error_components(self, components):
        """restrict errors() to the given component indices (A - TimeLoop / error quantities)"""
        mask = 0
        for c in components:
            mask |= 1 << int(c)
        self.o.lib.orc_sim_set_error_components(self.h, C.c_uint(mask))

    def set_initial_precomputed(self, values):
        """bathymetry of the shallow water equations, one value per node"""
        z = np.ascontiguousarray(values, dtype=np.float64)
        assert z.size == self.n
        self.o._check(self.o.lib.orc_sim_set_initial_precomputed(self.h, _dp(z)))

    def __init__(self, offline, handle):
        self.off = offline
        self.o = offline.o
        self.h = C.c_void_p(handle)
        self.k = int(self.o.lib.orc_sim_k(self.h))
        self.n_bounds = int(self.o.lib.orc_sim_n_bounds(self.h))
        self.n_precomputed = int(self.o.lib.orc_sim_n_precomputed(self.h))
        self.n = offline.n
        self.nnz = offline.nnz

    def __del__(self):
        try:
            self.o.lib.orc_sim_destroy(self.h)
        except Exception:
            pass

    def interpolate(self, t=0.0, slot=0):
        self.o._check(self.o.lib.orc_sim_interpolate(self.h, C.c_double(t), int(slot)))

    def set_U(self, U, slot=0):
        u = np.ascontiguousarray(U, dtype=np.float64)
        assert u.size == self.n * self.k
        self.o._check(self.o.lib.orc_sim_set_U(self.h, int(slot), _dp(u)))

    def get_U(self, slot=0):
        u = np.zeros((self.n, self.k))
        self.o._check(self.o.lib.orc_sim_get_U(self.h, int(slot), _dp(u)))
        return u

    def get_prec(self, slot=0):
        p = np.zeros((self.n, self.n_precomputed))
        self.o._check(self.o.lib.orc_sim_get_prec(self.h, int(slot), _dp(p)))
        return p

    def prepare_state_vector(self, slot=0, t=0.0):
        self.o._check(self.o.lib.orc_sim_prepare_state(self.h, int(slot), C.c_double(t)))

    def step(self, old_slot, stage_slots, weights, new_slot, tau=0.0, tau_max=np.finfo(np.float64).max):
        st = np.ascontiguousarray(stage_slots, dtype=np.int32)
        w = np.ascontiguousarray(weights, dtype=np.float64)
        tau_out = C.c_double()
        self.o._check(
            self.o.lib.orc_sim_hyp_step(
                self.h, int(old_slot), len(st), _ip(st), _dp(w), int(new_slot), C.c_double(tau), C.c_double(tau_max), C.byref(tau_out)
            )
        )
        return tau_out.value

    def configure_time_integrator(self, scheme="erk 33", cfl_min=0.45, cfl_max=0.90, recovery="bang bang control"):
        rec = {"none": 0, "bang bang control": 1}[recovery]
        self.o._check(self.o.lib.orc_sim_configure_ti(self.h, SCHEMES[scheme], C.c_double(cfl_min), C.c_double(cfl_max), rec))

    def time_step(self, t, t_final=np.finfo(np.float64).max):
        tau = C.c_double()
        self.o._check(self.o.lib.orc_sim_ti_step(self.h, C.c_double(t), C.c_double(t_final), C.byref(tau)))
        return tau.value

    def set_cfl(self, cfl):
        self.o._check(self.o.lib.orc_sim_set_cfl(self.h, C.c_double(cfl)))

    def set_raise_on_violation(self, flag):
        self.o._check(self.o.lib.orc_sim_set_strategy(self.h, int(bool(flag))))

    def scratch(self, name):
        sizes = {"alpha": self.n, "dij": self.nnz, "lij": self.nnz, "lij_next": self.nnz,
                 "pij": self.nnz * self.k, "r": self.n * self.k, "bounds": self.n * self.n_bounds}
        out = np.zeros(sizes[name])
        self.o._check(self.o.lib.orc_sim_get_scratch(self.h, SCRATCH[name], _dp(out)))
        return out

    def counters(self):
        r, w = C.c_uint(), C.c_uint()
        self.o._check(self.o.lib.orc_sim_counters(self.h, C.byref(r), C.byref(w)))
        return r.value, w.value

    def last_tau_max(self):
        return float(self.o.lib.orc_sim_last_tau_max(self.h))

    def errors(self, t):
        out = np.zeros(3)
        self.o._check(self.o.lib.orc_sim_errors(self.h, C.c_double(t), _dp(out)))
        return dict(Linf=out[0], L1=out[1], L2=out[2])
