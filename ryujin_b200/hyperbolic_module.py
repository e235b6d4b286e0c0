"""HyperbolicModule / StateVector / TimeIntegrator: host-side mirror of the
reference interface for the explicit hyperbolic update
(source/hyperbolic_module.h:110-278, source/time_integrator.h) on top of the
C ABI in include/ryujin_b200.h.  All compute happens in libryujin_b200.so.
"""
import ctypes as C
import enum

import numpy as np

from . import _lib
from ._lib import Restart, check, lib  # noqa: F401  (Restart re-exported)


class Equation(enum.IntEnum):
    euler = 0
    shallow_water = 1
    euler_aeos = 2


class EquationOfState(enum.IntEnum):
    """EquationOfStateLibrary of source/euler_aeos (analytic equations of state)"""

    polytropic_gas = 0
    noble_abel_stiffened_gas = 1
    van_der_waals = 2


class IDViolationStrategy(enum.IntEnum):
    """source/hyperbolic_module.h:37-52"""

    warn = 0
    raise_exception = 1


class TimeSteppingScheme(enum.IntEnum):
    ssprk_22 = 0
    ssprk_33 = 1
    erk_11 = 2
    erk_22 = 3
    erk_33 = 4
    erk_43 = 5
    erk_54 = 6


class CFLRecoveryStrategy(enum.IntEnum):
    none = 0
    bang_bang_control = 1


class Stepping(enum.IntEnum):
    """how TimeIntegrator drives a cycle (bit-identical results): host round trip per substep,
    device-resident cycle, or CUDA graph replay of the device-resident cycle"""

    host = 0
    device = 1
    graph = 2


class InitialState(enum.IntEnum):
    uniform = 0
    contrast = 1
    isentropic_vortex = 2
    sw_circular_dam_break = 16
    sw_uniform = 17
    sw_ritter_dam_break = 18
    sw_smooth_vortex = 19


class HyperbolicSystem:
    """Runtime parameters of the equation module (B - Equation;
    source/euler/hyperbolic_system.h:668-684, shallow_water/hyperbolic_system.h:643-674,
    euler_aeos/hyperbolic_system.h:456-490 plus the parameters of the selected equation of state:
    equation_of_state, gamma, covolume_b, reference_pressure, reference_specific_internal_energy,
    vdw_a, compute_strict_bounds)."""

    def __init__(self, equation=Equation.euler, **kw):
        self.desc = _lib.EquationDesc()
        lib.rj_equation_desc_default(C.byref(self.desc), int(equation))
        for key, val in kw.items():
            if not hasattr(self.desc, key):
                raise KeyError(key)
            setattr(self.desc, key, val)

    @property
    def equation(self):
        return Equation(self.desc.equation)


class ModuleParameters:
    """F - HyperbolicModule/{indicator, limiter, riemann solver}"""

    def __init__(self, **kw):
        self.desc = _lib.ModuleParams()
        lib.rj_module_params_default(C.byref(self.desc))
        for key, val in kw.items():
            if not hasattr(self.desc, key):
                raise KeyError(key)
            setattr(self.desc, key, val)


class InitialValues:
    """E - InitialValues (source/initial_values.template.h:20-62)"""

    def __init__(self, configuration=InitialState.uniform, direction=None, position=None, primitive=None,
                 primitive_right=None, mach_number=None, beta=None, sw_params=None):
        self.desc = _lib.InitialValuesDesc()
        lib.rj_initial_values_desc_default(C.byref(self.desc), int(configuration))

        def fill(field, values):
            if values is not None:
                for q, v in enumerate(values):
                    field[q] = float(v)

        fill(self.desc.direction, direction)
        if direction is not None:
            for q in range(len(direction), 3):
                self.desc.direction[q] = 0.0
        fill(self.desc.position, position)
        if position is not None:
            for q in range(len(position), 3):
                self.desc.position[q] = 0.0
        fill(self.desc.primitive, primitive)
        fill(self.desc.primitive_right, primitive_right)
        fill(self.desc.sw_params, sw_params)
        if mach_number is not None:
            self.desc.mach_number = float(mach_number)
        if beta is not None:
            self.desc.beta = float(beta)


class Communicator:
    """One process per GPU; replaces the MPI_Comm of the reference's module."""

    def __init__(self, unique_id, rank, n_ranks, device):
        buf = (C.c_uint8 * _lib.RJ_UNIQUE_ID_BYTES).from_buffer_copy(bytes(unique_id))
        h = C.c_void_p()
        check(lib.rj_comm_create(C.byref(h), buf, rank, n_ranks, device))
        self._h = h
        self.rank, self.n_ranks, self.device = rank, n_ranks, device

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * _lib.RJ_UNIQUE_ID_BYTES)()
        check(lib.rj_comm_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_torch_distributed(cls, device):
        """Broadcast the id created on rank 0 through the default process group."""
        import torch
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        backend = dist.get_backend()
        dev = torch.device("cuda", device) if backend == "nccl" else torch.device("cpu")
        t = torch.zeros(_lib.RJ_UNIQUE_ID_BYTES, dtype=torch.uint8, device=dev)
        if rank == 0:
            t = torch.tensor(list(cls.unique_id()), dtype=torch.uint8, device=dev)
        dist.broadcast(t, src=0)
        return cls(bytes(t.cpu().tolist()), rank, world, device)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.rj_comm_destroy(h)
            self._h = None


class StateVector:
    """(U, precomputed) on the device, component-major; host transfers use the
    reference's AoS layout (U_i)_c at i*k + c (source/multicomponent_vector.h:266)."""

    def __init__(self, module):
        self.module = module
        h = C.c_void_p()
        check(lib.rj_state_create(C.byref(h), module._h))
        self._h = h

    def upload(self, U):
        a = np.ascontiguousarray(U, dtype=np.float64)
        assert a.size == self.module.n_owned * self.module.k
        check(lib.rj_state_upload(self._h, a.ctypes.data))

    def upload_from_pointer(self, host_ptr):
        check(lib.rj_state_upload(self._h, host_ptr))

    def download(self, out=None):
        if out is None:
            out = np.zeros((self.module.n_owned, self.module.k))
        check(lib.rj_state_download(self._h, out.ctypes.data))
        return out

    def download_to_pointer(self, host_ptr):
        check(lib.rj_state_download(self._h, host_ptr))

    # streaming transfers (page-locked host memory; overlap with the sweeps of the module's stream)
    def upload_begin(self, host_ptr):
        check(lib.rj_state_upload_begin(self._h, host_ptr))

    def upload_end(self):
        check(lib.rj_state_upload_end(self._h))

    def download_begin(self, host_ptr):
        check(lib.rj_state_download_begin(self._h, host_ptr))

    def download_end(self):
        check(lib.rj_state_download_end(self._h))

    def download_precomputed(self):
        out = np.zeros((self.module.n_owned, self.module.n_precomputed))
        check(lib.rj_state_download_precomputed(self._h, out.ctypes.data))
        return out

    def interpolate(self, t=0.0):
        check(lib.rj_state_interpolate(self._h, float(t)))

    def sadd(self, s, b, src):
        check(lib.rj_state_sadd(self._h, float(s), float(b), src._h))

    def swap(self, other):
        check(lib.rj_state_swap(self._h, other._h))

    @property
    def device_U(self):
        return lib.rj_state_device_U(self._h)

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.rj_state_destroy(h)
            self._h = None


SCRATCH = {"alpha": 0, "dij": 1, "lij": 2, "lij_next": 3, "pij": 4, "r": 5, "bounds": 6}


class HyperbolicModule:
    """source/hyperbolic_module.h:110-278"""

    def __init__(self, offline_data, hyperbolic_system=None, initial_values=None, parameters=None, comm=None,
                 stream=None):
        self.offline_data = offline_data
        self.hyperbolic_system = hyperbolic_system or HyperbolicSystem()
        self.initial_values = initial_values or InitialValues()
        self.parameters = parameters or ModuleParameters()
        self.comm = comm
        h = C.c_void_p()
        check(
            lib.rj_hyperbolic_module_create(
                C.byref(h), offline_data._h, C.byref(self.hyperbolic_system.desc), C.byref(self.parameters.desc),
                C.byref(self.initial_values.desc), comm._h if comm is not None else None, stream,
            )
        )
        self._h = h
        sizes = (C.c_uint64 * 6)()
        check(lib.rj_hyperbolic_module_sizes(self._h, sizes))
        self.k, self.n_precomputed, self.n_bounds, self.n_owned, self.ld, self.n_entries = (int(x) for x in sizes)
        self.nnz = offline_data.host.nnz

    def set_initial_precomputed(self, values):
        """initial_precomputed_ of the reference (hyperbolic_module.h:309): the bathymetry for the
        shallow water equations, one value per owned row."""
        a = np.ascontiguousarray(values, dtype=np.float64)
        assert a.size == self.n_owned
        check(lib.rj_hyperbolic_module_set_initial_precomputed(self._h, a.ctypes.data))

    # -- reference API ---------------------------------------------------------
    def prepare_state_vector(self, state, t):
        check(lib.rj_hyperbolic_module_prepare_state_vector(self._h, state._h, float(t)))

    def step(self, old_state, stage_states, stage_weights, new_state, tau=0.0, tau_max=np.finfo(np.float64).max):
        """template<int stages> Number step(old, stage_state_vectors, stage_weights, new, tau, tau_max);
        raises Restart on an invariant domain violation with strategy raise_exception."""
        n = len(stage_states)
        arr = (C.c_void_p * max(n, 1))(*[s._h for s in stage_states])
        w = (C.c_double * max(n, 1))(*[float(x) for x in stage_weights])
        tau_out = C.c_double()
        check(
            lib.rj_hyperbolic_module_step(self._h, old_state._h, n, arr, w, new_state._h, float(tau), float(tau_max),
                                          C.byref(tau_out))
        )
        return tau_out.value

    @property
    def cfl(self):
        return lib.rj_hyperbolic_module_cfl(self._h)

    @cfl.setter
    def cfl(self, value):
        check(lib.rj_hyperbolic_module_set_cfl(self._h, float(value)))

    @property
    def id_violation_strategy(self):
        return self._strategy if hasattr(self, "_strategy") else IDViolationStrategy.warn

    @id_violation_strategy.setter
    def id_violation_strategy(self, value):
        check(lib.rj_hyperbolic_module_set_id_violation_strategy(self._h, int(value)))
        self._strategy = IDViolationStrategy(value)

    def n_restarts(self):
        return self._counters()[0]

    def n_warnings(self):
        return self._counters()[1]

    def _counters(self):
        r, w = C.c_uint32(), C.c_uint32()
        check(lib.rj_hyperbolic_module_counters(self._h, C.byref(r), C.byref(w)))
        return r.value, w.value

    # -- extras --------------------------------------------------------------------
    def state_vector(self):
        return StateVector(self)

    def synchronize(self):
        check(lib.rj_hyperbolic_module_synchronize(self._h))

    @property
    def stream(self):
        return lib.rj_hyperbolic_module_stream(self._h)

    def kernel_launches(self):
        return int(lib.rj_hyperbolic_module_kernel_launches(self._h))

    PROFILE_SLOTS = ("boundary_conditions", "precompute", "dij", "alpha", "boundary_dij", "dii_tau", "low_order",
                     "pij_lij", "high_order_next_lij", "high_order_final", "plumbing", "comm_wait", "allreduce")

    def set_profiling(self, on):
        """per-kernel CUDA-event timing on the module's stream (bench.py roofline)"""
        check(lib.rj_hyperbolic_module_set_profiling(self._h, int(bool(on))))

    def profile(self):
        """{kernel: (total_ms, launches)} accumulated since set_profiling(True)"""
        ms = (C.c_double * len(self.PROFILE_SLOTS))()
        cnt = (C.c_uint64 * len(self.PROFILE_SLOTS))()
        check(lib.rj_hyperbolic_module_get_profile(self._h, ms, cnt))
        return {name: (ms[q], int(cnt[q])) for q, name in enumerate(self.PROFILE_SLOTS)}

    def scratch(self, name):
        """alpha / dij / lij / lij_next / pij / r / bounds as host arrays in the
        host's CSR entry order (parity tests)."""
        sizes = {
            "alpha": self.n_owned, "dij": self.nnz, "lij": self.nnz, "lij_next": self.nnz,
            "pij": self.nnz * self.k, "r": self.n_owned * self.k, "bounds": self.n_owned * self.n_bounds,
        }
        out = np.zeros(sizes[name])
        check(lib.rj_hyperbolic_module_get_scratch(self._h, SCRATCH[name], out.ctypes.data))
        return out

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.rj_hyperbolic_module_destroy(h)
            self._h = None


class Postprocessor:
    """Postprocessor (source/postprocessor.h): schlieren and vorticity plots of a state vector on
    the exponential scale, computed on the device.  Parameters as in the reference's prm section
    ("schlieren beta" = 10, "schlieren quantities" = [rho], "vorticity quantities" = [])."""

    def __init__(self, module, schlieren_beta=10.0, schlieren_quantities=("rho",), vorticity_quantities=()):
        self.module = module
        self.beta = float(schlieren_beta)
        names = ["rho"] + ["m_%d" % (d + 1) for d in range(module.offline_data.dim)] + ["E"]
        if module.k == module.offline_data.dim + 1:  # shallow water: (h, m)
            names = ["h"] + ["m_%d" % (d + 1) for d in range(module.offline_data.dim)]
        self._index = {name: q for q, name in enumerate(names)}
        self._index["m"] = 1
        self.schlieren_quantities = list(schlieren_quantities)
        self.vorticity_quantities = list(vorticity_quantities)
        self.bounds = {}

    def compute(self, state):
        """-> {"schlieren_rho": array(n_owned), "vorticity_m": ...}"""
        out = {}
        for mode, prefix, quantities in ((0, "schlieren_", self.schlieren_quantities), (1, "vorticity_", self.vorticity_quantities)):
            for name in quantities:
                values = np.zeros(self.module.n_owned)
                bounds = (C.c_double * 2)()
                check(lib.rj_postprocessor_compute(self.module._h, state._h, mode, self._index[name], self.beta,
                                                   values.ctypes.data_as(_lib.f64p), bounds))
                out[prefix + name] = values
                self.bounds[prefix + name] = (bounds[0], bounds[1])
        return out


class TimeIntegrator:
    """source/time_integrator.template.h:207-443 (explicit schemes)"""

    def __init__(self, module, scheme=TimeSteppingScheme.erk_33, cfl_min=0.45, cfl_max=0.90,
                 cfl_recovery_strategy=CFLRecoveryStrategy.bang_bang_control):
        self.module = module
        h = C.c_void_p()
        check(lib.rj_time_integrator_create(C.byref(h), module._h, int(scheme), float(cfl_min), float(cfl_max),
                                            int(cfl_recovery_strategy)))
        self._h = h

    def step(self, state, t, t_final=None):
        tau = C.c_double()
        check(lib.rj_time_integrator_step(self._h, state._h, float(t), float("inf") if t_final is None else float(t_final),
                                          C.byref(tau)))
        return tau.value

    @property
    def efficiency(self):
        return lib.rj_time_integrator_efficiency(self._h)

    @property
    def stepping(self):
        return Stepping(lib.rj_time_integrator_stepping(self._h))

    @stepping.setter
    def stepping(self, value):
        check(lib.rj_time_integrator_set_stepping(self._h, int(value)))

    def graph_stats(self):
        """(graphs captured, graph launches)"""
        st = (C.c_uint64 * 2)()
        check(lib.rj_time_integrator_graph_stats(self._h, st))
        return int(st[0]), int(st[1])

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.rj_time_integrator_destroy(h)
            self._h = None
