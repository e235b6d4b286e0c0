"""ctypes binding of libryujin_b200.so (include/ryujin_b200.h).

The shared library holds the hand-written sm_100a kernels and the C ABI; this
module only declares the foreign signatures.  There is no Python or CPU
fallback: if the library is missing the import fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# RYUJIN_B200_LIB selects an alternative build of the same library (tuning experiments)
LIB_PATH = os.environ.get("RYUJIN_B200_LIB") or os.path.join(_HERE, "libryujin_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "ryujin_b200: %s is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
        "or `make -C ryujin_b200/csrc`. There is no CPU fallback." % LIB_PATH
    )



def _preload_nccl():
    """libryujin_b200.so links libnccl.so.2 by soname.  PyTorch bundles a newer NCCL than the
    system one; whichever copy is mapped first serves both, so map PyTorch's copy first (if it
    exists) -- otherwise a later `import torch` fails on symbols missing from the older build."""
    import importlib.util

    for name in ("nvidia.nccl", "nvidia"):
        try:
            spec = importlib.util.find_spec(name)
        except (ImportError, ValueError):
            spec = None
        if spec is None or not spec.submodule_search_locations:
            continue
        for base in spec.submodule_search_locations:
            for cand in (os.path.join(base, "lib", "libnccl.so.2"), os.path.join(base, "nccl", "lib", "libnccl.so.2")):
                if os.path.exists(cand):
                    try:
                        C.CDLL(cand, mode=C.RTLD_GLOBAL)
                        return cand
                    except OSError:
                        pass
    return None


NCCL_PATH = _preload_nccl()
lib = C.CDLL(LIB_PATH)

RJ_OK, RJ_RESTART = 0, 1
RJ_ERR_INVALID, RJ_ERR_CUDA, RJ_ERR_NCCL, RJ_ERR_NUMERIC, RJ_ERR_NO_DEVICE = -1, -2, -3, -4, -5
RJ_UNIQUE_ID_BYTES = 128

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
i32p = C.POINTER(C.c_int32)
f64p = C.POINTER(C.c_double)


class OfflineDesc(C.Structure):
    """rj_offline_desc"""

    _fields_ = [
        ("dim", C.c_int),
        ("n_export", C.c_uint32),
        ("n_internal", C.c_uint32),
        ("n_owned", C.c_uint32),
        ("n_relevant", C.c_uint32),
        ("simd_length", C.c_uint32),
        ("row_starts", u64p),
        ("column_indices", u32p),
        ("mass_matrix", f64p),
        ("cij_matrix", f64p),
        ("lumped_mass_matrix", f64p),
        ("measure_of_omega", C.c_double),
        ("n_boundary", C.c_uint32),
        ("boundary_dof", u32p),
        ("boundary_normal", f64p),
        ("boundary_id", i32p),
        ("boundary_position", f64p),
        ("n_coupling_pairs", C.c_uint32),
        ("coupling_pairs", u32p),
        ("coupling_cji", f64p),
        ("n_neighbors", C.c_uint32),
        ("neighbor_rank", i32p),
        ("ghost_count", u32p),
        ("send_offsets", u32p),
        ("send_indices", u32p),
        ("positions", f64p),
    ]


class MeshDesc(C.Structure):
    """rj_mesh_desc"""

    _fields_ = [
        ("dim", C.c_int),
        ("n_cells", C.c_uint32 * 3),
        ("origin", C.c_double * 3),
        ("h", C.c_double * 3),
        ("cell_mask", u8p),
        ("box_boundary_ids", C.c_int32 * 6),
        ("masked_boundary_id", C.c_int32),
        ("partition", C.c_uint32 * 3),
    ]


class EquationDesc(C.Structure):
    """rj_equation_desc"""

    _fields_ = [
        ("equation", C.c_int),
        ("gamma", C.c_double),
        ("reference_density", C.c_double),
        ("vacuum_state_relaxation_small", C.c_double),
        ("vacuum_state_relaxation_large", C.c_double),
        ("gravity", C.c_double),
        ("manning_friction_coefficient", C.c_double),
        ("reference_water_depth", C.c_double),
        ("dry_state_relaxation_factor", C.c_double),
        ("dry_state_relaxation_small", C.c_double),
        ("dry_state_relaxation_large", C.c_double),
        ("equation_of_state", C.c_int32),
        ("compute_strict_bounds", C.c_int32),
        ("covolume_b", C.c_double),
        ("reference_pressure", C.c_double),
        ("reference_specific_internal_energy", C.c_double),
        ("vdw_a", C.c_double),
    ]


class ModuleParams(C.Structure):
    """rj_module_params"""

    _fields_ = [
        ("indicator_evc_factor", C.c_double),
        ("limiter_iterations", C.c_uint32),
        ("limiter_newton_tolerance", C.c_double),
        ("limiter_newton_max_iterations", C.c_uint32),
        ("limiter_relaxation_factor", C.c_double),
        ("riemann_newton_tolerance", C.c_double),
        ("riemann_newton_max_iterations", C.c_uint32),
        ("limit_on_kinetic_energy", C.c_int32),
        ("limit_on_square_velocity", C.c_int32),
    ]


class InitialValuesDesc(C.Structure):
    """rj_initial_values_desc"""

    _fields_ = [
        ("kind", C.c_int),
        ("direction", C.c_double * 3),
        ("position", C.c_double * 3),
        ("primitive", C.c_double * 3),
        ("primitive_right", C.c_double * 3),
        ("mach_number", C.c_double),
        ("beta", C.c_double),
        ("sw_params", C.c_double * 4),
    ]


vp = C.c_void_p

_SIGNATURES = {
    "rj_last_error": (C.c_char_p, []),
    "rj_device_count": (C.c_int, []),
    "rj_version": (C.c_char_p, []),
    "rj_pow_implementation": (C.c_int, []),
    "rj_pow_evaluate": (C.c_int, [C.c_int, C.c_int, C.c_size_t, f64p, f64p, f64p]),
    "rj_offline_create": (C.c_int, [C.POINTER(vp), C.POINTER(OfflineDesc), C.c_int]),
    "rj_offline_destroy": (None, [vp]),
    "rj_offline_create_host_layout": (C.c_int, [C.POINTER(vp), C.POINTER(OfflineDesc)]),
    "rj_offline_exchange_sizes": (C.c_int, [vp, u64p]),
    "rj_offline_exchange_get": (C.c_int, [vp, u32p, u32p, u32p]),
    "rj_offline_layout_sizes": (C.c_int, [vp, u64p]),
    "rj_offline_layout_get": (C.c_int, [vp, u64p, u32p, u32p, u32p]),
    "rj_mesh_generate": (C.c_int, [C.POINTER(vp), C.POINTER(MeshDesc), C.c_int, C.c_int]),
    "rj_host_offline_destroy": (None, [vp]),
    "rj_host_offline_desc": (C.c_int, [vp, C.POINTER(OfflineDesc)]),
    "rj_host_offline_global_indices": (C.c_int, [vp, u64p]),
    "rj_host_offline_n_global_dofs": (C.c_uint64, [vp]),
    "rj_decode_simd_pattern": (C.c_int, [C.c_uint32, C.c_uint32, C.c_uint32, u64p, u32p, u64p, u32p]),
    "rj_comm_unique_id": (C.c_int, [u8p]),
    "rj_comm_create": (C.c_int, [C.POINTER(vp), u8p, C.c_int, C.c_int, C.c_int]),
    "rj_comm_destroy": (None, [vp]),
    "rj_equation_desc_default": (None, [C.POINTER(EquationDesc), C.c_int]),
    "rj_module_params_default": (None, [C.POINTER(ModuleParams)]),
    "rj_initial_values_desc_default": (None, [C.POINTER(InitialValuesDesc), C.c_int]),
    "rj_state_create": (C.c_int, [C.POINTER(vp), vp]),
    "rj_state_destroy": (None, [vp]),
    "rj_state_upload": (C.c_int, [vp, vp]),
    "rj_state_download": (C.c_int, [vp, vp]),
    "rj_state_download_precomputed": (C.c_int, [vp, vp]),
    "rj_state_upload_begin": (C.c_int, [vp, vp]),
    "rj_state_upload_end": (C.c_int, [vp]),
    "rj_state_download_begin": (C.c_int, [vp, vp]),
    "rj_state_download_end": (C.c_int, [vp]),
    "rj_state_interpolate": (C.c_int, [vp, C.c_double]),
    "rj_state_device_U": (vp, [vp]),
    "rj_state_device_precomputed": (vp, [vp]),
    "rj_state_sadd": (C.c_int, [vp, C.c_double, C.c_double, vp]),
    "rj_state_swap": (C.c_int, [vp, vp]),
    "rj_hyperbolic_module_create": (
        C.c_int,
        [C.POINTER(vp), vp, C.POINTER(EquationDesc), C.POINTER(ModuleParams), C.POINTER(InitialValuesDesc), vp, vp],
    ),
    "rj_hyperbolic_module_destroy": (None, [vp]),
    "rj_hyperbolic_module_set_initial_precomputed": (C.c_int, [vp, vp]),
    "rj_hyperbolic_module_prepare_state_vector": (C.c_int, [vp, vp, C.c_double]),
    "rj_hyperbolic_module_step": (
        C.c_int,
        [vp, vp, C.c_int, C.POINTER(vp), f64p, vp, C.c_double, C.c_double, f64p],
    ),
    "rj_hyperbolic_module_set_cfl": (C.c_int, [vp, C.c_double]),
    "rj_hyperbolic_module_cfl": (C.c_double, [vp]),
    "rj_hyperbolic_module_set_id_violation_strategy": (C.c_int, [vp, C.c_int]),
    "rj_hyperbolic_module_counters": (C.c_int, [vp, u32p, u32p]),
    "rj_hyperbolic_module_device_alpha": (vp, [vp]),
    "rj_hyperbolic_module_synchronize": (C.c_int, [vp]),
    "rj_hyperbolic_module_stream": (vp, [vp]),
    "rj_hyperbolic_module_sizes": (C.c_int, [vp, u64p]),
    "rj_hyperbolic_module_get_scratch": (C.c_int, [vp, C.c_int, vp]),
    "rj_hyperbolic_module_kernel_launches": (C.c_uint64, [vp]),
    "rj_hyperbolic_module_set_profiling": (C.c_int, [vp, C.c_int]),
    "rj_hyperbolic_module_get_profile": (C.c_int, [vp, f64p, u64p]),
    "rj_postprocessor_compute": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_double, f64p, f64p]),
    "rj_time_integrator_create": (C.c_int, [C.POINTER(vp), vp, C.c_int, C.c_double, C.c_double, C.c_int]),
    "rj_time_integrator_destroy": (None, [vp]),
    "rj_time_integrator_step": (C.c_int, [vp, vp, C.c_double, C.c_double, f64p]),
    "rj_time_integrator_efficiency": (C.c_double, [vp]),
    "rj_time_integrator_set_stepping": (C.c_int, [vp, C.c_int]),
    "rj_time_integrator_stepping": (C.c_int, [vp]),
    "rj_time_integrator_graph_stats": (C.c_int, [vp, u64p]),
}

EXPORTED_SYMBOLS = tuple(sorted(_SIGNATURES))

for _name, (_res, _args) in _SIGNATURES.items():
    _fn = getattr(lib, _name)  # AttributeError if the library does not export it
    _fn.restype = _res
    _fn.argtypes = _args


class RyujinError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("ryujin_b200 error %d: %s" % (code, message))
        self.code = code


class NoDeviceError(RyujinError):
    pass


class Restart(Exception):
    """Mirror of ryujin::Restart (source/hyperbolic_module.h:59-61)."""


def check(rc):
    if rc == RJ_OK:
        return
    if rc == RJ_RESTART:
        raise Restart()
    msg = lib.rj_last_error().decode(errors="replace")
    if rc == RJ_ERR_NO_DEVICE:
        raise NoDeviceError(rc, msg)
    raise RyujinError(rc, msg)
