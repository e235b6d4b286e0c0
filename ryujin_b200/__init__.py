"""ryujin_b200: B200-native explicit hyperbolic update (ryujin's HyperbolicModule::step()).

Hand-written sm_100a CUDA kernels behind a C ABI (include/ryujin_b200.h); this
package is the thin host-side mirror of the reference interface for that path:
OfflineData, HyperbolicModule, StateVector, TimeIntegrator.  No CPU fallback.
"""
from ._lib import EXPORTED_SYMBOLS, LIB_PATH, NoDeviceError, Restart, RyujinError, lib  # noqa: F401
from .hyperbolic_module import (  # noqa: F401
    CFLRecoveryStrategy,
    Communicator,
    Equation,
    EquationOfState,
    HyperbolicModule,
    HyperbolicSystem,
    IDViolationStrategy,
    InitialState,
    InitialValues,
    ModuleParameters,
    Postprocessor,
    StateVector,
    Stepping,
    TimeIntegrator,
    TimeSteppingScheme,
)
from .offline_data import Boundary, CartesianMesh, HostOfflineData, OfflineData  # noqa: F401
from .stencil_io import StencilData  # noqa: F401
from .time_loop import NotOnThePath, TimeLoop  # noqa: F401


def device_count():
    return int(lib.rj_device_count())
