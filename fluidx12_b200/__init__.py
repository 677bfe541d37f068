"""fluidx12_b200 — B200-native replacement of FluidX12's per-frame smoke-solver step.

The product is ``libfluidx_b200.so`` (C ABI in ``include/fluidx_b200.h``: C++ host code driving
hand-written sm_100a CUDA kernels).  This package is the thin ctypes binding used by the tests and
``bench.py`` plus a Python mirror of the reference ``Fluid`` class's simulation surface
(``Init`` / ``UpdateFrame`` / ``Simulate``, FluidX12/Content/Fluid.h:20-33).

There is no CPU fallback: importing works anywhere, but creating a ``Fluid`` raises unless the CUDA
library is built and an sm_100 device is present.
"""
from .binding import (  # noqa: F401
    ADDRESS_CLAMP,
    ADDRESS_MIRROR,
    FIELD_COLOR,
    FIELD_COLOR_PREV,
    FIELD_PRESSURE,
    FIELD_VELOCITY,
    FIELD_VELOCITY_ADVECTED,
    HALO_FUSED,
    HALO_NCCL,
    HALO_PEER,
    FluidError,
    FxbConfig,
    FxbLightParams,
    FxbStats,
    FxbViewParams,
    FxbVolumeHeader,
    dt_for_grid,
    lib,
    lib_path,
)
from .fluid import Fluid, FluidEZ  # noqa: F401
from .slab import gather_plan, halo_plan, slab_range  # noqa: F401
from . import volume  # noqa: F401

__all__ = [
    "Fluid", "FluidEZ", "FluidError", "FxbConfig", "FxbStats", "dt_for_grid", "lib", "lib_path", "slab_range", "halo_plan", "gather_plan", "volume", "FxbVolumeHeader", "FxbLightParams", "FxbViewParams",
    "ADDRESS_MIRROR", "ADDRESS_CLAMP", "FIELD_VELOCITY", "FIELD_COLOR", "FIELD_PRESSURE",
    "FIELD_VELOCITY_ADVECTED", "FIELD_COLOR_PREV", "HALO_PEER", "HALO_NCCL", "HALO_FUSED",
]
