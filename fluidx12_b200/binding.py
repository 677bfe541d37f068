"""ctypes declarations of the C ABI in ``include/fluidx_b200.h`` (kept in the same order)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_NAME = "libfluidx_b200.so"

ADDRESS_MIRROR, ADDRESS_CLAMP = 0, 1
HALO_PEER, HALO_NCCL, HALO_FUSED = 0, 1, 2
FIELD_VELOCITY, FIELD_COLOR, FIELD_PRESSURE, FIELD_VELOCITY_ADVECTED, FIELD_COLOR_PREV = range(5)

FXB_OK, FXB_ERR_INVALID, FXB_ERR_CUDA, FXB_ERR_NCCL, FXB_ERR_SIZE, FXB_ERR_HALO_OVERFLOW, FXB_ERR_IO = 0, -1, -2, -3, -4, -5, -6

# Every symbol include/fluidx_b200.h declares (tests check the library exports exactly these).
EXPORTS = (
    "fxb_config_default", "fxb_create", "fxb_destroy", "fxb_update_frame", "fxb_simulate", "fxb_sync",
    "fxb_dt_for_grid", "fxb_get_slab", "fxb_get_field", "fxb_set_field", "fxb_get_field_async", "fxb_get_stats",
    "fxb_post_stats", "fxb_wait_stats", "fxb_p2p_plan", "fxb_jacobi_schedule", "fxb_face_last_order", "fxb_emitter_box", "fxb_get_freeze_histogram", "fxb_profile_step", "fxb_get_phase_times", "fxb_state_checksum", "fxb_nccl_unique_id", "fxb_last_error", "fxb_abi_version",
    "fxb_volume_write", "fxb_volume_read_header", "fxb_volume_read", "fxb_export_field",
    "fxb_light_map", "fxb_get_light_map", "fxb_cube_visibility_mask", "fxb_estimate_cube_lod", "fxb_ray_march_v", "fxb_ray_march", "fxb_get_cube_map",
)


class FluidError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"fluidx_b200 error {code}: {message}")
        self.code = code


class FxbConfig(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("nx", C.c_uint32), ("ny", C.c_uint32), ("nz", C.c_uint32),
        ("address_mode", C.c_int32),
        ("early_exit", C.c_int32),
        ("jacobi_iters", C.c_int32),
        ("fuse_t", C.c_int32),
        ("device", C.c_int32),
        ("rank", C.c_int32), ("nranks", C.c_int32),
        ("h_adv", C.c_int32),
        ("use_graph", C.c_int32),
        ("kernel_path", C.c_int32),
        ("phase_timing", C.c_int32),
        ("halo_backend", C.c_int32),
        ("jacobi_group", C.c_int32),
        ("nccl_unique_id", C.c_void_p),
    ]


class FxbStats(C.Structure):
    _fields_ = [
        ("s_exec", C.c_int32),
        ("jacobi_passes", C.c_int32),
        ("fuse_t", C.c_int32),
        ("halo_overflow", C.c_int32),
        ("frame_parity", C.c_int32),
        ("kernels_per_step", C.c_int32),
        ("steps", C.c_uint64),
        ("active_after_first_sweep", C.c_uint64),
        ("total_sweeps", C.c_uint64),
        ("total_passes", C.c_uint64),
        ("bricks_processed", C.c_uint64),
        ("bricks_copied", C.c_uint64),
        ("brick_cells", C.c_uint64),
        ("bricks_per_pass", C.c_uint64),
        ("jacobi_fused", C.c_int32),
        ("tail_from", C.c_int32),
    ]


class FxbLightParams(C.Structure):
    """fxb_light_params: the constants CSRayMarchL reads (include/fluidx_b200.h); defaults = the reference's
    (Fluid.cpp:171-175, 182: light at (75, 75, -75), colour (1, .7, .3) x 3 pi, ambient (1, 1, 1) x 1.5 pi, volume
    scaled by 10, 64 light samples, no light probe)."""
    _fields_ = [("light_pt", C.c_float * 3), ("light_color", C.c_float * 4), ("ambient", C.c_float * 4),
                ("world_i", C.c_float * 12), ("world", C.c_float * 12), ("num_samples", C.c_uint32),
                ("has_light_probes", C.c_uint32), ("sh", (C.c_float * 3) * 9)]

    @classmethod
    def reference_defaults(cls) -> "FxbLightParams":
        import math
        p = cls()
        p.light_pt[:] = [75.0, 75.0, -75.0]
        p.light_color[:] = [1.0, 0.7, 0.3, math.pi * 3.0]
        p.ambient[:] = [1.0, 1.0, 1.0, math.pi * 1.5]
        p.world[:] = [10.0, 0, 0, 0, 0, 10.0, 0, 0, 0, 0, 10.0, 0]
        p.world_i[:] = [0.1, 0, 0, 0, 0, 0.1, 0, 0, 0, 0, 0.1, 0]
        p.num_samples, p.has_light_probes = 64, 0
        return p


class FxbViewParams(C.Structure):
    """fxb_view_params: the constants CSRayMarchV reads (include/fluidx_b200.h)."""
    _fields_ = [("eye_pt", C.c_float * 3), ("world_i", C.c_float * 12), ("num_samples", C.c_uint32),
                ("visibility_mask", C.c_uint32), ("cube_size", C.c_uint32)]


class FxbVolumeHeader(C.Structure):
    """fxb_volume_header: the 64-byte header of a volume file (include/fluidx_b200.h)."""
    _fields_ = [
        ("magic", C.c_char * 4),
        ("version", C.c_uint32),
        ("nx", C.c_uint32), ("ny", C.c_uint32), ("nz", C.c_uint32),
        ("z0", C.c_uint32), ("nz_local", C.c_uint32),
        ("field", C.c_uint32),
        ("format", C.c_uint32),
        ("flags", C.c_uint32),
        ("frame", C.c_uint64),
        ("dt", C.c_float),
        ("frame_parity", C.c_uint32),
        ("payload_bytes", C.c_uint64),
    ]


def lib_path() -> str:
    # FXB_LIB: another build of the same library (tools/timing_probe.py loads the -DFXB_TIMING debug build)
    return os.environ.get("FXB_LIB") or os.path.join(_HERE, _LIB_NAME)


_lib = None


def lib() -> C.CDLL:
    """Loads the CUDA library; fails loudly when it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        path = lib_path()
        if not os.path.exists(path):
            raise FluidError(FXB_ERR_CUDA, f"{path} is missing: run `python -c 'import __graft_entry__ as g; "
                                           "g.build()'` (make -C fluidx12_b200/csrc); there is no CPU fallback")
        L = C.CDLL(path)
        vp = C.c_void_p
        L.fxb_config_default.argtypes = [C.POINTER(FxbConfig)]
        L.fxb_create.argtypes = [C.POINTER(FxbConfig), C.POINTER(vp)]
        L.fxb_destroy.argtypes = [vp]
        L.fxb_destroy.restype = None
        L.fxb_update_frame.argtypes = [vp, C.c_float]
        L.fxb_simulate.argtypes = [vp, vp]
        L.fxb_sync.argtypes = [vp]
        L.fxb_dt_for_grid.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_float)]
        L.fxb_get_slab.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.fxb_get_field.argtypes = [vp, C.c_int, vp, C.c_size_t]
        L.fxb_set_field.argtypes = [vp, C.c_int, vp, C.c_size_t]
        L.fxb_get_field_async.argtypes = [vp, C.c_int, vp, C.c_size_t, vp]
        L.fxb_get_stats.argtypes = [vp, C.POINTER(FxbStats)]
        L.fxb_post_stats.argtypes = [vp, C.c_int]
        L.fxb_wait_stats.argtypes = [vp, C.c_int, C.POINTER(FxbStats)]
        L.fxb_p2p_plan.argtypes = [C.c_int32] * 5 + [C.POINTER(C.c_int64)]
        L.fxb_face_last_order.argtypes = [C.c_int32] * 5 + [C.POINTER(C.c_int32), C.c_int32]
        L.fxb_jacobi_schedule.argtypes = [C.c_int32] * 3 + [C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32]
        L.fxb_emitter_box.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_int32)]
        L.fxb_get_freeze_histogram.argtypes = [vp, vp, C.c_int]
        L.fxb_profile_step.argtypes = [vp, C.POINTER(C.c_float), C.c_int]
        L.fxb_state_checksum.argtypes = [vp, C.POINTER(C.c_uint64)]
        L.fxb_get_phase_times.argtypes = [vp, C.POINTER(C.c_double), C.c_int, C.c_int]
        L.fxb_nccl_unique_id.argtypes = [vp]
        L.fxb_volume_write.argtypes = [C.c_char_p, C.POINTER(FxbVolumeHeader), vp]
        L.fxb_volume_read_header.argtypes = [C.c_char_p, C.POINTER(FxbVolumeHeader)]
        L.fxb_volume_read.argtypes = [C.c_char_p, C.POINTER(FxbVolumeHeader), vp, C.c_size_t]
        L.fxb_export_field.argtypes = [vp, C.c_int, C.c_char_p]
        L.fxb_light_map.argtypes = [vp, C.POINTER(FxbLightParams), vp]
        L.fxb_get_light_map.argtypes = [vp, vp, C.c_size_t]
        L.fxb_cube_visibility_mask.argtypes = [C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
        L.fxb_estimate_cube_lod.argtypes = [C.POINTER(C.c_float), C.c_float, C.c_float, C.c_uint32, C.c_uint32, C.c_uint32,
                                            C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.fxb_ray_march_v.argtypes = [vp, C.POINTER(FxbViewParams), vp]
        L.fxb_ray_march.argtypes = [vp, C.POINTER(FxbViewParams), C.POINTER(FxbLightParams), vp]
        L.fxb_get_cube_map.argtypes = [vp, vp, C.c_size_t]
        L.fxb_last_error.restype = C.c_char_p
        L.fxb_abi_version.restype = C.c_int
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != FXB_OK:
        raise FluidError(rc, lib().fxb_last_error().decode("utf-8", "replace"))


def dt_for_grid(nx: int, ny: int, nz: int) -> float:
    """dt rule of FluidX::OnUpdate (FluidX12/FluidX12.cpp:266-267)."""
    out = C.c_float()
    check(lib().fxb_dt_for_grid(nx, ny, nz, C.byref(out)))
    return float(out.value)
