"""Python mirror of the reference ``Fluid`` class's simulation surface over the C ABI.

Reference: FluidX12/Content/Fluid.h:20-33 (``Init`` / ``UpdateFrame`` / ``Simulate``), callers
FluidX12/FluidX12.cpp:197-201, :282, :536.  Method names and argument order follow the reference;
the D3D12-only arguments (command list, descriptor-table library, render-target formats, view and
projection matrices) are accepted and ignored so reference-style call sites keep working.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import binding as B


class Fluid:
    FrameCount = 3  # Fluid.h:35 — the reference's constant-buffer ring; kept for call-site parity only

    def __init__(self):
        self._h: Optional[C.c_void_p] = None
        self._cfg: Optional[B.FxbConfig] = None
        self.m_gridSize = (0, 0, 0)
        self.m_timeStep = 0.0
        self.m_frameParity = 0
        self._slab = (0, 0)

    # -- Fluid::Init (Fluid.cpp:189-270) ------------------------------------------------------
    def Init(self, pCommandList=None, width: int = 0, height: int = 0, descriptorTableLib=None, uploaders=None,
             rtFormat=None, dsFormat=None, gridSize: Sequence[int] = (128, 128, 128), *,
             address_mode: int = B.ADDRESS_MIRROR, early_exit: bool = True, jacobi_iters: int = 64,
             fuse_t: int = 0, device: int = 0, rank: int = 0, nranks: int = 1, h_adv: int = 0,
             use_graph: bool = True, kernel_path: int = 0, phase_timing: bool = False,
             halo_backend: int = B.HALO_FUSED, jacobi_group: int = 0,
             nccl_unique_id: Optional[bytes] = None) -> bool:
        """Returns False on failure like the reference (XUSG_N_RETURN); ``last_error`` says why."""
        L = B.lib()
        if self._h:
            self.close()
        cfg = B.FxbConfig()
        B.check(L.fxb_config_default(C.byref(cfg)))
        cfg.nx, cfg.ny, cfg.nz = (int(v) for v in gridSize)
        cfg.address_mode = address_mode
        cfg.early_exit = int(early_exit)
        cfg.jacobi_iters = jacobi_iters
        cfg.fuse_t = fuse_t
        cfg.device = device
        cfg.rank, cfg.nranks = rank, nranks
        cfg.h_adv = h_adv
        cfg.use_graph = int(use_graph)
        cfg.kernel_path = kernel_path
        cfg.phase_timing = int(phase_timing)
        cfg.halo_backend = halo_backend
        cfg.jacobi_group = jacobi_group
        self._uid_buf = None
        if nccl_unique_id is not None:
            self._uid_buf = C.create_string_buffer(bytes(nccl_unique_id), 128)
            cfg.nccl_unique_id = C.cast(self._uid_buf, C.c_void_p)
        h = C.c_void_p()
        rc = L.fxb_create(C.byref(cfg), C.byref(h))
        if rc != B.FXB_OK:
            self.last_error = L.fxb_last_error().decode("utf-8", "replace")
            self.last_status = rc
            return False
        self._h, self._cfg = h, cfg
        self.m_gridSize = (cfg.nx, cfg.ny, cfg.nz)
        self.m_frameParity = 0
        z0, cnt = C.c_uint32(), C.c_uint32()
        B.check(L.fxb_get_slab(self._h, C.byref(z0), C.byref(cnt)))
        self._slab = (int(z0.value), int(cnt.value))
        return True

    def _handle(self):
        if not self._h:
            raise B.FluidError(B.FXB_ERR_INVALID, "Fluid.Init has not succeeded")
        return self._h

    def close(self):
        if self._h:
            B.lib().fxb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- Fluid::UpdateFrame (Fluid.cpp:283-346) -------------------------------------------------
    def UpdateFrame(self, timeStep: float, frameIndex: int = 0, view=None, proj=None, eyePt=None) -> None:
        B.check(B.lib().fxb_update_frame(self._handle(), float(timeStep)))
        self.m_timeStep = float(timeStep)
        if timeStep > 0.0:
            self.m_frameParity ^= 1

    # -- Fluid::Simulate (Fluid.cpp:348-410) ----------------------------------------------------
    def Simulate(self, pCommandList=None, frameIndex: int = 0) -> None:
        """``pCommandList`` is the CUDA stream to enqueue on (int/`cudaStream_t`, None = default)."""
        stream = C.c_void_p(int(pCommandList)) if pCommandList else C.c_void_p(0)
        B.check(B.lib().fxb_simulate(self._handle(), stream))

    # -- helpers outside the reference surface ---------------------------------------------------
    def step(self, dt: float, stream=None) -> None:
        self.UpdateFrame(dt)
        self.Simulate(stream)

    def sync(self) -> None:
        B.check(B.lib().fxb_sync(self._handle()))

    @property
    def slab(self):
        """(z0, count) of the planes this rank owns."""
        return self._slab

    def _host_array(self, field: int) -> np.ndarray:
        nx, ny, _ = self.m_gridSize
        nzl = self._slab[1]
        if field == B.FIELD_PRESSURE:
            return np.empty((nzl, ny, nx), np.float32)
        return np.empty((nzl, ny, nx, 4), np.float16)

    def get_field(self, field: int) -> np.ndarray:
        a = self._host_array(field)
        B.check(B.lib().fxb_get_field(self._handle(), field, a.ctypes.data_as(C.c_void_p), a.nbytes))
        return a

    def set_field(self, field: int, a: np.ndarray) -> None:
        ref = self._host_array(field)
        a = np.ascontiguousarray(a, dtype=ref.dtype)
        if a.shape != ref.shape:
            raise B.FluidError(B.FXB_ERR_SIZE, f"set_field: shape {a.shape} != {ref.shape}")
        B.check(B.lib().fxb_set_field(self._handle(), field, a.ctypes.data_as(C.c_void_p), a.nbytes))

    def get_field_async(self, field: int, host_ptr: int, nbytes: int, stream=None) -> None:
        stream = C.c_void_p(int(stream)) if stream else C.c_void_p(0)
        B.check(B.lib().fxb_get_field_async(self._handle(), field, C.c_void_p(host_ptr), nbytes, stream))

    # -- Fluid::rayMarchL (Fluid.cpp:857-878): the light-map pass of Fluid::Render's default mode ------------------
    def RayMarchL(self, params: "B.FxbLightParams" = None, pCommandList=None) -> None:
        """Enqueues CSRayMarchL over m_colors[m_frameParity] on the CUDA stream ``pCommandList``."""
        params = params if params is not None else B.FxbLightParams.reference_defaults()
        stream = C.c_void_p(int(pCommandList)) if pCommandList else C.c_void_p(0)
        B.check(B.lib().fxb_light_map(self._handle(), C.byref(params), stream))

    def get_light_map(self) -> np.ndarray:
        """m_lightMap (this rank's planes) as [z][y][x] uint32 words in R11G11B10_FLOAT packing."""
        nx, ny, _ = self.m_gridSize
        a = np.empty((self._slab[1], ny, nx), np.uint32)
        B.check(B.lib().fxb_get_light_map(self._handle(), a.ctypes.data_as(C.c_void_p), a.nbytes))
        return a

    # -- Fluid::rayMarchV (Fluid.cpp:880-908): view rays into one mip of the cube map -------------------------------
    def RayMarchV(self, params: "B.FxbViewParams", pCommandList=None) -> None:
        """Enqueues CSRayMarchV (needs the light map of ``RayMarchL``) on the CUDA stream ``pCommandList``."""
        stream = C.c_void_p(int(pCommandList)) if pCommandList else C.c_void_p(0)
        B.check(B.lib().fxb_ray_march_v(self._handle(), C.byref(params), stream))
        self._cube_size = int(params.cube_size)

    # -- Fluid::rayMarch (Fluid.cpp:825-855): the march without the separate light pass ------------------------------
    def RayMarch(self, view: "B.FxbViewParams", light: "B.FxbLightParams" = None, pCommandList=None) -> None:
        """Enqueues CSRayMarch: light (and occlusion) rays are cast at every view sample; ``light.num_samples`` is the
        light-ray sample count."""
        light = light if light is not None else B.FxbLightParams.reference_defaults()
        stream = C.c_void_p(int(pCommandList)) if pCommandList else C.c_void_p(0)
        B.check(B.lib().fxb_ray_march(self._handle(), C.byref(view), C.byref(light), stream))
        self._cube_size = int(view.cube_size)

    def get_cube_map(self) -> np.ndarray:
        """The cube-map mip last written: [6][S][S][4] UNORM8 (faces +X, -X, +Y, -Y, +Z, -Z)."""
        s = getattr(self, "_cube_size", 0)
        a = np.empty((6, s, s, 4), np.uint8)
        B.check(B.lib().fxb_get_cube_map(self._handle(), a.ctypes.data_as(C.c_void_p), a.nbytes))
        return a

    def export(self, path: str, field: int = B.FIELD_COLOR) -> None:
        """Writes this rank's slab of ``field`` as a volume file (fluidx12_b200/volume.py): by default the colour
        field ``Fluid::Render`` would sample, m_colors[m_frameParity] (Fluid.cpp:760-770, 841)."""
        B.check(B.lib().fxb_export_field(self._handle(), field, path.encode()))

    def stats(self) -> B.FxbStats:
        st = B.FxbStats()
        B.check(B.lib().fxb_get_stats(self._handle(), C.byref(st)))
        return st

    def post_stats(self, slot: int) -> None:
        """Enqueues a snapshot of the step record into pinned slot ``slot`` (0..3) behind the work enqueued so far."""
        B.check(B.lib().fxb_post_stats(self._handle(), slot))

    def wait_stats(self, slot: int) -> B.FxbStats:
        """Blocks until the snapshot posted to ``slot`` has landed and returns it (the device keeps running)."""
        st = B.FxbStats()
        B.check(B.lib().fxb_wait_stats(self._handle(), slot, C.byref(st)))
        return st

    def freeze_histogram(self, n: int = 64) -> np.ndarray:
        """Cells still active after sweep k+1 of the last step, k < n (the oracle's active_hist shifted by one)."""
        h = np.zeros(n, np.uint64)
        B.check(B.lib().fxb_get_freeze_histogram(self._handle(), h.ctypes.data_as(C.c_void_p), n))
        return h

    def profile_step(self):
        """One un-graphed step timed per phase: dict of milliseconds."""
        ms = (C.c_float * 6)()
        B.check(B.lib().fxb_profile_step(self._handle(), ms, 6))
        keys = ("advect", "divergence", "jacobi", "gradient", "halo", "step")
        return {k: float(ms[i]) for i, k in enumerate(keys)}

    def state_checksum(self) -> tuple:
        """Three 64-bit words (velocity.xyz, colour, pressure) of this rank's slab; rank sums mod 2^64 are
        decomposition-independent (fxb_state_checksum)."""
        out = (C.c_uint64 * 3)()
        B.check(B.lib().fxb_state_checksum(self._handle(), out))
        return tuple(int(v) for v in out)

    def phase_times(self, reset: bool = False) -> dict:
        """Device time per phase accumulated over the steps run so far (needs ``phase_timing=True`` at Init): ms."""
        ms = (C.c_double * 4)()
        B.check(B.lib().fxb_get_phase_times(self._handle(), ms, 4, int(reset)))
        return {k: float(ms[i]) for i, k in enumerate(("advect", "divergence", "jacobi", "gradient"))}


class FluidEZ(Fluid):
    """Mirror of the reference's ``FluidEZ`` (FluidX12/Content/FluidEZ.h:20-32), the app's default runtime path
    (FluidX12.cpp:40).  Its simulation is the same two dispatches (FluidEZ.cpp:373-449); the one difference on the
    hot path is the advection sampler, LINEAR_CLAMP instead of LINEAR_MIRROR (FluidEZ.cpp:406), and ``Init`` takes no
    descriptor-table library."""

    def Init(self, pCommandList=None, width: int = 0, height: int = 0, uploaders=None, rtFormat=None, dsFormat=None,
             gridSize: Sequence[int] = (128, 128, 128), **kw) -> bool:
        kw.setdefault("address_mode", B.ADDRESS_CLAMP)
        return super().Init(pCommandList, width, height, None, uploaders, rtFormat, dsFormat, gridSize, **kw)

