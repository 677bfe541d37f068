"""ctypes binding of the CPU oracle (``oracle/fluid_oracle.cpp``).

TEST INFRASTRUCTURE ONLY — imported by ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.  The product package
``fluidx12_b200`` never imports this module.  Parity is pinned to the reference's shipped DXBC (interpreted:
``tests/golden/dxbc_interp.py``, ``tests/test_dxbc_golden.py``) and unpinned for the D3D platform semantics
(see the header of ``fluid_oracle.cpp``): the reference ships no tests or golden vectors and cannot run here.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle_fluid.so")

FIELD_VEL, FIELD_COLOR, FIELD_PRESSURE, FIELD_VEL_ADVECTED, FIELD_COLOR_PREV = range(5)
ADDRESS_MIRROR, ADDRESS_CLAMP = 0, 1


def build(force: bool = False) -> str:
    """Compile the oracle with the committed Makefile (g++ + OpenMP)."""
    src = os.path.join(_HERE, "fluid_oracle.cpp")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "liboracle_fluid.so"] + (["-B"] if force else []))
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, f32, u16p, f32p = C.c_void_p, C.c_int, C.c_float, C.c_void_p, C.c_void_p
        L.fxo_create.restype = vp
        L.fxo_create.argtypes = [i32] * 6
        L.fxo_destroy.argtypes = [vp]
        L.fxo_update_frame.argtypes = [vp, f32]
        L.fxo_simulate.argtypes = [vp]
        L.fxo_get_field.argtypes = [vp, i32, vp, C.c_size_t]
        L.fxo_set_field.argtypes = [vp, i32, vp, C.c_size_t]
        L.fxo_s_exec.argtypes = [vp]
        L.fxo_active_hist.argtypes = [vp, vp, i32]
        L.fxo_dt_for_grid.restype = f32
        L.fxo_dt_for_grid.argtypes = [i32] * 3
        L.fxo_advect.argtypes = [i32, i32, i32, i32, f32, u16p, u16p, u16p, u16p]
        L.fxo_divergence2x.argtypes = [i32, i32, i32, u16p, f32p]
        L.fxo_jacobi.argtypes = [i32, i32, i32, f32p, f32p, i32, i32, vp, vp]
        L.fxo_gradient.argtypes = [i32, i32, i32, u16p, f32p, u16p]
        L.fxo_advect_slab.argtypes = [i32, i32, i32, i32, i32, i32, f32, u16p, u16p, u16p, u16p]
        L.fxo_divergence2x_slab.argtypes = [i32, i32, i32, i32, i32, u16p, f32p]
        L.fxo_jacobi_sweeps_slab.argtypes = [i32, i32, i32, i32, i32, f32p, f32p, vp, i32, i32, i32, i32, vp]
        L.fxo_gradient_slab.argtypes = [i32, i32, i32, i32, i32, u16p, f32p, u16p]
        L.fxo_f32_to_f16.restype = C.c_uint16
        L.fxo_f32_to_f16.argtypes = [f32]
        L.fxo_f16_to_f32.restype = f32
        L.fxo_f16_to_f32.argtypes = [C.c_uint16]
        L.fxo_address_tap.argtypes = [i32, i32, i32]
        L.fxo_sample_trilinear.argtypes = [u16p, i32, i32, i32, i32, f32, f32, f32, f32p]
        L.fxo_emitter_basis.restype = f32
        L.fxo_emitter_basis.argtypes = [i32] * 6
        L.fxo_constants.argtypes = [f32p, i32]
        L.fxo_light_map.argtypes = [i32, i32, i32, u16p, vp, vp]
        L.fxo_light_map.restype = None
        L.fxo_ray_march_v.argtypes = [i32, i32, i32, u16p, vp, vp, vp]
        L.fxo_ray_march_v.restype = None
        L.fxo_ray_march.argtypes = [i32, i32, i32, u16p, vp, vp, vp]
        L.fxo_ray_march.restype = None
        L.fxo_unpack_r11g11b10.argtypes = [C.c_uint32, f32p]
        L.fxo_unpack_r11g11b10.restype = None
        L.fxo_pack_r11g11b10.argtypes = [f32, f32, f32]
        L.fxo_pack_r11g11b10.restype = C.c_uint32
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def threads(n: int = 0) -> int:
    """OpenMP threads the oracle uses: ``n > 0`` sets the count; returns the count in effect."""
    L = lib()
    L.fxo_threads.argtypes = [C.c_int]
    L.fxo_threads.restype = C.c_int
    return int(L.fxo_threads(int(n)))


def dt_for_grid(nx: int, ny: int, nz: int) -> float:
    return float(lib().fxo_dt_for_grid(nx, ny, nz))


class FluidOracle:
    """Mirror of the reference ``Fluid`` simulation surface (Init / UpdateFrame / Simulate)."""

    def __init__(self, nx, ny, nz, address_mode=ADDRESS_MIRROR, early_exit=True, iters=64):
        self.shape = (nz, ny, nx)
        self._h = lib().fxo_create(nx, ny, nz, address_mode, int(early_exit), iters)
        if not self._h:
            raise ValueError("fxo_create rejected the configuration (nx must equal ny)")
        self.iters = iters

    def close(self):
        if self._h:
            lib().fxo_destroy(self._h)
            self._h = None

    __del__ = close

    def update_frame(self, dt: float):
        lib().fxo_update_frame(self._h, dt)

    def simulate(self):
        lib().fxo_simulate(self._h)

    def step(self, dt: float):
        self.update_frame(dt)
        self.simulate()

    def _field_array(self, field):
        if field == FIELD_PRESSURE:
            return np.empty(self.shape, np.float32)
        return np.empty(self.shape + (4,), np.float16)

    def get_field(self, field) -> np.ndarray:
        a = self._field_array(field)
        rc = lib().fxo_get_field(self._h, field, _ptr(a), a.nbytes)
        assert rc == 0
        return a

    def set_field(self, field, a: np.ndarray):
        ref = self._field_array(field)
        a = np.ascontiguousarray(a, dtype=ref.dtype)
        assert a.shape == ref.shape, (a.shape, ref.shape)
        rc = lib().fxo_set_field(self._h, field, _ptr(a), a.nbytes)
        assert rc == 0

    @property
    def s_exec(self) -> int:
        return lib().fxo_s_exec(self._h)

    def active_hist(self) -> np.ndarray:
        h = np.zeros(max(self.iters, 1), np.int64)
        lib().fxo_active_hist(self._h, _ptr(h), h.size)
        return h


# ---- stage-level wrappers (kernel-by-kernel parity) ------------------------------------------
def advect(vel, col, dt, address_mode=ADDRESS_MIRROR):
    nz, ny, nx, _ = vel.shape
    vel = np.ascontiguousarray(vel, np.float16)
    col = np.ascontiguousarray(col, np.float16)
    vo, co = np.empty_like(vel), np.empty_like(col)
    lib().fxo_advect(nx, ny, nz, address_mode, dt, _ptr(vel), _ptr(col), _ptr(vo), _ptr(co))
    return vo, co


def divergence2x(vel):
    nz, ny, nx, _ = vel.shape
    vel = np.ascontiguousarray(vel, np.float16)
    s = np.empty((nz, ny, nx), np.float32)
    lib().fxo_divergence2x(nx, ny, nz, _ptr(vel), _ptr(s))
    return s


def jacobi(s, p, iters=64, early_exit=True):
    """Returns (p_out, s_exec, hist, active)."""
    nz, ny, nx = s.shape
    s = np.ascontiguousarray(s, np.float32)
    p = np.array(p, np.float32, order="C", copy=True)
    hist = np.zeros(max(iters, 1), np.int64)
    active = np.empty(s.shape, np.uint8)
    n = lib().fxo_jacobi(nx, ny, nz, _ptr(s), _ptr(p), iters, int(early_exit), _ptr(hist), _ptr(active))
    return p, n, hist, active


def gradient(vel, p):
    nz, ny, nx, _ = vel.shape
    vel = np.ascontiguousarray(vel, np.float16)
    p = np.ascontiguousarray(p, np.float32)
    out = np.empty_like(vel)
    lib().fxo_gradient(nx, ny, nz, _ptr(vel), _ptr(p), _ptr(out))
    return out


# ---- z-slab window wrappers (arrays are windows [z0, z0 + nzl) of a grid with nzg planes) ----------------------
def advect_slab(vel, col, dt, nzg, z0, address_mode=ADDRESS_MIRROR):
    nzl, ny, nx, _ = vel.shape
    vel = np.ascontiguousarray(vel, np.float16)
    col = np.ascontiguousarray(col, np.float16)
    vo, co = np.empty_like(vel), np.empty_like(col)
    lib().fxo_advect_slab(nx, ny, nzl, nzg, z0, address_mode, dt, _ptr(vel), _ptr(col), _ptr(vo), _ptr(co))
    return vo, co


def divergence2x_slab(vel, nzg, z0):
    nzl, ny, nx, _ = vel.shape
    vel = np.ascontiguousarray(vel, np.float16)
    s = np.empty((nzl, ny, nx), np.float32)
    lib().fxo_divergence2x_slab(nx, ny, nzl, nzg, z0, _ptr(vel), _ptr(s))
    return s


def jacobi_sweeps_slab(s, p, active, nsweeps, nzg, z0, c0, c1, early_exit=True):
    """In-place on copies; returns (p, active, counts[nsweeps])."""
    nzl, ny, nx = s.shape
    s = np.ascontiguousarray(s, np.float32)
    p = np.array(p, np.float32, order="C", copy=True)
    active = np.array(active, np.uint8, order="C", copy=True)
    counts = np.zeros(max(nsweeps, 1), np.int64)
    lib().fxo_jacobi_sweeps_slab(nx, ny, nzl, nzg, z0, _ptr(s), _ptr(p), _ptr(active), nsweeps, int(early_exit), c0, c1,
                                 _ptr(counts))
    return p, active, counts[:nsweeps]


def gradient_slab(vel, p, nzg, z0):
    nzl, ny, nx, _ = vel.shape
    vel = np.ascontiguousarray(vel, np.float16)
    p = np.ascontiguousarray(p, np.float32)
    out = np.empty_like(vel)
    lib().fxo_gradient_slab(nx, ny, nzl, nzg, z0, _ptr(vel), _ptr(p), _ptr(out))
    return out


def sample_trilinear(field, cx, cy, cz, address_mode=ADDRESS_MIRROR):
    nz, ny, nx, _ = field.shape
    field = np.ascontiguousarray(field, np.float16)
    out = np.empty(4, np.float32)
    lib().fxo_sample_trilinear(_ptr(field), nx, ny, nz, address_mode, cx, cy, cz, _ptr(out))
    return out


class LightParams(C.Structure):
    """The light-map pass's constants (CSRayMarchL.hlsl; = fxb_light_params of include/fluidx_b200.h)."""
    _fields_ = [("light_pt", C.c_float * 3), ("light_color", C.c_float * 4), ("ambient", C.c_float * 4),
                ("world_i", C.c_float * 12), ("world", C.c_float * 12), ("num_samples", C.c_uint32),
                ("has_light_probes", C.c_uint32), ("sh", (C.c_float * 3) * 9)]


def light_map(colour, params: LightParams) -> np.ndarray:
    """Packed R11G11B10_FLOAT light map [nz][ny][nx] of a colour field [nz][ny][nx][4] half."""
    nz, ny, nx, _ = colour.shape
    colour = np.ascontiguousarray(colour, np.float16)
    out = np.empty((nz, ny, nx), np.uint32)
    lib().fxo_light_map(nx, ny, nz, _ptr(colour), C.byref(params), _ptr(out))
    return out


class ViewParams(C.Structure):
    """The cube-map ray march's constants (CSRayMarchV; = fxb_view_params of include/fluidx_b200.h)."""
    _fields_ = [("eye_pt", C.c_float * 3), ("world_i", C.c_float * 12), ("num_samples", C.c_uint32),
                ("visibility_mask", C.c_uint32), ("cube_size", C.c_uint32)]


def ray_march_v(colour, light_map_words, params: ViewParams, cube=None) -> np.ndarray:
    """[6][S][S][4] UNORM8 cube map; `cube` (optional) holds the previous contents, kept where nothing is written."""
    nz, ny, nx, _ = colour.shape
    colour = np.ascontiguousarray(colour, np.float16)
    lm = np.ascontiguousarray(light_map_words, np.uint32)
    s = int(params.cube_size)
    out = np.zeros((6, s, s, 4), np.uint8) if cube is None else np.ascontiguousarray(cube, np.uint8).copy()
    lib().fxo_ray_march_v(nx, ny, nz, _ptr(colour), _ptr(lm), C.byref(params), _ptr(out))
    return out


def ray_march(colour, view: ViewParams, light: LightParams, cube=None) -> np.ndarray:
    """The non-separated march (CSRayMarch): the light is computed at every view sample; light.num_samples = light-ray
    samples.  [6][S][S][4] UNORM8."""
    nz, ny, nx, _ = colour.shape
    colour = np.ascontiguousarray(colour, np.float16)
    s = int(view.cube_size)
    out = np.zeros((6, s, s, 4), np.uint8) if cube is None else np.ascontiguousarray(cube, np.uint8).copy()
    lib().fxo_ray_march(nx, ny, nz, _ptr(colour), C.byref(view), C.byref(light), _ptr(out))
    return out


def unpack_r11g11b10(word: int) -> np.ndarray:
    out = np.zeros(3, np.float32)
    lib().fxo_unpack_r11g11b10(int(word), _ptr(out))
    return out


def pack_r11g11b10(r, g, b) -> int:
    return int(lib().fxo_pack_r11g11b10(r, g, b))


def constants() -> np.ndarray:
    out = np.zeros(18, np.float32)
    n = lib().fxo_constants(_ptr(out), out.size)
    return out[:n]
