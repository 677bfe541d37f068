"""Volume files: the hand-off format of a field to a renderer (layout in include/fluidx_b200.h).

The reference keeps the colour field on the GPU: ``Fluid::Render`` binds ``m_colors[m_frameParity]`` as a Texture3D
(FluidX12/Content/Fluid.cpp:760-770) and the ray marchers sample it as premultiplied RGBA, density in ``.w``
(Shaders/RayMarch.hlsli:62-68, Shaders/CSRayMarch.hlsl:157, Shaders/PSVisualizeColor.hlsl:24-33).  A volume file holds
exactly that texture's logical contents — 64-byte header, then ``[z][y][x][4]`` half or ``[z][y][x]`` float — so an
external renderer (or a D3D12 upload into the reference's own ``m_colors``) can consume a frame.

``write`` / ``read`` go through the C library's host-only entry points; ``read_numpy`` is an independent reader that
needs nothing but numpy (what a consumer without the library would write), and ``assemble`` joins the per-rank slab
files of a multi-GPU run."""
from __future__ import annotations

import ctypes as C
import struct
from typing import Iterable, Tuple

import numpy as np

from . import binding as B

MAGIC = b"FXBV"
VERSION = 1
FORMAT_HALF4, FORMAT_FLOAT = 1, 2
FLAG_PREMULTIPLIED = 1
_HEADER = struct.Struct("<4s9IQfIQ")  # the 64 bytes of fxb_volume_header, little-endian
_KEYS = ("magic", "version", "nx", "ny", "nz", "z0", "nz_local", "field", "format", "flags", "frame", "dt",
         "frame_parity", "payload_bytes")


def _shape_dtype(h) -> Tuple[tuple, type]:
    if h["format"] == FORMAT_HALF4:
        return (h["nz_local"], h["ny"], h["nx"], 4), np.float16
    return (h["nz_local"], h["ny"], h["nx"]), np.float32


def _as_dict(h: B.FxbVolumeHeader) -> dict:
    return {k: (bytes(getattr(h, k)) if k == "magic" else getattr(h, k)) for k in _KEYS}


def write(path: str, array: np.ndarray, *, field: int, grid: Tuple[int, int, int], z0: int = 0, frame: int = 0,
          dt: float = 0.0, frame_parity: int = 0) -> None:
    """Writes ``array`` (a rank's slab of ``field`` in the layout of ``Fluid.get_field``) as a volume file."""
    half4 = field != B.FIELD_PRESSURE
    a = np.ascontiguousarray(array, dtype=np.float16 if half4 else np.float32)
    nx, ny, nz = grid
    want = (a.shape[0], ny, nx, 4) if half4 else (a.shape[0], ny, nx)
    if a.ndim != len(want) or a.shape != want:
        raise B.FluidError(B.FXB_ERR_SIZE, f"volume.write: shape {a.shape} does not fit grid {grid}")
    h = B.FxbVolumeHeader()
    h.nx, h.ny, h.nz, h.z0, h.nz_local = nx, ny, nz, z0, a.shape[0]
    h.field, h.format = field, FORMAT_HALF4 if half4 else FORMAT_FLOAT
    h.flags = FLAG_PREMULTIPLIED if field in (B.FIELD_COLOR, B.FIELD_COLOR_PREV) else 0
    h.frame, h.dt, h.frame_parity, h.payload_bytes = frame, dt, frame_parity, a.nbytes
    B.check(B.lib().fxb_volume_write(path.encode(), C.byref(h), a.ctypes.data_as(C.c_void_p)))


def read_header(path: str) -> dict:
    h = B.FxbVolumeHeader()
    B.check(B.lib().fxb_volume_read_header(path.encode(), C.byref(h)))
    return _as_dict(h)


def read(path: str) -> Tuple[np.ndarray, dict]:
    """(array, header) through the C library."""
    h = read_header(path)
    shape, dtype = _shape_dtype(h)
    a = np.empty(shape, dtype)
    hh = B.FxbVolumeHeader()
    B.check(B.lib().fxb_volume_read(path.encode(), C.byref(hh), a.ctypes.data_as(C.c_void_p), a.nbytes))
    return a, _as_dict(hh)


def read_numpy(path: str) -> Tuple[np.ndarray, dict]:
    """The same with numpy alone: what a consumer without libfluidx_b200.so needs to implement."""
    with open(path, "rb") as fh:
        raw = fh.read(_HEADER.size)
        if len(raw) != _HEADER.size:
            raise ValueError("file shorter than a volume header")
        h = dict(zip(_KEYS, _HEADER.unpack(raw)))
        if h["magic"] != MAGIC or h["version"] != VERSION or h["format"] not in (FORMAT_HALF4, FORMAT_FLOAT):
            raise ValueError("not a version-%d volume file" % VERSION)
        shape, dtype = _shape_dtype(h)
        a = np.fromfile(fh, dtype=np.dtype(dtype).newbyteorder("<"), count=int(np.prod(shape)))
        if a.nbytes != h["payload_bytes"] or fh.read(1):
            raise ValueError("payload size does not match the header")
    return a.reshape(shape), h


def assemble(paths: Iterable[str]) -> Tuple[np.ndarray, dict]:
    """Joins the z-slab files the ranks of one multi-GPU frame wrote into the whole field."""
    parts = sorted((read_numpy(p) for p in paths), key=lambda ah: ah[1]["z0"])
    if not parts:
        raise ValueError("no files")
    h0 = parts[0][1]
    z = 0
    for a, h in parts:
        same = all(h[k] == h0[k] for k in ("nx", "ny", "nz", "field", "format", "flags", "frame", "frame_parity"))
        if not same or h["z0"] != z:
            raise ValueError("slab files do not tile one frame of one field (z0 = %d, expected %d)" % (h["z0"], z))
        z += h["nz_local"]
    if z != h0["nz"]:
        raise ValueError("slab files cover %d of %d planes" % (z, h0["nz"]))
    whole = dict(h0, z0=0, nz_local=h0["nz"], payload_bytes=sum(h["payload_bytes"] for _, h in parts))
    return np.concatenate([a for a, _ in parts], axis=0), whole
