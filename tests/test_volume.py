"""Volume files: the renderer hand-off format.  Host-only entry points of the C library, the
independent numpy reader, slab assembly and the error behaviour need no GPU; the export of a live field does
(tests/test_zzy_gpu_volume.py)."""
import ctypes as C
import os

import numpy as np
import pytest

import fluidx12_b200 as fx
from fluidx12_b200 import binding as B
from fluidx12_b200 import volume


def colour(shape, seed=3):
    r = np.random.default_rng(seed)
    return r.random(shape + (4,), np.float32).astype(np.float16)


def test_header_is_the_64_byte_wire_structure():
    assert C.sizeof(B.FxbVolumeHeader) == 64 == volume._HEADER.size
    offs = {n: getattr(B.FxbVolumeHeader, n).offset for n, _ in B.FxbVolumeHeader._fields_}
    assert offs == {"magic": 0, "version": 4, "nx": 8, "ny": 12, "nz": 16, "z0": 20, "nz_local": 24, "field": 28,
                    "format": 32, "flags": 36, "frame": 40, "dt": 48, "frame_parity": 52, "payload_bytes": 56}


def test_round_trip_through_the_library_and_the_numpy_reader(tmp_path):
    a = colour((5, 6, 8))
    p = str(tmp_path / "c.fxbv")
    volume.write(p, a, field=fx.FIELD_COLOR, grid=(8, 6, 5), frame=17, dt=1 / 3, frame_parity=1)
    assert os.path.getsize(p) == 64 + a.nbytes and not os.path.exists(p + ".tmp")
    for reader in (volume.read, volume.read_numpy):
        b, h = reader(p)
        assert b.dtype == np.float16 and np.array_equal(a.view(np.uint16), b.view(np.uint16))
        assert h["magic"] == b"FXBV" and h["version"] == 1 and (h["nx"], h["ny"], h["nz"]) == (8, 6, 5)
        assert (h["z0"], h["nz_local"], h["field"], h["format"]) == (0, 5, fx.FIELD_COLOR, volume.FORMAT_HALF4)
        assert h["flags"] == volume.FLAG_PREMULTIPLIED and h["frame"] == 17 and h["frame_parity"] == 1
        assert h["dt"] == np.float32(1 / 3) and h["payload_bytes"] == a.nbytes
    # pressure: float payload, no colour flag
    q = np.random.default_rng(1).standard_normal((5, 6, 8)).astype(np.float32)
    volume.write(p, q, field=fx.FIELD_PRESSURE, grid=(8, 6, 5))
    b, h = volume.read_numpy(p)
    assert b.dtype == np.float32 and np.array_equal(q.view(np.uint32), b.view(np.uint32))
    assert h["format"] == volume.FORMAT_FLOAT and h["flags"] == 0
    # the raw bytes are the documented layout: payload starts at byte 64, x fastest
    raw = open(p, "rb").read()
    assert raw[:4] == b"FXBV" and np.frombuffer(raw, np.float32, offset=64)[8 * 6 + 8 + 3] == q[1, 1, 3]


def test_slab_files_assemble_into_the_whole_field(tmp_path):
    a = colour((12, 4, 8))
    paths = []
    for r in range(3):
        z0, z1 = fx.slab_range(12, r, 3)
        paths.append(str(tmp_path / ("c%d.fxbv" % r)))
        volume.write(paths[-1], a[z0:z1], field=fx.FIELD_COLOR, grid=(8, 4, 12), z0=z0, frame=9)
    whole, h = volume.assemble(reversed(paths))
    assert np.array_equal(whole.view(np.uint16), a.view(np.uint16)) and h["nz_local"] == 12 and h["z0"] == 0
    with pytest.raises(ValueError):
        volume.assemble(paths[:2])
    with pytest.raises(ValueError):
        volume.assemble([paths[0], paths[2]])
    volume.write(paths[1], a[4:8], field=fx.FIELD_COLOR, grid=(8, 4, 12), z0=4, frame=10)  # another frame
    with pytest.raises(ValueError):
        volume.assemble(paths)


def test_errors(tmp_path):
    L = fx.lib()
    a = colour((2, 4, 4))
    p = str(tmp_path / "x.fxbv")
    with pytest.raises(fx.FluidError) as e:   # shape does not fit the grid
        volume.write(p, a, field=fx.FIELD_COLOR, grid=(8, 4, 2))
    assert e.value.code == B.FXB_ERR_SIZE
    with pytest.raises(fx.FluidError) as e:   # slab outside the grid
        volume.write(p, a, field=fx.FIELD_COLOR, grid=(4, 4, 2), z0=1)
    assert e.value.code == B.FXB_ERR_INVALID
    with pytest.raises(fx.FluidError) as e:   # unwritable path
        volume.write(str(tmp_path / "no" / "dir.fxbv"), a, field=fx.FIELD_COLOR, grid=(4, 4, 2))
    assert e.value.code == B.FXB_ERR_IO
    h = B.FxbVolumeHeader()
    assert L.fxb_volume_write(None, C.byref(h), a.ctypes.data_as(C.c_void_p)) == B.FXB_ERR_INVALID
    assert L.fxb_volume_read_header(p.encode(), C.byref(h)) == B.FXB_ERR_IO        # missing file
    assert b"cannot open" in L.fxb_last_error()
    volume.write(p, a, field=fx.FIELD_COLOR, grid=(4, 4, 2))
    raw = open(p, "rb").read()
    open(p, "wb").write(raw[:-1])                                                   # truncated payload
    with pytest.raises(fx.FluidError) as e:
        volume.read(p)
    assert e.value.code == B.FXB_ERR_IO
    with pytest.raises(ValueError):
        volume.read_numpy(p)
    open(p, "wb").write(raw + b"\0")                                                # trailing bytes
    with pytest.raises(fx.FluidError):
        volume.read(p)
    with pytest.raises(ValueError):
        volume.read_numpy(p)
    open(p, "wb").write(b"XXXX" + raw[4:])                                          # wrong magic
    assert L.fxb_volume_read_header(p.encode(), C.byref(h)) == B.FXB_ERR_IO
    open(p, "wb").write(raw[:4] + (2).to_bytes(4, "little") + raw[8:])              # future version
    assert L.fxb_volume_read_header(p.encode(), C.byref(h)) == B.FXB_ERR_IO
    open(p, "wb").write(raw)
    small = np.empty(a.size - 1, np.float16)                                        # buffer too small
    assert L.fxb_volume_read(p.encode(), C.byref(h), small.ctypes.data_as(C.c_void_p), small.nbytes) == B.FXB_ERR_SIZE
    assert L.fxb_export_field(None, 1, p.encode()) == B.FXB_ERR_INVALID
