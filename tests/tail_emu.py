"""ctypes wrapper of tests/emu/libtail_emu.so — the CPU emulation of the tail kernel's body (test infrastructure)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-s", "-C", _HERE, "libtail_emu.so"])
        _LIB = C.CDLL(os.path.join(_HERE, "libtail_emu.so"))
        _LIB.tail_emu_launch.restype = C.c_int
    return _LIB


def _p(a):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


class BrickGrid:
    """Brick decomposition of jacobi_fused.cu: bricks of bx x by x bz output cells, brick id = (zc*nty + ty)*ntx + tx."""

    def __init__(self, nx, ny, nz, bx=120, by=12, bz=8):
        self.nx, self.ny, self.nz, self.bx, self.by, self.bz = nx, ny, nz, bx, by, bz
        self.ntx, self.nty, self.nzc = -(-nx // bx), -(-ny // by), -(-nz // bz)
        self.n = self.ntx * self.nty * self.nzc

    def region(self, brick):
        tx, ty, zc = brick % self.ntx, (brick // self.ntx) % self.nty, brick // (self.ntx * self.nty)
        return (slice(zc * self.bz, min((zc + 1) * self.bz, self.nz)), slice(ty * self.by, min((ty + 1) * self.by, self.ny)),
                slice(tx * self.bx, min((tx + 1) * self.bx, self.nx)))


def pack_mask(active):
    """[z][y][x] 0/1 -> bit-packed [z][y][nx/8], bit j of a byte = cell 8*byte + j (the layout of jacobi_fused.cu)."""
    return np.packbits(active.astype(np.uint8), axis=-1, bitorder="little")


def unpack_mask(mask, nx):
    return np.unpackbits(mask, axis=-1, bitorder="little")[..., :nx]


def thread_order(descending: bool):
    """Order in which the emulated threads run a barrier-free segment (a missing barrier shows up in one of the two)."""
    lib().tail_emu_thread_order(C.c_int(int(descending)))


def paths():
    """Work items that took the (copy, sparse, dense) path since the last call."""
    out = np.zeros(3, np.int64)
    lib().tail_emu_paths(_p(out))
    return tuple(int(v) for v in out)


def launch(g: BrickGrid, p_in, p_out, rhs, m_in, m_out, relax_in, copy_in, brick_state, hist_s0, first, early_exit=True,
           levels=4, tt=4, sparse_cap=-1, cp_async=1, slab=None, dense_mode=1):
    """One emulated launch.  Returns (relax_out, copy_out).  p_out / m_out / brick_state / hist_s0 are updated in place.

    slab = (nz_alloc, z_face_lo, z_face_hi, z_out0, z_out1) in local plane indices for a z-slab window (the arrays then
    hold nz_alloc planes and `g` describes the bricks of the owned planes only); None: the whole grid."""
    nz_alloc, z_face_lo, z_face_hi, z_out0, z_out1 = slab if slab else (g.nz, 0, g.nz, 0, g.nz)
    assert p_in.shape[0] == nz_alloc and z_out1 - z_out0 == g.nz
    geom = np.array([g.nx, g.ny, nz_alloc, z_face_lo, z_face_hi, z_out0, z_out1, g.bx, g.by, g.bz], np.int32)
    flags = np.array([int(first), int(early_exit), levels, tt, sparse_cap, cp_async, dense_mode], np.int32)
    relax_in = np.ascontiguousarray(relax_in, np.int32)
    copy_in = np.ascontiguousarray(copy_in, np.int32)
    relax_out, copy_out = np.full(g.n, -1, np.int32), np.full(g.n, -1, np.int32)
    nr, nc = np.zeros(1, np.int32), np.zeros(1, np.int32)
    rc = lib().tail_emu_launch(_p(geom), _p(flags), _p(p_in), _p(p_out), _p(rhs), _p(m_in), _p(m_out), _p(relax_in),
                               C.c_int(relax_in.size), _p(copy_in), C.c_int(copy_in.size), _p(relax_out), _p(nr),
                               _p(copy_out), _p(nc), _p(brick_state), _p(hist_s0))
    assert rc >= 0, "unsupported shape"
    lib().tail_emu_smem_overruns.restype = C.c_longlong
    overruns = lib().tail_emu_smem_overruns()
    assert overruns == 0, "%d work items wrote outside the shared memory the CUDA launch requests (S::kBytes)" % overruns
    return relax_out[:nr[0]].copy(), copy_out[:nc[0]].copy()
