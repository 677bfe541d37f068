"""The interior path of the second advection kernel (fluidx12_b200/csrc/advect_body.cuh), run on the CPU
(tests/emu/advect_emu.cpp compiles the same statements with g++) against the oracle's CSAdvect stage, bit for bit.

Voxels the interior path declines (a tap outside the grid) are left to the first kernel's general code and are
excluded here; the test also checks that those are a thin shell, and that the colour-free shortcut was exercised."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tests.util import smooth_state

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libadvect_emu.so"])
    lib = C.CDLL(os.path.join(_HERE, "libadvect_emu.so"))
    lib.advect_emu_run.restype = C.c_longlong
    return lib


def _p(a):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def run_interior(emu, oracle_mod, vel, col, dt, window=None):
    """window = (z_first, nz_alloc, z_own0, z_own1) of a z-slab (arrays hold the window), None = whole grid."""
    import fluidx12_b200 as fx
    nzl, ny, nx, _ = vel.shape
    nz = nzl if window is None else window[4]
    z_first, nz_alloc, z0, z1 = (0, nzl, 0, nzl) if window is None else window[:4]
    box = (C.c_int32 * 6)()
    assert fx.lib().fxb_emitter_box(nx, ny, nz, box) == 0
    ex0, ey0, ez0, ex1, ey1, ez1 = box
    basis = np.zeros((max(ez1 - ez0, 1), max(ey1 - ey0, 1), max(ex1 - ex0, 1)), np.float32)
    L = oracle_mod.lib()
    for z in range(ez0, ez1):
        for y in range(ey0, ey1):
            for x in range(ex0, ex1):
                basis[z - ez0, y - ey0, x - ex0] = L.fxo_emitter_basis(nx, ny, nz, x, y, z)
    pos = [((np.arange(n, dtype=np.float32) + np.float32(0.5)) / np.float32(n)).astype(np.float32) for n in (nx, ny, nz)]
    geom = np.array([nx, ny, nz, z_first, nz_alloc, z0, z1, ex0, ey0, ez0, ex1, ey1, ez1], np.int32)
    vo, co = np.full_like(vel, np.nan), np.full_like(col, np.nan)
    handled = np.zeros(vel.shape[:3], np.uint8)
    n = emu.advect_emu_run(_p(geom), _p(pos[0]), _p(pos[1]), _p(pos[2]), _p(basis), C.c_float(dt), _p(vel), _p(col),
                           _p(vo), _p(co), _p(handled))
    assert n == int(handled.sum())
    return vo, co, handled.astype(bool)


@pytest.mark.parametrize("n,umax,seed", [((48, 48, 40), 1.5, 3), ((64, 64, 24), 4.0, 4), ((32, 32, 32), 0.0, 5)])
def test_interior_path_matches_oracle(emu, oracle_mod, n, umax, seed):
    vel, col, _ = smooth_state(*n, seed=seed, umax=max(umax, 1e-3))
    if umax == 0.0:
        vel[...] = 0  # quiescent grid: every voxel takes the exact-texel case
    # smoke only in a blob, so that the colour-free shortcut and the smoke boundary are both exercised;
    # a few -0 texels to hit the exact-texel exception
    zz, yy, xx = np.meshgrid(*(np.arange(k) for k in (n[2], n[1], n[0])), indexing="ij")
    blob = (xx - n[0] / 2) ** 2 + (yy - n[1] / 3) ** 2 + (zz - n[2] / 2) ** 2 < (n[0] / 5) ** 2
    col[~blob] = 0
    col[2, 3, 4, 1] = np.float16(-0.0)
    vel[5, 6, 7, 0] = np.float16(-0.0)
    dt = oracle_mod.dt_for_grid(*n)
    want_v, want_c = oracle_mod.advect(vel, col, dt)
    got_v, got_c, handled = run_interior(emu, oracle_mod, vel, col, dt)
    assert handled.mean() > 0.55, handled.mean()
    assert handled[4:-4, 6:-6, 6:-6].mean() > (0.9 if umax <= 1.5 else 0.6)
    assert np.array_equal(got_v[handled][:, :3].view(np.uint16), want_v[handled][:, :3].view(np.uint16))
    assert (got_v[handled][:, 3].view(np.uint16) == 0).all()
    assert np.array_equal(got_c[handled].view(np.uint16), want_c[handled].view(np.uint16))
    assert np.isnan(got_v[~handled].astype(np.float32)).all()  # declined voxels are not written


def test_interior_path_emitter_and_developed_flow(emu, oracle_mod):
    """A developed emitter-driven state: emitter voxels, saturating colour, smoke boundary."""
    n = (64, 64, 64)
    f = oracle_mod.FluidOracle(*n)
    dt = oracle_mod.dt_for_grid(*n)
    for _ in range(20):
        f.step(dt)
    vel, col = f.get_field(oracle_mod.FIELD_VEL), f.get_field(oracle_mod.FIELD_COLOR)
    want_v, want_c = oracle_mod.advect(vel, col, dt)
    got_v, got_c, handled = run_interior(emu, oracle_mod, vel, col, dt)
    assert handled.mean() > 0.9
    assert np.array_equal(got_v[handled][:, :3].view(np.uint16), want_v[handled][:, :3].view(np.uint16))
    assert np.array_equal(got_c[handled].view(np.uint16), want_c[handled].view(np.uint16))
    assert (want_c[handled].astype(np.float32) > 0).any() and (want_c[handled].astype(np.float32) == 0).any()


def test_interior_path_on_a_slab_window(emu, oracle_mod):
    """z-slab window: taps must stay inside the local planes, results equal the oracle's slab-window stage."""
    n = (40, 40, 48)
    vel, col, _ = smooth_state(*n, seed=8, umax=1.0)
    dt = oracle_mod.dt_for_grid(*n)
    z_first, nz_alloc, z0, z1 = 10, 30, 16, 34
    wv, wc = vel[z_first:z_first + nz_alloc].copy(), col[z_first:z_first + nz_alloc].copy()
    want_v, want_c = oracle_mod.advect_slab(wv, wc, dt, n[2], z_first)
    got_v, got_c, handled = run_interior(emu, oracle_mod, wv, wc, dt, window=(z_first, nz_alloc, z0, z1, n[2]))
    assert not handled[:z0 - z_first].any() and not handled[z1 - z_first:].any()  # only owned planes are produced
    own = handled[z0 - z_first:z1 - z_first]
    assert own.mean() > 0.7
    assert np.array_equal(got_v[handled][:, :3].view(np.uint16), want_v[handled][:, :3].view(np.uint16))
    assert np.array_equal(got_c[handled].view(np.uint16), want_c[handled].view(np.uint16))
