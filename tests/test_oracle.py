"""CPU tests of the oracle (no GPU): unit, known-answer, property, cross-restatement, golden.

The reference has no tests — parity is unpinned — so the oracle is pinned by
(1) the fp32 literals of the reference's compiled DXBC blobs, (2) an independent numpy restatement,
(3) known-answer cases derived by hand from the HLSL, and (4) committed regression checksums.
"""
import json
import os
import struct

import numpy as np
import pytest

from tests.util import smooth_state

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def bits(f):
    return struct.unpack("<I", struct.pack("<f", float(f)))[0]


# ---- unit ------------------------------------------------------------------------------------
def test_constants_match_dxbc(oracle_mod):
    """Every literal the oracle uses is a literal of the reference's shipped shader blobs."""
    gold = json.load(open(os.path.join(GOLDEN, "dxbc_literals.json")))
    for blob in gold.values():
        assert all(blob["literals_found"].values())
    c = oracle_mod.constants()
    names = ["0.5/0.48", "1/0.03", "1/6", "0.25", "0.97", "0.001", "0.2", "-0.1", "log2e", "exp(-4)", "r2_3d",
             "r2_2d", "192", "48", "200", "8", "16", "40"]
    want = [0x3f855556, 0x42055556, 0x3e2aaaab, 0x3e800000, 0x3f7851ec, 0x3a83126f, 0x3e4ccccd, 0xbdcccccd,
            0x3fb8aa3b, 0x3c960aae, 0x3b800000, 0x3a800000, 0x43400000, 0x42400000, 0x43480000, 0x41000000,
            0x41800000, 0x42200000]
    for n, v, w in zip(names, c, want):
        assert bits(v) == w, n
    a3 = gold["CSProject3D.cso"]["literals_found"]
    for w in (0x3f855556, 0x42055556, 0x3e2aaaab, 0x3f7851ec, 0x3a83126f):
        assert a3["0x%08x" % w]
    adv = gold["CSAdvect.cso"]["literals_found"]
    for w in (0x3e4ccccd, 0xbdcccccd, 0x3fb8aa3b, 0x3c960aae, 0x3b800000, 0x3a800000, 0x43400000, 0x42400000):
        assert adv["0x%08x" % w]


def test_fp16_round_trip_all_values(oracle_mod):
    L = oracle_mod.lib()
    h = np.arange(65536, dtype=np.uint16)
    f = h.view(np.float16).astype(np.float32)
    finite = np.isfinite(f)
    for i in np.flatnonzero(finite)[::97]:
        assert L.fxo_f16_to_f32(int(h[i])) == f[i]
        assert L.fxo_f32_to_f16(float(f[i])) == h[i]
    # RNE on ties and near-ties
    for v in (1.0 + 2.0 ** -11, 1.0 + 3 * 2.0 ** -11, 2.0 ** -25, 3 * 2.0 ** -25, 65519.0, 65520.0, -1e-8, 0.1, 1 / 3):
        assert L.fxo_f32_to_f16(v) == int(np.float32(v).astype(np.float16).view(np.uint16)), v


def test_address_modes(oracle_mod):
    L = oracle_mod.lib()
    W = 8
    mirror = {-1: 0, -2: 1, -8: 7, -9: 7, -16: 0, -17: 0, 8: 7, 9: 6, 15: 0, 16: 0, 17: 1, 3: 3}
    for i, want in mirror.items():
        assert L.fxo_address_tap(i, W, 0) == want, i
    for i, want in {-5: 0, -1: 0, 0: 0, 7: 7, 8: 7, 100: 7}.items():
        assert L.fxo_address_tap(i, W, 1) == want
    assert L.fxo_address_tap(1, 1, 0) == 0 and L.fxo_address_tap(-1, 1, 0) == 0


def test_trilinear_reproduces_linear_ramps(oracle_mod):
    n = 16
    z, y, x = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    f = np.zeros((n, n, n, 4), np.float16)
    f[..., 0] = 3.0
    f[..., 1] = x
    f[..., 2] = y * 0.5
    f[..., 3] = z * 0.25
    rng = np.random.default_rng(0)
    for _ in range(200):
        c = rng.uniform(1.0 / n, 1 - 1.0 / n, 3).astype(np.float32)
        out = oracle_mod.sample_trilinear(f, c[0], c[1], c[2])
        t = c.astype(np.float64) * n - 0.5
        np.testing.assert_allclose(out, [3.0, t[0], 0.5 * t[1], 0.25 * t[2]], rtol=0, atol=2e-5)
    # texel centres are reproduced exactly
    out = oracle_mod.sample_trilinear(f, np.float32(5.5 / n), np.float32(2.5 / n), np.float32(9.5 / n))
    assert list(out) == [3.0, 5.0, 1.0, 2.25]


def test_emitter_basis(oracle_mod):
    from oracle import numpy_restatement as R
    for shape in ((32, 32, 32), (1, 64, 64)):
        nz, ny, nx = shape
        ref = R.emitter_basis_f64(shape)
        got = np.array([[[oracle_mod.lib().fxo_emitter_basis(nx, ny, nz, x, y, z) for x in range(nx)]
                         for y in range(ny)] for z in range(nz)], np.float32)
        big = ref > 1e-6
        assert np.abs(got[big] / ref[big] - 1).max() < 2e-5
        assert ((got >= np.float32(0.0183156393)) == (ref >= np.exp(-4.0)))[np.abs(ref - np.exp(-4.0)) > 1e-5].all()


# ---- known answers -----------------------------------------------------------------------------
def test_first_step_from_zero_state(oracle_mod):
    """Zero state -> after advect only emitter voxels are non-zero, with the hand-derived values."""
    n = 32
    dt = oracle_mod.dt_for_grid(n, n, n)
    assert dt == 2.0 / n
    o = oracle_mod.FluidOracle(n, n, n)
    o.step(dt)
    v1 = o.get_field(oracle_mod.FIELD_VEL_ADVECTED).astype(np.float32)
    c = o.get_field(oracle_mod.FIELD_COLOR).astype(np.float32)
    basis = np.array([[[oracle_mod.lib().fxo_emitter_basis(n, n, n, x, y, z) for x in range(n)] for y in range(n)]
                      for z in range(n)], np.float32)
    hit = basis >= np.float32(0.0183156393)
    assert hit.sum() > 10
    assert (v1[~hit] == 0).all() and (c[~hit] == 0).all()
    atten = np.float32(max(1.0 - np.float32(0.2) * np.float32(dt), 0))
    want_vy = (np.float32(192.0) * basis * np.float32(dt) * atten).astype(np.float16).astype(np.float32)
    np.testing.assert_array_equal(v1[..., 1][hit], want_vy[hit])
    want_a = (np.minimum(basis * np.float32(dt) * np.float32(40.0), 1).astype(np.float32) * atten).astype(np.float16)
    np.testing.assert_array_equal(c[..., 3][hit], want_a.astype(np.float32)[hit])
    # pressure: quiescent far-field cells froze at exactly 0
    p = o.get_field(oracle_mod.FIELD_PRESSURE)
    assert p[0, -1, 0] == 0 and np.abs(p).max() > 0
    assert 1 <= o.s_exec <= 64
    h = o.active_hist()
    assert h[0] == n ** 3 and (np.diff(h[:o.s_exec]) <= 0).all()


def test_dt_zero_is_identity(oracle_mod):
    """dt = 0: advect is the identity (atten 1, taps land on texel centres), project copies, parity holds."""
    n = 16
    vel, col, p = smooth_state(n, n, n)
    o = oracle_mod.FluidOracle(n, n, n)
    o.set_field(oracle_mod.FIELD_VEL, vel)
    o.set_field(oracle_mod.FIELD_COLOR_PREV, col)
    o.set_field(oracle_mod.FIELD_PRESSURE, p)
    o.step(0.0)
    np.testing.assert_array_equal(o.get_field(oracle_mod.FIELD_VEL)[..., :3], vel[..., :3])
    np.testing.assert_array_equal(o.get_field(oracle_mod.FIELD_COLOR), col)  # colour[p] <- colour[!p], no flip
    np.testing.assert_array_equal(o.get_field(oracle_mod.FIELD_PRESSURE), p)
    assert o.s_exec == 0


def test_create_rejects_non_square(oracle_mod):
    with pytest.raises(ValueError):
        oracle_mod.FluidOracle(32, 16, 8)


# ---- cross-restatement -----------------------------------------------------------------------------
@pytest.mark.parametrize("shape,clamp", [((12, 12, 12), False), ((10, 14, 14), True), ((1, 24, 24), False)])
def test_cpp_oracle_equals_numpy_restatement(oracle_mod, shape, clamp):
    from oracle import numpy_restatement as R
    nz, ny, nx = shape
    basis = np.array([[[oracle_mod.lib().fxo_emitter_basis(nx, ny, nz, x, y, z) for x in range(nx)]
                       for y in range(ny)] for z in range(nz)], np.float32)
    vel, col, p = smooth_state(nx, ny, nz, seed=7, umax=3.0)
    o = oracle_mod.FluidOracle(nx, ny, nz, address_mode=int(clamp))
    r = R.NumpyFluid(nx, ny, nz, basis, clamp=clamp)
    o.set_field(oracle_mod.FIELD_VEL, vel)
    o.set_field(oracle_mod.FIELD_COLOR, col)
    o.set_field(oracle_mod.FIELD_PRESSURE, p)
    r.vel[0], r.col[0], r.p = vel.copy(), col.copy(), p.copy()
    dt = oracle_mod.dt_for_grid(nx, ny, nz)
    for step in range(4):
        o.step(dt)
        r.step(dt)
        assert o.s_exec == r.s_exec, step
        np.testing.assert_array_equal(o.get_field(oracle_mod.FIELD_VEL_ADVECTED)[..., :3], r.vel[1][..., :3])
        np.testing.assert_array_equal(o.get_field(oracle_mod.FIELD_COLOR), r.col[r.parity])
        np.testing.assert_array_equal(o.get_field(oracle_mod.FIELD_PRESSURE), r.p)
        np.testing.assert_array_equal(o.get_field(oracle_mod.FIELD_VEL)[..., :3], r.vel[0][..., :3])


def test_stage_functions_equal_full_step(oracle_mod):
    n = 16
    vel, col, p = smooth_state(n, n, n, seed=3)
    dt = oracle_mod.dt_for_grid(n, n, n)
    o = oracle_mod.FluidOracle(n, n, n)
    o.set_field(oracle_mod.FIELD_VEL, vel)
    o.set_field(oracle_mod.FIELD_COLOR, col)
    o.set_field(oracle_mod.FIELD_PRESSURE, p)
    o.step(dt)
    v1, c1 = oracle_mod.advect(vel, col, dt)
    s = oracle_mod.divergence2x(v1)
    p1, s_exec, hist, active = oracle_mod.jacobi(s, p)
    v0 = oracle_mod.gradient(v1, p1)
    np.testing.assert_array_equal(o.get_field(oracle_mod.FIELD_VEL_ADVECTED), v1)
    np.testing.assert_array_equal(o.get_field(oracle_mod.FIELD_COLOR), c1)
    np.testing.assert_array_equal(o.get_field(oracle_mod.FIELD_PRESSURE), p1)
    np.testing.assert_array_equal(o.get_field(oracle_mod.FIELD_VEL), v0)
    assert s_exec == o.s_exec and not active.any()


# ---- properties ------------------------------------------------------------------------------------
def test_projection_reduces_divergence_and_colour_bounded(oracle_mod):
    n = 32
    o = oracle_mod.FluidOracle(n, n, n)
    dt = oracle_mod.dt_for_grid(n, n, n)
    for _ in range(12):
        o.step(dt)
    before = np.abs(oracle_mod.divergence2x(o.get_field(oracle_mod.FIELD_VEL_ADVECTED))[2:-2, 2:-2, 2:-2]).sum()
    after = np.abs(oracle_mod.divergence2x(o.get_field(oracle_mod.FIELD_VEL))[2:-2, 2:-2, 2:-2]).sum()
    assert after < 0.9 * before
    c = o.get_field(oracle_mod.FIELD_COLOR).astype(np.float32)
    assert c.min() >= 0 and c.max() <= 1.0


def test_2d_mirror_symmetry(oracle_mod):
    """2D emitter sits at x = 0.5 with no vortex term: u.y and colour symmetric, u.x antisymmetric in x."""
    n = 64
    o = oracle_mod.FluidOracle(n, n, 1)
    dt = oracle_mod.dt_for_grid(n, n, 1)
    assert dt == 1.0 / n
    for _ in range(20):
        o.step(dt)
    v = o.get_field(oracle_mod.FIELD_VEL).astype(np.float32)[0]
    c = o.get_field(oracle_mod.FIELD_COLOR).astype(np.float32)[0]
    assert np.abs(v[..., 1]).max() > 0.1
    assert np.abs(v[..., 1] - v[:, ::-1, 1]).max() < 2e-2 * np.abs(v[..., 1]).max()
    assert np.abs(v[..., 0] + v[:, ::-1, 0]).max() < 2e-2 * max(np.abs(v[..., 0]).max(), 1e-3) + 1e-3
    assert np.abs(c - c[:, ::-1]).max() < 2e-2
    assert (v[..., 2] == 0).all()


def test_early_exit_off_runs_all_sweeps(oracle_mod):
    n = 16
    o = oracle_mod.FluidOracle(n, n, n, early_exit=False, iters=20)
    o.step(oracle_mod.dt_for_grid(n, n, n))
    assert o.s_exec == 20 and (o.active_hist()[:20] == n ** 3).all()


# ---- regression checksums --------------------------------------------------------------------------
def _checksums(o, oracle_mod):
    import hashlib
    out = {}
    for name, f in (("vel", oracle_mod.FIELD_VEL), ("col", oracle_mod.FIELD_COLOR), ("p", oracle_mod.FIELD_PRESSURE)):
        a = o.get_field(f)
        if a.ndim == 4:
            a = np.ascontiguousarray(a[..., :3]) if name == "vel" else a
        out[name] = hashlib.sha256(a.tobytes()).hexdigest()
    return out


CASES = {
    "3d_32_mirror_10": dict(n=(32, 32, 32), mode=0, steps=10),
    "3d_24_clamp_10": dict(n=(24, 24, 24), mode=1, steps=10),
    "3d_30x30x18_mirror_8": dict(n=(30, 30, 18), mode=0, steps=8),
    "2d_64_mirror_20": dict(n=(64, 64, 1), mode=0, steps=20),
}


def run_case(oracle_mod, case):
    nx, ny, nz = case["n"]
    o = oracle_mod.FluidOracle(nx, ny, nz, address_mode=case["mode"])
    dt = oracle_mod.dt_for_grid(nx, ny, nz)
    s = []
    for _ in range(case["steps"]):
        o.step(dt)
        s.append(o.s_exec)
    return o, s


@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_checksums(oracle_mod, name):
    """Self-generated regression fixtures (tests/golden/make_oracle_golden.py); NOT reference outputs."""
    gold = json.load(open(os.path.join(GOLDEN, "oracle_checksums.json")))[name]
    o, s = run_case(oracle_mod, CASES[name])
    assert s == gold["s_exec"]
    assert _checksums(o, oracle_mod) == gold["sha256"]
