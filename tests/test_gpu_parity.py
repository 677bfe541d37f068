"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical inputs.

Gates (BASELINE.json north_star): after 1 step max|a-b| <= 1e-5 * max|b| per field; after 100 steps
relative L2 <= 1e-3 per field.  With fp16 storage the 1-step gate is effectively bit-exactness, so the
tests also report (and for most cases require) exact equality.  Velocity .w is excluded (never read by
the reference).
"""
import numpy as np
import pytest

from tests.util import expected_passes, max_abs_rel, rel_l2, smooth_state

pytestmark = pytest.mark.gpu

TOL_1STEP = 1e-5   # of the field's max magnitude (north_star)
TOL_100STEP = 1e-3  # relative L2 (north_star)


@pytest.fixture(scope="module")
def fx():
    import fluidx12_b200 as fx
    fx.lib()
    return fx


def make_pair(fx, oracle_mod, n, mode=0, early_exit=True, iters=64, **kw):
    nx, ny, nz = n
    f = fx.Fluid()
    assert f.Init(gridSize=n, address_mode=mode, early_exit=early_exit, jacobi_iters=iters, **kw), f.last_error
    o = oracle_mod.FluidOracle(nx, ny, nz, address_mode=mode, early_exit=early_exit, iters=iters)
    return f, o


def fields(fx, oracle_mod):
    return (("velocity", fx.FIELD_VELOCITY, oracle_mod.FIELD_VEL), ("colour", fx.FIELD_COLOR, oracle_mod.FIELD_COLOR),
            ("pressure", fx.FIELD_PRESSURE, oracle_mod.FIELD_PRESSURE),
            ("advected", fx.FIELD_VELOCITY_ADVECTED, oracle_mod.FIELD_VEL_ADVECTED))


def compare(fx, oracle_mod, f, o, tol, metric=max_abs_rel, exact=False):
    for name, gf, of in fields(fx, oracle_mod):
        a, b = f.get_field(gf), o.get_field(of)
        if a.ndim == 4 and name != "colour":
            a, b = a[..., :3], b[..., :3]
        a32, b32 = a.astype(np.float32), b.astype(np.float32)
        assert np.isfinite(a32).all(), name
        err = metric(a32, b32)
        assert err <= tol, (name, err)
        if exact:
            assert np.array_equal(a, b), (name, int((a != b).sum()))


def inject(fx, oracle_mod, f, o, n, seed=1234, umax=2.0):
    vel, col, p = smooth_state(*n, seed=seed, umax=umax)
    for gf, of, a in ((fx.FIELD_VELOCITY, oracle_mod.FIELD_VEL, vel), (fx.FIELD_COLOR, oracle_mod.FIELD_COLOR, col),
                      (fx.FIELD_PRESSURE, oracle_mod.FIELD_PRESSURE, p)):
        f.set_field(gf, a)
        o.set_field(of, a)


@pytest.mark.parametrize("n,mode", [((64, 64, 64), 0), ((50, 50, 30), 0), ((48, 48, 48), 1), ((40, 40, 7), 0)])
def test_one_step_from_smooth_random_state(fx, oracle_mod, n, mode):
    f, o = make_pair(fx, oracle_mod, n, mode)
    inject(fx, oracle_mod, f, o, n)
    dt = fx.dt_for_grid(*n)
    f.step(dt); o.step(dt)
    f.sync()
    assert f.stats().s_exec == o.s_exec
    compare(fx, oracle_mod, f, o, TOL_1STEP, exact=True)
    f.step(dt); o.step(dt)
    compare(fx, oracle_mod, f, o, TOL_1STEP, exact=True)


def test_mirror_beyond_first_tap(fx, oracle_mod):
    """|u| large enough that taps land several texels outside the grid (MIRROR != CLAMP there)."""
    n = (32, 32, 32)
    for mode in (0, 1):
        f, o = make_pair(fx, oracle_mod, n, mode)
        inject(fx, oracle_mod, f, o, n, seed=5, umax=12.0)
        dt = fx.dt_for_grid(*n)
        f.step(dt); o.step(dt)
        compare(fx, oracle_mod, f, o, TOL_1STEP, exact=True)


def test_emitter_driven_100_steps_3d(fx, oracle_mod):
    n = (64, 64, 64)
    f, o = make_pair(fx, oracle_mod, n)
    dt = fx.dt_for_grid(*n)
    f.step(dt); o.step(dt)
    compare(fx, oracle_mod, f, o, TOL_1STEP, exact=True)
    sx = []
    for _ in range(99):
        f.step(dt); o.step(dt)
        sx.append((f.stats().s_exec, o.s_exec))
    assert all(a == b for a, b in sx), sx
    compare(fx, oracle_mod, f, o, TOL_100STEP, metric=rel_l2)
    compare(fx, oracle_mod, f, o, TOL_1STEP, exact=True)  # in practice the whole trajectory is bit-exact


def test_default_grid_128_cubed(fx, oracle_mod):
    """BASELINE.json configs[1]: 3D 128^3, reference default emitter and Jacobi count."""
    n = (128, 128, 128)
    f, o = make_pair(fx, oracle_mod, n)
    dt = fx.dt_for_grid(*n)
    for _ in range(12):
        f.step(dt); o.step(dt)
    assert f.stats().s_exec == o.s_exec
    compare(fx, oracle_mod, f, o, TOL_1STEP, exact=True)


def test_2d_path_256_squared(fx, oracle_mod):
    """BASELINE.json configs[0]: 2D 256^2 (CSAdvect + CSProject2D), dt = 1/256."""
    n = (256, 256, 1)
    f, o = make_pair(fx, oracle_mod, n)
    dt = fx.dt_for_grid(*n)
    assert dt == 1.0 / 256
    for _ in range(40):
        f.step(dt); o.step(dt)
    assert f.stats().s_exec == o.s_exec
    compare(fx, oracle_mod, f, o, TOL_1STEP, exact=True)


def test_dt_zero_and_parity(fx, oracle_mod):
    n = (32, 32, 32)
    f, o = make_pair(fx, oracle_mod, n)
    inject(fx, oracle_mod, f, o, n)
    dt = fx.dt_for_grid(*n)
    for step_dt in (dt, 0.0, 0.0, dt, 0.0, dt):
        f.step(step_dt); o.step(step_dt)
        st = f.stats()
        assert st.s_exec == o.s_exec
        compare(fx, oracle_mod, f, o, TOL_1STEP, exact=True)
    assert f.m_frameParity == f.stats().frame_parity == 1


def test_early_exit_off_and_short_iteration_counts(fx, oracle_mod):
    n = (32, 32, 32)
    for early, iters in ((False, 64), (True, 5), (False, 7), (True, 0), (True, 1)):
        f, o = make_pair(fx, oracle_mod, n, early_exit=early, iters=iters)
        inject(fx, oracle_mod, f, o, n, seed=11)
        dt = fx.dt_for_grid(*n)
        for _ in range(2):
            f.step(dt); o.step(dt)
        assert f.stats().s_exec == o.s_exec == (iters if not early else o.s_exec)
        compare(fx, oracle_mod, f, o, TOL_1STEP, exact=True)


@pytest.mark.parametrize("fuse_t", [1, 2, 3, 4])
@pytest.mark.parametrize("n", [(64, 64, 64), (136, 136, 50), (248, 248, 36), (40, 40, 7), (128, 128, 3)])
def test_fused_jacobi_every_T_and_ragged_tiles(fx, oracle_mod, n, fuse_t):
    """The fused pass (TMA-staged tile, register z queue, brick skipping) against the oracle for every fusion
    depth, on grids that are not multiples of the 120 x (32-2T) x bz brick and thinner than a brick."""
    f, o = make_pair(fx, oracle_mod, n, fuse_t=fuse_t)
    assert f.stats().jacobi_fused == 1 and f.stats().fuse_t == fuse_t
    inject(fx, oracle_mod, f, o, n, seed=21)
    dt = fx.dt_for_grid(*n)
    for _ in range(3):
        f.step(dt); o.step(dt)
        assert f.stats().s_exec == o.s_exec
        compare(fx, oracle_mod, f, o, TOL_1STEP, exact=True)


@pytest.mark.parametrize("resident_from", ["0", "3", "999"])
@pytest.mark.parametrize("n", [(64, 64, 64), (136, 136, 50), (150, 150, 20)])
def test_pass_kernels_are_interchangeable(fx, oracle_mod, monkeypatch, n, resident_from):
    """The default schedule runs the first pass in the z-marching kernel and every later pass in the brick-resident one;
    either kernel can run any pass (FXB_RESIDENT_FROM): all of them resident, the first three marching, all marching."""
    monkeypatch.setenv("FXB_RESIDENT_FROM", resident_from)
    f, o = make_pair(fx, oracle_mod, n)
    monkeypatch.delenv("FXB_RESIDENT_FROM")
    assert f.stats().jacobi_fused == 1
    inject(fx, oracle_mod, f, o, n, seed=5)
    dt = fx.dt_for_grid(*n)
    for _ in range(4):
        f.step(dt); o.step(dt)
        assert f.stats().s_exec == o.s_exec
    compare(fx, oracle_mod, f, o, TOL_1STEP, exact=True)


@pytest.mark.parametrize("tile", ["64", "128"])
@pytest.mark.parametrize("n", [(64, 64, 64), (136, 136, 50), (248, 248, 36), (40, 40, 7), (128, 128, 3),
                               (150, 150, 20), (61, 61, 11), (118, 118, 9), (59, 59, 24)])
def test_default_schedule_on_both_tile_widths(fx, oracle_mod, monkeypatch, n, tile):
    """The default schedule (T = 2; bricks that froze in the first pass are copied only next to active bricks; the final
    pressure settles in the first pass's output buffer) with the tile width forced to 64 and to 128 cells (normally
    chosen per grid).  Widths that are not a multiple of 8 (or of 4: the grid's x face then cuts through a quad) run on
    pitched pressure rows (common.cuh Domain::pitch) — e.g. the 150-wide grid of the reference's Bin/FluidGI.bat."""
    monkeypatch.setenv("FXB_TILE", tile)
    f, o = make_pair(fx, oracle_mod, n)
    monkeypatch.delenv("FXB_TILE")
    st = f.stats()
    assert st.jacobi_fused == 1 and st.fuse_t == 2 and st.brick_cells == (56 * 28 if tile == "64" else 120 * 12) * min(8, n[2])
    inject(fx, oracle_mod, f, o, n, seed=33)
    dt = fx.dt_for_grid(*n)
    for _ in range(4):
        f.step(dt); o.step(dt)
        assert f.stats().s_exec == o.s_exec
        assert f.stats().jacobi_passes == expected_passes(o.s_exec, f.stats())
        compare(fx, oracle_mod, f, o, TOL_1STEP, exact=True)
    # cells still active after sweep k + 1 (GPU) = cells entering sweep k + 1 (oracle)
    s = o.s_exec
    assert np.array_equal(f.freeze_histogram(64)[:s - 1].astype(np.int64), o.active_hist()[1:s])


@pytest.mark.parametrize("early,iters", [(False, 64), (False, 7), (True, 5), (True, 1), (True, 64), (True, 3)])
def test_default_schedule_partial_passes(fx, oracle_mod, early, iters):
    n = (72, 72, 40)
    f, o = make_pair(fx, oracle_mod, n, early_exit=early, iters=iters)
    inject(fx, oracle_mod, f, o, n, seed=4)
    dt = fx.dt_for_grid(*n)
    for _ in range(2):
        f.step(dt); o.step(dt)
    assert f.stats().s_exec == o.s_exec
    compare(fx, oracle_mod, f, o, TOL_1STEP, exact=True)


@pytest.mark.parametrize("early,iters", [(False, 64), (False, 7), (True, 5), (True, 1), (True, 64)])
def test_fused_jacobi_partial_last_pass_and_no_early_exit(fx, oracle_mod, early, iters):
    n = (72, 72, 40)
    f, o = make_pair(fx, oracle_mod, n, early_exit=early, iters=iters, fuse_t=4)
    inject(fx, oracle_mod, f, o, n, seed=4)
    dt = fx.dt_for_grid(*n)
    for _ in range(2):
        f.step(dt); o.step(dt)
    st = f.stats()
    assert st.s_exec == o.s_exec and st.jacobi_passes == -(-o.s_exec // 4)
    compare(fx, oracle_mod, f, o, TOL_1STEP, exact=True)


@pytest.mark.parametrize("n,fused", [((30, 30, 30), 1), ((6, 6, 6), 0)])
def test_any_width_and_the_per_sweep_kernels_below_eight(fx, oracle_mod, n, fused):
    """A width that is not a multiple of 8 (like the 150^3 of Bin/FluidGI.bat) takes the tuned pressure solve on pitched
    rows, with the per-voxel divergence / gradient kernels around it; grids narrower than 8 cells use the per-sweep
    kernels (still CUDA, one sweep per launch)."""
    f, o = make_pair(fx, oracle_mod, n)
    assert f.stats().jacobi_fused == fused
    dt = fx.dt_for_grid(*n)
    for _ in range(3):
        f.step(dt); o.step(dt)
        assert f.stats().s_exec == o.s_exec
    compare(fx, oracle_mod, f, o, TOL_1STEP, exact=True)
    q = f.get_field(fx.FIELD_PRESSURE)
    f.set_field(fx.FIELD_PRESSURE, q)  # pitched rows <-> the dense host layout, both directions
    assert np.array_equal(f.get_field(fx.FIELD_PRESSURE), q)


def test_graph_and_stream_launch_paths_agree(fx, oracle_mod):
    import torch
    n = (48, 48, 48)
    a = fx.Fluid(); b = fx.Fluid(); c = fx.Fluid()
    assert a.Init(gridSize=n, use_graph=True) and b.Init(gridSize=n, use_graph=False)
    assert c.Init(gridSize=n, kernel_path=1)
    dt = fx.dt_for_grid(*n)
    stream = torch.cuda.Stream()
    for _ in range(15):
        a.UpdateFrame(dt); a.Simulate(stream.cuda_stream)
        b.step(dt)
        c.step(dt)
    stream.synchronize()
    for fld in (fx.FIELD_VELOCITY, fx.FIELD_COLOR, fx.FIELD_PRESSURE):
        assert np.array_equal(a.get_field(fld), b.get_field(fld))
        assert np.array_equal(a.get_field(fld), c.get_field(fld))
    assert a.stats().s_exec == b.stats().s_exec == c.stats().s_exec


def test_field_io_and_errors(fx):
    import ctypes as C
    n = (16, 16, 16)
    f = fx.Fluid()
    assert f.Init(gridSize=n)
    vel, col, p = smooth_state(*n)
    f.set_field(fx.FIELD_VELOCITY, vel); f.set_field(fx.FIELD_COLOR, col); f.set_field(fx.FIELD_PRESSURE, p)
    assert np.array_equal(f.get_field(fx.FIELD_VELOCITY), vel)
    assert np.array_equal(f.get_field(fx.FIELD_COLOR), col)
    assert np.array_equal(f.get_field(fx.FIELD_PRESSURE), p)
    assert f.slab == (0, 16)
    buf = np.zeros(10, np.float32)
    rc = fx.lib().fxb_get_field(f._h, fx.FIELD_PRESSURE, buf.ctypes.data_as(C.c_void_p), buf.nbytes)
    assert rc == -4 and b"size" in fx.lib().fxb_last_error()
    assert fx.lib().fxb_get_field(f._h, 99, buf.ctypes.data_as(C.c_void_p), buf.nbytes) == -1
    g = fx.Fluid()
    assert g.Init(gridSize=(32, 16, 16)) is False and "nx must equal ny" in g.last_error


def test_full_size_properties_256_cubed(fx):
    """BASELINE.json configs[2] size: size-independent properties (the oracle is too slow to run here
    for many steps): determinism, colour bounds, zero far field, projection reduces divergence."""
    n = (256, 256, 256)
    dt = fx.dt_for_grid(*n)
    runs = []
    for _ in range(2):
        f = fx.Fluid()
        assert f.Init(gridSize=n), f.last_error
        for _ in range(30):
            f.step(dt)
        st = f.stats()
        runs.append((st.s_exec, f.get_field(fx.FIELD_COLOR), f.get_field(fx.FIELD_PRESSURE)))
        if len(runs) == 2:
            v1 = f.get_field(fx.FIELD_VELOCITY_ADVECTED).astype(np.float32)
            v0 = f.get_field(fx.FIELD_VELOCITY).astype(np.float32)
        f.close()
    assert runs[0][0] == runs[1][0] and 1 <= runs[0][0] <= 64
    assert np.array_equal(runs[0][1], runs[1][1]) and np.array_equal(runs[0][2], runs[1][2])
    c = runs[0][1].astype(np.float32)
    assert c.min() >= 0 and c.max() <= 1 and c[..., 3].max() > 0.5
    assert (c[:, 200:] == 0).all()  # the plume has not reached the top yet

    def div(v):
        return (v[1:-1, 1:-1, 2:, 0] - v[1:-1, 1:-1, :-2, 0] + v[1:-1, 2:, 1:-1, 1] - v[1:-1, :-2, 1:-1, 1]
                + v[2:, 1:-1, 1:-1, 2] - v[:-2, 1:-1, 1:-1, 2])
    d1, d0 = np.abs(div(v1)), np.abs(div(v0))
    strong = d1 > 0.25  # where the advected field is clearly divergent the projection must reduce it
    assert strong.sum() > 100 and d0[strong].sum() < 0.7 * d1[strong].sum()
