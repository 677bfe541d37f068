"""GPU parity of the dynamic pressure-solve schedule (FXB_TAIL=1: bulk passes + tail launches, jacobi_tail.cu).

The tail kernel's body is pinned on the CPU by tests/test_tail_emu.py; here the CUDA build of the same body, the
device-side hand-over between the two kernels and the host schedule are checked against the oracle, bit for bit,
through the C ABI.  (The file sorts last on purpose: the schedule is opt-in, its tests must not mask the default path's.)"""
import os
from contextlib import contextmanager

import numpy as np
import pytest

from tests.util import smooth_state

pytestmark = pytest.mark.gpu


@contextmanager
def tail_env(**kw):
    env = {"FXB_TAIL": "1"}
    env.update({k: str(v) for k, v in kw.items()})
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    try:
        yield
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def run_pair(oracle_mod, n, steps, env, inject_seed=None, pause_at=None, **init_kw):
    import fluidx12_b200 as fx
    with tail_env(**env):
        f = fx.Fluid()
        assert f.Init(gridSize=n, **init_kw), f.last_error
    o = oracle_mod.FluidOracle(*n, early_exit=init_kw.get("early_exit", True), iters=init_kw.get("jacobi_iters", 64))
    assert f.tail_stats()["enabled"]
    if inject_seed is not None:
        vel, col, p = smooth_state(*n, seed=inject_seed, umax=1.5)
        for gf, of, a in ((fx.FIELD_VELOCITY, oracle_mod.FIELD_VEL, vel), (fx.FIELD_COLOR, oracle_mod.FIELD_COLOR, col),
                          (fx.FIELD_PRESSURE, oracle_mod.FIELD_PRESSURE, p)):
            f.set_field(gf, a)
            o.set_field(of, a)
    dt = fx.dt_for_grid(*n)
    launches = 0
    for k in range(steps):
        step_dt = 0.0 if k == pause_at else dt
        f.step(step_dt)
        o.step(step_dt)
        if step_dt > 0.0:
            f.sync()
            st = f.stats()
            assert st.s_exec == o.s_exec, (k, st.s_exec, o.s_exec)
            # GPU: cells still active after sweep k+1; oracle: cells active entering sweep k
            hist = f.freeze_histogram(o.iters)
            assert np.array_equal(hist[:o.s_exec - 1].astype(np.int64), o.active_hist()[1:o.s_exec]), k
            launches += f.tail_stats()["tail_launches_last_step"]
    for name, gf, of in (("velocity", fx.FIELD_VELOCITY, oracle_mod.FIELD_VEL), ("colour", fx.FIELD_COLOR, oracle_mod.FIELD_COLOR),
                         ("pressure", fx.FIELD_PRESSURE, oracle_mod.FIELD_PRESSURE)):
        a, b = f.get_field(gf), o.get_field(of)
        if name == "velocity":
            a, b = a[..., :3], b[..., :3]
        assert np.array_equal(a, b), (name, int((a != b).sum()))
    ts = f.tail_stats()
    f.close()
    return launches, ts


@pytest.mark.parametrize("n", [(64, 64, 64), (136, 136, 24), (128, 128, 40)])
def test_tail_schedule_matches_oracle(oracle_mod, n):
    launches, ts = run_pair(oracle_mod, n, 12, {})
    assert launches > 0 and ts["tail_bricks"] > 0


@pytest.mark.skipif(os.environ.get("FXB_TEST_EXPERIMENTAL") != "1",
                    reason="cp.async staging has not run on a GPU yet: FXB_TEST_EXPERIMENTAL=1 enables the test")
def test_tail_window_staged_with_cp_async(oracle_mod):
    """FXB_TAIL_CPASYNC=1: the sparse path stages its window with cp.async instead of through registers."""
    launches, _ = run_pair(oracle_mod, (128, 128, 40), 8, {"FXB_TAIL_CPASYNC": 1}, inject_seed=15)
    assert launches > 0


@pytest.mark.skipif(os.environ.get("FXB_TEST_EXPERIMENTAL") != "1",
                    reason="TMA staging of the tail window has not run on a GPU yet: FXB_TEST_EXPERIMENTAL=1 enables the test")
@pytest.mark.parametrize("n", [(128, 128, 40), (136, 136, 24)])
def test_tail_window_staged_with_tma(oracle_mod, n):
    """FXB_TAIL_CPASYNC=2: one cp.async.bulk.tensor.3d per window (zero fill outside the array by the copy engine)."""
    launches, _ = run_pair(oracle_mod, n, 8, {"FXB_TAIL_CPASYNC": 2}, inject_seed=16)
    assert launches > 0


def test_tail_dense_path_only(oracle_mod):
    """FXB_TAIL_SPARSE_CAP=0: every window that holds an active cell takes the register-column path."""
    _, ts = run_pair(oracle_mod, (64, 64, 40), 6, {"FXB_TAIL_SPARSE_CAP": 0})
    assert ts["tail_subblocks_dense"] == ts["tail_subblocks_relaxed"] > 0


@pytest.mark.skipif(os.environ.get("FXB_TEST_EXPERIMENTAL") != "1",
                    reason="the second dense path has not run on a GPU yet: FXB_TEST_EXPERIMENTAL=1 enables the test")
def test_tail_second_dense_path(oracle_mod):
    """FXB_TAIL_DENSE=2 with FXB_TAIL_SPARSE_CAP=0: every active window takes the two-phase all-quads path."""
    _, ts = run_pair(oracle_mod, (64, 64, 40), 6, {"FXB_TAIL_SPARSE_CAP": 0, "FXB_TAIL_DENSE": 2})
    assert ts["tail_subblocks_dense"] == ts["tail_subblocks_relaxed"] > 0


@pytest.mark.skipif(os.environ.get("FXB_TEST_EXPERIMENTAL") != "1",
                    reason="the block-resident pass-0 kernel has not run on a GPU yet: FXB_TEST_EXPERIMENTAL=1 enables the test")
@pytest.mark.parametrize("n", [(64, 64, 64), (136, 136, 24)])
def test_pass0_by_the_block_resident_kernel(oracle_mod, n):
    """FXB_PASS0=2: pass 0 of every frame runs on jacobi_tail_kernel<TailShape<2, 10, 12, 8>> instead of the bulk kernel."""
    launches, _ = run_pair(oracle_mod, n, 8, {"FXB_PASS0": 2}, inject_seed=17)
    assert launches > 0


def test_tail_takes_over_right_after_pass_zero(oracle_mod):
    launches, _ = run_pair(oracle_mod, (64, 64, 64), 8, {"FXB_TAIL_MAINS": 1}, inject_seed=11)
    assert launches > 0


def test_tail_never_qualifies_until_forced(oracle_mod):
    """Threshold 0: the interleaved tail launches return at once, bulk passes 0..4 run, then forced tail launches."""
    run_pair(oracle_mod, (64, 64, 64), 8, {"FXB_TAIL_THRESHOLD": 0}, inject_seed=12)


def test_tail_small_grid_of_ctas_and_eager_launch(oracle_mod):
    run_pair(oracle_mod, (136, 136, 24), 6, {"FXB_TAIL_GRID": 3, "FXB_TAIL_MAINS": 3}, inject_seed=13, use_graph=False)


def test_tail_odd_iteration_count_no_early_exit_and_pause(oracle_mod):
    run_pair(oracle_mod, (64, 64, 24), 6, {}, inject_seed=14, pause_at=3, jacobi_iters=23, early_exit=False)


def test_tail_same_bits_as_default_schedule(oracle_mod):
    """The dynamic schedule against the static one (both CUDA) on a larger developed grid: identical fields."""
    import fluidx12_b200 as fx
    n = (256, 256, 64)
    a = fx.Fluid()
    assert a.Init(gridSize=n), a.last_error
    with tail_env():
        b = fx.Fluid()
        assert b.Init(gridSize=n), b.last_error
    dt = fx.dt_for_grid(*n)
    for _ in range(30):
        a.step(dt)
        b.step(dt)
    a.sync(); b.sync()
    assert a.stats().s_exec == b.stats().s_exec
    for fld in (fx.FIELD_VELOCITY, fx.FIELD_COLOR, fx.FIELD_PRESSURE):
        assert np.array_equal(a.get_field(fld), b.get_field(fld)), fld
    assert b.tail_stats()["tail_launches_last_step"] > 0
    a.close(); b.close()


@pytest.mark.skipif(os.environ.get("FXB_TEST_EXPERIMENTAL") != "1",
                    reason="the second advection kernel has not run on a GPU yet: FXB_TEST_EXPERIMENTAL=1 enables the test")
@pytest.mark.parametrize("n,mode", [((64, 64, 64), 0), ((50, 50, 30), 0), ((48, 48, 48), 1), ((128, 128, 40), 0)])
def test_second_advection_kernel(oracle_mod, n, mode):
    """FXB_ADVECT=2 (advect_body.cuh, emulated on the CPU by tests/test_advect_emu.py) against the oracle."""
    import fluidx12_b200 as fx
    old = os.environ.get("FXB_ADVECT")
    os.environ["FXB_ADVECT"] = "2"
    try:
        f = fx.Fluid()
        assert f.Init(gridSize=n, address_mode=mode), f.last_error
    finally:
        if old is None:
            os.environ.pop("FXB_ADVECT", None)
        else:
            os.environ["FXB_ADVECT"] = old
    o = oracle_mod.FluidOracle(*n, address_mode=mode)
    vel, col, p = smooth_state(*n, seed=21, umax=2.0)
    for gf, of, a in ((fx.FIELD_VELOCITY, oracle_mod.FIELD_VEL, vel), (fx.FIELD_COLOR, oracle_mod.FIELD_COLOR, col),
                      (fx.FIELD_PRESSURE, oracle_mod.FIELD_PRESSURE, p)):
        f.set_field(gf, a)
        o.set_field(of, a)
    dt = fx.dt_for_grid(*n)
    for _ in range(10):
        f.step(dt)
        o.step(dt)
    for name, gf, of in (("velocity", fx.FIELD_VELOCITY, oracle_mod.FIELD_VEL), ("colour", fx.FIELD_COLOR, oracle_mod.FIELD_COLOR),
                         ("pressure", fx.FIELD_PRESSURE, oracle_mod.FIELD_PRESSURE)):
        a, b = f.get_field(gf), o.get_field(of)
        if name == "velocity":
            a, b = a[..., :3], b[..., :3]
        assert np.array_equal(a, b), (name, int((a != b).sum()))
    f.close()
