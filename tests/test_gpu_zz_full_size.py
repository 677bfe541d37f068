"""BASELINE.json configs[3] size, 512^3 — the workload bench.py times at N = 1 — where the oracle is far too slow to
step: size-independent properties of the CUDA path.  Two independent sets of kernels (the tuned path: fused TMA-staged
Jacobi passes, quad divergence / gradient; and the cross-check path: one simple kernel per logical pass, one sweep per
launch), each bit-exact against the oracle at the sizes the oracle can run, must agree bit for bit at full size,
including the number of executed sweeps; plus determinism-independent sanity of the fields."""
import numpy as np
import pytest

import fluidx12_b200 as fx

pytestmark = pytest.mark.gpu


def test_tuned_and_cross_check_kernel_paths_agree_at_512_cubed():
    n = (512, 512, 512)
    dt = fx.dt_for_grid(*n)
    a, b = fx.Fluid(), fx.Fluid()
    assert a.Init(gridSize=n, kernel_path=0), a.last_error
    assert b.Init(gridSize=n, kernel_path=1), b.last_error
    assert a.stats().jacobi_fused == 1 and b.stats().jacobi_fused == 0
    for k in range(12):
        step_dt = 0.0 if k == 7 else dt          # one paused frame
        a.step(step_dt)
        b.step(step_dt)
        assert a.stats().s_exec == b.stats().s_exec, k
    assert 1 <= a.stats().s_exec <= 64
    for fld in (fx.FIELD_PRESSURE, fx.FIELD_COLOR, fx.FIELD_VELOCITY):
        x, y = a.get_field(fld), b.get_field(fld)
        if fld == fx.FIELD_VELOCITY:
            x, y = x[..., :3], y[..., :3]
        assert np.array_equal(x, y), fld
        if fld == fx.FIELD_COLOR:
            c = x[..., 3].astype(np.float32)
            assert c.min() >= 0 and c.max() <= 1 and c.max() > 0.05    # smoke was emitted, colour stays in [0, 1]
            assert (x[:, 400:] == 0).all()                               # and has not reached the top of the grid
        if fld == fx.FIELD_PRESSURE:
            assert np.isfinite(x).all() and np.abs(x).max() > 0
        del x, y
    a.close()
    b.close()
