"""The CUDA path against the golden vectors made from the reference's own compiled shaders (tests/golden/dxbc_golden.npz,
see tests/test_dxbc_golden.py for what they are).  Kept in a file that sorts late: these grids are smaller than any
other GPU test's, and a surprise here must not stop the rest of the suite under -x."""
import numpy as np
import pytest

from tests.test_dxbc_golden import CASES, GOLDEN, check_inputs, frames

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_path_reproduces_interpreted_dxbc(golden, name):
    import fluidx12_b200 as fx
    grid, seed, steps, clamp, pause = CASES[name]
    v0, c0, p0 = check_inputs(golden, name, grid, seed)
    f = fx.Fluid()
    assert f.Init(gridSize=grid, address_mode=fx.ADDRESS_CLAMP if clamp else fx.ADDRESS_MIRROR), f.last_error
    f.set_field(fx.FIELD_VELOCITY, v0)
    f.set_field(fx.FIELD_COLOR, c0)
    f.set_field(fx.FIELD_PRESSURE, p0)
    for k, dt in enumerate(frames(grid, steps, pause)):
        f.step(dt)
        f.sync()
        trips = int(golden[name + "/loop_trips"][k])
        assert f.stats().s_exec == (0 if dt == 0.0 else min(trips + 1, 64)), (k, f.stats().s_exec, trips)
    for key, fld in (("velocity", fx.FIELD_VELOCITY), ("velocity_advected", fx.FIELD_VELOCITY_ADVECTED),
                     ("colour", fx.FIELD_COLOR), ("pressure", fx.FIELD_PRESSURE)):
        want, got = golden[name + "/" + key], f.get_field(fld)
        if key.startswith("velocity"):
            want, got = want[..., :3], got[..., :3]
        assert np.array_equal(want.view(np.uint16 if want.dtype == np.float16 else np.uint32),
                              got.view(np.uint16 if got.dtype == np.float16 else np.uint32)), (name, key)
    f.close()


