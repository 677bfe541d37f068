"""The CUDA light-map pass (fxb_light_map, lightmap.cu) through the C ABI: bit for bit against the
golden vectors made from the reference's compiled CSRayMarchL and against the oracle on a simulated plume.  In a file
that sorts late: the pass follows the simulation step and must not stop that step's tests under -x."""
import ctypes as C

import numpy as np
import pytest

import fluidx12_b200 as fx
import oracle
from tests.test_lightmap import CASES, GOLDEN, case_inputs, light_constants, oracle_params

pytestmark = pytest.mark.gpu


def fx_params(plain) -> fx.FxbLightParams:
    p = fx.FxbLightParams()
    C.memmove(C.byref(p), C.byref(oracle_params(plain)), C.sizeof(p))  # the two structures are the same 256 bytes
    return p


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_light_map_reproduces_the_interpreted_bytecode(name):
    golden = np.load(GOLDEN)
    col, plain = case_inputs(golden, name)
    grid = CASES[name][0]
    f = fx.Fluid()
    assert f.Init(gridSize=grid, kernel_path=1), f.last_error   # no step is taken here: the simulation kernels are irrelevant
    f.set_field(fx.FIELD_COLOR, col)
    f.RayMarchL(fx_params(plain))
    got = f.get_light_map()
    want = golden[name + "/light_map"]
    assert np.array_equal(got, want), (name, int((got != want).sum()))
    f.close()


@pytest.mark.parametrize("probes", [0, 1])
def test_light_map_of_a_simulated_plume_matches_the_oracle(probes):
    n = (64, 64, 48)
    f = fx.Fluid()
    assert f.Init(gridSize=n), f.last_error
    dt = fx.dt_for_grid(*n)
    for _ in range(60):
        f.step(dt)
    _, plain = light_constants(64, probes, (75.0, 75.0, -75.0), 3)
    f.RayMarchL(fx_params(plain))      # same (default) stream as the steps: sees the last step's colour field
    got = f.get_light_map()
    col = f.get_field(fx.FIELD_COLOR)
    assert (col[..., 3].astype(np.float32) >= 0.01).sum() > 1000
    want = oracle.light_map(col, oracle_params(plain))
    assert np.array_equal(got, want), int((got != want).sum())
    # A paused frame does not flip the parity, but its advection still rewrites m_colors[m_frameParity] from the other
    # buffer (Fluid.cpp:345, 362, 372): the pass must read that buffer's NEW contents, whatever they are.
    parity = f.stats().frame_parity
    f.step(0.0)
    assert f.stats().frame_parity == parity
    f.RayMarchL(fx_params(plain))
    col2 = f.get_field(fx.FIELD_COLOR)
    assert np.array_equal(f.get_light_map(), oracle.light_map(col2, oracle_params(plain)))
    f.close()


def test_reference_defaults_ragged_grid_and_errors():
    n = (40, 40, 7)
    f = fx.Fluid()
    assert f.Init(gridSize=n), f.last_error
    with pytest.raises(fx.FluidError):
        f.get_light_map()               # the pass has not run yet
    r = np.random.default_rng(2)
    col = np.zeros((7, 40, 40, 4), np.float16)
    col[..., 3] = (r.random((7, 40, 40)) < 0.3) * r.random((7, 40, 40))
    f.set_field(fx.FIELD_COLOR, col)
    p = fx.FxbLightParams.reference_defaults()
    f.RayMarchL(p)
    q = oracle.LightParams()
    C.memmove(C.byref(q), C.byref(p), C.sizeof(p))
    assert np.array_equal(f.get_light_map(), oracle.light_map(col, q))
    f.close()
    g = fx.Fluid()
    assert g.Init(gridSize=(64, 64, 1)), g.last_error
    with pytest.raises(fx.FluidError):
        g.RayMarchL()                   # 2D grids have no light map (Fluid.cpp:296)
    g.close()
