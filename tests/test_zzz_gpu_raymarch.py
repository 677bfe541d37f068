"""The CUDA cube-map ray march (fxb_ray_march_v, raymarch.cu) through the C ABI: bit for bit against
the golden vectors made from the reference's compiled CSRayMarchL + CSRayMarchV and against the oracle on a simulated
plume.  Late-sorting file (see tests/test_zzz_gpu_lightmap.py)."""
import ctypes as C

import numpy as np
import pytest

import fluidx12_b200 as fx
import oracle
from tests.test_lightmap import light_constants, oracle_params
from tests.test_raymarch import CASES, GOLDEN, case_inputs, view_params, visibility_mask

pytestmark = pytest.mark.gpu


def as_fx(p, cls):
    q = cls()
    assert C.sizeof(q) == C.sizeof(p)
    C.memmove(C.byref(q), C.byref(p), C.sizeof(p))
    return q


@pytest.mark.parametrize("name", sorted(CASES))
def test_cuda_ray_march_reproduces_the_interpreted_bytecode(name):
    golden = np.load(GOLDEN)
    col, plain_l, plain_v = case_inputs(golden, name)
    f = fx.Fluid()
    assert f.Init(gridSize=CASES[name][0], kernel_path=1), f.last_error   # no step is taken: only the colour field is set
    f.set_field(fx.FIELD_COLOR, col)
    f.RayMarchL(as_fx(oracle_params(plain_l), fx.FxbLightParams))
    assert np.array_equal(f.get_light_map(), golden[name + "/light_map"])
    f.RayMarchV(as_fx(view_params(plain_v), fx.FxbViewParams))
    got, want = f.get_cube_map(), golden[name + "/cube_map"]
    assert np.array_equal(got, want), (name, int((got != want).sum()))
    f.close()
    # the non-separated march (CSRayMarch) on a fresh handle: no light map needed, its own golden cube map
    f = fx.Fluid()
    assert f.Init(gridSize=CASES[name][0], kernel_path=1), f.last_error   # no step is taken: only the colour field is set
    f.set_field(fx.FIELD_COLOR, col)
    f.RayMarch(as_fx(view_params(plain_v), fx.FxbViewParams), as_fx(oracle_params(plain_l), fx.FxbLightParams))
    got, want = f.get_cube_map(), golden[name + "/cube_map_full"]
    assert np.array_equal(got, want), (name, "non-separated", int((got != want).sum()))
    f.close()


def test_cube_map_of_a_simulated_plume_matches_the_oracle():
    n = (64, 64, 48)
    f = fx.Fluid()
    assert f.Init(gridSize=n), f.last_error
    with pytest.raises(fx.FluidError):
        f.RayMarchV(fx.FxbViewParams())          # the light map is an input: RayMarchL first
    dt = fx.dt_for_grid(*n)
    for _ in range(60):
        f.step(dt)
    _, plain_l = light_constants(64, 1, (75.0, 75.0, -75.0), 3)
    plain_l["light_color"][3] = 2.0
    pl = oracle_params(plain_l)
    f.RayMarchL(as_fx(pl, fx.FxbLightParams))
    col = f.get_field(fx.FIELD_COLOR)
    lmap = oracle.light_map(col, pl)
    assert np.array_equal(f.get_light_map(), lmap)
    for eye, size in (((14.0, 22.0, -31.0), 64), ((0.5, 1.0, -1.5), 32), ((-60.0, 5.0, 3.0), 16)):
        wi = plain_l["world_i"].copy()
        wi[:, 3] = [0.01, 0.02, -0.03]
        mask = C.c_uint32()
        fx.binding.check(fx.lib().fxb_cube_visibility_mask(wi.ctypes.data_as(C.POINTER(C.c_float)),
                                                            np.array(eye, np.float32).ctypes.data_as(C.POINTER(C.c_float)),
                                                            C.byref(mask)))
        assert mask.value == visibility_mask(wi, eye)
        plain_v = {"eye_pt": np.array(eye, np.float32), "world_i": wi, "num_samples": 192, "visibility_mask": mask.value,
                   "cube_size": size}
        pv = view_params(plain_v)
        f.RayMarchV(as_fx(pv, fx.FxbViewParams))
        got = f.get_cube_map()
        want = oracle.ray_march_v(col, lmap, pv)   # a new cube size starts from zeros on both sides
        assert np.array_equal(got, want), (eye, int((got != want).sum()))
        assert (want[..., 3] > 0).sum() > 20
    f.close()
