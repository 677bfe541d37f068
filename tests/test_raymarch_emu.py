"""The cube-map ray-march kernel's per-texel body (fluidx12_b200/csrc/raymarch_body.cuh), run on the CPU
(tests/emu/raymarch_emu.cpp compiles the same statements with g++): bit for bit against the golden vectors made from
the reference's compiled CSRayMarchL + CSRayMarchV and against the oracle on a simulated plume and edge cases."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from tests.test_lightmap import light_constants, oracle_params
from tests.test_raymarch import CASES, GOLDEN, case_inputs, view_params, visibility_mask

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", _HERE, "libraymarch_emu.so"])
    lib = C.CDLL(os.path.join(_HERE, "libraymarch_emu.so"))
    lib.raymarch_emu_run.restype = None
    return lib


def run_body(emu, col, lmap, params, cube=None):
    nz, ny, nx, _ = col.shape
    col = np.ascontiguousarray(col, np.float16)
    lmap = np.ascontiguousarray(lmap, np.uint32)
    s = int(params.cube_size)
    out = np.zeros((6, s, s, 4), np.uint8) if cube is None else cube.copy()
    emu.raymarch_emu_run(nx, ny, nz, col.ctypes.data_as(C.c_void_p), lmap.ctypes.data_as(C.c_void_p), C.byref(params),
                         out.ctypes.data_as(C.c_void_p))
    return out


@pytest.mark.parametrize("name", sorted(CASES))
def test_body_reproduces_the_interpreted_bytecode(emu, name):
    golden = np.load(GOLDEN)
    col, plain_l, plain_v = case_inputs(golden, name)
    got = run_body(emu, col, golden[name + "/light_map"], view_params(plain_v))
    want = golden[name + "/cube_map"]
    assert np.array_equal(got, want), (name, int((got != want).sum()))


def run_body_full(emu, col, view, light):
    nz, ny, nx, _ = col.shape
    col = np.ascontiguousarray(col, np.float16)
    dens = np.ascontiguousarray(col[..., 3]).view(np.uint16)
    s = int(view.cube_size)
    out = np.zeros((6, s, s, 4), np.uint8)
    emu.raymarch_emu_run_full(nx, ny, nz, col.ctypes.data_as(C.c_void_p), dens.ctypes.data_as(C.c_void_p), C.byref(view),
                              C.byref(light), out.ctypes.data_as(C.c_void_p))
    return out


@pytest.mark.parametrize("name", sorted(CASES))
def test_non_separated_body_reproduces_the_interpreted_bytecode(emu, name):
    golden = np.load(GOLDEN)
    col, plain_l, plain_v = case_inputs(golden, name)
    got = run_body_full(emu, col, view_params(plain_v), oracle_params(plain_l))
    want = golden[name + "/cube_map_full"]
    assert np.array_equal(got, want), (name, int((got != want).sum()))


@pytest.mark.parametrize("probes", [0, 1])
def test_non_separated_body_matches_the_oracle_on_a_simulated_plume(emu, probes):
    n = (24, 24, 16)
    o = oracle.FluidOracle(*n)
    dt = oracle.dt_for_grid(*n)
    for _ in range(30):
        o.step(dt)
    col = o.get_field(oracle.FIELD_COLOR)
    _, plain_l = light_constants(16, probes, (75.0, 75.0, -75.0), 3)
    plain_l["light_color"][3] = 2.0
    wi = plain_l["world_i"].copy()
    wi[:, 3] = [0.01, 0.02, -0.03]
    eye = (14.0, 22.0, -31.0)
    plain_v = {"eye_pt": np.array(eye, np.float32), "world_i": wi, "num_samples": 48,
               "visibility_mask": visibility_mask(wi, eye), "cube_size": 16}
    pv, pl = view_params(plain_v), oracle_params(plain_l)
    want = oracle.ray_march(col, pv, pl)
    assert np.array_equal(run_body_full(emu, col, pv, pl), want)
    assert (want[..., 3] > 0).sum() > 100


@pytest.mark.parametrize("eye", [(14.0, 22.0, -31.0), (0.5, 1.0, -1.5), (-60.0, 0.0, 0.0)])
def test_body_matches_the_oracle_on_a_simulated_plume(emu, eye):
    n = (32, 32, 24)
    o = oracle.FluidOracle(*n)
    dt = oracle.dt_for_grid(*n)
    for _ in range(40):
        o.step(dt)
    col = o.get_field(oracle.FIELD_COLOR)
    _, plain_l = light_constants(32, 1, (75.0, 75.0, -75.0), 3)
    plain_l["light_color"][3] = 2.0
    lmap = oracle.light_map(col, oracle_params(plain_l))
    wi = plain_l["world_i"].copy()
    wi[:, 3] = [0.01, 0.02, -0.03]
    plain_v = {"eye_pt": np.array(eye, np.float32), "world_i": wi, "num_samples": 96,
               "visibility_mask": visibility_mask(wi, eye), "cube_size": 24}
    p = view_params(plain_v)
    prev = np.full((6, 24, 24, 4), 9, np.uint8)
    want = oracle.ray_march_v(col, lmap, p, cube=prev)
    assert np.array_equal(run_body(emu, col, lmap, p, cube=prev), want)
    assert (want[..., 3] != 9).sum() > 500


def test_body_matches_the_oracle_on_edge_cases(emu):
    r = np.random.default_rng(4)
    col = (r.random((6, 5, 7, 4)) * np.array([1, 1, 1, 1.5])).astype(np.float16)    # ragged grid, density above 1
    lmap = r.integers(0, 1 << 32, (6, 5, 7), dtype=np.uint64).astype(np.uint32)     # any words: denormals, INF, NaN
    wi = np.zeros((3, 4), np.float32)
    wi[[0, 1, 2], [0, 1, 2]] = 0.1
    for eye, ns, s in (((0.0, 0.0, -30.0), 1, 8), ((10.0, 10.0, 10.0), 500, 16), ((0.0, 0.0, 0.0), 40, 8),
                       ((10.0, 0.0, 25.0), 33, 8)):
        plain_v = {"eye_pt": np.array(eye, np.float32), "world_i": wi, "num_samples": ns, "visibility_mask": 0b111111,
                   "cube_size": s}
        p = view_params(plain_v)
        assert np.array_equal(run_body(emu, col, lmap, p), oracle.ray_march_v(col, lmap, p)), (eye, ns)


def test_bodies_match_the_oracle_on_degenerate_constants(emu):
    """Zero samples (an infinite step), a light at the origin or at infinity (NaN directions), a NaN eye: whatever the
    reference's arithmetic makes of them, the kernel bodies and the oracle make the same of them."""
    lm_emu = C.CDLL(os.path.join(_HERE, "liblightmap_emu.so"))
    lm_emu.lightmap_emu_run.restype = None
    r = np.random.default_rng(1)
    col = (r.random((6, 6, 6, 4)) * np.array([1, 1, 1, 0.6])).astype(np.float16)
    for light_samples, light_pt in ((0, (75.0, 75.0, -75.0)), (8, (0.0, 0.0, 0.0)), (8, (np.inf, 0.0, 0.0))):
        _, plain_l = light_constants(light_samples, 1, light_pt, 2)
        pl = oracle_params(plain_l)
        scratch, got = np.empty((6, 6, 6), np.uint16), np.empty((6, 6, 6), np.uint32)
        lm_emu.lightmap_emu_run(6, 6, 6, col.ctypes.data_as(C.c_void_p), C.byref(pl), scratch.ctypes.data_as(C.c_void_p),
                                got.ctypes.data_as(C.c_void_p))
        lmap = oracle.light_map(col, pl)
        assert np.array_equal(got, lmap), (light_samples, light_pt)
        for eye, ray_samples in (((14.0, 22.0, -31.0), 0), ((0.0, 0.0, 0.0), 16), ((np.nan, 1.0, 1.0), 16)):
            pv = view_params({"eye_pt": np.array(eye, np.float32), "world_i": plain_l["world_i"], "num_samples": ray_samples,
                              "visibility_mask": 63, "cube_size": 8})
            assert np.array_equal(run_body(emu, col, lmap, pv), oracle.ray_march_v(col, lmap, pv)), (eye, ray_samples)
            assert np.array_equal(run_body_full(emu, col, pv, pl), oracle.ray_march(col, pv, pl)), (eye, ray_samples)
