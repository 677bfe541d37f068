"""The light-map kernel's per-voxel body (fluidx12_b200/csrc/lightmap_body.cuh), run on the CPU
(tests/emu/lightmap_emu.cpp compiles the same statements with g++): bit for bit against the golden vectors made from
the reference's compiled CSRayMarchL and against the oracle on a simulated plume and on edge cases."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from tests.test_lightmap import CASES, GOLDEN, case_inputs, light_constants, oracle_params

_HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu")


@pytest.fixture(scope="module")
def emu():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liblightmap_emu.so"])
    lib = C.CDLL(os.path.join(_HERE, "liblightmap_emu.so"))
    lib.lightmap_emu_run.restype = None
    return lib


def run_body(emu, col, params):
    nz, ny, nx, _ = col.shape
    col = np.ascontiguousarray(col, np.float16)
    scratch = np.empty((nz, ny, nx), np.uint16)
    out = np.empty((nz, ny, nx), np.uint32)
    emu.lightmap_emu_run(nx, ny, nz, col.ctypes.data_as(C.c_void_p), C.byref(params),
                         scratch.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
    return out


@pytest.mark.parametrize("name", sorted(CASES))
def test_body_reproduces_the_interpreted_bytecode(emu, name):
    golden = np.load(GOLDEN)
    col, plain = case_inputs(golden, name)
    got = run_body(emu, col, oracle_params(plain))
    want = golden[name + "/light_map"]
    assert np.array_equal(got, want), (name, int((got != want).sum()))


@pytest.mark.parametrize("probes", [0, 1])
def test_body_matches_the_oracle_on_a_simulated_plume(emu, probes):
    n = (32, 32, 24)
    o = oracle.FluidOracle(*n)
    dt = oracle.dt_for_grid(*n)
    for _ in range(40):
        o.step(dt)
    col = o.get_field(oracle.FIELD_COLOR)
    assert (col[..., 3].astype(np.float32) >= 0.01).sum() > 500
    _, plain = light_constants(64, probes, (75.0, 75.0, -75.0), 3)
    p = oracle_params(plain)
    want = oracle.light_map(col, p)
    assert np.array_equal(run_body(emu, col, p), want)
    assert len(np.unique(want)) > 100


def test_body_matches_the_oracle_on_edge_cases(emu):
    r = np.random.default_rng(9)
    # ragged extents, density everywhere (every voxel marches), one sample per ray, light along an axis
    for grid, ns, lp in (((5, 7, 3), 1, (0.0, 0.0, 50.0)), ((9, 4, 6), 200, (-30.0, 1.0, 2.0)), ((4, 4, 4), 64, (1e-3, 0.0, 0.0))):
        nx, ny, nz = grid
        col = np.zeros((nz, ny, nx, 4), np.float16)
        col[..., 3] = r.random((nz, ny, nx)) * 0.3
        col[0, 0, 0, 3] = 0.0
        _, plain = light_constants(ns, 1, lp, 4)
        p = oracle_params(plain)
        assert np.array_equal(run_body(emu, col, p), oracle.light_map(col, p)), grid
    # density above 1 (negative GetStep factor) and a uniform block (zero gradient: the position is the AO direction)
    col = np.zeros((8, 8, 8, 4), np.float16)
    col[..., 3] = 0.25
    col[2:5, 2:5, 2:5, 3] = 1.75
    _, plain = light_constants(32, 1, (10.0, 20.0, 30.0), 5)
    p = oracle_params(plain)
    assert np.array_equal(run_body(emu, col, p), oracle.light_map(col, p))
