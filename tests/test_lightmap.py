"""Light-map pass (CSRayMarchL.hlsl): the oracle's restatement against golden vectors produced by
executing the reference's own Bin/CSRayMarchL.cso (tests/golden/make_lightmap_golden.py), the R11G11B10_FLOAT
packing, and known answers that need no bytecode.  The CUDA path's tests are in tests/test_zzz_gpu_lightmap.py."""
import hashlib
import os
import sys

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_lightmap_golden import CASES, colour_field, light_constants  # noqa: E402

GOLDEN = os.path.join(HERE, "golden", "lightmap_golden.npz")


def oracle_params(plain) -> oracle.LightParams:
    p = oracle.LightParams()
    p.light_pt[:] = plain["light_pt"].tolist()
    p.light_color[:] = plain["light_color"].tolist()
    p.ambient[:] = plain["ambient"].tolist()
    p.world_i[:] = plain["world_i"].reshape(-1).tolist()
    p.world[:] = plain["world"].reshape(-1).tolist()
    p.num_samples, p.has_light_probes = plain["num_samples"], plain["has_light_probes"]
    for i in range(9):
        p.sh[i][:] = plain["sh"][i].tolist()
    return p


def case_inputs(golden, name):
    grid, seed, ns, probes, lp = CASES[name]
    col = colour_field(grid, seed)
    _, plain = light_constants(ns, probes, lp, seed)
    digest = hashlib.sha256(col.tobytes() + plain["sh"].tobytes()).digest()
    assert bytes(golden[name + "/input_sha256"]) == digest, "the seeded inputs are not the ones the vectors were made from"
    return col, plain


def unpack(words):
    """R11G11B10_FLOAT words -> fp64 rgb (exact)."""
    w = np.asarray(words, np.uint32)
    out = []
    for mb, sh in ((6, 0), (6, 11), (5, 22)):
        f = (w >> np.uint32(sh)) & np.uint32((1 << (mb + 5)) - 1)
        e, m = (f >> np.uint32(mb)).astype(np.int64), (f & np.uint32((1 << mb) - 1)).astype(np.float64)
        v = np.where(e == 0, m * 2.0 ** (-14 - mb), (1 + m / (1 << mb)) * 2.0 ** (e - 15.0))
        out.append(np.where(e == 31, np.where(m == 0, np.inf, np.nan), v))
    return np.stack(out, -1)


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_the_interpreted_bytecode(golden, name):
    col, plain = case_inputs(golden, name)
    got = oracle.light_map(col, oracle_params(plain))
    want = golden[name + "/light_map"]
    assert got.shape == want.shape and np.array_equal(got, want), (name, int((got != want).sum()))
    assert len(np.unique(want)) > 100  # lit and shadowed voxels, many distinct values


def test_golden_file_names_the_blob_it_was_made_from(golden):
    ref = "/root/reference/Bin/CSRayMarchL.cso"
    if not os.path.exists(ref):
        pytest.skip("the reference tree is only present in the build container")
    assert bytes(golden["blob_sha256"]) == hashlib.sha256(open(ref, "rb").read()).digest()


def test_r11g11b10_packing_known_answers():
    P = oracle.pack_r11g11b10
    assert P(0.0, 0.0, 0.0) == 0 and P(-1.0, -0.0, -np.inf) == 0
    assert P(1.0, 1.0, 1.0) == (15 << 6) | (15 << 17) | (15 << 27)
    assert P(1.5, 0.0, 0.0) == (15 << 6) | 32 and P(0.0, 0.0, 1.5) == ((15 << 5) | 16) << 22
    # truncation toward zero: the largest fp32 below 2.0 stays in the binade of 1
    below2 = float(np.nextafter(np.float32(2.0), np.float32(0.0)))
    assert P(below2, 0.0, 0.0) == (15 << 6) | 63 and P(0.0, 0.0, below2) == ((15 << 5) | 31) << 22
    assert P(65024.0, 0.0, 0.0) == (30 << 6) | 63 and P(1e9, 0.0, 0.0) == (30 << 6) | 63       # finite overflow clamps
    assert P(np.inf, 0.0, 0.0) == 31 << 6 and (P(np.nan, 0.0, 0.0) >> 6) & 31 == 31 and P(np.nan, 0.0, 0.0) & 63
    assert P(2.0 ** -14, 0.0, 0.0) == 1 << 6 and P(2.0 ** -15, 0.0, 0.0) == 32 and P(2.0 ** -20, 0.0, 0.0) == 1
    assert P(2.0 ** -21, 0.0, 0.0) == 0
    r = np.random.default_rng(0)
    v = (r.random((2000, 3)) * np.array([4.0, 70000.0, 1e-4])).astype(np.float32)
    words = np.array([P(*map(float, row)) for row in v], np.uint32)
    back = unpack(words)
    lim = np.minimum(v.astype(np.float64), 65024.0)
    assert np.all(back <= lim) and np.all(back >= 0)
    ulp = np.maximum(back, 2.0 ** -14) * np.array([2.0 ** -6, 2.0 ** -6, 2.0 ** -5])
    assert np.all(lim - back < ulp * 1.0000001)
    import dxbc_interp as D
    assert np.array_equal(D.pack_r11g11b10(v), words)
    special = np.array([[np.inf, -np.inf, np.nan], [-0.0, 1e9, 2.0 ** -21], [6e-8, 65024.0, 64512.0]], np.float32)
    assert np.array_equal(D.pack_r11g11b10(special), np.array([P(*map(float, row)) for row in special], np.uint32))


def test_known_answers_without_bytecode():
    nx = ny = nz = 8
    _, plain = light_constants(16, 0, (0.0, 100.0, 0.0), 1)
    p = oracle_params(plain)
    lc = plain["light_color"][3] * plain["light_color"][:3]
    amb = plain["ambient"][3] * plain["ambient"][:3]
    # empty volume: every voxel is fully lit, light colour x 1 + ambient
    col = np.zeros((nz, ny, nx, 4), np.float16)
    out = oracle.light_map(col, p)
    full = oracle.pack_r11g11b10(*[float(np.float32(np.float32(1.0) * lc[c] + amb[c])) for c in range(3)])
    assert np.all(out == full)
    # a dense volume lit from +y through an axis-aligned transform: the light reaching a voxel never decreases
    # towards the light, the bottom of a column is in shadow (ambient only), the top is lit
    col[..., 3] = 1.0
    p.world_i[:] = [0.1, 0, 0, 0, 0, 0.1, 0, 0, 0, 0, 0.1, 0]
    p.world[:] = [10.0, 0, 0, 0, 0, 10.0, 0, 0, 0, 0, 10.0, 0]
    out = unpack(oracle.light_map(col, p))
    red = out[4, :, 4, 0]
    assert np.all(np.diff(red) >= 0) and red[0] < red[-1]
    assert abs(red[0] - amb[0]) < 0.1 and red[-1] > amb[0] + 0.1
    # light probes on, directional light off, only the constant SH term: irradiance = 0.886227 * sh0 where the density
    # reaches the 0.01 threshold, and the voxel stays black elsewhere (ao * irradiance with irradiance never evaluated)
    _, plain = light_constants(16, 1, (0.0, 100.0, 0.0), 1)
    plain["sh"][:] = 0
    plain["sh"][0] = [1.0, 2.0, 3.0]
    plain["light_color"][3] = 0.0
    p = oracle_params(plain)
    col[...] = 0
    col[2:6, 2:6, 2:6, 3] = 0.015
    out = unpack(oracle.light_map(col, p))
    assert np.all(out[0, 0, 0] == 0)
    assert np.allclose(out[3, 3, 3], 0.8862269520759583 * np.array([1.0, 2.0, 3.0]), rtol=0.05)
