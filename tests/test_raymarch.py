"""Cube-map ray march with the separate light pass (CSRayMarchV): the oracle's restatement against
golden vectors produced by executing the reference's own Bin/CSRayMarchL.cso + Bin/CSRayMarchV.cso
(tests/golden/make_raymarch_golden.py), plus known answers.  CUDA path: tests/test_zzz_gpu_raymarch.py."""
import hashlib
import os
import sys

import numpy as np
import pytest

import oracle
from tests.test_lightmap import oracle_params

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_lightmap_golden import colour_field, light_constants  # noqa: E402
from make_raymarch_golden import CASES, view_constants, visibility_mask  # noqa: E402

GOLDEN = os.path.join(HERE, "golden", "raymarch_golden.npz")


def view_params(plain) -> oracle.ViewParams:
    p = oracle.ViewParams()
    p.eye_pt[:] = plain["eye_pt"].tolist()
    p.world_i[:] = plain["world_i"].reshape(-1).tolist()
    p.num_samples, p.visibility_mask, p.cube_size = plain["num_samples"], plain["visibility_mask"], plain["cube_size"]
    return p


def case_inputs(golden, name):
    """(colour, light constants, view constants) exactly as make_raymarch_golden.run_case builds them."""
    grid, seed, light_samples, probes, ray_samples, cube_size, eye, faces_off = CASES[name]
    col = colour_field(grid, seed)
    col[..., :3] = (col[..., :3].astype(np.float32) * 1.5).astype(np.float16)
    cbs_l, plain_l = light_constants(light_samples, probes, (75.0, 75.0, -75.0), seed)
    plain_l["light_color"][3] = 2.0
    plain_l["ambient"][3] = 0.5
    _, plain_v = view_constants(cbs_l, plain_l, ray_samples, cube_size, eye, faces_off)
    digest = hashlib.sha256(col.tobytes() + plain_l["sh"].tobytes()).digest()
    assert bytes(golden[name + "/input_sha256"]) == digest, "the seeded inputs are not the ones the vectors were made from"
    return col, plain_l, plain_v


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_the_interpreted_bytecode(golden, name):
    col, plain_l, plain_v = case_inputs(golden, name)
    lmap = oracle.light_map(col, oracle_params(plain_l))
    assert np.array_equal(lmap, golden[name + "/light_map"])
    got = oracle.ray_march_v(col, lmap, view_params(plain_v))
    want = golden[name + "/cube_map"]
    assert got.shape == want.shape and np.array_equal(got, want), (name, int((got != want).sum()))
    assert (want[..., 3] > 0).sum() > 50 and len(np.unique(want.reshape(-1, 4), axis=0)) > 50


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_reproduces_the_interpreted_non_separated_march(golden, name):
    """Bin/CSRayMarch.cso (Fluid::rayMarch, Fluid.cpp:825-855): light and occlusion rays cast at every view sample."""
    col, plain_l, plain_v = case_inputs(golden, name)
    got = oracle.ray_march(col, view_params(plain_v), oracle_params(plain_l))
    want = golden[name + "/cube_map_full"]
    assert np.array_equal(got, want), (name, int((got != want).sum()))
    # and it is the same picture as the light-map version up to that map's quantisation and interpolation
    assert np.abs(want.astype(int) - golden[name + "/cube_map"].astype(int)).max() < 48


def test_unpack_is_the_inverse_of_pack_on_every_word_class():
    r = np.random.default_rng(3)
    words = np.concatenate([r.integers(0, 1 << 32, 3000, dtype=np.uint64).astype(np.uint32),
                            np.array([0, 1, 1 << 6, 63, 0x7FF, 31 << 6, (31 << 6) | 1, 0xFFFFFFFF, 1 << 22, 31 << 27], np.uint32)])
    import dxbc_interp as D
    ref = D.unpack_r11g11b10(words)
    for w, want in zip(words.tolist(), ref):
        got = oracle.unpack_r11g11b10(w)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)) or (np.isnan(got) == np.isnan(want)).all(), hex(w)
        if np.isfinite(got).all():
            assert oracle.pack_r11g11b10(*map(float, got)) == w


def test_known_answers_without_bytecode():
    n = 8
    col = np.zeros((n, n, n, 4), np.float16)
    _, plain_l = light_constants(16, 0, (0.0, 100.0, 0.0), 1)
    plain_l["world_i"][:] = 0
    plain_l["world_i"][[0, 1, 2], [0, 1, 2]] = 0.1
    lmap = oracle.light_map(col, oracle_params(plain_l))
    eye = (0.0, 0.0, -30.0)
    plain_v = {"eye_pt": np.array(eye, np.float32), "world_i": plain_l["world_i"], "num_samples": 32,
               "visibility_mask": visibility_mask(plain_l["world_i"], eye), "cube_size": 8}
    # The cube map holds what is seen THROUGH the volume on the far side: a ray runs from the eye to its texel on a cube
    # face, so an eye in front of the -z face marches every face but -z itself (IsCubeFaceVisible, Fluid.cpp:41-46).
    # An empty volume scatters nothing: every marched texel is written as 0, previous contents survive elsewhere.
    assert plain_v["visibility_mask"] == 0b011111
    prev = np.full((6, 8, 8, 4), 7, np.uint8)
    out = oracle.ray_march_v(col, lmap, view_params(plain_v), cube=prev)
    assert np.all(out[5] == 7)                          # -z face culled: untouched
    assert np.all(out[4] == 0)                          # +z face: every ray crosses the box
    # a uniform medium lit from +y: nearly opaque at the centre, mirror-symmetric in x, brighter towards the light
    col[..., 3] = 0.5
    col[..., :3] = 0.25
    lmap = oracle.light_map(col, oracle_params(plain_l))
    out = oracle.ray_march_v(col, lmap, view_params(plain_v))
    centre = out[4, 3:5, 3:5]
    assert np.all(centre[..., 3] > 200) and np.array_equal(centre[:, 0], centre[:, 1])
    assert np.all(centre[0, 0, :3] > centre[1, 0, :3])   # texel row 3 is above row 4 (pos.y = -y, CSRayMarch.hlsl:46)
    # fewer samples -> less accumulated opacity
    plain_v["num_samples"] = 2
    out2 = oracle.ray_march_v(col, lmap, view_params(plain_v))
    assert out2[4, 3, 3, 3] < out[4, 3, 3, 3]


def test_visibility_mask_of_the_library_is_the_references_rule():
    """fxb_cube_visibility_mask (host code of the product, needs no GPU) = GenVisibilityMask (Fluid.cpp:51-63)."""
    import ctypes as C

    import fluidx12_b200 as fx
    L = fx.lib()
    r = np.random.default_rng(8)
    fp = C.POINTER(C.c_float)
    for _ in range(200):
        wi = (r.standard_normal((3, 4)) * 0.1).astype(np.float32)
        eye = (r.standard_normal(3) * 20).astype(np.float32)
        m = C.c_uint32()
        assert L.fxb_cube_visibility_mask(wi.ctypes.data_as(fp), eye.ctypes.data_as(fp), C.byref(m)) == 0
        assert m.value == visibility_mask(wi, eye), (wi, eye)
    assert L.fxb_cube_visibility_mask(None, None, None) == fx.binding.FXB_ERR_INVALID
    assert L.fxb_ray_march_v(None, None, None) == fx.binding.FXB_ERR_INVALID
    assert L.fxb_get_cube_map(None, None, 0) == fx.binding.FXB_ERR_INVALID


def test_cube_lod_estimate_follows_the_reference_host_code():
    """fxb_estimate_cube_lod = EstimateCubeMapLOD (Fluid.cpp:141-166) with its two helpers, restated here in numpy
    float32 with the same operation order: ideal ray-sample count from the longest projected cube edge, clamped by the
    user's maximum, and the cube-map mip derived from it."""
    import ctypes as C

    import fluidx12_b200 as fx
    L = fx.lib()
    F = np.float32

    def look_at_lh(eye, at, up):
        z = (at - eye) / np.linalg.norm(at - eye)
        x = np.cross(up, z)
        x /= np.linalg.norm(x)
        y = np.cross(z, x)
        m = np.eye(4)
        m[:3, 0], m[:3, 1], m[:3, 2] = x, y, z
        m[3, :3] = [-x @ eye, -y @ eye, -z @ eye]
        return m

    def perspective_lh(fovy, aspect, zn, zf):
        h = 1.0 / np.tan(fovy / 2)
        m = np.zeros((4, 4))
        m[0, 0], m[1, 1], m[2, 2], m[2, 3], m[3, 2] = h / aspect, h, zf / (zf - zn), 1.0, -zn * zf / (zf - zn)
        return m

    def restated(m, vw, vh, max_samples, mips, size0):
        m = m.astype(F).reshape(-1)
        corner = np.array([[1, 1, 1], [-1, 1, 1], [1, -1, 1], [-1, -1, 1], [-1, 1, -1], [1, 1, -1], [-1, -1, -1], [1, -1, -1]], F)
        px, py = [], []
        for x, y, z in corner:
            rx = F(F(F(x * m[0]) + F(y * m[4])) + F(z * m[8])) + m[12]
            ry = F(F(F(x * m[1]) + F(y * m[5])) + F(z * m[9])) + m[13]
            rw = F(F(F(x * m[3]) + F(y * m[7])) + F(z * m[11])) + m[15]
            px.append(F(F(F(F(rx / rw) * F(0.5)) + F(0.5)) * F(vw)))
            py.append(F(F(F(F(ry / rw) * F(-0.5)) + F(0.5)) * F(vh)))
        edges = [(0, 1), (3, 2), (1, 3), (2, 0), (4, 5), (7, 6), (5, 7), (6, 4), (1, 4), (6, 3), (5, 0), (2, 7)]
        longest = F(0)
        for a, b in edges:
            ex, ey = F(px[b] - px[a]), F(py[b] - py[a])
            longest = max(np.sqrt(F(F(ex * ex) + F(ey * ey))), longest)
        s = F(longest / F(2))
        amount = F(F(F(2) * s) / np.sqrt(F(3)))
        samples = min(int(np.ceil(amount)), max_samples)
        amount = min(amount, F(samples))
        s = F(F(amount / F(2)) * np.sqrt(F(3)))
        level = max(np.log2(F(F(size0) / s)), F(0))
        return samples, min(int(level), mips - 1)

    r = np.random.default_rng(12)
    seen = set()
    for _ in range(300):
        eye = r.standard_normal(3) * r.uniform(15, 120)
        view = look_at_lh(eye, r.standard_normal(3) * 2, np.array([0.0, 1.0, 0.0]))
        vw, vh = float(r.integers(320, 3840)), float(r.integers(240, 2160))
        proj = perspective_lh(np.pi / 4, vw / vh, 1.0, 1000.0)
        wvp = ((np.eye(4) * [10, 10, 10, 1]) @ view @ proj).astype(F)
        size0 = int(r.choice([64, 128, 256, 512]))
        want = restated(wvp, vw, vh, 192, 5, size0)
        a, b = C.c_uint32(), C.c_uint32()
        assert L.fxb_estimate_cube_lod(wvp.ctypes.data_as(C.POINTER(C.c_float)), vw, vh, 192, 5, size0, C.byref(a), C.byref(b)) == 0
        assert (a.value, b.value) == want, (eye, vw, vh, size0)
        seen.add(want)
    assert len({s for s, _ in seen}) > 20 and {lvl for _, lvl in seen} >= {0, 1, 2}     # the cases are not all alike
    assert L.fxb_estimate_cube_lod(None, 1.0, 1.0, 1, 1, 1, None, None) == fx.binding.FXB_ERR_INVALID
