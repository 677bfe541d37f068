"""Golden vectors of the cube-map ray march with the separate light pass from the reference's own
compiled shaders: Bin/CSRayMarchL.cso produces the light map, Bin/CSRayMarchV.cso marches the view rays — both executed
by tests/golden/dxbc_interp.py — on seeded colour fields.  Writes tests/golden/raymarch_golden.npz (the [6][S][S][4]
UNORM8 cube map per case).  Needs /root/reference (this container only)."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import dxbc_interp as D  # noqa: E402
from make_lightmap_golden import colour_field, light_constants  # noqa: E402

REF = "/root/reference/Bin"
F32, U32 = np.float32, np.uint32

# name: (grid, seed, light samples, probes, ray samples, cube size, eye point (world), faces forced off)
CASES = {
    "outside_16": ((16, 16, 16), 11, 24, 1, 48, 16, (14.0, 9.0, -27.0), 0),
    "outside_ambient_lod1": ((16, 16, 16), 11, 24, 0, 32, 8, (-30.0, 4.0, 8.0), 0),
    "inside_24x24x8": ((24, 24, 8), 5, 32, 1, 64, 16, (1.5, -2.0, 3.0), 0),
    "culled_faces_few_samples": ((12, 12, 12), 7, 5, 1, 6, 8, (40.0, 40.0, 40.0), 0b001001),
}


def visibility_mask(world_i, eye):
    """GenVisibilityMask (Fluid.cpp:51-63) with IsCubeFaceVisible (:41-46): bit i set iff face i can be seen."""
    local = (world_i[:, :3].astype(np.float64) @ np.asarray(eye, np.float64)) + world_i[:, 3]
    mask = 0
    for face in range(6):
        v = local[face >> 1]
        mask |= (1 if ((v > -1.0) if (face & 1) else (v < 1.0)) else 0) << face
    return mask


def view_constants(cbs_light, plain_light, ray_samples, cube_size, eye, faces_off):
    world_i = plain_light["world_i"].copy()
    world_i[:, 3] = np.array([0.05, -0.03, 0.02], F32)   # a translated volume: the dp4's w terms are not zero
    cb0 = cbs_light[0].view(F32).copy()
    cb0[8:11] = world_i
    cb1 = cbs_light[1].view(F32).copy()
    cb1[0, :3] = eye
    cb2 = np.zeros((1, 4), U32)
    cb2[0, 0] = ray_samples
    cb3 = np.zeros((1, 4), U32)
    mask = visibility_mask(world_i, eye) & ~faces_off
    cb3[0, 0] = mask
    plain = {"eye_pt": np.array(eye, F32), "world_i": world_i, "num_samples": ray_samples, "visibility_mask": mask,
             "cube_size": cube_size}
    return {0: cb0.view(U32), 1: cb1.view(U32), 2: cb2, 3: cb3}, plain


def run_case(blobs, grid, seed, light_samples, probes, ray_samples, cube_size, eye, faces_off):
    nx, ny, nz = grid
    col = colour_field(grid, seed)
    col[..., :3] = (col[..., :3].astype(F32) * 1.5).astype(np.float16)
    cbs_l, plain_l = light_constants(light_samples, probes, (75.0, 75.0, -75.0), seed)
    plain_l["light_color"][3] = 2.0   # keeps a good part of the cube map below saturation
    plain_l["ambient"][3] = 0.5
    cb1 = cbs_l[1].view(F32).copy()
    cb1[2, 3], cb1[3, 3] = 2.0, 0.5
    cbs_l[1] = cb1.view(U32)
    lmap = np.zeros((nz, ny, nx), U32)
    D.Machine(blobs["CSRayMarchL"], grid, cbs_l, srv={0: D.Texture(col, "rgba16f"), 1: plain_l["sh"].view(U32)},
              uav={0: D.Texture(lmap, "r11g11b10f")}, clamp=True).run()
    cbs_v, plain_v = view_constants(cbs_l, plain_l, ray_samples, cube_size, eye, faces_off)
    cube = np.zeros((6, cube_size, cube_size, 4), np.uint8)
    m = D.Machine(blobs["CSRayMarchV"], (cube_size, cube_size, 6), cbs_v,
                  srv={0: D.Texture(col, "rgba16f"), 1: D.Texture(lmap, "r11g11b10f")},
                  uav={0: D.Texture(cube, "rgba8unorm")}, clamp=True).run()
    # The non-separated march (Fluid::rayMarch, Fluid.cpp:825-855): CSRayMarch.cso casts the light (and occlusion) ray
    # at every view sample instead of reading a light map; cbSampleRes = (ray samples, has probes, light samples).
    cbs_f = dict(cbs_v)
    cb2 = np.zeros((1, 4), U32)
    cb2[0, 0], cb2[0, 1], cb2[0, 2] = ray_samples, probes, light_samples
    cbs_f[2] = cb2
    cube_full = np.zeros((6, cube_size, cube_size, 4), np.uint8)
    D.Machine(blobs["CSRayMarch"], (cube_size, cube_size, 6), cbs_f,
              srv={0: D.Texture(col, "rgba16f"), 1: plain_l["sh"].view(U32)},
              uav={0: D.Texture(cube_full, "rgba8unorm")}, clamp=True).run()
    return col, plain_l, plain_v, lmap, cube, m, cube_full


def main():
    blobs = {n: open(os.path.join(REF, n + ".cso"), "rb").read() for n in ("CSRayMarchL", "CSRayMarchV", "CSRayMarch")}
    res = {"blob_sha256/" + n: np.frombuffer(hashlib.sha256(b).digest(), np.uint8) for n, b in blobs.items()}
    for name, case in CASES.items():
        col, plain_l, plain_v, lmap, cube, m, cube_full = run_case(blobs, *case)
        res[name + "/cube_map"] = cube
        res[name + "/cube_map_full"] = cube_full
        diff = np.abs(cube.astype(int) - cube_full.astype(int))
        print(name, "non-separated march: texels written", int((cube_full[..., 3] > 0).sum()), "max |difference| to the "
              "light-map version", int(diff.max()), "texels differing", int((diff.max(-1) > 0).sum()))
        res[name + "/light_map"] = lmap
        res[name + "/input_sha256"] = np.frombuffer(hashlib.sha256(col.tobytes() + plain_l["sh"].tobytes()).digest(), np.uint8)
        a = cube[..., 3]
        print(name, "mask", bin(plain_v["visibility_mask"]), "texels written (alpha > 0):", int((a > 0).sum()), "of", a.size,
              "saturated rgb:", int((cube[..., :3] == 255).sum()), "distinct texels:", len(np.unique(cube.reshape(-1, 4), axis=0)),
              "loop trips:", m.iterations)
    path = os.path.join(HERE, "raymarch_golden.npz")
    np.savez_compressed(path, **res)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
