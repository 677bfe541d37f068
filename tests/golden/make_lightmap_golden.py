"""Golden vectors of the light-map pass from the reference's own compiled shader.

Runs Bin/CSRayMarchL.cso through tests/golden/dxbc_interp.py on seeded colour fields and writes
tests/golden/lightmap_golden.npz: per case the packed R11G11B10_FLOAT light map the bytecode stores.  Needs
/root/reference (this container only); the vectors and this script are committed, the reference is not."""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import dxbc_interp as D  # noqa: E402

REF = "/root/reference/Bin"
F32, U32 = np.float32, np.uint32

# name: (grid, seed, num_samples, has_light_probes, light point)
CASES = {
    "plume_16_probes": ((16, 16, 16), 11, 24, 1, (75.0, 75.0, -75.0)),
    "plume_16_ambient": ((16, 16, 16), 11, 24, 0, (75.0, 75.0, -75.0)),
    "slab_24x24x8_probes": ((24, 24, 8), 5, 64, 1, (-20.0, 90.0, 35.0)),
    "dense_12_few_samples": ((12, 12, 12), 7, 5, 1, (0.0, 100.0, 0.0)),
}


def colour_field(grid, seed):
    """Premultiplied smoke-like colour: two Gaussian blobs of density (one saturating at 1), rgb = tint * density, and a
    uniform-density slab in one corner (zero gradient inside: the `any(abs(rayDir) > 0)` branch)."""
    nx, ny, nz = grid
    r = np.random.default_rng(seed)
    z, y, x = np.meshgrid((np.arange(nz) + 0.5) / nz, (np.arange(ny) + 0.5) / ny, (np.arange(nx) + 0.5) / nx, indexing="ij")
    dens = np.zeros((nz, ny, nx))
    for _ in range(2):
        c = r.uniform(0.25, 0.75, 3)
        s = r.uniform(0.08, 0.2)
        dens += r.uniform(0.6, 1.6) * np.exp(-((x - c[0]) ** 2 + (y - c[1]) ** 2 + (z - c[2]) ** 2) / (2 * s * s))
    dens += 0.02 * r.random((nz, ny, nx))
    dens[: max(nz // 4, 2), : max(ny // 4, 2), : max(nx // 4, 2)] = 0.5
    dens = np.clip(dens, 0.0, 1.0)
    dens[dens < 0.03] = 0.0
    col = np.zeros((nz, ny, nx, 4), np.float16)
    col[..., 0], col[..., 1], col[..., 2] = 0.2 * dens, 0.4 * dens, 1.0 * dens
    col[..., 3] = dens
    return col


def light_constants(num_samples, probes, light_pt, seed):
    """The constant buffers as Fluid::UpdateFrame / rayMarchL fill them (Fluid.cpp:171-182, 296-320, 872-874) for the
    default volume transform (scaling by 10) — plus a small rotation so that no matrix entry is trivially zero — and a
    seeded set of SH coefficients.  Returns (dict for the interpreter, dict of plain arrays for the other paths)."""
    r = np.random.default_rng(seed + 1000)
    a, b = 0.3, -0.2
    rz = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]])
    rx = np.array([[1, 0, 0], [0, np.cos(b), -np.sin(b)], [0, np.sin(b), np.cos(b)]])
    world3 = (10.0 * rz @ rx)
    world = np.zeros((3, 4), F32)
    world[:, :3] = world3.astype(F32)
    world_i = np.zeros((3, 4), F32)
    world_i[:, :3] = np.linalg.inv(world3).astype(F32)
    sh = (r.standard_normal((9, 3)) * np.array([1.0] + [0.35] * 8)[:, None]).astype(F32)
    sh[0] = np.abs(sh[0]) + 0.8
    light_color = np.array([1.0, 0.7, 0.3, np.pi * 3.0], F32)
    ambient = np.array([1.0, 1.0, 1.0, np.pi * 1.5], F32)
    cb0 = np.zeros((14, 4), F32)
    cb0[8:11], cb0[11:14] = world_i, world
    cb1 = np.zeros((4, 4), F32)
    cb1[1, :3], cb1[2], cb1[3] = light_pt, light_color, ambient
    cb2 = np.zeros((1, 4), U32)
    cb2[0, 0], cb2[0, 1] = num_samples, probes
    plain = {"light_pt": np.array(light_pt, F32), "light_color": light_color, "ambient": ambient, "world_i": world_i,
             "world": world, "num_samples": num_samples, "has_light_probes": probes, "sh": sh}
    return {0: cb0.view(U32), 1: cb1.view(U32), 2: cb2}, plain


def run_case(blob, grid, seed, num_samples, probes, light_pt):
    nx, ny, nz = grid
    col = colour_field(grid, seed)
    cbs, plain = light_constants(num_samples, probes, light_pt, seed)
    out = np.zeros((nz, ny, nx), U32)
    m = D.Machine(blob, grid, cbs, srv={0: D.Texture(col, "rgba16f"), 1: plain["sh"].view(U32)},
                  uav={0: D.Texture(out, "r11g11b10f")}, clamp=True).run()
    return col, plain, out, m


def main():
    blob = open(os.path.join(REF, "CSRayMarchL.cso"), "rb").read()
    res = {"blob_sha256": np.frombuffer(hashlib.sha256(blob).digest(), np.uint8)}
    for name, (grid, seed, ns, probes, lp) in CASES.items():
        col, plain, out, m = run_case(blob, grid, seed, ns, probes, lp)
        res[name + "/light_map"] = out
        res[name + "/input_sha256"] = np.frombuffer(hashlib.sha256(col.tobytes() + plain["sh"].tobytes()).digest(), np.uint8)
        lit = int((col[..., 3].astype(F32) >= 0.01).sum())
        print(name, grid, "voxels with density:", lit, "of", col[..., 3].size, "loop trips:", m.iterations,
              "distinct words:", len(np.unique(out)))
    path = os.path.join(HERE, "lightmap_golden.npz")
    np.savez_compressed(path, **res)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
