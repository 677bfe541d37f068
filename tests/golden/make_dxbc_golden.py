"""Generates tests/golden/dxbc_golden.npz: outputs of the reference's own compiled shaders, interpreted.

Run in the build container only (it reads /root/reference/Bin/*.cso, which does not exist on the GPU box):
    python tests/golden/make_dxbc_golden.py
Each case starts from a seeded state (tests/util.smooth_state) or from the reference's all-zero start, then runs
`steps` frames of  CSAdvect.cso -> CSProject3D.cso / CSProject2D.cso  through tests/golden/dxbc_interp.py with the
reference's bindings (Fluid.cpp:729-758: advect reads velocity[0] + colour[!parity], writes velocity[1] + colour[parity];
project reads velocity[1], writes velocity[0], relaxes the pressure texture in place) and its dt rule
(FluidX12.cpp:266-267).  The file stores the resulting fields, the number of loop trips per frame and a checksum of the
inputs; tests/test_dxbc_golden.py requires the CPU oracle (and, on a GPU, the CUDA path) to reproduce them bit for bit.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
import dxbc_interp as D  # noqa: E402
from tests.util import smooth_state  # noqa: E402

REF = "/root/reference/Bin"
# name: (grid (nx, ny, nz), seed or None for the zero state, steps, clamp sampler, paused frame index or None)
CASES = {
    "smooth3d": ((16, 16, 8), 1234, 2, False, None),
    "emitter3d_from_zero": ((16, 16, 16), None, 3, False, None),
    "smooth3d_clamp_pause": ((16, 16, 8), 77, 3, True, 1),
    "smooth2d": ((32, 32, 1), 5, 2, False, None),
    "emitter3d_32_cubed": ((32, 32, 32), None, 8, False, None),
    "emitter2d_from_zero": ((64, 64, 1), None, 6, False, None),
}


def start_state(grid, seed):
    nx, ny, nz = grid
    if seed is None:
        return (np.zeros((nz, ny, nx, 4), np.float16), np.zeros((nz, ny, nx, 4), np.float16),
                np.zeros((nz, ny, nx), np.float32))
    return smooth_state(nx, ny, nz, seed=seed, umax=1.0)


def run_case(blobs, grid, seed, steps, clamp, pause):
    nx, ny, nz = grid
    vel, col, p = start_state(grid, seed)
    vel = [vel.copy(), np.zeros_like(vel)]
    col = [np.zeros_like(col), np.zeros_like(col)]
    col[0] = start_state(grid, seed)[1].copy()  # m_colors[m_frameParity], parity starts at 0
    p = p.copy()
    parity, trips = 0, []
    dt_rule = np.float32(2.0 if nz > 1 else 1.0) / np.float32(ny)
    for k in range(steps):
        dt = np.float32(0.0) if k == pause else dt_rule
        if dt > 0:
            parity ^= 1  # Fluid.cpp:345
        cb0 = np.array([np.float32(dt).view(np.uint32), 0, 0, 0], np.uint32)
        D.Machine(blobs["CSAdvect"], grid, cb0, srv={0: D.Texture(vel[0], "rgba16f"), 1: D.Texture(col[parity ^ 1], "rgba16f")},
                  uav={0: D.Texture(vel[1], "rgba16f"), 1: D.Texture(col[parity], "rgba16f")}, clamp=clamp).run()
        proj = "CSProject3D" if nz > 1 else "CSProject2D"
        m = D.Machine(blobs[proj], grid, cb0, srv={0: D.Texture(vel[1], "rgba16f")},
                      uav={0: D.Texture(vel[0], "rgba16f"), 1: D.Texture(p, "r32f")}).run()
        trips.append(m.iterations)
    return {"velocity": vel[0], "velocity_advected": vel[1], "colour": col[parity], "pressure": p,
            "loop_trips": np.array(trips, np.int32)}


def main():
    blobs = {n: open(os.path.join(REF, n + ".cso"), "rb").read() for n in ("CSAdvect", "CSProject3D", "CSProject2D")}
    out = {}
    for name, (grid, seed, steps, clamp, pause) in CASES.items():
        res = run_case(blobs, grid, seed, steps, clamp, pause)
        v0, c0, p0 = start_state(grid, seed)
        out[name + "/input_sha256"] = np.frombuffer(
            hashlib.sha256(v0.tobytes() + c0.tobytes() + p0.tobytes()).digest(), np.uint8)
        for k, v in res.items():
            out[name + "/" + k] = v
        print(name, grid, "loop trips per frame", res["loop_trips"].tolist(), "max|p|", float(np.abs(res["pressure"]).max()))
    for n, b in blobs.items():
        out["blob_sha256/" + n] = np.frombuffer(hashlib.sha256(b).digest(), np.uint8)
    path = os.path.join(HERE, "dxbc_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
