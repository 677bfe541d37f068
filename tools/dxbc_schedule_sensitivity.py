"""How much does the reference's result depend on the thread schedule of its unsynchronised relaxation loop?

CSPoisson.hlsli:8-26 relaxes the pressure in place over a globallycoherent UAV with a device fence but no barrier, so a
real GPU interleaves the threads arbitrarily.  The oracle (and the CUDA path) take the lock-step reading (SURVEY.md
App. A.3).  This script runs the reference's own compiled CSProject3D.cso (tests/golden/dxbc_interp.py) under the other
extreme — thread groups one after the other, each running its whole loop before the next group starts, in ascending
or descending group order — and reports the distance to the lock-step result.  Build container only (reads
/root/reference).  Result quoted in DESIGN.md §2."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
import dxbc_interp as D  # noqa: E402
from tests.util import smooth_state  # noqa: E402


def project(blob, grid, vel1, p, dt, order):
    nx, ny, nz = grid
    vel0, p = np.zeros_like(vel1), p.copy()
    cb0 = np.array([np.float32(dt).view(np.uint32), 0, 0, 0], np.uint32)
    srv, uav = {0: D.Texture(vel1, "rgba16f")}, {0: D.Texture(vel0, "rgba16f"), 1: D.Texture(p, "r32f")}
    if order == "lockstep":
        D.Machine(blob, grid, cb0, srv, uav).run()
        return vel0, p
    gx, gy, gz = 4, 4, 4
    groups = [(bx, by, bz) for bz in range(nz // gz) for by in range(ny // gy) for bx in range(nx // gx)]
    if order == "descending":
        groups.reverse()
    for bx, by, bz in groups:
        z, y, x = np.meshgrid(np.arange(bz * gz, (bz + 1) * gz), np.arange(by * gy, (by + 1) * gy),
                              np.arange(bx * gx, (bx + 1) * gx), indexing="ij")
        ids = np.stack([x.reshape(-1), y.reshape(-1), z.reshape(-1)], 1)
        D.Machine(blob, grid, cb0, srv, uav, threads=ids).run()
    return vel0, p


def main():
    blob = open("/root/reference/Bin/CSProject3D.cso", "rb").read()
    adv = open("/root/reference/Bin/CSAdvect.cso", "rb").read()
    grid = (16, 16, 16)
    nx, ny, nz = grid
    dt = np.float32(2.0) / np.float32(ny)
    vel, col, p = smooth_state(nx, ny, nz, seed=1234, umax=1.0)
    vel1, col1 = np.zeros_like(vel), np.zeros_like(col)
    cb0 = np.array([np.float32(dt).view(np.uint32), 0, 0, 0], np.uint32)
    D.Machine(adv, grid, cb0, srv={0: D.Texture(vel, "rgba16f"), 1: D.Texture(col, "rgba16f")},
              uav={0: D.Texture(vel1, "rgba16f"), 1: D.Texture(col1, "rgba16f")}).run()
    ref_v, ref_p = project(blob, grid, vel1, p, dt, "lockstep")
    for order in ("ascending", "descending"):
        v, q = project(blob, grid, vel1, p, dt, order)
        dv = np.abs(v[..., :3].astype(np.float32) - ref_v[..., :3].astype(np.float32))
        dp = np.abs(q - ref_p)
        print("%-10s groups one after the other: pressure max|d| = %.3e (%.2e of max|p|), rel L2 = %.3e; "
              "velocity max|d| = %.3e (%.2e of max|u|), rel L2 = %.3e" %
              (order, dp.max(), dp.max() / np.abs(ref_p).max(), np.sqrt((dp ** 2).sum() / (ref_p ** 2).sum()),
               dv.max(), dv.max() / np.abs(ref_v[..., :3].astype(np.float32)).max(),
               np.sqrt((dv ** 2).sum() / (ref_v[..., :3].astype(np.float32) ** 2).sum())))


if __name__ == "__main__":
    main()
