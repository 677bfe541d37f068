"""Freeze statistics of one pressure solve, pass by pass, from the CPU oracle (no GPU needed).

Spins the oracle up `--spin` steps, then walks the next step's relaxation two sweeps at a time and prints, per fused
pass, the cells still active and the number of 120 x 12 x 8 bricks (the default brick of jacobi_fused.cu) that still
hold an active cell.  This is the data the tail strategy of the pressure solve is sized with."""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=256)
    ap.add_argument("--spin", type=int, default=100)
    ap.add_argument("--t", type=int, default=2)
    ap.add_argument("--brick", type=int, nargs=3, default=[120, 12, 8])
    a = ap.parse_args()
    n = a.grid
    dt = O.dt_for_grid(n, n, n)
    f = O.FluidOracle(n, n, n)
    t0 = time.time()
    for _ in range(a.spin):
        f.step(dt)
    print(f"spin-up {a.spin} steps of {n}^3: {time.time() - t0:.1f} s, s_exec {f.s_exec}", flush=True)
    vel, col, p = f.get_field(O.FIELD_VEL), f.get_field(O.FIELD_COLOR), f.get_field(O.FIELD_PRESSURE)
    vo, _ = O.advect(vel, col, dt)
    s = O.divergence2x(vo)
    active = np.ones(s.shape, np.uint8)
    bx, by, bz = a.brick
    nbx, nby, nbz = -(-n // bx), -(-n // by), -(-n // bz)
    print(f"bricks {nbx}x{nby}x{nbz} = {nbx * nby * nbz}")
    pad = np.zeros((nbz * bz, nby * by, nbx * bx), np.uint8)
    for k in range(64 // a.t):
        p, active, counts = O.jacobi_sweeps_slab(s, p, active, a.t, n, 0, 0, n)
        pad[:n, :n, :n] = active
        per_brick = pad.reshape(nbz, bz, nby, by, nbx, bx).any(axis=(1, 3, 5))
        nb = int(per_brick.sum())
        zz, yy, xx = np.nonzero(per_brick)
        box = (f"x[{xx.min()},{xx.max()}] y[{yy.min()},{yy.max()}] z[{zz.min()},{zz.max()}]" if nb else "-")
        print(f"pass {k:2d} sweeps {a.t * (k + 1):2d}: active cells {int(counts[-1]):10d} "
              f"({counts[-1] / active.size:8.5f})  bricks {nb:6d}  brick box {box}", flush=True)
        if counts[-1] == 0:
            break


if __name__ == "__main__":
    main()
