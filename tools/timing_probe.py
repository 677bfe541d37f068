"""Cycle stamps of the first CTA's first brick of one fused Jacobi pass (library built with EXTRA=-DFXB_TIMING=<pass>):
    make -C fluidx12_b200/csrc clean && make -C fluidx12_b200/csrc -j8 EXTRA=-DFXB_TIMING=16
    gpurun -- 'python tools/timing_probe.py 256'
Stamps per marching iteration: top, after the TMA issue, after the mbarrier wait, after the flag bytes arrived, end."""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluidx12_b200 as fx

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
f = fx.Fluid()
assert f.Init(gridSize=(n, n, n), use_graph=False), f.last_error
dt = fx.dt_for_grid(n, n, n)
for _ in range(101):
    f.step(dt)
f.sync()
# StepState layout: see common.cuh (dbg is the last member)




print("stats", f.stats().s_exec, f.stats().jacobi_passes)
hist = f.freeze_histogram(64)
# read dbg through the freeze histogram entry point is not possible: use fxb_debug_read
out = (C.c_longlong * 128)()
assert fx.lib().fxb_debug_stamps(f._h, out, 128) == 0
v = np.array(out[:], np.int64)
v = v[v != 0]
if len(v) < 4:
    print("no stamps (wrong pass number or not a debug build)", len(v))
    sys.exit(0)
t0 = v[0]
print("prologue (kernel start -> barriers initialised, copies done): %d cycles" % (v[1] - v[0]))
it = v[2:]
per = 5
for i in range(0, len(it) - per + 1, per):
    a = it[i:i + per]
    print("iter %2d: top +%6d | issue %5d | mbar wait %6d | flag bytes %6d | levels %6d | (sync -> next top %s)" %
          (i // per, a[0] - t0, a[1] - a[0], a[2] - a[1], a[3] - a[2], a[4] - a[3],
           (it[i + per] - a[4]) if i + per < len(it) else "-"))
