"""Cycle stamps of the first CTA's first brick of one fused Jacobi pass (library built with EXTRA=-DFXB_TIMING=<pass>):
    make -C fluidx12_b200/csrc -j8 EXTRA=-DFXB_TIMING=16 BUILD=build_timing OUT=../libfluidx_b200_timing.so
    gpurun -- 'FXB_LIB=$PWD/fluidx12_b200/libfluidx_b200_timing.so python tools/timing_probe.py 256'
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
raw = np.array(out[:], np.int64)
if raw[121]:
    t0g = raw[122]
    print("n_relax %d n_copy %d" % (raw[124], raw[125]))
    print("grid barrier, ns relative to CTA 0's arrival: last arrival CTA %d at %d, CTA 0 released at %d"
          % (raw[121] & 0xff, (raw[121] >> 8) - t0g, raw[123] - t0g))
    print("pass starts of CTA 0..39:  " + " ".join(str(int(x - t0g)) for x in raw[80:120] if x))
    print("arrivals of CTA 0..39:     " + " ".join(str(int(x - t0g)) for x in raw[40:80] if x))
    print("durations:                 " + " ".join(str(int(a - b)) for a, b in zip(raw[40:80], raw[80:120]) if a and b))
v = raw[:40]
v = v[v != 0]
if len(v) < 4:
    print("no stamps (wrong pass number or not a debug build)", len(v))
    sys.exit(0)
t0 = v[0]
if os.environ.get("FXB_PROBE_RESIDENT", "1") == "1":
    # jacobi_resident.cu: 0 start | 1 prologue loads | 2 copies | 3 window requested | per brick: flags requested,
    # window landed, level 1 done, written back, last level done, stores issued | last: counters
    names = ["pass start", "pass loads", "copies handed out", "window requested"]
    per = ["flags requested", "window landed", "level 1 done", "written back", "last level done", "stores issued"]
    print("clock 1965 MHz: 1 us = 1965 cycles")
    nb = (len(v) - 5) // 6
    for i, t in enumerate(v):
        if i < 4:
            nm = names[i]
        elif i >= 4 + 6 * nb:
            nm = "counters out"
        else:
            nm = "brick %d: %s" % ((i - 4) // 6, per[(i - 4) % 6])
        print("%-32s +%7d cycles (%6.2f us)  delta %6d" % (nm, t - t0, (t - t0) / 1965.0, t - (v[i - 1] if i else t0)))
    sys.exit(0)
t0 = v[0]
if os.environ.get("FXB_PROBE_RESIDENT", "1") == "1":
    # jacobi_resident.cu: 0 start | 1 prologue loads | 2 copies | 3 window requested | per brick: flags requested,
    # window landed, level 1 done, written back, last level done, stores issued | last: counters
    names = ["start", "prologue loads", "copies handed out", "window requested"]
    per = ["flags requested", "window landed", "level 1 done", "written back", "last level done", "stores issued"]
    print("clock 1965 MHz: 1 us = 1965 cycles")
    for i, t in enumerate(v):
        if i < 4:
            nm = names[i]
        elif i == len(v) - 1:
            nm = "counters out"
        else:
            nm = "brick %d: %s" % ((i - 4) // 6, per[(i - 4) % 6])
        print("%-32s +%7d cycles (%6.2f us)  delta %6d" % (nm, t - t0, (t - t0) / 1965.0, t - (v[i - 1] if i else t0)))
    sys.exit(0)
print("prologue (kernel start -> barriers initialised, copies done): %d cycles" % (v[1] - v[0]))
it = v[2:]
per = 5
for i in range(0, len(it) - per + 1, per):
    a = it[i:i + per]
    print("iter %2d: top +%6d | issue %5d | mbar wait %6d | flag bytes %6d | levels %6d | (sync -> next top %s)" %
          (i // per, a[0] - t0, a[1] - a[0], a[2] - a[1], a[3] - a[2], a[4] - a[3],
           (it[i + per] - a[4]) if i + per < len(it) else "-"))
