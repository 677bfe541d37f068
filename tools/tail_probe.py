"""Phase timing of the tail kernel (library built with `make -C fluidx12_b200/csrc -B EXTRA=-DFXB_TAIL_TIMING`).

Thread 0 of every CTA accumulates the cycles between the marks of tail_run_item (jacobi_tail_body.cuh) per path;
this prints the average per work item.  Marks: 0 ctrl reset, 1 flags, 2 scan, 3 window load / list build, 4 rhs gather,
5 the sweeps, 6 thread 0's brick bookkeeping (atomics, list append) followed by its share of the store (or copy), 7 (unused).  Debug aid, not a benchmark."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("FXB_TAIL", "1")
import fluidx12_b200 as fx  # noqa: E402

g = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 120
f = fx.Fluid()
assert f.Init(gridSize=(g, g, g)), f.last_error
assert f.tail_stats()["enabled"], "the dynamic schedule is not available for this grid"
dt = fx.dt_for_grid(g, g, g)
for _ in range(steps):
    f.step(dt)
f.sync()
m = f.freeze_histogram(128).astype(np.int64)[64:96]
print("tail stats", f.tail_stats(), "s_exec", f.stats().s_exec)
names = ("ctrl", "flags", "scan", "load/build", "gather", "sweeps", "bookkeeping+store", "-")
for path, label in enumerate(("copy", "sparse", "dense")):
    n = int(m[24 + path])
    if n == 0:
        print(f"{label}: no items (was the library built with -DFXB_TAIL_TIMING?)")
        continue
    avg = m[8 * path:8 * path + 8] / n
    print(f"{label}: {n} items, {avg.sum():.0f} cycles/item: " + ", ".join(f"{k} {v:.0f}" for k, v in zip(names, avg)))
