"""Debug probe (library built with EXTRA=-DFXB_TIMING): clock marks of warp 1 of CTA 0 inside each marching iteration
of the last brick it relaxed in the last pass of the last step.  Marks: 0 top, 1 after TMA wait, 2 after head(1) +
level-0 load, 3 after head(2) + tail(1), 4 after the remaining levels, 5 after the barrier."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import fluidx12_b200 as fx
g = int(sys.argv[1]) if len(sys.argv) > 1 else 256
f = fx.Fluid(); assert f.Init(gridSize=(g, g, g), use_graph=False), f.last_error
dt = fx.dt_for_grid(g, g, g)
for _ in range(120): f.step(dt)
f.sync()
h = f.freeze_histogram(128).astype(np.int64)[32:128].reshape(16, 6)
n = int((h[:, 0] > 0).sum())
print("s_exec", f.stats().s_exec, "iterations recorded", n)
d = np.diff(h[:n], axis=1)
print("segments per iteration (cycles): wait | head1+load0 | head2+tail1 | rest | barrier")
print(d)
print("iteration period:", np.diff(h[:n, 0]))
