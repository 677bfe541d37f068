"""Drives a few un-graphed steps for ncu (see profiles/README.md).  Not a benchmark: numbers under a profiler are
never reported as performance."""
import argparse
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluidx12_b200 as fx

ap = argparse.ArgumentParser()
ap.add_argument("--grid", type=int, nargs=3, default=[256, 256, 256])
ap.add_argument("--steps", type=int, default=101)
ap.add_argument("--fuse-t", type=int, default=0)
ap.add_argument("--graph", type=int, default=0)
a = ap.parse_args()
f = fx.Fluid()
assert f.Init(gridSize=tuple(a.grid), use_graph=bool(a.graph), fuse_t=a.fuse_t), f.last_error
dt = fx.dt_for_grid(*a.grid)
for _ in range(a.steps):
    f.step(dt)
f.sync()
st = f.stats()
print("kernels/step", st.kernels_per_step, "s_exec", st.s_exec, "passes", st.jacobi_passes)
