"""Step time of the graph-launched step at a few grids, in seconds of GPU time rather than the minutes bench.py takes:
    gpurun -- 'python tools/quick_time.py 256 512'   (environment knobs such as FXB_RESIDENT_FROM apply)
Spin-up 100 steps from the zero state, then 30 timed steps (host wall clock around a device sync; phase marks on).
Prints one line per grid with the checksum of the state, so variants can be compared bit for bit."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluidx12_b200 as fx

for arg in sys.argv[1:]:
    n = int(arg)
    f = fx.Fluid()
    assert f.Init(gridSize=(n, n, n), phase_timing=True), f.last_error
    dt = fx.dt_for_grid(n, n, n)
    for _ in range(100):
        f.step(dt)
    f.sync()
    f.phase_times(reset=True)
    steps = 30
    t0 = time.perf_counter()
    for _ in range(steps):
        f.step(dt)
    f.sync()
    ms = (time.perf_counter() - t0) * 1e3 / steps
    ph = {k: round(v / steps, 4) for k, v in f.phase_times().items()}
    st = f.stats()
    print("grid %d^3  ms/step %.4f  phases %s  s_exec %d passes %d  checksum %s  knobs %s" % (
        n, ms, ph, st.s_exec, st.jacobi_passes, "-".join("%016x" % v for v in f.state_checksum()),
        {k: v for k, v in os.environ.items() if k.startswith("FXB_") and k != "FXB_LIB"}), flush=True)
    f.close()
