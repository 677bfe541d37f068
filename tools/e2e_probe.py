"""Where the host-side time of the synchronous end-to-end loop goes (bench.py's e2e): variants of the per-step read-back.
    gpurun -- 'python tools/e2e_probe.py 512'"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluidx12_b200 as fx

n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
f = fx.Fluid()
assert f.Init(gridSize=(n, n, n)), f.last_error
dt = fx.dt_for_grid(n, n, n)
stream = torch.cuda.Stream()
for _ in range(100):
    f.UpdateFrame(dt); f.Simulate(stream.cuda_stream)
torch.cuda.synchronize()
K = 60


def loop(body):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(K):
        f.UpdateFrame(dt); f.Simulate(stream.cuda_stream)
        body()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / K


print("no read-back (device rate)      %.4f ms/step" % loop(lambda: None))
print("stream.synchronize() per step   %.4f ms/step" % loop(stream.synchronize))
print("f.stats() per step              %.4f ms/step" % loop(f.stats))
slot = [0]


def posted():
    f.post_stats(slot[0] & 3)
    f.wait_stats(slot[0] & 3)
    slot[0] += 1


print("post_stats + wait_stats (same)  %.4f ms/step" % loop(posted))
t0 = time.perf_counter()
for _ in range(2000):
    f.UpdateFrame(0.0)
print("UpdateFrame call                %.2f us" % ((time.perf_counter() - t0) * 1e6 / 2000))

# device time of a step launched into an idle stream vs. back to back, and the host-side gap
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K)]
torch.cuda.synchronize()
t_sync = []
for i in range(K):
    f.UpdateFrame(dt)
    ev[2 * i].record(stream)
    f.Simulate(stream.cuda_stream)
    ev[2 * i + 1].record(stream)
    t0 = time.perf_counter()
    stream.synchronize()
    t_sync.append(time.perf_counter() - t0)
dev = [ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(K)]
gap = [ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(K - 1)]
print("synchronised steps: device %.4f ms/step, idle gap between steps %.4f ms, host wait in synchronize %.4f ms"
      % (sum(dev) / K, sum(gap) / (K - 1), 1e3 * sum(t_sync) / K))
for i in range(K):
    f.UpdateFrame(dt)
    ev[2 * i].record(stream)
    f.Simulate(stream.cuda_stream)
    ev[2 * i + 1].record(stream)
torch.cuda.synchronize()
dev = [ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(K)]
gap = [ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(K - 1)]
print("back-to-back steps: device %.4f ms/step, gap %.4f ms" % (sum(dev) / K, sum(gap) / (K - 1)))
