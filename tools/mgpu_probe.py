"""Per-pass timeline of the pressure solve on every rank of a z-slab run (library built with -DFXB_TIMING=<pass>):
    make -C fluidx12_b200/csrc -j8 EXTRA=-DFXB_TIMING=16 BUILD=build_timing OUT=../libfluidx_b200_timing.so
    gpurun --gpus 2 -- 'FXB_LIB=$PWD/fluidx12_b200/libfluidx_b200_timing.so python -m torch.distributed.run --nnodes=1 \
        --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29671 tools/mgpu_probe.py'
Prints, for the last step, when each fused pass ended on each rank relative to the end of the divergence (us)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluidx12_b200 as fx
from fluidx12_b200 import binding as B

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
local = int(os.environ.get("LOCAL_RANK", rank))
dist.init_process_group("gloo")
torch.cuda.set_device(local)
grids = {1: (512, 512, 512), 2: (512, 512, 1024), 4: (1024, 1024, 512), 8: (1024, 1024, 1024)}
grid = grids[world]
uid = torch.zeros(128, dtype=torch.uint8)
if rank == 0:
    raw = (C.c_char * 128)()
    B.check(fx.lib().fxb_nccl_unique_id(C.cast(raw, C.c_void_p)))
    uid = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
dist.broadcast(uid, 0)
f = fx.Fluid()
assert f.Init(gridSize=grid, device=local, rank=rank, nranks=world, nccl_unique_id=uid.numpy().tobytes(),
              phase_timing=True), f.last_error
dt = fx.dt_for_grid(*grid)
for _ in range(int(os.environ.get("FXB_PROBE_STEPS", "110"))):
    f.step(dt)
f.sync()
out = (C.c_longlong * 128)()
assert fx.lib().fxb_debug_stamps(f._h, out, 128) == 0
raw = np.array(out[:], np.int64)
t0 = raw[106]
ends = [(int(raw[64 + k]) - int(t0)) / 1e3 for k in range(32)]
durs = [ends[0]] + [ends[k] - ends[k - 1] for k in range(1, 32)]
line = "rank %d: solve %.1f us | pass durations: %s" % (rank, (int(raw[105]) - int(t0)) / 1e3, " ".join("%.1f" % d for d in durs))
lines = [None] * world
dist.all_gather_object(lines, line)
if rank == 0:
    print("\n".join(lines), flush=True)
dist.barrier()
f.close()
dist.destroy_process_group()
