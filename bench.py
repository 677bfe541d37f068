#!/usr/bin/env python
"""bench.py — voxel-updates/s of the smoke-solver step (advect + project) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]                 # our CUDA path (one JSON line)
    python bench.py --impl reference [--gpus N] --steps K --warmup W    # the CPU oracle on the host cores

A "step" is one full frame of the hot path: Fluid::UpdateFrame + Fluid::Simulate (CSAdvect, then divergence, <= 64
Jacobi sweeps and gradient-subtract of CSProject3D) over the whole grid.  Workload (BASELINE.json): synthetic
emitter-driven smoke from the all-zero state, dt = 2/Ny, MIRROR addressing, ITER = 64 with the per-cell early exit.
The state is first spun up for --spinup steps (untimed state preparation: at step 0 nothing moves and the solver would
be trivially cheap), then W warm-up steps, then exactly K timed steps (CUDA events on the launching stream, max over
ranks).

Grid: N = 1 -> 512^3 (BASELINE config 4, its 1-GPU point: the largest single-GPU configuration; every field is far
larger than L2).  N > 1 -> weak scaling at 134 M voxels per GPU, z-slab decomposed: 512x512x1024 (2), 1024x1024x512 (4),
1024^3 (8, BASELINE config 5); `--scaling strong` keeps 512^3 at every N (config 4).  The per-phase device times and
the roofline come from phase marks INSIDE the timed steps (fxb_get_phase_times).  At N = 1 the line also carries "c3"
(the 256^3 roofline-characterisation config), "c2" (128^3), "c150" (150^3) and the CPU baseline; at N > 1 (weak) it carries
"c4_strong": config 4 on the same N ranks, whose state checksum must equal the N = 1 line's.  --grid overrides.

"e2e" is the same K steps host-timed through the public calls (UpdateFrame with the 8-byte constant upload, Simulate,
fxb_get_stats read-back every step) on the SAME frames as "value": a second simulator is spun up identically, because the
flow keeps developing and later frames cost more (its final state checksum is compared: e2e.same_state_as_value).
"roofline.traffic" is the ncu DRAM traffic per launch of the dominant phase's kernel(s) from profiles/traffic.json.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WEAK_GRIDS = {1: (512, 512, 512), 2: (512, 512, 1024), 4: (1024, 1024, 512), 8: (1024, 1024, 1024)}
STRONG_GRID = (512, 512, 512)
CPU_SAMPLE_GRID = (256, 256, 256)  # the same sample at every N and in both arms
METRIC = "voxel_updates_per_s"
UNIT = "voxel-updates/s"
# algorithmic bytes per voxel per launch
PHASE_BYTES = {"advect": 32.0, "divergence": 12.0, "gradient": 20.0}
PHASE_KERNEL = {"advect": "advect_kernel", "divergence": "divergence_quad_kernel", "jacobi": "jacobi_pass_kernel+jacobi_resident_kernel",
                "gradient": "gradient_quad_kernel"}


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def bytes_per_voxel_step(passes: float, mask_bytes_per_pass: float) -> float:
    """Nominal HBM bytes per voxel per step (SURVEY.md §8d / BASELINE.md §3): advect 32 + divergence 12
    + per executed Jacobi pass 12 (+ freeze mask) + gradient-subtract 20 — as if every pass touched every voxel."""
    return 32.0 + 12.0 + passes * (12.0 + mask_bytes_per_pass) + 20.0


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 25 ms while the GPU runs the warm-up and timed steps."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "25"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); smax.append(float(p[2])); power.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(smax), "power_w_max": max(power),
                "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def make_sim(fx, grid, args, rank, world, local_rank, uid):
    f = fx.Fluid()
    ok = f.Init(gridSize=grid, address_mode=fx.ADDRESS_MIRROR, early_exit=bool(args.early_exit), jacobi_iters=64,
                fuse_t=args.fuse_t, device=local_rank, rank=rank, nranks=world, use_graph=True,
                kernel_path=args.kernel_path, phase_timing=True,
                halo_backend={"nccl": fx.HALO_NCCL, "peer": fx.HALO_PEER, "fused": fx.HALO_FUSED}[args.halo],
                jacobi_group=args.jacobi_group,
                nccl_unique_id=uid)
    if not ok:
        raise RuntimeError("fluidx_b200 Init failed: " + f.last_error)
    return f


def timed_run(torch, dist, f, dt, steps, warmup, world, stream):
    """W warm-up + K timed steps, device-timed with CUDA events on the launching stream; max over ranks.  Also returns
    the per-phase device times accumulated by the phase marks inside those K steps (ms per step, this rank)."""
    for _ in range(warmup):
        f.UpdateFrame(dt); f.Simulate(stream.cuda_stream)
    stream.synchronize()
    st0 = f.stats()
    f.phase_times(reset=True)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            f.UpdateFrame(dt); f.Simulate(stream.cuda_stream)
        e1.record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    st1 = f.stats()
    phases = {k: v / max(steps, 1) for k, v in f.phase_times().items()}
    return ms, st0, st1, phases


def e2e_run(torch, dist, f, fx, dt, steps, world, stream, export):
    """The same K steps through the public call sequence a host application makes, host-timed:
    per step the frame constants go host->device (dt + parity, 8 bytes, the CBSimulation upload) and the
    step's result record comes back device->host (fxb_get_stats; with `export` also the whole colour field
    into pinned memory, the renderer hand-off format)."""
    import ctypes as C
    cb = torch.zeros(2, dtype=torch.float32).pin_memory()  # CBSimulation {TimeStep, BaseSeed} staging
    cb[0] = dt
    nbytes = 0
    host = None
    if export:
        nzl = f.slab[1]
        nbytes = f.m_gridSize[0] * f.m_gridSize[1] * nzl * 8
        host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    stats_bytes = C.sizeof(fx.FxbStats)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        f.UpdateFrame(float(cb[0]))
        f.Simulate(stream.cuda_stream)
        if export:
            f.get_field_async(fx.FIELD_COLOR, host.data_ptr(), nbytes, stream.cuda_stream)
        f.stats()  # D2H + stream sync: the step's result record
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sec = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([sec], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sec = float(t.item())
    return sec, 8, stats_bytes + nbytes


def jacobi_work_bytes(st0, st1, mask_bytes, voxels_local):
    """Bytes the Jacobi passes between two stats snapshots really had to move: a relaxed brick reads p + rhs and
    writes p (+ 2/8 B of freeze flags) per cell, a frozen brick is copied once (p in, p out), every other brick is
    skipped because its value is already final in both pressure buffers."""
    if st1.jacobi_fused:
        proc = st1.bricks_processed - st0.bricks_processed
        cop = st1.bricks_copied - st0.bricks_copied
        return (proc * 12.25 + cop * 8.0) * st1.brick_cells, proc, cop
    return (st1.total_passes - st0.total_passes) * (12.0 + mask_bytes) * voxels_local, 0, 0


def ncu_traffic(grid, phase):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the phase's kernel(s), averaged over every launch of
    whole steps, from the committed ncu pass over this grid (profiles/traffic.json, written by
    tools/summarize_profiles.py traffic), or None."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get("%dx%dx%d" % tuple(grid), {}).get(phase, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def rooflines(grid, phases, st0, st1, steps, voxels_local, peak, peak_src):
    """Per-phase rooflines from the phase marks of the timed steps, and the `roofline` object of the dominant kernel:
    algorithmic bytes per launch / average launch duration.  Advect, divergence and gradient are one launch per step;
    the Jacobi phase is `launches` launches of one kernel (its bytes: the bricks those very steps relaxed / copied)."""
    mask_bytes = 0.25 if st1.jacobi_fused else 2.0
    jb, proc, cop = jacobi_work_bytes(st0, st1, mask_bytes, voxels_local)
    jb /= max(steps, 1)
    launches = (st1.total_passes - st0.total_passes) / max(steps, 1)
    nbytes = {k: v * voxels_local for k, v in PHASE_BYTES.items()}
    nbytes["jacobi"] = jb
    per = {}
    for k in ("advect", "divergence", "jacobi", "gradient"):
        ms = phases.get(k, 0.0)
        gbs = nbytes[k] / max(ms * 1e-3, 1e-12) / 1e9
        per[k] = {"ms": round(ms, 4), "algorithmic_bytes": nbytes[k], "achieved_gbs": round(gbs, 1),
                  "frac": round(gbs / peak, 4)}
    per["jacobi"].update(launches_per_step=round(launches, 1), bricks_relaxed_per_step=round(proc / max(steps, 1), 1),
                         bricks_copied_per_step=round(cop / max(steps, 1), 1))
    dom = max(per, key=lambda k: per[k]["ms"])
    n_launch = launches if dom == "jacobi" else 1.0
    roof = {"bound": "hbm", "kernel": PHASE_KERNEL[dom], "phase": dom,
            "achieved": per[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": per[dom]["frac"],
            "peak_source": peak_src, "traffic": ncu_traffic(grid, dom),
            "bytes_per_launch": round(nbytes[dom] / max(n_launch, 1e-9)),
            "avg_launch_ms": round(per[dom]["ms"] / max(n_launch, 1e-9), 5), "launches_per_step": round(n_launch, 1),
            "how": "phase marks inside the timed graph-launched steps (fxb_get_phase_times); bytes of the same steps"}
    return roof, per, jb, proc, cop


def cpu_baseline_sample(f, fx, grid, dt, budget_s=25.0):
    """Times the OpenMP oracle on the host cores on a bounded sample of the same workload: the developed GPU
    state is copied into the oracle and stepped for a few frames; the result is also compared with the GPU
    (full-size parity spot check)."""
    import numpy as np
    import oracle
    cores = oracle.threads(os.cpu_count() or 1)
    nx, ny, nz = grid
    o = oracle.FluidOracle(nx, ny, nz)
    for gf, of in ((fx.FIELD_VELOCITY, oracle.FIELD_VEL), (fx.FIELD_COLOR, oracle.FIELD_COLOR),
                   (fx.FIELD_PRESSURE, oracle.FIELD_PRESSURE)):
        o.set_field(of, f.get_field(gf))
    n_steps, t_total = 0, 0.0
    while True:
        t0 = time.perf_counter()
        o.step(dt)
        t_total += time.perf_counter() - t0
        f.step(dt)
        n_steps += 1
        if t_total > budget_s or n_steps >= 8 or t_total + t_total / n_steps > 1.3 * budget_s:
            break
    f.sync()
    worst = 0.0
    for gf, of in ((fx.FIELD_VELOCITY, oracle.FIELD_VEL), (fx.FIELD_COLOR, oracle.FIELD_COLOR),
                   (fx.FIELD_PRESSURE, oracle.FIELD_PRESSURE)):
        a, b = f.get_field(gf).astype(np.float32), o.get_field(of).astype(np.float32)
        if gf == fx.FIELD_VELOCITY:
            a, b = a[..., :3], b[..., :3]
        worst = max(worst, float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)))
    return {"value": nx * ny * nz * n_steps / t_total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d oracle steps of the %dx%dx%d developed state (100 GPU steps from zero), OpenMP, %d threads, "
                      "%.1f s" % (n_steps, nx, ny, nz, cores, t_total),
            "parity_max_abs_rel_vs_gpu": worst}


def global_checksum(torch, dist, f, world):
    """Decomposition-independent checksum of the state: per field the sum over ranks (mod 2^64) of fxb_state_checksum."""
    words = list(f.state_checksum())
    if world > 1:
        # 64-bit wrap-around sums do not fit a signed all-reduce: add 16-bit limbs, then recombine mod 2^64
        limbs = [(w >> (16 * i)) & 0xffff for w in words for i in range(4)]
        t = torch.tensor(limbs, device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        limbs = [int(v) for v in t.tolist()]
        words = [sum(limbs[4 * k + i] << (16 * i) for i in range(4)) & ((1 << 64) - 1) for k in range(3)]
    return "%016x-%016x-%016x" % tuple(words)


def measure_config(torch, dist, fx, args, grid, rank, world, local_rank, uid, stream, peak, peak_src, sampler=None):
    """Spin-up + warm-up + timed steps of one grid; returns (sim, record)."""
    nx, ny, nz = grid
    voxels = nx * ny * nz
    dt = fx.dt_for_grid(*grid)
    f = make_sim(fx, grid, args, rank, world, local_rank, uid)
    for _ in range(args.spinup):
        f.UpdateFrame(dt); f.Simulate(stream.cuda_stream)
    stream.synchronize()
    if sampler is not None:
        sampler.start()
    ms, st0, st1, phases = timed_run(torch, dist, f, dt, args.steps, args.warmup, world, stream)
    checksum = global_checksum(torch, dist, f, world)  # state after spinup + warmup + steps frames from zero
    if st1.halo_overflow:
        raise SystemExit("advection back-trace left the z-halo: raise h_adv")
    voxels_local = nx * ny * f.slab[1]
    roof, per, jb, proc, cop = rooflines(grid, phases, st0, st1, args.steps, voxels_local, peak, peak_src)
    per_rank = None
    if world > 1:
        # every rank's own phase times and Jacobi work: the slab that holds the plume sets the pace of the step
        mine = torch.tensor([phases.get(k, 0.0) for k in ("advect", "divergence", "jacobi", "gradient")] +
                            [proc / max(args.steps, 1), cop / max(args.steps, 1)], device="cuda", dtype=torch.float64)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"advect_ms": round(float(v[0]), 4), "divergence_ms": round(float(v[1]), 4),
                     "jacobi_ms": round(float(v[2]), 4), "gradient_ms": round(float(v[3]), 4),
                     "bricks_relaxed_per_step": round(float(v[4]), 1), "bricks_copied_per_step": round(float(v[5]), 1)}
                    for v in allr]
        t = torch.tensor([jb, proc, cop], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        jb, proc, cop = (float(v) for v in t.tolist())
    passes = (st1.total_passes - st0.total_passes) / max(args.steps, 1)
    sweeps = (st1.total_sweeps - st0.total_sweeps) / max(args.steps, 1)
    t_step = ms * 1e-3 / args.steps
    work_bytes = 64.0 * voxels + jb
    nominal_bpv = bytes_per_voxel_step(passes, 0.25 if st1.jacobi_fused else 2.0)
    rec = {"value": voxels * args.steps / (ms * 1e-3), "ms_per_step": ms / args.steps, "state_checksum": checksum,
           "frames_from_zero": args.spinup + args.warmup + args.steps,
           "fuse_t": int(st1.fuse_t), "sweeps_per_step": round(sweeps, 2), "jacobi_passes_per_step": round(passes, 2),
           "bytes_per_voxel_step": round(work_bytes / voxels, 2), "jacobi_fused": int(st1.jacobi_fused),
           "bricks_relaxed_per_step": round(proc / args.steps, 1), "bricks_copied_per_step": round(cop / args.steps, 1),
           "brick_cells": int(st1.brick_cells), "kernels_per_step": int(st1.kernels_per_step),
           "step_roofline": {"bound": "hbm", "achieved": round(work_bytes / t_step / 1e9, 1), "peak": peak * world,
                             "unit": "GB/s", "frac": round(work_bytes / t_step / 1e9 / (peak * world), 4),
                             "peak_source": peak_src,
                             "definition": "(64 B x voxels + Jacobi bytes of the bricks actually relaxed/copied) / t_step",
                             "nominal_formula_frac": round(nominal_bpv * voxels / t_step / 1e9 / (peak * world), 4),
                             "nominal_note": "BASELINE.md formula 32+12+ceil(S/T)*12+20 assumes every pass touches "
                                             "every voxel; frozen bricks are skipped here, so it over-counts"},
           "roofline": roof, "phase_roofline": per, "phase_ms": {k: round(v, 4) for k, v in phases.items()}}
    if per_rank is not None:
        rec["per_rank"] = per_rank
    return f, rec


def run_ours(args):
    import torch
    import torch.distributed as dist
    import fluidx12_b200 as fx

    rank, local_rank, world = dist_env()
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: there is no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def fresh_uid():
        """A new ncclUniqueId from rank 0 for every simulator (an id serves one communicator)."""
        if world == 1:
            return None
        import ctypes as C
        buf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (C.c_char * 128)()
            from fluidx12_b200 import binding as B
            B.check(fx.lib().fxb_nccl_unique_id(C.cast(raw, C.c_void_p)))
            buf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
        buf = buf.cuda()
        dist.broadcast(buf, 0)
        return bytes(buf.cpu().numpy().tobytes())

    strong = args.scaling == "strong"
    grid = tuple(args.grid) if args.grid else (STRONG_GRID if strong else WEAK_GRIDS.get(world))
    if grid is None:
        raise SystemExit("--gpus must be 1, 2, 4 or 8 (or pass --grid)")
    nx, ny, nz = grid
    voxels = nx * ny * nz
    dt = fx.dt_for_grid(*grid)
    peak, peak_src = measured_peak_gbs()
    stream = torch.cuda.Stream()

    sampler = ClockSampler(local_rank) if rank == 0 else None
    f, rec = measure_config(torch, dist, fx, args, grid, rank, world, local_rank, fresh_uid(), stream, peak, peak_src, sampler)
    if rank == 0 and world == 1:
        # a very short timed region can end before nvidia-smi has answered once: keep the same load running
        # (untimed) until a few samples exist
        t_end = time.time() + 1.5
        while len(sampler.lines) < 4 and time.time() < t_end:
            for _ in range(5):
                f.UpdateFrame(dt); f.Simulate(stream.cuda_stream)
            stream.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    # End to end on the SAME frames as the device-timed run: the smoke keeps developing (the step of frame 400 costs
    # almost twice the step of frame 130), so a second simulator is spun up identically and the host-timed loop steps
    # frames spinup + warmup .. + steps again; its final state must carry the checksum of the device-timed run.
    f2 = make_sim(fx, grid, args, rank, world, local_rank, fresh_uid() if world > 1 else None)
    for _ in range(args.spinup + args.warmup):
        f2.UpdateFrame(dt); f2.Simulate(stream.cuda_stream)
    stream.synchronize()
    sec_e2e, h2d, d2h = e2e_run(torch, dist, f2, fx, dt, args.steps, world, stream, export=False)
    e2e_checksum = global_checksum(torch, dist, f2, world)
    f2.close()

    line = {
        "metric": METRIC, "value": rec["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": rec["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "3D %dx%dx%d emitter-driven smoke, dt=2/Ny, MIRROR, ITER=64%s, spun up %d steps from "
                               "the zero state; fields (%.0f MiB) far larger than L2, no L2 flush needed" %
                               (nx, ny, nz, " with per-cell early exit" if args.early_exit else ", early exit OFF",
                                args.spinup, voxels * 44 / 2 ** 20),
                   "grid": list(grid), "parallelism": "z-slab x%d" % world,
                   "halo_backend": (args.halo if world > 1 else None), "jacobi_group": args.jacobi_group,
                   "kernel_path": args.kernel_path,
                   **{k: rec[k] for k in ("fuse_t", "sweeps_per_step", "jacobi_passes_per_step", "bytes_per_voxel_step",
                                          "jacobi_fused", "bricks_relaxed_per_step", "bricks_copied_per_step",
                                          "brick_cells")},
                   "switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("FXB_")}},
        "state_checksum": rec["state_checksum"], "frames_from_zero": rec["frames_from_zero"],
        "step_roofline": rec["step_roofline"], "roofline": rec["roofline"], "phase_roofline": rec["phase_roofline"],
        "phase_ms": rec["phase_ms"], "per_rank": rec.get("per_rank"),
        "e2e": {"value": voxels * args.steps / sec_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * sec_e2e / args.steps,
                "what": "UpdateFrame(dt from pinned CB) + Simulate + fxb_get_stats readback, host-timed, on the same "
                        "frames as the device-timed run (second simulator, identical spin-up)",
                "same_state_as_value": e2e_checksum == rec["state_checksum"]},
        "gpu_launches": rec["kernels_per_step"] * args.steps,
        "clocks": clocks,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # CPU baseline on a bounded sample: the 256^3 version of the same workload, developed on the GPU
        cgrid = CPU_SAMPLE_GRID
        g = make_sim(fx, cgrid, args, 0, 1, local_rank, None) if cgrid != grid else f
        cdt = fx.dt_for_grid(*cgrid)
        if g is not f:
            for _ in range(100):
                g.step(cdt)
        line["cpu_baseline"] = cpu_baseline_sample(g, fx, cgrid, cdt)
        if g is not f:
            g.close()
    if args.export_e2e and world == 1:
        # the e2e loop with the renderer hand-off inside it: the colour field copied to pinned host memory every step
        k = min(args.steps, 20)
        sec_exp, _, d2h_exp = e2e_run(torch, dist, f, fx, dt, k, world, stream, export=True)
        line["e2e_export"] = {"value": voxels * k / sec_exp, "unit": UNIT, "d2h_bytes_per_step": d2h_exp,
                              "what": "as e2e plus the colour field copied to pinned host memory every step"}
    f.close()
    # BASELINE configs[2] (256^3: the roofline-characterisation config) and configs[1] (128^3, the reference's default
    # grid, FluidX12.cpp:44) on the same GPU, same method as the main workload
    for key, cn, label in (("c3", 256, "3D 256^3 (BASELINE config 3: roofline characterisation)"),
                           ("c2", 128, "3D 128^3 (BASELINE config 2: the reference's default grid; fits in L2)"),
                           ("c150", 150, "3D 150^3 (the reference's shipped Bin/FluidGI.bat grid: pitched pressure rows)")):
        if not (world == 1 and not args.no_c3 and grid != (cn, cn, cn)):
            continue
        g, crec = measure_config(torch, dist, fx, args, (cn, cn, cn), 0, 1, local_rank, None, stream, peak, peak_src)
        g.close()
        crec["workload"] = label
        line[key] = crec
    # config 4 (512^3 strong scaling) on the same N ranks: its checksum must equal the N = 1 line's state_checksum
    if world > 1 and not strong and not args.grid and not args.no_c4:
        g, crec = measure_config(torch, dist, fx, args, STRONG_GRID, rank, world, local_rank, fresh_uid(), stream, peak, peak_src)
        g.close()
        crec["workload"] = "3D 512^3 (BASELINE config 4) strong-scaled on %d ranks" % world
        line["c4_strong"] = {k: crec[k] for k in ("workload", "value", "ms_per_step", "state_checksum",
                                                  "frames_from_zero", "phase_ms", "sweeps_per_step")}
    if world > 1:
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(line, default=str), flush=True)


# ------------------------------------------------------------------------------------------------
# reference arm: the CPU restatement on the host cores (the reference itself needs Windows + D3D12)
# ------------------------------------------------------------------------------------------------
def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    import oracle
    oracle.build()
    # torch.distributed.run exports OMP_NUM_THREADS=1: ask for every host core explicitly and report what we got
    cores = oracle.threads(os.cpu_count() or 1)
    n = CPU_SAMPLE_GRID[0]
    spin = 100
    o = oracle.FluidOracle(n, n, n)
    dt = oracle.dt_for_grid(n, n, n)
    for _ in range(spin + args.warmup):
        o.step(dt)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        o.step(dt)
    sec = time.perf_counter() - t0
    value = n ** 3 * args.steps / sec
    # BASELINE.json configs[0] and [1] in full (they are small enough for the host): 2D 256^2 (Fluid2D.bat setup) and
    # 3D 128^3, emitter-driven from the zero state, timed after a short spin-up
    small = {}
    for key, g, spin_s, steps_s in (("c1_2d_256x256", (256, 256, 1), 50, 200), ("c2_3d_128_cubed", (128, 128, 128), 30, 20)):
        try:
            oc = oracle.FluidOracle(*g)
            cdt = oracle.dt_for_grid(*g)
            for _ in range(spin_s):
                oc.step(cdt)
            tc = time.perf_counter()
            for _ in range(steps_s):
                oc.step(cdt)
            tc = time.perf_counter() - tc
            small[key] = {"grid": list(g), "steps": steps_s, "spinup": spin_s, "ms_per_step": round(1e3 * tc / steps_s, 4),
                          "value": g[0] * g[1] * g[2] * steps_s / tc, "unit": UNIT, "s_exec_last": int(oc.s_exec)}
            oc.close()
        except Exception as e:
            small[key] = {"error": repr(e)[:200]}
    strong = args.scaling == "strong"
    grid = tuple(args.grid) if args.grid else (STRONG_GRID if strong else WEAK_GRIDS.get(world, WEAK_GRIDS[1]))
    sample = ("OpenMP C++ restatement of the reference HLSL (the reference needs Windows/D3D12 and cannot run "
              "here), %d threads (omp_get_max_threads), %d^3 sample of the workload, spun up %d steps" % (cores, n, spin))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True,
        "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "3D %dx%dx%d emitter-driven smoke (timed on a %d^3 sample), dt=2/Ny, MIRROR, ITER=64 "
                               "with per-cell early exit" % (grid + (n,)), "grid": list(grid),
                   "sample_grid": [n, n, n]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "baseline_configs": small,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, nargs=3, default=None)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N > 1: weak = 134 M voxels per GPU (default); strong = 512^3 at every N (BASELINE config 4)")
    ap.add_argument("--spinup", type=int, default=100)
    ap.add_argument("--fuse-t", type=int, default=0)
    ap.add_argument("--early-exit", type=int, default=1)
    ap.add_argument("--kernel-path", type=int, default=0)
    ap.add_argument("--halo", default="fused", choices=["fused", "peer", "nccl"],
                    help="N > 1: how slab-face halos travel (fused: stored by the kernels themselves; peer / nccl: exchanges)")
    ap.add_argument("--jacobi-group", type=int, default=0, help="N > 1: fused passes per pressure-halo exchange (0 = default)")
    ap.add_argument("--export-e2e", action="store_true", help="also time the e2e variant that copies the colour field out")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c3", action="store_true")
    ap.add_argument("--no-c4", action="store_true", help="N > 1: skip the 512^3 strong-scaling sub-run")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
