#!/usr/bin/env python
"""bench.py — voxel-updates/s of the smoke-solver step (advect + project) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our CUDA path (one JSON line)
    python bench.py --impl reference [--gpus N] --steps K --warmup W   # the CPU oracle on the host cores

A "step" is one full frame of the hot path: Fluid::UpdateFrame + Fluid::Simulate (CSAdvect, then
divergence, <=64 Jacobi sweeps and gradient-subtract of CSProject3D) over the whole grid.
Workload (BASELINE.json): synthetic emitter-driven smoke from the all-zero state, dt = 2/Ny, MIRROR
addressing, ITER = 64 with the per-cell early exit.  The state is first spun up for --spinup steps
(untimed state preparation: at step 0 nothing moves and the solver would be trivially cheap), then
W warm-up steps, then exactly K timed steps.

Grid: N = 1 -> 512^3 (BASELINE config "3D 512^3 ... at 1/2/4/8 B200", 1-GPU point; every field is far
larger than L2).  N > 1 -> weak scaling at 134 M voxels per GPU, z-slab decomposed: 512x512x1024 (2),
1024x1024x512 (4), 1024^3 (8, BASELINE config 5).  At N = 1 the line also carries "c3" (the 256^3
roofline-characterisation config) and "c2" (128^3), two further host-timed loops ("e2e_pipelined": one frame in
flight; "e2e_export": the colour field copied out every step) and "experiments": the opt-in variants of the step, the
light-map pass and the ray march, timed in a child process with a hard time limit AFTER everything above (--no-experiments
or FXB_BENCH_EXPERIMENTS=0 skips it; use that under ncu).  --grid overrides.  `--impl reference` also steps BASELINE's two small configs in full.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WEAK_GRIDS = {1: (512, 512, 512), 2: (512, 512, 1024), 4: (1024, 1024, 512), 8: (1024, 1024, 1024)}
METRIC = "voxel_updates_per_s"
UNIT = "voxel-updates/s"


def measured_peak_gbs():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def bytes_per_voxel_step(passes: float, mask_bytes_per_pass: float) -> float:
    """Algorithmic HBM bytes per voxel per step (SURVEY.md §8d / BASELINE.md §3): advect 32 + divergence 12
    + per executed Jacobi pass 12 (+ freeze mask) + gradient-subtract 20."""
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
                kernel_path=args.kernel_path, nccl_unique_id=uid)
    if not ok:
        raise RuntimeError("fluidx_b200 Init failed: " + f.last_error)
    return f


def timed_run(torch, dist, f, dt, steps, warmup, world, stream):
    """W warm-up + K timed steps, device-timed with CUDA events on the launching stream; max over ranks."""
    for _ in range(warmup):
        f.UpdateFrame(dt); f.Simulate(stream.cuda_stream)
    stream.synchronize()
    st0 = f.stats()
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
    return ms, st0, st1


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


def e2e_pipelined_run(torch, f, dt, steps, stream):
    """The frame loop with one frame in flight, as the reference keeps its own (FrameCount = 3, Fluid.h:35): per step the
    frame constants go in, the step is enqueued, a snapshot of its result record is posted (fxb_post_stats: async copy
    into pinned memory + event) and the host then waits for the PREVIOUS step's record — every step's record is read,
    each exactly once, and the device never idles."""
    cb = torch.zeros(2, dtype=torch.float32).pin_memory()
    cb[0] = dt
    torch.cuda.synchronize()
    read = 0
    t0 = time.perf_counter()
    for k in range(steps):
        f.UpdateFrame(float(cb[0]))
        f.Simulate(stream.cuda_stream)
        f.post_stats(k & 1)
        if k > 0:
            read += int(f.wait_stats((k - 1) & 1).s_exec >= 0)
    read += int(f.wait_stats((steps - 1) & 1).s_exec >= 0)
    torch.cuda.synchronize()
    sec = time.perf_counter() - t0
    assert read == steps
    return sec


def voxels_local_of(f):
    return f.m_gridSize[0] * f.m_gridSize[1] * f.slab[1]


def jacobi_work_bytes(st0, st1, mask_bytes, voxels_local):
    """Bytes the Jacobi passes between two stats snapshots really had to move: a relaxed brick reads p + rhs and
    writes p (+ 2/8 B of freeze flags) per cell, a frozen brick is copied once (p in, p out), every other brick is
    skipped because its value is already final in both pressure buffers."""
    if st1.jacobi_fused:
        proc = st1.bricks_processed - st0.bricks_processed
        cop = st1.bricks_copied - st0.bricks_copied
        return (proc * 12.25 + cop * 8.0) * st1.brick_cells, proc, cop
    return (st1.total_passes - st0.total_passes) * (12.0 + mask_bytes) * voxels_local, 0, 0


def phase_rooflines(f, dt, voxels_local, peak, peak_src, reps, mask_bytes):
    """Per-phase device time measured live with CUDA events between the kernels of un-graphed steps
    (fxb_profile_step) against each phase's algorithmic bytes.  The Jacobi passes are the dominant
    kernel: their bytes are counted from the bricks actually relaxed / copied during these very steps."""
    st0 = f.stats()
    phases = {}
    for _ in range(reps):
        f.UpdateFrame(dt)
        for k, v in f.profile_step().items():
            phases[k] = phases.get(k, 0.0) + v / reps
    st1 = f.stats()
    jb, proc, cop = jacobi_work_bytes(st0, st1, mask_bytes, voxels_local)
    jb /= reps
    launches = (st1.total_passes - st0.total_passes) / reps  # one kernel per executed pass
    per = {"advect": 32.0 * voxels_local, "divergence": 12.0 * voxels_local, "jacobi": jb,
           "gradient": 20.0 * voxels_local}
    out = {}
    for k, nbytes in per.items():
        gbs = nbytes / max(phases[k] * 1e-3, 1e-12) / 1e9
        out[k] = {"ms": round(phases[k], 4), "algorithmic_bytes": nbytes, "achieved_gbs": round(gbs, 1),
                  "frac": round(gbs / peak, 4)}
    j = out["jacobi"]
    roof = {"bound": "hbm", "kernel": "jacobi_pass_kernel (average over the executed passes of a step; the one-time "
                                      "copies of frozen bricks run inside it)",
            "achieved": j["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": j["frac"], "peak_source": peak_src,
            "traffic": None, "bytes_per_launch": round(jb / max(launches, 1e-9)),
            "avg_launch_ms": round(j["ms"] / max(launches, 1e-9), 5), "bytes_per_step": jb, "ms_per_step": j["ms"],
            "launches_per_step": round(launches, 1),
            "bricks_relaxed_per_step": round(proc / reps, 1), "bricks_copied_per_step": round(cop / reps, 1),
            "note": "issue/latency-bound, not HBM-bound: see DESIGN.md §5 and profiles/"}
    return roof, out, {k: round(v, 4) for k, v in phases.items()}


def cpu_baseline_sample(f, fx, grid, dt, budget_s=25.0):
    """Times the OpenMP oracle on the host cores on a bounded sample of the same workload: the developed GPU
    state is copied into the oracle (whole grid when it is small enough, else not run at this size) and
    stepped for a few frames; the result is also compared with the GPU (full-size parity spot check)."""
    import numpy as np
    import oracle
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
    cores = os.cpu_count() or 1
    return {"value": nx * ny * nz * n_steps / t_total, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d oracle steps of the same %dx%dx%d developed state (OpenMP, %d threads, %.1f s)" %
                      (n_steps, nx, ny, nz, cores, t_total),
            "parity_max_abs_rel_vs_gpu": worst}



# ------------------------------------------------------------------------------------------------
# experiments: after the measurement above, the opt-in variants of the same bit-exact step are timed in CHILD
# processes (own process group, hard time limit), so that every bench run also says what they would do.  Nothing
# here touches `value` / `e2e` / `roofline`: those were taken before, on the default path.
# ------------------------------------------------------------------------------------------------
def run_child(cmd, env, timeout_s):
    import signal
    p = subprocess.Popen(cmd, env=env, cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         start_new_session=True)
    try:
        out, _ = p.communicate(timeout=timeout_s)
        return p.returncode, out
    except subprocess.TimeoutExpired:
        try:
            os.killpg(p.pid, signal.SIGKILL)  # exactly the group started above
        except ProcessLookupError:
            pass
        try:
            out, _ = p.communicate(timeout=20)
        except subprocess.TimeoutExpired:  # a process stuck in the driver cannot be reaped: do not wait for it
            out = ""
        return None, out


def child_env(extra):
    drop = ("RANK", "LOCAL_RANK", "WORLD_SIZE", "LOCAL_WORLD_SIZE", "GROUP_RANK", "GROUP_WORLD_SIZE", "ROLE_RANK",
            "ROLE_WORLD_SIZE", "ROLE_NAME", "MASTER_ADDR", "MASTER_PORT")
    env = {k: v for k, v in os.environ.items() if k not in drop and not k.startswith("TORCHELASTIC_")}
    env.update({k: str(v) for k, v in extra.items()})
    return env


def experiments_single_gpu(budget_s):
    """tools/gpu_shot.py --bench: 256^3 and 512^3, state copied from a 100-step spin-up of the default schedule into
    each variant, host-timed graph-launched steps, every field compared bit for bit with the default schedule's."""
    path = os.path.join(ROOT, "gpurun_out", "bench_experiments.jsonl")
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if os.path.exists(path):
        os.remove(path)
    t0 = time.time()
    rc, out = run_child([sys.executable, os.path.join("tools", "gpu_shot.py"), "--bench"],
                        child_env({"FXB_SHOT_OUT": path}), budget_s)
    rows = []
    try:
        for ln in open(path):
            r = json.loads(ln)
            if r.get("stage") == "light_map":
                r.pop("stage"); r.pop("t", None)
                r["grid"] = "x".join(map(str, r["grid"]))
                r["variant"] = "light_map_pass"
                rows.append(r)
                continue
            if r.get("stage") != "timing":
                continue
            row = {"grid": "x".join(map(str, r["grid"])), "variant": r.get("variant", "default")}
            if "error" in r:
                row["error"] = r["error"][:200]
            elif "variant" in r:
                row.update(ms_per_step=r["ms"], jacobi_ms=r["phases"].get("jacobi"), advect_ms=r["phases"].get("advect"),
                           mismatched_elements_vs_default=sum(r["mismatch_vs_default"].values()),
                           tail_launches=r["tail"].get("tail_launches_last_step"))
            else:
                row.update(ms_per_step=r["default"], jacobi_ms=r["default_phases"].get("jacobi"),
                           advect_ms=r["default_phases"].get("advect"))
            rows.append(row)
    except Exception as e:  # the child wrote nothing usable
        rows.append({"error": repr(e)[:200]})
    return {"what": "opt-in variants (environment switches, DESIGN.md §5) timed in a child process after the measurement; "
                    "host-timed graph launches, same developed state for every variant; not the headline path",
            "seconds": round(time.time() - t0, 1), "exit": "timeout" if rc is None else rc, "results": rows,
            "child_tail": out[-400:] if rc not in (0,) else ""}


def experiments_multi_gpu(args, world, budget_s):
    """This same bench (short, without its extras) relaunched under torchrun with the multi-GPU switches; the state
    checksum of each variant must equal the default variant's (same number of steps from the same zero state)."""
    port = int(os.environ.get("MASTER_PORT", "29500"))
    t_end = time.time() + budget_s
    rows, ref = [], None
    for i, (label, extra) in enumerate((("default", {}), ("p2p_halos", {"FXB_P2P": 1, "FXB_P2P_TIMEOUT_S": 15}),
                                        ("tail_p2p", {"FXB_TAIL": 1, "FXB_P2P": 1, "FXB_P2P_TIMEOUT_S": 15}),
                                        ("tail", {"FXB_TAIL": 1}))):
        left = t_end - time.time()
        if left < 20:
            rows.append({"variant": label, "skipped": "time budget"})
            continue
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(port + 101 + i), os.path.join(ROOT, "bench.py"),
               "--gpus", str(world), "--steps", "20", "--warmup", "3", "--no-cpu-baseline", "--no-c3",
               "--no-experiments", "--checksum"]
        if args.grid:
            cmd += ["--grid"] + [str(v) for v in args.grid]
        rc, out = run_child(cmd, child_env(extra), min(left, 90.0))
        row = {"variant": label}
        got = None
        for ln in out.splitlines():
            if "{" in ln and '"metric"' in ln:  # torchrun may prefix a worker's output
                try:
                    got = json.loads(ln[ln.index("{"):])
                except Exception:
                    pass
        if got is None:
            row.update(error="timeout" if rc is None else "exit %s" % rc, child_tail=out[-300:])
        else:
            row.update(ms_per_step=round(got["ms_per_step"], 4), value=got["value"],
                       halo_ms=got.get("phase_ms", {}).get("halo"), jacobi_ms=got.get("phase_ms", {}).get("jacobi"))
            if label == "default":
                ref = got.get("state_checksum")
            row["state_equals_default_variant"] = (got.get("state_checksum") == ref) if ref is not None else None
        rows.append(row)
    return {"what": "multi-GPU opt-in variants (DESIGN.md §6): this bench relaunched in child torchrun jobs, 20 steps, "
                    "after the measurement; not the headline path", "results": rows}


def state_checksum(torch, dist, f, fx, world):
    """Order-sensitive 62-bit checksum of the rank's velocity.xyz, colour and pressure words, summed over ranks."""
    import numpy as np
    z0 = f.slab[0]
    total = 0
    for k, fld in enumerate((fx.FIELD_VELOCITY, fx.FIELD_COLOR, fx.FIELD_PRESSURE)):
        a = f.get_field(fld)
        w = a.view(np.uint16) if a.dtype == np.float16 else a.view(np.uint32)
        if fld == fx.FIELD_VELOCITY:
            w = w[..., :3]
        per_plane = np.add.reduce(w.reshape(w.shape[0], -1), axis=1, dtype=np.uint64)
        for j, v in enumerate(per_plane.tolist()):
            total = (total + (k + 1) * (z0 + j + 1) * (v % (1 << 40))) % (1 << 62)
    if world > 1:
        t = torch.tensor([total >> 31, total & ((1 << 31) - 1)], device="cuda", dtype=torch.int64)
        dist.all_reduce(t)
        total = ((int(t[0].item()) << 31) + int(t[1].item())) % (1 << 62)
    return total


def run_ours(args):
    t_bench0 = time.time()
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
    uid = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        import ctypes as C
        buf = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            raw = (C.c_char * 128)()
            from fluidx12_b200 import binding as B
            B.check(fx.lib().fxb_nccl_unique_id(C.cast(raw, C.c_void_p)))
            buf = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
        buf = buf.cuda()
        dist.broadcast(buf, 0)
        uid = bytes(buf.cpu().numpy().tobytes())

    grid = tuple(args.grid) if args.grid else WEAK_GRIDS.get(world)
    if grid is None:
        raise SystemExit("--gpus must be 1, 2, 4 or 8 (or pass --grid)")
    nx, ny, nz = grid
    voxels = nx * ny * nz
    dt = fx.dt_for_grid(*grid)
    peak, peak_src = measured_peak_gbs()
    stream = torch.cuda.Stream()

    f = make_sim(fx, grid, args, rank, world, local_rank, uid)
    for _ in range(args.spinup):
        f.UpdateFrame(dt); f.Simulate(stream.cuda_stream)
    stream.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, st0, st1 = timed_run(torch, dist, f, dt, args.steps, args.warmup, world, stream)
    if rank == 0 and world == 1:
        # a very short timed region can end before nvidia-smi has answered once: keep the same load running
        # (untimed) until a few samples exist
        t_end = time.time() + 1.5
        while len(sampler.lines) < 4 and time.time() < t_end:
            for _ in range(5):
                f.UpdateFrame(dt); f.Simulate(stream.cuda_stream)
            stream.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    sec_e2e, h2d, d2h = e2e_run(torch, dist, f, fx, dt, args.steps, world, stream, export=False)

    passes = (st1.total_passes - st0.total_passes) / max(args.steps, 1)
    sweeps = (st1.total_sweeps - st0.total_sweeps) / max(args.steps, 1)
    fuse_t = st1.fuse_t
    mask_bytes = 0.25 if st1.jacobi_fused else 2.0
    voxels_local = nx * ny * f.slab[1]
    value = voxels * args.steps / (ms * 1e-3)
    t_step = ms * 1e-3 / args.steps
    # bytes the step really needs (whole job): advect 32 + divergence 12 + gradient 20 per voxel + the Jacobi work
    jb, proc, cop = jacobi_work_bytes(st0, st1, mask_bytes, voxels_local)
    if world > 1:
        t = torch.tensor([jb, proc, cop], device="cuda", dtype=torch.float64)
        dist.all_reduce(t)
        jb, proc, cop = (float(v) for v in t.tolist())
    work_bytes = 64.0 * voxels + jb / args.steps
    work_gbs = work_bytes / t_step / 1e9
    nominal_bpv = bytes_per_voxel_step(passes, mask_bytes)
    nominal_gbs = nominal_bpv * voxels / t_step / 1e9

    roof, phase_roof, phases = phase_rooflines(f, dt, voxels_local, peak, peak_src, 5, mask_bytes)
    # `traffic` stays null: the ncu capture in profiles/ is of pass 0 alone (every brick relaxed), while `achieved`
    # averages over all passes of a step; the capture is quoted next to it instead.
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roof["ncu_pass0_capture"] = json.load(open(prof)).get("%dx%dx%d" % grid, {}).get("jacobi")
        except Exception:
            pass
    if st1.halo_overflow:
        raise SystemExit("advection back-trace left the z-halo: raise h_adv")

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "3D %dx%dx%d emitter-driven smoke, dt=2/Ny, MIRROR, ITER=64%s, spun up %d steps from "
                               "the zero state; fields (%.0f MiB) far larger than L2, no L2 flush needed" %
                               (nx, ny, nz, " with per-cell early exit" if args.early_exit else ", early exit OFF",
                                args.spinup, voxels * 44 / 2 ** 20),
                   "grid": list(grid), "parallelism": "z-slab x%d" % world, "fuse_t": fuse_t,
                   "sweeps_per_step": round(sweeps, 2), "jacobi_passes_per_step": round(passes, 2),
                   "bytes_per_voxel_step": round(work_bytes / voxels, 2), "kernel_path": args.kernel_path,
                   "jacobi_fused": int(st1.jacobi_fused), "bricks_relaxed_per_step": round(proc / args.steps, 1),
                   "bricks_copied_per_step": round(cop / args.steps, 1), "brick_cells": int(st1.brick_cells),
                   # opt-in variants of the same bit-exact step that were switched on through the environment
                   "switches": {k: v for k, v in sorted(os.environ.items()) if k.startswith("FXB_")}},
        "step_roofline": {"bound": "hbm", "achieved": round(work_gbs, 1), "peak": peak * world, "unit": "GB/s",
                          "frac": round(work_gbs / (peak * world), 4), "peak_source": peak_src,
                          "definition": "(64 B x voxels + Jacobi bytes of the bricks actually relaxed/copied) / t_step",
                          "nominal_bytes_step_formula": {
                              "bytes_per_voxel": round(nominal_bpv, 2), "achieved": round(nominal_gbs, 1),
                              "frac": round(nominal_gbs / (peak * world), 4),
                              "note": "BASELINE.md formula 32+12+ceil(S/T)*12+20: assumes every pass touches every "
                                      "voxel; frozen bricks are skipped here, so it over-counts the bytes moved"}},
        "roofline": roof,
        "phase_roofline": phase_roof,
        "phase_ms": phases,
        "e2e": {"value": voxels * args.steps / sec_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "ms_per_step": 1e3 * sec_e2e / args.steps,
                "what": "UpdateFrame(dt from pinned CB) + Simulate + fxb_get_stats readback, host-timed"},
        "gpu_launches": st1.kernels_per_step * args.steps,
        "clocks": clocks,
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # CPU baseline on a bounded sample: the 256^3 (or smaller) version of the same workload.
        cgrid = grid if voxels <= 256 ** 3 else (256, 256, 256)
        g = make_sim(fx, cgrid, args, 0, 1, local_rank, None) if cgrid != grid else f
        cdt = fx.dt_for_grid(*cgrid)
        if g is not f:
            for _ in range(args.spinup):
                g.step(cdt)
        line["cpu_baseline"] = cpu_baseline_sample(g, fx, cgrid, cdt)
        if g is not f:
            g.close()
    # BASELINE configs[2] (256^3: the roofline-characterisation config) and configs[1] (128^3, the reference's default
    # grid, FluidX12.cpp:44) on the same GPU, same method as the main workload
    for key, cn, label in (("c3", 256, "3D 256^3 (BASELINE config 3: roofline characterisation)"),
                           ("c2", 128, "3D 128^3 (BASELINE config 2: the reference's default grid; fits in L2)")):
        if not (rank == 0 and world == 1 and not args.no_c3 and grid != (cn, cn, cn)):
            continue
        g = make_sim(fx, (cn, cn, cn), args, 0, 1, local_rank, None)
        cdt = fx.dt_for_grid(cn, cn, cn)
        for _ in range(args.spinup):
            g.UpdateFrame(cdt); g.Simulate(stream.cuda_stream)
        cms, c0, c1 = timed_run(torch, dist, g, cdt, args.steps, args.warmup, 1, stream)
        cv = cn ** 3
        cjb, cproc, ccop = jacobi_work_bytes(c0, c1, mask_bytes, cv)
        cwork = 64.0 * cv + cjb / args.steps
        cnom = bytes_per_voxel_step((c1.total_passes - c0.total_passes) / args.steps, mask_bytes)
        croof, cphase_roof, cph = phase_rooflines(g, cdt, cv, peak, peak_src, 5, mask_bytes)
        ct = cms * 1e-3 / args.steps
        line[key] = {"workload": label,
                     "value": cv * args.steps / (cms * 1e-3), "ms_per_step": cms / args.steps,
                     "jacobi_passes_per_step": round((c1.total_passes - c0.total_passes) / args.steps, 2),
                     "sweeps_per_step": round((c1.total_sweeps - c0.total_sweeps) / args.steps, 2),
                     "bytes_per_voxel_step": round(cwork / cv, 2),
                     "step_roofline_frac": round(cwork / ct / 1e9 / peak, 4),
                     "nominal_formula_frac": round(cnom * cv / ct / 1e9 / peak, 4),
                     "roofline": croof, "phase_roofline": cphase_roof, "phase_ms": cph}
        g.close()
    if not args.no_export_e2e and world == 1:
        # extras, after everything the contract needs has been measured.  First the e2e loop with the renderer
        # hand-off inside it: the colour field copied to pinned host memory every step
        try:
            k = min(args.steps, 20)
            sec_exp, _, d2h_exp = e2e_run(torch, dist, f, fx, dt, k, world, stream, export=True)
            line["e2e_export"] = {"value": voxels * k / sec_exp, "unit": UNIT, "d2h_bytes_per_step": d2h_exp,
                                  "what": "as e2e plus the colour field copied to pinned host memory every step"}
        except Exception as e:
            line["e2e_export"] = {"error": repr(e)[:200]}
    if world == 1:
        # then the frame loop with one frame in flight
        try:
            import ctypes as C
            sec_pipe = e2e_pipelined_run(torch, f, dt, args.steps, stream)
            line["e2e_pipelined"] = {"value": voxels * args.steps / sec_pipe, "unit": UNIT, "h2d_bytes_per_step": 8,
                                     "d2h_bytes_per_step": C.sizeof(fx.FxbStats), "ms_per_step": 1e3 * sec_pipe / args.steps,
                                     "what": "as e2e, but the host waits for the PREVIOUS step's result record (posted "
                                             "asynchronously into pinned memory) while the current step runs: one "
                                             "frame in flight, every record still read every step"}
        except Exception as e:  # an extra: the line is complete without it
            line["e2e_pipelined"] = {"error": repr(e)[:200]}
    if args.checksum:
        line["state_checksum"] = state_checksum(torch, dist, f, fx, world)
    try:
        f.close()
    except Exception:
        pass
    if world > 1:
        dist.destroy_process_group()
    # Single GPU: on by default (every variant's kernels are plain compute kernels without device-side waits; a fault
    # in the child ends the child).  Several GPUs: the peer-memory halo exchange has never run on GPUs, and a
    # misbehaving multi-rank child must not be able to disturb scaling runs that follow — so it is opt-in
    # (--experiments-multi) except at 8 GPUs, the last point of a 1/2/4/8 scaling sequence, where it is on by default.
    want_exp = (not args.no_experiments and os.environ.get("FXB_BENCH_EXPERIMENTS", "1") != "0" and rank == 0
                and time.time() - t_bench0 < 300.0 and (world in (1, 8) or args.experiments_multi))
    if want_exp:
        # every rank has released its GPU memory and its communicators; ranks > 0 simply exit
        try:
            line["experiments"] = (experiments_single_gpu(args.experiments_budget) if world == 1 else
                                   experiments_multi_gpu(args, world, 2.0 * args.experiments_budget))
        except Exception as e:
            line["experiments"] = {"error": repr(e)[:300]}
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
    cores = os.cpu_count() or 1
    total = args.steps + args.warmup
    # Bounded sample: the same emitter-driven workload on the largest cubic grid whose K+W steps fit ~150 s.
    probe = oracle.FluidOracle(64, 64, 64)
    pdt = oracle.dt_for_grid(64, 64, 64)
    for _ in range(20):
        probe.step(pdt)
    t0 = time.perf_counter()
    for _ in range(5):
        probe.step(pdt)
    per_voxel_step = (time.perf_counter() - t0) / 5 / 64 ** 3
    n = 64
    for cand in (512, 384, 256, 192, 128, 96):
        spin = min(args.spinup, 100)
        if per_voxel_step * cand ** 3 * (total + spin) * 1.5 <= 150.0:
            n = cand
            break
    o = oracle.FluidOracle(n, n, n)
    dt = oracle.dt_for_grid(n, n, n)
    spin = min(args.spinup, 100)
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
    grid = tuple(args.grid) if args.grid else WEAK_GRIDS.get(world, WEAK_GRIDS[1])
    sample = ("OpenMP C++ restatement of the reference HLSL (the reference needs Windows/D3D12 and cannot run "
              "here), %d threads, %d^3 sample of the workload, spun up %d steps" % (cores, n, spin))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps, "higher_is_better": True, "scaling": "weak",
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
    ap.add_argument("--spinup", type=int, default=100)
    ap.add_argument("--fuse-t", type=int, default=0)
    ap.add_argument("--early-exit", type=int, default=1)
    ap.add_argument("--kernel-path", type=int, default=0)
    ap.add_argument("--no-export-e2e", action="store_true", help="skip the e2e variant that also copies the colour field out")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c3", action="store_true")
    ap.add_argument("--no-experiments", action="store_true",
                    help="skip the child-process runs of the opt-in variants after the measurement (use under ncu)")
    ap.add_argument("--experiments-multi", action="store_true",
                    help="N > 1: also relaunch this bench with the multi-GPU switches (peer-memory halos, dynamic schedule)")
    ap.add_argument("--experiments-budget", type=float, default=75.0, help="seconds (twice that for N > 1)")
    ap.add_argument("--checksum", action="store_true", help="add a checksum of the final state to the line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
