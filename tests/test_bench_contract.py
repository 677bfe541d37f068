"""bench.py's whole `ours` flow on a fake device: the JSON line it prints must carry every key of the bench contract
(metric, value, e2e, roofline, cpu-side extras, clocks, gpu_launches, c3, state checksum) — a typo in the code that runs
after the timed region would otherwise only show on the GPU box.  Nothing numerical is checked here."""
import contextlib
import json
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


class FakeStats:
    def __init__(self, steps):
        self.s_exec, self.jacobi_passes, self.fuse_t, self.halo_overflow, self.frame_parity = 57, 29, 2, 0, steps & 1
        self.kernels_per_step, self.steps = 38, steps
        self.total_sweeps, self.total_passes = 57 * steps, 29 * steps
        self.bricks_processed, self.bricks_copied = 200 * steps, 100 * steps
        self.brick_cells, self.bricks_per_pass, self.jacobi_fused = 11520, 64, 1


class FakeFluid:
    def Init(self, gridSize=(128, 128, 128), **kw):
        self.m_gridSize, self.steps, self.last_error = tuple(gridSize), 0, ""
        return True

    slab = property(lambda self: (0, self.m_gridSize[2]))

    def UpdateFrame(self, dt):
        pass

    def Simulate(self, stream=None):
        self.steps += 1

    def step(self, dt):
        self.steps += 1

    def sync(self):
        pass

    def stats(self):
        return FakeStats(self.steps)

    def profile_step(self):
        self.steps += 1
        return {"advect": 1.9, "divergence": 0.4, "jacobi": 2.9, "gradient": 0.5, "halo": 0.0, "step": 5.7}

    def phase_times(self, reset=False):
        return {"advect": 1.9 * self.steps, "divergence": 0.4 * self.steps, "jacobi": 2.9 * self.steps,
                "gradient": 0.5 * self.steps}

    def state_checksum(self):
        return (1, 2, 3)

    def get_field_async(self, field, ptr, nbytes, stream=None):
        assert nbytes == self.m_gridSize[0] * self.m_gridSize[1] * self.m_gridSize[2] * 8

    def close(self):
        pass


class FakeEvent:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 12.5


class FakeStream:
    cuda_stream = 0

    def synchronize(self):
        pass


def test_ours_prints_one_line_with_every_contract_key(monkeypatch, capsys):
    import ctypes as C

    import torch
    fake_fx = types.SimpleNamespace(Fluid=FakeFluid, ADDRESS_MIRROR=0, FIELD_COLOR=1, HALO_PEER=0, HALO_NCCL=1, HALO_FUSED=2,
                                    dt_for_grid=lambda *g: 2.0 / g[1],
                                    FxbStats=type("S", (C.Structure,), {"_fields_": [("x", C.c_int * 24)]}))
    monkeypatch.setitem(sys.modules, "fluidx12_b200", fake_fx)
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "Stream", FakeStream)
    monkeypatch.setattr(torch.cuda, "Event", FakeEvent)
    monkeypatch.setattr(torch.cuda, "stream", lambda s: contextlib.nullcontext())
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE"):
        monkeypatch.delenv(k, raising=False)

    monkeypatch.setattr(sys, "argv", ["bench.py", "--steps", "4", "--warmup", "3", "--spinup", "2", "--grid", "32", "32", "32",
                                      "--no-cpu-baseline", "--export-e2e"])
    bench.main()
    out = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("{")]
    assert len(out) == 1
    line = json.loads(out[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "e2e", "gpu_launches", "clocks", "c3", "c2",
                "e2e_export", "phase_roofline", "step_roofline", "state_checksum", "phase_ms"):
        assert key in line, key
    assert line["metric"] == "voxel_updates_per_s" and line["n_gpus"] == 1 and line["steps"] == 4 and line["warmup"] == 3
    assert line["config"]["workload"].startswith("3D 32x32x32") and line["vs_baseline"] is None
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(line["roofline"])
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(line["e2e"])
    assert line["e2e_export"]["d2h_bytes_per_step"] > 32 * 32 * 32 * 8
    assert line["gpu_launches"] == 38 * 4
    assert line["roofline"]["kernel"] == "jacobi_pass_kernel+jacobi_resident_kernel" and line["state_checksum"].count("-") == 2
    assert "experiments" not in line
    assert np.isclose(line["value"], 32 ** 3 * 4 / 12.5e-3)
