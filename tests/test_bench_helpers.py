"""bench.py's host-side helpers that need no GPU: the child-process runner of the experiments stage (own process
group, hard time limit), the parsing of what the children report, and the state checksum."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_run_child_returns_output_and_kills_its_group_on_timeout(tmp_path):
    rc, out = bench.run_child([sys.executable, "-c", "print('hello')"], dict(os.environ), 30)
    assert rc == 0 and "hello" in out
    # a child that starts a grandchild and hangs: both must be gone after the limit
    marker = tmp_path / "alive"
    grandchild = tmp_path / "grandchild.py"
    grandchild.write_text("import time\nwhile True:\n    open(%r, 'w').write(str(time.time()))\n    time.sleep(0.05)\n" % str(marker))
    child = tmp_path / "child.py"
    child.write_text("import subprocess, sys, time\nsubprocess.Popen([sys.executable, %r])\ntime.sleep(60)\n" % str(grandchild))
    t0 = time.time()
    rc, out = bench.run_child([sys.executable, str(child)], dict(os.environ), 1.5)
    assert rc is None and time.time() - t0 < 10
    assert marker.exists(), "the grandchild never ran"
    time.sleep(0.3)
    stamp = marker.read_text()
    time.sleep(0.4)
    assert marker.read_text() == stamp, "the grandchild is still running"


def test_child_env_drops_the_launcher_variables(monkeypatch):
    monkeypatch.setenv("RANK", "3")
    monkeypatch.setenv("MASTER_PORT", "1234")
    monkeypatch.setenv("TORCHELASTIC_RUN_ID", "x")
    monkeypatch.setenv("FXB_KEEP", "1")
    env = bench.child_env({"FXB_P2P": 1})
    assert "RANK" not in env and "MASTER_PORT" not in env and "TORCHELASTIC_RUN_ID" not in env
    assert env["FXB_KEEP"] == "1" and env["FXB_P2P"] == "1"


def test_single_gpu_experiments_are_summarised_from_the_childs_records(monkeypatch):
    rows = [
        {"stage": "import"},
        {"stage": "timing", "grid": [256, 256, 256], "default": 1.49, "default_phases": {"jacobi": 1.18, "advect": 0.29}},
        {"stage": "timing", "grid": [256, 256, 256], "variant": "tail", "ms": 1.40, "phases": {"jacobi": 1.0, "advect": 0.29},
         "tail": {"tail_launches_last_step": 13}, "passes": 17, "s_exec": 64,
         "mismatch_vs_default": {"0": 0, "1": 0, "2": 0}},
        {"stage": "timing", "grid": [256, 256, 256], "variant": "broken", "error": "RuntimeError('x')"},
        {"stage": "light_map", "grid": [256, 256, 256], "probes": 1, "ms": 2.5, "voxels_with_smoke": 123, "t": 9.1},
        {"stage": "light_map", "grid": [256, 256, 256], "pass_": "ray_march_v", "cube_size": 256, "ms": 0.4, "t": 9.3},
        {"stage": "done"},
    ]

    def fake(cmd, env, timeout_s):
        assert cmd[-1] == "--bench" and "FXB_SHOT_OUT" in env
        with open(env["FXB_SHOT_OUT"], "w") as fh:
            fh.write("".join(json.dumps(r) + "\n" for r in rows))
        return 0, ""

    monkeypatch.setattr(bench, "run_child", fake)
    res = bench.experiments_single_gpu(5)
    assert res["exit"] == 0 and len(res["results"]) == 5
    d, t, b, lm, rm = res["results"]
    assert lm == {"grid": "256x256x256", "probes": 1, "ms": 2.5, "voxels_with_smoke": 123, "variant": "light_map_pass"}
    assert rm["pass_"] == "ray_march_v" and rm["variant"] == "light_map_pass" and "t" not in rm
    assert d == {"grid": "256x256x256", "variant": "default", "ms_per_step": 1.49, "jacobi_ms": 1.18, "advect_ms": 0.29}
    assert t["variant"] == "tail" and t["mismatched_elements_vs_default"] == 0 and t["tail_launches"] == 13
    assert b["variant"] == "broken" and "error" in b


def test_multi_gpu_experiments_compare_state_checksums(monkeypatch):
    calls = []

    def fake(cmd, env, timeout_s):
        calls.append((cmd, env))
        assert "--no-experiments" in cmd and "--checksum" in cmd and cmd[cmd.index("--nproc-per-node") + 1] == "4"
        if env.get("FXB_TAIL") == "1" and "FXB_P2P" not in env:
            return None, "hung"
        line = {"metric": "voxel_updates_per_s", "ms_per_step": 8.0 if "FXB_P2P" not in env else 5.0, "value": 1.0,
                "phase_ms": {"halo": 0.0, "jacobi": 3.0},
                "state_checksum": 77 if env.get("FXB_TAIL") != "1" else 78}
        return 0, "noise\n[rank0]: " + json.dumps(line) + "\n"

    monkeypatch.setattr(bench, "run_child", fake)
    monkeypatch.setenv("MASTER_PORT", "29500")

    class A:
        grid = None

    res = bench.experiments_multi_gpu(A(), 4, 120.0)["results"]
    assert [r["variant"] for r in res] == ["default", "p2p_halos", "tail_p2p", "tail"]
    assert res[0]["state_equals_default_variant"] is True and res[1]["state_equals_default_variant"] is True
    assert res[1]["ms_per_step"] == 5.0
    assert res[2]["state_equals_default_variant"] is False
    assert res[3]["error"] == "timeout"
    ports = [c[0][c[0].index("--master-port") + 1] for c in calls]
    assert len(set(ports)) == 4 and "29500" not in ports


def test_state_checksum_sees_a_moved_or_changed_word():
    class FX:
        FIELD_VELOCITY, FIELD_COLOR, FIELD_PRESSURE = 0, 1, 2

    class F:
        slab = (4, 3)

        def __init__(self, seed, tweak=None):
            r = np.random.default_rng(seed)
            self.a = {0: r.standard_normal((3, 5, 8, 4)).astype(np.float16),
                      1: r.standard_normal((3, 5, 8, 4)).astype(np.float16),
                      2: r.standard_normal((3, 5, 8)).astype(np.float32)}
            if tweak == "w":
                self.a[0][..., 3] = 9.0   # velocity .w is a don't-care
            if tweak == "bit":
                self.a[2].view(np.uint32)[1, 2, 3] ^= 1
            if tweak == "swap":
                self.a[1][[0, 1]] = self.a[1][[1, 0]]

        def get_field(self, fld):
            return self.a[fld]

    base = bench.state_checksum(None, None, F(1), FX, 1)
    assert base == bench.state_checksum(None, None, F(1, "w"), FX, 1)
    assert base != bench.state_checksum(None, None, F(1, "bit"), FX, 1)
    assert base != bench.state_checksum(None, None, F(1, "swap"), FX, 1)
    assert 0 <= base < 1 << 62


def test_multi_gpu_experiments_are_opt_in():
    src = open(os.path.join(ROOT, "bench.py")).read()
    assert "--experiments-multi" in src and "world in (1, 8) or args.experiments_multi" in src


def test_gpu_shot_bench_mode_runs_to_the_end_on_a_fake_device(monkeypatch, tmp_path):
    """tools/gpu_shot.py --bench (the child of bench.py's experiments stage) against a fake Fluid: every stage must
    emit a record that bench.experiments_single_gpu can summarise — no typo may cost the one free measurement."""
    import importlib
    import types

    import fluidx12_b200 as real

    class F:
        def Init(self, gridSize):
            self.n, self.k, self.env = tuple(gridSize), 0, {k: v for k, v in os.environ.items() if k.startswith("FXB_")}
            self.last_error = ""
            return True

        def step(self, dt):
            self.k += 1

        UpdateFrame = step

        def sync(self):
            pass

        def close(self):
            pass

        def get_field(self, fld):
            return np.full((2, 2, 2) + ((4,) if fld != real.FIELD_PRESSURE else ()), 0.25, np.float16 if fld != real.FIELD_PRESSURE else np.float32)

        def set_field(self, fld, a):
            pass

        def profile_step(self):
            return {"advect": 0.3, "divergence": 0.1, "jacobi": 1.0, "gradient": 0.1, "halo": 0.0, "step": 1.5}

        def tail_stats(self):
            return {"enabled": True, "tail_launches_last_step": 13, "tail_bricks": 1, "tail_subblocks_relaxed": 1,
                    "tail_subblocks_dense": 0}

        def stats(self):
            return types.SimpleNamespace(jacobi_passes=17, s_exec=64)

        def RayMarchL(self, p):
            pass

        def RayMarchV(self, v):
            self.s = int(v.cube_size)

        def get_light_map(self):
            return np.arange(8 ** 3, dtype=np.uint32).reshape(8, 8, 8)

        def get_cube_map(self):
            return np.ones((6, 4, 4, 4), np.uint8)

    fake = types.SimpleNamespace(Fluid=F, dt_for_grid=real.dt_for_grid, lib=real.lib, FxbLightParams=real.FxbLightParams,
                                 FxbViewParams=real.FxbViewParams, FIELD_VELOCITY=0, FIELD_COLOR=1, FIELD_PRESSURE=2)
    out = tmp_path / "shot.jsonl"
    monkeypatch.setenv("FXB_SHOT_OUT", str(out))
    monkeypatch.setattr(sys, "argv", ["gpu_shot.py", "--bench"])
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    shot = importlib.import_module("gpu_shot")
    shot = importlib.reload(shot)          # picks up FXB_SHOT_OUT
    shot.bench_mode(fake)
    rows = [json.loads(ln) for ln in open(out)]
    assert rows[-1]["stage"] == "done" and not any("error" in r for r in rows), [r for r in rows if "error" in r]
    stages = [r["stage"] for r in rows]
    assert stages.count("light_map") == 6 and stages.count("timing") == 2 + 10 + 5
    monkeypatch.setattr(bench, "run_child", lambda cmd, env, t: (open(env["FXB_SHOT_OUT"], "w").write(open(out).read()), (0, ""))[1])
    res = bench.experiments_single_gpu(5)
    assert len(res["results"]) == 6 + 17 and {r["variant"] for r in res["results"]} >= {"default", "tail", "advect2_only",
                                                                                        "tail_pass0", "light_map_pass"}
