"""Multi-GPU (z-slab + halo exchange) parity: N ranks must reproduce the single-GPU fields bit for bit.

The default halo backend is fused into the step's kernels: every kernel stores its face planes straight into the
neighbours' memory (CUDA IPC over NVLink) and publishes an event counter (common.cuh PeerView).  The two exchange-kernel
backends (peer memory, NCCL send/recv) and their grouped pressure exchange are covered as well.  Needs `gpurun --gpus N`; the cases a
box cannot run are skipped."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


def _run(nproc, env_extra):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (nproc, nproc))
    env = dict(os.environ, **env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=420)
    log = out.stdout + out.stderr
    assert out.returncode == 0 and "MGPU_OK" in log, log[-3000:]
    return log


@pytest.mark.parametrize("grid,t", [("64,64,96", 2), ("72,72,64", 1), ("128,128,80", 4)])
def test_two_ranks_match_single_gpu(grid, t):
    _run(2, {"FXB_TEST_GRID": grid, "FXB_TEST_T": str(t)})


@pytest.mark.parametrize("backend", ["peer", "nccl"])
def test_two_ranks_exchange_backends(backend):
    _run(2, {"FXB_TEST_GRID": "64,64,96", "FXB_TEST_T": "2", "FXB_TEST_BACKEND": backend})


@pytest.mark.parametrize("backend", ["peer", "nccl"])
def test_two_ranks_grouped_pressure_exchange(backend):
    """Pressure halo exchanged every 4th pass, 4*T planes deep, halo planes relaxed redundantly in between."""
    _run(2, {"FXB_TEST_GRID": "64,64,96", "FXB_TEST_T": "2", "FXB_TEST_GROUP": "4", "FXB_TEST_BACKEND": backend})


def test_four_ranks_and_eager_launch():
    _run(4, {"FXB_TEST_GRID": "64,64,128", "FXB_TEST_T": "2", "FXB_TEST_GRAPH": "0"})


@pytest.mark.parametrize("backend,group", [("fused", 1), ("peer", 1), ("peer", 4)])
def test_four_ranks(backend, group):
    _run(4, {"FXB_TEST_GRID": "128,128,128", "FXB_TEST_T": "2", "FXB_TEST_GROUP": str(group), "FXB_TEST_BACKEND": backend})


@pytest.mark.parametrize("backend,group", [("fused", 1), ("peer", 4)])
def test_eight_ranks(backend, group):
    _run(8, {"FXB_TEST_GRID": "128,128,160", "FXB_TEST_T": "2", "FXB_TEST_GROUP": str(group), "FXB_TEST_BACKEND": backend})


@pytest.mark.parametrize("nproc,grid", [(2, "64,64,96"), (4, "128,128,128")])
def test_ranks_light_map_and_ray_marches_match_single_gpu(nproc, grid):
    log = _run(nproc, {"FXB_TEST_GRID": grid, "FXB_TEST_T": "2", "FXB_TEST_LIGHTMAP": "1"})
    assert "light map: identical" in log and "cube maps: identical" in log, log[-3000:]
