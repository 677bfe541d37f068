"""Multi-GPU (z-slab + NCCL halo exchange) parity: N ranks must reproduce the single-GPU fields bit for bit."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(nproc, env_extra):
    env = dict(os.environ, **env_extra)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", "29611", os.path.join(ROOT, "tests", "mgpu_worker.py")]
    out = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=600)
    return out.returncode, out.stdout + out.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("grid,t", [("64,64,96", 2), ("72,72,64", 1), ("128,128,80", 4)])
def test_two_ranks_match_single_gpu(grid, t):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    rc, log = _run(2, {"FXB_TEST_GRID": grid, "FXB_TEST_T": str(t)})
    assert rc == 0 and "MGPU_OK" in log, log[-3000:]


@pytest.mark.gpu
def test_two_ranks_grouped_pressure_exchange():
    """Opt-in schedule: pressure halo exchanged every 4th pass, 4*T planes deep, halo planes relaxed redundantly."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    rc, log = _run(2, {"FXB_TEST_GRID": "64,64,96", "FXB_TEST_T": "2", "FXB_JACOBI_GROUP": "4"})
    assert rc == 0 and "MGPU_OK" in log, log[-3000:]


@pytest.mark.gpu
def test_four_ranks_and_eager_launch():
    import torch
    if torch.cuda.device_count() < 4:
        pytest.skip("needs 4 GPUs (gpurun --gpus 4)")
    rc, log = _run(4, {"FXB_TEST_GRID": "64,64,128", "FXB_TEST_T": "2", "FXB_TEST_GRAPH": "0"})
    assert rc == 0 and "MGPU_OK" in log, log[-3000:]


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("FXB_TEST_EXPERIMENTAL") != "1",
                    reason="the dynamic schedule on z-slabs has not run on GPUs yet: FXB_TEST_EXPERIMENTAL=1 enables it")
@pytest.mark.parametrize("nproc,grid", [(2, "64,64,96"), (4, "128,128,128")])
def test_ranks_with_tail_schedule(nproc, grid):
    """FXB_TAIL=1 on z-slabs: bulk pass 0, then 16 unconditional tail launches, one pressure/mask halo exchange of 4
    planes before each (17 exchanges per step instead of 33).  Emulated on the CPU by
    tests/test_tail_emu.py::test_tail_on_z_slabs_matches_single_domain."""
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (nproc, nproc))
    rc, log = _run(nproc, {"FXB_TEST_GRID": grid, "FXB_TEST_T": "2", "FXB_TEST_TAIL": "1"})
    assert rc == 0 and "MGPU_OK" in log, log[-3000:]


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("FXB_TEST_EXPERIMENTAL") != "1",
                    reason="the peer-memory halo exchange has not run on GPUs yet: FXB_TEST_EXPERIMENTAL=1 enables it")
@pytest.mark.parametrize("nproc,grid,tail", [(2, "64,64,96", "0"), (2, "72,72,64", "0"), (4, "128,128,128", "0"),
                                             (2, "64,64,96", "1")])
def test_ranks_with_peer_memory_halos(nproc, grid, tail):
    """FXB_P2P=1: face planes are stored straight into the neighbours' halo planes (CUDA IPC + NVLink) by one kernel
    per exchange that also publishes / awaits an epoch flag; results must stay bit-identical to a single GPU."""
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (nproc, nproc))
    rc, log = _run(nproc, {"FXB_TEST_GRID": grid, "FXB_TEST_T": "2", "FXB_P2P": "1", "FXB_TEST_TAIL": tail})
    assert rc == 0 and "MGPU_OK" in log, log[-3000:]


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("FXB_TEST_EXPERIMENTAL") != "1",
                    reason="the z-slab light-map pass and ray marches have not run on GPUs yet (FXB_TEST_EXPERIMENTAL=1 enables them)")
@pytest.mark.parametrize("nproc,grid", [(2, "64,64,96"), (4, "128,128,128")])
def test_ranks_light_map_and_ray_marches_match_single_gpu(nproc, grid):
    import torch
    if torch.cuda.device_count() < nproc:
        pytest.skip("needs %d GPUs" % nproc)
    rc, log = _run(nproc, {"FXB_TEST_GRID": grid, "FXB_TEST_T": "2", "FXB_TEST_LIGHTMAP": "1"})
    assert rc == 0 and "MGPU_OK" in log and "light map: identical" in log and "cube maps: identical" in log, log[-3000:]
