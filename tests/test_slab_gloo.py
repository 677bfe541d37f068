"""CPU test of the multi-rank host logic (world_size 2 and 3, gloo): the z-slab rules of fluidx12_b200.slab.

Each rank holds the window [z_first, z_first + nz_alloc) that ``halo_plan`` prescribes, exchanges exactly the face
ranges the plan lists (torch.distributed send/recv over gloo) and advances its window with the oracle's slab-window
stage functions in the same order as the CUDA multi-GPU step (csrc/fxb_api.cu: exchange velocity+colour -> advect ->
exchange advected velocity -> divergence -> exchange rhs -> every `group` fused passes: exchange pressure (+ freeze flags);
T sweeps per pass -> all-reduce the freeze counters -> exchange pressure -> gradient).  The owned planes must equal the
single-domain oracle bit for bit, which pins the halo depths, the exchange schedule and the face handling.
"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests.util import smooth_state  # noqa: E402


def _exchange(arr, faces, z_first):
    """Apply a list of slab.Exchange records to a window array (planes along axis 0)."""
    reqs, recvs = [], []
    for e in faces:
        send = torch.from_numpy(np.ascontiguousarray(arr[e.send0 - z_first:e.send1 - z_first]).view(np.uint8).reshape(-1).copy())
        recv = torch.empty(int(np.prod(arr[e.recv0 - z_first:e.recv1 - z_first].shape)) * arr.itemsize, dtype=torch.uint8)
        reqs.append(dist.isend(send, e.peer))
        reqs.append(dist.irecv(recv, e.peer))
        recvs.append((e, recv))
    for r in reqs:
        r.wait()
    for e, recv in recvs:
        dst = arr[e.recv0 - z_first:e.recv1 - z_first]
        dst[...] = recv.numpy().view(arr.dtype).reshape(dst.shape)


def _faces(nz, rank, world, depth):
    from fluidx12_b200.slab import _faces as faces
    return faces(nz, rank, world, depth)


def _worker(rank, world, port, grid, steps, fuse_t, h_adv, group, out, tail=False):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from fluidx12_b200 import halo_plan
    nx, ny, nz = grid
    plan = halo_plan(nz, rank, world, fuse_t, h_adv, group)
    zf, nza = plan.z_first, plan.nz_alloc
    own = slice(plan.z0 - zf, plan.z1 - zf)
    vel_g, col_g, p_g = smooth_state(nx, ny, nz, seed=9, umax=1.0)
    vel0 = vel_g[zf:zf + nza].copy()
    col = [np.zeros_like(vel0), col_g[zf:zf + nza].copy()]  # col[parity]; parity starts at 0, flips before step 1
    col = [col[1], np.zeros_like(vel0)]
    p = p_g[zf:zf + nza].copy()
    parity, dt, iters = 0, O.dt_for_grid(nx, ny, nz), 64
    npass = -(-iters // fuse_t)
    s_exec = 0
    for _ in range(steps):
        parity ^= 1
        _exchange(vel0, plan.advect, zf)
        _exchange(col[1 - parity], plan.advect, zf)
        vel1, col[parity] = O.advect_slab(vel0, col[1 - parity], dt, nz, zf)
        _exchange(vel1, plan.stencil1, zf)
        s = O.divergence2x_slab(vel1, nz, zf)
        _exchange(s, plan.jacobi, zf)
        active = np.ones(s.shape, np.uint8)
        counts = np.zeros(npass * fuse_t, np.int64)
        if tail:
            # dynamic schedule on slabs (csrc/fxb_api.cu): bulk pass 0 (T sweeps), then tail launches of 4 sweeps; one
            # exchange before every launch, as deep as the launch's sweeps; the right-hand side once, 4 planes deep
            _exchange(s, _faces(nz, rank, world, 4), zf)
            done, launch = 0, 0
            while done < iters:
                n = min(fuse_t if launch == 0 else 4, iters - done)
                depth = fuse_t if launch == 0 else 4
                _exchange(p, _faces(nz, rank, world, depth), zf)
                if launch:
                    _exchange(active, _faces(nz, rank, world, depth), zf)
                p, active, c = O.jacobi_sweeps_slab(s, p, active, n, nz, zf, own.start, own.stop)
                counts[done:done + n] = c
                done += n
                launch += 1
        for k in range(0 if tail else npass):
            if k % plan.group == 0:  # group*T planes every `group` passes; the window is relaxed whole in between
                _exchange(p, plan.jacobi, zf)
                if k:
                    _exchange(active, plan.jacobi, zf)
            p, active, c = O.jacobi_sweeps_slab(s, p, active, fuse_t, nz, zf, own.start, own.stop)
            counts[k * fuse_t:(k + 1) * fuse_t] = c
        t = torch.from_numpy(counts)
        dist.all_reduce(t)
        counts = t.numpy()
        s_exec = 1 + int(np.argmax(np.append(counts[:iters - 1] == 0, True))) if iters else 0
        _exchange(p, plan.stencil1, zf)
        vel0 = O.gradient_slab(vel1, p, nz, zf)
    res = {"z0": plan.z0, "vel": vel0[own][..., :3].copy(), "col": col[parity][own].copy(), "p": p[own].copy(),
           "s_exec": s_exec}
    gathered = [None] * world
    dist.all_gather_object(gathered, res)
    if rank == 0:
        o = O.FluidOracle(nx, ny, nz)
        o.set_field(O.FIELD_VEL, vel_g); o.set_field(O.FIELD_COLOR, col_g); o.set_field(O.FIELD_PRESSURE, p_g)
        for _ in range(steps):
            o.step(dt)
        gathered.sort(key=lambda r: r["z0"])
        ok = (np.array_equal(np.concatenate([r["vel"] for r in gathered]), o.get_field(O.FIELD_VEL)[..., :3])
              and np.array_equal(np.concatenate([r["col"] for r in gathered]), o.get_field(O.FIELD_COLOR))
              and np.array_equal(np.concatenate([r["p"] for r in gathered]), o.get_field(O.FIELD_PRESSURE))
              and all(r["s_exec"] == o.s_exec for r in gathered))
        out.put(bool(ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,grid,fuse_t,h_adv,group", [(2, (16, 16, 24), 2, 3, 1), (3, (16, 16, 30), 4, 5, 1),
                                                           (2, (24, 24, 20), 1, 5, 1), (2, (16, 16, 24), 2, 7, 4),
                                                           (3, (16, 16, 36), 2, 7, 4)])
def test_slab_decomposition_matches_single_domain(world, grid, fuse_t, h_adv, group):
    import oracle
    oracle.build()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = 29700 + world * 10 + fuse_t + 3 * group
    procs = [ctx.Process(target=_worker, args=(r, world, port, grid, 3, fuse_t, h_adv, group, out)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=300)
        assert pr.exitcode == 0
    assert out.get(timeout=10) is True


@pytest.mark.parametrize("world,grid", [(2, (16, 16, 24)), (3, (16, 16, 36))])
def test_slab_decomposition_with_the_dynamic_schedule(world, grid):
    """The exchange plan of the dynamic pressure-solve schedule on slabs (bulk pass 0, then launches of 4 sweeps with
    4-plane halos: 17 exchanges per step instead of 33) at oracle level over gloo."""
    import oracle
    oracle.build()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, 29790 + world, grid, 3, 2, 5, 1, out, True)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=300)
        assert pr.exitcode == 0
    assert out.get(timeout=10) is True


def _lightmap_worker(rank, world, port, grid, out):
    """Light-map pass on z-slabs (csrc/lightmap.cu with nranks > 1): every rank extracts the density channel of its own
    planes, the ranks exchange them per slab.gather_plan, and each runs the kernel's per-voxel body (CPU emulation of
    lightmap_body.cuh) on its planes over the gathered array."""
    import ctypes as C
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle as O
    from fluidx12_b200 import gather_plan, slab_range
    from tests.test_lightmap import colour_field, light_constants, oracle_params
    nx, ny, nz = grid
    col = colour_field(grid, 21)
    _, plain = light_constants(24, 1, (75.0, 75.0, -75.0), 6)
    p = oracle_params(plain)
    z0, z1 = slab_range(nz, rank, world)
    dens = np.zeros((nz, ny, nx), np.uint16)              # the whole grid's density; only the own planes are known
    dens[z0:z1] = col[z0:z1, ..., 3].view(np.uint16)
    _exchange(dens, gather_plan(nz, rank, world), 0)
    emu = C.CDLL(os.path.join(ROOT, "tests", "emu", "liblightmap_emu.so"))
    emu.lightmap_emu_run_slab.restype = None
    mine = np.empty((z1 - z0, ny, nx), np.uint32)
    emu.lightmap_emu_run_slab(nx, ny, nz, dens.ctypes.data_as(C.c_void_p), C.byref(p), z0, z1, mine.ctypes.data_as(C.c_void_p))
    want = O.light_map(col, p)
    ok = bool(np.array_equal(dens.view(np.float16), col[..., 3])) and bool(np.array_equal(mine, want[z0:z1]))
    flag = torch.tensor([1 if ok else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        out.put(bool(flag.item()))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,grid", [(2, (16, 16, 12)), (3, (12, 12, 16))])
def test_light_map_on_z_slabs_matches_single_domain(world, grid):
    import subprocess

    import oracle
    oracle.build()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "tests", "emu"), "liblightmap_emu.so"])
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_lightmap_worker, args=(r, world, 29840 + world, grid, out)) for r in range(world)]
    for pr in procs:
        pr.start()
    for pr in procs:
        pr.join(timeout=300)
        assert pr.exitcode == 0
    assert out.get(timeout=10) is True
