"""Worker of tests/test_multi_gpu.py: one process per GPU (torchrun), z-slab run vs. a single-GPU run of the same grid.

Every rank steps its slab of the grid; rank 0 also steps the whole grid on its own GPU.  The slabs are gathered over
gloo and must be bit-identical to the single-GPU fields (halo exchange only moves data)."""
import ctypes as C
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fluidx12_b200 as fx  # noqa: E402
from tests.util import smooth_state  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    grid = tuple(int(v) for v in os.environ.get("FXB_TEST_GRID", "64,64,96").split(","))
    steps = int(os.environ.get("FXB_TEST_STEPS", "12"))
    fuse_t = int(os.environ.get("FXB_TEST_T", "2"))
    uid = torch.zeros(128, dtype=torch.uint8)
    if rank == 0:
        raw = (C.c_char * 128)()
        from fluidx12_b200 import binding as B
        B.check(fx.lib().fxb_nccl_unique_id(C.cast(raw, C.c_void_p)))
        uid = torch.frombuffer(bytearray(raw.raw), dtype=torch.uint8).clone()
    dist.broadcast(uid, 0)
    f = fx.Fluid()
    backend = {"fused": fx.HALO_FUSED, "peer": fx.HALO_PEER, "nccl": fx.HALO_NCCL}[os.environ.get("FXB_TEST_BACKEND", "fused")]
    assert f.Init(gridSize=grid, device=local, rank=rank, nranks=world, fuse_t=fuse_t, h_adv=int(os.environ.get("FXB_TEST_HADV", "8")),
                  halo_backend=backend, jacobi_group=int(os.environ.get("FXB_TEST_GROUP", "0")),
                  nccl_unique_id=uid.numpy().tobytes(), use_graph=bool(int(os.environ.get("FXB_TEST_GRAPH", "1")))), f.last_error
    z0, cnt = f.slab
    ref = None
    if rank == 0:
        ref = fx.Fluid()
        assert ref.Init(gridSize=grid, device=local, fuse_t=fuse_t), ref.last_error
    # identical smooth random start (exercises the advection halos immediately), then emitter-driven steps
    vel, col, p = smooth_state(*grid, seed=3, umax=1.5)
    for fld, a in ((fx.FIELD_VELOCITY, vel), (fx.FIELD_COLOR, col), (fx.FIELD_PRESSURE, p)):
        f.set_field(fld, a[z0:z0 + cnt])
        if ref:
            ref.set_field(fld, a)
    dt = fx.dt_for_grid(*grid)
    ok = True
    for step in range(steps):
        step_dt = 0.0 if step == 5 else dt  # one paused frame in the middle
        f.step(step_dt)
        if ref:
            ref.step(step_dt)
    f.sync()
    st = f.stats()
    for name, fld in (("velocity", fx.FIELD_VELOCITY), ("colour", fx.FIELD_COLOR), ("pressure", fx.FIELD_PRESSURE)):
        mine = torch.from_numpy(f.get_field(fld).view(np.uint8).reshape(-1).copy())
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([mine.numel()]))
        bufs = [torch.zeros(int(s.item()), dtype=torch.uint8) for s in sizes] if rank == 0 else None
        if rank == 0:
            bufs[0] = mine
            for r in range(1, world):
                dist.recv(bufs[r], src=r)
            whole = torch.cat(bufs).numpy()
            want = ref.get_field(fld)
            if name == "velocity":
                got = whole.view(np.float16).reshape(want.shape)[..., :3]
                want = want[..., :3]
            else:
                got = whole.view(want.dtype).reshape(want.shape)
            same = np.array_equal(got, want)
            print(f"{name}: {'identical' if same else 'MISMATCH %d' % int((got != want).sum())}")
            ok = ok and same
        else:
            dist.send(mine, dst=0)
    if os.environ.get("FXB_TEST_LIGHTMAP") == "1":
        # the light-map pass on z-slabs (density all-gathered over NCCL) against the single-GPU pass on rank 0
        lp = fx.FxbLightParams.reference_defaults()
        f.RayMarchL(lp)
        mine = torch.from_numpy(f.get_light_map().view(np.uint8).reshape(-1).copy())
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([mine.numel()]))
        if rank == 0:
            bufs = [mine] + [torch.zeros(int(s.item()), dtype=torch.uint8) for s in sizes[1:]]
            for r in range(1, world):
                dist.recv(bufs[r], src=r)
            ref.RayMarchL(lp)
            want = ref.get_light_map()
            got = torch.cat(bufs).numpy().view(np.uint32).reshape(want.shape)
            same = np.array_equal(got, want)
            print(f"light map: {'identical' if same else 'MISMATCH %d' % int((got != want).sum())}, "
                  f"{len(np.unique(want))} distinct words")
            ok = ok and same
        else:
            dist.send(mine, dst=0)
        # the cube-map marches on z-slabs (colour and light map of the whole grid gathered on every rank): every rank
        # ends up with the complete cube map, which must be the single-GPU one
        v = fx.FxbViewParams()
        v.eye_pt[:] = [4.0, 16.0, -40.0]
        v.world_i[:] = lp.world_i[:]
        v.num_samples, v.cube_size = 96, 32
        mask = C.c_uint32()
        fx.lib().fxb_cube_visibility_mask(v.world_i, v.eye_pt, C.byref(mask))
        v.visibility_mask = mask.value
        f.RayMarchV(v)
        cube_v = f.get_cube_map()
        v.num_samples, v.cube_size = 32, 16
        lp.num_samples = 16
        f.RayMarch(v, lp)
        cube_f = f.get_cube_map()
        if rank == 0:
            v.num_samples, v.cube_size = 96, 32
            ref.RayMarchV(v)
            same_v = np.array_equal(cube_v, ref.get_cube_map())
            v.num_samples, v.cube_size = 32, 16
            ref.RayMarch(v, lp)
            same_f = np.array_equal(cube_f, ref.get_cube_map())
            print(f"cube maps: {'identical' if same_v and same_f else 'MISMATCH'} "
                  f"({int((cube_v[..., 3] > 0).sum())} / {int((cube_f[..., 3] > 0).sum())} texels with smoke)")
            ok = ok and same_v and same_f
    if rank == 0:
        rs = ref.stats()
        print("s_exec", st.s_exec, rs.s_exec, "halo_overflow", st.halo_overflow)
        ok = ok and st.s_exec == rs.s_exec and st.halo_overflow == 0
    flag = torch.tensor([1 if ok else 0])
    dist.broadcast(flag, 0)
    dist.barrier()
    f.close()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_OK" if ok else "MGPU_FAIL")
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
