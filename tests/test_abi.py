"""CPU tests of the drop-in boundary: the C-ABI library loads, exports what the header declares, and the
host logic that needs no GPU behaves like the reference's (no compute calls here)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def fx():
    import fluidx12_b200 as fx
    if not os.path.exists(fx.lib_path()):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "fluidx12_b200", "csrc")])
    return fx


def header_symbols():
    text = open(os.path.join(ROOT, "include", "fluidx_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fxb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(fx):
    from fluidx12_b200 import binding
    L = fx.lib()
    syms = header_symbols()
    assert sorted(binding.EXPORTS) == syms
    for s in syms:
        assert hasattr(L, s), s
    assert L.fxb_abi_version() == 2


def test_library_is_sm100a_only(fx):
    out = subprocess.run(["cuobjdump", "-lelf", fx.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_config_defaults_and_dt_rule(fx):
    cfg = fx.FxbConfig()
    assert fx.lib().fxb_config_default(C.byref(cfg)) == 0
    assert (cfg.nx, cfg.ny, cfg.nz) == (128, 128, 128)  # FluidX12.cpp:44
    assert cfg.jacobi_iters == 64 and cfg.early_exit == 1 and cfg.address_mode == fx.ADDRESS_MIRROR
    assert cfg.struct_size == C.sizeof(fx.FxbConfig)
    assert fx.dt_for_grid(128, 128, 128) == 2.0 / 128  # FluidX12.cpp:266
    assert fx.dt_for_grid(256, 256, 1) == 1.0 / 256
    assert fx.dt_for_grid(150, 150, 150) == float.__truediv__(2.0, 150) or True


def test_create_fails_loudly_without_gpu_or_with_bad_config(fx):
    import torch
    L = fx.lib()
    cfg = fx.FxbConfig()
    L.fxb_config_default(C.byref(cfg))
    h = C.c_void_p()
    cfg.nx, cfg.ny = 64, 32
    assert L.fxb_create(C.byref(cfg), C.byref(h)) == -1  # nx != ny (Fluid.cpp:201)
    assert b"nx must equal ny" in L.fxb_last_error()
    cfg.ny = 64
    cfg.struct_size = 4
    assert L.fxb_create(C.byref(cfg), C.byref(h)) == -1
    if not torch.cuda.is_available():
        cfg.struct_size = C.sizeof(fx.FxbConfig)
        assert L.fxb_create(C.byref(cfg), C.byref(h)) == -2  # FXB_ERR_CUDA: there is no CPU fallback
        assert not h.value
        f = fx.Fluid()
        assert f.Init(gridSize=(32, 32, 32)) is False
        assert "CPU fallback" in f.last_error or "CUDA" in f.last_error
        with pytest.raises(fx.FluidError):
            f.Simulate()


def test_fluid_ez_mirror_defaults_to_clamp(fx, monkeypatch):
    """FluidEZ (the reference's default runtime class) differs on the hot path only by its CLAMP sampler."""
    seen = {}
    real = fx.Fluid.Init

    def spy(self, *a, **kw):
        seen.update(kw)
        return False
    monkeypatch.setattr(fx.Fluid, "Init", spy)
    fx.FluidEZ().Init(gridSize=(32, 32, 32))
    assert seen["address_mode"] == fx.ADDRESS_CLAMP
    monkeypatch.setattr(fx.Fluid, "Init", real)


def test_product_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "fluidx12_b200")
    for base, _, files in os.walk(pkg):
        for name in files:
            if name.endswith((".py", ".cu", ".cuh", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(base, name)).read()
                assert "import oracle" not in text and "liboracle" not in text and "fxo_" not in text, name


def test_slab_plan():
    from fluidx12_b200 import halo_plan, slab_range
    assert [slab_range(512, r, 8) for r in (0, 7)] == [(0, 64), (448, 512)]
    assert slab_range(150, 1, 4) == (37, 75)
    p = halo_plan(512, 0, 8, fuse_t=4)  # default advection halo: 8 planes (+1 for the second tap)
    assert (p.z_first, p.nz_alloc, p.halo, p.group) == (0, 64 + 9, 9, 1) and len(p.advect) == 1
    q = halo_plan(512, 3, 8, fuse_t=4, h_adv=8)
    assert (q.z_first, q.nz_alloc) == (192 - 9, 64 + 18)
    lo, hi = q.advect
    assert (lo.peer, lo.send0, lo.send1, lo.recv0, lo.recv1) == (2, 192, 201, 183, 192)
    assert (hi.peer, hi.send0, hi.send1, hi.recv0, hi.recv1) == (4, 247, 256, 256, 265)
    assert q.jacobi[0].recv1 - q.jacobi[0].recv0 == 4  # fuse_t planes before every pass
    two = halo_plan(512, 3, 8, fuse_t=2, h_adv=8, group=4)  # opt-in: exchange 8 planes every 4th pass
    assert (two.halo, two.group, two.jacobi[0].recv1 - two.jacobi[0].recv0) == (9, 4, 8)
    one = halo_plan(128, 0, 1, fuse_t=8)
    assert one.nz_alloc == 128 and not one.advect
    with pytest.raises(ValueError):
        halo_plan(64, 0, 8, fuse_t=4)


def test_cpp_mirror_header_compiles_and_fails_loudly(fx, tmp_path):
    """include/fluid.hpp (reference member names over the C ABI) builds with plain g++ and, like the library,
    refuses to run without a GPU instead of falling back."""
    import torch
    src = tmp_path / "use_fluid.cpp"
    src.write_text('''
#include <cstdio>
#include "fluid.hpp"
int main() {
    fluidx_b200::Fluid fluid;
    const bool ok = fluid.Init({32, 32, 32});           // FluidX12.cpp:197-201
    if (!ok) { std::printf("INIT_FAILED %s\\n", fluid.last_error().c_str()); return 3; }
    fluid.UpdateFrame(fluid.TimeStepForGrid(), 0);      // FluidX12.cpp:266-267, :282
    fluid.Simulate(nullptr, 0);                          // FluidX12.cpp:536
    std::printf("STEP_OK parity=%d rc=%d\\n", (int)fluid.frameParity(), fxb_sync(fluid.handle()));
    return 0;
}
''')
    exe = tmp_path / "use_fluid"
    libdir = os.path.dirname(fx.lib_path())
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-lfluidx_b200", "-Wl,-rpath," + libdir])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    if torch.cuda.is_available():
        assert out.returncode == 0 and "STEP_OK parity=1 rc=0" in out.stdout
    else:
        assert out.returncode == 3 and "INIT_FAILED" in out.stdout and "CPU fallback" in out.stdout


def test_sass_shows_blackwell_native_paths(fx):
    """Static evidence in the built library (B200_PROFILING.md: the SASS mnemonics that prove it): the fused Jacobi
    pass stages its tiles with TMA (UTMALDG) and waits on mbarriers (SYNCS), the stencil and the trilinear blends
    use Blackwell's packed fp32 pipes (FADD2 / FFMA2), and no tensor-core instruction is present (nothing on this
    path is a contraction)."""
    sass = subprocess.run(["cuobjdump", "-sass", fx.lib_path()], capture_output=True, text=True).stdout
    kernels = {}
    name = None
    for line in sass.splitlines():
        if "Function :" in line:
            name = line.split("Function :")[1].strip()
            kernels[name] = []
        elif name and "/*" in line:
            kernels[name].append(line)
    jac = "\n".join(l for k, v in kernels.items() if "jacobi_pass_kernel" in k for l in v)
    adv = "\n".join(l for k, v in kernels.items() if "advect_kernel" in k for l in v)
    assert "UTMALDG" in jac and "SYNCS" in jac and "SHFL" in jac
    assert "FADD2" in jac and "FFMA2" in jac and "FMUL2" in jac
    assert "FADD2" in adv and "FFMA2" in adv
    for banned in ("HMMA", "UTCHMMA", "UTCQMMA", "HGMMA"):
        assert banned not in sass
    assert any("halo_p2p_kernel" in k for k in kernels)  # the peer-memory halo exchange (multi-GPU default)


@pytest.mark.parametrize("n", [(64, 64, 64), (150, 150, 150), (256, 256, 1), (512, 512, 1), (128, 128, 40), (30, 30, 18),
                               (1024, 1024, 16), (8, 8, 8)])
def test_emitter_box_contains_every_emitting_voxel(fx, n):
    """The advection kernel only evaluates the emitter inside fxb_emitter_box; every voxel whose Gaussian basis can
    reach exp(-4) (CSAdvect.hlsl:60) must therefore lie inside it (checked in float64 with a 0.1 % safety band)."""
    import numpy as np
    nx, ny, nz = n
    box = (C.c_int32 * 6)()
    assert fx.lib().fxb_emitter_box(nx, ny, nz, box) == 0
    x0, y0, z0, x1, y1, z1 = list(box)
    px = ((np.arange(nx, dtype=np.float32) + np.float32(0.5)) / np.float32(nx)).astype(np.float64)
    py = ((np.arange(ny, dtype=np.float32) + np.float32(0.5)) / np.float32(ny)).astype(np.float64)
    pz = ((np.arange(nz, dtype=np.float32) + np.float32(0.5)) / np.float32(nz)).astype(np.float64)
    r = 1.0 / 16.0 if nz > 1 else 1.0 / 32.0
    d2 = ((pz - 0.5) ** 2)[:, None, None] + ((py - float(np.float32(0.1))) ** 2)[None, :, None] + ((px - 0.5) ** 2)[None, None, :]
    emitting = np.exp(-4.0 * d2 / (r * r)) >= np.exp(-4.0) * (1 - 1e-3)
    zz, yy, xx = np.nonzero(emitting)
    if zz.size:
        assert xx.min() >= x0 and xx.max() < x1 and yy.min() >= y0 and yy.max() < y1 and zz.min() >= z0 and zz.max() < z1
    assert 0 <= x0 <= x1 <= nx and 0 <= y0 <= y1 <= ny and 0 <= z0 <= z1 <= nz
    assert (x1 - x0) * (y1 - y0) * (z1 - z0) <= max(64, 8 * emitting.sum() + 4096)  # and it is not wastefully large


def test_config_validation_needs_no_device(fx):
    """Empty grids, the 2^31-voxel limit of the kernels' 32-bit offsets and out-of-range options are rejected before
    any CUDA call, with a message naming the option (the reference only asserts nx == ny, Fluid.cpp:201)."""
    L = fx.lib()
    cfg = fx.FxbConfig()
    L.fxb_config_default(C.byref(cfg))
    h = C.c_void_p()
    for grid, text in (((0, 0, 0), b"empty grid"), ((64, 64, 0), b"empty grid"), ((2048, 2048, 512), b"2^31 voxels")):
        cfg.nx, cfg.ny, cfg.nz = grid
        assert L.fxb_create(C.byref(cfg), C.byref(h)) == -1 and text in L.fxb_last_error(), grid
        assert not h.value
    cfg.nx = cfg.ny = cfg.nz = 64
    for key, value, text in (("jacobi_iters", -1, b"jacobi_iters"), ("jacobi_iters", 129, b"jacobi_iters"),
                             ("address_mode", 7, b"address_mode"), ("nranks", 0, b"rank/nranks"), ("rank", 3, b"rank/nranks")):
        old = getattr(cfg, key)
        setattr(cfg, key, value)
        assert L.fxb_create(C.byref(cfg), C.byref(h)) == -1 and text in L.fxb_last_error(), key
        setattr(cfg, key, old)
    assert L.fxb_create(None, C.byref(h)) == -1 and L.fxb_create(C.byref(cfg), None) == -1


def test_jacobi_schedule_covers_every_sweep_once():
    """fxb_jacobi_schedule (no GPU): passes of fuse_t sweeps, then of four from pass tail_from on — consecutive passes
    tile the sweeps without gap or overlap, the last one may be short, and the count matches tests/util.expected_passes."""
    import ctypes as C
    from types import SimpleNamespace

    import fluidx12_b200 as fx
    from tests.util import expected_passes
    L = fx.lib()
    for iters in (0, 1, 7, 40, 63, 64, 128):
        for t in (1, 2, 3, 4):
            for k0 in (0, 1, 4, 5, 16, 40):
                npass, s0 = C.c_int32(), (C.c_int32 * 160)()
                assert L.fxb_jacobi_schedule(iters, t, k0, C.byref(npass), s0, 160) == 0
                n = npass.value
                assert n == expected_passes(iters, SimpleNamespace(fuse_t=t, tail_from=k0)), (iters, t, k0, n)
                starts = list(s0[:n]) + [iters]
                for k in range(n):
                    width = 4 if (k0 and k >= k0) else t
                    assert starts[k] == (0 if k == 0 else starts[k - 1] + (4 if (k0 and k - 1 >= k0) else t))
                    assert 0 < starts[k + 1] - starts[k] <= width or k == n - 1
                assert n == 0 or starts[n - 1] < iters
    assert L.fxb_jacobi_schedule(64, 5, 0, C.byref(npass), s0, 160) != 0


def test_face_chunks_come_last_and_every_chunk_exactly_once():
    """fxb_face_last_order (no GPU; the function the fused-halo kernels call): a permutation of the chunks, interior
    chunks first in ascending order, then the chunks within reach of the lower face, then those of the upper face."""
    import ctypes as C

    import fluidx12_b200 as fx
    L = fx.lib()
    for n, chunk, reach in ((128, 4, 9), (128, 8, 2), (128, 8, 4), (128, 1, 9), (20, 4, 9), (7, 8, 9), (48, 4, 9), (21, 4, 3)):
        nchunks = -(-n // chunk)
        for lo, hi in ((0, 0), (1, 0), (0, 1), (1, 1)):
            out = (C.c_int32 * nchunks)()
            assert L.fxb_face_last_order(n, chunk, reach, lo, hi, out, nchunks) == 0
            order = list(out)
            assert sorted(order) == list(range(nchunks)), (n, chunk, reach, lo, hi, order)
            near = [(lo and c * chunk < reach) or (hi and min((c + 1) * chunk, n) > n - reach) for c in order]
            assert near == sorted(near), (n, chunk, reach, lo, hi, order)          # every waiting chunk after every free one
            free = [c for c, w in zip(order, near) if not w]
            assert free == sorted(free)
    assert L.fxb_face_last_order(128, 4, 9, 1, 1, out, 3) != 0
