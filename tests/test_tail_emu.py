"""The tail kernel's body (fluidx12_b200/csrc/jacobi_tail_body.cuh), emulated on the CPU, against the oracle.

The emulation runs the same statements the CUDA kernel runs (tests/emu/tail_emu.cpp), so these tests pin the
indexing, the clamp-to-edge rule, the shrinking-validity argument, the freeze flags, the bit-packed masks and the
work-list hand-over bit for bit without a GPU.  The GPU tests then only have to show that the CUDA build of the same
body agrees (tests/test_gpu_tail.py)."""
import numpy as np
import pytest

from tests import tail_emu as E
from tests.util import smooth_state


def solve_with_tail(oracle_mod, s2, p_start, grid, iters=64, tt=4, early_exit=True, check_each=True, sparse_cap=-1,
                    cp_async=1, dense_mode=1, pass0_kernel=False):
    """Whole pressure solve by emulated tail launches only (first launch: every brick, every cell active).
    Returns (p, s_exec, launches).  After every launch the output buffer must equal the oracle's state everywhere."""
    nz, ny, nx = s2.shape
    g = E.BrickGrid(nx, ny, nz, *grid)
    rhs = (-0.5 * s2).astype(np.float32)
    assert np.array_equal(rhs.astype(np.float64), -0.5 * s2.astype(np.float64))  # exact
    p = [p_start.copy(), np.full_like(p_start, np.nan)]   # NaN = never written: must not leak into a result
    m = [np.zeros((nz, ny, nx // 8), np.uint8), np.zeros((nz, ny, nx // 8), np.uint8)]
    brick_state = np.zeros(g.n, np.int32)
    hist = np.zeros(iters + 8, np.uint64)
    p_ref, act_ref = p_start.copy(), np.ones(s2.shape, np.uint8)
    relax, copy = np.arange(g.n, dtype=np.int32), np.zeros(0, np.int32)
    done, seq = 0, 0
    while done < iters:
        # pass0_kernel: the first launch is the experimental pass-0 kernel (TailShape<2, 10, 12, 8>: two sweeps, every
        # cell active, second dense path)
        wide = pass0_kernel and seq == 0
        levels = min(2 if wide else tt, iters - done)
        src, dst = seq & 1, (seq + 1) & 1
        relax, copy_next = E.launch(g, p[src], p[dst], rhs, m[src], m[dst], relax, copy, brick_state, hist[done:],
                                    first=(seq == 0), early_exit=early_exit, levels=levels, tt=2 if wide else tt,
                                    sparse_cap=sparse_cap, cp_async=cp_async, dense_mode=2 if wide else dense_mode)
        p_ref, act_ref, counts = oracle_mod.jacobi_sweeps_slab(s2, p_ref, act_ref, levels, nz, 0, 0, nz, early_exit)
        done += levels
        seq += 1
        assert not brick_state.any()
        assert np.array_equal(hist[done - levels:done].astype(np.int64), counts), (seq, hist[done - levels:done], counts)
        if check_each:
            assert np.array_equal(p[dst], p_ref), (seq, int((p[dst] != p_ref).sum()))
            # flags of every brick that was relaxed or copied in this launch are current in the output mask
            got = E.unpack_mask(m[dst], nx)
            if seq >= 2 or True:
                touched = np.zeros(s2.shape, bool)
                for b in np.concatenate([relax, copy_next]):
                    touched[g.region(int(b))] = True
                assert np.array_equal(got[touched], act_ref[touched] if early_exit else np.ones_like(got[touched]))
        # the next lists partition the bricks that were relaxed: still active -> relax, frozen -> copy
        for b in relax:
            assert act_ref[g.region(int(b))].any() or not early_exit
        for b in copy_next:
            assert not act_ref[g.region(int(b))].any()
        copy = copy_next
        if counts[-1] == 0:
            break
    s_exec = 1 + int(np.count_nonzero(hist[:iters - 1])) if iters else 0
    return p[seq & 1], min(s_exec, iters), seq


def developed_state(oracle_mod, n, steps):
    """Divergence (x2) and warm-start pressure of a developed emitter-driven flow: realistic freeze behaviour."""
    f = oracle_mod.FluidOracle(*n)
    dt = oracle_mod.dt_for_grid(*n)
    for _ in range(steps):
        f.step(dt)
    vel, col, p = f.get_field(oracle_mod.FIELD_VEL), f.get_field(oracle_mod.FIELD_COLOR), f.get_field(oracle_mod.FIELD_PRESSURE)
    vo, _ = oracle_mod.advect(vel, col, dt)
    return oracle_mod.divergence2x(vo), p


@pytest.fixture(params=[False, True], ids=["ascending", "descending"])
def thread_order(request):
    E.thread_order(request.param)
    yield request.param
    E.thread_order(False)


@pytest.mark.parametrize("n,steps", [((64, 64, 64), 12), ((136, 136, 24), 6)])
@pytest.mark.parametrize("cp_async", [0, 1, 2])
@pytest.mark.parametrize("sparse_cap,dense_mode", [(-1, 1), (0, 1), (300, 1), (0, 2), (300, 2)])
def test_tail_only_solve_matches_oracle(oracle_mod, n, steps, sparse_cap, dense_mode, thread_order, cp_async):
    """sparse_cap -1: sparse path wherever the list fits; 0: dense path only; 300: both in one solve.
    dense_mode 1: register z-columns; 2: two-phase update of all quads."""
    s2, p0 = developed_state(oracle_mod, n, steps)
    p_want, s_want, hist_want, _ = oracle_mod.jacobi(s2, p0, 64, True)
    E.paths()
    p_got, s_got, launches = solve_with_tail(oracle_mod, s2, p0, (120, 12, 8), sparse_cap=sparse_cap,
                                             dense_mode=dense_mode, cp_async=cp_async)
    assert np.array_equal(p_got, p_want)
    assert s_got == s_want
    assert launches == -(-s_want // 4)
    n_copy, n_sparse, n_dense = E.paths()
    assert n_copy > 0
    assert (n_sparse > 0) == (sparse_cap != 0) and (n_dense > 0) == (sparse_cap != -1 or n_dense > 0)
    if sparse_cap == 300:
        assert n_sparse > 0 and n_dense > 0


@pytest.mark.parametrize("n,steps", [((64, 64, 64), 12), ((136, 136, 24), 6), ((248, 248, 16), 3)])
def test_pass0_kernel_then_tail_launches(oracle_mod, n, steps, thread_order):
    """The experimental pass-0 kernel (TailShape<2, 10, 12, 8>, every cell active, no flags yet, second dense path)
    followed by ordinary tail launches: the whole solve must equal the oracle."""
    s2, p0 = developed_state(oracle_mod, n, steps)
    p_want, s_want, _, _ = oracle_mod.jacobi(s2, p0, 64, True)
    p_got, s_got, launches = solve_with_tail(oracle_mod, s2, p0, (120, 12, 8), pass0_kernel=True)
    assert np.array_equal(p_got, p_want)
    assert s_got == s_want
    assert launches == 1 + -(-(s_want - 2) // 4)


def test_tail_random_field_no_early_exit(oracle_mod, thread_order):
    """Dense activity (nothing freezes), a grid whose last brick column is 16 cells wide, 10 sweeps = 4 + 4 + 2."""
    n = (136, 136, 20)
    _, _, p0 = smooth_state(*n, seed=7)
    rng = np.random.default_rng(5)
    s2 = (rng.integers(-2 ** 20, 2 ** 20, size=p0.shape) * 2.0 ** -24).astype(np.float32)
    p_want, _, _, _ = oracle_mod.jacobi(s2, p0, 10, False)
    p_got, _, launches = solve_with_tail(oracle_mod, s2, p0, (120, 12, 8), iters=10, early_exit=False, dense_mode=2)
    assert launches == 3
    assert np.array_equal(p_got, p_want)


def test_tail_two_sweep_shape_and_thin_bricks(oracle_mod, thread_order):
    """TT = 2 instantiation, bricks thinner than the compiled sub-block (by = 10, bz = 5), odd plane count."""
    n = (64, 64, 13)
    s2, p0 = developed_state(oracle_mod, n, 8)
    p_want, s_want, _, _ = oracle_mod.jacobi(s2, p0, 64, True)
    p_got, s_got, _ = solve_with_tail(oracle_mod, s2, p0, (120, 10, 5), tt=2, cp_async=0)
    assert np.array_equal(p_got, p_want)
    assert s_got == s_want


@pytest.mark.parametrize("nranks,n", [(2, (64, 64, 40)), (3, (136, 136, 36))])
def test_tail_on_z_slabs_matches_single_domain(oracle_mod, nranks, n):
    """The multi-GPU form of the tail launches, emulated: every rank holds its slab plus `halo` planes, the pressure
    (and, from the second launch on, the freeze mask) halo of TT planes is exchanged before each launch, the launch
    relaxes the bricks of the owned planes only.  Owned planes must equal the single-domain oracle bit for bit, and
    the summed histograms must match — this pins the slab-window geometry of jacobi_tail_body.cuh (z_face_lo / z_face_hi
    / nz_alloc / z_out0 / z_out1) before the schedule is run on more than one GPU."""
    from fluidx12_b200 import halo_plan
    nx, ny, nz = n
    tt, iters = 4, 64
    s2, p0 = developed_state(oracle_mod, n, 8)
    rhs_g = (-0.5 * s2).astype(np.float32)
    p_want, s_want, hist_want, _ = oracle_mod.jacobi(s2, p0, iters, True)

    class Rank:
        pass

    ranks = []
    for r in range(nranks):
        k = Rank()
        k.plan = halo_plan(nz, r, nranks, 2, h_adv=5)
        zf, nza = k.plan.z_first, k.plan.nz_alloc
        assert k.plan.halo >= tt
        k.win = slice(zf, zf + nza)
        k.p = [p0[k.win].copy(), np.full((nza, ny, nx), np.nan, np.float32)]
        k.rhs = rhs_g[k.win].copy()      # the right-hand side halo is exchanged once per step (constant over the sweeps)
        k.m = [np.zeros((nza, ny, nx // 8), np.uint8), np.zeros((nza, ny, nx // 8), np.uint8)]
        k.g = E.BrickGrid(nx, ny, k.plan.z1 - k.plan.z0, 120, 12, 8)
        k.slab = (nza, -zf, nz - zf, k.plan.z0 - zf, k.plan.z1 - zf)
        k.state = np.zeros(k.g.n, np.int32)
        k.hist = np.zeros(iters + 8, np.uint64)
        k.relax, k.copy = np.arange(k.g.n, dtype=np.int32), np.zeros(0, np.int32)
        ranks.append(k)

    def exchange(field_of, depth):
        """Every rank receives `depth` planes beyond each interior face from the owner of those planes."""
        for r, k in enumerate(ranks):
            zf = k.plan.z_first
            for peer, lo, hi in ((r - 1, k.plan.z0 - depth, k.plan.z0), (r + 1, k.plan.z1, k.plan.z1 + depth)):
                if 0 <= peer < nranks:
                    o = ranks[peer]
                    assert o.plan.z0 <= lo and hi <= o.plan.z1  # halo not deeper than the neighbour's slab
                    field_of(k)[lo - zf:hi - zf] = field_of(o)[lo - o.plan.z_first:hi - o.plan.z_first]

    done = seq = 0
    while done < iters:
        levels = min(tt, iters - done)
        src, dst = seq & 1, (seq + 1) & 1
        exchange(lambda k: k.p[src], tt)
        if seq:
            exchange(lambda k: k.m[src], tt)
        for k in ranks:
            k.relax, k.copy = E.launch(k.g, k.p[src], k.p[dst], k.rhs, k.m[src], k.m[dst], k.relax, k.copy, k.state,
                                       k.hist[done:], first=(seq == 0), levels=levels, tt=tt, slab=k.slab)
        done += levels
        seq += 1
        total = sum(k.hist[:done].astype(np.int64) for k in ranks)
        if total[done - 1] == 0:
            break
    got = np.concatenate([k.p[seq & 1][k.slab[3]:k.slab[4]] for k in ranks])
    assert np.array_equal(got, p_want)
    total = sum(k.hist[:iters].astype(np.int64) for k in ranks)
    assert np.array_equal(total[:s_want - 1], hist_want[1:s_want])


@pytest.mark.parametrize("n,steps", [((72, 72, 24), 5), ((8, 8, 8), 6), ((24, 24, 9), 7), ((40, 40, 17), 9)])
@pytest.mark.parametrize("cp_async,dense_mode,sparse_cap,pass0", [(1, 2, 200, False), (2, 1, 0, True)])
def test_tail_odd_grids_and_mode_combinations(oracle_mod, n, steps, cp_async, dense_mode, sparse_cap, pass0, thread_order):
    """Grids that are smaller than a window, not multiples of the brick, or a single brick thick, under mixed modes."""
    s2, p0 = developed_state(oracle_mod, n, steps)
    p_want, s_want, _, _ = oracle_mod.jacobi(s2, p0, 64, True)
    p_got, s_got, _ = solve_with_tail(oracle_mod, s2, p0, (120, 12, 8), cp_async=cp_async, dense_mode=dense_mode,
                                      sparse_cap=sparse_cap, pass0_kernel=pass0)
    assert np.array_equal(p_got, p_want) and s_got == s_want
