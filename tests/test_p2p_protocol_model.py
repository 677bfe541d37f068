"""Model check of the peer-memory halo protocol (fluidx12_b200/csrc/halo.cu, FXB_P2P=1) — host logic only, no GPU.

halo_p2p_kernel stores a rank's face planes straight into its neighbours' halo planes, publishes an epoch and waits for
the neighbours' epochs; nothing acknowledges that a neighbour has finished READING a halo before it is overwritten.
The claim (header comment of halo_p2p_kernel) is that this is safe for the step's actual sequence of phases and
exchanges because a neighbour is never more than one exchange ahead and consecutive exchanges never touch the same
buffer.  This test replays that sequence (csrc/fxb_api.cu enqueue_phase: static schedule, and the dynamic schedule on
slabs) on R model ranks under thousands of random interleavings.  Every buffer carries a version; a phase checks, when
it ENDS, that each halo it read still holds exactly the version its neighbour published for it."""
import random

import pytest


def build_steps(nsteps, launches, p_cur=0, merged_rhs=False):
    """launches: relax kernels per step (32 fused passes, or 17 = bulk pass 0 + 16 tail launches).
    merged_rhs: the dynamic schedule on slabs sends the right-hand side together with launch 0's pressure halo."""
    prog = []
    parity = 0
    for _ in range(nsteps):
        parity ^= 1
        col_in, col_out = ("colA", "colB") if parity else ("colB", "colA")
        prog.append((("vel0", col_in), ("vel0", col_in), ("vel1", col_out)))          # advect
        prog.append((("vel1",), ("vel1",), ("rhs",)))                                 # divergence
        if not merged_rhs:
            prog.append((("rhs",), (), ()))                                           # rhs halo, no kernel of its own
        for k in range(launches):
            p_in, p_out = "p%d" % ((p_cur + k) & 1), "p%d" % ((p_cur + k + 1) & 1)
            m_in, m_out = "m%d" % (k & 1), "m%d" % ((k + 1) & 1)
            exch = ((p_in, "rhs") if merged_rhs else (p_in,)) if k == 0 else (p_in, m_in)
            reads = (p_in, "rhs") if k == 0 else (p_in, m_in, "rhs")
            prog.append((exch, reads, (p_out, m_out)))                                # relax kernel k
        p_cur = (p_cur + launches) & 1
        prog.append((("p%d" % p_cur,), ("p%d" % p_cur,), ("vel0",)))                  # final pressure halo, gradient
    return prog


class Rank:
    def __init__(self, r, nranks):
        self.r, self.n = r, nranks
        self.pc = 0               # index into the program
        self.stage = 0            # 0: exchange not issued, 1: waiting for the neighbours' epochs, 2: kernel running
        self.epoch = 0
        self.flag = {-1: 0, +1: 0}            # epochs published by the lower / upper neighbour
        self.version = {}                     # buffer -> version of the own planes
        self.halo = {-1: {}, +1: {}}          # side -> buffer -> version stored there by that neighbour
        self.expect = {}                      # buffer -> version the halos must hold while the current kernel runs

    def neighbours(self):
        return [d for d in (-1, +1) if 0 <= self.r + d < self.n]


def run(program, nranks, seed):
    rng = random.Random(seed)
    ranks = [Rank(r, nranks) for r in range(nranks)]
    while True:
        ready = []
        for k in ranks:
            if k.pc >= len(program):
                continue
            if k.stage == 1 and any(k.flag[d] < k.epoch for d in k.neighbours()):
                continue  # blocked in the wait of its exchange kernel
            ready.append(k)
        if not ready:
            assert all(k.pc >= len(program) for k in ranks), "deadlock"
            return
        k = rng.choice(ready)
        exch, reads, writes = program[k.pc]
        if k.stage == 0:      # exchange kernel: store the face planes into the neighbours' halos, publish the epoch
            k.epoch += 1
            for d in k.neighbours():
                o = ranks[k.r + d]
                for b in exch:
                    o.halo[-d][b] = k.version.get(b, 0)
                o.flag[-d] = k.epoch
            # SPMD: the neighbour's version of a buffer equals this rank's.  Expectations persist until the buffer is
            # exchanged again (the right-hand side's halo is sent once per step and read by every relax kernel).
            k.expect.update({b: k.version.get(b, 0) for b in exch})
            k.stage = 1
        elif k.stage == 1:    # the wait is over: the consumer kernel starts
            k.stage = 2
        else:                 # the consumer kernel ends: everything it read through a halo must still be what was sent
            for b in reads:
                if b in k.expect:
                    for d in k.neighbours():
                        assert k.halo[d].get(b) == k.expect[b], (
                            "rank %d: halo of %s from side %+d was overwritten while in use (phase %d)" % (k.r, b, d, k.pc))
            for b in writes:
                k.version[b] = k.version.get(b, 0) + 1
            k.pc += 1
            k.stage = 0


@pytest.mark.parametrize("nranks", [2, 3, 8])
@pytest.mark.parametrize("launches", [32, 17, 16, 1])
def test_unacknowledged_halo_stores_are_safe(nranks, launches):
    program = build_steps(3, launches, merged_rhs=(launches == 17))
    for seed in range(300 if nranks < 8 else 60):
        run(program, nranks, seed)


def test_the_model_detects_a_schedule_that_is_not_safe():
    """Control: exchanging the SAME buffer in two consecutive exchanges is exactly what the protocol cannot tolerate."""
    program = [(("p0",), ("p0",), ("p0",)), (("p0",), ("p0",), ("p0",))] * 20
    with pytest.raises(AssertionError):
        for seed in range(200):
            run(program, 3, seed)


@pytest.mark.parametrize("nz,nranks,halo", [(96, 2, 9), (128, 4, 13), (1024, 8, 13), (100, 3, 5), (64, 8, 4)])
def test_p2p_plane_arithmetic_matches_the_slab_plan(nz, nranks, halo):
    """The source / destination planes the peer-memory exchange computes (csrc/halo.cu p2p_planes, exported as
    fxb_p2p_plan) against the slab rules of fluidx12_b200/slab.py: what rank r stores into its neighbour's array must be
    exactly the planes that neighbour's exchange record says it receives, at the right local index."""
    import ctypes as C

    import fluidx12_b200 as fx
    from fluidx12_b200.slab import _faces, slab_range

    L = fx.lib()

    def z_first(r):
        return max(slab_range(nz, r, nranks)[0] - halo, 0)

    for depth in (1, 2, 4, min(halo, 9)):
        for r in range(nranks):
            out = (C.c_int64 * 4)()
            assert L.fxb_p2p_plan(nz, nranks, r, halo, depth, out) == 0
            send_lo, dst_lo, send_hi, dst_hi = out
            for e in _faces(nz, r, nranks, depth):
                # e: rank r sends own planes [send0, send1) to e.peer; the peer's record holds the matching receive
                back = [x for x in _faces(nz, e.peer, nranks, depth) if x.peer == r][0]
                assert (back.recv0, back.recv1) == (e.send0, e.send1)
                if e.peer == r - 1:
                    assert send_lo + z_first(r) == e.send0 and dst_lo + z_first(r - 1) == back.recv0
                else:
                    assert send_hi + z_first(r) == e.send0 and dst_hi + z_first(r + 1) == back.recv0
