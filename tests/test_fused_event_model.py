"""Model check of the fused-halo event protocol (fluidx12_b200/csrc/common.cuh PeerView, the default multi-GPU backend)
— host logic only, no GPU.

With fused halos there is no exchange: kernel m of a frame stores the planes next to an interior slab face into the
neighbour's halo while it runs, and publishes `epoch + m + 1` when it has completed; the part of kernel m that reads halo
planes or stores into a neighbour first waits until that neighbour has published `epoch + m` (it has finished kernel
m - 1).  Nothing acknowledges that a neighbour has finished READING a halo before it is overwritten.  The claim is that
this is safe for the step's actual kernel sequence (csrc/fxb_api.cu enqueue_phase) because a neighbour is never more than
one kernel ahead and no kernel stores into a buffer whose halo the kernel of the same index reads.

The model replays that sequence — advect, divergence, the fused passes (any count: 32 two-sweep passes, 24 or 18 with the
tail schedule), settle, gradient — on R ranks under random interleavings.  Every buffer carries a version; a kernel
checks when its face part STARTS that every halo it reads holds the version its neighbour produced for it (the data has
arrived) and when it ENDS that it still does (nobody overwrote it while in use)."""
import random

import pytest

EVENTS_PER_FRAME = 128  # common.cuh kEventsPerFrame


def frame_kernels(npass, p_cur, parity, unsafe=False):
    """The kernels of one frame as (halo buffers read, buffers pushed into the neighbours)."""
    col_in, col_out = ("colA", "colB") if parity else ("colB", "colA")
    ks = [(("vel0", col_in), ("vel1", col_out)),            # m = 0: advect
          (("vel1",), ("rhs",))]                            # m = 1: divergence
    for k in range(npass):
        p_in, p_out = "p%d" % ((p_cur + k) & 1), "p%d" % ((p_cur + k + 1) & 1)
        m_in, m_out = "m%d" % (k & 1), "m%d" % ((k + 1) & 1)
        reads = (p_in, "rhs") if k == 0 else (p_in, m_in, "rhs")
        ks.append((reads, (p_in, m_out) if unsafe else (p_out, m_out)))   # m = 2 + k: fused pass k
    y = "p%d" % ((p_cur + 1) & 1)                           # the first pass's output buffer holds the frame's result
    ks.append(((), (y,)))                                   # m = 2 + npass: settle (copies into Y, pushes them)
    ks.append(((y,), ("vel0",)))                            # m = 3 + npass: gradient
    return ks


def build_program(nframes, npass, unsafe=False):
    prog, p_cur, parity = [], 0, 0
    for f in range(nframes):
        parity ^= 1
        ks = frame_kernels(npass, p_cur, parity, unsafe)
        for m, (reads, pushes) in enumerate(ks):
            last = m == len(ks) - 1
            need = f * EVENTS_PER_FRAME + m
            publish = (f + 1) * EVENTS_PER_FRAME if last else need + 1  # the gradient publishes the next frame's base
            prog.append((reads, pushes, need, publish))
        p_cur ^= 1 if npass > 0 else 0
    return prog


class Rank:
    def __init__(self, r, nranks):
        self.r, self.n = r, nranks
        self.pc = 0
        self.stage = 0                        # 0 interior, 1 waiting, 2 face part running (pushes pending), 3 completing
        self.pending = []                     # (side, buffer) stores into the neighbours still to be issued
        self.events = {-1: 0, +1: 0}          # what the lower / upper neighbour has published
        self.version = {}                     # buffer -> times this rank has produced it
        self.halo = {-1: {}, +1: {}}          # side -> buffer -> version that neighbour stored there
        self.expect = {}

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
            if k.stage == 1 and any(k.events[d] < program[k.pc][2] for d in k.neighbours()):
                continue  # the face part waits for the neighbours' events
            ready.append(k)
        if not ready:
            assert all(k.pc >= len(program) for k in ranks), "deadlock"
            return
        k = rng.choice(ready)
        reads, pushes, need, publish = program[k.pc]
        if k.stage == 0:      # interior work: no halo, no neighbour (also the plain interior launch of the first pass)
            k.stage = 1
        elif k.stage == 1:    # the wait is over: the halos must hold what the neighbours produced for this kernel
            # SPMD: the neighbour's version of a buffer before its kernel m equals this rank's before its own kernel m
            k.expect = {b: k.version.get(b, 0) for b in reads}
            for b, v in k.expect.items():
                for d in k.neighbours():
                    assert k.halo[d].get(b, 0) == v, ("rank %d kernel %d: halo of %s from side %+d holds version %s, "
                                                      "expected %d" % (k.r, k.pc, b, d, k.halo[d].get(b), v))
            k.pending = [(d, b) for d in k.neighbours() for b in pushes]
            rng.shuffle(k.pending)
            k.stage = 2
        elif k.stage == 2:    # one store stream into one neighbour's halo (the kernel's own output, version + 1)
            if k.pending:
                d, b = k.pending.pop()
                ranks[k.r + d].halo[-d][b] = k.version.get(b, 0) + 1
            else:
                k.stage = 3
        else:                 # completion: nothing this kernel read through a halo was overwritten; publish the event
            for b, v in k.expect.items():
                for d in k.neighbours():
                    assert k.halo[d].get(b, 0) == v, ("rank %d kernel %d: halo of %s from side %+d was overwritten while "
                                                      "in use" % (k.r, k.pc, b, d))
            for b in pushes:
                k.version[b] = k.version.get(b, 0) + 1
            for d in k.neighbours():
                ranks[k.r + d].events[-d] = publish
            k.pc += 1
            k.stage = 0


@pytest.mark.parametrize("nranks", [2, 3, 8])
@pytest.mark.parametrize("npass", [32, 24, 18, 1])
def test_fused_halo_events_order_every_halo_access(nranks, npass):
    program = build_program(3, npass)
    for seed in range(200 if nranks < 8 else 40):
        run(program, nranks, seed)


def test_the_model_detects_a_kernel_that_stores_into_the_buffer_its_peer_reads():
    """Control: a pass that pushed its INPUT buffer would overwrite halos the neighbour's pass of the same index reads."""
    program = build_program(2, 8, unsafe=True)
    with pytest.raises(AssertionError):
        for seed in range(200):
            run(program, 3, seed)
