"""z-slab decomposition rules shared by the host code and the multi-rank tests.

The 3D grid is cut along z into ``nranks`` slabs (x-fastest storage makes a z-face contiguous, and
the plume rises along y so the back-trace is shortest along z — SURVEY.md §8e).  These functions are
the Python statement of the rules the C++ library applies in ``fxb_create`` (csrc/fxb_api.cu):
who owns which planes, how many halo planes are allocated, and which plane ranges are exchanged with
the z-1 / z+1 neighbours before each phase.  They contain no field arithmetic.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Tuple

DEFAULT_H_ADV = 8  # advection back-trace reach in planes (2*|u_z| voxels) before the +1 tap
DEFAULT_JACOBI_GROUP = 1  # fused passes per pressure-halo exchange; > 1 relaxes the halo planes in between redundantly (experimental)


def slab_range(nz: int, rank: int, nranks: int) -> Tuple[int, int]:
    """Global planes [z0, z1) owned by ``rank``."""
    if not (0 <= rank < nranks) or nranks > nz:
        raise ValueError("bad rank/nranks for this nz")
    return rank * nz // nranks, (rank + 1) * nz // nranks


@dataclass(frozen=True)
class Exchange:
    """One face exchange: send own planes [send0, send1) to ``peer``, receive [recv0, recv1) (global z)."""
    peer: int
    send0: int
    send1: int
    recv0: int
    recv1: int


@dataclass(frozen=True)
class HaloPlan:
    z0: int            # owned planes [z0, z1)
    z1: int
    z_first: int       # global z of local plane 0
    nz_alloc: int      # planes allocated (owned + halos, clipped at the global faces)
    halo: int          # halo depth allocated on interior faces
    advect: List[Exchange]   # velocity + colour before advect: h_adv + 1 planes
    stencil1: List[Exchange]  # 1 plane (advected velocity before divergence; pressure before gradient)
    jacobi: List[Exchange]   # group*fuse_t planes of rhs once, and of pressure (+ freeze mask) before every `group`-th pass
    group: int               # fused passes per pressure-halo exchange


def _faces(nz: int, rank: int, nranks: int, depth: int) -> List[Exchange]:
    z0, z1 = slab_range(nz, rank, nranks)
    out = []
    if rank > 0:  # lower neighbour owns [.., z0)
        lo = max(z0 - depth, slab_range(nz, rank - 1, nranks)[0])
        out.append(Exchange(rank - 1, z0, min(z0 + depth, z1), lo, z0))
    if rank < nranks - 1:
        hi = min(z1 + depth, slab_range(nz, rank + 1, nranks)[1])
        out.append(Exchange(rank + 1, max(z1 - depth, z0), z1, z1, hi))
    return out


def halo_plan(nz: int, rank: int, nranks: int, fuse_t: int, h_adv: int = 0, group: int = 0) -> HaloPlan:
    h_adv = h_adv or DEFAULT_H_ADV
    group = group or DEFAULT_JACOBI_GROUP
    z0, z1 = slab_range(nz, rank, nranks)
    if nranks == 1:
        return HaloPlan(z0, z1, 0, nz, 0, [], [], [], 1)
    halo = max(h_adv + 1, fuse_t)  # the group uses what the advection halo already provides
    group = max(1, min(group, halo // fuse_t))
    if halo > min(slab_range(nz, r, nranks)[1] - slab_range(nz, r, nranks)[0] for r in range(nranks)):
        raise ValueError("halo deeper than the thinnest slab: use fewer ranks or a smaller fuse_t/h_adv")
    z_first = max(z0 - halo, 0)
    z_last = min(z1 + halo, nz)
    return HaloPlan(z0, z1, z_first, z_last - z_first, halo,
                    _faces(nz, rank, nranks, h_adv + 1), _faces(nz, rank, nranks, 1),
                    _faces(nz, rank, nranks, group * fuse_t), group)


def gather_plan(nz: int, rank: int, nranks: int) -> List[Exchange]:
    """The light-map pass's exchange (csrc/halo.cu all_gather_slabs): a light ray crosses every slab, so every rank
    sends its owned planes of the density channel to every other rank and receives theirs — one send/recv pair per
    peer, all inside one group.  Global plane ranges."""
    z0, z1 = slab_range(nz, rank, nranks)
    return [Exchange(peer, z0, z1, *slab_range(nz, peer, nranks)) for peer in range(nranks) if peer != rank]
