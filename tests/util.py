"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np


def smooth_state(nx, ny, nz, seed=1234, umax=2.0):
    """Band-limited random velocity (|u| <= umax), colour in [0,1] and pressure."""
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.arange(nz) / max(nz, 1), np.arange(ny) / ny, np.arange(nx) / nx, indexing="ij")

    def field():
        f = np.zeros((nz, ny, nx))
        for _ in range(6):
            kx, ky, kz = rng.integers(1, 4, 3)
            ph = rng.uniform(0, 2 * np.pi, 3)
            f += rng.uniform(-1, 1) * np.sin(2 * np.pi * kx * x + ph[0]) * np.sin(2 * np.pi * ky * y + ph[1]) * \
                np.cos(2 * np.pi * kz * z + ph[2])
        return f / max(np.abs(f).max(), 1e-9)

    vel = np.zeros((nz, ny, nx, 4), np.float16)
    for c in range(3 if nz > 1 else 2):
        vel[..., c] = (umax * field()).astype(np.float16)
    col = np.zeros((nz, ny, nx, 4), np.float16)
    for c in range(4):
        col[..., c] = (0.5 + 0.5 * field()).astype(np.float16)
    p = (0.5 * field()).astype(np.float32)
    return vel, col, p


def max_abs_rel(a, b):
    """max|a-b| / max|b| per SURVEY.md §8: the 1-step gate is <= 1e-5."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def rel_l2(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.sqrt(((a - b) ** 2).sum()) / max(np.sqrt((b ** 2).sum()), 1e-30))


def expected_passes(s_exec, st):
    """Fused passes a frame of `s_exec` sweeps takes under the schedule the stats record describes: `fuse_t` sweeps per
    pass, four per pass from pass `tail_from` on (0: no tail schedule)."""
    t, k0 = st.fuse_t, st.tail_from
    n = -(-s_exec // t)
    if k0 and n > k0:
        n = k0 + -(-(s_exec - k0 * t) // 4)
    return n
