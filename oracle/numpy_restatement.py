"""Second, independent restatement of the smoke-solver step in vectorised numpy.

TEST INFRASTRUCTURE ONLY (same rule as ``fluid_oracle.cpp``).  Written directly from the HLSL
(FluidX12/Content/Shaders/CSAdvect.hlsl:41-79, CSProject3D.hlsl:68-113, CSProject2D.hlsl:64-106,
CSPoisson.hlsli:8-26) with whole-array operations instead of per-voxel loops, so that a transcription
slip in the C++ oracle shows up as a mismatch between the two.  fp32 fused multiply-adds are emulated
as float32(float64(a)*float64(b) + float64(c)): the product of two fp32 numbers is exact in fp64, so the
only deviation from a true FMA is a double-rounding tie (probability ~2^-29 per operation).
The emitter's exp2 comes in as a precomputed ``basis`` array (numpy's exp2 is not libm's).
"""
from __future__ import annotations

import numpy as np

F = np.float32


def fma(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F)


def _tap(i, w, clamp):
    if clamp:
        return np.clip(i, 0, w - 1)
    m = np.mod(i, 2 * w)
    return np.where(m < w, m, 2 * w - 1 - m)


def _positions(shape):
    nz, ny, nx = shape
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    px = ((x.astype(F) + F(0.5)) / F(nx)).astype(F)
    py = ((y.astype(F) + F(0.5)) / F(ny)).astype(F)
    pz = ((z.astype(F) + F(0.5)) / F(nz)).astype(F)
    return px, py, pz


def emitter_basis_f64(shape):
    """Reference value of the Gaussian basis in float64 (for a tolerance check of the fp32 one)."""
    nz, ny, nx = shape
    px, py, pz = _positions(shape)
    d2 = (px.astype(np.float64) - 0.5) ** 2 + (py.astype(np.float64) - np.float64(F(0.1))) ** 2 + (pz.astype(np.float64) - 0.5) ** 2
    r = 1.0 / 16.0 if nz > 1 else 1.0 / 32.0
    return np.exp(-4.0 * d2 / (r * r))


def sample(field, ax, ay, az, clamp):
    """Trilinear fetch of an [nz,ny,nx,4] fp16 field at normalised coords; x then y then z lerps."""
    nz, ny, nx, _ = field.shape
    f32 = field.astype(F)
    tx, ty, tz = fma(ax, F(nx), F(-0.5)), fma(ay, F(ny), F(-0.5)), fma(az, F(nz), F(-0.5))
    ix, iy, iz = np.floor(tx), np.floor(ty), np.floor(tz)
    fx, fy, fz = (tx - ix).astype(F), (ty - iy).astype(F), (tz - iz).astype(F)
    ix, iy, iz = ix.astype(np.int64), iy.astype(np.int64), iz.astype(np.int64)
    x0, x1 = _tap(ix, nx, clamp), _tap(ix + 1, nx, clamp)
    y0, y1 = _tap(iy, ny, clamp), _tap(iy + 1, ny, clamp)
    z0, z1 = _tap(iz, nz, clamp), _tap(iz + 1, nz, clamp)

    def lerp(a, b, f):
        return fma(f[..., None], (b - a).astype(F), a)

    x00 = lerp(f32[z0, y0, x0], f32[z0, y0, x1], fx)
    x10 = lerp(f32[z0, y1, x0], f32[z0, y1, x1], fx)
    x01 = lerp(f32[z1, y0, x0], f32[z1, y0, x1], fx)
    x11 = lerp(f32[z1, y1, x0], f32[z1, y1, x1], fx)
    return lerp(lerp(x00, x10, fy), lerp(x01, x11, fy), fz)


def advect(vel, col, dt, basis, clamp=False):
    nz, ny, nx, _ = vel.shape
    dt = F(dt)
    px, py, pz = _positions((nz, ny, nx))
    u0 = vel.astype(F)
    ax, ay, az = fma(-u0[..., 0], dt, px), fma(-u0[..., 1], dt, py), fma(-u0[..., 2], dt, pz)
    u = sample(vel, ax, ay, az, clamp)
    c = sample(col, ax, ay, az, clamp)
    dx, dz = (px + F(-0.5)).astype(F), (pz + F(-0.5)).astype(F)
    if nz > 1:
        force = np.stack([(dz * F(-200.0)).astype(F), (basis * F(192.0)).astype(F), (dx * F(200.0)).astype(F)], -1)
    else:
        force = np.stack([np.zeros_like(basis), (basis * F(48.0)).astype(F), np.zeros_like(basis)], -1)
    hit = basis >= F(0.0183156393)
    u3 = np.where(hit[..., None], fma(force, dt, u[..., :3]), u[..., :3])
    bdt = (basis * dt).astype(F)
    imp = np.array([8.0, 16.0, 40.0, 40.0], F)
    c = np.where(hit[..., None], np.clip(fma(bdt[..., None], imp, c), F(0), F(1)), c)
    atten = np.maximum(fma(-dt, F(0.200000003), F(1.0)), F(0.0))
    vo = np.zeros_like(vel)
    vo[..., :3] = (u3 * atten).astype(F).astype(np.float16)
    return vo, (c * atten).astype(F).astype(np.float16)


def _shift(a, axis, d):
    """a[clamp(i + d)] along axis (the reference's cellMin/cellMax neighbour rule)."""
    n = a.shape[axis]
    idx = np.clip(np.arange(n) + d, 0, n - 1)
    return np.take(a, idx, axis=axis)


def project(vel, p, dt, iters=64, early_exit=True):
    """Returns (vel_out, p_out, s_exec)."""
    nz = vel.shape[0]
    u = vel.astype(F)[..., :3]
    if not dt > 0:
        vo = np.zeros_like(vel)
        vo[..., :3] = vel[..., :3]
        return vo, p.copy(), 0
    is3d = nz > 1
    a = (-_shift(u[..., 0], 2, -1) + _shift(u[..., 0], 2, 1)).astype(F)
    b = (-_shift(u[..., 1], 1, -1) + _shift(u[..., 1], 1, 1)).astype(F)
    if is3d:
        b = (b + a).astype(F)
        c = (-_shift(u[..., 2], 0, -1) + _shift(u[..., 2], 0, 1)).astype(F)
        s = (c + b).astype(F)
    else:
        s = (a + b).astype(F)
    inv = F(0.166666672) if is3d else F(0.25)
    p = p.astype(F).copy()
    active = np.ones(p.shape, bool)
    s_exec = 0
    for _ in range(iters):
        if not active.any():
            break
        s_exec += 1
        acc = fma(-s, F(0.5), _shift(p, 2, -1))
        acc = (_shift(p, 2, 1) + acc).astype(F)
        acc = (_shift(p, 1, -1) + acc).astype(F)
        acc = (_shift(p, 1, 1) + acc).astype(F)
        if is3d:
            acc = (_shift(p, 0, -1) + acc).astype(F)
            acc = (_shift(p, 0, 1) + acc).astype(F)
        new = (acc * inv).astype(F)
        done = np.abs(fma(acc, inv, -p)) < F(0.00100000005)
        p = np.where(active, new, p)
        if early_exit:
            active = active & ~done
    gx = (-_shift(p, 2, -1) + _shift(p, 2, 1)).astype(F)
    gy = (-_shift(p, 1, -1) + _shift(p, 1, 1)).astype(F)
    px, py, pz = _positions(p.shape)
    if is3d:
        gz = (-_shift(p, 0, -1) + _shift(p, 0, 1)).astype(F)
        k = F(1.04166675)
        un = np.stack([fma(-gx, k, u[..., 0]), fma(-gy, k, u[..., 1]), fma(-gz, k, u[..., 2])], -1)
        bp = np.stack([fma(px, F(2), F(-1)), fma(py, F(2), F(-1)), fma(pz, F(2), F(-1))], -1)
    else:
        un = np.stack([fma(-gx, F(0.5), u[..., 0]), fma(-gy, F(0.5), u[..., 1]), u[..., 2]], -1)
        bp = np.stack([fma(px, F(2), F(-1)), fma(py, F(2), F(-1)), pz], -1)
    m = ((-np.abs(bp) + F(0.970000029)).astype(F) * F(33.3333359)).astype(F)
    m = np.minimum(np.maximum(m, F(-1)), F(1))
    m = np.where((un * bp).astype(F) > 0, m, F(1))
    vo = np.zeros_like(vel)
    vo[..., :3] = (un * m).astype(F).astype(np.float16)
    return vo, p, s_exec


class NumpyFluid:
    """Init / UpdateFrame / Simulate on the numpy restatement (small grids only)."""

    def __init__(self, nx, ny, nz, basis, clamp=False, early_exit=True, iters=64):
        self.vel = [np.zeros((nz, ny, nx, 4), np.float16) for _ in range(2)]
        self.col = [np.zeros((nz, ny, nx, 4), np.float16) for _ in range(2)]
        self.p = np.zeros((nz, ny, nx), F)
        self.basis, self.clamp, self.early_exit, self.iters = basis.astype(F), clamp, early_exit, iters
        self.parity, self.dt, self.s_exec = 0, F(0), 0

    def step(self, dt):
        self.dt = F(dt)
        if dt > 0:
            self.parity ^= 1
        p = self.parity
        self.vel[1], self.col[p] = advect(self.vel[0], self.col[1 - p], self.dt, self.basis, self.clamp)
        self.vel[0], self.p, self.s_exec = project(self.vel[1], self.p, self.dt, self.iters, self.early_exit)
