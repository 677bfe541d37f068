// advect.cu — CSAdvect as a software fp32 trilinear back-trace kernel (sm_100a).
//
// Replaces FluidX12/Content/Shaders/CSAdvect.hlsl:41-79 (dispatch Fluid.cpp:374).  Operation order
// follows the shipped DXBC; the two hardware SampleLevel fetches become 16
// gathered 8-byte texel loads blended with fp32 weights (no texture unit, no 8-bit weights).
//
// Mapping: a thread marches a short z column (kZ voxels); a CTA of 32 x 8 threads covers 32 x 8 x kZ voxels, so the
// tap footprints of neighbouring threads and of consecutive planes are shared through L1 (the back-trace is spatially
// coherent).  The kernel is bounded by instruction issue and by the latency of its two dependent memory round trips
// (the voxel's own velocity, then the taps) before HBM, so:
//   * the voxel's own velocity AND colour texel of plane z+1 are fetched while plane z is processed (software
//     pipelining along z): the first round trip is hidden, and a back-trace that lands exactly on the voxel's own
//     texel centre (the still-quiescent far field, most of a large grid) needs no further load at all — that path is a
//     pure stream, 16 B in + 16 B out per voxel;
//   * (i + 0.5) / N comes from per-axis tables (three IEEE divisions per voxel otherwise);
//   * taps that all fall inside the grid (the common case) take a fast path with 32-bit offsets and no
//     sampler-addressing arithmetic; MIRROR/CLAMP wrapping lives in a cold out-of-line path;
//   * the 14 lerps of the 7 used channels run as packed FADD2/FFMA2 on (x,y) and (z,w) pairs.
// Algorithmic traffic: 32 B/voxel (velocity in 8 + colour in 8 + velocity out 8 + colour out 8).
// Contract: fields are finite (an Inf/NaN texel next to an exactly-hit tap would make the general blend NaN where the
// exact-hit shortcut returns the tap; the reference's own fields never leave [-65504, 65504]).
#include "common.cuh"
#include "kernels.h"

namespace fxb {

namespace {

constexpr int kZ = 4;  // voxels a thread marches along z (measured on B200: 2 is 35 % slower at 512^3, 8 is 2 % faster there
                       // and 14 % slower at 256^3)

struct Pair4 {  // one RGBA16F texel widened to fp32 as two packed pairs
    float2 lo, hi;
};

__device__ __forceinline__ Pair4 widen(const uint2 r) {
    Pair4 t;
    t.lo = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    t.hi = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return t;
}

__device__ __forceinline__ float2 lerp2(float2 f, float2 a, float2 b) { return fma2(f, sub2(b, a), a); }

// x, then y, then z; each lerp is fma(f, b - a, a)
__device__ __forceinline__ Pair4 blend(const uint2 (&raw)[8], float fx, float fy, float fz) {
    Pair4 t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = widen(raw[k]);
    const float2 wx = make_float2(fx, fx), wy = make_float2(fy, fy), wz = make_float2(fz, fz);
    Pair4 r;
    {
        const float2 x00 = lerp2(wx, t[0].lo, t[1].lo), x10 = lerp2(wx, t[2].lo, t[3].lo);
        const float2 x01 = lerp2(wx, t[4].lo, t[5].lo), x11 = lerp2(wx, t[6].lo, t[7].lo);
        r.lo = lerp2(wz, lerp2(wy, x00, x10), lerp2(wy, x01, x11));
    }
    {
        const float2 x00 = lerp2(wx, t[0].hi, t[1].hi), x10 = lerp2(wx, t[2].hi, t[3].hi);
        const float2 x01 = lerp2(wx, t[4].hi, t[5].hi), x11 = lerp2(wx, t[6].hi, t[7].hi);
        r.hi = lerp2(wz, lerp2(wy, x00, x10), lerp2(wy, x01, x11));
    }
    return r;
}

// Cold path: sampler addressing of one axis when a tap may be outside the grid (or the coordinate is not
// finite).  Out of line and returned in registers so that the hot path carries none of its cost.
__device__ __noinline__ int2 wrapped_pair(float t, int w, int clamp_mode) {
    const int i = floor_to_tap(t);
    return make_int2(address_tap(i, w, clamp_mode), address_tap(i + 1, w, clamp_mode));
}

struct AdvectArgs {
    Domain d;
    AxisTables tab;
    Emitter em;
    int clamp_mode;
    int zv0, zv1;  // global planes [zv0, zv1) hold valid input: the owned planes plus the exchanged halo (or the grid's faces)
    // fused halos (common.cuh PeerView): planes within `reach` of an interior face read halo planes, and the same
    // planes of the colour output (the first / last plane of the velocity output) are also stored into the neighbour
    int reach;
    uint2 *vel_out_lo, *vel_out_hi;  // the neighbours' m_velocities[1]
    uint2 *col_lo[2], *col_hi[2];    // the neighbours' m_colors[0], m_colors[1]
};

struct Texels {
    uint2 vel, col;
};

// One voxel, any case (taps inside or outside the grid, either sampler addressing mode).  `sv` / `sc`: the voxel's
// own velocity / colour texel (already loaded).
__device__ __forceinline__ Texels advect_voxel(const AdvectArgs& A, const float dt, const float atten,
                                               const uint2* __restrict__ vel_in, const uint2* __restrict__ col_in,
                                               StepState* __restrict__ state, const int x, const int y, const int z,
                                               const float px, const float py, const unsigned self, const uint2 sv,
                                               const uint2 sc) {
    const Domain& d = A.d;
    const float pz = __ldg(A.tab.pos[2] + z);
    const float fnx = (float)d.nx, fny = (float)d.ny, fnz = (float)d.nz;

    const Pair4 u0 = widen(sv);
    const float tx = __fmaf_rn(__fmaf_rn(-u0.lo.x, dt, px), fnx, -0.5f);
    const float ty = __fmaf_rn(__fmaf_rn(-u0.lo.y, dt, py), fny, -0.5f);
    const float tz = __fmaf_rn(__fmaf_rn(-u0.hi.x, dt, pz), fnz, -0.5f);
    const float flx = floorf(tx), fly = floorf(ty), flz = floorf(tz);
    const float fx = tx - flx, fy = ty - fly, fz = tz - flz;

    unsigned o[8];
    const float zlo = (float)A.zv0, zhi = (float)(A.zv1 - 1);
    const bool inside = tx >= 0.0f && tx < fnx - 1.0f && ty >= 0.0f && ty < fny - 1.0f && tz >= 0.0f &&
                        tz < fnz - 1.0f && tz >= zlo && tz < zhi;
    if (inside) {  // both taps of every axis are inside the grid (and inside the valid planes of the local slab)
        const unsigned plane = (unsigned)d.nx * d.ny;
        const unsigned base = ((unsigned)((int)flz - d.z_first) * d.ny + (unsigned)(int)fly) * d.nx + (unsigned)(int)flx;
        o[0] = base; o[1] = base + 1; o[2] = base + d.nx; o[3] = base + d.nx + 1;
        o[4] = base + plane; o[5] = o[4] + 1; o[6] = o[4] + d.nx; o[7] = o[6] + 1;
    } else {
        const int2 xs = wrapped_pair(tx, d.nx, A.clamp_mode), ys = wrapped_pair(ty, d.ny, A.clamp_mode);
        int2 zs = wrapped_pair(tz, d.nz, A.clamp_mode);
        if (zs.x < A.zv0 || zs.x >= A.zv1 || zs.y < A.zv0 || zs.y >= A.zv1) {
            state->halo_overflow = 1;  // the back-trace left the exchanged z-halo (multi-GPU only)
            zs.x = min(max(zs.x, A.zv0), A.zv1 - 1);
            zs.y = min(max(zs.y, A.zv0), A.zv1 - 1);
        }
        zs.x -= d.z_first;
        zs.y -= d.z_first;
        const unsigned r00 = ((unsigned)zs.x * d.ny + ys.x) * d.nx, r10 = ((unsigned)zs.x * d.ny + ys.y) * d.nx;
        const unsigned r01 = ((unsigned)zs.y * d.ny + ys.x) * d.nx, r11 = ((unsigned)zs.y * d.ny + ys.y) * d.nx;
        o[0] = r00 + xs.x; o[1] = r00 + xs.y; o[2] = r10 + xs.x; o[3] = r10 + xs.y;
        o[4] = r01 + xs.x; o[5] = r01 + xs.y; o[6] = r11 + xs.x; o[7] = r11 + xs.y;
    }
    // Exact-texel fast path.  When the back-trace lands exactly on a texel centre (all three weights are 0: in
    // practice the voxels the flow has not reached, u = 0) every lerp is fma(0, b - a, a) = a, i.e. the fetch
    // returns the first tap unchanged.  That tap is normally the voxel's own texel, which is already here.  The one
    // case where fma(0, b - a, a) != a bitwise is a = -0 (the sum takes the sign of 0 * (b - a)); such texels take
    // the general path.
    Pair4 u, c;
    bool exact = inside && fx == 0.0f && fy == 0.0f && fz == 0.0f;
    if (exact) {
        uint2 rv = sv, rc = sc;
        if (o[0] != self) {
            rv = __ldg(vel_in + o[0]);
            rc = __ldg(col_in + o[0]);
        }
        auto neg_zero = [](unsigned w) { return (w & 0xffffu) == 0x8000u || (w >> 16) == 0x8000u; };
        if (neg_zero(rv.x) || neg_zero(rv.y) || neg_zero(rc.x) || neg_zero(rc.y)) {
            exact = false;
        } else {
            u = widen(rv);
            c = widen(rc);
        }
    }
    if (!exact) {
        // all 16 taps in flight together
        uint2 rv[8], rc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) rv[k] = __ldg(vel_in + o[k]);
#pragma unroll
        for (int k = 0; k < 8; ++k) rc[k] = __ldg(col_in + o[k]);
        u = blend(rv, fx, fy, fz);
        // Most moving voxels are outside the smoke: when all eight colour taps are +0 every lerp is fma(f, +0, +0) = +0
        // (the weights of an inside fetch are finite and in [0, 1)), so the blend is skipped.
        unsigned any = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) any |= rc[k].x | rc[k].y;
        if (inside && any == 0u) c.lo = c.hi = make_float2(0.0f, 0.0f);
        else c = blend(rc, fx, fy, fz);
    }

    // Emitter (CSAdvect.hlsl:57-68).  Outside the table's box the basis is below exp(-4) by construction.
    const Emitter& em = A.em;
    if (x >= em.x0 && x < em.x1 && y >= em.y0 && y < em.y1 && z >= em.z0 && z < em.z1) {
        const float basis =
            __ldg(em.basis + ((size_t)(z - em.z0) * (em.y1 - em.y0) + (y - em.y0)) * (em.x1 - em.x0) + (x - em.x0));
        if (basis >= 0.0183156393f) {
            float fx_, fy_, fz_;
            if (1.0f < fnz) {
                const float dx = px + -0.5f, dz = pz + -0.5f;
                fx_ = __fmaf_rn(basis, 0.0f, dz * -200.0f);
                fy_ = __fmaf_rn(basis, 192.0f, 0.0f);
                fz_ = __fmaf_rn(basis, 0.0f, dx * 200.0f);
            } else {
                fx_ = 0.0f; fy_ = basis * 48.0f; fz_ = 0.0f;
            }
            u.lo.x = __fmaf_rn(fx_, dt, u.lo.x);
            u.lo.y = __fmaf_rn(fy_, dt, u.lo.y);
            u.hi.x = __fmaf_rn(fz_, dt, u.hi.x);
            const float bdt = basis * dt;
            c.lo.x = __saturatef(__fmaf_rn(bdt, 8.0f, c.lo.x));
            c.lo.y = __saturatef(__fmaf_rn(bdt, 16.0f, c.lo.y));
            c.hi.x = __saturatef(__fmaf_rn(bdt, 40.0f, c.hi.x));
            c.hi.y = __saturatef(__fmaf_rn(bdt, 40.0f, c.hi.y));
        }
    }

    const float2 at2 = make_float2(atten, atten);
    u.lo = mul2(u.lo, at2);
    c.lo = mul2(c.lo, at2);
    c.hi = mul2(c.hi, at2);
    Texels out;
    out.vel = pack_texel4(u.lo.x, u.lo.y, u.hi.x * atten, 0.0f);
    out.col = pack_texel4(c.lo.x, c.lo.y, c.hi.x, c.hi.y);
    return out;
}

// FUSED: the multi-GPU instantiation with fused halos (common.cuh PeerView); the single-GPU one carries none of it.
template <bool FUSED>
__global__ void __launch_bounds__(256, 4)
advect_kernel(const __grid_constant__ AdvectArgs A, const __grid_constant__ PeerView pv,
              const FrameParams* __restrict__ frame, const uint2* __restrict__ vel_in, uint2* col0,
              uint2* col1,  // m_colors[0], m_colors[1]
              uint2* __restrict__ vel_out, StepState* __restrict__ state) {
    const Domain& d = A.d;
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    // (fused halos: the chunks next to an interior face come last, see face_last_chunk)
    const int zc = FUSED ? face_last_chunk(pv, blockIdx.z, gridDim.z, kZ, A.reach, d.z_own1 - d.z_own0) : (int)blockIdx.z;
    const int z0 = d.z_own0 + zc * kZ;  // global plane
    const int z1 = min(z0 + kZ, d.z_own1);
    // fused halos: the planes next to an interior face wait for the neighbour's previous frame (event m = 0)
    const bool near_lo = FUSED && pv.has_lo && z0 < d.z_own0 + A.reach;
    const bool near_hi = FUSED && pv.has_hi && z1 > d.z_own1 - A.reach;
    if (FUSED && (near_lo || near_hi)) {
        if (threadIdx.x == 0 && threadIdx.y == 0) peer_wait(pv, frame->epoch_base, near_lo, near_hi);
        __syncthreads();
    }
    if (x >= d.nx || y >= d.ny) return;

    const float dt = frame->dt;
    const int parity = frame->parity;
    const uint2* __restrict__ col_in = parity ? col0 : col1;  // colour[!parity] (Fluid.cpp:372)
    uint2* __restrict__ col_out = parity ? col1 : col0;       // colour[parity]
    const float px = __ldg(A.tab.pos[0] + x), py = __ldg(A.tab.pos[1] + y);
    const float atten = fmaxf(__fmaf_rn(-dt, 0.200000003f, 1.0f), 0.0f);
    const unsigned plane = (unsigned)d.nx * d.ny;
    unsigned self = ((unsigned)(z0 - d.z_first) * d.ny + y) * d.nx + x;
    // The own texels of all kZ planes are requested at once: 64 bytes in flight per thread — the kernel is a stream
    // over most of a large grid, and a stream needs the bytes in flight more than anything else.
    // (The field is written by the neighbours between frames with fused halos: no non-coherent loads of halo planes.)
    uint2 sv[kZ], sc[kZ];
#pragma unroll
    for (int k = 0; k < kZ; ++k) {
        sv[k] = sc[k] = make_uint2(0u, 0u);
        if (z0 + k < z1) {
            sv[k] = __ldg(vel_in + self + k * plane);
            sc[k] = __ldg(col_in + self + k * plane);
        }
    }
    // Rest shortcut.  A voxel whose velocity texel is all +0 back-traces onto its own texel centre wherever
    // fma(pos, N, -0.5) reproduces the index (the `still` tables; always for power-of-two grids), so both fetches return
    // the voxel's own texels: the new velocity is +0 * atten = +0 and the new colour its own colour * atten — exactly
    // what advect_voxel computes for it, minus the trace.  Texels holding a -0 and voxels in the emitter's box take the
    // general path.  This is the quiescent far field: most of a large grid.
    const bool still_xy = dt > 0.0f && __ldg(A.tab.still[0] + x) != 0.0f && __ldg(A.tab.still[1] + y) != 0.0f;
    const bool box_xy = x >= A.em.x0 && x < A.em.x1 && y >= A.em.y0 && y < A.em.y1;
    const float2 at2 = make_float2(atten, atten);
    auto neg_zero = [](unsigned w) { return (w & 0xffffu) == 0x8000u || (w >> 16) == 0x8000u; };
#pragma unroll
    for (int k = 0; k < kZ; ++k) {
        const int z = z0 + k;
        if (z >= z1) break;
        Texels t;
        const bool rest = still_xy && (sv[k].x | sv[k].y) == 0u && __ldg(A.tab.still[2] + z) != 0.0f && z >= A.zv0 &&
                          z < A.zv1 - 1 && !(box_xy && z >= A.em.z0 && z < A.em.z1) && !neg_zero(sc[k].x) &&
                          !neg_zero(sc[k].y);
        if (rest) {
            Pair4 c = widen(sc[k]);
            c.lo = mul2(c.lo, at2);
            c.hi = mul2(c.hi, at2);
            t.vel = make_uint2(0u, 0u);
            t.col = pack_texel4(c.lo.x, c.lo.y, c.hi.x, c.hi.y);
        } else {
            t = advect_voxel(A, dt, atten, vel_in, col_in, state, x, y, z, px, py, self, sv[k], sc[k]);
        }
        vel_out[self] = t.vel;
        col_out[self] = t.col;
        if (FUSED && near_lo) {
            const long long at = (long long)self + (long long)pv.dz_lo * plane;
            if (z == d.z_own0) A.vel_out_lo[at] = t.vel;
            if (z < d.z_own0 + A.reach) A.col_lo[parity][at] = t.col;
        }
        if (FUSED && near_hi) {
            const long long at = (long long)self + (long long)pv.dz_hi * plane;
            if (z == d.z_own1 - 1) A.vel_out_hi[at] = t.vel;
            if (z >= d.z_own1 - A.reach) A.col_hi[parity][at] = t.col;
        }
        self += plane;
    }
}

}  // namespace

void launch_advect(const Domain& d, const AxisTables& tab, const FrameParams* frame, const void* vel_in,
                   void* const col[2], void* vel_out, const Emitter& em, int clamp_mode, StepState* state, int h_adv,
                   const PeerView& pv, const AdvectPeers& peers, cudaStream_t stream) {
    AdvectArgs A;
    A.d = d; A.tab = tab; A.em = em; A.clamp_mode = clamp_mode;
    // valid input planes: the owned ones plus the h_adv + 1 exchanged on each interior face (clipped to the grid)
    A.zv0 = d.z_own0 > 0 ? (d.z_own0 - (h_adv + 1) > 0 ? d.z_own0 - (h_adv + 1) : 0) : 0;
    A.zv1 = d.z_own1 < d.nz ? (d.z_own1 + h_adv + 1 < d.nz ? d.z_own1 + h_adv + 1 : d.nz) : d.nz;
    if (A.zv0 < d.z_first) A.zv0 = d.z_first;
    if (A.zv1 > d.z_first + d.nz_alloc) A.zv1 = d.z_first + d.nz_alloc;
    A.reach = h_adv + 1;
    A.vel_out_lo = (uint2*)peers.vel_out[0]; A.vel_out_hi = (uint2*)peers.vel_out[1];
    for (int i = 0; i < 2; ++i) {
        A.col_lo[i] = (uint2*)peers.col[0][i];
        A.col_hi[i] = (uint2*)peers.col[1][i];
    }
    const dim3 block(32, 8, 1);
    const dim3 grid((d.nx + 31) / 32, (d.ny + 7) / 8, (d.z_own1 - d.z_own0 + kZ - 1) / kZ);
    if (pv.has_lo || pv.has_hi)
        advect_kernel<true><<<grid, block, 0, stream>>>(A, pv, frame, (const uint2*)vel_in, (uint2*)col[0], (uint2*)col[1],
                                                        (uint2*)vel_out, state);
    else
        advect_kernel<false><<<grid, block, 0, stream>>>(A, pv, frame, (const uint2*)vel_in, (uint2*)col[0],
                                                         (uint2*)col[1], (uint2*)vel_out, state);
}

}  // namespace fxb
