// advect.cu — CSAdvect as a software fp32 trilinear back-trace kernel (sm_100a).
//
// Replaces FluidX12/Content/Shaders/CSAdvect.hlsl:41-79 (dispatch Fluid.cpp:374).  Operation order
// follows the shipped DXBC; the two hardware SampleLevel fetches become 16
// gathered 8-byte texel loads blended with fp32 weights (no texture unit, no 8-bit weights).
//
// Mapping: one thread per voxel, CTA = 32 x 4 x 4 voxels so that the 33 x 5 x 5 tap footprint of a
// CTA is shared through L1 (the back-trace is spatially coherent).  The kernel is bounded by
// instruction issue before HBM, so the hot path is kept lean:
//   * (i + 0.5) / N comes from per-axis tables (three IEEE divisions per voxel otherwise);
//   * taps that all fall inside the grid (the common case) take a fast path with 32-bit offsets and no
//     sampler-addressing arithmetic; MIRROR/CLAMP wrapping lives in a cold out-of-line path;
//   * the 14 lerps of the 7 used channels run as packed FADD2/FFMA2 on (x,y) and (z,w) pairs;
//   * a back-trace that lands exactly on a texel centre (still-quiescent voxels) needs 2 loads, not 16.
// Algorithmic traffic: 32 B/voxel (velocity in 8 + colour in 8 + velocity out 8 + colour out 8).
#include "advect_body.cuh"
#include "common.cuh"
#include "kernels.h"

namespace fxb {

namespace {

struct Pair4 {  // one RGBA16F texel widened to fp32 as two packed pairs
    float2 lo, hi;
};

__device__ __forceinline__ Pair4 load_pairs(const uint2* __restrict__ f, unsigned i) {
    const uint2 r = __ldg(f + i);
    Pair4 t;
    t.lo = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    t.hi = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return t;
}

__device__ __forceinline__ float2 lerp2(float2 f, float2 a, float2 b) { return fma2(f, sub2(b, a), a); }

// x, then y, then z; each lerp is fma(f, b - a, a)
__device__ __forceinline__ Pair4 gather(const uint2* __restrict__ f, const unsigned (&o)[8], float fx, float fy,
                                        float fz) {
    Pair4 t[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) t[k] = load_pairs(f, o[k]);
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

}  // namespace

// WHAT: 3 = velocity and colour in one pass; 1 = velocity only; 2 = colour only.  The colour field is not an
// input of the projection, so the step advects it on a side branch of the graph, under the Jacobi passes.
enum { kAdvectVelocity = 1, kAdvectColour = 2, kAdvectBoth = 3, kAdvectSecondKernel = 4 };

// One voxel, any case (taps inside or outside the grid, either sampler addressing mode).
template <int WHAT>
__device__ __forceinline__ void advect_voxel(const Domain& d, const AxisTables& tab, const float dt,
                                             const uint2* __restrict__ vel_in, const uint2* __restrict__ col_in,
                                             uint2* __restrict__ col_out, uint2* __restrict__ vel_out, const Emitter& em,
                                             const int clamp_mode, StepState* __restrict__ state, const int x,
                                             const int y, const int z) {
    const float px = __ldg(tab.pos[0] + x), py = __ldg(tab.pos[1] + y), pz = __ldg(tab.pos[2] + z);
    const float fnx = (float)d.nx, fny = (float)d.ny, fnz = (float)d.nz;

    const unsigned self = ((unsigned)(z - d.z_first) * d.ny + y) * d.nx + x;
    const Pair4 u0 = load_pairs(vel_in, self);
    const float tx = __fmaf_rn(__fmaf_rn(-u0.lo.x, dt, px), fnx, -0.5f);
    const float ty = __fmaf_rn(__fmaf_rn(-u0.lo.y, dt, py), fny, -0.5f);
    const float tz = __fmaf_rn(__fmaf_rn(-u0.hi.x, dt, pz), fnz, -0.5f);
    const float flx = floorf(tx), fly = floorf(ty), flz = floorf(tz);
    const float fx = tx - flx, fy = ty - fly, fz = tz - flz;

    unsigned o[8];
    const float zlo = (float)d.z_first, zhi = (float)(d.z_first + d.nz_alloc - 1);
    const bool inside = tx >= 0.0f && tx < fnx - 1.0f && ty >= 0.0f && ty < fny - 1.0f && tz >= 0.0f &&
                        tz < fnz - 1.0f && tz >= zlo && tz < zhi;
    if (inside) {  // both taps of every axis are inside the grid (and inside the local slab)
        const unsigned plane = (unsigned)d.nx * d.ny;
        const unsigned base = ((unsigned)((int)flz - d.z_first) * d.ny + (unsigned)(int)fly) * d.nx + (unsigned)(int)flx;
        o[0] = base; o[1] = base + 1; o[2] = base + d.nx; o[3] = base + d.nx + 1;
        o[4] = base + plane; o[5] = o[4] + 1; o[6] = o[4] + d.nx; o[7] = o[6] + 1;
    } else {
        const int2 xs = wrapped_pair(tx, d.nx, clamp_mode), ys = wrapped_pair(ty, d.ny, clamp_mode);
        int2 zs = wrapped_pair(tz, d.nz, clamp_mode);
        zs.x -= d.z_first;
        zs.y -= d.z_first;
        if ((unsigned)zs.x >= (unsigned)d.nz_alloc || (unsigned)zs.y >= (unsigned)d.nz_alloc) {
            state->halo_overflow = 1;  // the back-trace left the exchanged z-halo (multi-GPU only)
            zs.x = min(max(zs.x, 0), d.nz_alloc - 1);
            zs.y = min(max(zs.y, 0), d.nz_alloc - 1);
        }
        const unsigned r00 = ((unsigned)zs.x * d.ny + ys.x) * d.nx, r10 = ((unsigned)zs.x * d.ny + ys.y) * d.nx;
        const unsigned r01 = ((unsigned)zs.y * d.ny + ys.x) * d.nx, r11 = ((unsigned)zs.y * d.ny + ys.y) * d.nx;
        o[0] = r00 + xs.x; o[1] = r00 + xs.y; o[2] = r10 + xs.x; o[3] = r10 + xs.y;
        o[4] = r01 + xs.x; o[5] = r01 + xs.y; o[6] = r11 + xs.x; o[7] = r11 + xs.y;
    }
    // Exact-texel fast path.  When the back-trace lands exactly on a texel centre (all three weights are 0: in
    // practice the voxels the flow has not reached, u = 0) every lerp is fma(0, b - a, a) = a, i.e. the fetch
    // returns the first tap unchanged, so 2 loads replace 16.  The one case where fma(0, b - a, a) != a bitwise is
    // a = -0 (the sum takes the sign of 0 * (b - a)); such texels take the general path.
    constexpr bool kVel = (WHAT & kAdvectVelocity) != 0, kCol = (WHAT & kAdvectColour) != 0;
    Pair4 u, c;
    u.lo = u.hi = c.lo = c.hi = make_float2(0.0f, 0.0f);
    bool exact = inside && fx == 0.0f && fy == 0.0f && fz == 0.0f;
    if (exact) {
        uint2 rv = make_uint2(0u, 0u), rc = make_uint2(0u, 0u);
        if (kVel) rv = __ldg(vel_in + o[0]);
        if (kCol) rc = __ldg(col_in + o[0]);
        auto neg_zero = [](unsigned w) { return (w & 0xffffu) == 0x8000u || (w >> 16) == 0x8000u; };
        if (neg_zero(rv.x) || neg_zero(rv.y) || neg_zero(rc.x) || neg_zero(rc.y)) {
            exact = false;
        } else {
            u.lo = __half22float2(*reinterpret_cast<const __half2*>(&rv.x));
            u.hi = __half22float2(*reinterpret_cast<const __half2*>(&rv.y));
            c.lo = __half22float2(*reinterpret_cast<const __half2*>(&rc.x));
            c.hi = __half22float2(*reinterpret_cast<const __half2*>(&rc.y));
        }
    }
    if (!exact) {
        if (kVel) u = gather(vel_in, o, fx, fy, fz);
        if (kCol) c = gather(col_in, o, fx, fy, fz);
    }

    // Emitter (CSAdvect.hlsl:57-68).  Outside the table's box the basis is below exp(-4) by construction.
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
            if (kVel) {
                u.lo.x = __fmaf_rn(fx_, dt, u.lo.x);
                u.lo.y = __fmaf_rn(fy_, dt, u.lo.y);
                u.hi.x = __fmaf_rn(fz_, dt, u.hi.x);
            }
            if (kCol) {
                const float bdt = basis * dt;
                c.lo.x = __saturatef(__fmaf_rn(bdt, 8.0f, c.lo.x));
                c.lo.y = __saturatef(__fmaf_rn(bdt, 16.0f, c.lo.y));
                c.hi.x = __saturatef(__fmaf_rn(bdt, 40.0f, c.hi.x));
                c.hi.y = __saturatef(__fmaf_rn(bdt, 40.0f, c.hi.y));
            }
        }
    }

    const float atten = fmaxf(__fmaf_rn(-dt, 0.200000003f, 1.0f), 0.0f);
    const float2 at2 = make_float2(atten, atten);
    if (kVel) {
        u.lo = mul2(u.lo, at2);
        vel_out[self] = pack_texel4(u.lo.x, u.lo.y, u.hi.x * atten, 0.0f);
    }
    if (kCol) {
        c.lo = mul2(c.lo, at2);
        c.hi = mul2(c.hi, at2);
        col_out[self] = pack_texel4(c.lo.x, c.lo.y, c.hi.x, c.hi.y);
    }
}

template <int WHAT>
__global__ void __launch_bounds__(512, 2)
advect_kernel(Domain d, AxisTables tab, const FrameParams* __restrict__ frame, const uint2* __restrict__ vel_in,
              uint2* col0, uint2* col1,  // m_colors[0], m_colors[1]
              uint2* __restrict__ vel_out, Emitter em, int clamp_mode, StepState* __restrict__ state) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int z = d.z_own0 + blockIdx.z * 4 + threadIdx.z;  // global plane
    if (x >= d.nx || y >= d.ny || z >= d.z_own1) return;

    const float dt = frame->dt;
    const int parity = frame->parity;
    const uint2* __restrict__ col_in = parity ? col0 : col1;  // colour[!parity] (Fluid.cpp:372)
    uint2* __restrict__ col_out = parity ? col1 : col0;       // colour[parity]
    advect_voxel<WHAT>(d, tab, dt, vel_in, col_in, col_out, vel_out, em, clamp_mode, state, x, y, z);
}

// Second kernel (experimental, FXB_ADVECT=2): voxels whose taps all lie inside the grid — the bulk of every grid —
// take the leaner interior path of advect_body.cuh (row pointers, all loads in flight, no convert of velocity .w,
// colour-free shortcut); the shell of voxels whose back-trace reaches a face runs the code above, out of line.
template <int WHAT>
__device__ __noinline__ void advect_voxel_cold(const Domain& d, const AxisTables& tab, const float dt,
                                               const uint2* __restrict__ vel_in, const uint2* __restrict__ col_in,
                                               uint2* __restrict__ col_out, uint2* __restrict__ vel_out,
                                               const Emitter& em, const int clamp_mode, StepState* __restrict__ state,
                                               const int x, const int y, const int z) {
    advect_voxel<WHAT>(d, tab, dt, vel_in, col_in, col_out, vel_out, em, clamp_mode, state, x, y, z);
}

__global__ void __launch_bounds__(512, 2)
advect2_kernel(Domain d, AxisTables tab, const FrameParams* __restrict__ frame, const uint2* __restrict__ vel_in,
               uint2* col0, uint2* col1, uint2* __restrict__ vel_out, Emitter em, int clamp_mode,
               StepState* __restrict__ state) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int z = d.z_own0 + blockIdx.z * 4 + threadIdx.z;  // global plane
    if (x >= d.nx || y >= d.ny || z >= d.z_own1) return;
    const float dt = frame->dt;
    const int parity = frame->parity;
    const uint2* __restrict__ col_in = parity ? col0 : col1;  // colour[!parity] (Fluid.cpp:372)
    uint2* __restrict__ col_out = parity ? col1 : col0;       // colour[parity]
    AdvectGeom g;
    g.nx = d.nx; g.ny = d.ny; g.nz = d.nz; g.z_first = d.z_first; g.nz_alloc = d.nz_alloc;
    g.pos[0] = tab.pos[0]; g.pos[1] = tab.pos[1]; g.pos[2] = tab.pos[2];
    g.ex0 = em.x0; g.ey0 = em.y0; g.ez0 = em.z0; g.ex1 = em.x1; g.ey1 = em.y1; g.ez1 = em.z1;
    g.basis = em.basis;
    if (advect_interior_voxel(g, dt, vel_in, col_in, vel_out, col_out, x, y, z)) return;
    advect_voxel_cold<kAdvectBoth>(d, tab, dt, vel_in, col_in, col_out, vel_out, em, clamp_mode, state, x, y, z);
}

// what: 1 = velocity only, 2 = colour only, 3 = both, 3 | 4 = both with the second kernel (see the enum above)
void launch_advect(const Domain& d, const AxisTables& tab, const FrameParams* frame, const void* vel_in,
                   void* const col[2], void* vel_out, const Emitter& em, int clamp_mode, StepState* state, int what,
                   cudaStream_t stream) {
    const dim3 block(32, 4, 4);
    const dim3 grid((d.nx + 31) / 32, (d.ny + 3) / 4, (d.z_own1 - d.z_own0 + 3) / 4);
    auto* v = (const uint2*)vel_in;
    auto *c0 = (uint2*)col[0], *c1 = (uint2*)col[1], *vo = (uint2*)vel_out;
    if (what == (kAdvectBoth | kAdvectSecondKernel) && d.nz > 1)
        advect2_kernel<<<grid, block, 0, stream>>>(d, tab, frame, v, c0, c1, vo, em, clamp_mode, state);
    else if ((what & 3) == kAdvectVelocity)
        advect_kernel<kAdvectVelocity><<<grid, block, 0, stream>>>(d, tab, frame, v, c0, c1, vo, em, clamp_mode, state);
    else if ((what & 3) == kAdvectColour)
        advect_kernel<kAdvectColour><<<grid, block, 0, stream>>>(d, tab, frame, v, c0, c1, vo, em, clamp_mode, state);
    else
        advect_kernel<kAdvectBoth><<<grid, block, 0, stream>>>(d, tab, frame, v, c0, c1, vo, em, clamp_mode, state);
}

}  // namespace fxb
