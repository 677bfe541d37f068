// advect.cu — CSAdvect as a software fp32 trilinear back-trace kernel (sm_100a).
//
// Replaces FluidX12/Content/Shaders/CSAdvect.hlsl:41-79 (dispatch Fluid.cpp:374).  Operation order
// follows the shipped DXBC; the two hardware SampleLevel fetches become 16
// gathered 8-byte texel loads with fp32 weights (no texture unit, no 8-bit weights).
//
// Mapping: one thread per voxel, CTA = 32 x 4 x 4 voxels so that the 33 x 5 x 5 tap footprint of a
// CTA is shared through L1 (the back-trace is spatially coherent).  Algorithmic traffic: 32 B/voxel
// (velocity in 8 + colour in 8 + velocity out 8 + colour out 8); HBM-bound.
#include "common.cuh"
#include "kernels.h"

namespace fxb {

namespace {

struct Taps {
    size_t o[8];  // texel offsets of the 8 taps, order (x0|x1) fastest, then y, then z
    float fx, fy, fz;
};

__device__ __forceinline__ float lerp3(float a000, float a100, float a010, float a110, float a001, float a101,
                                       float a011, float a111, float fx, float fy, float fz) {
    const float x00 = __fmaf_rn(fx, a100 - a000, a000);
    const float x10 = __fmaf_rn(fx, a110 - a010, a010);
    const float x01 = __fmaf_rn(fx, a101 - a001, a001);
    const float x11 = __fmaf_rn(fx, a111 - a011, a011);
    const float y0 = __fmaf_rn(fy, x10 - x00, x00);
    const float y1 = __fmaf_rn(fy, x11 - x01, x01);
    return __fmaf_rn(fz, y1 - y0, y0);
}

__device__ __forceinline__ float4 gather4(const uint2* __restrict__ f, const Taps& t) {
    float4 a[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) a[k] = load_texel4(f, t.o[k]);
    float4 r;
    r.x = lerp3(a[0].x, a[1].x, a[2].x, a[3].x, a[4].x, a[5].x, a[6].x, a[7].x, t.fx, t.fy, t.fz);
    r.y = lerp3(a[0].y, a[1].y, a[2].y, a[3].y, a[4].y, a[5].y, a[6].y, a[7].y, t.fx, t.fy, t.fz);
    r.z = lerp3(a[0].z, a[1].z, a[2].z, a[3].z, a[4].z, a[5].z, a[6].z, a[7].z, t.fx, t.fy, t.fz);
    r.w = lerp3(a[0].w, a[1].w, a[2].w, a[3].w, a[4].w, a[5].w, a[6].w, a[7].w, t.fx, t.fy, t.fz);
    return r;
}

}  // namespace

__global__ void __launch_bounds__(512) advect_kernel(Domain d, const FrameParams* __restrict__ frame,
                                                     const uint2* __restrict__ vel_in,
                                                     uint2* col0, uint2* col1,  // m_colors[0], m_colors[1]
                                                     uint2* __restrict__ vel_out, Emitter em, int clamp_mode,
                                                     StepState* __restrict__ state) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int z = d.z_own0 + blockIdx.z * 4 + threadIdx.z;  // global plane
    if (x >= d.nx || y >= d.ny || z >= d.z_own1) return;

    const float dt = frame->dt;
    const int parity = frame->parity;
    const uint2* __restrict__ col_in = parity ? col0 : col1;  // colour[!parity] (Fluid.cpp:372)
    uint2* __restrict__ col_out = parity ? col1 : col0;       // colour[parity]

    const float fnx = (float)d.nx, fny = (float)d.ny, fnz = (float)d.nz;
    const float px = ((float)x + 0.5f) / fnx;
    const float py = ((float)y + 0.5f) / fny;
    const float pz = ((float)z + 0.5f) / fnz;

    const size_t self = ((size_t)(z - d.z_first) * d.ny + y) * d.nx + x;
    const float4 u0 = load_texel4(vel_in, self);
    const float ax = __fmaf_rn(-u0.x, dt, px);
    const float ay = __fmaf_rn(-u0.y, dt, py);
    const float az = __fmaf_rn(-u0.z, dt, pz);

    Taps t;
    {
        const float tx = __fmaf_rn(ax, fnx, -0.5f);
        const float ty = __fmaf_rn(ay, fny, -0.5f);
        const float tz = __fmaf_rn(az, fnz, -0.5f);
        const int ix = floor_to_tap(tx), iy = floor_to_tap(ty), iz = floor_to_tap(tz);
        t.fx = tx - floorf(tx);
        t.fy = ty - floorf(ty);
        t.fz = tz - floorf(tz);
        const int x0 = address_tap(ix, d.nx, clamp_mode), x1 = address_tap(ix + 1, d.nx, clamp_mode);
        const int y0 = address_tap(iy, d.ny, clamp_mode), y1 = address_tap(iy + 1, d.ny, clamp_mode);
        int z0 = address_tap(iz, d.nz, clamp_mode) - d.z_first;
        int z1 = address_tap(iz + 1, d.nz, clamp_mode) - d.z_first;
        if ((unsigned)z0 >= (unsigned)d.nz_alloc || (unsigned)z1 >= (unsigned)d.nz_alloc) {
            // The back-trace left the exchanged halo (multi-GPU only): flag it, keep addresses legal.
            state->halo_overflow = 1;
            z0 = min(max(z0, 0), d.nz_alloc - 1);
            z1 = min(max(z1, 0), d.nz_alloc - 1);
        }
        const size_t r00 = ((size_t)z0 * d.ny + y0) * d.nx, r10 = ((size_t)z0 * d.ny + y1) * d.nx;
        const size_t r01 = ((size_t)z1 * d.ny + y0) * d.nx, r11 = ((size_t)z1 * d.ny + y1) * d.nx;
        t.o[0] = r00 + x0; t.o[1] = r00 + x1; t.o[2] = r10 + x0; t.o[3] = r10 + x1;
        t.o[4] = r01 + x0; t.o[5] = r01 + x1; t.o[6] = r11 + x0; t.o[7] = r11 + x1;
    }
    float4 u = gather4(vel_in, t);
    float4 c = gather4(col_in, t);

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
            u.x = __fmaf_rn(fx_, dt, u.x);
            u.y = __fmaf_rn(fy_, dt, u.y);
            u.z = __fmaf_rn(fz_, dt, u.z);
            const float bdt = basis * dt;
            c.x = __saturatef(__fmaf_rn(bdt, 8.0f, c.x));
            c.y = __saturatef(__fmaf_rn(bdt, 16.0f, c.y));
            c.z = __saturatef(__fmaf_rn(bdt, 40.0f, c.z));
            c.w = __saturatef(__fmaf_rn(bdt, 40.0f, c.w));
        }
    }

    const float atten = fmaxf(__fmaf_rn(-dt, 0.200000003f, 1.0f), 0.0f);
    vel_out[self] = pack_texel4(u.x * atten, u.y * atten, u.z * atten, 0.0f);
    col_out[self] = pack_texel4(c.x * atten, c.y * atten, c.z * atten, c.w * atten);
}

void launch_advect(const Domain& d, const FrameParams* frame, const void* vel_in, void* const col[2], void* vel_out,
                   const Emitter& em, int clamp_mode, StepState* state, cudaStream_t stream) {
    const dim3 block(32, 4, 4);
    const dim3 grid((d.nx + 31) / 32, (d.ny + 3) / 4, (d.z_own1 - d.z_own0 + 3) / 4);
    advect_kernel<<<grid, block, 0, stream>>>(d, frame, (const uint2*)vel_in, (uint2*)col[0], (uint2*)col[1],
                                              (uint2*)vel_out, em, clamp_mode, state);
}

}  // namespace fxb
