// advect_body.cuh — interior path of the second advection kernel (advect.cu, FXB_ADVECT=2; experimental).
//
// Same arithmetic as advect_kernel (CSAdvect.hlsl:41-79 in the DXBC's operation order) for a voxel
// whose eight taps all lie inside the grid and the local slab — the only case handled here; everything else stays
// with the first kernel's code.  What differs is the instruction stream, which is what bounds the kernel
// (profiles/README.md: 283 instructions per voxel, issue-bound at a third of the HBM roofline):
//   * the taps are addressed as four row pointers per field with the x+1 tap at an immediate offset;
//   * all sixteen texels are requested before the first one is converted;
//   * velocity .w is never converted (it is a don't-care of the reference);
//   * a voxel whose eight colour taps are all +0 (no smoke anywhere near: 60-80 % of the moving voxels of a developed
//     flow, tools measurement in DESIGN.md §7) skips the colour converts and lerps — every lerp is
//     fma(f, 0 - 0, 0) = +0 for the finite weights of the interior path.
// Written against a small portability layer so that tests/emu/advect_emu.cpp can run the same statements on the CPU
// against the oracle (test infrastructure; never part of libfluidx_b200.so).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda_fp16.h>
#define FXA_FN __device__ __forceinline__
namespace fxb {
typedef float2 AF2;
typedef uint2 AU2;
FXA_FN AF2 fxa_h2f(unsigned w) { return __half22float2(*reinterpret_cast<const __half2*>(&w)); }
FXA_FN float fxa_hlo2f(unsigned w) { return __low2float(*reinterpret_cast<const __half2*>(&w)); }
FXA_FN unsigned fxa_f2h(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const unsigned*>(&h);
}
FXA_FN AF2 fxa_sub2(AF2 a, AF2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
FXA_FN AF2 fxa_mul2(AF2 a, AF2 b) { return __fmul2_rn(a, b); }
FXA_FN AF2 fxa_fma2(AF2 a, AF2 b, AF2 c) { return __ffma2_rn(a, b, c); }
FXA_FN float fxa_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
FXA_FN float fxa_sat(float v) { return __saturatef(v); }
FXA_FN AU2 fxa_ldg8(const AU2* p) { return __ldg(p); }
FXA_FN float fxa_ldgf(const float* p) { return __ldg(p); }
FXA_FN AF2 fxa_f2(float x, float y) { return make_float2(x, y); }
}  // namespace fxb
#else
#include <cmath>
#define FXA_FN inline
namespace fxb {
struct AF2 { float x, y; };
struct alignas(8) AU2 { unsigned x, y; };
inline float fxa_half_bits_to_float(unsigned short h) {
    _Float16 v;
    __builtin_memcpy(&v, &h, 2);
    return (float)v;
}
inline unsigned short fxa_float_to_half_bits(float f) {
    const _Float16 v = (_Float16)f;  // round to nearest even, like cvt.rn.f16.f32
    unsigned short h;
    __builtin_memcpy(&h, &v, 2);
    return h;
}
FXA_FN AF2 fxa_h2f(unsigned w) { return AF2{fxa_half_bits_to_float((unsigned short)(w & 0xffffu)), fxa_half_bits_to_float((unsigned short)(w >> 16))}; }
FXA_FN float fxa_hlo2f(unsigned w) { return fxa_half_bits_to_float((unsigned short)(w & 0xffffu)); }
FXA_FN unsigned fxa_f2h(float a, float b) { return (unsigned)fxa_float_to_half_bits(a) | ((unsigned)fxa_float_to_half_bits(b) << 16); }
FXA_FN AF2 fxa_sub2(AF2 a, AF2 b) { return AF2{a.x + -b.x, a.y + -b.y}; }
FXA_FN AF2 fxa_mul2(AF2 a, AF2 b) { return AF2{a.x * b.x, a.y * b.y}; }
FXA_FN AF2 fxa_fma2(AF2 a, AF2 b, AF2 c) { return AF2{std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
FXA_FN float fxa_fma(float a, float b, float c) { return std::fmaf(a, b, c); }
FXA_FN float fxa_sat(float v) { return std::fmin(std::fmax(v, 0.0f), 1.0f); }
FXA_FN AU2 fxa_ldg8(const AU2* p) { return *p; }
FXA_FN float fxa_ldgf(const float* p) { return *p; }
FXA_FN AF2 fxa_f2(float x, float y) { return AF2{x, y}; }
}  // namespace fxb
#endif

namespace fxb {

// Everything the interior path needs (uniform over the launch).
struct AdvectGeom {
    int nx, ny, nz;           // global grid
    int z_first, nz_alloc;    // local slab: plane 0 of the arrays is global plane z_first
    const float* pos[3];      // (i + 0.5) / N per axis, indexed by the global coordinate
    int ex0, ey0, ez0, ex1, ey1, ez1;  // emitter box
    const float* basis;       // emitter table
};

FXA_FN AF2 fxa_lerp2(AF2 f, AF2 a, AF2 b) { return fxa_fma2(f, fxa_sub2(b, a), a); }
FXA_FN float fxa_lerp(float f, float a, float b) { return fxa_fma(f, b + -a, a); }

// x, then y, then z; t[k]: tap k = x + 2 y + 4 z
FXA_FN AF2 fxa_trilerp2(const AF2 (&t)[8], AF2 wx, AF2 wy, AF2 wz) {
    const AF2 x00 = fxa_lerp2(wx, t[0], t[1]), x10 = fxa_lerp2(wx, t[2], t[3]);
    const AF2 x01 = fxa_lerp2(wx, t[4], t[5]), x11 = fxa_lerp2(wx, t[6], t[7]);
    return fxa_lerp2(wz, fxa_lerp2(wy, x00, x10), fxa_lerp2(wy, x01, x11));
}
FXA_FN float fxa_trilerp(const float (&t)[8], float fx, float fy, float fz) {
    const float x00 = fxa_lerp(fx, t[0], t[1]), x10 = fxa_lerp(fx, t[2], t[3]);
    const float x01 = fxa_lerp(fx, t[4], t[5]), x11 = fxa_lerp(fx, t[6], t[7]);
    return fxa_lerp(fz, fxa_lerp(fy, x00, x10), fxa_lerp(fy, x01, x11));
}

// One voxel (x, y, z global) on the interior path.  Returns false — and writes nothing — when a tap would leave the
// grid or the local slab, or a coordinate is not finite: the caller then runs the general code.
FXA_FN bool advect_interior_voxel(const AdvectGeom& g, float dt, const AU2* __restrict__ vel_in,
                                  const AU2* __restrict__ col_in, AU2* __restrict__ vel_out,
                                  AU2* __restrict__ col_out, int x, int y, int z) {
    const float px = fxa_ldgf(g.pos[0] + x), py = fxa_ldgf(g.pos[1] + y), pz = fxa_ldgf(g.pos[2] + z);
    const float fnx = (float)g.nx, fny = (float)g.ny, fnz = (float)g.nz;
    const unsigned self = ((unsigned)(z - g.z_first) * g.ny + y) * g.nx + x;
    const AU2 u0r = fxa_ldg8(vel_in + self);
    const AF2 u0xy = fxa_h2f(u0r.x);
    const float u0z = fxa_hlo2f(u0r.y);
    const float tx = fxa_fma(fxa_fma(-u0xy.x, dt, px), fnx, -0.5f);
    const float ty = fxa_fma(fxa_fma(-u0xy.y, dt, py), fny, -0.5f);
    const float tz = fxa_fma(fxa_fma(-u0z, dt, pz), fnz, -0.5f);
    const float zlo = (float)g.z_first, zhi = (float)(g.z_first + g.nz_alloc - 1);
    const bool inside = tx >= 0.0f && tx < fnx - 1.0f && ty >= 0.0f && ty < fny - 1.0f && tz >= 0.0f &&
                        tz < fnz - 1.0f && tz >= zlo && tz < zhi;
    if (!inside) return false;
    const float flx = floorf(tx), fly = floorf(ty), flz = floorf(tz);
    const float fx = tx - flx, fy = ty - fly, fz = tz - flz;
    const unsigned plane = (unsigned)g.nx * g.ny;
    const unsigned base = ((unsigned)((int)flz - g.z_first) * g.ny + (unsigned)(int)fly) * g.nx + (unsigned)(int)flx;
    // four row pointers per field; the x+1 tap is [1]
    const AU2* v00 = vel_in + base;
    const AU2* v10 = v00 + g.nx;
    const AU2* v01 = v00 + plane;
    const AU2* v11 = v01 + g.nx;
    const AU2* c00 = col_in + base;
    const AU2* c10 = c00 + g.nx;
    const AU2* c01 = c00 + plane;
    const AU2* c11 = c01 + g.nx;

    AF2 uxy, cxy, czw;
    float uz;
    // Exact-texel case: the back-trace lands on a texel centre, so every lerp is fma(0, b - a, a) = a and two loads
    // replace sixteen — except for a = -0, whose sum takes the sign of 0 * (b - a): such texels need the neighbours.
    bool exact = fx == 0.0f && fy == 0.0f && fz == 0.0f;
    if (exact) {
        const AU2 rv = fxa_ldg8(v00), rc = fxa_ldg8(c00);
        const unsigned w[4] = {rv.x, rv.y, rc.x, rc.y};
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if ((w[k] & 0xffffu) == 0x8000u || (w[k] >> 16) == 0x8000u) exact = false;
        uxy = fxa_h2f(rv.x);
        uz = fxa_hlo2f(rv.y);
        cxy = fxa_h2f(rc.x);
        czw = fxa_h2f(rc.y);
    }
    if (!exact) {
        // all sixteen texels in flight before the first convert
        AU2 rv[8], rc[8];
        rv[0] = fxa_ldg8(v00); rv[1] = fxa_ldg8(v00 + 1); rv[2] = fxa_ldg8(v10); rv[3] = fxa_ldg8(v10 + 1);
        rv[4] = fxa_ldg8(v01); rv[5] = fxa_ldg8(v01 + 1); rv[6] = fxa_ldg8(v11); rv[7] = fxa_ldg8(v11 + 1);
        rc[0] = fxa_ldg8(c00); rc[1] = fxa_ldg8(c00 + 1); rc[2] = fxa_ldg8(c10); rc[3] = fxa_ldg8(c10 + 1);
        rc[4] = fxa_ldg8(c01); rc[5] = fxa_ldg8(c01 + 1); rc[6] = fxa_ldg8(c11); rc[7] = fxa_ldg8(c11 + 1);
        const AF2 wx = fxa_f2(fx, fx), wy = fxa_f2(fy, fy), wz = fxa_f2(fz, fz);
        {
            AF2 t[8];
            float tzc[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                t[k] = fxa_h2f(rv[k].x);
                tzc[k] = fxa_hlo2f(rv[k].y);
            }
            uxy = fxa_trilerp2(t, wx, wy, wz);
            uz = fxa_trilerp(tzc, fx, fy, fz);
        }
        unsigned any = 0u;
#pragma unroll
        for (int k = 0; k < 8; ++k) any |= rc[k].x | rc[k].y;
        if (any == 0u) {  // eight +0 texels: the fetch is +0 in every channel
            cxy = fxa_f2(0.0f, 0.0f);
            czw = fxa_f2(0.0f, 0.0f);
        } else {
            AF2 t[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = fxa_h2f(rc[k].x);
            cxy = fxa_trilerp2(t, wx, wy, wz);
#pragma unroll
            for (int k = 0; k < 8; ++k) t[k] = fxa_h2f(rc[k].y);
            czw = fxa_trilerp2(t, wx, wy, wz);
        }
    }
    // Emitter (CSAdvect.hlsl:57-68).  Outside the table's box the basis is below exp(-4) by construction.
    if (x >= g.ex0 && x < g.ex1 && y >= g.ey0 && y < g.ey1 && z >= g.ez0 && z < g.ez1) {
        const float basis =
            fxa_ldgf(g.basis + ((size_t)(z - g.ez0) * (g.ey1 - g.ey0) + (y - g.ey0)) * (g.ex1 - g.ex0) + (x - g.ex0));
        if (basis >= 0.0183156393f) {
            float fx_, fy_, fz_;
            if (1.0f < fnz) {
                const float dx = px + -0.5f, dz = pz + -0.5f;
                fx_ = fxa_fma(basis, 0.0f, dz * -200.0f);
                fy_ = fxa_fma(basis, 192.0f, 0.0f);
                fz_ = fxa_fma(basis, 0.0f, dx * 200.0f);
            } else {
                fx_ = 0.0f; fy_ = basis * 48.0f; fz_ = 0.0f;
            }
            uxy.x = fxa_fma(fx_, dt, uxy.x);
            uxy.y = fxa_fma(fy_, dt, uxy.y);
            uz = fxa_fma(fz_, dt, uz);
            const float bdt = basis * dt;
            cxy.x = fxa_sat(fxa_fma(bdt, 8.0f, cxy.x));
            cxy.y = fxa_sat(fxa_fma(bdt, 16.0f, cxy.y));
            czw.x = fxa_sat(fxa_fma(bdt, 40.0f, czw.x));
            czw.y = fxa_sat(fxa_fma(bdt, 40.0f, czw.y));
        }
    }
    const float atten = fmaxf(fxa_fma(-dt, 0.200000003f, 1.0f), 0.0f);
    const AF2 at2 = fxa_f2(atten, atten);
    uxy = fxa_mul2(uxy, at2);
    cxy = fxa_mul2(cxy, at2);
    czw = fxa_mul2(czw, at2);
    AU2 ov, oc;
    ov.x = fxa_f2h(uxy.x, uxy.y);
    ov.y = fxa_f2h(uz * atten, 0.0f);
    oc.x = fxa_f2h(cxy.x, cxy.y);
    oc.y = fxa_f2h(czw.x, czw.y);
    vel_out[self] = ov;
    col_out[self] = oc;
    return true;
}

}  // namespace fxb
