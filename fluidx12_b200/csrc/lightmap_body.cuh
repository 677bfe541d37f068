// lightmap_body.cuh — one voxel of the light-map pass (lightmap.cu).
//
// Reference: FluidX12/Content/Shaders/CSRayMarchL.hlsl:15-80 with RayMarch.hlsli:62-68 (GetSample), :75-98
// (GetDensityGradient), :203-210 (LocalToTex3DSpace), :215-228 (GetStep), :233-268 (CastLightRay); dispatched by
// Fluid::rayMarchL (Fluid.cpp:857-878) over the colour field Fluid::Render binds, m_colors[m_frameParity].  The
// arithmetic follows the shipped Bin/CSRayMarchL.cso instruction by instruction (fused where the bytecode has `mad`,
// separately rounded elsewhere; the library is built with -fmad=false -prec-div=true -prec-sqrt=true), including the
// order-3 SH irradiance the blob inlines from XUSG's SHIrradiance.hlsli.  Platform semantics restated as in the oracle:
// LINEAR_CLAMP sampler (Fluid.cpp:475) = fp32-weight trilinear, t = fma(coord, W, -0.5), lerps x, y, z as
// fma(f, b - a, a), texel offsets added before clamping; rsq = 1 / sqrt; min16float carried out in fp32;
// R11G11B10_FLOAT store truncates toward zero.
//
// Only colour.w (the density) is ever sampled, so the kernel reads a compact half-precision copy of that channel
// (2 bytes per voxel instead of the 8-byte texel: four times as many taps per 32-byte sector).
// Written against a small portability layer so that tests/emu/lightmap_emu.cpp runs the same statements on the CPU
// against the oracle (test infrastructure; never part of libfluidx_b200.so).
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#include <cuda_fp16.h>
#define FXL_FN __device__ __forceinline__
namespace fxb {
FXL_FN float fxl_fma(float a, float b, float c) { return __fmaf_rn(a, b, c); }
FXL_FN float fxl_h2f(const unsigned short* p) { return __half2float(__ushort_as_half(__ldg(p))); }
FXL_FN float fxl_floor(float v) { return floorf(v); }
FXL_FN float fxl_abs(float v) { return fabsf(v); }
FXL_FN float fxl_min(float a, float b) { return fminf(a, b); }
FXL_FN float fxl_max(float a, float b) { return fmaxf(a, b); }
FXL_FN float fxl_rsq(float v) { return 1.0f / sqrtf(v); }
FXL_FN unsigned fxl_bits(float v) { return __float_as_uint(v); }
}  // namespace fxb
#else
#include <cmath>
#include <cstring>
#define FXL_FN inline
namespace fxb {
FXL_FN float fxl_fma(float a, float b, float c) { return std::fmaf(a, b, c); }
FXL_FN float fxl_h2f(const unsigned short* p) {
    _Float16 v;
    __builtin_memcpy(&v, p, 2);
    return (float)v;
}
FXL_FN float fxl_floor(float v) { return std::floor(v); }
FXL_FN float fxl_abs(float v) { return std::fabs(v); }
FXL_FN float fxl_min(float a, float b) { return std::fmin(a, b); }
FXL_FN float fxl_max(float a, float b) { return std::fmax(a, b); }
FXL_FN float fxl_rsq(float v) { return 1.0f / std::sqrt(v); }
FXL_FN unsigned fxl_bits(float v) {
    unsigned u;
    __builtin_memcpy(&u, &v, 4);
    return u;
}
}  // namespace fxb
#endif

namespace fxb {

// = fxb_light_params (include/fluidx_b200.h), the constants CSRayMarchL reads
struct LightConsts {
    float light_pt[3];
    float light_color[4];
    float ambient[4];
    float world_i[12];
    float world[12];
    unsigned num_samples;
    unsigned has_light_probes;
    float sh[9][3];
};

struct LightGeom {
    int nx, ny, nz;
};

FXL_FN int fxl_tap(float t) {  // floor(t) saturated to +-2^30 (NaN -> -2^30), as the oracle's floor_to_tap
    const float lim = 1073741824.0f;
    if (!(t > -lim)) return -(1 << 30);
    if (t > lim) return 1 << 30;
    return (int)fxl_floor(t);
}
FXL_FN int fxl_clamp(int i, int w) { return i < 0 ? 0 : (i > w - 1 ? w - 1 : i); }

// colour.w at the normalised coordinate (cx, cy, cz), taps shifted by (ox, oy, oz) texels, LINEAR_CLAMP
FXL_FN float fxl_density(const unsigned short* __restrict__ dens, const LightGeom& g, const float cx, const float cy,
                         const float cz, const int ox, const int oy, const int oz) {
    const float tx = fxl_fma(cx, (float)g.nx, -0.5f), ty = fxl_fma(cy, (float)g.ny, -0.5f);
    const float tz = fxl_fma(cz, (float)g.nz, -0.5f);
    const int ix = fxl_tap(tx) + ox, iy = fxl_tap(ty) + oy, iz = fxl_tap(tz) + oz;
    const float fx = tx - fxl_floor(tx), fy = ty - fxl_floor(ty), fz = tz - fxl_floor(tz);
    const int x0 = fxl_clamp(ix, g.nx), x1 = fxl_clamp(ix + 1, g.nx);
    const int y0 = fxl_clamp(iy, g.ny), y1 = fxl_clamp(iy + 1, g.ny);
    const int z0 = fxl_clamp(iz, g.nz), z1 = fxl_clamp(iz + 1, g.nz);
    const unsigned short* r00 = dens + ((size_t)z0 * g.ny + y0) * g.nx;
    const unsigned short* r10 = dens + ((size_t)z0 * g.ny + y1) * g.nx;
    const unsigned short* r01 = dens + ((size_t)z1 * g.ny + y0) * g.nx;
    const unsigned short* r11 = dens + ((size_t)z1 * g.ny + y1) * g.nx;
    // all eight taps requested before the first lerp
    const float a000 = fxl_h2f(r00 + x0), a100 = fxl_h2f(r00 + x1), a010 = fxl_h2f(r10 + x0), a110 = fxl_h2f(r10 + x1);
    const float a001 = fxl_h2f(r01 + x0), a101 = fxl_h2f(r01 + x1), a011 = fxl_h2f(r11 + x0), a111 = fxl_h2f(r11 + x1);
    const float x00 = fxl_fma(fx, a100 - a000, a000), x10 = fxl_fma(fx, a110 - a010, a010);
    const float x01 = fxl_fma(fx, a101 - a001, a001), x11 = fxl_fma(fx, a111 - a011, a011);
    const float y0v = fxl_fma(fy, x10 - x00, x00), y1v = fxl_fma(fy, x11 - x01, x01);
    return fxl_fma(fz, y1v - y0v, y0v);
}

FXL_FN float fxl_dp3(const float a0, const float a1, const float a2, const float* b) {
    return (a0 * b[0] + a1 * b[1]) + a2 * b[2];
}

// CastLightRay as compiled (blob instructions 17-54 and 107-144): transmittance along `d` from `o`
FXL_FN float fxl_cast_ray(const unsigned short* __restrict__ dens, const LightGeom& g, const float o0, const float o1,
                          const float o2, const float d0, const float d1, const float d2, const float step,
                          const unsigned num_samples) {
    float transm = 1.0f, t = step, prev = 0.0f;
    for (unsigned i = 0; i < num_samples; ++i) {
        const float p0 = fxl_fma(d0, t, o0), p1 = fxl_fma(d1, t, o1), p2 = fxl_fma(d2, t, o2);
        if (1.0f < fxl_abs(p0) || 1.0f < fxl_abs(p1) || 1.0f < fxl_abs(p2)) break;
        const float d = fxl_density(dens, g, fxl_fma(p0, 0.5f, 0.5f), fxl_fma(p1, 0.5f, 0.5f), fxl_fma(p2, 0.5f, 0.5f),
                                    0, 0, 0);
        const float tr = fxl_fma(-d, 0.8f, 1.0f) * transm;   // transm *= 1 - density * ABSORPTION
        if (tr < 0.01f) return tr;                            // ZERO_THRESHOLD
        const float ev = fxl_min(0.00390625f / fxl_abs(-prev + d), 2.0f);  // GetStep: 1/256/|dDensity|, capped
        const float ui = fxl_min(-d + 1.0f, 1.0f);
        const float th = -transm + 1.0f;                      // the transmittance BEFORE this sample
        const float grow = fxl_max(th * (ui * (ev * 1.5f)), 1.0f);
        t = fxl_fma(step, grow, t);
        transm = tr;
        prev = d;
    }
    return transm;
}

// DXGI_FORMAT_R11G11B10_FLOAT: R bits 0-10, G 11-21, B 22-31; truncation toward zero (see the header comment)
FXL_FN unsigned fxl_pack_r11g11b10(const float r, const float g, const float b) {
    const float v[3] = {r, g, b};
    unsigned out = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const unsigned w = fxl_bits(v[k]);
        const unsigned sign = w >> 31, e = (w >> 23) & 0xFFu, m = w & 0x7FFFFFu;
        const int mb = k == 2 ? 5 : 6, drop = 23 - mb;
        const unsigned maxfin = (30u << mb) | ((1u << mb) - 1u);
        unsigned q;
        if (e == 0xFFu) q = m ? ((31u << mb) | ((1u << mb) - 1u)) : (sign ? 0u : (31u << mb));
        else if (sign) q = 0u;
        else if (e >= 143u) q = maxfin;
        else if (e >= 113u) q = ((e - 112u) << mb) | (m >> drop);
        else {
            const unsigned sh = 113u - e < 24u ? 113u - e : 24u;
            q = ((m | 0x800000u) >> sh) >> drop;
        }
        out |= q << (k == 0 ? 0 : (k == 1 ? 11 : 22));
    }
    return out;
}

// The light reaching the point o (volume space; u = its texture coordinate) where there is smoke — GetLight of
// RayMarch.hlsli:280-313 / the body of CSRayMarchL.hlsl:44-77: transmittance towards the light, and with light probes
// the occlusion ray along the density gradient times the SH irradiance instead of the constant ambient term.
// l0..l2: the normalised light direction in volume space; step / num_samples: of the light rays.
FXL_FN void fxl_light_at(const unsigned short* __restrict__ dens, const LightGeom& g, const LightConsts& P, const float o0,
                         const float o1, const float o2, const float u0, const float u1, const float u2, const float l0,
                         const float l1, const float l2, const float step, float& shadow, float& ao, float (&irr)[3]) {
    shadow = fxl_cast_ray(dens, g, o0, o1, o2, l0, l1, l2, step, P.num_samples);
    ao = 1.0f;
    irr[0] = irr[1] = irr[2] = 0.0f;
    if (P.has_light_probes) {
        const float q0 = fxl_density(dens, g, u0, u1, u2, -1, 0, 0), q1 = fxl_density(dens, g, u0, u1, u2, 1, 0, 0);
        const float q2 = fxl_density(dens, g, u0, u1, u2, 0, -1, 0), q3 = fxl_density(dens, g, u0, u1, u2, 0, 1, 0);
        const float q4 = fxl_density(dens, g, u0, u1, u2, 0, 0, -1), q5 = fxl_density(dens, g, u0, u1, u2, 0, 0, 1);
        const float g0 = -q0 + q1, g1 = -q2 + q3, g2 = -q4 + q5;
        const bool any = 0.0f < fxl_abs(g0) || 0.0f < fxl_abs(g1) || 0.0f < fxl_abs(g2);
        const float r0 = any ? -g0 : o0, r1 = any ? -g1 : o1, r2 = any ? -g2 : o2;  // uniform density: use the position
        const float w0 = fxl_dp3(r0, r1, r2, P.world + 0), w1 = fxl_dp3(r0, r1, r2, P.world + 4);
        const float w2 = fxl_dp3(r0, r1, r2, P.world + 8);
        const float wv[3] = {w0, w1, w2};
        const float winv = fxl_rsq(fxl_dp3(w0, w1, w2, wv));
        const float n0 = winv * w0, n1 = winv * w1, n2 = winv * w2;
        const float yy = n1 * n1, zz = n2 * n2;
        const float a = fxl_fma(n0, n0, -yy) * 0.4290427565574646f;    // c1 (x^2 - y^2)
        const float b = fxl_fma(zz, 3.0f, -1.0f) * 0.24770796298980713f;  // c5 (3 z^2 - 1)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float t10 = P.sh[6][c] * b;
            t10 = fxl_fma(a, P.sh[8][c], t10);
            float t3 = fxl_fma(P.sh[0][c], 0.8862269520759583f, t10);   // c4 L00
            float t8 = P.sh[4][c] * -n0;
            t10 = P.sh[7][c] * -n0;
            t10 = n2 * t10;
            t8 = fxl_fma(t8, -n1, t10);
            const float t9 = P.sh[5][c] * -n1;
            t8 = fxl_fma(t9, n2, t8);
            t3 = fxl_fma(t8, 0.8580855131149292f, t3);                   // 2 c1 (xy, xz, yz terms)
            float t4 = P.sh[1][c] * -n1;
            t4 = fxl_fma(P.sh[3][c], -n0, t4);
            t4 = fxl_fma(P.sh[2][c], n2, t4);
            t3 = fxl_fma(t4, 1.0233267545700073f, t3);                   // 2 c2 (linear terms)
            irr[c] = fxl_max(t3, 0.0f);
        }
        const float rv[3] = {r0, r1, r2};
        const float rinv = fxl_rsq(fxl_dp3(r0, r1, r2, rv));
        ao = fxl_cast_ray(dens, g, o0, o1, o2, rinv * r0, rinv * r1, rinv * r2, step, P.num_samples);
    }
}

// light colour x shadow + (occlusion x irradiance | ambient), one channel
FXL_FN float fxl_combine(const LightConsts& P, const int c, const float shadow, const float ao, const float irr) {
    const float lc = P.light_color[3] * P.light_color[c];
    const float amb = P.has_light_probes ? ao * irr : P.ambient[3] * P.ambient[c];
    return fxl_fma(shadow, lc, amb);
}

// The normalised light direction in volume space (mul(g_lightPt, (float3x3)g_worldI), normalize)
FXL_FN void fxl_light_dir(const LightConsts& P, float& l0, float& l1, float& l2) {
    const float a0 = fxl_dp3(P.light_pt[0], P.light_pt[1], P.light_pt[2], P.world_i + 0);
    const float a1 = fxl_dp3(P.light_pt[0], P.light_pt[1], P.light_pt[2], P.world_i + 4);
    const float a2 = fxl_dp3(P.light_pt[0], P.light_pt[1], P.light_pt[2], P.world_i + 8);
    const float lv[3] = {a0, a1, a2};
    const float inv = fxl_rsq(fxl_dp3(a0, a1, a2, lv));
    l0 = inv * a0; l1 = inv * a1; l2 = inv * a2;
}

// The light-map word of voxel (x, y, z).
FXL_FN unsigned light_map_voxel(const unsigned short* __restrict__ dens, const LightGeom& g, const LightConsts& P,
                                const int x, const int y, const int z) {
    const float o0 = fxl_fma(((float)x + 0.5f) / (float)g.nx, 2.0f, -1.0f);
    const float o1 = fxl_fma(((float)y + 0.5f) / (float)g.ny, 2.0f, -1.0f);
    const float o2 = fxl_fma(((float)z + 0.5f) / (float)g.nz, 2.0f, -1.0f);
    const float u0 = fxl_fma(o0, 0.5f, 0.5f), u1 = fxl_fma(o1, 0.5f, 0.5f), u2 = fxl_fma(o2, 0.5f, 0.5f);
    float shadow = 1.0f, ao = 1.0f, irr[3] = {0.0f, 0.0f, 0.0f};
    if (fxl_density(dens, g, u0, u1, u2, 0, 0, 0) >= 0.01f) {
        const float step = 3.464101552963257f / (float)P.num_samples;  // g_maxDist = 2 sqrt(3) (RayMarch.hlsli:29-30)
        float l0, l1, l2;
        fxl_light_dir(P, l0, l1, l2);
        fxl_light_at(dens, g, P, o0, o1, o2, u0, u1, u2, l0, l1, l2, step, shadow, ao, irr);
    }
    return fxl_pack_r11g11b10(fxl_combine(P, 0, shadow, ao, irr[0]), fxl_combine(P, 1, shadow, ao, irr[1]),
                              fxl_combine(P, 2, shadow, ao, irr[2]));
}

}  // namespace fxb
