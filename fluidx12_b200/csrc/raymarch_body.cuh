// raymarch_body.cuh — one cube-map texel of the view-ray march with the separate light pass (SURVEY.md §8 f3;
// raymarch.cu).
//
// Reference: FluidX12/Content/Shaders/CSRayMarchV.hlsl = CSRayMarch.hlsl:98-196 compiled with _LIGHT_PASS_
// (GetLocalPos :40-66; ComputeRayOrigin RayMarch.hlsli:146-177; ComputeTargetHit :182-187; GetStep :215-228; GetLight
// = light-map fetch :273-278), dispatched by Fluid::rayMarchV (Fluid.cpp:880-908) into mip m_cubeMapLOD of the
// R8G8B8A8_UNORM cube map (Fluid.cpp:229-232).  The arithmetic follows the shipped Bin/CSRayMarchV.cso instruction by
// instruction; that blob is built with _CPU_CUBE_FACE_CULL_ == 1, i.e. a face is marched iff its bit is set in the
// visibility mask (Fluid.cpp:51-63).  Restated platform semantics as in lightmap_body.cuh, plus: the light map is
// sampled LINEAR_CLAMP on its decoded fp32 values; the UNORM8 store is NaN -> 0, clamp to [0, 1], * 255, + 0.5,
// truncate.  A texel whose face is culled or whose ray misses the volume is not written (stale contents stay, as in
// the reference).  Same portability layer as lightmap_body.cuh (tests/emu/raymarch_emu.cpp runs it on the CPU).
#pragma once

#include "lightmap_body.cuh"

#if defined(__CUDACC__)
namespace fxb {
typedef uint2 RU2;
FXL_FN RU2 fxr_ld8(const RU2* p) { return __ldg(p); }
FXL_FN unsigned fxr_ld4(const unsigned* p) { return __ldg(p); }
FXL_FN float fxr_half_bits(unsigned h) { return __half2float(__ushort_as_half((unsigned short)h)); }
FXL_FN float fxr_from_bits(unsigned u) { return __uint_as_float(u); }
FXL_FN bool fxr_isnan(float v) { return v != v; }
}  // namespace fxb
#else
namespace fxb {
struct alignas(8) RU2 { unsigned x, y; };
FXL_FN RU2 fxr_ld8(const RU2* p) { return *p; }
FXL_FN unsigned fxr_ld4(const unsigned* p) { return *p; }
FXL_FN float fxr_half_bits(unsigned h) {
    const unsigned short s = (unsigned short)h;
    return fxl_h2f(&s);
}
FXL_FN float fxr_from_bits(unsigned u) {
    float f;
    __builtin_memcpy(&f, &u, 4);
    return f;
}
FXL_FN bool fxr_isnan(float v) { return v != v; }
}  // namespace fxb
#endif

namespace fxb {

// = fxb_view_params (include/fluidx_b200.h)
struct ViewConsts {
    float eye_pt[3];
    float world_i[12];
    unsigned num_samples;
    unsigned visibility_mask;
    unsigned cube_size;
};

FXL_FN float fxr_unpack(unsigned f, int mb) {  // one R11G11B10_FLOAT component (mb mantissa bits) -> fp32, exact
    const unsigned e = f >> mb, m = f & ((1u << mb) - 1u);
    if (e == 0u) return (float)m * (mb == 6 ? 9.5367431640625e-07f : 1.9073486328125e-06f);  // m * 2^(-14 - mb)
    if (e == 31u) return fxr_from_bits(0x7F800000u | (m << (23 - mb)));
    return fxr_from_bits(((e + 112u) << 23) | (m << (23 - mb)));
}

FXL_FN float fxr_lerp3(const float fx, const float fy, const float fz, const float (&a)[8]) {  // tap k = x + 2 y + 4 z
    const float x00 = fxl_fma(fx, a[1] - a[0], a[0]), x10 = fxl_fma(fx, a[3] - a[2], a[2]);
    const float x01 = fxl_fma(fx, a[5] - a[4], a[4]), x11 = fxl_fma(fx, a[7] - a[6], a[6]);
    const float y0v = fxl_fma(fy, x10 - x00, x00), y1v = fxl_fma(fy, x11 - x01, x01);
    return fxl_fma(fz, y1v - y0v, y0v);
}

// Marches the ray of texel (x, y) of `face`; returns false when nothing is to be written.
// SEPARATE = true: CSRayMarchV — the light at a sample is fetched from the light map `lmap` (Fluid::rayMarchV after
// Fluid::rayMarchL).  SEPARATE = false: CSRayMarch (Fluid::rayMarch, Fluid.cpp:825-855; Bin/CSRayMarch.cso) — the
// light is computed on the spot by fxl_light_at over the compact density array `dens` with the light constants `LP`
// (LP->num_samples = g_numLightSamples).
template <bool SEPARATE>
FXL_FN bool ray_march_texel(const RU2* __restrict__ col, const unsigned* __restrict__ lmap,
                            const unsigned short* __restrict__ dens, const LightGeom& g, const ViewConsts& P,
                            const LightConsts* LP, const int x, const int y, const int face, unsigned* rgba8) {
    if (((1u << face) & P.visibility_mask) == 0u) return false;
    float ro[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float* w = P.world_i + 4 * k;
        ro[k] = ((P.eye_pt[0] * w[0] + P.eye_pt[1] * w[1]) + P.eye_pt[2] * w[2]) + 1.0f * w[3];
    }
    const float S = (float)P.cube_size;
    const float X = fxl_fma(((float)x + 0.5f) / S, 2.0f, -1.0f);
    const float Z = fxl_fma(((float)y + 0.5f) / S, 2.0f, -1.0f);  // = -pos.y of GetLocalPos
    float tg[3];
    switch (face) {
        case 0: tg[0] = 1.0f; tg[1] = -Z; tg[2] = -X; break;
        case 1: tg[0] = -1.0f; tg[1] = -Z; tg[2] = X; break;
        case 2: tg[0] = X; tg[1] = 1.0f; tg[2] = Z; break;
        case 3: tg[0] = X; tg[1] = -1.0f; tg[2] = -Z; break;
        case 4: tg[0] = X; tg[1] = -Z; tg[2] = 1.0f; break;
        default: tg[0] = -X; tg[1] = -Z; tg[2] = -1.0f; break;
    }
    float dir[3];
    {
        const float d0 = -ro[0] + tg[0], d1 = -ro[1] + tg[1], d2 = -ro[2] + tg[2];
        const float dv[3] = {d0, d1, d2};
        const float inv = fxl_rsq(fxl_dp3(d0, d1, d2, dv));
        dir[0] = inv * d0; dir[1] = inv * d1; dir[2] = inv * d2;
    }
    if (!(1.0f >= fxl_abs(ro[0]) && 1.0f >= fxl_abs(ro[1]) && 1.0f >= fxl_abs(ro[2]))) {
        // ComputeRayOrigin: nearest entry point on the box (the eye is outside)
        const float FMAX = 3.402823466e+38f;
        float r3[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int sg = (0.0f < dir[k] ? -1 : 0) - (dir[k] < 0.0f ? -1 : 0);  // -sign(dir)
            r3[k] = -ro[k] + (float)sg;
        }
        const float u0 = r3[0] / dir[0], u1 = r3[1] / dir[1];
        float U = FMAX;
        bool hit = false;
        if (u0 >= 0.0f && 1.0f >= fxl_abs(fxl_fma(dir[1], u0, ro[1])) && 1.0f >= fxl_abs(fxl_fma(dir[2], u0, ro[2]))) {
            hit = u0 < FMAX;
            U = fxl_min(u0, FMAX);
        }
        if (u1 >= 0.0f && 1.0f >= fxl_abs(fxl_fma(dir[2], u1, ro[2])) && 1.0f >= fxl_abs(fxl_fma(dir[0], u1, ro[0])) &&
            u1 < U) {
            U = u1;
            hit = true;
        }
        const float u2 = r3[2] / dir[2];
        if (u2 >= 0.0f && 1.0f >= fxl_abs(fxl_fma(dir[0], u2, ro[0])) && 1.0f >= fxl_abs(fxl_fma(dir[1], u2, ro[1])) &&
            u2 < U) {
            U = u2;
            hit = true;
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) ro[k] = fxl_min(fxl_max(fxl_fma(dir[k], U, ro[k]), -1.0f), 1.0f);
        if (!hit) return false;
    }
    const float step = 3.464101552963257f / (float)P.num_samples;
    const float tmax = fxl_max((tg[2] + -ro[2]) / dir[2], fxl_max((tg[1] + -ro[1]) / dir[1], (tg[0] + -ro[0]) / dir[0]));
    float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, lstep = 0.0f;
    if (!SEPARATE) {
        fxl_light_dir(*LP, l0, l1, l2);
        lstep = 3.464101552963257f / (float)LP->num_samples;  // g_lightStep (RayMarch.hlsli:31)
    }
    float sc[4] = {0.0f, 0.0f, 0.0f, 0.0f}, t = 0.0f, prev = 0.0f;
    for (unsigned i = 0; i < P.num_samples; ++i) {
        const float p0 = fxl_fma(dir[0], t, ro[0]), p1 = fxl_fma(dir[1], t, ro[1]), p2 = fxl_fma(dir[2], t, ro[2]);
        if (1.0f < fxl_abs(p0) || 1.0f < fxl_abs(p1) || 1.0f < fxl_abs(p2)) break;
        const float uu0 = fxl_fma(p0, 0.5f, 0.5f), uu1 = fxl_fma(p1, 0.5f, 0.5f), uu2 = fxl_fma(p2, 0.5f, 0.5f);
        const float tx = fxl_fma(uu0, (float)g.nx, -0.5f);
        const float ty = fxl_fma(uu1, (float)g.ny, -0.5f);
        const float tz = fxl_fma(uu2, (float)g.nz, -0.5f);
        const int ix = fxl_tap(tx), iy = fxl_tap(ty), iz = fxl_tap(tz);
        const float fx = tx - fxl_floor(tx), fy = ty - fxl_floor(ty), fz = tz - fxl_floor(tz);
        const int x0 = fxl_clamp(ix, g.nx), x1 = fxl_clamp(ix + 1, g.nx);
        const int y0 = fxl_clamp(iy, g.ny), y1 = fxl_clamp(iy + 1, g.ny);
        const int z0 = fxl_clamp(iz, g.nz), z1 = fxl_clamp(iz + 1, g.nz);
        const size_t r00 = ((size_t)z0 * g.ny + y0) * g.nx, r10 = ((size_t)z0 * g.ny + y1) * g.nx;
        const size_t r01 = ((size_t)z1 * g.ny + y0) * g.nx, r11 = ((size_t)z1 * g.ny + y1) * g.nx;
        const size_t idx[8] = {r00 + x0, r00 + x1, r10 + x0, r10 + x1, r01 + x0, r01 + x1, r11 + x0, r11 + x1};
        RU2 tex[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) tex[k] = fxr_ld8(col + idx[k]);
        float c[4], a[8];
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const unsigned w = ch < 2 ? tex[k].x : tex[k].y;
                a[k] = fxr_half_bits((ch & 1) ? (w >> 16) : (w & 0xFFFFu));
            }
            c[ch] = fxr_lerp3(fx, fy, fz, a);
        }
        float r5[4], new_step;
        if (0.01f < c[3]) {
            float L[3];
            if (SEPARATE) {
                unsigned lw[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) lw[k] = fxr_ld4(lmap + idx[k]);
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
#pragma unroll
                    for (int k = 0; k < 8; ++k)
                        a[k] = ch == 0 ? fxr_unpack(lw[k] & 0x7FFu, 6) : (ch == 1 ? fxr_unpack((lw[k] >> 11) & 0x7FFu, 6)
                                                                                  : fxr_unpack(lw[k] >> 22, 5));
                    L[ch] = fxr_lerp3(fx, fy, fz, a);
                }
            } else {
                float shadow, ao, irr[3];
                fxl_light_at(dens, g, *LP, p0, p1, p2, uu0, uu1, uu2, l0, l1, l2, lstep, shadow, ao, irr);
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) L[ch] = fxl_combine(*LP, ch, shadow, ao, irr[ch]);
            }
            const float transm = -sc[3] + 1.0f;
            const float ev = fxl_min(0.00390625f / fxl_abs(-prev + c[3]), 2.0f);
            const float ui = fxl_min(-c[3] + 1.0f, 1.0f);
            const float th = -transm + 1.0f;
            new_step = fxl_max(th * (ui * (ev * 1.5f)), 1.0f) * step;
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) r5[ch] = fxl_fma(transm * (L[ch] * c[ch]), 0.8f, sc[ch]);
            r5[3] = fxl_fma(0.8f * c[3], transm, sc[3]);
            if (transm < 0.01f) {
#pragma unroll
                for (int ch = 0; ch < 4; ++ch) sc[ch] = r5[ch];
                break;
            }
            prev = c[3];
        } else {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) r5[ch] = sc[ch];
            new_step = step;
        }
        t = t + new_step;
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) sc[ch] = r5[ch];
        if (tmax < t) break;
    }
    unsigned out = 0;
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) {
        float v = ch < 3 ? sc[ch] * 0.15915493667125702f : sc[ch];  // scatter.xyz /= 2 pi
        v = fxr_isnan(v) ? 0.0f : fxl_min(fxl_max(v, 0.0f), 1.0f);
        out |= (unsigned)(v * 255.0f + 0.5f) << (8 * ch);
    }
    *rgba8 = out;
    return true;
}

}  // namespace fxb
