// fluid_oracle.cpp — CPU restatement (OpenMP C++) of FluidX12's per-frame smoke-solver step.
//
// TEST INFRASTRUCTURE ONLY.  Nothing under fluidx12_b200/ links, loads or calls this file; only
// tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it,
// and only as the checker or as the reported CPU baseline.
//
// PARITY: PINNED TO THE SHIPPED BYTECODE, UNPINNED FOR THE PLATFORM.  The reference (Windows + D3D12 + closed
// XUSG.dll) ships no tests, golden vectors or CPU path and cannot be built or run here.  What it does ship is the
// fxc-compiled DXBC of its shaders; tests/golden/dxbc_interp.py executes those blobs (operation order, swizzles,
// folded constants and control flow come from the bytecode) and tests/test_dxbc_golden.py requires this file to
// reproduce the resulting vectors bit for bit (3D, 2D, MIRROR, CLAMP, paused frame, zero start).  What no file of the
// reference defines — D3D's sampler and format conversions, exp2, and the thread interleaving of the racy
// relaxation loop — is restated (decisions below) in the interpreter exactly as it is here, so those remain
// unpinned.  This file states the deterministic semantics of the path from the HLSL source and the DXBC
// (SURVEY.md Appendix A, B, D):
//
//   CSAdvect      FluidX12/Content/Shaders/CSAdvect.hlsl:41-79   (+ Simulation.hlsli:8-19,
//                 Impulse.hlsli:8-18; sampler LINEAR_MIRROR Content/Fluid.cpp:452,
//                 LINEAR_CLAMP Content/FluidEZ.cpp:406)
//   CSProject3D   FluidX12/Content/Shaders/CSProject3D.hlsl:68-113
//   CSProject2D   FluidX12/Content/Shaders/CSProject2D.hlsl:64-106
//   Poisson       FluidX12/Content/Shaders/CSPoisson.hlsli:8-26
//   schedule      FluidX12/Content/Fluid.cpp:283-291,344-345 (UpdateFrame), :348-410 (Simulate)
//   dt rule       FluidX12/FluidX12.cpp:266-267
//   formats       FluidX12/Content/Fluid.cpp:204-221 (RGBA16F velocity/colour, R32F pressure)
// and, for the rows past the step (each pinned the same way, by executing the shipped blob —
// tests/test_lightmap.py, tests/test_raymarch.py):
//   CSRayMarchL   FluidX12/Content/Shaders/CSRayMarchL.hlsl:15-80 (+ RayMarch.hlsli)   light_map()
//   CSRayMarchV   FluidX12/Content/Shaders/CSRayMarch.hlsl:98-196 with _LIGHT_PASS_     ray_march_v(lmap)
//   CSRayMarch    the same file without it                                              ray_march_v(nullptr, LP)
//
// Decisions where the platform is implementation-defined:
//   D1 pressure solve = synchronous double-buffered Jacobi, <=64 sweeps, per-cell freeze after the
//      sweep in which |x - x0| < 0.001, freeze flags reset every frame;
//   D4 trilinear = fp32 weights, t = fma(coord, W, -0.5), lerp x then y then z as fma(f, b-a, a);
//   D5 DXBC `mad` is fused, `dp3` is (x*x + y*y) + z*z unfused; no re-association;
//   D6 fp32->fp16 stores round to nearest even;  D7 exp2 = libm exp2f on the host (the CUDA build
//      receives the same values as a host-computed table);  D8 velocity .w is a don't-care.
//
// Build: g++ -O2 -fopenmp -ffp-contract=off -fno-fast-math -march=x86-64-v3 (see oracle/Makefile).

#ifdef _OPENMP
#include <omp.h>
#endif
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#if defined(__F16C__)
#include <immintrin.h>
#endif

namespace {

// ---------------------------------------------------------------------------------------------
// fp16 storage conversion (App. B.3): loads are exact, stores are RNE.
// ---------------------------------------------------------------------------------------------
inline float half_to_float(uint16_t h) {
#if defined(__F16C__)
    return _cvtsh_ss(h);
#else
    const uint32_t sign = uint32_t(h & 0x8000u) << 16;
    uint32_t exp = (h >> 10) & 0x1fu, man = h & 0x3ffu, bits;
    if (exp == 0) {
        if (man == 0) bits = sign;
        else {  // subnormal half -> normal float
            int e = -1;
            do { ++e; man <<= 1; } while (!(man & 0x400u));
            bits = sign | uint32_t(127 - 15 - e) << 23 | (man & 0x3ffu) << 13;
        }
    } else if (exp == 31) bits = sign | 0x7f800000u | man << 13;
    else bits = sign | (exp + 112u) << 23 | man << 13;
    float f; std::memcpy(&f, &bits, 4); return f;
#endif
}

inline uint16_t float_to_half(float f) {
#if defined(__F16C__)
    return (uint16_t)_cvtss_sh(f, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
#else
    uint32_t x; std::memcpy(&x, &f, 4);
    const uint16_t sign = uint16_t((x >> 16) & 0x8000u);
    x &= 0x7fffffffu;
    if (x >= 0x7f800000u) return sign | 0x7c00u | (x > 0x7f800000u ? 0x200u | ((x >> 13) & 0x3ffu) : 0);
    if (x >= 0x477ff000u) return sign | 0x7c00u;             // rounds to inf
    if (x < 0x33000001u) return sign;                        // rounds to zero
    int exp = int(x >> 23) - 127;
    uint32_t man = (x & 0x7fffffu) | 0x800000u;
    int shift = exp < -14 ? 13 + (-14 - exp) : 13;
    uint32_t half_man = man >> shift, rem = man & ((1u << shift) - 1), halfway = 1u << (shift - 1);
    uint32_t out = exp < -14 ? half_man : (uint32_t(exp + 15) << 10) + (half_man - 0x400u);
    if (rem > halfway || (rem == halfway && (out & 1u))) ++out;
    return sign | uint16_t(out);
#endif
}

// ---------------------------------------------------------------------------------------------
// Sampler addressing (App. B.2).  MIRROR has period 2W: -1 -> 0, W -> W-1, -2 -> 1, ...
// ---------------------------------------------------------------------------------------------
enum { ADDRESS_MIRROR = 0, ADDRESS_CLAMP = 1 };

inline int address_tap(int i, int W, int mode) {
    if (mode == ADDRESS_CLAMP) return std::min(std::max(i, 0), W - 1);
    const int period = 2 * W;
    int m = i % period;
    if (m < 0) m += period;
    return m < W ? m : period - 1 - m;
}

// floor(t) as an int, saturated to +-2^30 so that tap+1 never overflows; NaN -> -2^30.
inline int floor_to_tap(float t) {
    const float lim = 1073741824.0f;
    if (!(t > -lim)) return -(1 << 30);
    if (t > lim) return 1 << 30;
    return (int)std::floor(t);
}

// nx, ny, nz: extent of the array the function works on.  For the z-slab tests the array is a window of a taller
// grid: nzg = global plane count (0 = same as nz) and z0 = global index of local plane 0; positions, sampler
// addressing, the emitter and the clamp-to-edge neighbour rule are then evaluated in GLOBAL coordinates.
struct Grid {
    int nx, ny, nz;
    int nzg = 0, z0 = 0;
    int gnz() const { return nzg > 0 ? nzg : nz; }
    size_t idx(int x, int y, int z) const { return (size_t(z) * ny + y) * nx + x; }
    size_t voxels() const { return size_t(nx) * ny * nz; }
    int local_z(int zg) const { return std::min(std::max(zg - z0, 0), nz - 1); }  // window-clamped
};

// One trilinear fetch of an RGBA16F field at normalised coordinate (cx,cy,cz) -> out[4].
// CSAdvect.hlsl:53-54 `SampleLevel(g_smpLinear, adv, 0)`.
inline void sample_trilinear(const uint16_t* field, const Grid& g, int mode,
                             float cx, float cy, float cz, float out[4]) {
    const float tx = std::fmaf(cx, (float)g.nx, -0.5f);
    const float ty = std::fmaf(cy, (float)g.ny, -0.5f);
    const float tz = std::fmaf(cz, (float)g.gnz(), -0.5f);
    const int ix = floor_to_tap(tx), iy = floor_to_tap(ty), iz = floor_to_tap(tz);
    const float fx = tx - std::floor(tx), fy = ty - std::floor(ty), fz = tz - std::floor(tz);
    const int x0 = address_tap(ix, g.nx, mode), x1 = address_tap(ix + 1, g.nx, mode);
    const int y0 = address_tap(iy, g.ny, mode), y1 = address_tap(iy + 1, g.ny, mode);
    const int z0 = g.local_z(address_tap(iz, g.gnz(), mode)), z1 = g.local_z(address_tap(iz + 1, g.gnz(), mode));
    const uint16_t* t000 = field + 4 * g.idx(x0, y0, z0);
    const uint16_t* t100 = field + 4 * g.idx(x1, y0, z0);
    const uint16_t* t010 = field + 4 * g.idx(x0, y1, z0);
    const uint16_t* t110 = field + 4 * g.idx(x1, y1, z0);
    const uint16_t* t001 = field + 4 * g.idx(x0, y0, z1);
    const uint16_t* t101 = field + 4 * g.idx(x1, y0, z1);
    const uint16_t* t011 = field + 4 * g.idx(x0, y1, z1);
    const uint16_t* t111 = field + 4 * g.idx(x1, y1, z1);
    for (int c = 0; c < 4; ++c) {
        const float a000 = half_to_float(t000[c]), a100 = half_to_float(t100[c]);
        const float a010 = half_to_float(t010[c]), a110 = half_to_float(t110[c]);
        const float a001 = half_to_float(t001[c]), a101 = half_to_float(t101[c]);
        const float a011 = half_to_float(t011[c]), a111 = half_to_float(t111[c]);
        const float x00 = std::fmaf(fx, a100 - a000, a000);
        const float x10 = std::fmaf(fx, a110 - a010, a010);
        const float x01 = std::fmaf(fx, a101 - a001, a001);
        const float x11 = std::fmaf(fx, a111 - a011, a011);
        const float y0v = std::fmaf(fy, x10 - x00, x00);
        const float y1v = std::fmaf(fy, x11 - x01, x01);
        out[c] = std::fmaf(fz, y1v - y0v, y0v);
    }
}

// Gaussian emitter basis of a voxel (CSAdvect.hlsl:33-36, :57-59; DXBC order in App. A.1).
inline float emitter_basis(const Grid& g, int x, int y, int z, float disp[3]) {
    const float px = ((float)x + 0.5f) / (float)g.nx;
    const float py = ((float)y + 0.5f) / (float)g.ny;
    const float pz = ((float)(z + g.z0) + 0.5f) / (float)g.gnz();
    disp[0] = px + -0.5f;
    disp[1] = py + -0.100000001f;
    disp[2] = pz + -0.5f;
    const float d2 = (disp[0] * disp[0] + disp[1] * disp[1]) + disp[2] * disp[2];
    const float r2 = (1.0f < (float)g.gnz()) ? 0.00390625f : 0.0009765625f;
    const float e = ((d2 * -4.0f) / r2) * 1.44269502f;
    return std::exp2f(e);
}

// ---------------------------------------------------------------------------------------------
// CSAdvect (App. A.1)
// ---------------------------------------------------------------------------------------------
void advect(const Grid& g, int mode, float dt, const uint16_t* vel_in, const uint16_t* col_in,
            uint16_t* vel_out, uint16_t* col_out) {
    const bool is3d = 1.0f < (float)g.gnz();
    const float atten = std::max(std::fmaf(-dt, 0.200000003f, 1.0f), 0.0f);
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < g.nz; ++z)
        for (int y = 0; y < g.ny; ++y)
            for (int x = 0; x < g.nx; ++x) {
                const size_t i = g.idx(x, y, z);
                const float px = ((float)x + 0.5f) / (float)g.nx;
                const float py = ((float)y + 0.5f) / (float)g.ny;
                const float pz = ((float)(z + g.z0) + 0.5f) / (float)g.gnz();
                float disp[3];
                const float basis = emitter_basis(g, x, y, z, disp);
                float F[3];
                if (is3d) {
                    F[0] = std::fmaf(basis, 0.0f, disp[2] * -200.0f);
                    F[1] = std::fmaf(basis, 192.0f, 0.0f);
                    F[2] = std::fmaf(basis, 0.0f, disp[0] * 200.0f);
                } else {
                    F[0] = 0.0f; F[1] = basis * 48.0f; F[2] = 0.0f;
                }
                const float u0x = half_to_float(vel_in[4 * i + 0]);
                const float u0y = half_to_float(vel_in[4 * i + 1]);
                const float u0z = half_to_float(vel_in[4 * i + 2]);
                const float ax = std::fmaf(-u0x, dt, px);
                const float ay = std::fmaf(-u0y, dt, py);
                const float az = std::fmaf(-u0z, dt, pz);
                float u[4], c[4];
                sample_trilinear(vel_in, g, mode, ax, ay, az, u);
                sample_trilinear(col_in, g, mode, ax, ay, az, c);
                if (basis >= 0.0183156393f) {
                    for (int k = 0; k < 3; ++k) u[k] = std::fmaf(F[k], dt, u[k]);
                    const float bdt = basis * dt;
                    const float imp[4] = {8.0f, 16.0f, 40.0f, 40.0f};
                    for (int k = 0; k < 4; ++k)
                        c[k] = std::min(std::max(std::fmaf(bdt, imp[k], c[k]), 0.0f), 1.0f);
                }
                for (int k = 0; k < 3; ++k) vel_out[4 * i + k] = float_to_half(u[k] * atten);
                vel_out[4 * i + 3] = 0;  // don't-care lane (D8)
                for (int k = 0; k < 4; ++k) col_out[4 * i + k] = float_to_half(c[k] * atten);
            }
}

// ---------------------------------------------------------------------------------------------
// CSProject2D / CSProject3D (App. A.2, A.3)
// ---------------------------------------------------------------------------------------------
struct Nbr { size_t L, R, U, D, F, B; };

inline Nbr neighbours(const Grid& g, int x, int y, int z, bool is3d) {
    Nbr n;
    n.L = g.idx(std::max(x, 1) - 1, y, z);
    n.R = g.idx(std::min(x + 1, g.nx - 1), y, z);
    n.U = g.idx(x, std::max(y, 1) - 1, z);
    n.D = g.idx(x, std::min(y + 1, g.ny - 1), z);
    const int zg = z + g.z0;
    n.F = is3d ? g.idx(x, y, g.local_z(std::max(zg, 1) - 1)) : 0;
    n.B = is3d ? g.idx(x, y, g.local_z(std::min(zg + 1, g.gnz() - 1))) : 0;
    return n;
}

// s = 2*divergence exactly as the DXBC sums it (the 0.5 is applied inside the Poisson loop).
void divergence2x(const Grid& g, const uint16_t* vel, float* s) {
    const bool is3d = g.gnz() > 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < g.nz; ++z)
        for (int y = 0; y < g.ny; ++y)
            for (int x = 0; x < g.nx; ++x) {
                const Nbr n = neighbours(g, x, y, z, is3d);
                const float a = -half_to_float(vel[4 * n.L + 0]) + half_to_float(vel[4 * n.R + 0]);
                float b = -half_to_float(vel[4 * n.U + 1]) + half_to_float(vel[4 * n.D + 1]);
                if (is3d) {
                    b = b + a;
                    const float c = -half_to_float(vel[4 * n.F + 2]) + half_to_float(vel[4 * n.B + 2]);
                    s[g.idx(x, y, z)] = c + b;
                } else {
                    s[g.idx(x, y, z)] = a + b;
                }
            }
}

// Synchronous Jacobi with per-cell freeze.  p is the persistent pressure (in/out), q a scratch
// buffer of the same size.  Returns the number of sweeps in which at least one cell was active
// (S_exec); hist[k] = number of active cells entering sweep k.
int jacobi(const Grid& g, const float* s, float* p, float* q, uint8_t* active, int iters,
           int early_exit, int64_t* hist) {
    const bool is3d = g.gnz() > 1;
    const float inv = is3d ? 0.166666672f : 0.25f;
    const size_t n = g.voxels();
    std::memset(active, 1, n);
    if (hist) std::fill(hist, hist + iters, int64_t(0));
    float* cur = p;
    float* nxt = q;
    int sweeps = 0;
    for (int k = 0; k < iters; ++k) {
        int64_t n_active = 0;
#pragma omp parallel for collapse(2) schedule(static) reduction(+ : n_active)
        for (int z = 0; z < g.nz; ++z)
            for (int y = 0; y < g.ny; ++y)
                for (int x = 0; x < g.nx; ++x) {
                    const size_t i = g.idx(x, y, z);
                    if (!active[i]) { nxt[i] = cur[i]; continue; }
                    ++n_active;
                    const Nbr nb = neighbours(g, x, y, z, is3d);
                    float acc = std::fmaf(-s[i], 0.5f, cur[nb.L]);
                    acc = cur[nb.R] + acc;
                    acc = cur[nb.U] + acc;
                    acc = cur[nb.D] + acc;
                    if (is3d) {
                        acc = cur[nb.F] + acc;
                        acc = cur[nb.B] + acc;
                    }
                    nxt[i] = acc * inv;
                    if (early_exit && std::fabs(std::fmaf(acc, inv, -cur[i])) < 0.00100000005f)
                        active[i] = 0;
                }
        if (hist) hist[k] = n_active;
        if (n_active == 0) break;  // the sweep was an identity copy: cur already holds the result
        ++sweeps;
        std::swap(cur, nxt);
    }
    if (cur != p) std::memcpy(p, cur, n * sizeof(float));
    return sweeps;
}

// Gradient subtract + soft-wall damping + fp16 store (CSProject3D.hlsl:55-63,:106-112).
void gradient(const Grid& g, const uint16_t* vel_in, const float* p, uint16_t* vel_out) {
    const bool is3d = g.gnz() > 1;
#pragma omp parallel for collapse(2) schedule(static)
    for (int z = 0; z < g.nz; ++z)
        for (int y = 0; y < g.ny; ++y)
            for (int x = 0; x < g.nx; ++x) {
                const size_t i = g.idx(x, y, z);
                const Nbr n = neighbours(g, x, y, z, is3d);
                float u[3] = {half_to_float(vel_in[4 * i + 0]), half_to_float(vel_in[4 * i + 1]),
                              half_to_float(vel_in[4 * i + 2])};
                const float gx = -p[n.L] + p[n.R];
                const float gy = -p[n.U] + p[n.D];
                float bp[3];
                const float px = ((float)x + 0.5f) / (float)g.nx;
                const float py = ((float)y + 0.5f) / (float)g.ny;
                const float pz = ((float)(z + g.z0) + 0.5f) / (float)g.gnz();
                if (is3d) {
                    const float gz = -p[n.F] + p[n.B];
                    u[0] = std::fmaf(-gx, 1.04166675f, u[0]);
                    u[1] = std::fmaf(-gy, 1.04166675f, u[1]);
                    u[2] = std::fmaf(-gz, 1.04166675f, u[2]);
                    bp[0] = std::fmaf(px, 2.0f, -1.0f);
                    bp[1] = std::fmaf(py, 2.0f, -1.0f);
                    bp[2] = std::fmaf(pz, 2.0f, -1.0f);
                } else {
                    u[0] = std::fmaf(-gx, 0.5f, u[0]);
                    u[1] = std::fmaf(-gy, 0.5f, u[1]);
                    bp[0] = std::fmaf(px, 2.0f, -1.0f);
                    bp[1] = std::fmaf(py, 2.0f, -1.0f);
                    bp[2] = std::fmaf(pz, 1.0f, 0.0f);
                }
                for (int k = 0; k < 3; ++k) {
                    float m = (-std::fabs(bp[k]) + 0.970000029f) * 33.3333359f;
                    m = std::min(std::max(m, -1.0f), 1.0f);
                    if (!(0.0f < u[k] * bp[k])) m = 1.0f;
                    u[k] = u[k] * m;
                }
                vel_out[4 * i + 0] = float_to_half(u[0]);
                vel_out[4 * i + 1] = float_to_half(u[1]);
                vel_out[4 * i + 2] = float_to_half(u[2]);
                vel_out[4 * i + 3] = 0;  // reference stores u.x here; never read (D8)
            }
}

// ---------------------------------------------------------------------------------------------
// Light-map pass: CSRayMarchL.hlsl:15-80 with RayMarch.hlsli:62-68 (GetSample), :75-98
// (GetDensityGradient), :203-210 (LocalToTex3DSpace), :215-228 (GetStep), :233-268 (CastLightRay) and the order-3 SH
// irradiance of XUSG's SHIrradiance.hlsli (not in the reference tree: restated from the shipped bytecode, whose five
// constants are Ramamoorthi & Hanrahan's c1..c5 — 0.429043, 2*0.511664 = 1.023328, 2*0.429043 = 0.858086, 0.886227,
// 0.247708).  Operation order, fused mads and folded constants follow Bin/CSRayMarchL.cso instruction by instruction
// (tests/golden/dxbc_interp.py executes that blob; tests/test_lightmap.py requires this function to reproduce it).
// Restated because no file of the reference defines them: the LINEAR_CLAMP sampler (Fluid.cpp:475) as in
// sample_trilinear above, with the instruction's texel offsets added to both taps before clamping; `rsq` as
// 1 / sqrt(x), both correctly rounded; min16float arithmetic carried out in fp32; the R11G11B10_FLOAT store
// (Fluid.cpp:225-227) truncating toward zero, negative -> 0, finite overflow -> largest finite value.
// ---------------------------------------------------------------------------------------------
struct LightParams {      // = fxb_light_params of include/fluidx_b200.h
    float light_pt[3];    // cbPerFrame g_lightPt, cb1[1].xyz (Fluid.cpp:304)
    float light_color[4]; // cb1[2] (Fluid.cpp:305)
    float ambient[4];     // cb1[3] (Fluid.cpp:306)
    float world_i[12];    // cbPerObject g_worldI, cb0[8..10] as XMStoreFloat3x4 wrote them (Fluid.cpp:318)
    float world[12];      // g_world, cb0[11..13] (Fluid.cpp:319)
    uint32_t num_samples; // cbSampleRes g_numSamples = m_maxLightSamples (Fluid.cpp:872)
    uint32_t has_light_probes;  // (Fluid.cpp:873)
    float sh[9][3];       // g_roSHCoeffs (Fluid.cpp:874)
};

inline uint32_t pack_r11g11b10(const float rgb[3]) {
    uint32_t out = 0;
    const int mbits[3] = {6, 6, 5}, shift[3] = {0, 11, 22};
    for (int k = 0; k < 3; ++k) {
        uint32_t w;
        std::memcpy(&w, &rgb[k], 4);
        const uint32_t sign = w >> 31, e = (w >> 23) & 0xFFu, m = w & 0x7FFFFFu;
        const int mb = mbits[k], drop = 23 - mb;
        const uint32_t maxfin = (30u << mb) | ((1u << mb) - 1u);
        uint32_t r;
        if (e == 0xFFu) r = m ? ((31u << mb) | ((1u << mb) - 1u)) : (sign ? 0u : (31u << mb));
        else if (sign) r = 0u;
        else if (e >= 143u) r = maxfin;
        else if (e >= 113u) r = ((e - 112u) << mb) | (m >> drop);
        else { const uint32_t sh = std::min(113u - e, 24u); r = ((m | 0x800000u) >> sh) >> drop; }
        out |= r << shift[k];
    }
    return out;
}

// colour.w at normalised (cx, cy, cz) + integer texel offset, LINEAR_CLAMP
inline float sample_density(const uint16_t* col, const Grid& g, float cx, float cy, float cz, int ox, int oy, int oz) {
    const float tx = std::fmaf(cx, (float)g.nx, -0.5f), ty = std::fmaf(cy, (float)g.ny, -0.5f);
    const float tz = std::fmaf(cz, (float)g.nz, -0.5f);
    const int ix = floor_to_tap(tx) + ox, iy = floor_to_tap(ty) + oy, iz = floor_to_tap(tz) + oz;
    const float fx = tx - std::floor(tx), fy = ty - std::floor(ty), fz = tz - std::floor(tz);
    const int x0 = address_tap(ix, g.nx, ADDRESS_CLAMP), x1 = address_tap(ix + 1, g.nx, ADDRESS_CLAMP);
    const int y0 = address_tap(iy, g.ny, ADDRESS_CLAMP), y1 = address_tap(iy + 1, g.ny, ADDRESS_CLAMP);
    const int z0 = address_tap(iz, g.nz, ADDRESS_CLAMP), z1 = address_tap(iz + 1, g.nz, ADDRESS_CLAMP);
    auto w = [&](int x, int y, int z) { return half_to_float(col[4 * g.idx(x, y, z) + 3]); };
    const float a000 = w(x0, y0, z0), a100 = w(x1, y0, z0), a010 = w(x0, y1, z0), a110 = w(x1, y1, z0);
    const float a001 = w(x0, y0, z1), a101 = w(x1, y0, z1), a011 = w(x0, y1, z1), a111 = w(x1, y1, z1);
    const float x00 = std::fmaf(fx, a100 - a000, a000), x10 = std::fmaf(fx, a110 - a010, a010);
    const float x01 = std::fmaf(fx, a101 - a001, a001), x11 = std::fmaf(fx, a111 - a011, a011);
    const float y0v = std::fmaf(fy, x10 - x00, x00), y1v = std::fmaf(fy, x11 - x01, x01);
    return std::fmaf(fz, y1v - y0v, y0v);
}

inline float dp3(const float a[3], const float b[3]) { return (a[0] * b[0] + a[1] * b[1]) + a[2] * b[2]; }
inline float rsq(float v) { return 1.0f / std::sqrt(v); }

// CastLightRay (RayMarch.hlsli:233-268) as compiled: returns the transmittance along `dir` from `o`.
inline float cast_light_ray(const uint16_t* col, const Grid& g, const float o[3], const float dir[3], float step,
                            uint32_t num_samples) {
    float transm = 1.0f, t = step, prev = 0.0f;
    for (uint32_t i = 0; i < num_samples; ++i) {
        float pos[3];
        for (int k = 0; k < 3; ++k) pos[k] = std::fmaf(dir[k], t, o[k]);
        if (1.0f < std::fabs(pos[0]) || 1.0f < std::fabs(pos[1]) || 1.0f < std::fabs(pos[2])) break;
        const float d = sample_density(col, g, std::fmaf(pos[0], 0.5f, 0.5f), std::fmaf(pos[1], 0.5f, 0.5f),
                                       std::fmaf(pos[2], 0.5f, 0.5f), 0, 0, 0);
        const float tr = std::fmaf(-d, 0.8f, 1.0f) * transm;
        if (tr < 0.01f) return tr;
        const float ev = std::fmin(0.00390625f / std::fabs(-prev + d), 2.0f);
        const float ui = std::fmin(-d + 1.0f, 1.0f);
        const float th = -transm + 1.0f;
        const float grow = std::fmax(th * (ui * (ev * 1.5f)), 1.0f);
        t = std::fmaf(step, grow, t);
        transm = tr;
        prev = d;
    }
    return transm;
}

void light_map(const Grid& g, const uint16_t* col, const LightParams& P, uint32_t* out) {
    const float fn[3] = {(float)g.nx, (float)g.ny, (float)g.nz};
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int z = 0; z < g.nz; ++z)
        for (int y = 0; y < g.ny; ++y)
            for (int x = 0; x < g.nx; ++x) {
                const int id[3] = {x, y, z};
                float o[3], uvw[3];
                for (int k = 0; k < 3; ++k) {
                    o[k] = std::fmaf(((float)id[k] + 0.5f) / fn[k], 2.0f, -1.0f);
                    uvw[k] = std::fmaf(o[k], 0.5f, 0.5f);
                }
                float shadow = 1.0f, ao = 1.0f, irr[3] = {0.0f, 0.0f, 0.0f};
                if (sample_density(col, g, uvw[0], uvw[1], uvw[2], 0, 0, 0) >= 0.01f) {
                    const float step = 3.464101552963257f / (float)P.num_samples;  // 0x405db3d7 = 2 sqrt(3)
                    float L[3], dir[3];
                    for (int k = 0; k < 3; ++k) L[k] = dp3(P.light_pt, P.world_i + 4 * k);
                    const float inv = rsq(dp3(L, L));
                    for (int k = 0; k < 3; ++k) dir[k] = inv * L[k];
                    shadow = cast_light_ray(col, g, o, dir, step, P.num_samples);
                    if (P.has_light_probes) {
                        const float q0 = sample_density(col, g, uvw[0], uvw[1], uvw[2], -1, 0, 0);
                        const float q1 = sample_density(col, g, uvw[0], uvw[1], uvw[2], 1, 0, 0);
                        const float q2 = sample_density(col, g, uvw[0], uvw[1], uvw[2], 0, -1, 0);
                        const float q3 = sample_density(col, g, uvw[0], uvw[1], uvw[2], 0, 1, 0);
                        const float q4 = sample_density(col, g, uvw[0], uvw[1], uvw[2], 0, 0, -1);
                        const float q5 = sample_density(col, g, uvw[0], uvw[1], uvw[2], 0, 0, 1);
                        const float grad[3] = {-q0 + q1, -q2 + q3, -q4 + q5};
                        const bool any = 0.0f < std::fabs(grad[0]) || 0.0f < std::fabs(grad[1]) || 0.0f < std::fabs(grad[2]);
                        float rd[3], wd[3], n[3];
                        for (int k = 0; k < 3; ++k) rd[k] = any ? -grad[k] : o[k];
                        for (int k = 0; k < 3; ++k) wd[k] = dp3(rd, P.world + 4 * k);
                        const float winv = rsq(dp3(wd, wd));
                        for (int k = 0; k < 3; ++k) n[k] = winv * wd[k];
                        const float yy = n[1] * n[1], zz = n[2] * n[2];
                        const float a = std::fmaf(n[0], n[0], -yy) * 0.4290427565574646f;   // 0x3edbab7e
                        const float b = std::fmaf(zz, 3.0f, -1.0f) * 0.24770796298980713f;    // 0x3e7da728
                        for (int c = 0; c < 3; ++c) {
                            float r10 = P.sh[6][c] * b;
                            r10 = std::fmaf(a, P.sh[8][c], r10);
                            float r3 = std::fmaf(P.sh[0][c], 0.8862269520759583f, r10);     // 0x3f62dfc5
                            float r8 = P.sh[4][c] * -n[0];
                            r10 = P.sh[7][c] * -n[0];
                            r10 = n[2] * r10;
                            r8 = std::fmaf(r8, -n[1], r10);
                            const float r9 = P.sh[5][c] * -n[1];
                            r8 = std::fmaf(r9, n[2], r8);
                            r3 = std::fmaf(r8, 0.8580855131149292f, r3);                    // 0x3f5bab7e
                            float r4 = P.sh[1][c] * -n[1];
                            r4 = std::fmaf(P.sh[3][c], -n[0], r4);
                            r4 = std::fmaf(P.sh[2][c], n[2], r4);
                            r3 = std::fmaf(r4, 1.0233267545700073f, r3);                    // 0x3f82fc5f
                            irr[c] = std::fmax(r3, 0.0f);
                        }
                        const float rinv = rsq(dp3(rd, rd));
                        float rdn[3];
                        for (int k = 0; k < 3; ++k) rdn[k] = rinv * rd[k];
                        ao = cast_light_ray(col, g, o, rdn, step, P.num_samples);
                    }
                }
                float rgb[3];
                for (int c = 0; c < 3; ++c) {
                    const float lc = P.light_color[3] * P.light_color[c];
                    const float amb = P.has_light_probes ? ao * irr[c] : P.ambient[3] * P.ambient[c];
                    rgb[c] = std::fmaf(shadow, lc, amb);
                }
                out[g.idx(x, y, z)] = pack_r11g11b10(rgb);
            }
}

// ---------------------------------------------------------------------------------------------
// Cube-map-space ray march with the separate light pass: CSRayMarchV.hlsl (= CSRayMarch.hlsl:98-196
// compiled with _LIGHT_PASS_), GetLocalPos :40-66, ComputeRayOrigin RayMarch.hlsli:146-177, ComputeTargetHit
// :182-187, GetLight (light-map fetch) :273-278; dispatched by Fluid::rayMarchV (Fluid.cpp:880-908) into mip
// m_cubeMapLOD of the R8G8B8A8_UNORM cube map (Fluid.cpp:229-232).  Follows Bin/CSRayMarchV.cso instruction by
// instruction (the shipped blob is built with _CPU_CUBE_FACE_CULL_ == 1: a face is marched iff its bit is set in the
// visibility mask, Fluid.cpp:51-63).  Restated platform semantics as for the light map, plus: the light map is sampled
// LINEAR_CLAMP on its decoded fp32 values; the UNORM8 store = NaN -> 0, clamp to [0, 1], * 255, + 0.5, truncate.
// A texel whose face is culled or whose ray misses the volume is NOT written (the reference leaves stale contents).
// ---------------------------------------------------------------------------------------------
struct ViewParams {          // = fxb_view_params of include/fluidx_b200.h
    float eye_pt[3];         // cbPerFrame g_eyePt, cb1[0].xyz (Fluid.cpp:302)
    float world_i[12];       // cbPerObject g_worldI, cb0[8..10] (Fluid.cpp:318)
    uint32_t num_samples;    // cbSampleRes g_numSamples = m_raySampleCount (Fluid.cpp:898)
    uint32_t visibility_mask;  // (Fluid.cpp:900)
    uint32_t cube_size;      // m_gridSize.x >> m_cubeMapLOD (Fluid.cpp:906)
};

inline void unpack_r11g11b10(uint32_t w, float out[3]) {
    const int mbits[3] = {6, 6, 5}, shift[3] = {0, 11, 22};
    for (int k = 0; k < 3; ++k) {
        const int mb = mbits[k];
        const uint32_t f = (w >> shift[k]) & ((1u << (mb + 5)) - 1u), e = f >> mb, m = f & ((1u << mb) - 1u);
        uint32_t bits;
        if (e == 31u) bits = 0x7F800000u | (m << (23 - mb));
        else if (e != 0u) bits = ((e + 112u) << 23) | (m << (23 - mb));
        else if (m == 0u) bits = 0u;
        else {  // denormal: normalise
            int sh = 0;
            uint32_t mm = m;
            while (!(mm & (1u << mb))) { mm <<= 1; ++sh; }
            bits = ((113u - sh) << 23) | ((mm & ((1u << mb) - 1u)) << (23 - mb));
        }
        std::memcpy(&out[k], &bits, 4);
    }
}

struct Taps { int x0, x1, y0, y1, z0, z1; float fx, fy, fz; };
inline Taps clamp_taps(const Grid& g, float cx, float cy, float cz) {
    const float tx = std::fmaf(cx, (float)g.nx, -0.5f), ty = std::fmaf(cy, (float)g.ny, -0.5f);
    const float tz = std::fmaf(cz, (float)g.nz, -0.5f);
    const int ix = floor_to_tap(tx), iy = floor_to_tap(ty), iz = floor_to_tap(tz);
    Taps t;
    t.fx = tx - std::floor(tx); t.fy = ty - std::floor(ty); t.fz = tz - std::floor(tz);
    t.x0 = address_tap(ix, g.nx, ADDRESS_CLAMP); t.x1 = address_tap(ix + 1, g.nx, ADDRESS_CLAMP);
    t.y0 = address_tap(iy, g.ny, ADDRESS_CLAMP); t.y1 = address_tap(iy + 1, g.ny, ADDRESS_CLAMP);
    t.z0 = address_tap(iz, g.nz, ADDRESS_CLAMP); t.z1 = address_tap(iz + 1, g.nz, ADDRESS_CLAMP);
    return t;
}
inline float trilerp(const Taps& t, const float a[8]) {  // a[k]: tap k = x + 2 y + 4 z
    const float x00 = std::fmaf(t.fx, a[1] - a[0], a[0]), x10 = std::fmaf(t.fx, a[3] - a[2], a[2]);
    const float x01 = std::fmaf(t.fx, a[5] - a[4], a[4]), x11 = std::fmaf(t.fx, a[7] - a[6], a[6]);
    const float y0v = std::fmaf(t.fy, x10 - x00, x00), y1v = std::fmaf(t.fy, x11 - x01, x01);
    return std::fmaf(t.fz, y1v - y0v, y0v);
}

// GetLight of the non-separated march (RayMarch.hlsli:280-313 with _HAS_LIGHT_PROBE_; Bin/CSRayMarch.cso instructions
// 164-298): the light reaching `pos`, computed on the spot — the light ray (and, with probes, the occlusion ray along
// the density gradient and the SH irradiance) cast with g_lightStep / g_numLightSamples — instead of read from the
// light map.  `ldir`: the normalised light direction in volume space (per ray, instructions 136-141).
inline void full_light(const uint16_t* col, const Grid& g, const LightParams& P, const float pos[3], const float uvw[3],
                       const float ldir[3], float light[3]) {
    const float lstep = 3.464101552963257f / (float)P.num_samples;
    const float shadow = cast_light_ray(col, g, pos, ldir, lstep, P.num_samples);
    float ao = 1.0f, irr[3] = {0.0f, 0.0f, 0.0f};
    if (P.has_light_probes) {
        const float q0 = sample_density(col, g, uvw[0], uvw[1], uvw[2], -1, 0, 0);
        const float q1 = sample_density(col, g, uvw[0], uvw[1], uvw[2], 1, 0, 0);
        const float q2 = sample_density(col, g, uvw[0], uvw[1], uvw[2], 0, -1, 0);
        const float q3 = sample_density(col, g, uvw[0], uvw[1], uvw[2], 0, 1, 0);
        const float q4 = sample_density(col, g, uvw[0], uvw[1], uvw[2], 0, 0, -1);
        const float q5 = sample_density(col, g, uvw[0], uvw[1], uvw[2], 0, 0, 1);
        const float grad[3] = {-q0 + q1, -q2 + q3, -q4 + q5};
        const bool any = 0.0f < std::fabs(grad[0]) || 0.0f < std::fabs(grad[1]) || 0.0f < std::fabs(grad[2]);
        float rd[3], wd[3], n[3];
        for (int k = 0; k < 3; ++k) rd[k] = any ? -grad[k] : pos[k];
        for (int k = 0; k < 3; ++k) wd[k] = dp3(rd, P.world + 4 * k);
        const float winv = rsq(dp3(wd, wd));
        for (int k = 0; k < 3; ++k) n[k] = winv * wd[k];
        const float yy = n[1] * n[1], zz = n[2] * n[2];
        const float a = std::fmaf(n[0], n[0], -yy) * 0.4290427565574646f;
        const float b = std::fmaf(zz, 3.0f, -1.0f) * 0.24770796298980713f;
        for (int c = 0; c < 3; ++c) {
            float r10 = P.sh[6][c] * b;
            r10 = std::fmaf(a, P.sh[8][c], r10);
            float r3 = std::fmaf(P.sh[0][c], 0.8862269520759583f, r10);
            float r8 = P.sh[4][c] * -n[0];
            r10 = P.sh[7][c] * -n[0];
            r10 = n[2] * r10;
            r8 = std::fmaf(r8, -n[1], r10);
            const float r9 = P.sh[5][c] * -n[1];
            r8 = std::fmaf(r9, n[2], r8);
            r3 = std::fmaf(r8, 0.8580855131149292f, r3);
            float r4 = P.sh[1][c] * -n[1];
            r4 = std::fmaf(P.sh[3][c], -n[0], r4);
            r4 = std::fmaf(P.sh[2][c], n[2], r4);
            r3 = std::fmaf(r4, 1.0233267545700073f, r3);
            irr[c] = std::fmax(r3, 0.0f);
        }
        const float rinv = rsq(dp3(rd, rd));
        float rdn[3];
        for (int k = 0; k < 3; ++k) rdn[k] = rinv * rd[k];
        ao = cast_light_ray(col, g, pos, rdn, lstep, P.num_samples);
    }
    for (int c = 0; c < 3; ++c) {
        const float lc = P.light_color[3] * P.light_color[c];
        const float amb = P.has_light_probes ? ao * irr[c] : P.ambient[3] * P.ambient[c];
        light[c] = std::fmaf(lc, shadow, amb);
    }
}

// lmap != nullptr: CSRayMarchV (light read from the light map); lmap == nullptr: CSRayMarch (light computed per sample, LP)
void ray_march_v(const Grid& g, const uint16_t* col, const uint32_t* lmap, const ViewParams& P, uint8_t* cube,
                 const LightParams* LP = nullptr) {
    const int S = (int)P.cube_size;
    const float FMAX = 3.402823466e+38f;
#pragma omp parallel for collapse(2) schedule(dynamic, 4)
    for (int face = 0; face < 6; ++face)
        for (int y = 0; y < S; ++y)
            for (int x = 0; x < S; ++x) {
                if (((1u << face) & P.visibility_mask) == 0u) continue;
                float ro[3];
                for (int k = 0; k < 3; ++k) {
                    const float* w = P.world_i + 4 * k;
                    ro[k] = ((P.eye_pt[0] * w[0] + P.eye_pt[1] * w[1]) + P.eye_pt[2] * w[2]) + 1.0f * w[3];
                }
                const float X = std::fmaf(((float)x + 0.5f) / (float)S, 2.0f, -1.0f);
                const float Z = std::fmaf(((float)y + 0.5f) / (float)S, 2.0f, -1.0f);
                float tg[3];
                switch (face) {
                    case 0: tg[0] = 1.0f; tg[1] = -Z; tg[2] = -X; break;
                    case 1: tg[0] = -1.0f; tg[1] = -Z; tg[2] = X; break;
                    case 2: tg[0] = X; tg[1] = 1.0f; tg[2] = Z; break;
                    case 3: tg[0] = X * 1.0f; tg[1] = -1.0f; tg[2] = Z * -1.0f; break;
                    case 4: tg[0] = X * 1.0f; tg[1] = Z * -1.0f; tg[2] = 1.0f; break;
                    default: tg[0] = -X; tg[1] = -Z; tg[2] = -1.0f; break;
                }
                float d[3], dir[3];
                for (int k = 0; k < 3; ++k) d[k] = -ro[k] + tg[k];
                const float inv = rsq(dp3(d, d));
                for (int k = 0; k < 3; ++k) dir[k] = inv * d[k];
                bool hit = true;
                if (!(1.0f >= std::fabs(ro[0]) && 1.0f >= std::fabs(ro[1]) && 1.0f >= std::fabs(ro[2]))) {
                    float r3[3], u[3];
                    for (int k = 0; k < 3; ++k) {
                        const int sg = (0.0f < dir[k] ? -1 : 0) - (dir[k] < 0.0f ? -1 : 0);  // = -sign(dir)
                        r3[k] = -ro[k] + (float)sg;
                    }
                    u[0] = r3[0] / dir[0];
                    u[1] = r3[1] / dir[1];
                    float U = FMAX;
                    hit = false;
                    if (u[0] >= 0.0f && 1.0f >= std::fabs(std::fmaf(dir[1], u[0], ro[1]))) {
                        if (1.0f >= std::fabs(std::fmaf(dir[2], u[0], ro[2]))) { hit = u[0] < FMAX; U = std::fmin(u[0], FMAX); }
                    }
                    if (u[1] >= 0.0f && 1.0f >= std::fabs(std::fmaf(dir[2], u[1], ro[2]))) {
                        const bool bb = 1.0f >= std::fabs(std::fmaf(dir[0], u[1], ro[0]));
                        if (bb && u[1] < U) { U = u[1]; hit = true; }
                    }
                    u[2] = r3[2] / dir[2];
                    if (u[2] >= 0.0f && 1.0f >= std::fabs(std::fmaf(dir[0], u[2], ro[0]))) {
                        const bool bb = 1.0f >= std::fabs(std::fmaf(dir[1], u[2], ro[1]));
                        if (bb && u[2] < U) { U = u[2]; hit = true; }
                    }
                    for (int k = 0; k < 3; ++k) ro[k] = std::fmin(std::fmax(std::fmaf(dir[k], U, ro[k]), -1.0f), 1.0f);
                }
                if (!hit) continue;
                const float step = 3.464101552963257f / (float)P.num_samples;
                float tm[3];
                for (int k = 0; k < 3; ++k) tm[k] = (tg[k] + -ro[k]) / dir[k];
                const float tmax = std::fmax(tm[2], std::fmax(tm[1], tm[0]));
                float ldir[3] = {0.0f, 0.0f, 0.0f};
                if (!lmap) {
                    float Lv[3];
                    for (int k = 0; k < 3; ++k) Lv[k] = dp3(LP->light_pt, LP->world_i + 4 * k);
                    const float linv = rsq(dp3(Lv, Lv));
                    for (int k = 0; k < 3; ++k) ldir[k] = linv * Lv[k];
                }
                float sc[4] = {0.0f, 0.0f, 0.0f, 0.0f}, t = 0.0f, prev = 0.0f;
                for (uint32_t i = 0; i < P.num_samples; ++i) {
                    float pos[3];
                    for (int k = 0; k < 3; ++k) pos[k] = std::fmaf(dir[k], t, ro[k]);
                    if (1.0f < std::fabs(pos[0]) || 1.0f < std::fabs(pos[1]) || 1.0f < std::fabs(pos[2])) break;
                    const float uvw[3] = {std::fmaf(pos[0], 0.5f, 0.5f), std::fmaf(pos[1], 0.5f, 0.5f),
                                          std::fmaf(pos[2], 0.5f, 0.5f)};
                    const Taps tp = clamp_taps(g, uvw[0], uvw[1], uvw[2]);
                    const size_t idx[8] = {g.idx(tp.x0, tp.y0, tp.z0), g.idx(tp.x1, tp.y0, tp.z0), g.idx(tp.x0, tp.y1, tp.z0),
                                           g.idx(tp.x1, tp.y1, tp.z0), g.idx(tp.x0, tp.y0, tp.z1), g.idx(tp.x1, tp.y0, tp.z1),
                                           g.idx(tp.x0, tp.y1, tp.z1), g.idx(tp.x1, tp.y1, tp.z1)};
                    float c[4], a[8];
                    for (int ch = 0; ch < 4; ++ch) {
                        for (int k = 0; k < 8; ++k) a[k] = half_to_float(col[4 * idx[k] + ch]);
                        c[ch] = trilerp(tp, a);
                    }
                    float r5[4], new_step;
                    if (0.01f < c[3]) {
                        float L[3], tex[8][3];
                        if (lmap) {
                            for (int k = 0; k < 8; ++k) unpack_r11g11b10(lmap[idx[k]], tex[k]);
                            for (int ch = 0; ch < 3; ++ch) {
                                for (int k = 0; k < 8; ++k) a[k] = tex[k][ch];
                                L[ch] = trilerp(tp, a);
                            }
                        } else {
                            full_light(col, g, *LP, pos, uvw, ldir, L);
                        }
                        const float transm = -sc[3] + 1.0f;
                        const float ev = std::fmin(0.00390625f / std::fabs(-prev + c[3]), 2.0f);
                        const float ui = std::fmin(-c[3] + 1.0f, 1.0f);
                        const float th = -transm + 1.0f;
                        new_step = std::fmax(th * (ui * (ev * 1.5f)), 1.0f) * step;
                        for (int ch = 0; ch < 3; ++ch) r5[ch] = std::fmaf(transm * (L[ch] * c[ch]), 0.8f, sc[ch]);
                        r5[3] = std::fmaf(0.8f * c[3], transm, sc[3]);
                        if (transm < 0.01f) { std::memcpy(sc, r5, sizeof sc); break; }
                        prev = c[3];
                    } else {
                        std::memcpy(r5, sc, sizeof sc);
                        new_step = step;
                    }
                    t = t + new_step;
                    std::memcpy(sc, r5, sizeof sc);
                    if (tmax < t) break;
                }
                uint8_t* o = cube + 4 * (((size_t)face * S + y) * S + x);
                for (int ch = 0; ch < 4; ++ch) {
                    float v = ch < 3 ? sc[ch] * 0.15915493667125702f : sc[ch];  // 0x3e22f983 = 1 / (2 pi)
                    v = std::isnan(v) ? 0.0f : std::fmin(std::fmax(v, 0.0f), 1.0f);
                    o[ch] = (uint8_t)(v * 255.0f + 0.5f);
                }
            }
}

struct Oracle {
    Grid g;
    int address_mode, early_exit, iters;
    std::vector<uint16_t> vel[2], col[2];
    std::vector<float> p, q, s;
    std::vector<uint8_t> active;
    std::vector<int64_t> hist;
    int parity = 0;
    float dt = 0.0f;
    int s_exec = 0;
};

}  // namespace

extern "C" {

// OpenMP thread count used by every entry point below: n > 0 sets it; returns the count now in effect.  (Under
// torch.distributed.run the environment says OMP_NUM_THREADS=1, so callers that time the oracle set it explicitly.)
int fxo_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}


enum { FXO_VEL = 0, FXO_COLOR = 1, FXO_PRESSURE = 2, FXO_VEL_ADVECTED = 3, FXO_COLOR_PREV = 4 };

void* fxo_create(int nx, int ny, int nz, int address_mode, int early_exit, int iters) {
    if (nx < 1 || ny < 1 || nz < 1 || nx != ny || iters < 0 || iters > 4096) return nullptr;  // Fluid.cpp:201
    Oracle* o = new Oracle;
    o->g = Grid{nx, ny, nz};
    o->address_mode = address_mode;
    o->early_exit = early_exit;
    o->iters = iters;
    const size_t n = o->g.voxels();
    for (int i = 0; i < 2; ++i) { o->vel[i].assign(4 * n, 0); o->col[i].assign(4 * n, 0); }  // App. B.1
    o->p.assign(n, 0.0f); o->q.assign(n, 0.0f); o->s.assign(n, 0.0f);
    o->active.assign(n, 1);
    o->hist.assign(std::max(iters, 1), 0);
    return o;
}

void fxo_destroy(void* h) { delete static_cast<Oracle*>(h); }

// Fluid::UpdateFrame, simulation part (Fluid.cpp:288-290, 344-345).
void fxo_update_frame(void* h, float dt) {
    Oracle* o = static_cast<Oracle*>(h);
    o->dt = dt;
    if (dt > 0.0f) o->parity ^= 1;
}

// Fluid::Simulate (Fluid.cpp:348-410): advect vel[0],color[!p] -> vel[1],color[p]; project vel[1] -> vel[0].
void fxo_simulate(void* h) {
    Oracle* o = static_cast<Oracle*>(h);
    const int p = o->parity;
    advect(o->g, o->address_mode, o->dt, o->vel[0].data(), o->col[!p].data(), o->vel[1].data(), o->col[p].data());
    if (0.0f < o->dt) {
        divergence2x(o->g, o->vel[1].data(), o->s.data());
        o->s_exec = jacobi(o->g, o->s.data(), o->p.data(), o->q.data(), o->active.data(), o->iters,
                           o->early_exit, o->hist.data());
        gradient(o->g, o->vel[1].data(), o->p.data(), o->vel[0].data());
    } else {
        o->s_exec = 0;
        o->vel[0] = o->vel[1];  // CSProject3D.hlsl:88,112 — identity copy of .xyz
    }
}

static void* field_ptr(Oracle* o, int field, size_t* bytes) {
    const size_t n = o->g.voxels();
    switch (field) {
        case FXO_VEL: *bytes = 8 * n; return o->vel[0].data();
        case FXO_COLOR: *bytes = 8 * n; return o->col[o->parity].data();
        case FXO_PRESSURE: *bytes = 4 * n; return o->p.data();
        case FXO_VEL_ADVECTED: *bytes = 8 * n; return o->vel[1].data();
        case FXO_COLOR_PREV: *bytes = 8 * n; return o->col[!o->parity].data();
    }
    *bytes = 0;
    return nullptr;
}

int fxo_get_field(void* h, int field, void* out, size_t bytes) {
    size_t need; void* src = field_ptr(static_cast<Oracle*>(h), field, &need);
    if (!src || bytes != need) return -1;
    std::memcpy(out, src, need);
    return 0;
}

int fxo_set_field(void* h, int field, const void* in, size_t bytes) {
    size_t need; void* dst = field_ptr(static_cast<Oracle*>(h), field, &need);
    if (!dst || bytes != need) return -1;
    std::memcpy(dst, in, need);
    return 0;
}

int fxo_s_exec(void* h) { return static_cast<Oracle*>(h)->s_exec; }

int fxo_active_hist(void* h, int64_t* out, int n) {
    Oracle* o = static_cast<Oracle*>(h);
    const int m = std::min<int>(n, (int)o->hist.size());
    std::copy(o->hist.begin(), o->hist.begin() + m, out);
    return m;
}

// FluidX12.cpp:266-267
float fxo_dt_for_grid(int nx, int ny, int nz) { (void)nx; return (nz > 1 ? 2.0f : 1.0f) / (float)ny; }

// ---- stage-level entry points (kernel-by-kernel parity tests) --------------------------------
void fxo_advect(int nx, int ny, int nz, int mode, float dt, const uint16_t* vel_in, const uint16_t* col_in,
                uint16_t* vel_out, uint16_t* col_out) {
    advect(Grid{nx, ny, nz}, mode, dt, vel_in, col_in, vel_out, col_out);
}
void fxo_divergence2x(int nx, int ny, int nz, const uint16_t* vel, float* s) { divergence2x(Grid{nx, ny, nz}, vel, s); }
int fxo_jacobi(int nx, int ny, int nz, const float* s, float* p, int iters, int early_exit, int64_t* hist,
               uint8_t* active_out) {
    const Grid g{nx, ny, nz};
    std::vector<float> q(g.voxels());
    std::vector<uint8_t> active(g.voxels());
    const int r = jacobi(g, s, p, q.data(), active.data(), iters, early_exit, hist);
    if (active_out) std::memcpy(active_out, active.data(), active.size());
    return r;
}
void fxo_gradient(int nx, int ny, int nz, const uint16_t* vel_in, const float* p, uint16_t* vel_out) {
    gradient(Grid{nx, ny, nz}, vel_in, p, vel_out);
}

// ---- z-slab window entry points (decomposition tests: tests/test_slab_gloo.py) ----------------------------------
void fxo_advect_slab(int nx, int ny, int nz, int nzg, int z0, int mode, float dt, const uint16_t* vel_in,
                     const uint16_t* col_in, uint16_t* vel_out, uint16_t* col_out) {
    Grid g{nx, ny, nz}; g.nzg = nzg; g.z0 = z0;
    advect(g, mode, dt, vel_in, col_in, vel_out, col_out);
}
void fxo_divergence2x_slab(int nx, int ny, int nz, int nzg, int z0, const uint16_t* vel, float* s) {
    Grid g{nx, ny, nz}; g.nzg = nzg; g.z0 = z0;
    divergence2x(g, vel, s);
}
// Exactly `nsweeps` synchronous sweeps over the whole window starting from the given freeze flags (in/out);
// counts[k] = cells of local planes [c0, c1) still active after sweep k.  Cells near a window end that is not a
// grid face become invalid one plane per sweep, exactly like the halo of a fused pass.
void fxo_jacobi_sweeps_slab(int nx, int ny, int nz, int nzg, int z0, const float* s, float* p, uint8_t* active,
                            int nsweeps, int early_exit, int c0, int c1, int64_t* counts) {
    Grid g{nx, ny, nz}; g.nzg = nzg; g.z0 = z0;
    const float inv = 0.166666672f;
    std::vector<float> q(g.voxels());
    float* cur = p; float* nxt = q.data();
    for (int k = 0; k < nsweeps; ++k) {
        int64_t still = 0;
#pragma omp parallel for collapse(2) schedule(static) reduction(+ : still)
        for (int z = 0; z < nz; ++z)
            for (int y = 0; y < ny; ++y)
                for (int x = 0; x < nx; ++x) {
                    const size_t i = g.idx(x, y, z);
                    if (!active[i]) { nxt[i] = cur[i]; continue; }
                    const Nbr nb = neighbours(g, x, y, z, true);
                    float acc = std::fmaf(-s[i], 0.5f, cur[nb.L]);
                    acc = cur[nb.R] + acc; acc = cur[nb.U] + acc; acc = cur[nb.D] + acc;
                    acc = cur[nb.F] + acc; acc = cur[nb.B] + acc;
                    nxt[i] = acc * inv;
                    if (early_exit && std::fabs(std::fmaf(acc, inv, -cur[i])) < 0.00100000005f) active[i] = 0;
                    else if (z >= c0 && z < c1) ++still;
                }
        counts[k] = still;
        std::swap(cur, nxt);
    }
    if (cur != p) std::memcpy(p, cur, g.voxels() * sizeof(float));
}
void fxo_gradient_slab(int nx, int ny, int nz, int nzg, int z0, const uint16_t* vel_in, const float* p,
                       uint16_t* vel_out) {
    Grid g{nx, ny, nz}; g.nzg = nzg; g.z0 = z0;
    gradient(g, vel_in, p, vel_out);
}

// ---- unit helpers -----------------------------------------------------------------------------
// Light-map pass on a whole grid: colour = [nz][ny][nx][4] half, params = LightParams, out = [nz][ny][nx] packed
// R11G11B10_FLOAT words.
void fxo_light_map(int nx, int ny, int nz, const uint16_t* colour, const void* params, uint32_t* out) {
    light_map(Grid{nx, ny, nz}, colour, *static_cast<const LightParams*>(params), out);
}
// Cube-map ray march: light_map = fxo_light_map's output; cube = [6][S][S][4] UNORM8, S = params.cube_size; texels of
// culled faces / rays that miss the volume are left untouched.
void fxo_ray_march_v(int nx, int ny, int nz, const uint16_t* colour, const uint32_t* light_map, const void* params,
                     uint8_t* cube) {
    ray_march_v(Grid{nx, ny, nz}, colour, light_map, *static_cast<const ViewParams*>(params), cube);
}
// The non-separated march (CSRayMarch): `light` = LightParams, whose num_samples is the light-ray sample count.
void fxo_ray_march(int nx, int ny, int nz, const uint16_t* colour, const void* view, const void* light, uint8_t* cube) {
    ray_march_v(Grid{nx, ny, nz}, colour, nullptr, *static_cast<const ViewParams*>(view), cube,
                static_cast<const LightParams*>(light));
}
void fxo_unpack_r11g11b10(uint32_t w, float* out3) { unpack_r11g11b10(w, out3); }
uint32_t fxo_pack_r11g11b10(float r, float g, float b) {
    const float rgb[3] = {r, g, b};
    return pack_r11g11b10(rgb);
}

uint16_t fxo_f32_to_f16(float f) { return float_to_half(f); }
float fxo_f16_to_f32(uint16_t h) { return half_to_float(h); }
int fxo_address_tap(int i, int w, int mode) { return address_tap(i, w, mode); }
void fxo_sample_trilinear(const uint16_t* field, int nx, int ny, int nz, int mode, float cx, float cy, float cz,
                          float* out4) {
    sample_trilinear(field, Grid{nx, ny, nz}, mode, cx, cy, cz, out4);
}
float fxo_emitter_basis(int nx, int ny, int nz, int x, int y, int z) {
    float d[3];
    return emitter_basis(Grid{nx, ny, nz}, x, y, z, d);
}
// fp32 literals the restatement uses, in a fixed order, for the DXBC-literal test.
int fxo_constants(float* out, int n) {
    const float c[] = {1.04166675f, 33.3333359f, 0.166666672f, 0.25f, 0.970000029f, 0.00100000005f,
                       0.200000003f, -0.100000001f, 1.44269502f, 0.0183156393f, 0.00390625f, 0.0009765625f,
                       192.0f, 48.0f, 200.0f, 8.0f, 16.0f, 40.0f};
    const int m = std::min<int>(n, int(sizeof(c) / sizeof(c[0])));
    std::copy(c, c + m, out);
    return m;
}
int fxo_has_f16c(void) {
#if defined(__F16C__)
    return 1;
#else
    return 0;
#endif
}

}  // extern "C"
