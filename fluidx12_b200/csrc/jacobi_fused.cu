// jacobi_fused.cu — T Jacobi sweeps of the pressure solve fused into one HBM pass (sm_100a).
//
// Replaces the relaxation loop of FluidX12/Content/Shaders/CSPoisson.hlsli:8-26 as called from
// CSProject3D.hlsl:93 (dispatch Fluid.cpp:394-408), under the deterministic restatement of SURVEY.md
// App. A.3: synchronous Jacobi, per-cell freeze once |x - x0| < 0.001, at most ITER sweeps.
//
// Scheme: z-marching 2.5-D blocking with temporal fusion.  A CTA owns a brick of 120 x (TILE_Y - 2T) x BZ output
// cells (TILE_Y = rows per thread x warps; the default shape is T = 2, 2 rows x 8 warps = 16 rows, BZ = 8, two
// CTAs per SM).  It streams the xy tile (128 x TILE_Y cells: halo 4 in x, T in y) plane by plane along z; level l
// (= number of sweeps applied) of plane k-l is produced in iteration k, so the T sweeps advance in lock-step, each
// one plane behind the previous:
//   * level-0 pressure planes and the right-hand-side planes are staged into shared memory by TMA
//     (cp.async.bulk.tensor.3d + mbarrier), DEPTH iterations ahead; out-of-grid tile parts are zero-filled;
//   * every thread keeps its own column (ROWS rows x 4 cells) of the two most recent planes of every level in
//     registers (the z queue), so the z neighbours never touch memory; a sweep is split into a 5-addition head and
//     a 1-addition tail so that a level's new plane overwrites the queue slot its consumer has just read;
//   * x neighbours come from warp shuffles (a warp spans the 128-cell tile row), y neighbours from the thread's own
//     rows or from the edge rows every warp publishes in shared memory;
//   * the reference's clamp-to-edge neighbour rule (CSProject3D.hlsl:76-83) is applied by index (x, y) or by
//     reusing the centre value (z), never by TMA fill;
//   * the per-cell freeze flags travel with the values (4 bits per quad per level); a warp whose cells are all
//     frozen at a level skips that level's arithmetic; flags persist between passes in a bit-packed array
//     (1 bit per cell, 0.25 B/voxel/pass of traffic);
//   * persistent CTAs take bricks from device-side work lists; a brick whose cells are all frozen is copied once to
//     the other pressure buffer and skipped for the rest of the frame (final in both ping-pong buffers).
// Algorithmic traffic per relaxed cell per pass: p in 4 + rhs in 4 + p out 4 (+ 2/8 mask) bytes; 8 per copied cell.
#include <cuda.h>

#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "kernels.h"

namespace fxb {

namespace {

constexpr int kLanes = 32;
constexpr int kTileX = 4 * kLanes;       // 128 cells per tile row: one warp spans a row, 4 cells per lane
constexpr int kHaloX = 4;                // one quad
constexpr int kOutX = kTileX - 2 * kHaloX;  // 120
constexpr float kInv6 = 0.166666672f;
constexpr float kEps = 0.00100000005f;

// Compile-time shape of one kernel variant.
//   T     sweeps fused per pass;  ROWS rows per thread;  WARPS warps per CTA (tile = 128 x ROWS*WARPS cells);
//   DEPTH TMA bundles in flight ahead of the one being consumed.
template <int T_, int ROWS, int WARPS, int DEPTH, int CTAS = 1>
struct Shape {
    static constexpr int T = T_, kRows = ROWS, kWarps = WARPS, kPrefetch = DEPTH;
    static constexpr int kThreads = kLanes * WARPS;
    static constexpr int kCtasPerSm = CTAS;
    static constexpr int kTileY = ROWS * WARPS;
    static constexpr int kPlane = kTileX * kTileY;       // floats per TMA-staged plane
    static constexpr int kEdgePlane = WARPS * 2 * kTileX;  // floats per published level plane: first and last row of every warp
    static constexpr int kP0Slots = 2 + DEPTH;   // planes k-1 (neighbours), k (own), k+1.. (in flight)
    static constexpr int kRhsSlots = T_ + DEPTH;  // planes k-T .. k-1 in use, k .. in flight
    static constexpr size_t kFloats = (size_t)(kP0Slots + kRhsSlots) * kPlane + (size_t)2 * (T_ - 1) * kEdgePlane;
    static constexpr size_t kBytes = kFloats * sizeof(float) + 96;  // + barriers (<= 6) and counters
    static_assert(DEPTH + 1 <= 6, "barrier slots");
    static_assert(kBytes * CTAS <= 232448, "shared memory budget (227 KB per SM)");
    static_assert(ROWS >= 2, "a thread publishes its first and last row separately");
};

#define FXB_SHAPE_CONSTANTS(S)                                                                              \
    constexpr int T = S::T, kRows = S::kRows, kWarps = S::kWarps, kPrefetch = S::kPrefetch;                   \
    constexpr int kThreads = S::kThreads, kTileY = S::kTileY, kPlane = S::kPlane, kEdgePlane = S::kEdgePlane; \
    (void)kWarps; (void)kThreads; (void)kTileY; (void)kPlane; (void)kEdgePlane; (void)kPrefetch; (void)kRows; (void)T

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int x, int y, int z, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}

struct PassParams {
    int nx, ny;            // grid extent in x, y
    int nz_alloc;          // local planes allocated
    int z_face_lo;         // local index of global plane 0 (or very negative when it is on another rank)
    int z_face_hi;         // local index one past global plane nz-1 (informational: the array ends at that face)
    int z_out0, z_out1;    // local planes this rank must produce
    int bz;                // planes per brick
    int ntx, nty, nzc;     // brick grid
    int pass;              // index of this fused pass in the frame
    int levels_total;      // ITER
    int early_exit;
    int run_all;           // multi-GPU: never end the solve on this rank's own freeze counters
    int ext_lo, ext_hi;    // multi-GPU: planes below / above the owned range to relax redundantly in this pass
};

struct WorkLists {
    int* relax[2];     // bricks that still hold an active cell, ping-pong by pass parity
    int* copy[2];      // bricks that froze in the previous pass: one copy into the other pressure buffer
    int* relax_count;  // [pass]
    int* copy_count;   // [pass]
};

// Position of a launch in the frame's relax sequence: the static schedule keeps it in the launch parameters (constant
// bank), the dynamic one (DYN) reads it from StepState::seq at run time.
template <bool DYN>
__device__ __forceinline__ int pass_of(const PassParams& P, const int seq) {
    if constexpr (DYN) return seq;
    else return P.pass;
}

// A brick whose cells all froze during pass p-1 holds its final values in that pass's output buffer only.  Pass p
// copies its own region once into the other buffer (and clears the other mask buffer), after which the brick is
// final in both ping-pong buffers and is never touched again in this frame.  Pure streaming (8 B/cell), eight
// independent 16-byte loads in flight per thread.
template <class S>
__device__ __noinline__ void copy_frozen_brick(const float* __restrict__ p_in, float* __restrict__ p_out,
                                               unsigned char* __restrict__ m_out, const PassParams& P,
                                               const int brick) {
    FXB_SHAPE_CONSTANTS(S);
    constexpr int kOutY = kTileY - 2 * T;
    const int tid = threadIdx.x, nxb = P.nx >> 3;
    const int tx = brick % P.ntx, ty = (brick / P.ntx) % P.nty, zc_idx = brick / (P.ntx * P.nty);
    const int zs = P.z_out0 + zc_idx * P.bz, ze = min(zs + P.bz, P.z_out1);
    const int x_lo = tx * kOutX, y_lo = ty * kOutY;
    const int rows = min(kOutY, P.ny - y_lo), planes = ze - zs;
    const int qpr = min(kOutX, P.nx - x_lo) >> 2;  // float4 per row inside the grid
    const int total = planes * rows * qpr;
    for (int base = tid; base < total; base += 8 * kThreads) {
        float4 v[8];
        size_t at[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * kThreads;
            const int xq = i % qpr, rz = i / qpr;
            at[u] = ((size_t)(zs + rz / rows) * P.ny + (y_lo + rz % rows)) * P.nx + x_lo + 4 * xq;
            if (i < total) v[u] = __ldcs(reinterpret_cast<const float4*>(p_in + at[u]));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (base + u * kThreads < total) *reinterpret_cast<float4*>(p_out + at[u]) = v[u];
    }
    const int bpr = qpr >> 1;  // mask bytes per row
    for (int i = tid; i < planes * rows * bpr; i += kThreads) {
        const int xb = i % bpr, rz = i / bpr;
        m_out[((size_t)(zs + rz / rows) * P.ny + (y_lo + rz % rows)) * nxb + (x_lo >> 3) + xb] = 0;
    }
}

// First five additions of one relaxation of a quad, in the DXBC's order:
// acc = p[L] + rhs; acc = p[R] + acc; acc = p[U] + acc; acc = p[D] + acc; acc = p[F] + acc.
// The x neighbours live in other registers of the same quad, so those two additions are scalar; the rest are
// packed FADD2 on the (x,y) / (z,w) halves.
__device__ __forceinline__ float4 relax_head(const float4 c, const float4 lo, const float4 up, const float4 dn,
                                             const float left, const float right, const float4 rhs) {
    float2 a = make_float2(left + rhs.x, c.x + rhs.y);
    float2 b = make_float2(c.y + rhs.z, c.z + rhs.w);
    a = make_float2(c.y + a.x, c.z + a.y);
    b = make_float2(c.w + b.x, right + b.y);
    a = add2(make_float2(up.x, up.y), a);
    b = add2(make_float2(up.z, up.w), b);
    a = add2(make_float2(dn.x, dn.y), a);
    b = add2(make_float2(dn.z, dn.w), b);
    a = add2(make_float2(lo.x, lo.y), a);
    b = add2(make_float2(lo.z, lo.w), b);
    return make_float4(a.x, a.y, b.x, b.y);
}

// Last addition (acc = p[B] + acc), x = acc * (1/6), freeze test |fma(acc, 1/6, -x0)| < eps (eps < 0: never),
// frozen cells keep x0.  `act` / `still`: 4 flag bits of the quad before / after this sweep.
__device__ __forceinline__ void relax_tail(const float4 head, const float4 hi, const float4 c, const unsigned act,
                                           const float eps, float4& out, unsigned& still) {
    const float2 inv2 = make_float2(kInv6, kInv6);
    const float2 a = add2(make_float2(hi.x, hi.y), make_float2(head.x, head.y));
    const float2 b = add2(make_float2(hi.z, hi.w), make_float2(head.z, head.w));
    const float2 na = mul2(a, inv2), nb = mul2(b, inv2);
    const float2 da = fma2(a, inv2, make_float2(-c.x, -c.y)), db = fma2(b, inv2, make_float2(-c.z, -c.w));
    unsigned s = act;
    if (fabsf(da.x) < eps) s &= ~1u;
    if (fabsf(da.y) < eps) s &= ~2u;
    if (fabsf(db.x) < eps) s &= ~4u;
    if (fabsf(db.y) < eps) s &= ~8u;
    out.x = (act & 1u) ? na.x : c.x;
    out.y = (act & 2u) ? na.y : c.y;
    out.z = (act & 4u) ? nb.x : c.z;
    out.w = (act & 8u) ? nb.y : c.w;
    still = s;
}

// Relaxes one brick: T fused sweeps over its 120 x (TILE_Y - 2T) x (ze - zs) output cells (see the file header).
template <class S, bool DYN = false>
__device__ __forceinline__ void relax_brick(const CUtensorMap* map_in, const CUtensorMap* map_rhs_p,
                                         float* __restrict__ p_out, const unsigned char* __restrict__ m_in,
                                         unsigned char* __restrict__ m_out, StepState* __restrict__ state,
                                         const WorkLists& W, const PassParams& P, const int brick, const int tx,
                                         const int ty, const int zs, const int ze, const int levels, const int s0,
                                         const int seq = 0) {
    FXB_SHAPE_CONSTANTS(S);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int kOutY = kTileY - 2 * T;
    extern __shared__ __align__(1024) float sm[];                  // TMA destinations need 128-byte alignment
    float* sm_p0 = sm;                                             // [kP0Slots][kPlane]   level-0 planes (TMA)
    float* sm_rhs = sm_p0 + S::kP0Slots * kPlane;            // [kRhsSlots][kPlane]  rhs planes (TMA)
    float* sm_lev = sm_rhs + S::kRhsSlots * kPlane;          // [T-1][2][kEdgePlane] levels 1..T-1, edge rows
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::kFloats);  // [kPrefetch + 1]
    unsigned* s_cnt = reinterpret_cast<unsigned*>(bars + 6);               // [T]

    // brick >= 0: a brick of this rank's own planes (tracked in the work lists and freeze counters);
    // brick < 0: planes of the z-halo relaxed redundantly between two exchanges (multi-GPU): never listed or counted.
    const int gx0 = tx * kOutX - kHaloX;
    const int gy0 = ty * kOutY - T;
    const int nxb = P.nx >> 3;  // mask bytes per row

    const int gx = gx0 + 4 * lane;
    const bool qin = gx >= 0 && gx < P.nx;
    const bool own_lane = lane >= 1 && lane <= 30 && qin;
    const int gyb = gy0 + kRows * warp;  // grid y of this thread's row 0; row r is gyb + r
    unsigned own_bits = 0;  // bit (4r + j): cell j of row r belongs to this brick's output region
    unsigned dom_bits = 0;  // bit (4r + j): cell lies inside the grid
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int ry = kRows * warp + r;
        const bool rin = gyb + r >= 0 && gyb + r < P.ny;
        if (rin && qin) dom_bits |= 0xFu << (4 * r);
        if (rin && own_lane && ry >= T && ry < kTileY - T) own_bits |= 0xFu << (4 * r);
    }

    // planes available for loading and the z range each level must cover (trapezoid in z)
    const int zl0 = max(zs - T, 0), zl1 = min(ze + T, P.nz_alloc);
    auto lev_lo = [&](int l) { return max(zs - (T - l), 0); };
    auto lev_hi = [&](int l) { return min(ze + (T - l), P.nz_alloc); };

    // per-thread shared-memory offsets (floats)
    const int off0 = kRows * warp * kTileX + 4 * lane;  // own quad of row 0 inside a full (TMA) plane; row r: + r * kTileX
    const bool clamp_u = warp == 0 || gyb <= 0;                         // no row above inside the grid/tile
    const bool clamp_d = warp == kWarps - 1 || (gyb + kRows - 1) >= P.ny - 1;  // no row below
    const int off_up = clamp_u ? off0 : off0 - kTileX;
    const int off_dn = clamp_d ? (off0 + (kRows - 1) * kTileX) : (off0 + (kRows - 1) * kTileX) + kTileX;
    // published level planes keep only rows 0 and 3 of every warp: [warp][top|bottom][128]
    const int eoff_top = (warp * 2 + 0) * kTileX + 4 * lane, eoff_bot = (warp * 2 + 1) * kTileX + 4 * lane;
    const int eoff_up = clamp_u ? eoff_top : eoff_bot - 2 * kTileX;  // bottom row of the warp above
    const int eoff_dn = clamp_d ? eoff_bot : eoff_top + 2 * kTileX;  // top row of the warp below
    const bool clamp_l = lane == 0 || gx == 0;
    const bool clamp_r = lane == 31 || gx + 4 == P.nx;
    // the grid's y faces may cut through this warp's rows: then the in-register y neighbours need clamping
    bool y_edge = false;
#pragma unroll
    for (int r = 0; r < kRows; ++r) y_edge |= ((gyb + r) == 0 && r > 0) || ((gyb + r) == P.ny - 1 && r < kRows - 1);

    auto issue_bundle = [&](int k) {  // p plane k and rhs plane k-1 -> shared memory (thread 0 only)
        const bool has_p = k >= zl0 && k < zl1, has_r = k - 1 >= zl0 && k - 1 < zl1;
        if (!has_p && !has_r) return;
        uint64_t* bar = &bars[(k - zl0) % (kPrefetch + 1)];
        mbar_expect_tx(bar, (uint32_t)((has_p ? 1 : 0) + (has_r ? 1 : 0)) * kPlane * 4u);
        if (has_p) tma_load_3d(sm_p0 + ((k - zl0) % S::kP0Slots) * kPlane, map_in, gx0, gy0, k, bar);
        if (has_r)
            tma_load_3d(sm_rhs + ((k - 1 - zl0) % S::kRhsSlots) * kPlane, map_rhs_p, gx0, gy0, k - 1, bar);
    };
    // Freeze flags of the level-0 planes (the previous pass's output mask).  The raw bytes are fetched two
    // iterations ahead and only decoded when their plane is consumed, so the load latency stays hidden.
    auto fetch_flag_bytes = [&](int z, unsigned (&raw)[kRows]) {
        if (pass_of<DYN>(P, seq) == 0 || z >= zl1) return;
#pragma unroll
        for (int r = 0; r < kRows; ++r)
            if ((dom_bits >> (4 * r)) & 1u) raw[r] = __ldg(m_in + ((size_t)z * P.ny + (gyb + r)) * nxb + (gx >> 3));
    };
    auto decode_flags = [&](const unsigned (&raw)[kRows]) -> unsigned {
        if (pass_of<DYN>(P, seq) == 0) return dom_bits;
        unsigned f = 0;
#pragma unroll
        for (int r = 0; r < kRows; ++r) f |= ((raw[r] >> (gx & 4)) & 0xFu) << (4 * r);
        return f & dom_bits;
    };

    // z queue.  Level l (0..T-1) keeps two planes per cell in registers: at iteration `it` slot [it & 1] holds the
    // older plane (the "F" neighbour of the consumer) and slot [~it & 1] the centre plane; the level's newest
    // plane overwrites the older slot once the consumer has taken it, so the roles swap every iteration and no
    // value is ever moved between registers (the loop is unrolled by two to make the slots compile-time).
    float4 q[T][2][kRows];
    unsigned fl[T][2];
#pragma unroll
    for (int l = 0; l < T; ++l) {
        fl[l][0] = fl[l][1] = 0;
#pragma unroll
        for (int r = 0; r < kRows; ++r) q[l][0][r] = q[l][1][r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    unsigned cnt[T + 1];
#pragma unroll
    for (int l = 0; l <= T; ++l) cnt[l] = 0;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < kPrefetch; ++i) issue_bundle(zl0 + i);
    }
    unsigned raw_flags[2][kRows];  // [parity of the plane's iteration]
#pragma unroll
    for (int r = 0; r < kRows; ++r) raw_flags[0][r] = raw_flags[1][r] = 0;
    fetch_flag_bytes(zl0, raw_flags[0]);
    fetch_flag_bytes(zl0 + 1, raw_flags[1]);
    const float eps = P.early_exit ? kEps : -1.0f;
    const int k_end = ze - 1 + T;

    // Rows of this thread that lie just outside the grid's y faces mirror the adjacent inside row (kept up to
    // date after every update), so the in-register y neighbours obey the clamp rule without any select.
    auto fix_ghost_rows = [&](float4 (&v)[kRows]) {
        if (y_edge) {
#pragma unroll
            for (int r = 1; r < kRows; ++r)
                if ((gyb + r) == 0) v[r - 1] = v[r];
#pragma unroll
            for (int r = 0; r < kRows - 1; ++r)
                if ((gyb + r) == P.ny - 1) v[r + 1] = v[r];
        }
    };

    // One marching step.  PH = (k - zl0) & 1 selects the register slots; FACE = this iteration may touch the
    // grid's lower z face (only the first T iterations of the bottom bricks), where the "F" neighbour is the
    // cell itself.
    auto step = [&](auto ph, auto face, const int k) {
        constexpr int PH = decltype(ph)::value;   // slot of the older plane / of the plane produced now
        constexpr int MID = PH ^ 1;               // slot of the centre plane
        constexpr bool FACE = decltype(face)::value;
        const int it = k - zl0;
#ifdef FXB_TIMING
#define FXB_MARK(m) if (tid == 32 && blockIdx.x == 0 && it < 16) state->active_after[32 + 6 * it + (m)] = clock64()
#else
#define FXB_MARK(m)
#endif
        FXB_MARK(0);
        if (tid == 0) issue_bundle(k + kPrefetch);
        if (k <= zl1) mbar_wait(&bars[it % (kPrefetch + 1)], (uint32_t)(it / (kPrefetch + 1)) & 1u);
        FXB_MARK(1);

        float4 head[2][kRows];  // relax_head results of the level being finished and of the next one
        bool any[T + 2];
        bool run[T + 2];

        // first five additions of level l (consumes queue l-1: older plane, centre plane, its xy neighbours)
        auto do_head = [&](auto lc) {
            constexpr int l = decltype(lc)::value;
            const int zc = k - l;
            run[l] = zc >= lev_lo(l) && zc < lev_hi(l);
            any[l] = false;
            if (!run[l]) return;
            const unsigned act = (l <= levels) ? fl[l - 1][MID] : 0u;
            any[l] = __any_sync(0xffffffffu, act != 0u);
            if (!any[l]) return;
            float4 up_s, dn_s;
            if (l == 1) {
                const float* nb = sm_p0 + ((it + S::kP0Slots - 1) % S::kP0Slots) * kPlane;  // plane k-1
                up_s = *reinterpret_cast<const float4*>(nb + off_up);
                dn_s = *reinterpret_cast<const float4*>(nb + off_dn);
            } else {
                const float* nb = sm_lev + ((l - 2) * 2 + ((it + 1) & 1)) * kEdgePlane;
                up_s = *reinterpret_cast<const float4*>(nb + eoff_up);
                dn_s = *reinterpret_cast<const float4*>(nb + eoff_dn);
            }
            const float* rb = sm_rhs + ((zc - zl0) % S::kRhsSlots) * kPlane;
            const bool lo_is_c = FACE && zc == P.z_face_lo;
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                const float4 c = q[l - 1][MID][r];
                float left = __shfl_up_sync(0xffffffffu, c.w, 1);
                float right = __shfl_down_sync(0xffffffffu, c.x, 1);
                if (clamp_l) left = c.x;
                if (clamp_r) right = c.w;
                const float4 rhs = *reinterpret_cast<const float4*>(rb + (off0 + r * kTileX));
                if (FACE) {
                    const float4 o = q[l - 1][PH][r];
                    const float4 lo = lo_is_c ? c : o;
                    head[l & 1][r] = relax_head(c, lo, r == 0 ? up_s : q[l - 1][MID][r > 0 ? r - 1 : 0],
                                                r == kRows - 1 ? dn_s : q[l - 1][MID][r < kRows - 1 ? r + 1 : r], left,
                                                right, rhs);
                } else {
                    head[l & 1][r] = relax_head(c, q[l - 1][PH][r], r == 0 ? up_s : q[l - 1][MID][r > 0 ? r - 1 : 0],
                                                r == kRows - 1 ? dn_s : q[l - 1][MID][r < kRows - 1 ? r + 1 : r], left,
                                                right, rhs);
                }
            }
        };

        // last addition + update of level l; the new plane goes into the older slot of queue l (or to global)
        auto do_tail = [&](auto lc) {
            constexpr int l = decltype(lc)::value;
            const int zc = k - l;
            if constexpr (l < T) {
                unsigned rf = 0;
                if (run[l]) {
                    if (any[l]) {
                        const unsigned act = (l <= levels) ? fl[l - 1][MID] : 0u;
#pragma unroll
                        for (int r = 0; r < kRows; ++r) {
                            unsigned st;
                            relax_tail(head[l & 1][r], q[l - 1][PH][r], q[l - 1][MID][r], (act >> (4 * r)) & 0xFu, eps,
                                       q[l][PH][r], st);
                            rf |= st << (4 * r);
                        }
                        fix_ghost_rows(q[l][PH]);
                        if (zc >= zs && zc < ze) cnt[l] += __popc(rf & own_bits);
                    } else {
#pragma unroll
                        for (int r = 0; r < kRows; ++r) q[l][PH][r] = q[l - 1][MID][r];
                    }
                    fl[l][PH] = rf;
                } else {  // outside this level's range or beyond the grid's top face: repeat the last plane
#pragma unroll
                    for (int r = 0; r < kRows; ++r) q[l][PH][r] = q[l][MID][r];
                    fl[l][PH] = fl[l][MID];
                }
                // publish the edge rows of the new plane for the warps above / below (read next iteration)
                float* dst = sm_lev + ((l - 1) * 2 + (it & 1)) * kEdgePlane;
                *reinterpret_cast<float4*>(dst + eoff_top) = q[l][PH][0];
                *reinterpret_cast<float4*>(dst + eoff_bot) = q[l][PH][kRows - 1];
            } else if (run[l] && zc >= zs && zc < ze) {  // level T: the pass's output
                float4 res[kRows];
                unsigned rf = 0;
                if (any[l]) {
                    const unsigned act = (l <= levels) ? fl[l - 1][MID] : 0u;
#pragma unroll
                    for (int r = 0; r < kRows; ++r) {
                        unsigned st;
                        relax_tail(head[l & 1][r], q[l - 1][PH][r], q[l - 1][MID][r], (act >> (4 * r)) & 0xFu, eps,
                                   res[r], st);
                        rf |= st << (4 * r);
                    }
                    cnt[l] += __popc(rf & own_bits);
                } else {
#pragma unroll
                    for (int r = 0; r < kRows; ++r) res[r] = q[l - 1][MID][r];
                }
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    if ((own_bits >> (4 * r)) & 1u) {
                        const size_t row = (size_t)zc * P.ny + (gyb + r);
                        *reinterpret_cast<float4*>(p_out + row * P.nx + gx) = res[r];
                    }
                    // bit-packed freeze flags: two quads (8 cells) per byte, written by the odd lane
                    const unsigned nib = (rf >> (4 * r)) & 0xFu;
                    const unsigned hi = __shfl_down_sync(0xffffffffu, nib, 1);
                    if (((own_bits >> (4 * r)) & 1u) && (lane & 1))
                        m_out[((size_t)zc * P.ny + (gyb + r)) * nxb + (gx >> 3)] = (unsigned char)(nib | (hi << 4));
                }
            }
        };

        do_head(std::integral_constant<int, 1>{});
        // level 0: the newest plane replaces the older slot of queue 0 (its consumer has taken it above)
        if (k < zl1) {
            const float* src = sm_p0 + (it % S::kP0Slots) * kPlane;
#pragma unroll
            for (int r = 0; r < kRows; ++r) q[0][PH][r] = *reinterpret_cast<const float4*>(src + (off0 + r * kTileX));
            fix_ghost_rows(q[0][PH]);
            fl[0][PH] = decode_flags(raw_flags[PH]);
            fetch_flag_bytes(k + 2, raw_flags[PH]);
        } else {  // beyond the grid's top face: ghost plane = last plane (clamp rule)
#pragma unroll
            for (int r = 0; r < kRows; ++r) q[0][PH][r] = q[0][MID][r];
            fl[0][PH] = fl[0][MID];
        }
        FXB_MARK(2);
        if constexpr (T >= 2) do_head(std::integral_constant<int, 2>{});
        do_tail(std::integral_constant<int, 1>{});
        FXB_MARK(3);
        if constexpr (T >= 3) do_head(std::integral_constant<int, 3>{});
        if constexpr (T >= 2) do_tail(std::integral_constant<int, 2>{});
        if constexpr (T >= 4) do_head(std::integral_constant<int, 4>{});
        if constexpr (T >= 3) do_tail(std::integral_constant<int, 3>{});
        if constexpr (T >= 4) do_tail(std::integral_constant<int, 4>{});
        FXB_MARK(4);
        __syncthreads();
        FXB_MARK(5);
    };

    {
        using I0 = std::integral_constant<int, 0>;
        using I1 = std::integral_constant<int, 1>;
        int k = zl0;
        // iterations that can touch the lower z face (plane z_face_lo is consumed at k = z_face_lo + 1 .. + T)
        const int k_face = min(k_end, P.z_face_lo + T);
        for (; k <= k_face; ++k) {
            if ((k - zl0) & 1) step(I1{}, std::true_type{}, k);
            else step(I0{}, std::true_type{}, k);
        }
        if ((k - zl0) & 1) {
            if (k <= k_end) step(I1{}, std::false_type{}, k);
            ++k;
        }
        for (; k <= k_end; k += 2) {
            step(I0{}, std::false_type{}, k);
            if (k + 1 > k_end) break;
            step(I1{}, std::false_type{}, k + 1);
        }
    }

    // ---- per-level active counts of this brick -> global counters; brick state --------------------------
#pragma unroll
    for (int l = 1; l <= T; ++l) {
        unsigned v = cnt[l];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0 && v) atomicAdd(&s_cnt[l - 1], v);
    }
    __syncthreads();
    if (brick < 0) return;
    if (tid < T && tid < levels) {
        const unsigned v = s_cnt[tid];
        if (v) atomicAdd(&state->active_after[s0 + tid], (unsigned long long)v);
    }
    if (tid == 0) {
        // still active -> relax again next pass; just frozen -> one copy into the other pressure buffer next pass
        if (s_cnt[levels - 1] != 0u) W.relax[(pass_of<DYN>(P, seq) + 1) & 1][atomicAdd(&W.relax_count[pass_of<DYN>(P, seq) + 1], 1)] = brick;
        else W.copy[(pass_of<DYN>(P, seq) + 1) & 1][atomicAdd(&W.copy_count[pass_of<DYN>(P, seq) + 1], 1)] = brick;
        atomicAdd(&state->bricks_processed, 1ull);
    }
}

// One fused pass.  P.pass is the position in the frame's relax sequence (it selects the work lists, the freeze masks
// and the side of the pressure ping-pong), s0 the number of sweeps completed before it.  Returns the sweeps applied
// (0 when the pass had nothing to do).
template <class S, bool DYN = false>
__device__ __forceinline__ int jacobi_pass_body(const CUtensorMap& map_p0, const CUtensorMap& map_p1,
                                                const CUtensorMap& map_rhs, const FrameParams* __restrict__ frame,
                                                StepState* __restrict__ state, float* p0, float* p1, unsigned char* m0,
                                                unsigned char* m1, const WorkLists& W, const PassParams& P,
                                                const int s0, const int seq = 0) {
    FXB_SHAPE_CONSTANTS(S);
    // independent loads first (one round trip instead of a chain), then the decisions
    const float dt = frame->dt;
    const int p_cur = state->p_cur;
    const unsigned long long still = pass_of<DYN>(P, seq) > 0 ? state->active_after[s0 - 1] : 1ull;
    const int n_relax = W.relax_count[pass_of<DYN>(P, seq)], n_copy = W.copy_count[pass_of<DYN>(P, seq)];
    if (!(0.0f < dt)) return 0;
    if (pass_of<DYN>(P, seq) > 0 && !P.run_all && still == 0ull) return 0;
    const int levels = min(T, P.levels_total - s0);

    const int sel = (p_cur + pass_of<DYN>(P, seq)) & 1;
    const CUtensorMap* map_in = sel ? &map_p1 : &map_p0;
    const float* p_in = sel ? p1 : p0;
    float* p_out = sel ? p0 : p1;
    const unsigned char* m_in = (pass_of<DYN>(P, seq) & 1) ? m1 : m0;
    unsigned char* m_out = (pass_of<DYN>(P, seq) & 1) ? m0 : m1;

    const int tid = threadIdx.x;
    extern __shared__ __align__(1024) float sm[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::kFloats);  // [kPrefetch + 1]
    unsigned* s_cnt = reinterpret_cast<unsigned*>(bars + 6);               // [T]
    if ((smem_u32(sm) & 127u) != 0u) __trap();

    // Work lists of this pass: every brick in pass 0; afterwards the bricks that froze in the previous pass (one
    // copy each) and the bricks that still hold an active cell (relaxed again).  CTAs are persistent
    // (kCtasPerSm per SM) and take list entries round-robin, so that the entry index — and in pass 0 the brick
    // coordinates — stay warp-uniform values.
    if (pass_of<DYN>(P, seq) > 0) {
        const int* __restrict__ copy_list = W.copy[pass_of<DYN>(P, seq) & 1];
        for (int w = blockIdx.x; w < n_copy; w += gridDim.x) copy_frozen_brick<S>(p_in, p_out, m_out, P, copy_list[w]);
        if (tid == 0 && blockIdx.x == 0 && n_copy) atomicAdd(&state->bricks_copied, (unsigned long long)n_copy);
    }
    const int n_work = pass_of<DYN>(P, seq) == 0 ? P.ntx * P.nty * P.nzc : n_relax;
    const int* __restrict__ list_in = W.relax[pass_of<DYN>(P, seq) & 1];
    bool bars_live = false;

    for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
        __syncthreads();  // the previous brick is completely finished (shared memory and barriers are idle)
        if (tid == 0) {
            if (bars_live) {
#pragma unroll
                for (int i = 0; i <= kPrefetch; ++i)
                    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars[i])) : "memory");
            }
#pragma unroll
            for (int i = 0; i <= kPrefetch; ++i) mbar_init(&bars[i], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        bars_live = true;
        if (tid < T) s_cnt[tid] = 0;
        __syncthreads();
        const int brick = pass_of<DYN>(P, seq) == 0 ? work : list_in[work];
        const int tx = brick % P.ntx, ty = (brick / P.ntx) % P.nty, zc_idx = brick / (P.ntx * P.nty);
        const int zs = P.z_out0 + zc_idx * P.bz;
        relax_brick<S, DYN>(map_in, &map_rhs, p_out, m_in, m_out, state, W, P, brick, tx, ty, zs, min(zs + P.bz, P.z_out1),
                            levels, s0, seq);
    }

    // Multi-GPU: the pressure halo is exchanged only every few passes, deep enough that in between the planes next
    // to the slab faces can be relaxed here as well (redundantly with their owner, bit-identically).  These halo
    // bricks are always processed — their cells carry the owner's freeze flags, so frozen regions cost only the copy.
    const int tiles = P.ntx * P.nty;
    const int lo_chunks = (P.ext_lo + P.bz - 1) / P.bz, hi_chunks = (P.ext_hi + P.bz - 1) / P.bz;
    for (int work = blockIdx.x; work < tiles * (lo_chunks + hi_chunks); work += gridDim.x) {
        __syncthreads();
        if (tid == 0) {
            if (bars_live) {
#pragma unroll
                for (int i = 0; i <= kPrefetch; ++i)
                    asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(&bars[i])) : "memory");
            }
#pragma unroll
            for (int i = 0; i <= kPrefetch; ++i) mbar_init(&bars[i], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        bars_live = true;
        if (tid < T) s_cnt[tid] = 0;
        __syncthreads();
        const int tile = work % tiles, chunk = work / tiles;
        int zs, ze;
        if (chunk < lo_chunks) {
            zs = P.z_out0 - P.ext_lo + chunk * P.bz;
            ze = min(zs + P.bz, P.z_out0);
        } else {
            zs = P.z_out1 + (chunk - lo_chunks) * P.bz;
            ze = min(zs + P.bz, P.z_out1 + P.ext_hi);
        }
        relax_brick<S, DYN>(map_in, &map_rhs, p_out, m_in, m_out, state, W, P, -1, tile % P.ntx, tile / P.ntx, zs, ze, levels,
                            s0, seq);
    }
    return levels;
}

// DYN = false: the static schedule (launch k is pass k of the frame).  DYN = true: the schedule is shared with the
// tail kernel (jacobi_tail.cu): the position in the frame's relax sequence and the sweeps completed so far come from
// StepState, and this launch only runs when the solve stands exactly at the sweep count its static index expects
// (otherwise a tail launch has taken over, or will).
template <class S, bool DYN>
__global__ void __launch_bounds__(S::kThreads, S::kCtasPerSm)
jacobi_pass_kernel(const __grid_constant__ CUtensorMap map_p0, const __grid_constant__ CUtensorMap map_p1,
                   const __grid_constant__ CUtensorMap map_rhs, const FrameParams* __restrict__ frame,
                   StepState* __restrict__ state, float* p0, float* p1, unsigned char* m0, unsigned char* m1,
                   const __grid_constant__ WorkLists W, const __grid_constant__ PassParams P) {
    if constexpr (!DYN) {
        jacobi_pass_body<S>(map_p0, map_p1, map_rhs, frame, state, p0, p1, m0, m1, W, P, P.pass * S::T);
    } else {
        const int seq = state->seq, s0 = state->sweeps_done;
        if (s0 != P.pass * S::T) return;
        const int levels = jacobi_pass_body<S, true>(map_p0, map_p1, map_rhs, frame, state, p0, p1, m0, m1, W, P, s0, seq);
        if (levels == 0) return;
        // the last CTA to finish advances the shared schedule (every CTA has read it long before)
        __syncthreads();
        if (threadIdx.x == 0) {
            __threadfence();
            if (atomicAdd(&state->done_ctas, 1) == (int)gridDim.x - 1) {
                state->done_ctas = 0;
                state->seq = seq + 1;
                state->sweeps_done = s0 + levels;
            }
        }
    }
}


typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

bool make_plane_map(CUtensorMap* map, float* base, int nx, int ny, int nz_alloc, int tile_y) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz_alloc};
    const cuuint64_t strides[2] = {(cuuint64_t)nx * 4, (cuuint64_t)nx * ny * 4};
    const cuuint32_t box[3] = {(cuuint32_t)kTileX, (cuuint32_t)tile_y, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class S, bool DYN = false>
cudaError_t launch_shape(const FusedJacobi& J, const Domain& d, const FrameParams* frame, StepState* state, int pass,
                         int iters, int early_exit, bool run_all, int ext_lo, int ext_hi, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(jacobi_pass_kernel<S, DYN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)S::kBytes);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    PassParams P;
    P.nx = d.nx; P.ny = d.ny; P.nz_alloc = d.nz_alloc;
    P.z_face_lo = 0 - d.z_first;
    P.z_face_hi = d.nz - d.z_first;
    P.z_out0 = d.z_own0 - d.z_first; P.z_out1 = d.z_own1 - d.z_first;
    P.bz = J.bz; P.ntx = J.ntx; P.nty = J.nty; P.nzc = J.nzc;
    P.pass = pass; P.levels_total = iters; P.early_exit = early_exit; P.run_all = run_all ? 1 : 0;
    P.ext_lo = ext_lo; P.ext_hi = ext_hi;
    const int nbricks = J.ntx * J.nty * J.nzc;
    const int slots = J.num_sms * S::kCtasPerSm;  // persistent CTAs
    const int grid = nbricks < slots ? nbricks : slots;
    WorkLists W;
    const int np = FusedJacobi::kMaxPasses + 1;
    W.relax[0] = J.work_list[0]; W.relax[1] = J.work_list[1];
    W.copy[0] = J.work_list[0] + nbricks; W.copy[1] = J.work_list[1] + nbricks;
    W.relax_count = J.work_count; W.copy_count = J.work_count + np;
    jacobi_pass_kernel<S, DYN><<<grid, S::kThreads, S::kBytes, stream>>>(
        *reinterpret_cast<const CUtensorMap*>(J.map_p[0]), *reinterpret_cast<const CUtensorMap*>(J.map_p[1]),
        *reinterpret_cast<const CUtensorMap*>(J.map_rhs), frame, state, J.p[0], J.p[1], J.mask[0], J.mask[1], W, P);
    return cudaGetLastError();
}

// Kernel shapes (tile rows = ROWS * WARPS).  Measured on B200 (profiles/): the small two-CTA-per-SM shape wins
// because the marching loop is bound by dependent-issue latency, not by HBM; FXB_VARIANT selects the others.
//   0: 2 rows/thread, 8 warps (tile 128 x 16), TMA depth 2, two CTAs per SM   (T <= 2; the default)
//   1: 2 rows/thread, 16 warps (tile 128 x 32), TMA depth 1
//   2: 4 rows/thread,  8 warps (tile 128 x 32), TMA depth 2
template <int T>
cudaError_t launch_T(const FusedJacobi& J, const Domain& d, const FrameParams* frame, StepState* state, int pass,
                     int iters, int early_exit, bool run_all, int ext_lo, int ext_hi, cudaStream_t stream) {
    switch (J.variant) {
        case 0:
            if constexpr (T <= 2) {
                if (J.dynamic)  // schedule shared with the tail kernel (jacobi_tail.cu)
                    return launch_shape<Shape<T, 2, 8, 2, 2>, true>(J, d, frame, state, pass, iters, early_exit, run_all, ext_lo, ext_hi, stream);
                return launch_shape<Shape<T, 2, 8, 2, 2>>(J, d, frame, state, pass, iters, early_exit, run_all, ext_lo, ext_hi, stream);
            }
            break;
        case 1: return launch_shape<Shape<T, 2, 16, 1>>(J, d, frame, state, pass, iters, early_exit, run_all, ext_lo, ext_hi, stream);
        case 2: return launch_shape<Shape<T, 4, 8, 2>>(J, d, frame, state, pass, iters, early_exit, run_all, ext_lo, ext_hi, stream);
    }
    return cudaErrorInvalidValue;
}

int variant_tile_y(int variant) { return variant == 0 ? 16 : 32; }

}  // namespace

bool fused_jacobi_supported(const Domain& d) { return d.nz > 1 && (d.nx % 8) == 0 && d.nx >= 8; }

int fused_jacobi_plan(FusedJacobi* J, const Domain& d, int fuse_t, float* p0, float* p1, float* rhs) {
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    J->T = fuse_t;
    J->variant = fuse_t <= 2 ? 0 : 2;
    if (const char* e = getenv("FXB_VARIANT")) {
        const int v = atoi(e);
        if (v >= 0 && v <= 2 && (v != 0 || fuse_t <= 2)) J->variant = v;
    }
    J->tile_y = variant_tile_y(J->variant);
    const int out_y = J->tile_y - 2 * fuse_t;
    J->ntx = (d.nx + kOutX - 1) / kOutX;
    J->nty = (d.ny + out_y - 1) / out_y;
    const int nz_out = d.z_own1 - d.z_own0;
    J->bz = nz_out >= 8 ? 8 : nz_out;
    if (const char* e = getenv("FXB_BZ")) {  // tuning knob: planes per brick
        const int v = atoi(e);
        if (v >= 1 && v <= nz_out) J->bz = v;
    }
    J->nzc = (nz_out + J->bz - 1) / J->bz;
    J->p[0] = p0; J->p[1] = p1; J->rhs = rhs;
    if (!make_plane_map(reinterpret_cast<CUtensorMap*>(J->map_p[0]), p0, d.nx, d.ny, d.nz_alloc, J->tile_y)) return -1;
    if (!make_plane_map(reinterpret_cast<CUtensorMap*>(J->map_p[1]), p1, d.nx, d.ny, d.nz_alloc, J->tile_y)) return -1;
    if (!make_plane_map(reinterpret_cast<CUtensorMap*>(J->map_rhs), rhs, d.nx, d.ny, d.nz_alloc, J->tile_y)) return -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&J->num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    return 0;
}

// Tensor map of a pressure buffer with an arbitrary box (the tail kernel's TMA-staged window, jacobi_tail.cu).
bool fused_make_box_map(void* map128, float* base, int nx, int ny, int nz_alloc, int box_x, int box_y, int box_z) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz_alloc};
    const cuuint64_t strides[2] = {(cuuint64_t)nx * 4, (cuuint64_t)nx * ny * 4};
    const cuuint32_t box[3] = {(cuuint32_t)box_x, (cuuint32_t)box_y, (cuuint32_t)box_z};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(reinterpret_cast<CUtensorMap*>(map128), CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;  // out-of-bounds elements arrive as zeros
}

size_t fused_jacobi_bricks(const FusedJacobi& J) { return (size_t)J.ntx * J.nty * J.nzc; }

size_t fused_jacobi_brick_cells(const FusedJacobi& J) { return (size_t)kOutX * (J.tile_y - 2 * J.T) * J.bz; }

void fused_jacobi_brick_extent(const FusedJacobi& J, int out[3]) {
    out[0] = kOutX; out[1] = J.tile_y - 2 * J.T; out[2] = J.bz;
}

cudaError_t launch_jacobi_pass_fused(const FusedJacobi& J, const Domain& d, const FrameParams* frame, StepState* state,
                                     int pass, int iters, int early_exit, bool run_all, int ext_lo, int ext_hi,
                                     cudaStream_t stream) {
    switch (J.T) {
        case 1: return launch_T<1>(J, d, frame, state, pass, iters, early_exit, run_all, ext_lo, ext_hi, stream);
        case 2: return launch_T<2>(J, d, frame, state, pass, iters, early_exit, run_all, ext_lo, ext_hi, stream);
        case 3: return launch_T<3>(J, d, frame, state, pass, iters, early_exit, run_all, ext_lo, ext_hi, stream);
        case 4: return launch_T<4>(J, d, frame, state, pass, iters, early_exit, run_all, ext_lo, ext_hi, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace fxb
