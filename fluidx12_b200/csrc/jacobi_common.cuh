// jacobi_common.cuh — what the two fused Jacobi kernels share (jacobi_fused.cu: z-marching passes; jacobi_resident.cu:
// brick-resident passes): tile shapes, TMA / mbarrier wrappers, the pass parameters, work lists and work items, the
// one-time copy of a frozen brick and the relaxation of one quad in the reference's operation order
// (CSPoisson.hlsli:8-26 as restated in SURVEY.md App. A.3).
#pragma once

#include <cuda.h>

#include "common.cuh"
#include "kernels.h"

namespace fxb {

namespace {

constexpr float kInv6 = 0.166666672f;
constexpr float kEps = 0.00100000005f;
constexpr int kHaloX = 4;  // one quad

// Compile-time shape of one kernel variant.
//   T     sweeps fused per pass;  LX lanes per tile row (32: tile 128 wide, 16: tile 64 wide; a warp then covers
//   32/LX row groups);  WARPS warps per CTA;  DEPTH TMA bundles in flight ahead of the one being consumed;  ROWS rows
//   per thread (2: fewer instructions per cell, for the passes that relax many bricks; 1: twice the warps on a tile,
//   half the dependent work per warp, for the passes in which one brick chain per SM sets the pace).
template <int T_, int LX_, int WARPS_, int DEPTH_, int CTAS_, int ROWS_ = 2>
struct Shape {
    static constexpr int T = T_, LX = LX_, kWarps = WARPS_, kDepth = DEPTH_, kCtasPerSm = CTAS_, kRows = ROWS_;
    static constexpr int kSub = 32 / LX_;               // row groups per warp
    static constexpr int kWarpRows = ROWS_ * kSub;
    static constexpr int kThreads = 32 * WARPS_;
    static constexpr int kTileX = 4 * LX_, kTileY = WARPS_ * kWarpRows;
    static constexpr int kOutX = kTileX - 2 * kHaloX, kOutY = kTileY - 2 * T_;
    static constexpr int kPlane = kTileX * kTileY;       // floats per staged plane
    static constexpr int kPSlots = DEPTH_ + 2;           // planes k-1 (y neighbours of level 1), k, in flight
    static constexpr int kRSlots = T_ + DEPTH_ + 1;      // rhs planes k-T .. k-1 in use, k, in flight
    static constexpr int kBars = DEPTH_ + 1;
    static constexpr int kPubPlanes = 2 * (T_ - 1);      // levels 1..T-1, double-buffered by iteration parity
    static constexpr size_t kFloats = (size_t)(kPSlots + kRSlots + kPubPlanes) * kPlane;
    static constexpr size_t kBytes = kFloats * sizeof(float) + 64;  // + barriers
    static_assert(LX_ == 32 || LX_ == 16, "tile rows are 32 or 16 lanes wide");
    static_assert(ROWS_ == 1 || ROWS_ == 2, "rows per thread");
    static_assert(kBars <= 8, "barrier slots");
    static_assert((kBytes + 1024) * CTAS_ <= 233472, "shared memory budget (228 KB per SM, 1 KB reserved per CTA)");
    static_assert(kOutY > 0 && kOutX % 8 == 0, "own region must be a whole number of mask bytes wide");
};

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
    int pitch;             // floats per row of the pressure / rhs arrays (multiple of 8); mask rows are pitch/8 bytes
    int nz_alloc;          // local planes allocated
    int z_face_lo;         // local index of global plane 0 (negative when it is on another rank)
    int z_face_hi;         // local index one past global plane nz-1
    int z_out0, z_out1;    // local planes this rank must produce
    int bz;                // planes per brick
    int ntx, nty, nzc;     // brick grid
    int pass;              // index of this fused pass in the frame
    int s0;                // sweeps completed before this pass
    int levels_total;      // ITER
    int early_exit;
    int run_all;           // multi-GPU: never end the solve on this rank's own freeze counters
    int ext_lo, ext_hi;    // multi-GPU: planes below / above the owned range to relax redundantly in this pass
    int copy_all;          // 1: every brick that froze in the first pass is copied; 0: only those next to an active brick
    int keep_lo, keep_hi;  // multi-GPU: bricks of the lowest / highest layer are always copied (a neighbour rank reads them)
    int first_brick, first_count;  // first pass: this launch relaxes the bricks (first_brick + w) mod bricks, w < first_count
    int event;             // fused halos: this kernel's number m in the frame (common.cuh PeerView)
    int push_depth;        // fused halos: own planes next to an interior face that are also stored into the neighbour
};

struct WorkLists {
    int* relax[2];     // bricks that still hold an active cell, ping-pong by pass parity
    int* copy[2];      // bricks that froze in the previous pass: one copy into the other pressure buffer
    int* relax_count;  // [pass]
    int* copy_count;   // [pass]
    int* brick_flag;   // [bricks] first pass: bit 0 = all cells froze, bit 1 = a brick next to it is still active
    int* brick_half;   // [bricks] passes on half bricks: where the two halves of a brick meet (zero between passes)
};

// Fused halos: the neighbours' copies of the pressure and freeze-mask ping-pong buffers ([side][buffer], side 0 =
// rank - 1, side 1 = rank + 1; nullptr at a grid face / on a single GPU).
struct JacobiPeers {
    float* p[2][2];
    unsigned char* m[2][2];
};

// One work item of a pass: a brick of this rank's own planes (brick >= 0; tracked in the work lists and the freeze
// counters) or planes of the z-halo relaxed redundantly between two exchanges (multi-GPU; brick < 0, never listed).
struct Item {
    int brick, gx0, gy0, zs, ze;
};

template <class S>
__device__ __forceinline__ Item own_item(const PassParams& P, const int brick) {
    Item it;
    const int tx = brick % P.ntx, ty = (brick / P.ntx) % P.nty, zc = brick / (P.ntx * P.nty);
    it.brick = brick;
    it.gx0 = tx * S::kOutX - kHaloX;
    it.gy0 = ty * S::kOutY - S::T;
    it.zs = P.z_out0 + zc * P.bz;
    it.ze = min(it.zs + P.bz, P.z_out1);
    return it;
}

template <class S>
__device__ __forceinline__ Item ext_item(const PassParams& P, const int e) {
    const int tiles = P.ntx * P.nty, lo_chunks = (P.ext_lo + P.bz - 1) / P.bz;
    const int tile = e % tiles, chunk = e / tiles;
    Item it;
    it.brick = -1;
    it.gx0 = (tile % P.ntx) * S::kOutX - kHaloX;
    it.gy0 = (tile / P.ntx) * S::kOutY - S::T;
    if (chunk < lo_chunks) {
        it.zs = P.z_out0 - P.ext_lo + chunk * P.bz;
        it.ze = min(it.zs + P.bz, P.z_out0);
    } else {
        it.zs = P.z_out1 + (chunk - lo_chunks) * P.bz;
        it.ze = min(it.zs + P.bz, P.z_out1 + P.ext_hi);
    }
    return it;
}

// Where the stores of plane z of a pass's output also go (fused halos): side 0 / 1 when the plane is within
// push_depth of the interior face below / above.
__device__ __forceinline__ bool pushes_lo(const PeerView& pv, const PassParams& P, int z) {
    return pv.has_lo && z < P.z_out0 + P.push_depth;
}
__device__ __forceinline__ bool pushes_hi(const PeerView& pv, const PassParams& P, int z) {
    return pv.has_hi && z >= P.z_out1 - P.push_depth;
}

// A brick whose cells all froze during pass p-1 holds its final values in that pass's output buffer only.  Pass p
// copies its own region once into the other buffer (and clears the other mask buffer), after which the brick is
// final in both ping-pong buffers and is never touched again in this frame.  Pure streaming (8 B/cell), eight
// independent 16-byte loads in flight per thread.  Planes next to an interior slab face go to the neighbour as well.
template <class S, bool FUSED>
__device__ __noinline__ void copy_frozen_brick(const float* __restrict__ p_in, float* __restrict__ p_out,
                                               unsigned char* __restrict__ m_out, const PassParams& P, const int brick,
                                               const PeerView& pv, const JacobiPeers& peers, const int pi,
                                               const int mi) {  // pi / mi: which of the neighbours' p / mask buffers
    const int tid = threadIdx.x, nxb = P.pitch >> 3;
    const Item it = own_item<S>(P, brick);
    const int x_lo = it.gx0 + kHaloX, y_lo = it.gy0 + S::T;
    const int rows = min(S::kOutY, P.ny - y_lo), planes = it.ze - it.zs;
    const int qpr = (min(S::kOutX, P.nx - x_lo) + 3) >> 2;  // float4 per row (the last one may reach into the row padding)
    const int total = planes * rows * qpr;
    const long long plane_f = (long long)P.ny * P.pitch, plane_b = (long long)P.ny * nxb;
    for (int base = tid; base < total; base += 8 * S::kThreads) {
        float4 v[8];
        size_t at[8];
        int zz[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = base + u * S::kThreads;
            const int xq = i % qpr, rz = i / qpr;
            zz[u] = it.zs + rz / rows;
            at[u] = ((size_t)zz[u] * P.ny + (y_lo + rz % rows)) * P.pitch + x_lo + 4 * xq;
            if (i < total) v[u] = __ldcg(reinterpret_cast<const float4*>(p_in + at[u]));  // L2: written by another CTA, possibly in this launch
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (base + u * S::kThreads < total) {
                *reinterpret_cast<float4*>(p_out + at[u]) = v[u];
                if constexpr (!FUSED) continue;
                if (pushes_lo(pv, P, zz[u])) *reinterpret_cast<float4*>(peers.p[0][pi] + (long long)at[u] + pv.dz_lo * plane_f) = v[u];
                if (pushes_hi(pv, P, zz[u])) *reinterpret_cast<float4*>(peers.p[1][pi] + (long long)at[u] + pv.dz_hi * plane_f) = v[u];
            }
    }
    const int bpr = (qpr + 1) >> 1;  // mask bytes per row
    for (int i = tid; i < planes * rows * bpr; i += S::kThreads) {
        const int xb = i % bpr, rz = i / bpr, z = it.zs + rz / rows;
        const size_t at = ((size_t)z * P.ny + (y_lo + rz % rows)) * nxb + (x_lo >> 3) + xb;
        m_out[at] = 0;
        if constexpr (!FUSED) continue;
        if (pushes_lo(pv, P, z)) peers.m[0][mi][(long long)at + pv.dz_lo * plane_b] = 0;
        if (pushes_hi(pv, P, z)) peers.m[1][mi][(long long)at + pv.dz_hi * plane_b] = 0;
    }
}

// One relaxation of a quad in the DXBC's order:
//   acc = p[L] + rhs; acc = p[R] + acc; acc = p[U] + acc; acc = p[D] + acc; acc = p[F] + acc; acc = p[B] + acc;
//   x = acc * (1/6); frozen after this sweep iff |fma(acc, 1/6, -x0)| < eps (eps < 0: never); frozen cells keep x0.
// The x neighbours live in other registers of the same quad, so those two additions are scalar; the rest are packed
// FADD2 / FMUL2 / FFMA2 on the (x,y) / (z,w) halves.  `act` / return: 4 flag bits of the quad before / after.
__device__ __forceinline__ unsigned relax_quad(const float4 c, const float4 lo, const float4 hi, const float4 up,
                                               const float4 dn, const float left, const float right, const float4 rhs,
                                               const unsigned act, const float eps, float4& out) {
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
    a = add2(make_float2(hi.x, hi.y), a);
    b = add2(make_float2(hi.z, hi.w), b);
    const float2 inv2 = make_float2(kInv6, kInv6);
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
    return s;
}

// ---- host side: the kernel arguments of one pass, shared by the launchers of both kernels ----
inline PassParams make_pass_params(const FusedJacobi& J, const Domain& d, int pass, int iters, int early_exit, bool run_all,
                                   int ext_lo, int ext_hi) {
    PassParams P;
    P.nx = d.nx; P.ny = d.ny; P.pitch = d.pitch; P.nz_alloc = d.nz_alloc;
    P.z_face_lo = 0 - d.z_first;
    P.z_face_hi = d.nz - d.z_first;
    P.z_out0 = d.z_own0 - d.z_first; P.z_out1 = d.z_own1 - d.z_first;
    P.bz = J.bz; P.ntx = J.ntx; P.nty = J.nty; P.nzc = J.nzc;
    P.pass = pass; P.s0 = fused_jacobi_s0(J, pass); P.levels_total = iters; P.early_exit = early_exit; P.run_all = run_all ? 1 : 0;
    P.ext_lo = ext_lo; P.ext_hi = ext_hi;
    P.copy_all = J.copy_all ? 1 : 0;
    P.keep_lo = d.z_own0 > 0 ? 1 : 0;
    P.keep_hi = d.z_own1 < d.nz ? 1 : 0;
    P.first_brick = 0;
    P.first_count = J.ntx * J.nty * J.nzc;
    P.event = 2 + pass;
    P.push_depth = J.push_depth;
    return P;
}

inline JacobiPeers make_jacobi_peers(const FusedJacobi& J) {
    JacobiPeers peers;
    for (int side = 0; side < 2; ++side)
        for (int i = 0; i < 2; ++i) {
            peers.p[side][i] = J.peer_p[side][i];
            peers.m[side][i] = J.peer_m[side][i];
        }
    return peers;
}

inline WorkLists make_work_lists(const FusedJacobi& J) {
    WorkLists W;
    const int np = FusedJacobi::kMaxPasses + 1;
    W.relax[0] = J.work_list[0]; W.relax[1] = J.work_list[1];
    W.copy[0] = J.work_list[0] + J.list_stride; W.copy[1] = J.work_list[1] + J.list_stride;
    W.relax_count = J.work_count; W.copy_count = J.work_count + np;
    W.brick_flag = J.brick_flag;
    W.brick_half = J.brick_flag + (size_t)J.ntx * J.nty * J.nzc;
    return W;
}

}  // namespace

}  // namespace fxb
