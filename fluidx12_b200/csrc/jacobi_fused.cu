// jacobi_fused.cu — T Jacobi sweeps of the pressure solve fused into one HBM pass (sm_100a).
//
// Replaces the relaxation loop of FluidX12/Content/Shaders/CSPoisson.hlsli:8-26 as called from
// CSProject3D.hlsl:93 (dispatch Fluid.cpp:394-408), under the deterministic restatement of SURVEY.md
// App. A.3: synchronous Jacobi, per-cell freeze once |x - x0| < 0.001, at most ITER sweeps.
//
// Scheme: z-marching 2.5-D blocking with temporal fusion.  A CTA owns a brick of 120 x (32-2T) x BZ
// output cells.  It streams the xy tile (128 x 32 cells, halo 4 in x / T in y) plane by plane along z;
// level l (= number of sweeps applied) of plane k-l is produced in iteration k, so T sweeps advance in
// lock-step, each one plane behind the previous:
//   * level-0 pressure planes and the right-hand-side planes are staged into shared memory by TMA
//     (cp.async.bulk.tensor.3d + mbarrier), one iteration ahead; out-of-grid tile parts are zero-filled;
//   * every thread keeps its own column (4 rows x 4 cells) of the two most recent planes of every level
//     in registers (the z queue), so the z neighbours never touch memory;
//   * x neighbours come from warp shuffles (a warp spans the 128-cell tile row), y neighbours from the
//     thread's own rows or from the level's plane in shared memory;
//   * the reference's clamp-to-edge neighbour rule (CSProject3D.hlsl:76-83) is applied by index (x, y)
//     or by reusing the centre value (z), never by TMA fill;
//   * the per-cell freeze flags travel with the values (4 bits per quad per level); a warp whose 512
//     cells are all frozen at a level skips that level's arithmetic; flags persist between passes in a
//     bit-packed array (1 bit per cell, 0.25 B/voxel/pass of traffic);
//   * a brick whose cells are all frozen is copied once to the other pressure buffer and skipped for the
//     rest of the frame (its value is final in both ping-pong buffers).
// Algorithmic traffic per processed cell per pass: p in 4 + rhs in 4 + p out 4 (+ 2/8 mask) bytes.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"

namespace fxb {

namespace {

constexpr int kLanes = 32;
constexpr int kRows = 4;                 // rows per thread
constexpr int kWarps = 8;                // warps per CTA
constexpr int kThreads = kLanes * kWarps;
constexpr int kTileX = 4 * kLanes;       // 128 cells per tile row
constexpr int kTileY = kRows * kWarps;   // 32 rows per tile
constexpr int kHaloX = 4;                // one quad
constexpr int kOutX = kTileX - 2 * kHaloX;  // 120
constexpr int kPlane = kTileX * kTileY;  // floats per staged plane (16 KB)
constexpr float kInv6 = 0.166666672f;
constexpr float kEps = 0.00100000005f;

template <int T>
struct Smem {
    static constexpr int kP0Slots = 3;
    static constexpr int kRhsSlots = T + 1;
    static constexpr int kPlanes = kP0Slots + 2 * (T - 1) + kRhsSlots;
    static constexpr size_t kBytes = (size_t)kPlanes * kPlane * sizeof(float) + 64;  // + barriers and counters
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
    int nz_alloc;          // local planes allocated
    int z_face_lo;         // local index of global plane 0 (or very negative when it is on another rank)
    int z_face_hi;         // local index one past global plane nz-1 (or very large)
    int z_out0, z_out1;    // local planes this rank must produce
    int bz;                // planes per brick
    int ntx, nty, nzc;     // brick grid
    int pass;              // index of this fused pass in the frame
    int levels_total;      // ITER
    int early_exit;
};

// One relaxation of a quad (4 x-adjacent cells) as two packed pairs.  Operation order = the DXBC's (SURVEY.md
// App. A.3): acc = p[L] + rhs; += p[R]; += p[U]; += p[D]; += p[F]; += p[B]; x = acc * (1/6);
// freeze when |fma(acc, 1/6, -x0)| < eps (eps < 0 disables the test: early_exit = 0).
__device__ __forceinline__ void relax_quad(const float4 c, const float4 lo, const float4 hi, const float4 up,
                                           const float4 dn, const float left, const float right, const float4 rhs,
                                           const unsigned act, const float eps, float4& out, unsigned& still) {
    const float2 inv2 = make_float2(kInv6, kInv6);
    const float2 cA = make_float2(c.x, c.y), cB = make_float2(c.z, c.w);
    const float2 mid = make_float2(c.y, c.z);  // right neighbours of pair A = left neighbours of pair B
    float2 accA = add2(make_float2(left, c.x), make_float2(rhs.x, rhs.y));
    float2 accB = add2(mid, make_float2(rhs.z, rhs.w));
    accA = add2(mid, accA);
    accB = add2(make_float2(c.w, right), accB);
    accA = add2(make_float2(up.x, up.y), accA);
    accB = add2(make_float2(up.z, up.w), accB);
    accA = add2(make_float2(dn.x, dn.y), accA);
    accB = add2(make_float2(dn.z, dn.w), accB);
    accA = add2(make_float2(lo.x, lo.y), accA);
    accB = add2(make_float2(lo.z, lo.w), accB);
    accA = add2(make_float2(hi.x, hi.y), accA);
    accB = add2(make_float2(hi.z, hi.w), accB);
    const float2 nA = mul2(accA, inv2), nB = mul2(accB, inv2);
    const float2 dA = fma2(accA, inv2, make_float2(-cA.x, -cA.y)), dB = fma2(accB, inv2, make_float2(-cB.x, -cB.y));
    unsigned s = act;
    if (fabsf(dA.x) < eps) s &= ~1u;
    if (fabsf(dA.y) < eps) s &= ~2u;
    if (fabsf(dB.x) < eps) s &= ~4u;
    if (fabsf(dB.y) < eps) s &= ~8u;
    out.x = (act & 1u) ? nA.x : c.x;
    out.y = (act & 2u) ? nA.y : c.y;
    out.z = (act & 4u) ? nB.x : c.z;
    out.w = (act & 8u) ? nB.y : c.w;
    still = s;
}

template <int T>
__global__ void __launch_bounds__(kThreads, 1)
jacobi_pass_kernel(const __grid_constant__ CUtensorMap map_p0, const __grid_constant__ CUtensorMap map_p1,
                   const __grid_constant__ CUtensorMap map_rhs, const FrameParams* __restrict__ frame,
                   StepState* __restrict__ state, float* p0, float* p1, unsigned char* m0, unsigned char* m1,
                   int* __restrict__ brick_state, const PassParams P) {
    if (!(0.0f < frame->dt)) return;
    const int s0 = P.pass * T;  // sweeps completed before this pass
    if (P.pass > 0 && state->active_after[s0 - 1] == 0ull) return;
    const int levels = min(T, P.levels_total - s0);

    const int sel = (state->p_cur + P.pass) & 1;
    const CUtensorMap* map_in = sel ? &map_p1 : &map_p0;
    const float* __restrict__ p_in = sel ? p1 : p0;
    float* __restrict__ p_out = sel ? p0 : p1;
    const unsigned char* __restrict__ m_in = (P.pass & 1) ? m1 : m0;
    unsigned char* __restrict__ m_out = (P.pass & 1) ? m0 : m1;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = blockIdx.x, ty = blockIdx.y, zc_idx = blockIdx.z;
    const int brick = (zc_idx * P.nty + ty) * P.ntx + tx;
    constexpr int kOutY = kTileY - 2 * T;

    const int gx0 = tx * kOutX - kHaloX;
    const int gy0 = ty * kOutY - T;
    const int zs = P.z_out0 + zc_idx * P.bz;
    const int ze = min(zs + P.bz, P.z_out1);
    const int nxb = P.nx >> 3;  // mask bytes per row

    const int gx = gx0 + 4 * lane;
    const bool qin = gx >= 0 && gx < P.nx;
    const bool own_lane = lane >= 1 && lane <= 30 && qin;
    int gy[kRows];
    unsigned own_bits = 0;  // bit (4r + j): cell j of row r belongs to this brick's output region
    unsigned dom_bits = 0;  // bit (4r + j): cell lies inside the grid
#pragma unroll
    for (int r = 0; r < kRows; ++r) {
        const int ry = kRows * warp + r;
        gy[r] = gy0 + ry;
        const bool rin = gy[r] >= 0 && gy[r] < P.ny;
        if (rin && qin) dom_bits |= 0xFu << (4 * r);
        if (rin && own_lane && ry >= T && ry < kTileY - T) own_bits |= 0xFu << (4 * r);
    }

    // ---- frozen bricks ---------------------------------------------------------------------------------
    if (P.pass > 0) {
        const int bs = brick_state[brick];
        if (bs == 2) return;  // final in both pressure buffers
        if (bs == 1) {        // became fully frozen in the previous pass: copy once, clear the other mask buffer
            for (int z = zs; z < ze; ++z) {
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    if ((own_bits >> (4 * r)) & 1u) {
                        const size_t row = ((size_t)z * P.ny + gy[r]);
                        *reinterpret_cast<float4*>(p_out + row * P.nx + gx) =
                            *reinterpret_cast<const float4*>(p_in + row * P.nx + gx);
                        if (lane & 1) m_out[row * nxb + (gx >> 3)] = 0;
                    }
                }
            }
            if (tid == 0) {
                brick_state[brick] = 2;
                atomicAdd(&state->bricks_copied, 1ull);
            }
            return;
        }
    }

    extern __shared__ __align__(1024) float sm[];                  // TMA destinations need 128-byte alignment
    float* sm_p0 = sm;                                             // [3][kPlane]
    float* sm_lev = sm_p0 + Smem<T>::kP0Slots * kPlane;            // [T-1][2][kPlane]  (levels 1..T-1)
    float* sm_rhs = sm_lev + 2 * (T - 1) * kPlane;                 // [T+1][kPlane]
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm_rhs + Smem<T>::kRhsSlots * kPlane);  // [3]
    unsigned* s_cnt = reinterpret_cast<unsigned*>(bars + 4);                                // [T]
    if ((smem_u32(sm) & 127u) != 0u) __trap();

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < 3; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < T) s_cnt[tid] = 0;
    __syncthreads();

    // planes available for loading and the z range each level must cover (trapezoid in z)
    const int zl0 = max(zs - T, 0), zl1 = min(ze + T, P.nz_alloc);
    int lev_lo[T + 1], lev_hi[T + 1];
#pragma unroll
    for (int l = 1; l <= T; ++l) {
        lev_lo[l] = max(zs - (T - l), 0);
        lev_hi[l] = min(ze + (T - l), P.nz_alloc);
    }

    // per-thread shared-memory offsets (floats) inside a plane
    int off[kRows];
#pragma unroll
    for (int r = 0; r < kRows; ++r) off[r] = (kRows * warp + r) * kTileX + 4 * lane;
    const int off_up = (warp == 0 || gy[0] <= 0) ? off[0] : off[0] - kTileX;
    const int off_dn = (warp == kWarps - 1 || gy[kRows - 1] >= P.ny - 1) ? off[kRows - 1] : off[kRows - 1] + kTileX;
    const bool clamp_l = lane == 0 || gx == 0;
    const bool clamp_r = lane == 31 || gx + 4 == P.nx;
    // the grid's y faces may cut through this warp's rows: then the in-register y neighbours need clamping
    bool y_edge = false;
#pragma unroll
    for (int r = 0; r < kRows; ++r) y_edge |= (gy[r] == 0 && r > 0) || (gy[r] == P.ny - 1 && r < kRows - 1);

    auto issue_bundle = [&](int k) {  // p plane k and rhs plane k-1 -> shared memory (thread 0 only)
        const bool has_p = k >= zl0 && k < zl1, has_r = k - 1 >= zl0 && k - 1 < zl1;
        if (!has_p && !has_r) return;
        uint64_t* bar = &bars[(k - zl0) % 3];
        mbar_expect_tx(bar, (uint32_t)((has_p ? 1 : 0) + (has_r ? 1 : 0)) * kPlane * 4u);
        if (has_p) tma_load_3d(sm_p0 + ((k - zl0) % 3) * kPlane, map_in, gx0, gy0, k, bar);
        if (has_r) tma_load_3d(sm_rhs + ((k - 1 - zl0) % (T + 1)) * kPlane, &map_rhs, gx0, gy0, k - 1, bar);
    };
    auto load_flags = [&](int z) -> unsigned {  // freeze flags of plane z for my 4 rows (level-0 input)
        if (P.pass == 0) return dom_bits;
        unsigned f = 0;
        if (z < zl1) {
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                if ((dom_bits >> (4 * r)) & 1u) {
                    const unsigned b = m_in[((size_t)z * P.ny + gy[r]) * nxb + (gx >> 3)];
                    f |= ((b >> (gx & 4)) & 0xFu) << (4 * r);
                }
            }
        }
        return f;
    };

    float4 q_lo[T][kRows], q_mid[T][kRows];  // z queue: levels 0..T-1, planes (zc-1, zc) of the consumer level
    unsigned f_mid[T];
#pragma unroll
    for (int l = 0; l < T; ++l) {
        f_mid[l] = 0;
#pragma unroll
        for (int r = 0; r < kRows; ++r) q_lo[l][r] = q_mid[l][r] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    unsigned cnt[T + 1];
#pragma unroll
    for (int l = 0; l <= T; ++l) cnt[l] = 0;

    if (tid == 0) issue_bundle(zl0);
    unsigned next_flags = load_flags(zl0);
    const float eps = P.early_exit ? kEps : -1.0f;
    const int k_end = ze - 1 + T;

    for (int k = zl0; k <= k_end; ++k) {
        const int it = k - zl0;
        if (tid == 0) issue_bundle(k + 1);
        if (k <= zl1) mbar_wait(&bars[it % 3], (uint32_t)(it / 3) & 1u);  // bundle k exists for k in [zl0, zl1]

        float4 nw[kRows];  // newest plane of the previous level (the "B" neighbour of the consumer)
        unsigned nf;
        if (k < zl1) {
            const float* src = sm_p0 + (it % 3) * kPlane;
#pragma unroll
            for (int r = 0; r < kRows; ++r) nw[r] = *reinterpret_cast<const float4*>(src + off[r]);
            nf = next_flags;
            next_flags = load_flags(k + 1);
        } else {  // beyond the grid's top face: ghost plane = last plane (clamp rule)
#pragma unroll
            for (int r = 0; r < kRows; ++r) nw[r] = q_mid[0][r];
            nf = f_mid[0];
        }

#pragma unroll
        for (int l = 1; l <= T; ++l) {
            const int zc = k - l;  // plane produced at level l in this iteration
            float4 res[kRows];
            unsigned rf;
            const bool run = zc >= lev_lo[l] && zc < lev_hi[l];
            if (run) {
                const unsigned act = (l <= levels) ? f_mid[l - 1] : 0u;
                if (__any_sync(0xffffffffu, act != 0u)) {
                    const float* nb = (l == 1) ? sm_p0 + ((it + 2) % 3) * kPlane  // plane k-1
                                               : sm_lev + ((l - 2) * 2 + ((it + 1) & 1)) * kPlane;
                    const float* rb = sm_rhs + ((zc - zl0) % (T + 1)) * kPlane;
                    const float4 up_s = *reinterpret_cast<const float4*>(nb + off_up);
                    const float4 dn_s = *reinterpret_cast<const float4*>(nb + off_dn);
                    const bool lo_clamp = zc == P.z_face_lo;
                    rf = 0;
#pragma unroll
                    for (int r = 0; r < kRows; ++r) {
                        const float4 c = q_mid[l - 1][r];
                        const float4 lo = lo_clamp ? c : q_lo[l - 1][r];
                        float4 up = (r == 0) ? up_s : q_mid[l - 1][r > 0 ? r - 1 : 0];
                        float4 dn = (r == kRows - 1) ? dn_s : q_mid[l - 1][r < kRows - 1 ? r + 1 : r];
                        if (y_edge) {
                            if (gy[r] == 0) up = c;
                            if (gy[r] == P.ny - 1) dn = c;
                        }
                        float left = __shfl_up_sync(0xffffffffu, c.w, 1);
                        float right = __shfl_down_sync(0xffffffffu, c.x, 1);
                        if (clamp_l) left = c.x;
                        if (clamp_r) right = c.w;
                        const float4 rhs = *reinterpret_cast<const float4*>(rb + off[r]);
                        unsigned st;
                        relax_quad(c, lo, nw[r], up, dn, left, right, rhs, (act >> (4 * r)) & 0xFu, eps, res[r], st);
                        rf |= st << (4 * r);
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < kRows; ++r) res[r] = q_mid[l - 1][r];
                    rf = 0;
                }
                if (zc >= zs && zc < ze) cnt[l] += __popc(rf & own_bits);
            } else if (l < T) {  // outside this level's range (or beyond the top face): repeat the last plane
#pragma unroll
                for (int r = 0; r < kRows; ++r) res[r] = q_mid[l][r];
                rf = f_mid[l];
            } else {
#pragma unroll
                for (int r = 0; r < kRows; ++r) res[r] = nw[r];
                rf = 0;
            }
            // shift the queue of level l-1 (its only consumer is done) and take its newest plane
#pragma unroll
            for (int r = 0; r < kRows; ++r) {
                q_lo[l - 1][r] = q_mid[l - 1][r];
                q_mid[l - 1][r] = nw[r];
                nw[r] = res[r];
            }
            f_mid[l - 1] = nf;
            nf = rf;
            if (l < T) {  // publish the new plane of level l for the y neighbours of the next iteration
                float* dst = sm_lev + ((l - 1) * 2 + (it & 1)) * kPlane;
#pragma unroll
                for (int r = 0; r < kRows; ++r) *reinterpret_cast<float4*>(dst + off[r]) = nw[r];
            } else if (run && zc >= zs && zc < ze) {  // level T: the pass's output
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    if ((own_bits >> (4 * r)) & 1u) {
                        const size_t row = (size_t)zc * P.ny + gy[r];
                        *reinterpret_cast<float4*>(p_out + row * P.nx + gx) = nw[r];
                    }
                    // bit-packed freeze flags: two quads (8 cells) per byte, written by the odd lane
                    const unsigned nib = (nf >> (4 * r)) & 0xFu;
                    const unsigned hi = __shfl_down_sync(0xffffffffu, nib, 1);
                    if (((own_bits >> (4 * r)) & 1u) && (lane & 1))
                        m_out[((size_t)zc * P.ny + gy[r]) * nxb + (gx >> 3)] = (unsigned char)(nib | (hi << 4));
                }
            }
        }
        __syncthreads();
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
    if (tid < T && tid < levels) {
        const unsigned v = s_cnt[tid];
        if (v) atomicAdd(&state->active_after[s0 + tid], (unsigned long long)v);
    }
    if (tid == 0) {
        brick_state[brick] = (s_cnt[levels - 1] == 0u) ? 1 : 0;
        atomicAdd(&state->bricks_processed, 1ull);
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

bool make_plane_map(CUtensorMap* map, float* base, int nx, int ny, int nz_alloc) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz_alloc};
    const cuuint64_t strides[2] = {(cuuint64_t)nx * 4, (cuuint64_t)nx * ny * 4};
    const cuuint32_t box[3] = {(cuuint32_t)kTileX, (cuuint32_t)kTileY, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int T>
cudaError_t launch_T(const FusedJacobi& J, const Domain& d, const FrameParams* frame, StepState* state, int pass,
                     int iters, int early_exit, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(jacobi_pass_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)Smem<T>::kBytes);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    PassParams P;
    P.nx = d.nx; P.ny = d.ny; P.nz_alloc = d.nz_alloc;
    P.z_face_lo = 0 - d.z_first;
    P.z_face_hi = d.nz - d.z_first;
    P.z_out0 = d.z_own0 - d.z_first; P.z_out1 = d.z_own1 - d.z_first;
    P.bz = J.bz; P.ntx = J.ntx; P.nty = J.nty; P.nzc = J.nzc;
    P.pass = pass; P.levels_total = iters; P.early_exit = early_exit;
    const dim3 grid(J.ntx, J.nty, J.nzc);
    jacobi_pass_kernel<T><<<grid, kThreads, Smem<T>::kBytes, stream>>>(
        *reinterpret_cast<const CUtensorMap*>(J.map_p[0]), *reinterpret_cast<const CUtensorMap*>(J.map_p[1]),
        *reinterpret_cast<const CUtensorMap*>(J.map_rhs), frame, state, J.p[0], J.p[1], J.mask[0], J.mask[1],
        J.brick_state, P);
    return cudaGetLastError();
}

}  // namespace

bool fused_jacobi_supported(const Domain& d) { return d.nz > 1 && (d.nx % 8) == 0 && d.nx >= 8; }

int fused_jacobi_plan(FusedJacobi* J, const Domain& d, int fuse_t, float* p0, float* p1, float* rhs) {
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    J->T = fuse_t;
    const int out_y = kTileY - 2 * fuse_t;
    J->ntx = (d.nx + kOutX - 1) / kOutX;
    J->nty = (d.ny + out_y - 1) / out_y;
    const int nz_out = d.z_own1 - d.z_own0;
    J->bz = nz_out >= 64 ? 32 : (nz_out >= 16 ? 16 : nz_out);
    if (const char* e = getenv("FXB_BZ")) {  // tuning knob: planes per brick
        const int v = atoi(e);
        if (v >= 1 && v <= nz_out) J->bz = v;
    }
    J->nzc = (nz_out + J->bz - 1) / J->bz;
    J->p[0] = p0; J->p[1] = p1; J->rhs = rhs;
    if (!make_plane_map(reinterpret_cast<CUtensorMap*>(J->map_p[0]), p0, d.nx, d.ny, d.nz_alloc)) return -1;
    if (!make_plane_map(reinterpret_cast<CUtensorMap*>(J->map_p[1]), p1, d.nx, d.ny, d.nz_alloc)) return -1;
    if (!make_plane_map(reinterpret_cast<CUtensorMap*>(J->map_rhs), rhs, d.nx, d.ny, d.nz_alloc)) return -1;
    return 0;
}

size_t fused_jacobi_bricks(const FusedJacobi& J) { return (size_t)J.ntx * J.nty * J.nzc; }

cudaError_t launch_jacobi_pass_fused(const FusedJacobi& J, const Domain& d, const FrameParams* frame, StepState* state,
                                     int pass, int iters, int early_exit, cudaStream_t stream) {
    switch (J.T) {
        case 1: return launch_T<1>(J, d, frame, state, pass, iters, early_exit, stream);
        case 2: return launch_T<2>(J, d, frame, state, pass, iters, early_exit, stream);
        case 3: return launch_T<3>(J, d, frame, state, pass, iters, early_exit, stream);
        case 4: return launch_T<4>(J, d, frame, state, pass, iters, early_exit, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace fxb
