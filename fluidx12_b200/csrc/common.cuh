// common.cuh — shared device-side types and helpers for the fluidx_b200 kernels (sm_100a).
//
// Numerics contract: every kernel reproduces the operation order of the shipped
// DXBC.  The library is compiled with -fmad=false, so a fused multiply-add
// happens only where __fmaf_rn is written (a DXBC `mad`); `/` is IEEE (-prec-div=true), denormals
// are kept (-ftz=false), fp32->fp16 stores are round-to-nearest-even.
#pragma once

#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fxb {

// Local (per-rank) view of the grid.  Every field is stored x-fastest as [plane][y][x]; local plane
// index lz holds global plane z = z_first + lz.  Single GPU: z_first = 0, nz_alloc = nz.
struct Domain {
    int nx, ny, nz;      // global grid
    int pitch;           // floats per row of the PRESSURE-SIDE arrays (pressure ping-pong, right-hand side; freeze masks:
                         // pitch / 8 bytes per row): nx rounded up to a multiple of 8 on the tuned 3D path, so that rows
                         // start 32-byte aligned for TMA and 128-bit accesses whatever the grid width (e.g. the 150^3 of
                         // Bin/FluidGI.bat); nx otherwise.  Velocity and colour rows are nx texels apart.
    int z_first;         // global z of local plane 0 (may be negative: halo below the global face)
    int nz_alloc;        // planes allocated locally (owned + halos)
    int z_own0, z_own1;  // owned global planes [z_own0, z_own1)
};

// Per-frame parameters: the replacement of the reference's CBSimulation constant buffer
// (Fluid.cpp:12-16) plus the frame parity (Fluid.h:124).  Uploaded host->device once per
// fxb_simulate; kernels read it through a const pointer so one captured graph serves every frame.
struct FrameParams {
    float dt;
    int parity;  // m_frameParity: advect writes colour[parity], reads colour[!parity]
    unsigned long long epoch_base;  // multi-GPU: kEventsPerFrame x the frame's index (see PeerView)
};

// Device-side per-step solver state and counters.
struct StepState {
    int p_cur;                           // which pressure buffer currently holds P (persists across frames)
    int s_exec;                          // sweeps executed in the last step
    int passes;                          // fused passes executed in the last step
    int halo_overflow;                   // sticky
    unsigned long long total_sweeps;     // cumulative over all steps
    unsigned long long total_passes;
    unsigned long long bricks_processed;  // cumulative: bricks fully relaxed by fused passes
    unsigned long long bricks_copied;     // cumulative: frozen bricks copied once to the other buffer
    unsigned long long active_after[128];  // [k] = cells still active after sweep k (this rank)
    int done_ctas;                       // CTAs of the running fused pass that have finished (multi-GPU event publish)
    int pad0;
    unsigned long long phase_ns[8];      // cumulative device time per phase (phase marks; fxb_config.phase_timing)
    unsigned long long mark_ns;          // %globaltimer of the last phase mark
#ifdef FXB_TIMING
    long long dbg[128];                  // debug build: cycle stamps (jacobi_fused.cu FXB_STAMP)
#endif
};

// ---- fused halos (multi-GPU, fxb_config.halo_backend = FXB_HALO_FUSED) ------------------------------------------
// Every kernel of the step writes the planes next to an interior slab face twice: into its own array and — the same
// store instruction stream, over NVLink — into the halo planes of the neighbouring rank's array (mapped through CUDA
// IPC).  There is no exchange kernel and no collective on the data path.  Ordering uses ONE monotone event counter per
// rank: the kernels of a frame are numbered m = 0 (advect), 1 (divergence), 2 + k (fused pass k), 2 + n (settle),
// 3 + n (gradient), the same on every rank; a rank publishes epoch_base + m + 1 into both neighbours' memory when
// kernel m has completed (the gradient publishes the next frame's base), and the parts of kernel m that read halo
// planes or write into a neighbour wait until that neighbour's counter has reached epoch_base + m: the neighbour has
// then finished kernel m - 1, i.e. its pushes into this rank's halos have landed and it no longer reads the halo planes
// this kernel is about to overwrite.  Everything away from the slab faces runs without waiting.
constexpr unsigned long long kEventsPerFrame = 128;

struct PeerView {
    int has_lo, has_hi;          // an interior face below / above (both 0 on a single GPU: all of this is skipped)
    int dz_lo, dz_hi;            // local plane index in the neighbour's arrays = local plane index here + dz
    unsigned long long* events;  // [2] in this rank's memory: the counters rank - 1 / rank + 1 publish
    unsigned long long* ev_lo;   // rank - 1's event words (this rank writes [1] there: it is its upper neighbour)
    unsigned long long* ev_hi;   // rank + 1's event words (this rank writes [0])
    unsigned long long* error;   // sticky time-out flag (this rank's memory)
    long long timeout_cycles;
};

// One thread waits until the neighbours named have completed every kernel before `need`.  Never hangs the device: a
// wait that exceeds the limit records the failure (fxb_sync reports it) and goes on.
__device__ __forceinline__ void peer_wait(const PeerView& pv, const unsigned long long need, const bool lo, const bool hi) {
    const long long t0 = clock64();
    for (int side = 0; side < 2; ++side) {
        if (!(side == 0 ? (lo && pv.has_lo) : (hi && pv.has_hi))) continue;
        const unsigned long long* w = pv.events + side;
        for (;;) {
            // acquire load at system scope: what the neighbour wrote before publishing is visible to what follows
            // (measured on B200: ~2 us per wait cheaper than a relaxed poll followed by a full system fence)
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(w) : "memory");
            if (v >= need) break;
            if (clock64() - t0 > pv.timeout_cycles) {
                *pv.error = 1ull;
                break;
            }
            __nanosleep(32);
        }
    }
    // the neighbour's stores are also read through the async proxy (TMA) by the thread that waited
    asm volatile("fence.proxy.async;" ::: "memory");
}

// One thread, after everything the kernel wrote (here and into the neighbours) is complete.
__device__ __forceinline__ void peer_publish(const PeerView& pv, const unsigned long long value) {
    __threadfence_system();
    if (pv.has_lo) *reinterpret_cast<volatile unsigned long long*>(pv.ev_lo + 1) = value;
    if (pv.has_hi) *reinterpret_cast<volatile unsigned long long*>(pv.ev_hi + 0) = value;
}

// Fused halos: the z chunks of a kernel that lie next to an interior slab face wait for the neighbour rank.  CTAs are
// scheduled in blockIdx order, so those chunks are handed out LAST: the wait then overlaps the interior chunks instead
// of occupying every SM with spinning CTAs.  b: blockIdx.z; returns the chunk it works on (interior chunks first, then
// the n_lo chunks at the lower face, then the n_hi at the upper face).  `chunk` planes per chunk, `reach` planes next
// to a face that wait, n owned planes.
__host__ __device__ __forceinline__ int face_last_chunk_of(int has_lo, int has_hi, int b, int nchunks, int chunk,
                                                           int reach, int n) {
    int n_lo = has_lo ? (reach + chunk - 1) / chunk : 0;
    int n_hi = has_hi ? nchunks - (n - reach > 0 ? n - reach : 0) / chunk : 0;
    if (n_lo > nchunks) n_lo = nchunks;
    if (n_hi > nchunks - n_lo) n_hi = nchunks - n_lo;
    const int n_int = nchunks - n_lo - n_hi;
    return b < n_int ? b + n_lo : (b < n_int + n_lo ? b - n_int : b);
}
__device__ __forceinline__ int face_last_chunk(const PeerView& pv, int b, int nchunks, int chunk, int reach, int n) {
    return face_last_chunk_of(pv.has_lo, pv.has_hi, b, nchunks, chunk, reach, n);
}

// Phase marks: the device time since the previous mark is added to StepState::phase_ns[slot] (slot < 0: start of a
// step).  The step's own one-thread kernels (frame constants, begin-step, finish-solve) carry three of the five marks;
// with fxb_config.phase_timing two more one-thread kernels close the divergence and the gradient phase.
__device__ __forceinline__ void phase_mark(StepState* state, int slot) {
    unsigned long long now;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
    if (slot >= 0) state->phase_ns[slot] += now - state->mark_ns;
    state->mark_ns = now;
}

// Static emitter table (Impulse.hlsli:14-18 is time-independent): basis values of the voxels in a
// conservative bounding box of the sphere basis >= exp(-4).  Computed on the host at fxb_create
// with libm exp2f and indexed by global voxel coordinates.
struct Emitter {
    int x0, y0, z0, x1, y1, z1;  // box [x0,x1) x [y0,y1) x [z0,z1)
    const float* basis;          // [(z-z0)][(y-y0)][(x-x0)]
};

// Per-axis tables indexed by the global voxel coordinate, computed once on the host with the shader's own
// arithmetic: pos = (i + 0.5) / N (a true IEEE division), bp = fma(pos, 2, -1) and
// wall = clamp((-|bp| + 0.97) * (1/0.03), -1, 1), the soft-wall damping factor of CSProject3D.hlsl:106-108.
// `still[a][i]` (as a float, 1 or 0): a voxel at rest back-traces exactly onto its own texel centre along axis a, i.e.
// fma(pos[i], N, -0.5) == i, and both taps are inside the grid (i <= N - 2) — the advection's rest shortcut.
struct AxisTables {
    const float* pos[3];
    const float* bp[3];
    const float* wall[3];
    const float* still[3];
};

// Packed fp32 (Blackwell FADD2/FMUL2/FFMA2): two IEEE round-to-nearest operations per instruction, bit-identical
// to the scalar forms.
__device__ __forceinline__ float2 sub2(float2 a, float2 b) { return __fadd2_rn(a, make_float2(-b.x, -b.y)); }
__device__ __forceinline__ float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

__device__ __forceinline__ float4 load_texel4(const uint2* __restrict__ f, size_t i) {
    const uint2 r = __ldg(f + i);
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&r.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&r.y));
    return make_float4(a.x, a.y, b.x, b.y);
}

__device__ __forceinline__ uint2 pack_texel4(float x, float y, float z, float w) {
    const __half2 a = __floats2half2_rn(x, y);
    const __half2 b = __floats2half2_rn(z, w);
    uint2 r;
    r.x = *reinterpret_cast<const uint32_t*>(&a);
    r.y = *reinterpret_cast<const uint32_t*>(&b);
    return r;
}

__device__ __forceinline__ float half_bits_to_float(unsigned short h) {
    return __half2float(__ushort_as_half(h));
}

// Sampler addressing.  Fast path for taps within one period of the grid.
__device__ __forceinline__ int address_tap(int i, int w, int clamp_mode) {
    if (clamp_mode) return min(max(i, 0), w - 1);
    if ((unsigned)i < (unsigned)w) return i;
    const int period = 2 * w;
    int m = i % period;
    if (m < 0) m += period;
    return m < w ? m : period - 1 - m;
}

// floor(t) saturated to +-2^30 (NaN -> -2^30), identical to the oracle's floor_to_tap.
__device__ __forceinline__ int floor_to_tap(float t) {
    const float lim = 1073741824.0f;
    if (!(t > -lim)) return -(1 << 30);
    if (t > lim) return 1 << 30;
    return __float2int_rd(t);
}

}  // namespace fxb
