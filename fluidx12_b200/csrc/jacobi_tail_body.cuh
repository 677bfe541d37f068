// jacobi_tail_body.cuh — body of the block-resident multi-sweep Jacobi kernel ("tail kernel", jacobi_tail.cu).
//
// Same relaxation as jacobi_fused.cu (CSPoisson.hlsli:8-26 under SURVEY.md App. A.3), for the long tail of the
// solve in which only a few bricks still hold an active cell and the z-marching bulk kernel is bound by the latency
// of one brick chain per launch.  Here a CTA loads one sub-block of an active brick — 40 x 12 x 8 output cells plus a
// halo of TT cells per side — once, keeps it on chip (every thread owns the z column of one quad in registers, the
// xy neighbours are exchanged through shared memory) and applies TT sweeps before it stores the result, so a launch
// advances the solve by TT sweeps along a dependency chain of TT short phases instead of bz + 2T marching steps.
// Validity shrinks by one cell per sweep from every window edge that is not a grid face (the halo of TT cells
// absorbs exactly that); at a grid face the reference's clamp-to-edge rule applies (CSProject3D.hlsl:76-83).
//
// Two ways to run the sweeps on a window, chosen per window from the number of active cells in it:
//   * dense  (crowded windows): every thread owns the z column of one quad in registers, xy neighbours through
//     shared memory, a warp skips a plane only when none of its 128 cells is active;
//   * sparse (the usual case in the tail: a few percent of the cells are active, scattered): the active cells are
//     compacted into a list in shared memory and only list entries are relaxed — new values go to a side array and
//     are committed after a barrier (two-phase Jacobi on one window copy), a frozen cell leaves the list.
//     Measured reason (profiles/README.md, round 1): with the dense form alone a launch cost ~85 us at 256^3 because
//     almost all lanes of every executed warp instruction belonged to frozen cells.
//
// The file is written against a tiny portability layer (FXT_*) so that the SAME statements compile under nvcc
// (threads = CUDA threads, phases separated by __syncthreads) and under g++ in tests/emu/tail_emu.cpp (threads = a
// loop, phases = consecutive loops).  The emulation is test infrastructure: it lets the CPU suite check this file
// bit for bit against the oracle; it is never linked into libfluidx_b200.so.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define FXT_FN __device__ __forceinline__
#define FXT_COMPILER_FENCE() asm volatile("" ::: "memory")  // keeps a batch of loads ahead of the stores that follow
#define FXT_FMA(a, b, c) __fmaf_rn((a), (b), (c))
#define FXT_POPC(v) __popc(v)
#define FXT_FFS(v) __ffs(v)
#define FXT_ATOMIC_ADD_U32(p, v) atomicAdd((p), (v))
#define FXT_ATOMIC_ADD_I32(p, v) atomicAdd((p), (v))
#define FXT_ATOMIC_ADD_U64(p, v) atomicAdd((p), (v))
#define FXT_LDG_U8(p) __ldg(p)
#define FXT_LDG_F32(p) __ldg(p)
#define FXT_ATOMIC_AND_U32(p, v) atomicAnd((p), (v))
// 16-byte asynchronous copy global -> shared (LDGSTS): no register in between, so a thread's copies are all in flight
// at once; `valid` false zero-fills the destination (the source address must still be a mapped one).
#define FXT_CP_ASYNC_16(dst, src, valid)                                                                       \
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((unsigned)__cvta_generic_to_shared(dst)), \
                 "l"(src), "r"((valid) ? 16 : 0)                                                               \
                 : "memory")
#define FXT_CP_ASYNC_WAIT() asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory")
// TMA staging of a whole window (cp_async mode 2): ONE cp.async.bulk.tensor.3d, issued by one thread, brings the
// 48 x 20 x 16 box into shared memory — cells outside the array are zero-filled by the copy engine, which is exactly
// the window's definition — and completes on an mbarrier every thread then waits on.
#define FXT_TMA_ISSUE(tma, dst, x, y, z, bytes, p_in, P, it)                                                          \
    do {                                                                                                               \
        const unsigned bar_ = (unsigned)__cvta_generic_to_shared((tma).bar);                                           \
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_), "r"((unsigned)(bytes))      \
                     : "memory");                                                                                      \
        asm volatile(                                                                                                  \
            "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" \
            ::"r"((unsigned)__cvta_generic_to_shared(dst)), "l"((tma).map), "r"(x), "r"(y), "r"(z), "r"(bar_)          \
            : "memory");                                                                                               \
    } while (0)
#define FXT_TMA_WAIT(tma)                                                                            \
    do {                                                                                             \
        const unsigned bar_ = (unsigned)__cvta_generic_to_shared((tma).bar);                         \
        const unsigned parity_ = (tma).uses & 1u;                                                    \
        asm volatile(                                                                                \
            "{\n"                                                                                    \
            ".reg .pred p;\n"                                                                        \
            "FXT_TMA_WAIT_LOOP:\n"                                                                   \
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"                                \
            "@p bra FXT_TMA_WAIT_DONE;\n"                                                            \
            "bra FXT_TMA_WAIT_LOOP;\n"                                                               \
            "FXT_TMA_WAIT_DONE:\n"                                                                   \
            "}\n" ::"r"(bar_),                                                                       \
            "r"(parity_)                                                                             \
            : "memory");                                                                             \
    } while (0)
// A work item is a sequence of phases separated by barriers (tail_run_item).  Under nvcc a phase is just the
// statement, executed by the calling thread `tid` with its state `t`.
#define FXT_CTX(S) const int tid
#define FXT_THREAD_STATE(S) TailThread<S> t
#define FXT_PHASE(...) do { __VA_ARGS__; } while (0)
#define FXT_SYNC() __syncthreads()
#define FXT_END() do { } while (0)
// Phase timing (library built with EXTRA=-DFXB_TAIL_TIMING, read by tools/tail_probe.py): thread 0 of every CTA
// accumulates the cycles between consecutive marks of an item and adds them to marks[8 * path + k] at the end,
// plus one count in marks[24 + path].
#ifdef FXB_TAIL_TIMING
#define FXT_MARK_BEGIN() long long fxt_t0 = clock64(), fxt_d[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define FXT_MARK(k) do { const long long c_ = clock64(); fxt_d[k] += c_ - fxt_t0; fxt_t0 = c_; } while (0)
#define FXT_MARK_END(marks, path)                                                                      \
    do {                                                                                               \
        if (tid == 0 && (marks) != nullptr) {                                                          \
            for (int k_ = 0; k_ < 8; ++k_)                                                             \
                if (fxt_d[k_]) atomicAdd(&(marks)[8 * (path) + k_], (unsigned long long)fxt_d[k_]);    \
            atomicAdd(&(marks)[24 + (path)], 1ull);                                                    \
        }                                                                                              \
    } while (0)
#else
#define FXT_MARK_BEGIN() do { } while (0)
#define FXT_MARK(k) do { } while (0)
#define FXT_MARK_END(marks, path) do { } while (0)
#endif
namespace fxb { typedef float4 Quad; }
#else
#include <cmath>
#define FXT_FN inline
#define FXT_COMPILER_FENCE() asm volatile("" ::: "memory")
#define FXT_FMA(a, b, c) std::fmaf((a), (b), (c))
#define FXT_POPC(v) __builtin_popcount(v)
#define FXT_FFS(v) __builtin_ffs(v)
#define FXT_ATOMIC_ADD_U32(p, v) (*(p) += (v))
static inline int fxt_fetch_add_i32(int* p, int v) { const int o = *p; *p = o + v; return o; }
#define FXT_ATOMIC_ADD_I32(p, v) fxt_fetch_add_i32((p), (v))
#define FXT_ATOMIC_ADD_U64(p, v) (*(p) += (v))
#define FXT_LDG_U8(p) (*(p))
#define FXT_LDG_F32(p) (*(p))
#define FXT_ATOMIC_AND_U32(p, v) (*(p) &= (v))
#define FXT_CP_ASYNC_16(dst, src, valid)                                      \
    do {                                                                      \
        const fxb::Quad zero_ = {0.f, 0.f, 0.f, 0.f};                         \
        *reinterpret_cast<fxb::Quad*>(dst) = (valid) ? *reinterpret_cast<const fxb::Quad*>(src) : zero_; \
    } while (0)
#define FXT_CP_ASYNC_WAIT() (void)0
// TMA staging, emulated: the box copy with zero fill outside the array, done by the issuing thread
#define FXT_TMA_ISSUE(tma, dst, x, y, z, bytes, p_in, P, it) fxb::tail_emulated_box_copy<S>(dst, p_in, P, x, y, z)
#define FXT_TMA_WAIT(tma) (void)0
// Emulation: phases are queued and run at the next barrier, thread after thread — every thread runs ALL phases of
// the barrier-free segment before the next thread starts (in ascending or descending thread order).  That is the
// most skewed interleaving a missing __syncthreads() would permit, so a phase that needs another thread's result
// from the same segment reads stale data and the parity tests fail.
#include <functional>
#include <vector>
#define FXT_CTX(S) TailEmu<S>& emu
#define FXT_THREAD_STATE(S) (void)0
#define FXT_PHASE(...) emu.seg.push_back([&](int tid) { TailThread<S>& t = emu.th[tid]; (void)t; (void)tid; __VA_ARGS__; })
#define FXT_SYNC() emu.flush()
#define FXT_END() emu.flush()
#define FXT_MARK_BEGIN() (void)0
#define FXT_MARK(k) (void)0
#define FXT_MARK_END(marks, path) (void)0
namespace fxb { struct alignas(16) Quad { float x, y, z, w; }; }
#endif

namespace fxb {

constexpr float kTailInv6 = 0.166666672f;   // 0x3e2aaaab
constexpr float kTailEps = 0.00100000005f;  // 0x3a83126f

// Compile-time shape: TT sweeps per launch, output sub-block OXQ quads x OY rows x OZ planes.
template <int TT_, int OXQ_, int OY_, int OZ_>
struct TailShape {
    static constexpr int TT = TT_, OXQ = OXQ_, OY = OY_, OZ = OZ_;
    static_assert(TT_ >= 1 && TT_ <= 4, "the x halo is one quad");
    static_assert(OXQ_ % 2 == 0, "a sub-block owns whole bytes of the bit-packed masks: its x extent is a multiple of 8 cells");
    static constexpr int OX = 4 * OXQ_;
    static constexpr int LXQ = OXQ_ + 2, LX = 4 * LXQ;  // window: one halo quad per side in x
    static constexpr int LY = OY_ + 2 * TT_, LZ = OZ_ + 2 * TT_;
    static constexpr int kUsed = LXQ * LY;               // threads that own a column
    static constexpr int kThreads = (kUsed + 31) / 32 * 32;
    static constexpr int kPlane = LX * LY;                 // floats per window plane
    static constexpr int kRhsPlane = LX * (LY - 2);
    static constexpr int kFlagWords = (4 * LZ + 31) / 32;
    static_assert(LZ <= 16, "flag words / dirty mask are sized for 16 planes");
    // shared memory (floats): window values, right-hand side of the cells that can be relaxed, then bytes
    static constexpr int kPFloats = LZ * kPlane;
    static constexpr int kRhsFloats = (LZ - 2) * kRhsPlane;
    static constexpr int kNibBytes = LZ * LY * LXQ;
    static constexpr int kCtrlWords = 8 + TT_;  // any-own flag, per-level counters, listable-cell count, spare
    static constexpr int kCtrlTotal = TT_ + 1;  // ctrl index of the number of active cells that can be relaxed at all
    static constexpr int kCtrlBar = (TT_ + 3) & ~1;  // ctrl index (even: 8-byte aligned) of the TMA mbarrier; never reset
    // sparse path: list entries (u32) + their right-hand sides (f32) + new values (f32) live where the dense path
    // stages the right-hand side of the whole window
    static constexpr int kListCap = (kRhsFloats * 4 / 12) / 32 * 32;
    static constexpr int kScanWords = kThreads + kThreads / 32;
    // second dense path: the quads that can be relaxed at all (window planes 1..LZ-2, rows 1..LY-2), a fixed share per
    // thread; their new values go to a side array that takes the place of the staged right-hand side
    static constexpr int kQuads = (LZ - 2) * (LY - 2) * LXQ;
    static constexpr int kQuadsPerThread = (kQuads + kThreads - 1) / kThreads;
    static_assert(kQuads * 4 <= kRhsFloats, "side array of the second dense path");
    static_assert(kPFloats <= (1 << 14), "list entries keep the window index in 14 bits");
    static_assert(kNibBytes % 4 == 0, "flag bytes are cleared with 32-bit atomics");
    static constexpr size_t kBytes =
        (size_t)(kPFloats + kRhsFloats) * 4 + kCtrlWords * 4 + kScanWords * 4 + kNibBytes;
};

// list entry of the sparse path
constexpr unsigned kTailIdxMask = 0x3FFFu;  // window index of the cell
constexpr int kTailDepthShift = 14;        // 3 bits: sweeps for which the cell is a valid output (1..TT)
constexpr unsigned kTailOwn = 1u << 17;     // the cell belongs to the CTA's output region
constexpr unsigned kTailFresh = 1u << 18;   // a new value waits in the side array
constexpr unsigned kTailFroze = 1u << 19;   // ... and the cell froze with it
constexpr int kTailClampShift = 20;         // 6 bits: the L, R, U, D, F, B neighbour is the cell itself (grid face)
constexpr unsigned kTailDead = 0xFFFFFFFFu;

// What one launch needs to know (uniform over the grid).
struct TailParams {
    int nx, ny;          // grid extent in x, y
    int nz_alloc;        // local planes allocated
    int z_face_lo;       // local index of global plane 0 (may be negative: on another rank)
    int z_face_hi;       // local index one past global plane nz-1 (may exceed nz_alloc)
    int z_out0, z_out1;  // local planes this rank owns
    int bx, by, bz;      // brick extent (jacobi_fused.cu: 120 x (TILE_Y - 2T) x bz)
    int ntx, nty;        // brick grid in x, y
    int nsub;            // sub-blocks per brick in x
    int first;           // 1: no flags exist yet (first kernel of a frame): every cell is active
    int early_exit;
    int levels;          // sweeps to apply (<= TT); the CUDA kernel passes the run-time value separately
    int sparse_cap;      // windows with at most this many relaxable active cells take the sparse path (<= kListCap)
    int cp_async;        // window staging: 0 through registers; 1 cp.async, requested together with the flags; 2 TMA box copy
    int dense_mode;      // crowded windows: 1 = register z-columns, 2 = two-phase update of all quads (side array)
};

// TMA staging state (mode 2): the tensor map of the pressure buffer that is this launch's input, the mbarrier in shared
// memory, and how many copies have completed on it so far (its phase parity).
struct TailTma {
    const void* map = nullptr;
    unsigned long long* bar = nullptr;
    unsigned uses = 0;
};

#if !defined(__CUDACC__)
// What the copy engine does for a box at (x0, y0, z0): elements outside the array [0,nx) x [0,ny) x [0,nz_alloc) are 0.
template <class S>
inline void tail_emulated_box_copy(float* dst, const float* p_in, const TailParams& P, int x0, int y0, int z0);
#endif

// Shared-memory view.
template <class S>
struct TailShared {
    float* p;            // [LZ][LY][LX]
    float* rhs;          // [LZ-2][LY-2][LX]  (window planes 1..LZ-2, rows 1..LY-2)
    unsigned char* nib;  // [LZ][LY][LXQ] freeze flags, one nibble per quad
    unsigned* ctrl;      // [0] any active cell in the own region; [1 + l] active own cells after level l+1; [kCtrlTotal]
    int* scan;           // [kThreads] list entries per thread, then [kThreads / 32] per warp
    unsigned* list;      // [kListCap] sparse path (aliases rhs)
    float* newv;         // [kListCap] sparse path (aliases rhs)
    float* rhsv;         // [kListCap] sparse path (aliases rhs)
};

// Carves the views out of one block of S::kBytes bytes (16-byte aligned).
template <class S>
FXT_FN TailShared<S> tail_shared(void* base) {
    TailShared<S> sh;
    sh.p = reinterpret_cast<float*>(base);
    sh.rhs = sh.p + S::kPFloats;
    sh.ctrl = reinterpret_cast<unsigned*>(sh.rhs + S::kRhsFloats);
    sh.scan = reinterpret_cast<int*>(sh.ctrl + S::kCtrlWords);
    sh.nib = reinterpret_cast<unsigned char*>(sh.scan + S::kScanWords);
    sh.list = reinterpret_cast<unsigned*>(sh.rhs);
    sh.newv = reinterpret_cast<float*>(sh.list + S::kListCap);
    sh.rhsv = sh.newv + S::kListCap;
    return sh;
}

#if !defined(__CUDACC__)
template <class S>
inline void tail_emulated_box_copy(float* dst, const float* p_in, const TailParams& P, int x0, int y0, int z0) {
    for (int z = 0; z < S::LZ; ++z)
        for (int y = 0; y < S::LY; ++y)
            for (int x = 0; x < S::LX; ++x) {
                const int gx = x0 + x, gy = y0 + y, gz = z0 + z;
                const bool in = gx >= 0 && gx < P.nx && gy >= 0 && gy < P.ny && gz >= 0 && gz < P.nz_alloc;
                dst[(z * S::LY + y) * S::LX + x] = in ? p_in[((size_t)gz * P.ny + gy) * P.nx + gx] : 0.0f;
            }
}
#endif

// Per-thread state that lives across phases (registers under nvcc).
template <class S>
struct TailThread {
    Quad v[S::LZ];
    unsigned fl[2];     // freeze flags of the column: nibble z of the 64-bit word (1 = active)
    unsigned own[2];    // nibbles of the cells this CTA must produce
    unsigned dirty;     // bit z: v[z] changed in the current sweep
    int qx, y;          // window coordinates of the column
    int gx, gy;         // grid coordinates
    bool used;          // owns a column at all
    bool in_xy;         // column lies inside the grid
    bool own_xy;        // column belongs to the output region
    Quad rq[S::kQuadsPerThread];  // second dense path: right-hand sides of the thread's quads (loaded once per item)
    unsigned qinfo[S::kQuadsPerThread];  // ... and what the sweeps need to know about each of them (see kQuad* below)
    int nlist;          // active cells of the column that can be relaxed (sparse path: its list entries)
    unsigned okx;       // 0x11111111 * (4-bit mask of the column's cells that can be relaxed as far as x and y go)
};

// Geometry of one work item (uniform over the CTA).
template <class S>
struct TailItem {
    int ox, oy, oz;     // first output cell (grid x, grid y, local plane)
    int ex, ey, ez;     // output extent (<= OX, OY, OZ; <= 0: nothing to do)
    int wx, wy, wz;     // window origin
    int zvl, zvh;       // window planes inside the array and the grid: [zvl, zvh)
    int yvl, yvh;       // window rows inside the grid
    bool zlo_face, zhi_face, ylo_face, yhi_face;  // the valid range ends at a grid face (clamp rule) rather than at a window edge
    bool xlo_face, xhi_face;
    unsigned zok[2];    // nibble z = 0xF when cells of window plane z can be relaxed at all (tail_depth >= 1 along z)
};

template <class S>
FXT_FN TailItem<S> tail_item(const TailParams& P, int brick, int sub) {
    TailItem<S> it;
    const int tx = brick % P.ntx, ty = (brick / P.ntx) % P.nty, zc = brick / (P.ntx * P.nty);
    it.ox = tx * P.bx + sub * S::OX;
    it.oy = ty * P.by;
    it.oz = P.z_out0 + zc * P.bz;
    const int bx_end = (tx + 1) * P.bx < P.nx ? (tx + 1) * P.bx : P.nx;
    it.ex = bx_end - it.ox < S::OX ? bx_end - it.ox : S::OX;
    it.ey = P.ny - it.oy < P.by ? P.ny - it.oy : P.by;
    it.ez = P.z_out1 - it.oz < P.bz ? P.z_out1 - it.oz : P.bz;
    it.wx = it.ox - 4;
    it.wy = it.oy - S::TT;
    it.wz = it.oz - S::TT;
    const int zlo = P.z_face_lo > 0 ? P.z_face_lo : 0, zhi = P.z_face_hi < P.nz_alloc ? P.z_face_hi : P.nz_alloc;
    it.zvl = zlo - it.wz > 0 ? zlo - it.wz : 0;
    it.zvh = zhi - it.wz < S::LZ ? zhi - it.wz : S::LZ;
    it.yvl = -it.wy > 0 ? -it.wy : 0;
    it.yvh = P.ny - it.wy < S::LY ? P.ny - it.wy : S::LY;
    // A face that coincides with the first / last window plane or row is treated like a plain window edge: the cells
    // on it have no right-hand side staged, and nothing that far out is needed (validity shrinks from there).
    it.zlo_face = it.wz + it.zvl == P.z_face_lo && it.zvl >= 1;
    it.zhi_face = it.wz + it.zvh == P.z_face_hi && it.zvh <= S::LZ - 1;
    it.ylo_face = it.wy + it.yvl == 0 && it.yvl >= 1;
    it.yhi_face = it.wy + it.yvh == P.ny && it.yvh <= S::LY - 1;
    it.xlo_face = it.wx < 0;
    it.xhi_face = it.wx + S::LX > P.nx;
    it.zok[0] = it.zok[1] = 0u;
    for (int z = 0; z < S::LZ; ++z) {
        const bool ok = z >= it.zvl && z < it.zvh && (it.zlo_face || z - it.zvl >= 1) && (it.zhi_face || it.zvh - 1 - z >= 1);
        if (ok) it.zok[z >> 3] |= 0xFu << (4 * (z & 7));
    }
    return it;
}

// Number of sweeps after which the window cell (x, y, z) is still a valid result: its distance to the nearest window
// edge that is not a grid face, at most TT.  0: the cell can never be relaxed here (it sits on such an edge).
template <class S>
FXT_FN int tail_depth(const TailItem<S>& it, int x, int y, int z) {
    int d = S::TT;
    if (!it.xlo_face && x < d) d = x;
    if (!it.xhi_face && S::LX - 1 - x < d) d = S::LX - 1 - x;
    if (!it.ylo_face && y - it.yvl < d) d = y - it.yvl;
    if (!it.yhi_face && it.yvh - 1 - y < d) d = it.yvh - 1 - y;
    if (!it.zlo_face && z - it.zvl < d) d = z - it.zvl;
    if (!it.zhi_face && it.zvh - 1 - z < d) d = it.zvh - 1 - z;
    return d;
}

FXT_FN unsigned tail_nib(const unsigned (&fl)[2], int z) { return (fl[z >> 3] >> (4 * (z & 7))) & 0xFu; }
FXT_FN void tail_set_nib(unsigned (&fl)[2], int z, unsigned n) {
    fl[z >> 3] = (fl[z >> 3] & ~(0xFu << (4 * (z & 7)))) | (n << (4 * (z & 7)));
}

// ---- phase 0: thread geometry, freeze flags of the column, "is anything in the own region still active" ----------
template <class S>
FXT_FN void tail_phase_flags(int tid, TailThread<S>& t, const TailShared<S>& sh, const TailItem<S>& it,
                             const TailParams& P, const unsigned char* __restrict__ m_in,
                             const float* __restrict__ p_in) {
    t.used = tid < S::kUsed;
    t.qx = tid % S::LXQ;
    t.y = tid / S::LXQ;
    t.gx = it.wx + 4 * t.qx;
    t.gy = it.wy + t.y;
    t.in_xy = t.used && t.gx >= 0 && t.gx < P.nx && t.gy >= 0 && t.gy < P.ny;
    t.own_xy = t.in_xy && t.qx >= 1 && 4 * (t.qx - 1) < it.ex && t.y >= S::TT && t.y - S::TT < it.ey;
    t.fl[0] = t.fl[1] = 0u;
    t.own[0] = t.own[1] = 0u;
    t.dirty = 0u;
    if (P.cp_async == 1 && t.used) {
        // cp.async mode: the window is requested here, together with the flags, so that every path of the item
        // (copy, sparse, dense) finds it in shared memory after ONE memory round trip; no register is involved
#pragma unroll
        for (int z = 0; z < S::LZ; ++z) {
            const bool in = t.in_xy && z >= it.zvl && z < it.zvh;
            FXT_CP_ASYNC_16(sh.p + z * S::kPlane + t.y * S::LX + 4 * t.qx,
                            in ? p_in + ((size_t)(it.wz + z) * P.ny + t.gy) * P.nx + t.gx : p_in, in);
        }
    }
    const int nxb = P.nx >> 3;
    // All flag bytes of the column are requested before the first one is decoded (an in-order core would otherwise
    // pay one memory round trip per plane); planes outside the grid re-read byte 0 of the mask and are ignored.
    unsigned raw[S::LZ];
#pragma unroll
    for (int z = 0; z < S::LZ; ++z) {
        const bool in = t.in_xy && z >= it.zvl && z < it.zvh && !P.first;
        raw[z] = FXT_LDG_U8(in ? m_in + ((size_t)(it.wz + z) * P.ny + t.gy) * nxb + (t.gx >> 3) : m_in);
    }
    FXT_COMPILER_FENCE();
#pragma unroll
    for (int z = 0; z < S::LZ; ++z) {
        if (!t.in_xy || z < it.zvl || z >= it.zvh) continue;
        const unsigned n = P.first ? 0xFu : (raw[z] >> (t.gx & 4)) & 0xFu;
        t.fl[z >> 3] |= n << (4 * (z & 7));
        if (t.own_xy && z >= S::TT && z - S::TT < it.ez) t.own[z >> 3] |= 0xFu << (4 * (z & 7));
    }
    if (((t.fl[0] & t.own[0]) | (t.fl[1] & t.own[1])) != 0u) sh.ctrl[0] = 1u;  // benign race: everybody stores 1
    // how many list entries the column would contribute (sparse path): active cells with tail_depth >= 1
    unsigned okx = 0u;
    if (t.in_xy && (it.ylo_face || t.y - it.yvl >= 1) && (it.yhi_face || it.yvh - 1 - t.y >= 1)) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int x = 4 * t.qx + j;
            if ((it.xlo_face || x >= 1) && (it.xhi_face || S::LX - 1 - x >= 1)) okx |= 1u << j;
        }
    }
    t.okx = okx * 0x11111111u;
    const int nl = FXT_POPC(t.fl[0] & it.zok[0] & t.okx) + FXT_POPC(t.fl[1] & it.zok[1] & t.okx);
    t.nlist = nl;
    sh.scan[tid] = nl;
    if (nl) FXT_ATOMIC_ADD_U32(&sh.ctrl[S::kCtrlTotal], (unsigned)nl);
    if (P.cp_async == 1) FXT_CP_ASYNC_WAIT();  // the window has landed before the barrier that follows this phase
}

// ---- copy path: no active cell in the own region, so the output equals the input ---------------------------------
template <class S>
FXT_FN void tail_phase_copy(TailThread<S>& t, const TailShared<S>& sh, const TailItem<S>& it, const TailParams& P,
                            const float* __restrict__ p_in, float* __restrict__ p_out,
                            unsigned char* __restrict__ m_out) {
    if (!t.own_xy) return;
    const int nxb = P.nx >> 3;
    Quad q[S::OZ];  // all loads in flight before the first store (planes past the end re-read the first one)
    if (P.cp_async) {  // the window is already in shared memory
#pragma unroll
        for (int z = S::TT; z < S::TT + S::OZ; ++z) {
            const int zz = z - S::TT < it.ez ? z : S::TT;
            q[z - S::TT] = *reinterpret_cast<const Quad*>(sh.p + zz * S::kPlane + t.y * S::LX + 4 * t.qx);
        }
    } else {
#pragma unroll
        for (int z = S::TT; z < S::TT + S::OZ; ++z) {
            const int zz = z - S::TT < it.ez ? z : S::TT;
            q[z - S::TT] = *reinterpret_cast<const Quad*>(p_in + ((size_t)(it.wz + zz) * P.ny + t.gy) * P.nx + t.gx);
        }
    }
#pragma unroll
    for (int z = S::TT; z < S::TT + S::OZ; ++z) {
        if (z - S::TT >= it.ez) continue;
        const size_t row = (size_t)(it.wz + z) * P.ny + t.gy;
        *reinterpret_cast<Quad*>(p_out + row * P.nx + t.gx) = q[z - S::TT];
        if (t.qx & 1) m_out[row * nxb + (t.gx >> 3)] = 0;
    }
}

// ---- phase 1: window values -> registers + shared memory, right-hand side -> shared memory -----------------------
template <class S>
FXT_FN void tail_phase_load(TailThread<S>& t, const TailShared<S>& sh, const TailItem<S>& it, const TailParams& P,
                            const float* __restrict__ p_in, const float* __restrict__ rhs) {
    if (!t.used) return;
    const Quad zero = {0.f, 0.f, 0.f, 0.f};
    if (P.cp_async) {  // the flags phase has staged the window: only the register copy is missing
#pragma unroll
        for (int z = 0; z < S::LZ; ++z)
            t.v[z] = *reinterpret_cast<const Quad*>(sh.p + z * S::kPlane + t.y * S::LX + 4 * t.qx);
    } else {  // (the mode test stays outside the loop so that the sixteen loads are issued back to back)
#pragma unroll
        for (int z = 0; z < S::LZ; ++z) {
            const bool in = t.in_xy && z >= it.zvl && z < it.zvh;
            t.v[z] = *reinterpret_cast<const Quad*>(in ? p_in + ((size_t)(it.wz + z) * P.ny + t.gy) * P.nx + t.gx : p_in);
        }
#pragma unroll
        for (int z = 0; z < S::LZ; ++z) {
            const bool in = t.in_xy && z >= it.zvl && z < it.zvh;
            if (!in) t.v[z] = zero;
            *reinterpret_cast<Quad*>(sh.p + z * S::kPlane + t.y * S::LX + 4 * t.qx) = t.v[z];
        }
    }
    if (t.y >= 1 && t.y <= S::LY - 2) {
#pragma unroll
        for (int z = 1; z <= S::LZ - 2; ++z) {
            Quad q = zero;
            if (t.in_xy && z >= it.zvl && z < it.zvh)
                q = *reinterpret_cast<const Quad*>(rhs + ((size_t)(it.wz + z) * P.ny + t.gy) * P.nx + t.gx);
            *reinterpret_cast<Quad*>(sh.rhs + (z - 1) * S::kRhsPlane + (t.y - 1) * S::LX + 4 * t.qx) = q;
        }
    }
}

// One cell: the six additions in the DXBC's order, x = acc * (1/6), freeze test on the fused
// difference.  Returns the new value (active cells) or the old one (frozen cells); clears the cell's flag on freeze.
FXT_FN float tail_cell(float c, float l, float r, float u, float d, float f, float b, float rhs, unsigned bit,
                       float eps, unsigned& act) {
    float acc = l + rhs;
    acc = r + acc;
    acc = u + acc;
    acc = d + acc;
    acc = f + acc;
    acc = b + acc;
    const float xn = acc * kTailInv6;
    const float diff = FXT_FMA(acc, kTailInv6, -c);
    const bool is_active = (act & bit) != 0u;
    act &= (is_active && fabsf(diff) < eps) ? ~bit : ~0u;  // branch-free: selects, not jumps
    return is_active ? xn : c;
}

// ---- phase A of sweep s (1-based): new values of the column into registers (reads shared memory only) ------------
template <class S>
FXT_FN void tail_phase_relax(TailThread<S>& t, const TailShared<S>& sh, const TailItem<S>& it, const TailParams& P,
                             int s) {
    t.dirty = 0u;
    if (!t.in_xy || (t.fl[0] | t.fl[1]) == 0u) return;
    // rows and planes that are still valid inputs of this sweep (see the file header)
    const int yc0 = it.ylo_face ? it.yvl : it.yvl + s, yc1 = it.yhi_face ? it.yvh : it.yvh - s;
    if (t.y < yc0 || t.y >= yc1) return;
    const int zc0 = it.zlo_face ? it.zvl : it.zvl + s, zc1 = it.zhi_face ? it.zvh : it.zvh - s;
    const float eps = P.early_exit ? kTailEps : -1.0f;
    const int col = t.y * S::LX + 4 * t.qx;
    const int up = (it.ylo_face && t.y == it.yvl) ? col : col - S::LX;          // clamp rule in y
    const int dn = (it.yhi_face && t.y == it.yvh - 1) ? col : col + S::LX;
    const bool clamp_l = t.qx == 0 || t.gx == 0;                                  // clamp rule / window edge in x
    const bool clamp_r = t.qx == S::LXQ - 1 || t.gx + 4 == P.nx;
    const int rcol = (t.y - 1) * S::LX + 4 * t.qx;
    Quad below = t.v[0];  // previous sweep's value of plane z-1
#pragma unroll
    for (int z = 0; z < S::LZ; ++z) {
        const Quad c = t.v[z];
        unsigned a = tail_nib(t.fl, z);
        if (a != 0u && z >= zc0 && z < zc1) {
            const float* pl = sh.p + z * S::kPlane;
            const Quad f = (it.zlo_face && z == it.zvl) ? c : below;
            const Quad b = (it.zhi_face && z == it.zvh - 1) ? c : t.v[z + 1 < S::LZ ? z + 1 : z];
            const Quad u = *reinterpret_cast<const Quad*>(pl + up);
            const Quad d = *reinterpret_cast<const Quad*>(pl + dn);
            const float left = clamp_l ? c.x : pl[col - 1];
            const float right = clamp_r ? c.w : pl[col + 4];
            const Quad r = *reinterpret_cast<const Quad*>(sh.rhs + (z - 1) * S::kRhsPlane + rcol);
            Quad n;
            n.x = tail_cell(c.x, left, c.y, u.x, d.x, f.x, b.x, r.x, 1u, eps, a);
            n.y = tail_cell(c.y, c.x, c.z, u.y, d.y, f.y, b.y, r.y, 2u, eps, a);
            n.z = tail_cell(c.z, c.y, c.w, u.z, d.z, f.z, b.z, r.z, 4u, eps, a);
            n.w = tail_cell(c.w, c.z, right, u.w, d.w, f.w, b.w, r.w, 8u, eps, a);
            t.v[z] = n;
            tail_set_nib(t.fl, z, a);
            t.dirty |= 1u << z;
        }
        below = c;
    }
}

// ---- phase B of sweep s: publish the changed planes, count the own cells that are still active -------------------
template <class S>
FXT_FN void tail_phase_publish(TailThread<S>& t, const TailShared<S>& sh, int s, bool last) {
    if (!t.used) return;
    if (!last && t.dirty) {
#pragma unroll
        for (int z = 0; z < S::LZ; ++z)
            if ((t.dirty >> z) & 1u) *reinterpret_cast<Quad*>(sh.p + z * S::kPlane + t.y * S::LX + 4 * t.qx) = t.v[z];
    }
    const unsigned n = FXT_POPC(t.fl[0] & t.own[0]) + FXT_POPC(t.fl[1] & t.own[1]);
    if (n) FXT_ATOMIC_ADD_U32(&sh.ctrl[s], n);
}

// ---- final phases: own cells -> the other pressure buffer; flags -> bit-packed mask (two quads per byte) ----------
template <class S>
FXT_FN void tail_phase_store(TailThread<S>& t, const TailShared<S>& sh, const TailItem<S>& it, const TailParams& P,
                             float* __restrict__ p_out) {
    if (!t.used) return;
#pragma unroll
    for (int z = S::TT; z < S::TT + S::OZ; ++z) {
        sh.nib[(z * S::LY + t.y) * S::LXQ + t.qx] = (unsigned char)tail_nib(t.fl, z);
        if (t.own_xy && z - S::TT < it.ez)
            *reinterpret_cast<Quad*>(p_out + ((size_t)(it.wz + z) * P.ny + t.gy) * P.nx + t.gx) = t.v[z];
    }
}

template <class S>
FXT_FN void tail_phase_store_mask(TailThread<S>& t, const TailShared<S>& sh, const TailItem<S>& it, const TailParams& P,
                                  unsigned char* __restrict__ m_out) {
    // own quads start at window quad 1 (grid x a multiple of 8) and come in pairs: the odd one writes the byte
    if (!t.own_xy || !(t.qx & 1)) return;
    const int nxb = P.nx >> 3;
#pragma unroll
    for (int z = S::TT; z < S::TT + S::OZ; ++z) {
        if (z - S::TT >= it.ez) continue;
        const unsigned lo = tail_nib(t.fl, z), hi = sh.nib[(z * S::LY + t.y) * S::LXQ + t.qx + 1];
        m_out[((size_t)(it.wz + z) * P.ny + t.gy) * nxb + (t.gx >> 3)] = (unsigned char)(lo | (hi << 4));
    }
}

// =====================================================================================================================
// Sparse path
// =====================================================================================================================

// ---- sparse phase 1: entries per warp (exclusive scan of the per-thread counts, two levels) ------------------------
template <class S>
FXT_FN void tail_sparse_scan(int tid, const TailShared<S>& sh) {
    if ((tid & 31) != 0) return;
    int sum = 0;
    for (int l = 0; l < 32; ++l) sum += sh.scan[tid + l];
    sh.scan[S::kThreads + (tid >> 5)] = sum;
}

// ---- sparse phase 2: window values -> shared memory, flags -> nibble array, active cells -> list -------------------
template <class S>
FXT_FN void tail_sparse_build(int tid, TailThread<S>& t, const TailShared<S>& sh, const TailItem<S>& it,
                              const TailParams& P, const float* __restrict__ p_in, const bool with_list = true) {
    if (!t.used) return;
    if (!P.cp_async) {  // (in cp.async mode the flags phase has staged the window already)
        const Quad zero = {0.f, 0.f, 0.f, 0.f};
        Quad q[S::LZ];  // cells outside the grid read p_in[0..3] and store zero
#pragma unroll
        for (int z = 0; z < S::LZ; ++z) {
            const bool in = t.in_xy && z >= it.zvl && z < it.zvh;
            q[z] = *reinterpret_cast<const Quad*>(in ? p_in + ((size_t)(it.wz + z) * P.ny + t.gy) * P.nx + t.gx : p_in);
        }
#pragma unroll
        for (int z = 0; z < S::LZ; ++z) {
            const bool in = t.in_xy && z >= it.zvl && z < it.zvh;
            *reinterpret_cast<Quad*>(sh.p + z * S::kPlane + t.y * S::LX + 4 * t.qx) = in ? q[z] : zero;
        }
    }
#pragma unroll
    for (int z = 0; z < S::LZ; ++z) sh.nib[(z * S::LY + t.y) * S::LXQ + t.qx] = (unsigned char)tail_nib(t.fl, z);
    if (t.nlist == 0 || !with_list) return;
    int at = 0;  // entries of the threads before this one
    for (int w = 0; w < (tid >> 5); ++w) at += sh.scan[S::kThreads + w];
    for (int l = tid & ~31; l < tid; ++l) at += sh.scan[l];
#pragma unroll
    for (int w = 0; w < 2; ++w) {
        unsigned m = t.fl[w] & it.zok[w] & t.okx;
        while (m) {
            const int bit = FXT_FFS(m) - 1;
            m &= m - 1u;
            const int z = 8 * w + (bit >> 2), j = bit & 3;
            const bool own = t.own_xy && z >= S::TT && z - S::TT < it.ez;
            const int depth = tail_depth<S>(it, 4 * t.qx + j, t.y, z);
            // clamp-to-edge at the grid faces (CSProject3D.hlsl:76-83), decided once per entry
            const int gx = t.gx + j, gz = it.wz + z;
            const unsigned clamp = (gx == 0 ? 1u : 0u) | (gx == P.nx - 1 ? 2u : 0u) | (t.gy == 0 ? 4u : 0u) |
                                   (t.gy == P.ny - 1 ? 8u : 0u) | (gz == P.z_face_lo ? 16u : 0u) |
                                   (gz == P.z_face_hi - 1 ? 32u : 0u);
            sh.list[at++] = (unsigned)(z * S::kPlane + t.y * S::LX + 4 * t.qx + j) | ((unsigned)depth << kTailDepthShift) |
                            (own ? kTailOwn : 0u) | (clamp << kTailClampShift);
        }
    }
}

// ---- sparse phase 3: right-hand sides of the listed cells -> side array (eight independent loads in flight) --------
template <class S>
FXT_FN void tail_sparse_gather(int tid, const TailShared<S>& sh, const TailItem<S>& it, const TailParams& P,
                               const float* __restrict__ rhs, int n) {
    for (int base = tid; base < n; base += 8 * S::kThreads) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int e = base + u * S::kThreads;
            // entries past the end re-read entry `base` (in range): no branch around the load, nothing out of bounds
            const int idx = (int)(sh.list[e < n ? e : base] & kTailIdxMask);
            const int z = idx / S::kPlane, r = idx - z * S::kPlane, y = r / S::LX, x = r - y * S::LX;
            v[u] = FXT_LDG_F32(rhs + ((size_t)(it.wz + z) * P.ny + (it.wy + y)) * P.nx + (it.wx + x));
        }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (base + u * S::kThreads < n) sh.rhsv[base + u * S::kThreads] = v[u];
    }
}

// ---- sparse phase A of sweep s: new values of the listed cells into the side array ---------------------------------
template <class S>
FXT_FN void tail_sparse_relax(int tid, const TailShared<S>& sh, const TailItem<S>& it, const TailParams& P, int n,
                              int s) {
    const float eps = P.early_exit ? kTailEps : -1.0f;
    // Four entries per trip: their reads are independent, so the latencies overlap instead of adding up.
    for (int base = tid; base < n; base += 4 * S::kThreads) {
        unsigned ent[4];
        float nv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = base + u * S::kThreads;
            ent[u] = e < n ? sh.list[e] : kTailDead;
            if (ent[u] != kTailDead && (int)((ent[u] >> kTailDepthShift) & 7u) < s) ent[u] = kTailDead;  // skipped this sweep
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            nv[u] = 0.f;
            if (ent[u] == kTailDead) continue;
            const int e = base + u * S::kThreads;
            const int idx = (int)(ent[u] & kTailIdxMask);
            const unsigned open_ = ~(ent[u] >> kTailClampShift);  // bit k set: neighbour k is a different cell
            const float c = sh.p[idx];
            // a clamped neighbour is the cell itself (offset 0); elsewhere the neighbour is inside the window
            const float l = sh.p[idx - (int)(open_ & 1u)];
            const float rr = sh.p[idx + (int)((open_ >> 1) & 1u)];
            const float up = sh.p[idx - S::LX * (int)((open_ >> 2) & 1u)];
            const float dn = sh.p[idx + S::LX * (int)((open_ >> 3) & 1u)];
            const float f = sh.p[idx - S::kPlane * (int)((open_ >> 4) & 1u)];
            const float b = sh.p[idx + S::kPlane * (int)((open_ >> 5) & 1u)];
            unsigned act = 1u;
            nv[u] = tail_cell(c, l, rr, up, dn, f, b, sh.rhsv[e], 1u, eps, act);
            ent[u] |= kTailFresh | (act ? 0u : kTailFroze);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (ent[u] == kTailDead) continue;  // nothing computed: the list entry stays as it is
            const int e = base + u * S::kThreads;
            sh.newv[e] = nv[u];
            sh.list[e] = ent[u];
        }
    }
}

// ---- sparse phase B of sweep s: commit the new values, retire frozen cells, count the live own cells ---------------
template <class S>
FXT_FN void tail_sparse_commit(int tid, const TailShared<S>& sh, int n, int s) {
    unsigned live = 0;
    for (int base = tid; base < n; base += 4 * S::kThreads) {
        unsigned ent[4];
        float nv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int e = base + u * S::kThreads;
            ent[u] = e < n ? sh.list[e] : kTailDead;
            nv[u] = sh.newv[e < n ? e : base];
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (ent[u] == kTailDead || !(ent[u] & kTailFresh)) continue;
            const int e = base + u * S::kThreads;
            const int idx = (int)(ent[u] & kTailIdxMask);
            sh.p[idx] = nv[u];
            if (ent[u] & kTailFroze) {
                const int z = idx / S::kPlane, r = idx - z * S::kPlane, y = r / S::LX, x = r - y * S::LX;
                const int byte = (z * S::LY + y) * S::LXQ + (x >> 2);
                FXT_ATOMIC_AND_U32(reinterpret_cast<unsigned*>(sh.nib) + (byte >> 2), ~(1u << (8 * (byte & 3) + (x & 3))));
                sh.list[e] = kTailDead;
            } else {
                sh.list[e] = ent[u] & ~kTailFresh;
                if (ent[u] & kTailOwn) ++live;
            }
        }
    }
    if (live) FXT_ATOMIC_ADD_U32(&sh.ctrl[s], live);
}

// ---- sparse final phase: own cells and their flags from shared memory to the other buffers --------------------------
template <class S>
FXT_FN void tail_sparse_store(TailThread<S>& t, const TailShared<S>& sh, const TailItem<S>& it, const TailParams& P,
                              float* __restrict__ p_out, unsigned char* __restrict__ m_out) {
    if (!t.own_xy) return;
    const int nxb = P.nx >> 3;
#pragma unroll
    for (int z = S::TT; z < S::TT + S::OZ; ++z) {
        if (z - S::TT >= it.ez) continue;
        const size_t row = (size_t)(it.wz + z) * P.ny + t.gy;
        *reinterpret_cast<Quad*>(p_out + row * P.nx + t.gx) =
            *reinterpret_cast<const Quad*>(sh.p + z * S::kPlane + t.y * S::LX + 4 * t.qx);
        if (t.qx & 1) {
            const unsigned char* nb = sh.nib + (z * S::LY + t.y) * S::LXQ + t.qx;
            m_out[row * nxb + (t.gx >> 3)] = (unsigned char)(nb[0] | (nb[1] << 4));
        }
    }
}

// =====================================================================================================================
// Second dense path (crowded windows): every thread relaxes a fixed share of the window's quads, new values go to a
// side array and are committed after a barrier — no register-resident columns, so no spills, and the same two-phase
// structure as the sparse path.  The window and the flag nibbles are staged by tail_sparse_build (without a list).
// =====================================================================================================================

// quad number q -> window coordinates (planes 1..LZ-2, rows 1..LY-2, every quad of a row)
template <class S>
FXT_FN void tail_quad_coords(int q, int& z, int& y, int& qx) {
    z = 1 + q / ((S::LY - 2) * S::LXQ);
    const int r = q - (z - 1) * ((S::LY - 2) * S::LXQ);
    y = 1 + r / S::LXQ;
    qx = r - (y - 1) * S::LXQ;
}

// Per-quad word computed once per item: quad index in the window (= index into the nibble array; x4 = float index),
// the sweeps for which the quad is a valid output (its y/z distance to window edges that are not grid faces), which
// neighbours are clamped to the cell itself (grid faces; for L/R also the window's x edges, whose cells are spoilt
// anyway), whether the quad belongs to the output region, and whether the slot holds a quad at all.
constexpr unsigned kQuadIdxMask = 0xFFFu;
constexpr int kQuadDepthShift = 12, kQuadClampShift = 15;
constexpr unsigned kQuadOwn = 1u << 21, kQuadValid = 1u << 22;

// ---- right-hand sides of the thread's quads -> registers (all loads in flight; quads outside the grid read rhs[0..3]);
// ---- the per-quad words ------------------------------------------------------------------------------------------------
template <class S>
FXT_FN void tail_dense2_load_rhs(int tid, TailThread<S>& t, const TailItem<S>& it, const TailParams& P,
                                 const float* __restrict__ rhs) {
    static_assert(S::LZ * S::LY * S::LXQ <= 4096, "quad index field");
#pragma unroll
    for (int k = 0; k < S::kQuadsPerThread; ++k) {
        const int q = tid + k * S::kThreads;
        int z, y, qx;
        tail_quad_coords<S>(q < S::kQuads ? q : 0, z, y, qx);
        const int gx = it.wx + 4 * qx, gy = it.wy + y, gz = it.wz + z;
        const bool in = q < S::kQuads && gx >= 0 && gx < P.nx && gy >= 0 && gy < P.ny && z >= it.zvl && z < it.zvh;
        t.rq[k] = *reinterpret_cast<const Quad*>(in ? rhs + ((size_t)gz * P.ny + gy) * P.nx + gx : rhs);
        const unsigned clamp = ((qx == 0 || gx == 0) ? 1u : 0u) | ((qx == S::LXQ - 1 || gx + 4 == P.nx) ? 2u : 0u) |
                               (gy == 0 ? 4u : 0u) | (gy == P.ny - 1 ? 8u : 0u) | (gz == P.z_face_lo ? 16u : 0u) |
                               (gz == P.z_face_hi - 1 ? 32u : 0u);
        const bool own = qx >= 1 && 4 * (qx - 1) < it.ex && y >= S::TT && y - S::TT < it.ey && z >= S::TT && z - S::TT < it.ez;
        const int depth = tail_depth<S>(it, S::TT, y, z);  // x = TT: no limit from the x edges (handled by the clamps)
        t.qinfo[k] = (unsigned)((z * S::LY + y) * S::LXQ + qx) | ((unsigned)(depth < 0 ? 0 : depth) << kQuadDepthShift) |
                     (clamp << kQuadClampShift) | (own ? kQuadOwn : 0u) | (in ? kQuadValid : 0u);
    }
}

// ---- phase A of sweep s: new values of the thread's quads -> side array; new flags -> high nibble of the flag byte ---
template <class S>
FXT_FN void tail_dense2_relax(int tid, TailThread<S>& t, const TailShared<S>& sh, const TailParams& P, int s) {
    const float eps = P.early_exit ? kTailEps : -1.0f;
    Quad* side = reinterpret_cast<Quad*>(sh.rhs);
    // not unrolled: the per-quad words and right-hand sides are indexed at run time (they live in local memory, which
    // L1 serves), in exchange the loop body keeps its temporaries in registers
#pragma unroll 1
    for (int k = 0; k < S::kQuadsPerThread; ++k) {
        const unsigned info = t.qinfo[k];
        if (!(info & kQuadValid) || (int)((info >> kQuadDepthShift) & 7u) < s) continue;
        const int nb = (int)(info & kQuadIdxMask);
        unsigned a = sh.nib[nb] & 0xFu;
        if (a == 0u) continue;
        const unsigned open_ = ~(info >> kQuadClampShift);  // bit k set: neighbour k is a different cell
        const float* pc = sh.p + 4 * nb;
        const Quad c = *reinterpret_cast<const Quad*>(pc);
        // a clamped neighbour is the cell itself: offset 0 (CSProject3D.hlsl:76-83)
        const Quad u = *reinterpret_cast<const Quad*>(pc - S::LX * (int)((open_ >> 2) & 1u));
        const Quad d = *reinterpret_cast<const Quad*>(pc + S::LX * (int)((open_ >> 3) & 1u));
        const Quad f = *reinterpret_cast<const Quad*>(pc - S::kPlane * (int)((open_ >> 4) & 1u));
        const Quad b = *reinterpret_cast<const Quad*>(pc + S::kPlane * (int)((open_ >> 5) & 1u));
        const float left = pc[-(int)(open_ & 1u)];
        const float right = pc[3 + (int)((open_ >> 1) & 1u)];
        const Quad r = t.rq[k];
        Quad n;
        n.x = tail_cell(c.x, left, c.y, u.x, d.x, f.x, b.x, r.x, 1u, eps, a);
        n.y = tail_cell(c.y, c.x, c.z, u.y, d.y, f.y, b.y, r.y, 2u, eps, a);
        n.z = tail_cell(c.z, c.y, c.w, u.z, d.z, f.z, b.z, r.z, 4u, eps, a);
        n.w = tail_cell(c.w, c.z, right, u.w, d.w, f.w, b.w, r.w, 8u, eps, a);
        side[tid + k * S::kThreads] = n;
        sh.nib[nb] = (unsigned char)((sh.nib[nb] & 0xFu) | (a << 4));  // old flags stay in the low nibble until the commit
    }
}

// ---- phase B of sweep s: commit the quads relaxed in phase A, count the own cells that are still active -------------
template <class S>
FXT_FN void tail_dense2_commit(int tid, TailThread<S>& t, const TailShared<S>& sh, int s) {
    const Quad* side = reinterpret_cast<const Quad*>(sh.rhs);
    unsigned live = 0;
#pragma unroll 1
    for (int k = 0; k < S::kQuadsPerThread; ++k) {
        const unsigned info = t.qinfo[k];
        if (!(info & kQuadValid)) continue;
        const int nb = (int)(info & kQuadIdxMask);
        const unsigned byte = sh.nib[nb];
        if ((byte & 0xFu) != 0u && (int)((info >> kQuadDepthShift) & 7u) >= s) {  // the same predicate as in phase A
            *reinterpret_cast<Quad*>(sh.p + 4 * nb) = side[tid + k * S::kThreads];
            sh.nib[nb] = (unsigned char)(byte >> 4);
            if (info & kQuadOwn) live += FXT_POPC(byte >> 4);
        } else if (info & kQuadOwn) {
            live += FXT_POPC(byte & 0xFu);
        }
    }
    if (live) FXT_ATOMIC_ADD_U32(&sh.ctrl[s], live);
}

// Work lists of one launch (same lists as jacobi_fused.cu: bricks that still hold an active cell, and bricks that
// froze in the previous kernel and need one copy into the other pressure buffer).
struct TailWork {
    const int* relax_in;
    const int* copy_in;
    int n_relax, n_copy;
    int* relax_out;
    int* copy_out;
    int* relax_out_count;
    int* copy_out_count;
    int* brick_state;  // per brick: arrivals of its sub-blocks (low byte) and "still active" votes; zero between launches
};

// ---- after the last phase of a relax item (one thread): histogram, and the brick's entry in the next lists --------
template <class S>
FXT_FN void tail_finish_item(const TailShared<S>& sh, const TailParams& P, const TailWork& W, int brick,
                             unsigned long long* active_after_s0, const int levels) {
    for (int l = 0; l < levels; ++l)
        if (sh.ctrl[1 + l]) FXT_ATOMIC_ADD_U64(&active_after_s0[l], (unsigned long long)sh.ctrl[1 + l]);
    const bool active = sh.ctrl[levels] != 0u;
    const int old = FXT_ATOMIC_ADD_I32(&W.brick_state[brick], 1 + (active ? 256 : 0));
    if ((old & 255) == P.nsub - 1) {  // last sub-block of the brick: all votes are in
        W.brick_state[brick] = 0;
        if (active || (old >> 8) != 0) W.relax_out[FXT_ATOMIC_ADD_I32(W.relax_out_count, 1)] = brick;
        else W.copy_out[FXT_ATOMIC_ADD_I32(W.copy_out_count, 1)] = brick;
    }
}

#if !defined(__CUDACC__)
// Emulated CTA: per-thread state plus the phases queued since the last barrier (see FXT_PHASE above).
template <class S>
struct TailEmu {
    std::vector<TailThread<S>> th = std::vector<TailThread<S>>(S::kThreads);
    std::vector<std::function<void(int)>> seg;
    bool descending = false;
    void flush() {
        for (int i = 0; i < S::kThreads; ++i) {
            const int tid = descending ? S::kThreads - 1 - i : i;
            for (auto& f : seg) f(tid);
        }
        seg.clear();
    }
};
#endif

// ---- one work item of the relax list: sub-block `sub` of `brick`, all phases and barriers ---------------------------
// Returns the path taken: 0 = nothing active in the own region (copy) or empty sub-block, 1 = sparse, 2 = dense.
// DENSE: which dense path is compiled in — 1 or 2 — or 0 for both with P.dense_mode choosing (the emulation; the
// CUDA build instantiates one kernel per dense path so that neither carries the other's register pressure).
template <class S, int DENSE = 0>
FXT_FN int tail_run_item(FXT_CTX(S), const TailShared<S>& sh, const TailParams& P, const TailWork& W, const int brick,
                         const int sub, const float* __restrict__ p_in, float* __restrict__ p_out,
                         const float* __restrict__ rhs, const unsigned char* __restrict__ m_in,
                         unsigned char* __restrict__ m_out, unsigned long long* active_after_s0,
                         unsigned long long* marks, TailTma& tma, const int levels) {
    const TailItem<S> it = tail_item<S>(P, brick, sub);
    FXT_THREAD_STATE(S);
    FXT_MARK_BEGIN();
    int path = 0;
    FXT_SYNC();  // the previous item is finished with shared memory
    FXT_PHASE(if (tid < S::kCtrlBar) sh.ctrl[tid] = 0u);  // (the words from kCtrlBar on hold the TMA mbarrier)
    FXT_SYNC();
    FXT_MARK(0);
    if (it.ex > 0) {
        if (P.cp_async == 2)
            FXT_PHASE(if (tid == 0) FXT_TMA_ISSUE(tma, sh.p, it.wx, it.wy, it.wz, S::kPFloats * 4, p_in, P, it));
        FXT_PHASE(tail_phase_flags<S>(tid, t, sh, it, P, m_in, p_in));
        if (P.cp_async == 2) {
            FXT_PHASE(FXT_TMA_WAIT(tma));
            tma.uses += 1u;
        }
        FXT_SYNC();
        FXT_MARK(1);
        const int n_list = (int)sh.ctrl[S::kCtrlTotal];
        if (sh.ctrl[0] == 0u) {  // nothing active in the own region: the output equals the input
            FXT_PHASE(if (tid == 0) tail_finish_item<S>(sh, P, W, brick, active_after_s0, levels));  // (its atomics overlap the copy)
            FXT_PHASE(tail_phase_copy<S>(t, sh, it, P, p_in, p_out, m_out));
            FXT_MARK(6);
        } else if (n_list <= P.sparse_cap) {  // relax a compacted list of the active cells
            path = 1;
            FXT_PHASE(tail_sparse_scan<S>(tid, sh));
            FXT_SYNC();
            FXT_MARK(2);
            FXT_PHASE(tail_sparse_build<S>(tid, t, sh, it, P, p_in));
            FXT_SYNC();
            FXT_MARK(3);
            FXT_PHASE(tail_sparse_gather<S>(tid, sh, it, P, rhs, n_list));
            FXT_SYNC();
            FXT_MARK(4);
            for (int s = 1; s <= levels; ++s) {
                FXT_PHASE(tail_sparse_relax<S>(tid, sh, it, P, n_list, s));
                FXT_SYNC();
                FXT_PHASE(tail_sparse_commit<S>(tid, sh, n_list, s));
                FXT_SYNC();
            }
            FXT_MARK(5);
            // the counters are final: thread 0 starts the brick's bookkeeping (two dependent atomics) while the other
            // warps store the result
            FXT_PHASE(if (tid == 0) tail_finish_item<S>(sh, P, W, brick, active_after_s0, levels));
            FXT_PHASE(tail_sparse_store<S>(t, sh, it, P, p_out, m_out));
            FXT_MARK(6);
        } else if (DENSE == 2 || (DENSE == 0 && P.dense_mode == 2)) {  // crowded window: two-phase update of all quads
            path = 2;
            FXT_PHASE(tail_sparse_build<S>(tid, t, sh, it, P, p_in, false); tail_dense2_load_rhs<S>(tid, t, it, P, rhs));
            FXT_SYNC();
            FXT_MARK(3);
            for (int s = 1; s <= levels; ++s) {
                FXT_PHASE(tail_dense2_relax<S>(tid, t, sh, P, s));
                FXT_SYNC();
                FXT_PHASE(tail_dense2_commit<S>(tid, t, sh, s));
                FXT_SYNC();
            }
            FXT_MARK(5);
            FXT_PHASE(if (tid == 0) tail_finish_item<S>(sh, P, W, brick, active_after_s0, levels));
            FXT_PHASE(tail_sparse_store<S>(t, sh, it, P, p_out, m_out));
            FXT_MARK(6);
        } else {  // crowded window: register columns
            path = 2;
            FXT_PHASE(tail_phase_load<S>(t, sh, it, P, p_in, rhs));
            FXT_SYNC();
            FXT_MARK(3);
            for (int s = 1; s <= levels; ++s) {
                FXT_PHASE(tail_phase_relax<S>(t, sh, it, P, s));
                FXT_SYNC();
                FXT_PHASE(tail_phase_publish<S>(t, sh, s, s == levels));
                FXT_SYNC();
            }
            FXT_MARK(5);
            FXT_PHASE(if (tid == 0) tail_finish_item<S>(sh, P, W, brick, active_after_s0, levels));
            FXT_PHASE(tail_phase_store<S>(t, sh, it, P, p_out));
            FXT_SYNC();
            FXT_PHASE(tail_phase_store_mask<S>(t, sh, it, P, m_out));
            FXT_MARK(6);
        }
    } else {
        FXT_PHASE(if (tid == 0) tail_finish_item<S>(sh, P, W, brick, active_after_s0, levels));  // an empty sub-block (outside the grid) still casts its brick's vote
    }
    FXT_MARK(7);
    FXT_MARK_END(marks, path);
    FXT_END();
    return path;
}

// ---- a brick of the copy list: its values are final; bring the other buffer (and mask) up to date ----------------
// `tid` of `nthreads` threads; pure streaming, 16 bytes per access.
FXT_FN void tail_copy_brick(int tid, int nthreads, const TailParams& P, int brick, const float* __restrict__ p_in,
                            float* __restrict__ p_out, unsigned char* __restrict__ m_out) {
    const int tx = brick % P.ntx, ty = (brick / P.ntx) % P.nty, zc = brick / (P.ntx * P.nty);
    const int x_lo = tx * P.bx, y_lo = ty * P.by, zs = P.z_out0 + zc * P.bz;
    const int planes = (P.z_out1 - zs < P.bz ? P.z_out1 - zs : P.bz);
    const int rows = P.ny - y_lo < P.by ? P.ny - y_lo : P.by;
    const int qpr = (P.nx - x_lo < P.bx ? P.nx - x_lo : P.bx) >> 2;  // quads per row inside the grid
    const int nxb = P.nx >> 3;
    const int total = planes * rows * qpr;
    for (int base = tid; base < total; base += 4 * nthreads) {
        Quad v[4];
        size_t at[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // four independent loads in flight; slots past the end repeat slot 0
            const int i = base + u * nthreads < total ? base + u * nthreads : base;
            const int xq = i % qpr, rz = i / qpr;
            const size_t row = (size_t)(zs + rz / rows) * P.ny + (y_lo + rz % rows);
            at[u] = row * P.nx + x_lo + 4 * xq;
            v[u] = *reinterpret_cast<const Quad*>(p_in + at[u]);
            if (xq & 1) m_out[row * nxb + ((x_lo + 4 * xq) >> 3)] = 0;
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (u == 0 || base + u * nthreads < total) *reinterpret_cast<Quad*>(p_out + at[u]) = v[u];
    }
}

}  // namespace fxb
