// jacobi_resident.cu — a fused Jacobi pass in brick-resident form (sm_100a): the latency shape of the pressure solve.
//
// Same arithmetic, work lists, freeze flags and buffers as the z-marching pass of jacobi_fused.cu (CSPoisson.hlsli:8-26
// under the restatement of SURVEY.md App. A.3), different schedule inside a brick.  After the first passes of a frame
// only a few hundred bricks still hold an active cell (profiles/: at 256^3 about 130 of 2112 from pass 4 on), every
// persistent CTA gets at most one of them, and a pass costs what ONE brick costs from its first load to its last
// store.  The marching kernel walks a brick plane by plane — 8 + 2T dependent iterations, each with a TMA issue, an
// mbarrier wait, the levels and a CTA barrier: ~20 us per pass however little work there is.  Here a brick is loaded
// whole instead:
//   * ONE cp.async.bulk.tensor.3d per array brings the brick's window (tile x tile x (8 + 2T) planes of pressure and
//     of right-hand side, 96 KB each at T = 2) into shared memory — one round trip instead of 12;
//   * a thread owns one quad column: its 8 + 2T pressure quads live in registers, so the z neighbours are registers,
//     the x neighbours warp shuffles, the y neighbours two shared-memory loads per quad and level;
//   * a level (sweep) runs over all its planes without a barrier in between — the planes are independent, so their
//     dependent chains overlap in the pipeline; levels are separated by two CTA barriers (level l is written back to
//     shared memory for the rows above and below);  T = 2: three barriers per brick instead of twelve;
//   * a warp whose quads are all frozen in a plane skips the plane's arithmetic (one vote).
// Shared memory: 2 x 96 KB, one CTA of 512 threads per SM.  Clamp-to-edge by index (x, y) or by the centre value (z),
// never by TMA fill, as in the marching kernel.  Launched with programmatic dependent launch: the next pass's CTAs are
// resident and past their prologue when this pass drains.
// Two shapes: RShape (a work item is a brick, T <= 2 sweeps per pass) and RShapeHalf4 (a work item is half a brick, FOUR
// sweeps per pass: the tail schedule, see the struct).  Measured on B200 (tools/timing_probe.py, 256^3, pass 16, first
// CTA): 0.65 us pass loads, 0.6 us item + TMA issue, 2.2 us until the window has landed, 1.1 + 0.7 us the two levels,
// 0.9 us stores and counters = 6.2 us, against ~20 us for the marching chain.
#include <cuda.h>

#include <type_traits>

#include "common.cuh"
#include "jacobi_common.cuh"
#include "kernels.h"

namespace fxb {

namespace {

// Tile of the resident kernel: the same own region as the marching shapes (so bricks, lists and masks are shared), one
// thread per quad of a plane.  `Geo` is the brick geometry the work lists are expressed in.
template <int T_, int LX_>
struct RShape {
    using Geo = RShape;
    static constexpr int T = T_, LX = LX_;
    static constexpr bool kHalf = false;           // a work item is a whole brick
    static constexpr int kTileX = 4 * LX_, kTileY = 2048 / kTileX;
    static constexpr int kThreads = LX_ * kTileY;  // 512
    static constexpr int kOutX = kTileX - 2 * kHaloX, kOutY = kTileY - 2 * T_;
    static constexpr int kPlane = kTileX * kTileY;
    static constexpr int kBz = 8;                  // planes per brick at most
    static constexpr int kPlanes = kBz + 2 * T_;   // planes of a window
    static constexpr size_t kFloats = (size_t)2 * kPlanes * kPlane;
    static constexpr size_t kBytes = kFloats * sizeof(float) + 64;
    static_assert(kThreads == 512, "one thread per quad of a 2048-cell plane");
    static_assert(kBytes + 1024 <= 233472, "shared memory budget");
};

// FOUR sweeps per pass on HALF bricks (the tail of the solve: half as many passes).  The bricks stay those of the T = 2
// schedule (56 x 28 x 8 own cells: lists, masks and copies are shared); a work item is the lower or upper 14 rows of one,
// so that the window with its halo of four — 64 x 22 x 16 cells of pressure and of right-hand side, 2 x 88 KB — fits
// beside nothing else in shared memory.  The two halves of a brick run in different CTAs; whichever finishes second
// decides whether the brick is relaxed again or copied.  352 threads, one quad column of 16 planes each.
struct RShapeHalf4 {
    struct Geo {  // the brick geometry of the T = 2 shapes with 64-wide tiles
        static constexpr int T = 2, kOutX = 56, kOutY = 28, kThreads = 352;
    };
    static constexpr int T = 4, LX = 16;
    static constexpr bool kHalf = true;
    static constexpr int kTileX = 64, kTileY = Geo::kOutY / 2 + 2 * T;  // 22
    static constexpr int kThreads = LX * kTileY;                         // 352
    static constexpr int kPlane = kTileX * kTileY;
    static constexpr int kBz = 8;
    static constexpr int kPlanes = kBz + 2 * T;                          // 16: one flag nibble per plane in 64 bits
    static constexpr size_t kFloats = (size_t)2 * kPlanes * kPlane;
    static constexpr size_t kBytes = kFloats * sizeof(float) + 64;
    static_assert(kThreads == Geo::kThreads && kThreads % 32 == 0, "whole warps");
    static_assert(kBytes + 1024 <= 233472, "shared memory budget");
};

// L2 loads of what other CTAs of this launch (or a neighbour rank) wrote: the L1 is not coherent.
__device__ __forceinline__ int ld_l2(const int* p) { return __ldcg(p); }
__device__ __forceinline__ unsigned long long ld_l2(const unsigned long long* p) { return __ldcg(p); }

template <class S, bool FUSED>
__global__ void __launch_bounds__(S::kThreads, 1)
jacobi_resident_kernel(const __grid_constant__ CUtensorMap map_p0, const __grid_constant__ CUtensorMap map_p1,
                       const __grid_constant__ CUtensorMap map_rhs, const FrameParams* __restrict__ frame,
                       StepState* state, float* p0, float* p1, unsigned char* m0, unsigned char* m1,
                       const __grid_constant__ WorkLists W, const __grid_constant__ PassParams P,
                       const __grid_constant__ PeerView pv, const __grid_constant__ JacobiPeers peers) {
    constexpr int T = S::T, LX = S::LX, kTileX = S::kTileX, kTileY = S::kTileY, kPlane = S::kPlane, NP = S::kPlanes;
    constexpr unsigned kFull = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31;
    const int li = tid % LX, row = tid / LX;
#ifdef FXB_TIMING
    // debug build: cycle stamps of the first CTA into StepState::dbg (tools/timing_probe.py)
    int dbg_n = 0;
    int dbg_pass = -1;
#define FXB_STAMP() do { if (tid == 0 && blockIdx.x == 0 && dbg_pass == FXB_TIMING && dbg_n < 120) state->dbg[dbg_n++] = clock64(); } while (0)
#else
#define FXB_STAMP() do {} while (0)
#endif

    // Programmatic dependent launch: the next pass's CTAs may be scheduled as soon as every CTA of this one has
    // started (they take the SMs of the CTAs that have nothing to do and exit at once); they wait below, before they
    // touch anything this pass writes.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    extern __shared__ __align__(1024) float sm[];                  // TMA destinations need 128-byte alignment
    float* sm_p = sm;                                              // [NP][kPlane] pressure window: level 0, then level 1 ..
    float* sm_rhs = sm + NP * kPlane;                              // [NP][kPlane] right-hand side window
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + S::kFloats);  // one barrier, one phase per window
    __shared__ unsigned s_cnt[T];
    __shared__ int s_copied;
    __shared__ unsigned s_todo;
    if (tid == 0) {
        if ((smem_u32(sm) & 127u) != 0u) __trap();
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_copied = 0;
    }
    __syncthreads();
    // everything the previous kernel wrote (pressure, masks, lists, counters) is complete and visible from here on
    asm volatile("griddepcontrol.wait;" ::: "memory");

    const float dt = frame->dt;
    const unsigned long long epoch = frame->epoch_base;
    const int p_cur = state->p_cur;  // constant during the solve (jacobi_flip_kernel flips it afterwards)
    const int layer = P.ntx * P.nty;
    const int nxb = P.pitch >> 3;
    const size_t mplane = (size_t)P.ny * nxb;
    const float eps = P.early_exit ? kEps : -1.0f;
    const bool cut_x = (P.nx & 3) != 0;  // the grid's x face cuts through a quad
    const int off0 = row * kTileX + 4 * li;
    unsigned round = 0;        // windows consumed so far: the mbarrier's phase
    unsigned n_done = 0;       // own bricks finished by this CTA
    unsigned n_copied = 0;     // thread 0 of CTA 0: bricks copied from the copy list

    {
        const int pass = P.pass;
#ifdef FXB_TIMING
        dbg_pass = pass;
#endif
        FXB_STAMP();  // 0: pass start
        const int s0 = P.s0;  // sweeps completed before this pass (the schedule may mix passes of 2 and of 4 sweeps)
        const unsigned long long need = epoch + 2ull + (unsigned long long)pass;  // this pass's event number (PassParams::event)
        // independent loads first (one round trip instead of a chain), then the decisions
        const unsigned long long still_prev = pass > 0 ? ld_l2(&state->active_after[s0 - 1]) : 1ull;
        const int n_relax = pass > 0 ? ld_l2(&W.relax_count[pass]) : P.first_count;
        const int n_copy = pass > 0 ? ld_l2(&W.copy_count[pass]) : 0;
        const int* list_in = W.relax[pass & 1];
        // speculative: the list entries of this CTA's first two work items (garbage beyond the list, then unused; the
        // lists are padded)
        constexpr int kItemShift = S::kHalf ? 1 : 0;  // two work items per listed brick when they are half bricks
        int listed = pass > 0 ? ld_l2(&list_in[blockIdx.x >> kItemShift]) : (int)blockIdx.x;
        int listed_next = pass > 0 ? ld_l2(&list_in[(blockIdx.x + gridDim.x) >> kItemShift]) : (int)(blockIdx.x + gridDim.x);
        // Fused halos: when the last CTA is done, this kernel's event is published to the neighbours — on every path.
        bool pushed = false;  // this CTA stored into a neighbour rank (uniform)
        auto finish = [&]() {
            if constexpr (!FUSED) return;
            __syncthreads();
            if (tid == 0) {
                // what this CTA pushed over NVLink has arrived before its arrival is counted (the publishing CTA fences
                // again before the event); a CTA that pushed nothing has nothing to wait for
                if (pushed) __threadfence_system();
                if (atomicAdd(&state->done_ctas, 1) == (int)gridDim.x - 1) {
                    state->done_ctas = 0;
                    peer_publish(pv, need + 1ull);
#ifdef FXB_TIMING
                    {  // debug build: when each pass of the frame ended on this rank (tools/mgpu_probe.py)
                        unsigned long long now;
                        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
                        state->dbg[64 + (pass < 40 ? pass : 39)] = (long long)now;
                    }
#endif
                }
            }
        };
        if (!(0.0f < dt) || (pass > 0 && !P.run_all && still_prev == 0ull)) {  // nothing left in this frame
            finish();
            return;
        }
        const int levels = min(T, P.levels_total - s0);
        FXB_STAMP();  // 1: the pass's loads have arrived

        const int sel = (p_cur + pass) & 1;
        const CUtensorMap* map_in = sel ? &map_p1 : &map_p0;
        const float* p_in = sel ? p1 : p0;
        float* p_out = sel ? p0 : p1;
        const unsigned char* m_in = (pass & 1) ? m1 : m0;
        unsigned char* m_out = (pass & 1) ? m0 : m1;
        const int pi = sel ? 0 : 1, mi = (pass & 1) ? 0 : 1;  // the neighbours' copies of the two output buffers

        // Fused halos: a brick of the lowest / highest layer reads halo planes the neighbour's previous kernel wrote and
        // stores into the neighbour's halo planes that its previous kernel still read: wait for that pass (once per side).
        bool waited_lo = !FUSED || !pv.has_lo, waited_hi = !FUSED || !pv.has_hi;
        auto peer_sync = [&](const bool lo, const bool hi) -> bool {  // uniform; thread 0 polls; true: a wait happened
            if constexpr (!FUSED) return false;
            const bool wl = lo && !waited_lo, wh = hi && !waited_hi;
            if (!(wl || wh)) return false;
            if (tid == 0) peer_wait(pv, need, wl, wh);
            waited_lo |= wl;
            waited_hi |= wh;
            return true;
        };
        auto brick_faces = [&](const int brick, bool& lo, bool& hi) {  // called for every brick this CTA copies or relaxes
            lo = FUSED && pv.has_lo && brick < layer;
            hi = FUSED && pv.has_hi && brick >= layer * (P.nzc - 1);
            pushed |= lo || hi;
        };

        // The frozen bricks of the previous pass: one copy each into the other pressure buffer (as in jacobi_fused.cu,
        // including the first pass's special case: this kernel also runs pass 1).
        if (pass == 1) {
            const int nbricks = layer * P.nzc;
            auto rot = [&](const int i) { return !FUSED ? i : (i + layer < nbricks ? i + layer : i + layer - nbricks); };
            for (int b0 = blockIdx.x; b0 < nbricks; b0 += gridDim.x * 32) {
                // 32 candidate bricks of this CTA at a time: one round trip for their flags
                // (fused halos: candidates rotated by one layer, the face layers last, as in the first pass)
                const int bi = b0 + (tid & 31) * gridDim.x;
                const int b = rot(bi);
                const int f = (tid < 32 && bi < nbricks) ? ld_l2(&W.brick_flag[b]) : 0;
                const bool edge = bi < nbricks && ((P.keep_lo && b < layer) || (P.keep_hi && b >= nbricks - layer));
                const bool want = (f & 1) && ((f & 2) || P.copy_all || edge);
                const unsigned todo = __ballot_sync(kFull, tid < 32 && want);
                if (tid == 0) s_todo = todo;
                __syncthreads();
                unsigned m = s_todo;
                while (m) {
                    const int j = __ffs(m) - 1;
                    m &= m - 1;
                    const int brick = rot(b0 + j * gridDim.x);
                    bool lo, hi;
                    brick_faces(brick, lo, hi);
                    if (peer_sync(lo, hi)) __syncthreads();
                    copy_frozen_brick<typename S::Geo, FUSED>(p_in, p_out, m_out, P, brick, pv, peers, pi, mi);
                    if (tid == 0) ++s_copied;
                }
                __syncthreads();
            }
        } else if (n_copy > 0) {
            const int* copy_list = W.copy[pass & 1];
            // handed out from the LAST CTA down: the first CTAs are the ones that relax bricks
            for (int w = gridDim.x - 1 - blockIdx.x; w < n_copy; w += gridDim.x) {
                const int brick = ld_l2(&copy_list[w]);
                bool lo, hi;
                brick_faces(brick, lo, hi);
                if (peer_sync(lo, hi)) __syncthreads();
                copy_frozen_brick<typename S::Geo, FUSED>(p_in, p_out, m_out, P, brick, pv, peers, pi, mi);
            }
            if (blockIdx.x == 0) n_copied += (unsigned)n_copy;
        }
        FXB_STAMP();  // 2: copies handed out

        const int n_ext = layer * ((P.ext_lo + P.bz - 1) / P.bz + (P.ext_hi + P.bz - 1) / P.bz);
        const int n_items = n_relax << kItemShift;
        const int n_work = n_items + n_ext;

        // A work item: a listed brick, or (kHalf) its lower / upper half: items 2i and 2i + 1 are the halves of entry i.
        constexpr bool HALF = S::kHalf;
        auto item_of = [&](const int w, const int entry) -> Item {
            if (w < n_items) {
                int brick = entry;
                if (pass == 0) {  // the bricks (first_brick + w) mod bricks (see PassParams)
                    brick = w + P.first_brick;
                    if (brick >= layer * P.nzc) brick -= layer * P.nzc;
                }
                Item it = own_item<typename S::Geo>(P, brick);
                if constexpr (HALF) it.gy0 += S::Geo::T - T + (w & 1) * (S::Geo::kOutY / 2);  // halo of T rows, 14 own rows
                return it;
            }
            return ext_item<typename S::Geo>(P, w - n_items);
        };
        // the window of a work item: before anything of a face brick is staged, the neighbour's data must be there
        auto stage = [&](const Item& it) {
            if (FUSED && it.brick >= 0) {
                bool lo, hi;
                brick_faces(it.brick, lo, hi);
                if (peer_sync(lo, hi)) __syncthreads();  // also orders the other threads' flag loads after the wait
            }
            if (tid == 0) {
                // the previous window was read and written through the generic proxy (all of it before the CTA barrier
                // the caller has just passed); the copy engine writes through the async proxy
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, 2u * NP * kPlane * 4u);
                tma_load_3d(sm_p, map_in, it.gx0, it.gy0, it.zs - T, bar);
                tma_load_3d(sm_rhs, &map_rhs, it.gx0, it.gy0, it.zs - T, bar);
            }
        };

        unsigned tot[T + 1];  // cells of this CTA's own bricks still active after each level
#pragma unroll
        for (int l = 0; l <= T; ++l) tot[l] = 0;
        // list append of the previous brick (thread 0): the atomic's result is consumed one brick later
        int pend_brick = -1, pend_slot = 0;
        int* pend_list = nullptr;

        Item it = item_of((int)blockIdx.x < n_work ? (int)blockIdx.x : 0, listed);
        if ((int)blockIdx.x < n_work) stage(it);
        FXB_STAMP();  // 3: first window requested

        for (int work = blockIdx.x; work < n_work; work += gridDim.x, ++round) {
            // the next work item (its list entry was fetched one round ago)
            const int work_next = work + gridDim.x;
            listed = listed_next;
            {
                const int w2 = work_next + gridDim.x;
                listed_next = (pass > 0 && w2 < n_items) ? ld_l2(&list_in[w2 >> kItemShift]) : w2;
            }
            const int zs = it.zs, ze = it.ze, z0 = zs - T;
            const int zl0 = max(z0, 0), zl1 = min(ze + T, P.nz_alloc);

            // ---- per-thread geometry of this tile ----
            const int gx = it.gx0 + 4 * li, gy = it.gy0 + row;
            const int xe = (cut_x && gx < P.nx && gx + 4 > P.nx) ? P.nx - gx : 0;  // cells of a quad cut by the x face
            const unsigned qmask = gx >= 0 && gx < P.nx ? (xe ? (1u << xe) - 1u : 0xFu) : 0u;
            const unsigned dom = (gy >= 0 && gy < P.ny) ? qmask : 0u;                                     // cells inside the grid
            const unsigned own = (li >= 1 && li <= LX - 2 && row >= T && row < kTileY - T) ? dom : 0u;  // this brick's output
            const bool mine = own & 1u;
            const bool clamp_u = row == 0 || gy <= 0;
            const bool clamp_d = row == kTileY - 1 || gy >= P.ny - 1;
            const int off_up = clamp_u ? off0 : off0 - kTileX;
            const int off_dn = clamp_d ? off0 : off0 + kTileX;
            const bool clamp_l = li == 0 || gx == 0;
            const bool clamp_r = li == LX - 1 || gx + 4 >= P.nx;
            auto fix_ghosts = [&](float4& v) {  // cells of a cut quad beyond the face mirror the last inside cell
                if (cut_x) {
                    if (xe == 1) v.y = v.x;
                    if (xe == 2) v.z = v.y;
                    if (xe == 3) v.w = v.z;
                }
            };

            // ---- freeze flags ----
            // fl: level-0 window (the previous pass's output mask), nibble j = plane z0 + j.
            // keep: own planes whose cells were all frozen two passes ago already: the output buffer (written two passes
            // ago) and its mask hold their final value, nothing is stored.  (L2 loads: other CTAs / ranks wrote them.)
            unsigned long long fl = 0ull;
            unsigned stale = 0;  // bit j: this thread's quad of own plane j must be stored
            {
                unsigned raw[NP], old[NP];
                const size_t mrow0 = (size_t)max(gy, 0) * nxb + (max(gx, 0) >> 3);
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const int z = z0 + j;
                    raw[j] = 0u;
                    old[j] = 0xffu;
                    if (pass > 0 && dom && z >= zl0 && z < zl1) raw[j] = __ldcg(m_in + (size_t)z * mplane + mrow0);
                    if (j >= T && j < NP - T && pass > 1 && mine && z < ze) old[j] = __ldcg(m_out + (size_t)z * mplane + mrow0);
                }
                const int nib_shift = gx & 4;
#pragma unroll
                for (int j = 0; j < NP; ++j) {
                    const int z = z0 + j;
                    const unsigned nib = pass == 0 ? ((z >= zl0 && z < zl1) ? dom : 0u) : ((raw[j] >> nib_shift) & dom);
                    fl |= (unsigned long long)nib << (4 * j);
                    if (j >= T && j < NP - T && mine && z < ze && ((old[j] >> nib_shift) & 0xFu) != 0u) stale |= 1u << j;
                }
            }
            // a warp without an active cell in the window and without a stale output quad has nothing to do
            constexpr unsigned long long kLevel1Planes = ((1ull << (4 * (NP - 2))) - 1ull) << 4;
            const bool warp_busy = __any_sync(kFull, (fl & kLevel1Planes) != 0ull || stale != 0u);

            // ---- the window has landed: this thread's quad column into registers ----
            FXB_STAMP();  // brick + 0: flag bytes requested
            mbar_wait(bar, round & 1u);
            FXB_STAMP();  // brick + 1: window landed
            // (planes 1 .. NP-2: the outermost two are only ever the z neighbour of the first level and stay in shared memory)
            float4 col[NP];
            if (warp_busy) {
#pragma unroll
                for (int j = 1; j < NP - 1; ++j) {
                    col[j] = *reinterpret_cast<const float4*>(sm_p + j * kPlane + off0);
                    fix_ghosts(col[j]);
                }
            }

            unsigned alive = 0;  // an own cell of this brick is still active after the pass's last sweep
            auto level = [&](auto lc) {
                constexpr int l = decltype(lc)::value;
                const int lo_z = max(zs - (T - l), 0), hi_z = min(ze + (T - l), P.nz_alloc);  // planes level l must produce
                unsigned changed = 0;  // planes this warp relaxed (uniform per warp)
                if (warp_busy) {
                    float4 prev;  // level l-1 of the plane below (this level leaves plane l-1 alone)
                    if constexpr (l == 1) {
                        prev = *reinterpret_cast<const float4*>(sm_p + off0);
                        fix_ghosts(prev);
                    } else {
                        prev = col[l - 1];
                    }
#pragma unroll
                    for (int j = l; j <= NP - 1 - l; ++j) {
                        const int z = z0 + j;
                        const float4 c = col[j];
                        const bool run = z >= lo_z && z < hi_z;  // uniform in the CTA
                        const unsigned act = (l <= levels && run) ? (unsigned)(fl >> (4 * j)) & 0xFu : 0u;
                        unsigned st = 0;
                        if (__any_sync(kFull, act != 0u)) {       // uniform in the warp
                            const float* nb = sm_p + j * kPlane;  // level l-1 of this plane: the rows above / below
                            const float4 up = *reinterpret_cast<const float4*>(nb + off_up);
                            const float4 dn = *reinterpret_cast<const float4*>(nb + off_dn);
                            const float4 rhs = *reinterpret_cast<const float4*>(sm_rhs + j * kPlane + off0);
                            // clamp rule at the grid's z faces: the missing neighbour plane is the centre plane itself
                            const float4 zlo = z == P.z_face_lo ? c : prev;
                            float4 above;
                            if constexpr (l == 1) {
                                if (j == NP - 2) {
                                    above = *reinterpret_cast<const float4*>(sm_p + (NP - 1) * kPlane + off0);
                                    fix_ghosts(above);
                                } else {
                                    above = col[j + 1 < NP - 1 ? j + 1 : j];
                                }
                            } else {
                                above = col[j + 1];
                            }
                            const float4 zhi = z + 1 == P.z_face_hi ? c : above;
                            float left = __shfl_up_sync(kFull, c.w, 1, LX), right = __shfl_down_sync(kFull, c.x, 1, LX);
                            if (clamp_l) left = c.x;
                            if (clamp_r) right = c.w;
                            float4 out;
                            st = relax_quad(c, zlo, zhi, up, dn, left, right, rhs, act, eps, out);
                            fix_ghosts(out);
                            col[j] = out;
                            changed |= 1u << j;
                            if (z >= zs && z < ze && it.brick >= 0) {  // halo bricks are counted by their owner
                                tot[l] += __popc(st & own);
                                if (l == levels) alive |= st & own;
                            }
                        }
                        if (run) fl = (fl & ~(0xFull << (4 * j))) | ((unsigned long long)st << (4 * j));
                        prev = c;
                    }
                }
                if constexpr (l < T) {
                    // level l replaces level l-1 in shared memory for the rows above / below (unchanged planes are equal)
                    __syncthreads();
                    FXB_STAMP();  // brick + 2: level l done by everybody
#pragma unroll
                    for (int j = l; j <= NP - 1 - l; ++j)
                        if (changed & (1u << j)) *reinterpret_cast<float4*>(sm_p + j * kPlane + off0) = col[j];
                    __syncthreads();
                    FXB_STAMP();  // brick + 3: written back
                }
            };
            level(std::integral_constant<int, 1>{});
            if constexpr (T >= 2) level(std::integral_constant<int, 2>{});
            if constexpr (T >= 3) level(std::integral_constant<int, 3>{});
            if constexpr (T >= 4) level(std::integral_constant<int, 4>{});

            // everybody is done with the window; the next one is staged while this brick's output goes out
            const int any_alive = __syncthreads_or(alive != 0u);
            FXB_STAMP();  // brick + 4: last level done by everybody
            const Item cur = it;
            if (work_next < n_work) {
                it = item_of(work_next, listed);
                stage(it);
            }

            // ---- level T of the own planes: the pass's output ----
            if (warp_busy) {
#pragma unroll
                for (int j = T; j <= NP - 1 - T; ++j) {
                    const int z = z0 + j;
                    if (z < ze) {  // uniform (z >= zs always)
                        // bit-packed freeze flags: two quads (8 cells) per byte, written by the odd lane
                        const unsigned nib = (unsigned)(fl >> (4 * j)) & 0xFu;
                        const unsigned hi = __shfl_down_sync(kFull, nib, 1, LX);
                        const bool hi_stale = __shfl_down_sync(kFull, (stale >> j) & 1u, 1, LX);
                        const bool st_p = mine && ((stale >> j) & 1u);
                        const bool st_m = mine && (li & 1) && (((stale >> j) & 1u) || hi_stale);
                        const size_t at = ((size_t)z * P.ny + gy) * P.pitch + gx;
                        const size_t mat = ((size_t)z * P.ny + gy) * nxb + (gx >> 3);
                        const unsigned char byte = (unsigned char)(nib | (hi << 4));
                        if (st_p) *reinterpret_cast<float4*>(p_out + at) = col[j];
                        if (st_m) m_out[mat] = byte;
                        if constexpr (FUSED) {
                            const bool to_lo = cur.brick >= 0 && pushes_lo(pv, P, z), to_hi = cur.brick >= 0 && pushes_hi(pv, P, z);
                            if (to_lo || to_hi) {  // the same stores into the neighbour's halo planes
                                const long long plane_f = (long long)P.ny * P.pitch, plane_b = (long long)P.ny * nxb;
                                if (st_p && to_lo) *reinterpret_cast<float4*>(peers.p[0][pi] + (long long)at + pv.dz_lo * plane_f) = col[j];
                                if (st_p && to_hi) *reinterpret_cast<float4*>(peers.p[1][pi] + (long long)at + pv.dz_hi * plane_f) = col[j];
                                if (st_m && to_lo) peers.m[0][mi][(long long)mat + pv.dz_lo * plane_b] = byte;
                                if (st_m && to_hi) peers.m[1][mi][(long long)mat + pv.dz_hi * plane_b] = byte;
                            }
                        }
                    }
                }
            }

            // ---- brick state: still active -> relax again next pass; just frozen -> one copy next pass ----
            if (cur.brick >= 0) {
                if (pass == 0) {
                    // first pass: a frozen brick is only flagged; a still-active one flags the 26 bricks around it
                    if (!any_alive) {
                        if (tid == 0) atomicOr(&W.brick_flag[cur.brick], 1);
                    } else if (tid < 27) {
                        const int tx = cur.brick % P.ntx, ty = (cur.brick / P.ntx) % P.nty, tz = cur.brick / (P.ntx * P.nty);
                        const int nx_ = tx + tid % 3 - 1, ny_ = ty + (tid / 3) % 3 - 1, nz_ = tz + tid / 9 - 1;
                        if (tid != 13 && nx_ >= 0 && nx_ < P.ntx && ny_ >= 0 && ny_ < P.nty && nz_ >= 0 && nz_ < P.nzc)
                            atomicOr(&W.brick_flag[(nz_ * P.nty + ny_) * P.ntx + nx_], 2);
                    }
                }
                if (tid == 0 && (any_alive || pass > 0)) {
                    // Half bricks: the two halves meet in the brick's state word (zero between passes): the half that
                    // arrives second knows whether either is still active, lists the brick and clears the word.
                    bool mine_to_list = true, brick_alive = any_alive != 0;
                    if constexpr (HALF) {
                        const int old = atomicAdd(&W.brick_half[cur.brick], 1 + (any_alive ? 0x10000 : 0));
                        mine_to_list = (old & 0xffff) == 1;
                        brick_alive = brick_alive || (old >> 16) != 0;
                        if (mine_to_list) W.brick_half[cur.brick] = 0;
                    }
                    if (mine_to_list) {
                        if (pend_brick >= 0) pend_list[pend_slot] = pend_brick;
                        pend_brick = cur.brick;
                        if (brick_alive) {
                            pend_list = W.relax[(pass + 1) & 1];
                            pend_slot = atomicAdd(&W.relax_count[pass + 1], 1);
                        } else {
                            pend_list = W.copy[(pass + 1) & 1];
                            pend_slot = atomicAdd(&W.copy_count[pass + 1], 1);
                        }
                        ++n_done;
                    }
                } else if (!HALF || pass == 0) {
                    if (tid == 0) ++n_done;
                }
            }
            FXB_STAMP();  // brick + 5: output stores issued, brick listed
        }
        if (tid == 0 && pend_brick >= 0) pend_list[pend_slot] = pend_brick;

        // ---- per-level active counts of this CTA's bricks -> global counters ----
        if (__syncthreads_or(tot[1] != 0u)) {  // (tot[l] is non-increasing in l)
            if (tid < T) s_cnt[tid] = 0;
            __syncthreads();
#pragma unroll
            for (int l = 1; l <= T; ++l) {
                unsigned v = tot[l];
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(kFull, v, o);
                if (lane == 0 && v) atomicAdd(&s_cnt[l - 1], v);
            }
            __syncthreads();
            if (tid < T && tid < levels) {
                const unsigned v = s_cnt[tid];
                if (v) atomicAdd(&state->active_after[s0 + tid], (unsigned long long)v);
            }
        }
        FXB_STAMP();  // counters out

        if (tid == 0) {  // (n_done is thread 0's count)
            if (n_done) atomicAdd(&state->bricks_processed, (unsigned long long)n_done);
            const unsigned c = n_copied + (unsigned)s_copied;
            if (c) atomicAdd(&state->bricks_copied, (unsigned long long)c);
        }
        finish();
    }
}

template <class S>
cudaError_t launch_resident_shape(const FusedJacobi& J, const Domain& d, const FrameParams* frame, StepState* state,
                                  int pass, int iters, int early_exit, bool run_all, int ext_lo, int ext_hi,
                                  int first_brick, int first_count, const PeerView& pv, cudaStream_t stream) {
    // the opt-in above the 48 KB default is per device: set it whenever the device changes (cheap, idempotent)
    static int attr_device = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_device != dev) {
        cudaError_t e = cudaFuncSetAttribute(jacobi_resident_kernel<S, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)S::kBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(jacobi_resident_kernel<S, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kBytes);
        if (e != cudaSuccess) return e;
        attr_device = dev;
    }
    PassParams P = make_pass_params(J, d, pass, iters, early_exit, run_all, ext_lo, ext_hi);
    if (pass == 0 && first_count >= 0) {
        P.first_brick = first_brick;
        P.first_count = first_count;
    }
    const JacobiPeers peers = make_jacobi_peers(J);
    const WorkLists W = make_work_lists(J);
    const int nbricks = pass == 0 ? P.first_count : J.ntx * J.nty * J.nzc;
    const int grid = nbricks < J.num_sms ? (nbricks > 0 ? nbricks : 1) : J.num_sms;  // persistent, one CTA per SM
    const CUtensorMap& mp0 = *reinterpret_cast<const CUtensorMap*>(S::kHalf ? J.map4_p[0] : J.map3_p[0]);
    const CUtensorMap& mp1 = *reinterpret_cast<const CUtensorMap*>(S::kHalf ? J.map4_p[1] : J.map3_p[1]);
    const CUtensorMap& mr = *reinterpret_cast<const CUtensorMap*>(S::kHalf ? J.map4_rhs : J.map3_rhs);
    // programmatic dependent launch: this kernel may start while its predecessor in the stream drains (it waits,
    // griddepcontrol.wait, before it reads anything); the predecessor is the previous pass — resident kernels release
    // their dependents at once, any other kernel when it completes
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)S::kThreads);
    cfg.dynamicSmemBytes = S::kBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = J.pdl ? 1 : 0;
    if (pv.has_lo || pv.has_hi)
        return cudaLaunchKernelEx(&cfg, jacobi_resident_kernel<S, true>, mp0, mp1, mr, frame, state, J.p[0], J.p[1], J.mask[0],
                                  J.mask[1], W, P, pv, peers);
    return cudaLaunchKernelEx(&cfg, jacobi_resident_kernel<S, false>, mp0, mp1, mr, frame, state, J.p[0], J.p[1], J.mask[0],
                              J.mask[1], W, P, pv, peers);
}

}  // namespace

bool resident_jacobi_supported(const FusedJacobi& J) { return J.T >= 1 && J.T <= 2 && J.bz <= 8; }

int resident_jacobi_window_planes(const FusedJacobi& J) { return 8 + 2 * J.T; }

cudaError_t launch_jacobi_pass_resident(const FusedJacobi& J, const Domain& d, const FrameParams* frame, StepState* state,
                                        int pass, int iters, int early_exit, bool run_all, int ext_lo, int ext_hi,
                                        int first_brick, int first_count, const PeerView& pv, cudaStream_t stream) {
#define FXB_LAUNCH(S) return launch_resident_shape<S>(J, d, frame, state, pass, iters, early_exit, run_all, ext_lo, ext_hi, first_brick, first_count, pv, stream)
    if (pass >= J.tail_from) FXB_LAUNCH(RShapeHalf4);
    using N1 = RShape<1, 16>;
    using W1 = RShape<1, 32>;
    using N2 = RShape<2, 16>;
    using W2 = RShape<2, 32>;
    switch (J.T) {
        case 1: if (J.narrow) FXB_LAUNCH(N1); FXB_LAUNCH(W1);
        case 2: if (J.narrow) FXB_LAUNCH(N2); FXB_LAUNCH(W2);
    }
#undef FXB_LAUNCH
    return cudaErrorInvalidValue;
}

}  // namespace fxb
