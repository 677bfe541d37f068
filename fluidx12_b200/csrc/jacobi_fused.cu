// jacobi_fused.cu — T Jacobi sweeps of the pressure solve fused into one HBM pass (sm_100a).
//
// Replaces the relaxation loop of FluidX12/Content/Shaders/CSPoisson.hlsli:8-26 as called from
// CSProject3D.hlsl:93 (dispatch Fluid.cpp:394-408), under the deterministic restatement of SURVEY.md
// App. A.3: synchronous Jacobi, per-cell freeze once |x - x0| < 0.001, at most ITER sweeps.
//
// Scheme: z-marching 2.5-D blocking with temporal fusion.  A CTA owns a brick of kOutX x kOutY x bz output cells and
// streams the xy tile (kTileX x kTileY cells: halo of one quad in x, T rows in y) plane by plane along z.  Level l
// (= number of sweeps applied) of plane k-l is produced in iteration k, so the T sweeps advance in lock-step, each one
// plane behind the previous:
//   * level-0 pressure planes and right-hand-side planes are staged into shared-memory rings by TMA
//     (cp.async.bulk.tensor.3d + mbarrier).  The bundles of all bricks a persistent CTA works on form ONE stream that
//     runs DEPTH bundles ahead of the consumer, also across brick boundaries, so a new brick never starts cold;
//     barriers are indexed by the running bundle count and never re-initialised;
//   * every thread keeps its column (2 rows x 4 cells) of the previous and the centre plane of every level in
//     registers (the z queue), so the z neighbours never touch memory;
//   * x neighbours come from warp shuffles (a tile row is LX lanes x 4 cells), y neighbours from the thread's other
//     row or from the plane each level publishes in shared memory (read one iteration later: one barrier per
//     iteration);
//   * the reference's clamp-to-edge neighbour rule (CSProject3D.hlsl:76-83) is applied by index (x, y) or by reusing
//     the centre value (z), never by TMA fill; rows / cells of a tile just outside the grid mirror the adjacent
//     inside row / cell after every update, so the in-register neighbours obey the rule as well (any nx, ny);
//   * the per-cell freeze flags travel with the values (4 bits per quad per level); a warp whose cells are all
//     frozen at a level skips that level's arithmetic; flags persist between passes in a bit-packed array
//     (1 bit per cell, 0.25 B/voxel/pass of traffic);
//   * persistent CTAs take bricks from device-side work lists; a brick whose cells are all frozen is copied once to
//     the other pressure buffer and skipped for the rest of the frame (final in both ping-pong buffers).
// Rows are `pitch` floats apart (a multiple of 8 >= nx: the "pitched device buffers" of the north star), so any grid
// width takes this path.
// Algorithmic traffic per relaxed cell per pass: p in 4 + rhs in 4 + p out 4 (+ 2/8 mask) bytes; 8 per copied cell.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "jacobi_common.cuh"
#include "kernels.h"

namespace fxb {

namespace {

// One fused pass (see the file header).  FUSED: the multi-GPU instantiation with fused halos (common.cuh PeerView);
// the single-GPU one carries none of that code.
template <class S, bool FUSED, bool FIRST>
__global__ void __launch_bounds__(S::kThreads, S::kCtasPerSm)
jacobi_pass_kernel(const __grid_constant__ CUtensorMap map_p0, const __grid_constant__ CUtensorMap map_p1,
                   const __grid_constant__ CUtensorMap map_rhs, const FrameParams* __restrict__ frame,
                   StepState* __restrict__ state, float* p0, float* p1, unsigned char* m0, unsigned char* m1,
                   const __grid_constant__ WorkLists W, const __grid_constant__ PassParams P,
                   const __grid_constant__ PeerView pv, const __grid_constant__ JacobiPeers peers) {
    constexpr int T = S::T, LX = S::LX, kRows = S::kRows, kTileX = S::kTileX, kTileY = S::kTileY, kPlane = S::kPlane;
    constexpr unsigned kFull = 0xffffffffu;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int li = lane & (LX - 1), sub = lane / LX;
    // (A ninth warp that only feeds the TMA stream was measured on B200: it takes the ~300 cycles a bundle costs its
    // issuing thread off the relaxing warps' path, but 9 warps per CTA split unevenly over the 4 SM sub-partitions and
    // cap the kernel at 96 registers — spills, 2x slower.  Thread 0 issues.)
    constexpr bool producer = false;

    // independent loads first (one round trip instead of a chain), then the decisions
    // FIRST: the instantiation for the first pass of a frame — with the brick-resident kernel taking every later pass it
    // is the only one the default schedule launches, and everything that depends on `pass` folds away in it (no lists,
    // no flag bytes to fetch, nothing to copy).
    const int pass = FIRST ? 0 : P.pass, s0 = FIRST ? 0 : P.s0;
    const float dt = frame->dt;
    const unsigned long long need = frame->epoch_base + (unsigned long long)P.event;
    const int p_cur = state->p_cur;
    const unsigned long long still_prev = pass > 0 ? state->active_after[s0 - 1] : 1ull;
    const int n_relax = pass > 0 ? W.relax_count[pass] : P.first_count;
    const int n_copy = pass > 0 ? W.copy_count[pass] : 0;
    const int* __restrict__ list_in = W.relax[pass & 1];
    // speculative: the first two list entries of this CTA (garbage beyond n_relax, then unused)
    int pre_a = pass > 0 ? list_in[blockIdx.x] : (int)blockIdx.x;
    int pre_b = pass > 0 ? list_in[blockIdx.x + gridDim.x] : (int)(blockIdx.x + gridDim.x);
    constexpr bool fused_halos = FUSED;
#ifdef FXB_TIMING
    // debug build: cycle stamps of the first CTA's first brick into StepState::dbg (tools/timing_probe.py)
    int dbg_n = 0;
#define FXB_STAMP() do { if (tid == 0 && blockIdx.x == 0 && pass == FXB_TIMING && dbg_n < 120) state->dbg[dbg_n++] = clock64(); } while (0)
#else
#define FXB_STAMP() do {} while (0)
#endif
    FXB_STAMP();
    // Fused halos: when the last CTA is done, this kernel's event is published to the neighbours — on every path.
    auto finish = [&]() {
        if constexpr (!FUSED) return;
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
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
    if (!(0.0f < dt) || (pass > 0 && !P.run_all && still_prev == 0ull)) {
        finish();
        return;
    }
    const int levels = min(T, P.levels_total - s0);

    const int sel = (p_cur + pass) & 1;
    const CUtensorMap* map_in = sel ? &map_p1 : &map_p0;
    const float* p_in = sel ? p1 : p0;
    float* p_out = sel ? p0 : p1;
    const unsigned char* m_in = (pass & 1) ? m1 : m0;
    unsigned char* m_out = (pass & 1) ? m0 : m1;
    const int pi = sel ? 0 : 1, mi = (pass & 1) ? 0 : 1;  // the neighbours' copies of the two output buffers

    extern __shared__ __align__(1024) float sm[];                   // TMA destinations need 128-byte alignment
    float* sm_p = sm;                                               // [kPSlots][kPlane]  level-0 planes (TMA)
    float* sm_rhs = sm_p + S::kPSlots * kPlane;                     // [kRSlots][kPlane]  rhs planes (TMA)
    float* sm_pub = sm_rhs + S::kRSlots * kPlane;                   // [T-1][2][kPlane]   levels 1..T-1
    uint64_t* bars = reinterpret_cast<uint64_t*>(sm + S::kFloats);  // [kBars]
    if (tid == 0) {
        if ((smem_u32(sm) & 127u) != 0u) __trap();
#pragma unroll
        for (int i = 0; i < S::kBars; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }

    // Fused halos: a brick of the lowest / highest layer reads halo planes the neighbour's previous kernel wrote and
    // stores into the neighbour's halo planes that its previous kernel still read: wait for that kernel (once per side).
    bool waited_lo = !FUSED || !pv.has_lo, waited_hi = !FUSED || !pv.has_hi;
    auto peer_sync = [&](const bool lo, const bool hi) {  // uniform; thread 0 polls
        if constexpr (!FUSED) return;
        const bool wl = lo && !waited_lo, wh = hi && !waited_hi;
        if (!(wl || wh)) return;
        if (tid == 0) peer_wait(pv, need, wl, wh);
        waited_lo |= wl;
        waited_hi |= wh;
    };
    const int layer = P.ntx * P.nty;
    auto brick_faces = [&](const int brick, bool& lo, bool& hi) {
        lo = FUSED && pv.has_lo && brick < layer;
        hi = FUSED && pv.has_hi && brick >= layer * (P.nzc - 1);
    };

    // The frozen bricks of the previous pass: one copy each into the other pressure buffer.  Of the bricks that froze
    // in the FIRST pass (the far field: most of a large grid) only those within reach of a still-active brick are ever
    // read again in this frame, so only those are copied (brick_flag == 3); the others stay final in the first pass's
    // output buffer, which jacobi_settle_kernel makes the frame's final buffer.
    if (pass == 1) {
        __shared__ int s_copied;
        __shared__ unsigned s_todo;
        if (tid == 0) s_copied = 0;
        __syncthreads();
        const int nbricks = layer * P.nzc;
            auto rot = [&](const int i) { return !FUSED ? i : (i + layer < nbricks ? i + layer : i + layer - nbricks); };
        for (int b0 = blockIdx.x; b0 < nbricks; b0 += gridDim.x * 32) {
            // 32 candidate bricks of this CTA at a time: one round trip for their flags
            // (fused halos: candidates rotated by one layer, the face layers last, as in the first pass)
            const int bi = b0 + (tid & 31) * gridDim.x;
            const int b = rot(bi);
            const int f = (tid < 32 && bi < nbricks) ? W.brick_flag[b] : 0;
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
                if (lo || hi) {
                    peer_sync(lo, hi);
                    __syncthreads();
                }
                copy_frozen_brick<S, FUSED>(p_in, p_out, m_out, P, brick, pv, peers, pi, mi);
                if (tid == 0) ++s_copied;
            }
            __syncthreads();
        }
        if (tid == 0 && s_copied) atomicAdd(&state->bricks_copied, (unsigned long long)s_copied);
    } else if (n_copy > 0) {
        const int* __restrict__ copy_list = W.copy[pass & 1];
        // handed out from the LAST CTA down: in the late passes the first CTAs are the ones that relax bricks
        for (int w = gridDim.x - 1 - blockIdx.x; w < n_copy; w += gridDim.x) {
            const int brick = copy_list[w];
            bool lo, hi;
            brick_faces(brick, lo, hi);
            if (lo || hi) {
                peer_sync(lo, hi);
                __syncthreads();
            }
            copy_frozen_brick<S, FUSED>(p_in, p_out, m_out, P, brick, pv, peers, pi, mi);
        }
        if (tid == 0 && blockIdx.x == 0) atomicAdd(&state->bricks_copied, (unsigned long long)n_copy);
    }
    __syncthreads();  // barriers initialised
    FXB_STAMP();

    const int n_ext = P.ntx * P.nty * ((P.ext_lo + P.bz - 1) / P.bz + (P.ext_hi + P.bz - 1) / P.bz);
    const int n_work = n_relax + n_ext;
    const int nxb = P.pitch >> 3;
    const float eps = P.early_exit ? kEps : -1.0f;

    // First pass: the bricks (first_brick + w) mod bricks, w < first_count — all of them on a single GPU; with fused
    // halos the launch is split (launch_jacobi_pass_fused): the interior layers here, without any of the halo code, and
    // the layers at the slab's faces in a second, small launch.
    const int nbricks_all = layer * P.nzc;
    auto item_of = [&](const int w, const int listed) -> Item {
        if (w < n_relax) {
            int brick = listed;
            if (pass == 0) {
                brick = w + P.first_brick;
                if (brick >= nbricks_all) brick -= nbricks_all;
            }
            return own_item<S>(P, brick);
        }
        return ext_item<S>(P, w - n_relax);
    };

    // ---- producer: one stream of (p plane, rhs plane) bundles over all work items of this CTA ------------------
    int pw = blockIdx.x;       // work item the producer stands in
    int p_listed = pre_a;      // its list entry, and the next one's (prefetched)
    int p_listed_next = pre_b;
    Item pit = item_of(pw < n_work ? pw : 0, p_listed);
    int pk = max(pit.zs - T, 0), pk_end = min(pit.ze + T, P.nz_alloc);
    bool p_first = true;       // the next bundle is the first of its work item
    unsigned issued = 0;       // bundles issued so far
    auto issue_next = [&]() {  // uniform in every thread; thread 0 talks to the copy engine
        if (pw >= n_work) return;
        if (fused_halos && p_first && pit.brick >= 0) {  // before anything of a face brick is staged: the neighbour's data is there
            bool lo, hi;
            brick_faces(pit.brick, lo, hi);
            peer_sync(lo, hi);
        }
        p_first = false;
        if (tid == 0) {
            uint64_t* bar = &bars[issued % S::kBars];
            mbar_expect_tx(bar, 2u * kPlane * 4u);
            tma_load_3d(sm_p + (issued % S::kPSlots) * kPlane, map_in, pit.gx0, pit.gy0, pk, bar);
            tma_load_3d(sm_rhs + (issued % S::kRSlots) * kPlane, &map_rhs, pit.gx0, pit.gy0, pk, bar);
        }
        ++issued;
        if (++pk == pk_end) {
            pw += gridDim.x;
            p_first = true;
            p_listed = p_listed_next;
            const int w2 = pw + gridDim.x;
            p_listed_next = (pass > 0 && w2 < n_relax) ? list_in[w2] : w2;
            if (pw < n_work) {
                pit = item_of(pw, p_listed);
                pk = max(pit.zs - T, 0);
                pk_end = min(pit.ze + T, P.nz_alloc);
            }
        }
    };
#pragma unroll
    for (int i = 0; i < S::kDepth; ++i) issue_next();
    if constexpr (FUSED) __syncthreads();  // thread 0's waits above come before anybody's loads of the neighbours' data

    // ---- consumer ------------------------------------------------------------------------------------------------
    unsigned consumed = 0;        // bundles consumed so far = flat index of the next bundle
    unsigned tot[T + 1];          // cells of this CTA's own bricks still active after each level
#pragma unroll
    for (int l = 0; l <= T; ++l) tot[l] = 0;
    unsigned n_done = 0;          // own bricks finished by this CTA
    // list append of the previous brick (thread 0): the atomic's result is consumed one brick later, off the critical path
    int pend_brick = -1, pend_slot = 0;
    int* pend_list = nullptr;
    int c_listed = pre_a, c_listed_next = pre_b;

    for (int work = blockIdx.x; work < n_work; work += gridDim.x) {
        const Item it = item_of(work, c_listed);
        c_listed = c_listed_next;
        {
            const int w2 = work + 2 * gridDim.x;
            c_listed_next = (pass > 0 && w2 < n_relax) ? list_in[w2] : w2;
        }
        const int zs = it.zs, ze = it.ze;
        const int zl0 = max(zs - T, 0), zl1 = min(ze + T, P.nz_alloc);
        const unsigned c0 = consumed;  // flat index of the bundle of plane zl0

        // ---- per-thread geometry of this tile ----
        const int gx = it.gx0 + 4 * li;
        const int ry0 = S::kWarpRows * warp + kRows * sub;  // tile row of this thread's row 0
        const int gyb = it.gy0 + ry0;                       // grid y of this thread's row 0; row r is gyb + r
        const int xe = (gx < P.nx && gx + 4 > P.nx) ? P.nx - gx : 0;  // cells of a quad cut by the grid's x face
        const unsigned qmask = gx >= 0 && gx < P.nx ? (xe ? (1u << xe) - 1u : 0xFu) : 0u;
        unsigned dom_bits = 0, own_bits = 0;  // bit (4r + j): cell j of row r lies inside the grid / is this brick's output
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            if (gyb + r >= 0 && gyb + r < P.ny) dom_bits |= qmask << (4 * r);
            if (li >= 1 && li <= LX - 2 && ry0 + r >= T && ry0 + r < kTileY - T) own_bits |= 0xFu << (4 * r);
        }
        if (producer) dom_bits = 0;  // (its row index lies outside the tile)
        own_bits &= dom_bits;
        int off0 = ry0 * kTileX + 4 * li;  // own quad of row 0 inside a staged plane; row r: + r * kTileX
        const bool clamp_u = ry0 == 0 || gyb <= 0;                                       // no row above inside the grid / tile
        const bool clamp_d = ry0 + kRows - 1 == kTileY - 1 || gyb + kRows - 1 >= P.ny - 1;  // no row below
        int off_up = clamp_u ? off0 : off0 - kTileX;
        int off_dn = clamp_d ? off0 + (kRows - 1) * kTileX : off0 + kRows * kTileX;
        const bool clamp_l = li == 0 || gx == 0;
        const bool clamp_r = li == LX - 1 || gx + 4 >= P.nx;
        // per-thread constants of the brick the compiler would otherwise re-derive from the thread index at every use
        asm volatile("" : "+r"(off_up), "+r"(off_dn), "+r"(own_bits), "+r"(dom_bits));
        // the grid's faces may cut through this thread's rows / four cells: ghosts mirror the adjacent inside
        // row / cell after every update so that the in-register neighbours obey the clamp rule
        const bool y_ghost_lo = kRows == 2 && gyb == -1, y_ghost_hi = kRows == 2 && gyb == P.ny - 1;
        const bool any_ghost = y_ghost_lo || y_ghost_hi || xe != 0;
        auto fix_ghosts = [&](float4 (&v)[kRows]) {
            if (any_ghost) {
                if constexpr (kRows == 2) {
                    if (y_ghost_lo) v[0] = v[1];
                    if (y_ghost_hi) v[1] = v[0];
                }
#pragma unroll
                for (int r = 0; r < kRows; ++r) {
                    if (xe == 1) v[r].y = v[r].x;
                    if (xe == 2) v[r].z = v[r].y;
                    if (xe == 3) v[r].w = v[r].z;
                }
            }
        };

        // Freeze flags of the level-0 planes (the previous pass's output mask): raw bytes are fetched two iterations
        // ahead and decoded when their plane is consumed.  (L2 loads: a neighbour rank may have written halo planes.)
        const size_t mrow0 = (size_t)max(gyb, 0) * nxb + (max(gx, 0) >> 3);
        const size_t mplane = (size_t)P.ny * nxb;
        auto fetch_flags = [&](const int z, unsigned& raw0, unsigned& raw1) {
            if (pass == 0 || z >= zl1) return;
            if (dom_bits & 0x0Fu) raw0 = __ldcg(m_in + (size_t)z * mplane + mrow0);
            if (kRows == 2 && (dom_bits & 0xF0u)) raw1 = __ldcg(m_in + (size_t)z * mplane + mrow0 + ((dom_bits & 0x0Fu) ? nxb : 0));
        };
        const int nib_shift = gx & 4;
        unsigned raw_a0 = 0, raw_a1 = 0, raw_b0 = 0, raw_b1 = 0;  // bytes of the plane consumed next / the one after
        if (!producer) {
            fetch_flags(zl0, raw_a0, raw_a1);
            fetch_flags(zl0 + 1, raw_b0, raw_b1);
        }

        // z queue: level l (0..T-1) keeps the previous and the centre plane of this thread's column, with the flags of
        // the centre plane
        float4 qp[T][kRows], qc[T][kRows];
        unsigned qf[T];
#pragma unroll
        for (int l = 0; l < T; ++l) {
            qf[l] = 0;
#pragma unroll
            for (int r = 0; r < kRows; ++r) qp[l][r] = qc[l][r] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        unsigned alive = 0;  // an own cell of this brick is still active after the pass's last sweep
        const int k_end = ze - 1 + T;
        int lo_of[T + 1], hi_of[T + 1];  // planes level l must produce (trapezoid in z)
#pragma unroll
        for (int l = 1; l <= T; ++l) {
            lo_of[l] = max(zs - (T - l), 0);
            hi_of[l] = min(ze + (T - l), P.nz_alloc);
        }
        // ring positions (in floats), carried through the loop: the staged plane k and k-1, the rhs plane of every level
        int po_new = (int)(c0 % S::kPSlots) * kPlane, po_prev = po_new;
        int ro[T + 1];
        ro[0] = (int)(c0 % S::kRSlots) * kPlane;
#pragma unroll
        for (int l = 1; l <= T; ++l) ro[l] = ro[0];
        int pub_w = 0;  // half of the published planes written in this iteration (read in the next one)

#pragma unroll 1  // (unrolling by 2 or 3 was measured on B200: same time, 2-3x the code)
        for (int k = zl0; k <= k_end; ++k) {
            float4 nw[kRows];   // the plane the level below produced in this iteration
            unsigned nfl = 0;
            bool have = k < zl1;
            FXB_STAMP();
            if (have) {
                issue_next();
                FXB_STAMP();
                if (!producer) {
                    mbar_wait(&bars[consumed % S::kBars], (consumed / S::kBars) & 1u);
                    FXB_STAMP();
                    const float* src = sm_p + po_new + off0;
#pragma unroll
                    for (int r = 0; r < kRows; ++r) nw[r] = *reinterpret_cast<const float4*>(src + r * kTileX);
                    fix_ghosts(nw);
                    nfl = pass == 0 ? dom_bits
                                    : ((((raw_a0 >> nib_shift) & 0xFu) | (((raw_a1 >> nib_shift) & 0xFu) << 4)) & dom_bits);
                    raw_a0 = raw_b0;
                    raw_a1 = raw_b1;
                    FXB_STAMP();
                    fetch_flags(k + 2, raw_b0, raw_b1);
                }
                ++consumed;
            } else {
#pragma unroll
                for (int r = 0; r < kRows; ++r) nw[r] = make_float4(0.f, 0.f, 0.f, 0.f);
            }

            auto level = [&](auto lc) {
                constexpr int l = decltype(lc)::value;
                const int z = k - l;
                const bool run = z >= lo_of[l] && z < hi_of[l];
                float4 out[kRows];
#pragma unroll
                for (int r = 0; r < kRows; ++r) out[r] = nw[r];
                unsigned st = 0;
                if (run) {
                    const unsigned act = (l <= levels) ? qf[l - 1] : 0u;
                    if (__any_sync(kFull, act != 0u)) {
                        // y neighbours of the centre plane (level l-1, plane z): staged plane for level 0, published otherwise
                        const float* nb = (l == 1) ? sm_p + po_prev : sm_pub + ((l - 2) * 2 + (pub_w ^ 1)) * kPlane;
                        const float4 up = *reinterpret_cast<const float4*>(nb + off_up);
                        const float4 dn = *reinterpret_cast<const float4*>(nb + off_dn);
                        const float* rb = sm_rhs + ro[l] + off0;
                        float4 rhs[kRows];
#pragma unroll
                        for (int r = 0; r < kRows; ++r) rhs[r] = *reinterpret_cast<const float4*>(rb + r * kTileX);
                        // clamp rule at the grid's z faces: the missing neighbour plane is the centre plane itself (the
                        // overwritten queue entries are dead: no plane follows the face, none precedes it)
                        if (z == P.z_face_lo) {
#pragma unroll
                            for (int r = 0; r < kRows; ++r) qp[l - 1][r] = qc[l - 1][r];
                        }
                        if (z + 1 == P.z_face_hi) {
#pragma unroll
                            for (int r = 0; r < kRows; ++r) nw[r] = qc[l - 1][r];
                        }
#pragma unroll
                        for (int r = 0; r < kRows; ++r) {
                            const float4 c = qc[l - 1][r];
                            float left = __shfl_up_sync(kFull, c.w, 1, LX), right = __shfl_down_sync(kFull, c.x, 1, LX);
                            if (clamp_l) left = c.x;
                            if (clamp_r) right = c.w;
                            const float4 u = r == 0 ? up : qc[l - 1][r > 0 ? r - 1 : 0];
                            const float4 d = r == kRows - 1 ? dn : qc[l - 1][r < kRows - 1 ? r + 1 : r];
                            st |= relax_quad(c, qp[l - 1][r], nw[r], u, d, left, right, rhs[r], (act >> (4 * r)) & 0xFu, eps,
                                             out[r]) << (4 * r);
                        }
                        fix_ghosts(out);
                        if (z >= zs && z < ze && it.brick >= 0) {  // halo bricks are counted by their owner
                            tot[l] += __popc(st & own_bits);
                            if (l == levels) alive |= st & own_bits;
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < kRows; ++r) out[r] = qc[l - 1][r];
                    }
                    if constexpr (l < T) {  // publish the new plane for the rows above / below (read next iteration)
                        float* dst = sm_pub + ((l - 1) * 2 + pub_w) * kPlane + off0;
#pragma unroll
                        for (int r = 0; r < kRows; ++r) *reinterpret_cast<float4*>(dst + r * kTileX) = out[r];
                    } else if (z >= zs && z < ze) {  // level T: the pass's output
                        const bool to_lo = fused_halos && it.brick >= 0 && pushes_lo(pv, P, z);
                        const bool to_hi = fused_halos && it.brick >= 0 && pushes_hi(pv, P, z);
#pragma unroll
                        for (int r = 0; r < kRows; ++r) {
                            const bool mine = (own_bits >> (4 * r)) & 1u;
                            const size_t at = ((size_t)z * P.ny + (gyb + r)) * P.pitch + gx;
                            if (mine) *reinterpret_cast<float4*>(p_out + at) = out[r];
                            // bit-packed freeze flags: two quads (8 cells) per byte, written by the odd lane
                            const unsigned nib = (st >> (4 * r)) & 0xFu;
                            const unsigned hi = __shfl_down_sync(kFull, nib, 1, LX);
                            const size_t mat = ((size_t)z * P.ny + (gyb + r)) * nxb + (gx >> 3);
                            const unsigned char byte = (unsigned char)(nib | (hi << 4));
                            if (mine && (li & 1)) m_out[mat] = byte;
                            if (FUSED && (to_lo || to_hi)) {  // the same stores into the neighbour's halo planes
                                const long long plane_f = (long long)P.ny * P.pitch, plane_b = (long long)P.ny * nxb;
                                if (mine && to_lo) *reinterpret_cast<float4*>(peers.p[0][pi] + (long long)at + pv.dz_lo * plane_f) = out[r];
                                if (mine && to_hi) *reinterpret_cast<float4*>(peers.p[1][pi] + (long long)at + pv.dz_hi * plane_f) = out[r];
                                if (mine && (li & 1) && to_lo) peers.m[0][mi][(long long)mat + pv.dz_lo * plane_b] = byte;
                                if (mine && (li & 1) && to_hi) peers.m[1][mi][(long long)mat + pv.dz_hi * plane_b] = byte;
                            }
                        }
                    }
                }
                // the level below moves on: its centre plane becomes the previous one
                if (have) {
#pragma unroll
                    for (int r = 0; r < kRows; ++r) {
                        qp[l - 1][r] = qc[l - 1][r];
                        qc[l - 1][r] = nw[r];
                    }
                    qf[l - 1] = nfl;
                }
#pragma unroll
                for (int r = 0; r < kRows; ++r) nw[r] = out[r];
                nfl = st;
                have = run;
            };
            if (!producer) {
                level(std::integral_constant<int, 1>{});
                if constexpr (T >= 2) level(std::integral_constant<int, 2>{});
                if constexpr (T >= 3) level(std::integral_constant<int, 3>{});
                if constexpr (T >= 4) level(std::integral_constant<int, 4>{});
            }
            // ring positions of the next iteration
            po_prev = po_new;
            po_new = po_new + kPlane == S::kPSlots * kPlane ? 0 : po_new + kPlane;
#pragma unroll
            for (int l = T; l >= 1; --l) ro[l] = ro[l - 1];
            ro[0] = ro[0] + kPlane == S::kRSlots * kPlane ? 0 : ro[0] + kPlane;
            pub_w ^= 1;
            FXB_STAMP();
            __syncthreads();
        }

        // ---- brick state: still active -> relax again next pass; just frozen -> one copy next pass ----
        if (it.brick >= 0) {
            const int any_alive = __syncthreads_or(alive != 0u);
            if (pass == 0) {
                // first pass: a frozen brick is only flagged; a still-active one flags the 26 bricks around it (their
                // cells are within reach of its tile in x, y or z), which pass 1 then copies if they froze
                if (!any_alive) {
                    if (tid == 0) atomicOr(&W.brick_flag[it.brick], 1);
                } else if (tid < 27) {
                    const int tx = it.brick % P.ntx, ty = (it.brick / P.ntx) % P.nty, tz = it.brick / (P.ntx * P.nty);
                    const int nx_ = tx + tid % 3 - 1, ny_ = ty + (tid / 3) % 3 - 1, nz_ = tz + tid / 9 - 1;
                    if (tid != 13 && nx_ >= 0 && nx_ < P.ntx && ny_ >= 0 && ny_ < P.nty && nz_ >= 0 && nz_ < P.nzc)
                        atomicOr(&W.brick_flag[(nz_ * P.nty + ny_) * P.ntx + nx_], 2);
                }
            }
            if (tid == 0 && (any_alive || pass > 0)) {
                if (pend_brick >= 0) pend_list[pend_slot] = pend_brick;
                pend_brick = it.brick;
                if (any_alive) {
                    pend_list = W.relax[(pass + 1) & 1];
                    pend_slot = atomicAdd(&W.relax_count[pass + 1], 1);
                } else {
                    pend_list = W.copy[(pass + 1) & 1];
                    pend_slot = atomicAdd(&W.copy_count[pass + 1], 1);
                }
            }
            ++n_done;
        }
    }
    if (tid == 0 && pend_brick >= 0) pend_list[pend_slot] = pend_brick;

    // ---- per-level active counts of this CTA's bricks -> global counters --------------------------------------
    __shared__ unsigned s_cnt[T];
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
    if (tid == 0 && n_done) atomicAdd(&state->bricks_processed, (unsigned long long)n_done);
    finish();
}

// After the last pass.  The frame's final pressure must be complete in ONE buffer: the first pass's output buffer Y,
// the only one that holds the bricks that froze in the first pass and were never copied.  Every later brick is final
// in both buffers once copied — except the bricks relaxed by the last executed pass L when that pass wrote the other
// buffer (L odd): those (pass L + 1's relax and copy lists) are copied into Y here.  Also the solve's bookkeeping that
// finish_solve_kernel does for the per-sweep path: s_exec, pass count (which buffer holds P: jacobi_flip_kernel).
template <class S, bool FUSED>
__global__ void __launch_bounds__(S::kThreads)
jacobi_settle_kernel(const FrameParams* __restrict__ frame, StepState* __restrict__ state, float* p0, float* p1,
                     unsigned char* m0, unsigned char* m1, const __grid_constant__ WorkLists W,
                     const __grid_constant__ PassParams P, const int fuse_t, const int tail_from, const int force_passes,
                     const __grid_constant__ PeerView pv,
                     const __grid_constant__ JacobiPeers peers) {
    const int iters = P.levels_total;
    const bool live = 0.0f < frame->dt && iters > 0;
    int s = 0;
    if (live) {
        s = 1;
        while (s < iters && state->active_after[s - 1] != 0ull) ++s;
    }
    // passes executed: T sweeps each up to pass tail_from, four each from there on
    int passes = (s + fuse_t - 1) / fuse_t;
    if (passes > tail_from) passes = tail_from + (s - tail_from * fuse_t + 3) / 4;
    if (force_passes >= 0 && live) passes = force_passes;
    const int p_cur = state->p_cur;  // still the frame's input buffer X; Y is the other one
    if (passes > 0 && ((passes - 1) & 1)) {
        const int L = passes - 1;
        const float* src = p_cur ? p1 : p0;   // out(L) = X for odd L
        float* dst = p_cur ? p0 : p1;
        unsigned char* m_dst = ((L + 1) & 1) ? m0 : m1;  // as pass L + 1 would have cleared it (unused after the frame)
        const int pi = p_cur ? 0 : 1, mi = ((L + 1) & 1) ? 0 : 1;
        const int n_r = W.relax_count[L + 1], n_c = W.copy_count[L + 1];
        const int* __restrict__ lr = W.relax[(L + 1) & 1];
        const int* __restrict__ lc = W.copy[(L + 1) & 1];
        if (FUSED && n_r + n_c > 0) {  // the neighbours' last pass no longer reads the halos written here
            if (threadIdx.x == 0) peer_wait(pv, frame->epoch_base + (unsigned long long)P.event, true, true);
            __syncthreads();
        }
        for (int w = blockIdx.x; w < n_r + n_c; w += gridDim.x)
            copy_frozen_brick<S, FUSED>(src, dst, m_dst, P, w < n_r ? lr[w] : lc[w - n_r], pv, peers, pi, mi);
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {  // (p_cur itself is flipped by the next kernel: other CTAs still read it)
        if (force_passes < 0) {  // (multi-GPU: the global figure comes from the summed counters, fxb_api.cu)
            state->s_exec = s;
            state->total_sweeps += (unsigned long long)s;
        }
        state->passes = passes;
        state->total_passes += (unsigned long long)passes;
    }
}

// p_cur flips once per frame with a pressure solve (to the first pass's output buffer); a kernel of its own so that
// no CTA of jacobi_settle_kernel can observe the new value.  Fused halos: the settle step's event.
__global__ void jacobi_flip_kernel(const FrameParams* __restrict__ frame, StepState* __restrict__ state, int iters,
                                   const __grid_constant__ PeerView pv, int event) {
    if (0.0f < frame->dt && iters > 0 && state->passes > 0) state->p_cur ^= 1;
    phase_mark(state, 2);  // the pressure solve ends here
#ifdef FXB_TIMING
    {
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        state->dbg[105] = (long long)now;  // the solve ended
    }
#endif
    if (pv.has_lo || pv.has_hi) peer_publish(pv, frame->epoch_base + (unsigned long long)event + 1ull);
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

// Tensor map of a pitched [nz_alloc][ny][pitch] float array seen as (nx, ny, nz_alloc): elements beyond nx / ny (and
// tile parts at negative coordinates) arrive as zeros; the kernel never uses them (clamp by index).
bool make_plane_map(CUtensorMap* map, float* base, int nx, int ny, int pitch, int nz_alloc, int tile_x, int tile_y,
                    int tile_z = 1) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz_alloc};
    const cuuint64_t strides[2] = {(cuuint64_t)pitch * 4, (cuuint64_t)pitch * ny * 4};
    const cuuint32_t box[3] = {(cuuint32_t)tile_x, (cuuint32_t)tile_y, (cuuint32_t)tile_z};
    const cuuint32_t estr[3] = {1, 1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// The shapes in use.  Wide: tile rows of 32 lanes (128 cells); narrow: 16 lanes (64 cells, a warp covers two row groups),
// chosen per grid by fused_jacobi_plan so that the tiles overhang the grid's faces as little as possible.  T <= 2: two
// rows per thread, 8 warps, two CTAs per SM, TMA depth 3 (the default, T = 2); T > 2: one CTA per SM.  Measured on B200
// (profiles/): fusing 4 sweeps in the later passes, or spreading a tile over 16 one-row warps there, made the
// latency-bound tail of the solve slower, not faster, so one shape runs every pass.
template <int T> using WideU = Shape<T, 32, 8, (T <= 2 ? 3 : 2), (T <= 2 ? 2 : 1)>;
template <int T> using NarrowU = Shape<T, 16, 8, (T <= 2 ? 3 : 2), (T <= 2 ? 2 : 1)>;

template <class S>
cudaError_t launch_shape(const FusedJacobi& J, const Domain& d, const FrameParams* frame, StepState* state, int pass,
                         int s0, int iters, int early_exit, bool run_all, int ext_lo, int ext_hi, int first_brick,
                         int first_count, bool plain, const PeerView& pv, cudaStream_t stream) {
    // the opt-in above the 48 KB default is per device: set it whenever the device changes (cheap, idempotent)
    static int attr_device = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (attr_device != dev) {
        cudaError_t e = cudaSuccess;
        const void* fns[4] = {(const void*)jacobi_pass_kernel<S, false, false>, (const void*)jacobi_pass_kernel<S, true, false>,
                              (const void*)jacobi_pass_kernel<S, false, true>, (const void*)jacobi_pass_kernel<S, true, true>};
        for (int i = 0; i < 4 && e == cudaSuccess; ++i)
            e = cudaFuncSetAttribute(fns[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S::kBytes);
        if (e != cudaSuccess) return e;
        attr_device = dev;
    }
    PassParams P = make_pass_params(J, d, pass, iters, early_exit, run_all, ext_lo, ext_hi);
    if (pass == 0 && first_count >= 0) {
        P.first_brick = first_brick;
        P.first_count = first_count;
    }
    const JacobiPeers peers = make_jacobi_peers(J);
    const int nbricks = pass == 0 ? P.first_count : J.ntx * J.nty * J.nzc;
    const int slots = J.num_sms * S::kCtasPerSm;  // persistent CTAs
    const int grid = nbricks < slots ? (nbricks > 0 ? nbricks : 1) : slots;
    const WorkLists W = make_work_lists(J);
    const CUtensorMap& mp0 = *reinterpret_cast<const CUtensorMap*>(J.map_p[0]);
    const CUtensorMap& mp1 = *reinterpret_cast<const CUtensorMap*>(J.map_p[1]);
    const CUtensorMap& mr = *reinterpret_cast<const CUtensorMap*>(J.map_rhs);
    const bool fused = (pv.has_lo || pv.has_hi) && !plain;
#define FXB_GO(FU, FI) jacobi_pass_kernel<S, FU, FI><<<grid, S::kThreads, S::kBytes, stream>>>(mp0, mp1, mr, frame, state, J.p[0], \
                                                                                        J.p[1], J.mask[0], J.mask[1], W, P, pv, peers)
    if (pass == 0) {
        if (fused) FXB_GO(true, true); else FXB_GO(false, true);
    } else {
        if (fused) FXB_GO(true, false); else FXB_GO(false, false);
    }
#undef FXB_GO
    return cudaGetLastError();
}

// Tiles needed to cover n cells with own regions of `out` cells.
int tiles_for(int n, int out) { return (n + out - 1) / out; }

}  // namespace

bool fused_jacobi_supported(const Domain& d) { return d.nz > 1 && d.nx >= 8 && (d.pitch % 8) == 0; }

int fused_jacobi_plan(FusedJacobi* J, const Domain& d, int fuse_t, float* p0, float* p1, float* rhs, bool allow_tail) {
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    J->T = fuse_t == 0 ? 2 : fuse_t;  // 0 = library default
    const int T = J->T;
    // tile shape: the one whose tiles cover the least area beyond the grid (compute and staging are per tile cell)
    const long long wide = (long long)tiles_for(d.nx, 120) * 128 * tiles_for(d.ny, 16 - 2 * T) * 16;
    const long long narrow = (long long)tiles_for(d.nx, 56) * 64 * tiles_for(d.ny, 32 - 2 * T) * 32;
    J->narrow = narrow < wide;
    if (const char* e = getenv("FXB_TILE")) J->narrow = atoi(e) == 64;  // tuning knob: 64 or 128
    J->tile_x = J->narrow ? 64 : 128;
    J->tile_y = J->narrow ? 32 : 16;  // of the first pass's kernel
    const int out_x = J->tile_x - 2 * kHaloX, out_y = J->tile_y - 2 * T;
    J->ntx = tiles_for(d.nx, out_x);
    J->nty = tiles_for(d.ny, out_y);
    const int nz_out = d.z_own1 - d.z_own0;
    J->bz = nz_out >= 8 ? 8 : nz_out;
    J->nzc = (nz_out + J->bz - 1) / J->bz;
    J->p[0] = p0; J->p[1] = p1; J->rhs = rhs;
    if (!make_plane_map(reinterpret_cast<CUtensorMap*>(J->map_p[0]), p0, d.nx, d.ny, d.pitch, d.nz_alloc, J->tile_x, J->tile_y)) return -1;
    if (!make_plane_map(reinterpret_cast<CUtensorMap*>(J->map_p[1]), p1, d.nx, d.ny, d.pitch, d.nz_alloc, J->tile_x, J->tile_y)) return -1;
    if (!make_plane_map(reinterpret_cast<CUtensorMap*>(J->map_rhs), rhs, d.nx, d.ny, d.pitch, d.nz_alloc, J->tile_x, J->tile_y)) return -1;
    // Brick-resident form for the latency-bound part of the solve.  Measured on B200 (profiles/): from the second pass
    // on the resident form is the faster one at 128^3, 256^3 and 512^3 (the first pass, every brick dense, is not).
    J->resident_from = FusedJacobi::kMaxPasses + 1;
    J->tail_from = FusedJacobi::kMaxPasses + 1;
    if (resident_jacobi_supported(*J)) {
        const int planes = resident_jacobi_window_planes(*J);
        if (!make_plane_map(reinterpret_cast<CUtensorMap*>(J->map3_p[0]), p0, d.nx, d.ny, d.pitch, d.nz_alloc, J->tile_x, J->tile_y, planes)) return -1;
        if (!make_plane_map(reinterpret_cast<CUtensorMap*>(J->map3_p[1]), p1, d.nx, d.ny, d.pitch, d.nz_alloc, J->tile_x, J->tile_y, planes)) return -1;
        if (!make_plane_map(reinterpret_cast<CUtensorMap*>(J->map3_rhs), rhs, d.nx, d.ny, d.pitch, d.nz_alloc, J->tile_x, J->tile_y, planes)) return -1;
        J->resident_from = 1;
        if (const char* e = getenv("FXB_RESIDENT_FROM")) J->resident_from = atoi(e);  // tuning knob: first resident pass
        // Tail schedule (T = 2, 64-wide tiles): four sweeps per pass on half bricks once the passes are latency-bound.
        // (z-slabs: the tail passes read four halo planes, which the halos must provide and the backend must push)
        const int halo_lo = d.z_own0 - d.z_first, halo_hi = d.z_first + d.nz_alloc - d.z_own1;
        const bool halo_ok = (d.z_own0 == 0 || halo_lo >= 4) && (d.z_own1 == d.nz || halo_hi >= 4);
        if (T == 2 && J->narrow && allow_tail && halo_ok) {
            // Measured on B200 (tools/quick_time.py, FXB_TAIL_FROM sweep): a four-sweep pass on half bricks costs about as
            // much as 1.7 two-sweep passes (halo of four: 44 plane-updates per column instead of 2 x 18, 22 rows for 14
            // own), so it only pays where a pass is pure latency — early on small grids (128^3: 0.346 -> 0.312 ms per
            // step from pass 5), late on large ones (256^3: 0.908 -> 0.876 ms from pass 16; 512^3: -0.6 %).
            // (decided from the GLOBAL grid: every rank of a z-slab run must follow the same schedule)
            J->tail_from = (long long)J->ntx * J->nty * ((d.nz + 7) / 8) < 1000 ? 5 : 16;
            if (const char* e = getenv("FXB_TAIL_FROM")) J->tail_from = atoi(e);  // tuning knob: first four-sweep pass
            if (J->tail_from < J->resident_from) J->tail_from = J->resident_from;
            if (J->tail_from < 1) J->tail_from = 1;
            if (J->tail_from > FusedJacobi::kMaxPasses) J->tail_from = FusedJacobi::kMaxPasses + 1;
            if (!make_plane_map(reinterpret_cast<CUtensorMap*>(J->map4_p[0]), p0, d.nx, d.ny, d.pitch, d.nz_alloc, 64, 22, 16)) return -1;
            if (!make_plane_map(reinterpret_cast<CUtensorMap*>(J->map4_p[1]), p1, d.nx, d.ny, d.pitch, d.nz_alloc, 64, 22, 16)) return -1;
            if (!make_plane_map(reinterpret_cast<CUtensorMap*>(J->map4_rhs), rhs, d.nx, d.ny, d.pitch, d.nz_alloc, 64, 22, 16)) return -1;
        }
    }
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&J->num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return -1;
    J->push_depth = J->tail_from <= FusedJacobi::kMaxPasses ? 4 : T;
    // the kernel prefetches list entries up to two grid strides ahead: the lists are padded accordingly
    J->list_stride = (int)fused_jacobi_bricks(*J) + 4 * J->num_sms + 8;
    return 0;
}

int fused_jacobi_passes(const FusedJacobi& J, int iters) {
    if (iters <= 0) return 0;
    const int head = J.tail_from * J.T;  // sweeps of the passes before the tail schedule
    if (iters <= head) return (iters + J.T - 1) / J.T;
    return J.tail_from + (iters - head + 3) / 4;
}

size_t fused_jacobi_bricks(const FusedJacobi& J) { return (size_t)J.ntx * J.nty * J.nzc; }

size_t fused_jacobi_brick_cells(const FusedJacobi& J) { return (size_t)(J.tile_x - 2 * kHaloX) * (J.tile_y - 2 * J.T) * J.bz; }

void fused_jacobi_brick_extent(const FusedJacobi& J, int out[3]) {
    out[0] = J.tile_x - 2 * kHaloX; out[1] = J.tile_y - 2 * J.T; out[2] = J.bz;
}

// With fused halos the first pass is split when the resident kernel is available: the interior layers run in the
// marching kernel WITHOUT the halo code (measured: the instantiation with it is ~25 % slower on every brick), the one
// or two layers at the slab's interior faces follow in the resident kernel, which pushes their planes to the
// neighbours and publishes the pass's event.
static bool first_pass_split(const FusedJacobi& J, const PeerView& pv) {
    return (pv.has_lo || pv.has_hi) && resident_jacobi_supported(J);
}

int fused_jacobi_launches(const FusedJacobi& J, const PeerView& pv, int pass) {
    if (pass != 0 || pass >= J.resident_from || !first_pass_split(J, pv)) return 1;
    const int layer = J.ntx * J.nty, faces = std::min(J.nzc, (pv.has_lo ? 1 : 0) + (pv.has_hi ? 1 : 0));
    return layer * J.nzc - faces * layer > 0 ? 2 : 1;
}

cudaError_t launch_jacobi_pass_fused(const FusedJacobi& J, const Domain& d, const FrameParams* frame, StepState* state,
                                     int pass, int iters, int early_exit, bool run_all, int ext_lo, int ext_hi,
                                     const PeerView& pv, cudaStream_t stream) {
    if (pass >= J.resident_from)
        return launch_jacobi_pass_resident(J, d, frame, state, pass, iters, early_exit, run_all, ext_lo, ext_hi, 0, -1, pv,
                                           stream);
    const int s0 = fused_jacobi_s0(J, pass);
    int first_brick = 0, first_count = -1;
    bool plain = false;
    if (pass == 0 && first_pass_split(J, pv)) {
        // bricks in rotated order: interior layers, then the top layer, then the bottom layer
        const int layer = J.ntx * J.nty, nbricks = layer * J.nzc;
        const int faces = std::min(J.nzc, (pv.has_lo ? 1 : 0) + (pv.has_hi ? 1 : 0));
        const int interior = nbricks - faces * layer;
        const int face_first = (nbricks - (pv.has_hi ? layer : 0)) % nbricks;
        if (interior > 0) {
            first_brick = pv.has_lo ? layer : 0;
            first_count = interior;
            plain = true;
        }
        // (launch order: the interior first — whatever the neighbours lag behind is absorbed there)
        cudaError_t e = cudaSuccess;
        if (interior > 0) {
#define FXB_LAUNCH(S) e = launch_shape<S>(J, d, frame, state, pass, s0, iters, early_exit, run_all, ext_lo, ext_hi, first_brick, first_count, plain, pv, stream)
            switch (J.T) {
                case 1: if (J.narrow) FXB_LAUNCH(NarrowU<1>); else FXB_LAUNCH(WideU<1>); break;
                default: if (J.narrow) FXB_LAUNCH(NarrowU<2>); else FXB_LAUNCH(WideU<2>); break;
            }
#undef FXB_LAUNCH
            if (e != cudaSuccess) return e;
        }
        return launch_jacobi_pass_resident(J, d, frame, state, pass, iters, early_exit, run_all, ext_lo, ext_hi, face_first,
                                           faces * layer, pv, stream);
    }
#define FXB_LAUNCH(S) return launch_shape<S>(J, d, frame, state, pass, s0, iters, early_exit, run_all, ext_lo, ext_hi, first_brick, first_count, plain, pv, stream)
    switch (J.T) {
        case 1: if (J.narrow) FXB_LAUNCH(NarrowU<1>); FXB_LAUNCH(WideU<1>);
        case 2: if (J.narrow) FXB_LAUNCH(NarrowU<2>); FXB_LAUNCH(WideU<2>);
        case 3: if (J.narrow) FXB_LAUNCH(NarrowU<3>); FXB_LAUNCH(WideU<3>);
        case 4: if (J.narrow) FXB_LAUNCH(NarrowU<4>); FXB_LAUNCH(WideU<4>);
    }
#undef FXB_LAUNCH
    return cudaErrorInvalidValue;
}

cudaError_t launch_jacobi_settle(const FusedJacobi& J, const Domain& d, const FrameParams* frame, StepState* state,
                                 int iters, int force_passes, const PeerView& pv, cudaStream_t stream) {
    PassParams P{};
    P.nx = d.nx; P.ny = d.ny; P.pitch = d.pitch; P.nz_alloc = d.nz_alloc;
    P.z_out0 = d.z_own0 - d.z_first; P.z_out1 = d.z_own1 - d.z_first;
    P.bz = J.bz; P.ntx = J.ntx; P.nty = J.nty; P.nzc = J.nzc;
    P.levels_total = iters;
    const int npass = fused_jacobi_passes(J, iters);
    P.event = 2 + npass;
    P.push_depth = J.push_depth;
    JacobiPeers peers;
    for (int side = 0; side < 2; ++side)
        for (int i = 0; i < 2; ++i) {
            peers.p[side][i] = J.peer_p[side][i];
            peers.m[side][i] = J.peer_m[side][i];
        }
    WorkLists W;
    const int np = FusedJacobi::kMaxPasses + 1;
    W.relax[0] = J.work_list[0]; W.relax[1] = J.work_list[1];
    W.copy[0] = J.work_list[0] + J.list_stride; W.copy[1] = J.work_list[1] + J.list_stride;
    W.relax_count = J.work_count; W.copy_count = J.work_count + np;
    W.brick_flag = J.brick_flag;
    const int grid = J.num_sms * 2;
    // geometry of the own region only (shared by every shape of the schedule)
#define FXB_SETTLE(S)                                                                                                     \
    if (pv.has_lo || pv.has_hi)                                                                                           \
        jacobi_settle_kernel<S, true><<<grid, S::kThreads, 0, stream>>>(frame, state, J.p[0], J.p[1], J.mask[0], J.mask[1], W, P, J.T, J.tail_from, force_passes, pv, peers); \
    else                                                                                                                  \
        jacobi_settle_kernel<S, false><<<grid, S::kThreads, 0, stream>>>(frame, state, J.p[0], J.p[1], J.mask[0], J.mask[1], W, P, J.T, J.tail_from, force_passes, pv, peers)
    if (J.narrow) {
        switch (J.T) {
            case 1: FXB_SETTLE(NarrowU<1>); break;
            case 2: FXB_SETTLE(NarrowU<2>); break;
            case 3: FXB_SETTLE(NarrowU<3>); break;
            default: FXB_SETTLE(NarrowU<4>); break;
        }
    } else {
        switch (J.T) {
            case 1: FXB_SETTLE(WideU<1>); break;
            case 2: FXB_SETTLE(WideU<2>); break;
            case 3: FXB_SETTLE(WideU<3>); break;
            default: FXB_SETTLE(WideU<4>); break;
        }
    }
#undef FXB_SETTLE
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    jacobi_flip_kernel<<<1, 1, 0, stream>>>(frame, state, iters, pv, 2 + npass);
    return cudaGetLastError();
}

}  // namespace fxb
