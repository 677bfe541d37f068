// jacobi_tail.cu — block-resident multi-sweep Jacobi kernel for the tail of the pressure solve (sm_100a).
//
// Replaces, like jacobi_fused.cu, the relaxation loop of FluidX12/Content/Shaders/CSPoisson.hlsli:8-26 (called from
// CSProject3D.hlsl:93) under the deterministic restatement of SURVEY.md App. A.3.  Why a second kernel: after the first
// couple of sweeps only a few percent of the cells are still active (tools/tail_stats.py: 256^3 after 100 steps has
// 9.6 % active cells after sweep 2, 0.6 % after sweep 18, none after sweep 50), yet every further fused pass of the
// bulk kernel costs a launch plus the latency of one z-marching brick chain (~25-35 us measured), ~30 times per step.
// This kernel advances kSweeps = 4 sweeps per launch on 40 x 12 x 8 sub-blocks of the still-active bricks, each held
// on chip by one CTA for all four sweeps (body and data layout: jacobi_tail_body.cuh), so the tail needs half the
// launches of the T = 2 bulk kernel and the critical path of a launch is four short phases.
//
// Schedule ("dynamic", fxb_api.cu): bulk pass 0 always runs; afterwards tail launches and bulk passes are interleaved
// in the captured graph and decide on the device which of them does the work: a tail launch runs when at most
// `threshold` bricks are listed (or always, once the bulk passes of the schedule are used up), a bulk pass runs only
// when StepState::sweeps_done equals the sweep count its static index stands for.  Both kernels use the same work
// lists, freeze masks and ping-pong buffers, indexed by StepState::seq (relax kernels executed so far in the frame).
// Results are bit-identical whichever kernel relaxes a brick.
#include <cuda.h>

#include "common.cuh"
#include "jacobi_tail_body.cuh"
#include "kernels.h"

namespace fxb {

namespace {

using TailS = TailShape<4, 10, 12, 8>;  // 4 sweeps, sub-block 40 x 12 x 8, window 48 x 20 x 16, 256 threads
// Experimental pass-0 kernel (FXB_PASS0=2): the same body with 2 sweeps, every cell active and no flags to read; every
// window takes the second dense path.  Per quad update it executes about a third of the instructions of the z-marching
// bulk kernel (estimate from the SASS; unmeasured), at the price of loading each window with its halo (2.4x the cells).
using TailP0 = TailShape<2, 10, 12, 8>;  // window 48 x 16 x 12, 192 threads

struct TailLaunch {
    TailParams P;      // levels / first are filled in on the device
    int iters;         // ITER
    int threshold;     // run only when at most this many bricks are listed; < 0: always
    int run_all;       // multi-GPU: never end the solve on this rank's own freeze counters
    int first;         // 1: this launch is pass 0 of the frame (no lists, no flags: every brick, every cell active)
    int nbricks;
    int* list[2];      // [2 * bricks] per parity of seq: bricks to relax, then bricks to copy (jacobi_fused.cu)
    int* relax_count;  // [seq]
    int* copy_count;   // [seq]
    int* brick_state;
};

template <class S, int DENSE>
__global__ void __launch_bounds__(S::kThreads, S::kBytes > 80 * 1024 ? 2 : 3)
jacobi_tail_kernel(const FrameParams* __restrict__ frame, StepState* __restrict__ state, float* p0, float* p1,
                   const float* __restrict__ rhs, unsigned char* m0, unsigned char* m1,
                   const __grid_constant__ TailLaunch L, const __grid_constant__ CUtensorMap win0,
                   const __grid_constant__ CUtensorMap win1) {
    const float dt = frame->dt;
    const int seq = state->seq, s0 = state->sweeps_done, p_cur = state->p_cur;
    if (!(0.0f < dt)) return;
    if (L.first) {
        if (seq != 0 || s0 != 0 || L.iters <= 0) return;
    } else {
        if (seq == 0 || s0 <= 0 || s0 >= L.iters) return;         // pass 0 builds the lists; nothing left to do
        if (!L.run_all && state->active_after[s0 - 1] == 0ull) return;  // every cell is frozen: the solve is over
    }
    const int n_relax = L.first ? L.nbricks : L.relax_count[seq], n_copy = L.first ? 0 : L.copy_count[seq];
    if (!L.first && L.threshold >= 0 && n_relax > L.threshold) return;  // too many bricks: the bulk kernel is the better tool

    const TailParams& P = L.P;  // stays in the constant bank; only the sweep count is a run-time value
    const int levels = min(S::TT, L.iters - s0);
    const int sel = (p_cur + seq) & 1;
    const float* p_in = sel ? p1 : p0;
    float* p_out = sel ? p0 : p1;
    const unsigned char* m_in = (seq & 1) ? m1 : m0;
    unsigned char* m_out = (seq & 1) ? m0 : m1;
    TailWork W;
    W.relax_in = L.list[seq & 1];
    W.copy_in = L.list[seq & 1] + L.nbricks;
    W.n_relax = n_relax;
    W.n_copy = n_copy;
    W.relax_out = L.list[(seq + 1) & 1];
    W.copy_out = L.list[(seq + 1) & 1] + L.nbricks;
    W.relax_out_count = L.relax_count + seq + 1;
    W.copy_out_count = L.copy_count + seq + 1;
    W.brick_state = L.brick_state;

    extern __shared__ __align__(128) float tail_sm[];  // a TMA destination needs 128-byte alignment
    const TailShared<S> sh = tail_shared<S>(tail_sm);
    TailTma tma;
    tma.map = sel ? &win1 : &win0;
    tma.bar = reinterpret_cast<unsigned long long*>(sh.ctrl + S::kCtrlBar);
    if (P.cp_async == 2) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(tma.bar)), "r"(1)
                         : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
    }

    const int tid = threadIdx.x;
    const int items = n_copy + n_relax * P.nsub;
    unsigned relaxed = 0, dense = 0;  // sub-blocks of this CTA that held an active cell / took the dense path (thread 0)
    for (int item = blockIdx.x; item < items; item += gridDim.x) {
        if (item < n_copy) {  // a brick that froze in the previous kernel: one copy into the other buffer
            tail_copy_brick(tid, S::kThreads, P, W.copy_in[item], p_in, p_out, m_out);
            continue;
        }
        const int r = item - n_copy;
        const int path = tail_run_item<S, DENSE>(tid, sh, P, W, L.first ? r / P.nsub : W.relax_in[r / P.nsub], r % P.nsub,
                                                 p_in, p_out, rhs, m_in, m_out,
                                                 state->active_after + s0, state->active_after + 64, tma, levels);
        if (tid == 0 && path != 0) {
            ++relaxed;
            if (path == 2) ++dense;
        }
    }

    // the last CTA to finish advances the shared schedule (every CTA has read it long before)
    __syncthreads();
    if (tid == 0) {
        if (relaxed) atomicAdd(&state->tail_subblocks_relaxed, (unsigned long long)relaxed);
        if (dense) atomicAdd(&state->tail_subblocks_dense, (unsigned long long)dense);
        __threadfence();
        if (atomicAdd(&state->done_ctas, 1) == (int)gridDim.x - 1) {
            state->done_ctas = 0;
            state->seq = seq + 1;
            state->sweeps_done = s0 + levels;
            state->tail_launches += 1;
            state->tail_bricks += (unsigned long long)n_relax;
            state->bricks_processed += (unsigned long long)n_relax;  // one HBM pass of the brick, like a bulk pass
            state->bricks_copied += (unsigned long long)n_copy;
        }
    }
}

}  // namespace

int jacobi_tail_sweeps() { return TailS::TT; }

bool jacobi_tail_make_window_maps(FusedJacobi* J, const Domain& d) {
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap size");
    J->win_maps = fused_make_box_map(J->map_win[0], J->p[0], d.nx, d.ny, d.nz_alloc, TailS::LX, TailS::LY, TailS::LZ) &&
                  fused_make_box_map(J->map_win[1], J->p[1], d.nx, d.ny, d.nz_alloc, TailS::LX, TailS::LY, TailS::LZ);
    return J->win_maps;
}

bool jacobi_tail_supported(const FusedJacobi& J, const Domain& d) {
    int ext[3];
    fused_jacobi_brick_extent(J, ext);
    return (d.nx % 8) == 0 && ext[0] % TailS::OX == 0 && ext[0] % 8 == 0 && ext[1] <= TailS::OY && ext[2] <= TailS::OZ &&
           ext[0] / TailS::OX <= 200;
}

namespace {

// Everything of TailLaunch that does not depend on the shape.
TailLaunch tail_launch_params(const FusedJacobi& J, const Domain& d, int iters, int early_exit) {
    int ext[3];
    fused_jacobi_brick_extent(J, ext);
    TailLaunch L;
    L.P.nx = d.nx; L.P.ny = d.ny; L.P.nz_alloc = d.nz_alloc;
    L.P.z_face_lo = 0 - d.z_first;
    L.P.z_face_hi = d.nz - d.z_first;
    L.P.z_out0 = d.z_own0 - d.z_first; L.P.z_out1 = d.z_own1 - d.z_first;
    L.P.bx = ext[0]; L.P.by = ext[1]; L.P.bz = ext[2];
    L.P.ntx = J.ntx; L.P.nty = J.nty;
    L.P.nsub = ext[0] / TailS::OX;  // TailP0 has the same sub-block width
    L.P.first = 0; L.P.early_exit = early_exit; L.P.levels = 0;  // (levels: computed on the device)
    L.P.sparse_cap = 0;
    L.P.cp_async = J.tail_cp_async == 2 ? (J.win_maps ? 2 : 1) : (J.tail_cp_async ? 1 : 0);
    L.P.dense_mode = J.tail_dense_mode == 2 ? 2 : 1;
    L.iters = iters;
    L.threshold = -1;
    L.run_all = 0;
    L.first = 0;
    L.nbricks = J.ntx * J.nty * J.nzc;
    L.list[0] = J.work_list[0]; L.list[1] = J.work_list[1];
    const int np = FusedJacobi::kMaxPasses + 1;
    L.relax_count = J.work_count; L.copy_count = J.work_count + np;
    L.brick_state = J.brick_state;
    return L;
}

}  // namespace

cudaError_t launch_jacobi_tail(const FusedJacobi& J, const Domain& d, const FrameParams* frame, StepState* state,
                               int iters, int early_exit, int threshold, bool run_all, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(jacobi_tail_kernel<TailS, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)TailS::kBytes);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(jacobi_tail_kernel<TailS, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)TailS::kBytes);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    TailLaunch L = tail_launch_params(J, d, iters, early_exit);
    L.P.levels = TailS::TT;
    L.P.sparse_cap = J.tail_sparse_cap < 0 || J.tail_sparse_cap > TailS::kListCap ? TailS::kListCap : J.tail_sparse_cap;
    L.threshold = threshold;
    L.run_all = run_all ? 1 : 0;
    const CUtensorMap& w0 = *reinterpret_cast<const CUtensorMap*>(J.map_win[0]);
    const CUtensorMap& w1 = *reinterpret_cast<const CUtensorMap*>(J.map_win[1]);
    if (L.P.dense_mode == 2)
        jacobi_tail_kernel<TailS, 2><<<J.tail_grid, TailS::kThreads, TailS::kBytes, stream>>>(
            frame, state, J.p[0], J.p[1], J.rhs, J.mask[0], J.mask[1], L, w0, w1);
    else
        jacobi_tail_kernel<TailS, 1><<<J.tail_grid, TailS::kThreads, TailS::kBytes, stream>>>(
            frame, state, J.p[0], J.p[1], J.rhs, J.mask[0], J.mask[1], L, w0, w1);
    return cudaGetLastError();
}

// Pass 0 of the frame by the block-resident kernel (experimental, FXB_PASS0=2): every brick, every cell active.
cudaError_t launch_jacobi_pass0_tail(const FusedJacobi& J, const Domain& d, const FrameParams* frame, StepState* state,
                                     int iters, int early_exit, cudaStream_t stream) {
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(jacobi_tail_kernel<TailP0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)TailP0::kBytes);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    TailLaunch L = tail_launch_params(J, d, iters, early_exit);
    L.P.levels = TailP0::TT;
    L.P.sparse_cap = 0;      // nothing is sparse in pass 0
    L.P.dense_mode = 2;
    if (L.P.cp_async == 2) L.P.cp_async = 1;  // the window maps are built for the 4-sweep shape only
    L.first = 1;
    L.P.first = 1;
    const int items = L.nbricks * L.P.nsub;
    const int slots = J.num_sms * 12;  // three CTAs per SM resident; a few rounds per launch keep the tail short
    jacobi_tail_kernel<TailP0, 2><<<items < slots ? items : slots, TailP0::kThreads, TailP0::kBytes, stream>>>(
        frame, state, J.p[0], J.p[1], J.rhs, J.mask[0], J.mask[1], L, *reinterpret_cast<const CUtensorMap*>(J.map_win[0]),
        *reinterpret_cast<const CUtensorMap*>(J.map_win[1]));
    return cudaGetLastError();
}

}  // namespace fxb
