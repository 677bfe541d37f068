// kernels.h — host-side launchers of the fluidx_b200 CUDA kernels.
#pragma once

#include <cuda_runtime.h>

#include "common.cuh"

namespace fxb {

// advect.cu
// Fused halos (common.cuh PeerView): the neighbours' copies of the arrays a kernel also stores into, [0] = rank - 1,
// [1] = rank + 1 (nullptr at a grid face / on a single GPU).
struct AdvectPeers {
    void* vel_out[2] = {nullptr, nullptr};                       // m_velocities[1]
    void* col[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [side][m_colors index]
};
void launch_advect(const Domain& d, const AxisTables& tab, const FrameParams* frame, const void* vel_in,
                   void* const col[2], void* vel_out, const Emitter& em, int clamp_mode, StepState* state, int h_adv,
                   const PeerView& pv, const AdvectPeers& peers, cudaStream_t stream);

// lightmap.cu — the light-map pass after the step (CSRayMarchL); consts = fxb_light_params
struct HaloComm;
cudaError_t launch_light_map(const Domain& d, const void* colour_own, unsigned short* dens, unsigned* out,
                             const void* consts, HaloComm* comm, cudaStream_t stream);

cudaError_t launch_extract_density(const void* colour, unsigned short* dens, size_t n, cudaStream_t stream);

// raymarch.cu — the view-ray march into the cube map (CSRayMarchV / CSRayMarch); consts = fxb_view_params
cudaError_t launch_ray_march(const Domain& d, const void* colour, const unsigned short* dens, unsigned* cube,
                             const void* view, const void* light, cudaStream_t stream);
cudaError_t launch_ray_march_v(const Domain& d, const void* colour, const unsigned* light_map, unsigned* cube,
                               const void* consts, cudaStream_t stream);

// project_simple.cu — one kernel per logical pass (cross-check path, kernel_path = 1)
void launch_begin_step(const FrameParams* frame, StepState* state, int iters, const PeerView& pv, cudaStream_t stream);
void launch_divergence(const Domain& d, const FrameParams* frame, const void* vel, float* rhs, cudaStream_t stream);
void launch_jacobi_sweep_simple(const Domain& d, const FrameParams* frame, const float* rhs, float* p0, float* p1,
                                unsigned char* active, StepState* state, int sweep, int early_exit,
                                cudaStream_t stream);
void launch_finish_solve(const FrameParams* frame, StepState* state, int iters, cudaStream_t stream);
void launch_gradient(const Domain& d, const FrameParams* frame, const void* vel_in, const float* p0, const float* p1,
                     void* vel_out, const StepState* state, cudaStream_t stream);

// project_quad.cu — 4 cells per thread (tuned path; 3D grids with nx % 8 == 0)
bool quad_kernels_supported(const Domain& d);
void launch_divergence_quad(const Domain& d, const FrameParams* frame, const void* vel, float* rhs, const PeerView& pv,
                            float* rhs_lo, float* rhs_hi, int push_depth, cudaStream_t stream);
void launch_gradient_quad(const Domain& d, const AxisTables& tab, const FrameParams* frame, const void* vel_in,
                          const float* p0, const float* p1, void* vel_out, const StepState* state, const PeerView& pv,
                          void* vel_out_lo, void* vel_out_hi, int reach, int event, cudaStream_t stream);

// jacobi_fused.cu — T sweeps fused per HBM pass (tuned path, kernel_path = 0)
struct FusedJacobi {
    int T = 0;                 // sweeps fused per pass (1..4)
    bool copy_all = false;     // copy every brick that froze in the first pass (grouped multi-GPU exchange), not only those
                               // next to an active brick
    int* brick_flag = nullptr; // [bricks] per frame: bit 0 = froze in the first pass, bit 1 = next to a still-active brick
    // fused halos: the neighbours' copies of p[] / mask[] ([side][buffer]; nullptr at a grid face / on a single GPU)
    float* peer_p[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    unsigned char* peer_m[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
    bool narrow = false;       // tile 64 x 32 (a warp covers two row pairs) instead of 128 x 16
    int tile_x = 128, tile_y = 16;
    int ntx = 0, nty = 0, nzc = 0, bz = 0;  // brick grid and planes per brick
    float* p[2] = {nullptr, nullptr};
    float* rhs = nullptr;
    unsigned char* mask[2] = {nullptr, nullptr};  // bit-packed freeze flags, ping-pong by pass parity
    int* work_list[2] = {nullptr, nullptr};       // [2 * list_stride] per pass parity: bricks to relax, then bricks to copy
    int list_stride = 0;                          // entries per list (bricks + padding for the kernel's look-ahead)
    int* work_count = nullptr;                    // [3][kMaxPasses + 1]: relax count and copy count per pass (+ spare)
    int num_sms = 0;
    static constexpr int kMaxPasses = 130;
    alignas(64) unsigned char map_p[2][128];      // CUtensorMap of each pressure buffer
    alignas(64) unsigned char map_rhs[128];
    // brick-resident passes (jacobi_resident.cu): tensor maps whose box is a brick's whole window (tile x tile x 8 + 2T
    // planes), and the first pass of a frame that runs in that form (kMaxPasses + 1: none)
    alignas(64) unsigned char map3_p[2][128];
    alignas(64) unsigned char map3_rhs[128];
    int resident_from = kMaxPasses + 1;
    bool pdl = true;           // resident passes are launched with programmatic stream serialization
    // Tail schedule: from pass `tail_from` on a pass fuses FOUR sweeps and runs on half bricks (jacobi_resident.cu
    // RShapeHalf4; tensor maps with its box); kMaxPasses + 1: none.  `push_depth`: planes next to an interior slab face
    // every pass (and the divergence) also stores into the neighbour: the deepest halo any pass of the schedule reads.
    alignas(64) unsigned char map4_p[2][128];
    alignas(64) unsigned char map4_rhs[128];
    int tail_from = kMaxPasses + 1;
    int push_depth = 0;
};
// The schedule of a frame: sweeps completed before pass k, sweeps pass k fuses, passes needed for `iters` sweeps.
inline int fused_jacobi_pass_t(const FusedJacobi& J, int pass) { return pass >= J.tail_from ? 4 : J.T; }
inline int fused_jacobi_s0(const FusedJacobi& J, int pass) {
    return pass <= J.tail_from ? pass * J.T : J.tail_from * J.T + (pass - J.tail_from) * 4;
}
int fused_jacobi_passes(const FusedJacobi& J, int iters);  // launches per frame
bool fused_jacobi_supported(const Domain& d);
// allow_tail: the four-sweep tail schedule may be used (single GPU or fused halos: nothing on the host separates passes)
int fused_jacobi_plan(FusedJacobi* J, const Domain& d, int fuse_t, float* p0, float* p1, float* rhs, bool allow_tail);
size_t fused_jacobi_bricks(const FusedJacobi& J);
size_t fused_jacobi_brick_cells(const FusedJacobi& J);
void fused_jacobi_brick_extent(const FusedJacobi& J, int out[3]);
// After the last pass: copies what the last executed pass left in the wrong buffer, s_exec / pass count, flips p_cur.
cudaError_t launch_jacobi_settle(const FusedJacobi& J, const Domain& d, const FrameParams* frame, StepState* state,
                                 int iters, int force_passes, const PeerView& pv, cudaStream_t stream);
cudaError_t launch_jacobi_pass_fused(const FusedJacobi& J, const Domain& d, const FrameParams* frame, StepState* state,
                                     int pass, int iters, int early_exit, bool run_all_passes, int ext_lo, int ext_hi,
                                     const PeerView& pv, cudaStream_t stream);

// jacobi_resident.cu — the same pass with the brick's window resident in shared memory (the latency shape)
bool resident_jacobi_supported(const FusedJacobi& J);
int resident_jacobi_window_planes(const FusedJacobi& J);
// first_count >= 0 (first pass only): the launch relaxes the bricks (first_brick + w) mod bricks, w < first_count.
cudaError_t launch_jacobi_pass_resident(const FusedJacobi& J, const Domain& d, const FrameParams* frame, StepState* state,
                                        int pass, int iters, int early_exit, bool run_all_passes, int ext_lo, int ext_hi,
                                        int first_brick, int first_count, const PeerView& pv, cudaStream_t stream);
// Kernels launch_jacobi_pass_fused enqueues for this pass (2 for the first pass with fused halos: interior + face layers).
int fused_jacobi_launches(const FusedJacobi& J, const PeerView& pv, int pass);

}  // namespace fxb
