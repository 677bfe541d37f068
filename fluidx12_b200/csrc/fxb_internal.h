// fxb_internal.h — what the translation units behind the C ABI share: the simulator handle, the per-thread error
// string and the CUDA error macro.  Not installed; include/fluidx_b200.h is the public boundary.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/fluidx_b200.h"
#include "common.cuh"
#include "halo.h"
#include "kernels.h"

namespace fxb {
// Stores `msg` as the calling thread's last error (fxb_last_error) and returns `code`.
int api_fail(int code, const std::string& msg);
}  // namespace fxb

#define FXB_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e_ = (expr);                                                                    \
        if (e_ != cudaSuccess)                                                                      \
            return fxb::api_fail(FXB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_)); \
    } while (0)

struct fxb_sim {
    fxb_config cfg{};
    fxb::Domain dom{};
    int fuse_t = 1;
    int parity = 0;  // m_frameParity (Fluid.h:124)
    float dt = 0.0f;  // m_timeStep (Fluid.h:126)
    uint64_t steps = 0;
    int kernels_per_step = 0;

    void* vel[2] = {nullptr, nullptr};  // m_velocities (Fluid.h:94), RGBA16F
    void* col[2] = {nullptr, nullptr};  // m_colors (Fluid.h:95), RGBA16F
    float* p[2] = {nullptr, nullptr};   // m_incompress (Fluid.h:93), R32F, ping-pong
    float* rhs = nullptr;               // -0.5 * (2*divergence)
    unsigned char* active = nullptr;    // per-cell freeze flags of the simple path
    bool fused = false;                 // tuned Jacobi path in use
    fxb::FusedJacobi jac;
    float* emitter_basis = nullptr;
    float* axis_tables = nullptr;       // pos / bp / wall per axis, one allocation
    fxb::AxisTables tab{};
    bool quad = false;                  // 4-cells-per-thread divergence / gradient kernels in use
    fxb::Emitter emitter{};
    fxb::FrameParams* d_frame = nullptr;
    fxb::StepState* d_state = nullptr;

    cudaStream_t own_stream = nullptr;
    cudaStream_t last_stream = nullptr;
    cudaStream_t side_stream = nullptr;  // multi-GPU: the all-reduce of the freeze counters runs beside the step's tail
    bool side_forked = false;
    // One captured graph per (frame parity, pressure-buffer parity): both select pointers that the halo exchange
    // of the multi-GPU step needs on the host side.  Single GPU uses slot [0][0] only (its kernels select on device).
    cudaGraph_t graph[2][2] = {};
    cudaGraphExec_t graph_exec[2][2] = {};
    fxb::HaloComm comm;       // z-slab neighbours (nranks > 1)
    fxb::PeerView pv{};       // fused halos (FXB_HALO_FUSED): neighbours' event words and plane offsets; zero otherwise
    fxb::AdvectPeers advect_peers;
    bool halo_stale = false;  // fxb_set_field since the last step: the neighbours' halo copies need one plain exchange
    int halo = 0;             // halo planes allocated on interior faces
    int h_adv = 0;            // advection halo (back-trace reach in planes)
    int jacobi_group = 1;     // multi-GPU: fused passes per pressure-halo exchange (fxb_config.jacobi_group)
    int p_cur_host = 0;       // host mirror of StepState::p_cur (multi-GPU: the pass count per step is fixed)
    unsigned* light_map = nullptr;         // m_lightMap (Fluid.h), R11G11B10_FLOAT words; allocated by fxb_light_map
    unsigned* cube_map = nullptr;          // one mip of m_cubeMap (Fluid.cpp:229-232): [6][S][S] RGBA8 words
    uint32_t cube_size = 0;
    // fxb_post_stats / fxb_wait_stats: a ring of pinned snapshots of StepState with one event each, so that a frame
    // loop can read step k's record while step k + 1 runs (the reference keeps FrameCount = 3 frames in flight, Fluid.h:35)
    static constexpr int kStatsSlots = 4;
    fxb::StepState* stats_ring = nullptr;  // pinned host memory, kStatsSlots entries
    cudaEvent_t stats_event[kStatsSlots] = {};
    uint64_t stats_steps[kStatsSlots] = {};
    int stats_parity[kStatsSlots] = {};
    void* whole_colour = nullptr;          // nranks > 1: the colour field / light map of the WHOLE grid, gathered from all
    unsigned* whole_light_map = nullptr;   // ranks for the view-ray march (a view ray crosses every z-slab)
    unsigned short* light_density = nullptr;  // colour.w of every voxel, the channel the light-map pass samples
    bool multi() const { return cfg.nranks > 1; }
    cudaEvent_t ev[8] = {};

    size_t plane_voxels() const { return (size_t)dom.nx * dom.ny; }
    size_t alloc_voxels() const { return plane_voxels() * dom.nz_alloc; }
    size_t own_voxels() const { return plane_voxels() * (dom.z_own1 - dom.z_own0); }
    size_t own_offset() const { return plane_voxels() * (dom.z_own0 - dom.z_first); }
};

namespace fxb {
// Device pointer (plane 0 = the rank's first allocated plane) and element size of a field of include/fluidx_b200.h's
// fxb_field; nullptr with *err set for an unknown id.
void* field_device_ptr(fxb_sim* s, int field, size_t* elem_bytes, int* err);
}  // namespace fxb
