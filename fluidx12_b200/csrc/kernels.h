// kernels.h — host-side launchers of the fluidx_b200 CUDA kernels.
#pragma once

#include <cuda_runtime.h>

#include "common.cuh"

namespace fxb {

// advect.cu
void launch_advect(const Domain& d, const FrameParams* frame, const void* vel_in, void* const col[2], void* vel_out,
                   const Emitter& em, int clamp_mode, StepState* state, cudaStream_t stream);

// project_simple.cu — one kernel per logical pass (cross-check path, kernel_path = 1)
void launch_begin_step(const FrameParams* frame, StepState* state, int iters, cudaStream_t stream);
void launch_divergence(const Domain& d, const FrameParams* frame, const void* vel, float* rhs, cudaStream_t stream);
void launch_jacobi_sweep_simple(const Domain& d, const FrameParams* frame, const float* rhs, float* p0, float* p1,
                                unsigned char* active, StepState* state, int sweep, int early_exit,
                                cudaStream_t stream);
void launch_finish_solve(const FrameParams* frame, StepState* state, int iters, int fuse_t, cudaStream_t stream);
void launch_gradient(const Domain& d, const FrameParams* frame, const void* vel_in, const float* p0, const float* p1,
                     void* vel_out, const StepState* state, cudaStream_t stream);

}  // namespace fxb
