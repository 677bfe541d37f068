// project_simple.cu — CSProject2D/3D split into one straightforward kernel per logical pass.
//
// Replaces FluidX12/Content/Shaders/CSProject3D.hlsl:68-113, CSProject2D.hlsl:64-106 and
// CSPoisson.hlsli:8-26 (dispatch Fluid.cpp:394-408).  The reference's in-place racy relaxation is
// restated as synchronous double-buffered Jacobi with a per-cell freeze flag.
// divergence and gradient kernels are shared with the tuned path; the one-sweep-per-launch Jacobi
// here is the cross-check path (fxb_config.kernel_path = 1) the fused kernels are tested against.
//
// rhs = -0.5 * s is stored instead of s: s is a sum of differences of fp16 values, hence a multiple
// of 2^-24, so the product is exact and `p[L] + rhs` is bit-identical to the DXBC's
// `mad(-s, 0.5, p[L])`.
#include "common.cuh"
#include "kernels.h"

namespace fxb {

namespace {

struct Nbr {
    size_t c, L, R, U, D, F, B;
};

// `stride`: elements per row of the array indexed — d.nx for velocity / colour, d.pitch for pressure / right-hand side.
__device__ __forceinline__ Nbr neighbours(const Domain& d, int x, int y, int z, int stride) {
    Nbr n;
    const size_t plane = (size_t)stride * d.ny;
    const size_t zc = (size_t)(z - d.z_first) * plane;
    const size_t row = zc + (size_t)y * stride;
    n.c = row + x;
    n.L = row + (max(x, 1) - 1);
    n.R = row + min(x + 1, d.nx - 1);
    n.U = zc + (size_t)(max(y, 1) - 1) * stride + x;
    n.D = zc + (size_t)min(y + 1, d.ny - 1) * stride + x;
    n.F = (size_t)(max(z, 1) - 1 - d.z_first) * plane + (size_t)y * stride + x;
    n.B = (size_t)(min(z + 1, d.nz - 1) - d.z_first) * plane + (size_t)y * stride + x;
    return n;
}

__device__ __forceinline__ float comp(const uint2* __restrict__ vel, size_t i, int c) {
    const unsigned short* h = reinterpret_cast<const unsigned short*>(vel + i);
    return half_bits_to_float(__ldg(h + c));
}

__global__ void begin_step_kernel(const FrameParams* __restrict__ frame, StepState* __restrict__ state, int iters,
                                  const __grid_constant__ PeerView pv) {
    const int k = threadIdx.x;
    if (k < iters && k < 128) state->active_after[k] = 0ull;
    if (k == 0) {
        state->s_exec = 0; state->passes = 0;
        phase_mark(state, 0);  // the advection phase ends here
        if (pv.has_lo || pv.has_hi) peer_publish(pv, frame->epoch_base + 1ull);  // fused halos: advect (m = 0) is complete
    }
}

__global__ void __launch_bounds__(256) divergence_kernel(Domain d, const FrameParams* __restrict__ frame,
                                                         const uint2* __restrict__ vel, float* __restrict__ rhs) {
    if (!(0.0f < frame->dt)) return;
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int z = d.z_own0 + blockIdx.z;
    if (x >= d.nx || y >= d.ny) return;
    const Nbr n = neighbours(d, x, y, z, d.nx);
    const float a = -comp(vel, n.L, 0) + comp(vel, n.R, 0);
    float b = -comp(vel, n.U, 1) + comp(vel, n.D, 1);
    float s;
    if (d.nz > 1) {
        b = b + a;
        const float c = -comp(vel, n.F, 2) + comp(vel, n.B, 2);
        s = c + b;
    } else {
        s = a + b;
    }
    rhs[((size_t)(z - d.z_first) * d.ny + y) * d.pitch + x] = -0.5f * s;
}

__global__ void __launch_bounds__(256) jacobi_sweep_simple_kernel(Domain d, const FrameParams* __restrict__ frame,
                                                                  const float* __restrict__ rhs, float* p0, float* p1,
                                                                  unsigned char* __restrict__ active,
                                                                  StepState* __restrict__ state, int sweep,
                                                                  int early_exit) {
    if (!(0.0f < frame->dt)) return;
    if (sweep > 0 && state->active_after[sweep - 1] == 0ull) return;
    const int sel = (state->p_cur + sweep) & 1;
    const float* __restrict__ in = sel ? p1 : p0;
    float* __restrict__ out = sel ? p0 : p1;

    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int z = d.z_own0 + blockIdx.z;
    int still = 0;
    if (x < d.nx && y < d.ny) {
        const Nbr n = neighbours(d, x, y, z, d.pitch);
        const bool act = sweep == 0 ? true : active[n.c] != 0;
        const float x0 = in[n.c];
        if (!act) {
            out[n.c] = x0;
        } else {
            const bool is3d = d.nz > 1;
            float acc = in[n.L] + rhs[n.c];
            acc = in[n.R] + acc;
            acc = in[n.U] + acc;
            acc = in[n.D] + acc;
            if (is3d) {
                acc = in[n.F] + acc;
                acc = in[n.B] + acc;
            }
            const float inv = is3d ? 0.166666672f : 0.25f;
            out[n.c] = acc * inv;
            still = 1;
            if (early_exit && fabsf(__fmaf_rn(acc, inv, -x0)) < 0.00100000005f) still = 0;
        }
        active[n.c] = (unsigned char)still;
    }
    const int cnt = __syncthreads_count(still);
    if (threadIdx.x == 0 && threadIdx.y == 0 && cnt) atomicAdd(&state->active_after[sweep], (unsigned long long)cnt);
}

// The per-sweep path flips the pressure ping-pong once per sweep.
__global__ void finish_solve_kernel(const FrameParams* __restrict__ frame, StepState* __restrict__ state, int iters) {
    if (threadIdx.x != 0) return;
    int s = 0;
    if (0.0f < frame->dt && iters > 0) {
        s = 1;
        while (s < iters && state->active_after[s - 1] != 0ull) ++s;
    }
    state->s_exec = s;
    state->passes = s;
    state->p_cur = (state->p_cur + s) & 1;
    state->total_sweeps += (unsigned long long)s;
    state->total_passes += (unsigned long long)s;
    phase_mark(state, 2);  // the pressure solve ends here
}

__global__ void __launch_bounds__(256) gradient_kernel(Domain d, const FrameParams* __restrict__ frame,
                                                       const uint2* __restrict__ vel_in, const float* p0,
                                                       const float* p1, uint2* __restrict__ vel_out,
                                                       const StepState* __restrict__ state) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int z = d.z_own0 + blockIdx.z;
    if (x >= d.nx || y >= d.ny) return;
    const Nbr nv = neighbours(d, x, y, z, d.nx), n = neighbours(d, x, y, z, d.pitch);
    const float4 v = load_texel4(vel_in, nv.c);
    float u[3] = {v.x, v.y, v.z};
    if (0.0f < frame->dt) {
        const float* __restrict__ p = state->p_cur ? p1 : p0;
        const float gx = -p[n.L] + p[n.R];
        const float gy = -p[n.U] + p[n.D];
        const float px = ((float)x + 0.5f) / (float)d.nx;
        const float py = ((float)y + 0.5f) / (float)d.ny;
        const float pz = ((float)z + 0.5f) / (float)d.nz;
        float bp[3];
        if (d.nz > 1) {
            const float gz = -p[n.F] + p[n.B];
            u[0] = __fmaf_rn(-gx, 1.04166675f, u[0]);
            u[1] = __fmaf_rn(-gy, 1.04166675f, u[1]);
            u[2] = __fmaf_rn(-gz, 1.04166675f, u[2]);
            bp[0] = __fmaf_rn(px, 2.0f, -1.0f);
            bp[1] = __fmaf_rn(py, 2.0f, -1.0f);
            bp[2] = __fmaf_rn(pz, 2.0f, -1.0f);
        } else {
            u[0] = __fmaf_rn(-gx, 0.5f, u[0]);
            u[1] = __fmaf_rn(-gy, 0.5f, u[1]);
            bp[0] = __fmaf_rn(px, 2.0f, -1.0f);
            bp[1] = __fmaf_rn(py, 2.0f, -1.0f);
            bp[2] = __fmaf_rn(pz, 1.0f, 0.0f);
        }
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float m = (-fabsf(bp[k]) + 0.970000029f) * 33.3333359f;
            m = fminf(fmaxf(m, -1.0f), 1.0f);
            if (!(0.0f < u[k] * bp[k])) m = 1.0f;
            u[k] = u[k] * m;
        }
    }
    vel_out[nv.c] = pack_texel4(u[0], u[1], u[2], 0.0f);
}

inline dim3 plane_grid(const Domain& d) { return dim3((d.nx + 31) / 32, (d.ny + 7) / 8, d.z_own1 - d.z_own0); }

}  // namespace

void launch_begin_step(const FrameParams* frame, StepState* state, int iters, const PeerView& pv, cudaStream_t stream) {
    begin_step_kernel<<<1, 128, 0, stream>>>(frame, state, iters, pv);
}

void launch_divergence(const Domain& d, const FrameParams* frame, const void* vel, float* rhs, cudaStream_t stream) {
    divergence_kernel<<<plane_grid(d), dim3(32, 8), 0, stream>>>(d, frame, (const uint2*)vel, rhs);
}

void launch_jacobi_sweep_simple(const Domain& d, const FrameParams* frame, const float* rhs, float* p0, float* p1,
                                unsigned char* active, StepState* state, int sweep, int early_exit,
                                cudaStream_t stream) {
    jacobi_sweep_simple_kernel<<<plane_grid(d), dim3(32, 8), 0, stream>>>(d, frame, rhs, p0, p1, active, state, sweep,
                                                                          early_exit);
}

void launch_finish_solve(const FrameParams* frame, StepState* state, int iters, cudaStream_t stream) {
    finish_solve_kernel<<<1, 32, 0, stream>>>(frame, state, iters);
}

void launch_gradient(const Domain& d, const FrameParams* frame, const void* vel_in, const float* p0, const float* p1,
                     void* vel_out, const StepState* state, cudaStream_t stream) {
    gradient_kernel<<<plane_grid(d), dim3(32, 8), 0, stream>>>(d, frame, (const uint2*)vel_in, p0, p1, (uint2*)vel_out,
                                                               state);
}

}  // namespace fxb
