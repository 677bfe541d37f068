// fxb_api.cu — C ABI and host-side step driver of fluidx_b200 (see include/fluidx_b200.h).
//
// Host mirror of the simulation half of the reference's Fluid class:
//   Fluid::Init        FluidX12/Content/Fluid.cpp:189-270  -> fxb_create
//   Fluid::UpdateFrame FluidX12/Content/Fluid.cpp:283-346  -> fxb_update_frame
//   Fluid::Simulate    FluidX12/Content/Fluid.cpp:348-410  -> fxb_simulate
// The XUSG Texture3D resources become plain device buffers, the per-frame constant buffer becomes a
// device-resident FrameParams written by a one-thread kernel whose arguments carry dt and the frame
// parity (so a single captured CUDA graph serves every frame), and the two dispatches become the
// captured graph advect -> divergence -> Jacobi passes -> gradient-subtract.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "fxb_internal.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string& msg) { return fxb::api_fail(code, msg); }

__global__ void set_frame_kernel(fxb::FrameParams* frame, fxb::StepState* state, float dt, int parity,
                                 unsigned long long frame_index) {
    frame->dt = dt;
    frame->parity = parity;
    frame->epoch_base = frame_index * fxb::kEventsPerFrame;  // fused halos: the event numbers of this frame's kernels
    fxb::phase_mark(state, -1);  // a step starts here
}

}  // namespace

int fxb::api_fail(int code, const std::string& msg) {
    g_last_error = msg;
    return code;
}


namespace {

// Emitter table (Impulse.hlsli:14-15, CSAdvect.hlsl:58-60): basis values of the voxels in the conservative box,
// computed exactly as the shader orders the arithmetic with libm exp2f.
// Voxel box [lo, hi) that contains every voxel whose emitter basis can reach exp(-4): the sphere
// |pos - (0.5, 0.1, 0.5)| <= r with a 0.1 % radius margin plus one voxel per side, clipped to the grid.
void emitter_box(int nx, int ny, int nz, int lo[3], int hi[3]) {
    const bool is3d = nz > 1;
    const float r = is3d ? 1.0f / 16.0f : 1.0f / 32.0f;
    const float centre[3] = {0.5f, 0.1f, 0.5f};
    const int n[3] = {nx, ny, nz};
    for (int a = 0; a < 3; ++a) {
        lo[a] = (int)std::floor((centre[a] - r * 1.001f) * n[a] - 0.5f) - 1;
        hi[a] = (int)std::ceil((centre[a] + r * 1.001f) * n[a] - 0.5f) + 2;
        lo[a] = std::max(lo[a], 0);
        hi[a] = std::min(hi[a], n[a]);
        if (hi[a] < lo[a]) hi[a] = lo[a];
    }
    if (!is3d) { lo[2] = 0; hi[2] = 1; }
}

int build_emitter(fxb_sim* s) {
    const int nx = s->dom.nx, ny = s->dom.ny, nz = s->dom.nz;
    int lo[3], hi[3];
    emitter_box(nx, ny, nz, lo, hi);
    fxb::Emitter& em = s->emitter;
    em.x0 = lo[0]; em.y0 = lo[1]; em.z0 = lo[2];
    em.x1 = hi[0]; em.y1 = hi[1]; em.z1 = hi[2];
    const size_t ex = em.x1 - em.x0, ey = em.y1 - em.y0, ez = em.z1 - em.z0;
    std::vector<float> table(std::max<size_t>(ex * ey * ez, 1), 0.0f);
    const float r2 = (1.0f < (float)nz) ? 0.00390625f : 0.0009765625f;
    for (size_t kz = 0; kz < ez; ++kz)
        for (size_t ky = 0; ky < ey; ++ky)
            for (size_t kx = 0; kx < ex; ++kx) {
                const float px = ((float)(em.x0 + (int)kx) + 0.5f) / (float)nx;
                const float py = ((float)(em.y0 + (int)ky) + 0.5f) / (float)ny;
                const float pz = ((float)(em.z0 + (int)kz) + 0.5f) / (float)nz;
                const float dx = px + -0.5f, dy = py + -0.100000001f, dz = pz + -0.5f;
                const float d2 = (dx * dx + dy * dy) + dz * dz;
                const float e = ((d2 * -4.0f) / r2) * 1.44269502f;
                table[(kz * ey + ky) * ex + kx] = exp2f(e);
            }
    FXB_CUDA(cudaMalloc(&s->emitter_basis, table.size() * sizeof(float)));
    FXB_CUDA(cudaMemcpy(s->emitter_basis, table.data(), table.size() * sizeof(float), cudaMemcpyHostToDevice));
    em.basis = s->emitter_basis;
    return FXB_OK;
}

// Per-axis tables (see fxb::AxisTables).  Same arithmetic and operation order as the shaders (App. A.1/A.2).
int build_axis_tables(fxb_sim* s) {
    const int n[3] = {s->dom.nx, s->dom.ny, s->dom.nz};
    const bool is3d = s->dom.nz > 1;
    size_t off[3], total = 0;
    for (int a = 0; a < 3; ++a) { off[a] = total; total += ((size_t)n[a] + 3) / 4 * 4; }
    std::vector<float> host(4 * total, 0.0f);
    for (int a = 0; a < 3; ++a)
        for (int i = 0; i < n[a]; ++i) {
            const float pos = ((float)i + 0.5f) / (float)n[a];
            const float bp = (a == 2 && !is3d) ? fmaf(pos, 1.0f, 0.0f) : fmaf(pos, 2.0f, -1.0f);
            float w = (-fabsf(bp) + 0.970000029f) * 33.3333359f;
            w = fminf(fmaxf(w, -1.0f), 1.0f);
            host[off[a] + i] = pos;
            host[total + off[a] + i] = bp;
            host[2 * total + off[a] + i] = w;
            // (a == 2 on a 2D grid: W = 1, the trace is fma(pos, 1, -0.5) = 0 and the single plane has no second tap)
            host[3 * total + off[a] + i] = (fmaf(pos, (float)n[a], -0.5f) == (float)i && i <= n[a] - 2) ? 1.0f : 0.0f;
        }
    FXB_CUDA(cudaMalloc((void**)&s->axis_tables, host.size() * sizeof(float)));
    FXB_CUDA(cudaMemcpy(s->axis_tables, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice));
    for (int a = 0; a < 3; ++a) {
        s->tab.pos[a] = s->axis_tables + off[a];
        s->tab.bp[a] = s->axis_tables + total + off[a];
        s->tab.wall[a] = s->axis_tables + 2 * total + off[a];
        s->tab.still[a] = s->axis_tables + 3 * total + off[a];
    }
    return FXB_OK;
}

// Pressure ping-pong flips of one (dt > 0) step as the HOST must know them (multi-GPU: they select the buffers whose
// halos are exchanged): every rank runs every pass, and every pass flips once.
int flips_per_step(const fxb_sim* s) {
    return s->fused && s->cfg.jacobi_iters > 0 ? 1 : 0;  // to the first pass's output buffer (jacobi_settle_kernel)
}

enum Phase { PH_ADVECT = 0, PH_DIVERGENCE, PH_JACOBI, PH_GRADIENT, PH_COUNT };

// Phase marks (common.cuh): with fxb_config.phase_timing a one-thread kernel closes the divergence and the gradient
// phase (the other marks ride on the step's own one-thread kernels), so the per-phase device times of the very steps a
// caller times are available afterwards (fxb_get_phase_times) — also when the step runs as a captured graph.
// With fused halos the same kernel publishes the event of the kernel before it (divergence: epoch_base + 2, gradient:
// the next frame's base).
__global__ void phase_mark_kernel(fxb::StepState* state, int slot, const fxb::FrameParams* frame,
                                  const __grid_constant__ fxb::PeerView pv, unsigned long long event_offset) {
    if (slot >= 0) fxb::phase_mark(state, slot);
    if (pv.has_lo || pv.has_hi) fxb::peer_publish(pv, frame->epoch_base + event_offset);
#ifdef FXB_TIMING
    if (event_offset == 2) {  // debug build: the divergence ended = the pressure solve starts (tools/mgpu_probe.py)
        unsigned long long now;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
        state->dbg[106] = (long long)now;
    }
#endif
}

// State checksum (fxb_state_checksum): per field the wrap-around sum over the rank's own voxels of a 64-bit mix of the
// voxel's GLOBAL linear index and its bits, so the sums of all ranks of a z-slab run add up to the single-GPU value.
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return x;
}
__global__ void __launch_bounds__(256) checksum_kernel(fxb::Domain d, const uint2* __restrict__ vel,
                                                       const uint2* __restrict__ col, const float* __restrict__ p,
                                                       unsigned long long* __restrict__ out) {
    const size_t n = (size_t)d.nx * d.ny * (d.z_own1 - d.z_own0);
    unsigned long long a = 0, b = 0, c = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const int x = (int)(i % d.nx), y = (int)((i / d.nx) % d.ny), lz = (int)(i / ((size_t)d.nx * d.ny));
        const size_t row = (size_t)(d.z_own0 - d.z_first + lz) * d.ny + y;
        const size_t at = row * d.nx + x, at_p = row * d.pitch + x;
        const unsigned long long g = ((unsigned long long)(d.z_own0 + lz) * d.ny + y) * d.nx + x;
        const uint2 v = vel[at], k = col[at];
        a += mix64(g * 3 + 0 + (((unsigned long long)(v.y & 0xffffu) << 32 | v.x) << 20));  // velocity .w is a don't-care
        b += mix64(g * 3 + 1 + (((unsigned long long)k.y << 32 | k.x) << 20) + (unsigned long long)(k.y >> 12));
        c += mix64(g * 3 + 2 + ((unsigned long long)__float_as_uint(p[at_p]) << 24));
    }
    for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_down_sync(0xffffffffu, a, o);
        b += __shfl_down_sync(0xffffffffu, b, o);
        c += __shfl_down_sync(0xffffffffu, c, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(&out[0], a);
        atomicAdd(&out[1], b);
        atomicAdd(&out[2], c);
    }
}

// Multi-GPU: the global number of sweeps of the frame from the summed freeze counters (jacobi_settle_kernel leaves
// s_exec alone there: it only sees this rank's counters).
__global__ void global_sweeps_kernel(const fxb::FrameParams* frame, fxb::StepState* state, int iters) {
    int s = 0;
    if (0.0f < frame->dt && iters > 0) {
        s = 1;
        while (s < iters && state->active_after[s - 1] != 0ull) ++s;
    }
    state->s_exec = s;
    state->total_sweeps += (unsigned long long)s;
}

struct Enqueue {
    fxb_sim* s;
    cudaStream_t st;
    int launches = 0;
    bool ok = true;
    std::string err;
    void halo(const fxb::HaloField* f, int n) {
        if (fused()) return;  // every kernel stores its face planes into the neighbours itself (common.cuh PeerView)
        if (ok && !s->comm.exchange(s->dom, f, n, st)) { ok = false; err = "halo exchange: " + fxb::halo_last_error(); }
        if (s->comm.p2p.enabled) ++launches;
    }
    void launched(cudaError_t e, const char* what, int n = 1) {
        if (ok && e != cudaSuccess) { ok = false; err = std::string(what) + ": " + cudaGetErrorString(e); }
        launches += n;
    }
    bool fused() const { return s->multi() && s->cfg.halo_backend == FXB_HALO_FUSED; }
    // closes a phase: the phase mark (fxb_config.phase_timing) and, with fused halos, the event of the kernel before it
    void mark(int slot, unsigned long long event_offset) {
        if (!s->cfg.phase_timing && !fused()) return;
        phase_mark_kernel<<<1, 1, 0, st>>>(s->d_state, s->cfg.phase_timing ? slot : -1, s->d_frame, s->pv, event_offset);
        launched(cudaGetLastError(), "phase_mark_kernel");
    }
};

// Enqueues one phase of the step.
void enqueue_phase(Enqueue& q, int phase) {
    fxb_sim* s = q.s;
    cudaStream_t st = q.st;
    const fxb::Domain& d = s->dom;
    switch (phase) {
        case PH_ADVECT:
            if (s->multi()) {  // back-trace reach + 1 tap of the inputs
                const fxb::HaloField f[2] = {{s->vel[0], s->plane_voxels() * 8, s->h_adv + 1},
                                             {s->col[!s->parity], s->plane_voxels() * 8, s->h_adv + 1}};
                q.halo(f, 2);
            }
            // Fluid.cpp:358-375: vel[0], colour[!p] -> vel[1], colour[p]
            fxb::launch_advect(d, s->tab, s->d_frame, s->vel[0], s->col, s->vel[1], s->emitter, s->cfg.address_mode,
                               s->d_state, s->h_adv, s->pv, s->advect_peers, st);
            // also the advection phase's mark and, with fused halos, the event "advect complete"
            fxb::launch_begin_step(s->d_frame, s->d_state, s->cfg.jacobi_iters, s->pv, st);
            q.launched(cudaGetLastError(), "advect_kernel", 2);
            break;
        case PH_DIVERGENCE:
            if (s->multi() && s->dt > 0.0f) {  // z neighbours of the advected velocity
                const fxb::HaloField f[1] = {{s->vel[1], s->plane_voxels() * 8, 1}};
                q.halo(f, 1);
            }
            if (s->quad)
                fxb::launch_divergence_quad(d, s->d_frame, s->vel[1], s->rhs, s->pv, (float*)s->comm.peer_of(s->rhs, 0),
                                            (float*)s->comm.peer_of(s->rhs, 1), s->jac.push_depth, st);
            else fxb::launch_divergence(d, s->d_frame, s->vel[1], s->rhs, st);
            q.launched(cudaGetLastError(), "divergence_kernel");
            q.mark(1, 2);
            break;
        case PH_JACOBI:
            if (s->fused) {
                const int npass = fxb::fused_jacobi_passes(s->jac, s->cfg.jacobi_iters);
                if (cudaMemsetAsync(s->jac.work_count, 0, 3 * (fxb::FusedJacobi::kMaxPasses + 1) * sizeof(int), st) != cudaSuccess ||
                    cudaMemsetAsync(s->jac.brick_flag, 0, 2 * fxb::fused_jacobi_bricks(s->jac) * sizeof(int), st) != cudaSuccess)
                    q.launched(cudaGetLastError(), "cudaMemsetAsync(work lists)", 0);
                const bool mg = s->multi() && s->dt > 0.0f;
                // Multi-GPU: the pressure (+ freeze flag) halo is exchanged every G passes, G*T planes deep; in
                // between, pass j of a group also relaxes the (G-1-j)*T halo planes next to each interior face.
                const int G = s->multi() ? std::max(1, std::min(s->jacobi_group, s->halo / s->fuse_t)) : 1;
                if (mg) {  // the right-hand side is constant over the sweeps: one exchange, as deep as the group
                    const fxb::HaloField f[1] = {{s->rhs, s->plane_voxels() * 4, G * s->fuse_t}};
                    q.halo(f, 1);
                }
                for (int k = 0; k < npass; ++k) {
                    int ext_lo = 0, ext_hi = 0;
                    if (mg) {
                        const int Tk = s->fuse_t;
                        if (k % G == 0) {
                            const fxb::HaloField f[2] = {
                                {s->p[(s->p_cur_host + k) & 1], s->plane_voxels() * 4, G * Tk},
                                {s->jac.mask[k & 1], s->plane_voxels() / 8, G * Tk}};
                            q.halo(f, k == 0 ? 1 : 2);
                        }
                        const int ext = (G - 1 - k % G) * Tk;
                        ext_lo = s->cfg.rank > 0 ? ext : 0;
                        ext_hi = s->cfg.rank < s->cfg.nranks - 1 ? ext : 0;
                    }
                    q.launched(fxb::launch_jacobi_pass_fused(s->jac, d, s->d_frame, s->d_state, k, s->cfg.jacobi_iters,
                                                             s->cfg.early_exit, s->multi(), ext_lo, ext_hi, s->pv, st),
                               "jacobi_pass_kernel", fxb::fused_jacobi_launches(s->jac, s->pv, k));
                }
                if (mg || q.fused()) {  // (fused halos: ONE graph serves every frame, paused ones included)
                    // Freeze counters are per rank: their sum over the ranks (NCCL all-reduce, in place) gives the global
                    // s_exec.  Nothing on the data path needs it — every rank runs all passes, so the buffer parity stays
                    // in step — hence it runs on a side stream beside the settle step and the gradient, and joins the
                    // step's stream after the gradient (the next frame resets the counters).
                    // (With the NCCL halo backend the communicator is busy on the step's stream: no second stream then.)
                    const bool fork = s->cfg.halo_backend != FXB_HALO_NCCL;
                    cudaStream_t ar = fork ? s->side_stream : st;
                    if (fork) {
                        cudaError_t e = cudaEventRecord(s->ev[6], st);
                        if (e == cudaSuccess) e = cudaStreamWaitEvent(ar, s->ev[6], 0);
                        if (e != cudaSuccess) q.launched(e, "fork to the side stream", 0);
                    }
                    if (q.ok && !s->comm.all_reduce_sum_u64(s->d_state->active_after, 128, ar)) {
                        q.ok = false;
                        q.err = "all-reduce of the freeze counters: " + fxb::halo_last_error();
                    }
                    global_sweeps_kernel<<<1, 1, 0, ar>>>(s->d_frame, s->d_state, s->cfg.jacobi_iters);
                    q.launched(cudaGetLastError(), "global_sweeps_kernel");
                    if (fork) {
                        q.launched(cudaEventRecord(s->ev[7], ar), "cudaEventRecord", 0);
                        s->side_forked = true;
                    }
                }
                q.launched(fxb::launch_jacobi_settle(s->jac, d, s->d_frame, s->d_state, s->cfg.jacobi_iters,
                                                     s->multi() ? npass : -1, s->pv, st), "jacobi_settle_kernel", 2);
                if (mg) {  // z neighbours of the final pressure (the first pass's output buffer) for the gradient
                    const fxb::HaloField f[1] = {{s->p[(s->p_cur_host + 1) & 1], s->plane_voxels() * 4, 1}};
                    q.halo(f, 1);
                }
            } else {
                for (int k = 0; k < s->cfg.jacobi_iters; ++k)
                    fxb::launch_jacobi_sweep_simple(d, s->d_frame, s->rhs, s->p[0], s->p[1], s->active, s->d_state, k,
                                                    s->cfg.early_exit, st);
                fxb::launch_finish_solve(s->d_frame, s->d_state, s->cfg.jacobi_iters, st);
                q.launched(cudaGetLastError(), "jacobi_sweep_simple_kernel", s->cfg.jacobi_iters + 1);
            }
            break;
        case PH_GRADIENT:
            // Fluid.cpp:378-408: vel[1] -> vel[0]
            if (s->quad)
                fxb::launch_gradient_quad(d, s->tab, s->d_frame, s->vel[1], s->p[0], s->p[1], s->vel[0], s->d_state, s->pv,
                                          s->comm.peer_of(s->vel[0], 0), s->comm.peer_of(s->vel[0], 1), s->h_adv + 1,
                                          3 + (s->fused ? fxb::fused_jacobi_passes(s->jac, s->cfg.jacobi_iters) : 0), st);
            else
                fxb::launch_gradient(d, s->d_frame, s->vel[1], s->p[0], s->p[1], s->vel[0], s->d_state, st);
            q.launched(cudaGetLastError(), "gradient_kernel");
            q.mark(3, fxb::kEventsPerFrame);
            if (s->side_forked) {  // the side stream (global freeze counters) joins here
                q.launched(cudaStreamWaitEvent(st, s->ev[7], 0), "join of the side stream", 0);
                s->side_forked = false;
            }
            break;
    }
}

// The whole step: advect -> divergence -> Jacobi passes -> gradient-subtract, with a phase mark after each.
// Returns FXB_OK or the first failure (NCCL or launch); `launches` receives the number of kernels enqueued.
int enqueue_step(fxb_sim* s, cudaStream_t st, int* launches) {
    Enqueue q{s, st};
    for (int ph = 0; ph < PH_COUNT && q.ok; ++ph) enqueue_phase(q, ph);
    if (launches) *launches = q.launches;
    if (!q.ok) return fail(s->multi() && q.err.compare(0, 4, "halo") == 0 ? FXB_ERR_NCCL : FXB_ERR_CUDA, "fxb_simulate: " + q.err);
    return FXB_OK;
}

// Captures the step for the current (frame parity, pressure parity) key; single GPU always uses key [0][0].
int capture_graph(fxb_sim* s, int a, int b) {
    FXB_CUDA(cudaStreamBeginCapture(s->own_stream, cudaStreamCaptureModeThreadLocal));
    int launches = 0;
    const int rc = enqueue_step(s, s->own_stream, &launches);
    cudaGraph_t g = nullptr;
    cudaError_t e = cudaStreamEndCapture(s->own_stream, &g);  // always end the capture, also after a failure
    if (rc != FXB_OK) {
        if (g) cudaGraphDestroy(g);
        return rc;
    }
    if (e != cudaSuccess) return fail(FXB_ERR_CUDA, std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    s->graph[a][b] = g;
    FXB_CUDA(cudaGraphInstantiate(&s->graph_exec[a][b], s->graph[a][b], 0));
    s->kernels_per_step = launches + 1;  // + set_frame_kernel
    return FXB_OK;
}

}  // namespace

void* fxb::field_device_ptr(fxb_sim* s, int field, size_t* elem_bytes, int* err) {
    *err = FXB_OK;
    switch (field) {
        case FXB_FIELD_VELOCITY: *elem_bytes = 8; return s->vel[0];
        case FXB_FIELD_COLOR: *elem_bytes = 8; return s->col[s->parity];
        case FXB_FIELD_VELOCITY_ADVECTED: *elem_bytes = 8; return s->vel[1];
        case FXB_FIELD_COLOR_PREV: *elem_bytes = 8; return s->col[!s->parity];
        case FXB_FIELD_PRESSURE: {
            *elem_bytes = 4;
            int p_cur = 0;
            if (cudaMemcpy(&p_cur, &s->d_state->p_cur, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) {
                *err = FXB_ERR_CUDA;
                return nullptr;
            }
            return s->p[p_cur & 1];
        }
    }
    *err = FXB_ERR_INVALID;
    return nullptr;
}

using fxb::field_device_ptr;

extern "C" {

int fxb_abi_version(void) { return FXB_ABI_VERSION; }

const char* fxb_last_error(void) { return g_last_error.c_str(); }

int fxb_config_default(fxb_config* cfg) {
    if (!cfg) return fail(FXB_ERR_INVALID, "cfg is null");
    std::memset(cfg, 0, sizeof(*cfg));
    cfg->struct_size = sizeof(fxb_config);
    cfg->nx = cfg->ny = cfg->nz = 128;  // FluidX12.cpp:44
    cfg->address_mode = FXB_ADDRESS_MIRROR;
    cfg->early_exit = 1;
    cfg->jacobi_iters = 64;
    cfg->fuse_t = 0;
    cfg->device = 0;
    cfg->rank = 0;
    cfg->nranks = 1;
    cfg->h_adv = 0;
    cfg->use_graph = 1;
    cfg->kernel_path = 0;
    cfg->phase_timing = 0;
    cfg->halo_backend = FXB_HALO_FUSED;
    cfg->jacobi_group = 0;
    cfg->nccl_unique_id = nullptr;
    return FXB_OK;
}

int fxb_dt_for_grid(uint32_t nx, uint32_t ny, uint32_t nz, float* dt) {
    (void)nx;
    if (!dt || ny == 0) return fail(FXB_ERR_INVALID, "fxb_dt_for_grid: bad argument");
    *dt = (nz > 1 ? 2.0f : 1.0f) / (float)ny;
    return FXB_OK;
}

int fxb_create(const fxb_config* cfg, fxb_sim** out) {
    if (!cfg || !out) return fail(FXB_ERR_INVALID, "fxb_create: null argument");
    *out = nullptr;
    if (cfg->struct_size != sizeof(fxb_config)) return fail(FXB_ERR_INVALID, "fxb_create: struct_size mismatch");
    if (cfg->nx == 0 || cfg->ny == 0 || cfg->nz == 0) return fail(FXB_ERR_INVALID, "fxb_create: empty grid");
    if (cfg->nx != cfg->ny) return fail(FXB_ERR_INVALID, "fxb_create: nx must equal ny (Fluid.cpp:201)");
    if (cfg->nx > 4096 || cfg->nz > 4096) return fail(FXB_ERR_INVALID, "fxb_create: grid dimension > 4096");
    if ((uint64_t)cfg->nx * cfg->ny * cfg->nz >= (1ull << 31))
        return fail(FXB_ERR_INVALID, "fxb_create: more than 2^31 voxels per device (kernels use 32-bit offsets)");
    if (cfg->jacobi_iters < 0 || cfg->jacobi_iters > 128)
        return fail(FXB_ERR_INVALID, "fxb_create: jacobi_iters must be in [0, 128]");
    if (cfg->address_mode != FXB_ADDRESS_MIRROR && cfg->address_mode != FXB_ADDRESS_CLAMP)
        return fail(FXB_ERR_INVALID, "fxb_create: bad address_mode");
    if (cfg->nranks < 1 || cfg->rank < 0 || cfg->rank >= cfg->nranks)
        return fail(FXB_ERR_INVALID, "fxb_create: bad rank/nranks");
    if (cfg->nranks > 1 && !cfg->nccl_unique_id)
        return fail(FXB_ERR_INVALID, "fxb_create: nranks > 1 needs nccl_unique_id (fxb_nccl_unique_id on rank 0)");
    if (cfg->nranks > 1 && (cfg->kernel_path != 0 || cfg->nz <= 1 || cfg->nx % 8 != 0 || cfg->jacobi_iters < 1))
        return fail(FXB_ERR_INVALID, "fxb_create: the z-slab multi-GPU step needs the tuned 3D path (nx % 8 == 0)");

    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(FXB_ERR_CUDA, "fxb_create: no CUDA device (there is no CPU fallback)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(FXB_ERR_INVALID, "fxb_create: bad device ordinal");
    FXB_CUDA(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    FXB_CUDA(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(FXB_ERR_CUDA, "fxb_create: device is not sm_100 (kernels are built for sm_100a only)");

    fxb_sim* s = new fxb_sim;
    s->cfg = *cfg;
    s->cfg.nccl_unique_id = nullptr;
    s->dom.nx = (int)cfg->nx; s->dom.ny = (int)cfg->ny; s->dom.nz = (int)cfg->nz;
    s->dom.z_first = 0; s->dom.nz_alloc = (int)cfg->nz;
    s->dom.z_own0 = 0; s->dom.z_own1 = (int)cfg->nz;
    // Pitched pressure-side arrays: on the tuned 3D path their rows are padded to a multiple of 8 floats, so every grid
    // width takes the TMA-staged fused Jacobi kernels (multi-GPU slabs require nx % 8 == 0 anyway: pitch == nx there).
    s->dom.pitch = (int)cfg->nx;
    if (cfg->kernel_path == 0 && cfg->nz > 1 && cfg->nx >= 8 && cfg->jacobi_iters > 0 && cfg->nranks == 1)
        s->dom.pitch = ((int)cfg->nx + 7) / 8 * 8;
    s->fuse_t = 1;
    if (cfg->nranks > 1) {
        // z-slab decomposition (fluidx12_b200/slab.py states the same rules): rank r owns planes
        // [r*nz/R, (r+1)*nz/R); interior faces carry `halo` extra planes, the grid's own faces none
        const int nz = (int)cfg->nz, R = cfg->nranks, r = cfg->rank;
        const int fuse = cfg->fuse_t ? cfg->fuse_t : 2;
        s->h_adv = cfg->h_adv > 0 ? cfg->h_adv : 8;  // 2|u_z| voxels; |u_z| stayed below 3 in every run
        s->jacobi_group = cfg->jacobi_group > 0 ? cfg->jacobi_group : 1;
        if (cfg->halo_backend == FXB_HALO_FUSED) s->jacobi_group = 1;  // fused halos: every pass pushes its own face planes
        s->halo = std::max(s->h_adv + 1, fuse);  // the Jacobi group uses what the advection halo provides
        int thinnest = nz;
        for (int q = 0; q < R; ++q) thinnest = std::min(thinnest, (q + 1) * nz / R - q * nz / R);
        if (R > nz || s->halo > thinnest) {
            delete s;
            return fail(FXB_ERR_INVALID, "fxb_create: halo deeper than the thinnest slab (fewer ranks or smaller h_adv)");
        }
        s->dom.z_own0 = r * nz / R;
        s->dom.z_own1 = (r + 1) * nz / R;
        s->dom.z_first = std::max(s->dom.z_own0 - s->halo, 0);
        s->dom.nz_alloc = std::min(s->dom.z_own1 + s->halo, nz) - s->dom.z_first;
    }

    auto cleanup_fail = [&](int rc) { fxb_destroy(s); return rc; };
    const size_t n = s->alloc_voxels();
    const size_t np = (size_t)s->dom.pitch * s->dom.ny * s->dom.nz_alloc;  // elements of a pressure-side array
    // Peer-memory halos (FXB_P2P=1) map these buffers into the neighbours through CUDA IPC, and an IPC handle maps a
    // whole underlying allocation: small buffers are then given at least 2 MiB so that each is an allocation of its own.
    if (cfg->halo_backend != FXB_HALO_FUSED && cfg->halo_backend != FXB_HALO_PEER && cfg->halo_backend != FXB_HALO_NCCL) {
        delete s;
        return fail(FXB_ERR_INVALID, "fxb_create: bad halo_backend");
    }
    const bool peer_halos = cfg->nranks > 1 && cfg->halo_backend != FXB_HALO_NCCL;
    const size_t ipc_min = peer_halos ? ((size_t)2 << 20) : 0;
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
        e = cudaMalloc(&s->vel[i], std::max(n * 8, ipc_min));
        if (e == cudaSuccess) e = cudaMemset(s->vel[i], 0, n * 8);  // zero-filled like new D3D12 resources
        if (e == cudaSuccess) e = cudaMalloc(&s->col[i], std::max(n * 8, ipc_min));
        if (e == cudaSuccess) e = cudaMemset(s->col[i], 0, n * 8);
        if (e == cudaSuccess) e = cudaMalloc((void**)&s->p[i], std::max(np * 4, ipc_min));
        if (e == cudaSuccess) e = cudaMemset(s->p[i], 0, np * 4);
    }
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->rhs, std::max(np * 4, ipc_min));
    if (e == cudaSuccess) e = cudaMemset(s->rhs, 0, np * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->active, n);
    if (e == cudaSuccess) e = cudaMemset(s->active, 0, n);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_frame, sizeof(fxb::FrameParams));
    if (e == cudaSuccess) e = cudaMemset(s->d_frame, 0, sizeof(fxb::FrameParams));
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->d_state, sizeof(fxb::StepState));
    if (e == cudaSuccess) e = cudaMemset(s->d_state, 0, sizeof(fxb::StepState));
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&s->own_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess && cfg->nranks > 1) e = cudaStreamCreateWithFlags(&s->side_stream, cudaStreamNonBlocking);
    for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaEventCreate(&s->ev[i]);
    if (e != cudaSuccess)
        return cleanup_fail(fail(FXB_ERR_CUDA, std::string("fxb_create: allocation failed: ") + cudaGetErrorString(e)));

    if (s->multi() && !s->comm.init(cfg->nccl_unique_id, cfg->rank, cfg->nranks))
        return cleanup_fail(fail(FXB_ERR_NCCL, "fxb_create: " + fxb::halo_last_error()));
    int rc = build_emitter(s);
    if (rc != FXB_OK) return cleanup_fail(rc);
    rc = build_axis_tables(s);
    if (rc != FXB_OK) return cleanup_fail(rc);
    s->quad = s->cfg.kernel_path == 0 && fxb::quad_kernels_supported(s->dom);
    if (cfg->fuse_t < 0 || cfg->fuse_t > 4) return cleanup_fail(fail(FXB_ERR_INVALID, "fxb_create: fuse_t must be 0..4"));
    if (s->cfg.kernel_path == 0 && fxb::fused_jacobi_supported(s->dom) && s->cfg.jacobi_iters > 0) {
        // Tuned path: T sweeps fused per HBM pass.  Grids whose nx is not a multiple of 8 (e.g. the 150^3 of
        // Bin/FluidGI.bat) and the 2D path use the one-sweep-per-launch kernels instead.
        const size_t mask_bytes = np / 8;
        for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
            e = cudaMalloc((void**)&s->jac.mask[i], std::max(mask_bytes, ipc_min));
            if (e == cudaSuccess) e = cudaMemset(s->jac.mask[i], 0, mask_bytes);
        }
        if (e == cudaSuccess && fxb::fused_jacobi_plan(&s->jac, s->dom, cfg->fuse_t, s->p[0], s->p[1], s->rhs,
                                                            !s->multi() || cfg->halo_backend == FXB_HALO_FUSED) != 0)
            return cleanup_fail(fail(FXB_ERR_CUDA, "fxb_create: cuTensorMapEncodeTiled failed"));
        s->fuse_t = s->jac.T;
        const size_t nc = 3 * (fxb::FusedJacobi::kMaxPasses + 1);
        for (int i = 0; i < 2 && e == cudaSuccess; ++i) {
            e = cudaMalloc((void**)&s->jac.work_list[i], 2 * (size_t)s->jac.list_stride * sizeof(int));
            if (e == cudaSuccess) e = cudaMemset(s->jac.work_list[i], 0, 2 * (size_t)s->jac.list_stride * sizeof(int));
        }
        const size_t nbricks = fxb::fused_jacobi_bricks(s->jac);
        // (two words per brick: the first pass's flags, and where the halves of a brick meet in the tail passes)
        if (e == cudaSuccess) e = cudaMalloc((void**)&s->jac.brick_flag, 2 * nbricks * sizeof(int));
        if (e == cudaSuccess) e = cudaMemset(s->jac.brick_flag, 0, 2 * nbricks * sizeof(int));
        s->jac.copy_all = s->multi() && s->jacobi_group > 1;
        if (e == cudaSuccess) e = cudaMalloc((void**)&s->jac.work_count, nc * sizeof(int));
        if (e == cudaSuccess) e = cudaMemset(s->jac.work_count, 0, nc * sizeof(int));
        if (e != cudaSuccess)
            return cleanup_fail(fail(FXB_ERR_CUDA, std::string("fxb_create: ") + cudaGetErrorString(e)));
        s->fused = true;
    }
    if (peer_halos) {
        // Halos through peer memory (CUDA IPC + one store/flag kernel per exchange) instead of NCCL send/recv; the
        // all-reduce of the freeze counters stays on NCCL.
        std::vector<void*> bufs = {s->vel[0], s->vel[1], s->col[0], s->col[1], (void*)s->p[0], (void*)s->p[1], (void*)s->rhs};
        if (s->fused) { bufs.push_back(s->jac.mask[0]); bufs.push_back(s->jac.mask[1]); }
        const int nz = (int)cfg->nz, R = cfg->nranks, r = cfg->rank;
        const int zf_lo = r > 0 ? std::max((r - 1) * nz / R - s->halo, 0) : 0;
        const int zf_hi = r < R - 1 ? std::max((r + 1) * nz / R - s->halo, 0) : 0;
        if (!s->comm.p2p_init(bufs.data(), (int)bufs.size(), zf_lo, zf_hi, s->own_stream))
            return cleanup_fail(fail(FXB_ERR_NCCL, "fxb_create: peer-memory halo setup failed: " + fxb::halo_last_error()));
        if (cfg->halo_backend == FXB_HALO_FUSED) {
            // fused halos: every kernel also stores its face planes into the neighbours' arrays (common.cuh PeerView)
            s->pv = s->comm.peer_view(s->dom);
            for (int side = 0; side < 2; ++side) {
                s->advect_peers.vel_out[side] = s->comm.peer_of(s->vel[1], side);
                for (int i = 0; i < 2; ++i) {
                    s->advect_peers.col[side][i] = s->comm.peer_of(s->col[i], side);
                    s->jac.peer_p[side][i] = (float*)s->comm.peer_of(s->p[i], side);
                    s->jac.peer_m[side][i] = (unsigned char*)s->comm.peer_of(s->jac.mask[i], side);
                }
            }
        }
    }
    const bool one_graph = !s->multi() || cfg->halo_backend == FXB_HALO_FUSED;  // no host-side buffer selection in the step
    if (s->cfg.use_graph && one_graph) {
        rc = capture_graph(s, 0, 0);
        if (rc != FXB_OK) return cleanup_fail(rc);
    } else {
        const int jl = s->fused ? fxb::fused_jacobi_passes(s->jac, s->cfg.jacobi_iters) - 1 + fxb::fused_jacobi_launches(s->jac, s->pv, 0)
                                : s->cfg.jacobi_iters;
        s->kernels_per_step = 1 + 1 + 2 + jl + (s->fused ? 2 : 1) + 1 + (s->multi() && s->fused ? 1 : 0) +
                              ((s->cfg.phase_timing || (s->multi() && cfg->halo_backend == FXB_HALO_FUSED)) ? 2 : 0);
    }
    if (s->multi()) {
        // establish the NCCL connections now (outside any graph capture): one throw-away exchange and reduction
        const fxb::HaloField f[1] = {{s->rhs, s->plane_voxels() * 4, 1}};
        if (!s->comm.exchange(s->dom, f, 1, s->own_stream) ||
            !s->comm.all_reduce_sum_u64(s->d_state->active_after, 128, s->own_stream))
            return cleanup_fail(fail(FXB_ERR_NCCL, "fxb_create: " + fxb::halo_last_error()));
        if (cudaStreamSynchronize(s->own_stream) == cudaSuccess) cudaMemset(s->d_state, 0, sizeof(fxb::StepState));
    }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess)
        return cleanup_fail(fail(FXB_ERR_CUDA, std::string("fxb_create: ") + cudaGetErrorString(e)));
    *out = s;
    return FXB_OK;
}

void fxb_destroy(fxb_sim* s) {
    if (!s) return;
    cudaSetDevice(s->cfg.device);
    cudaDeviceSynchronize();
    for (int a = 0; a < 2; ++a)
        for (int b = 0; b < 2; ++b) {
            if (s->graph_exec[a][b]) cudaGraphExecDestroy(s->graph_exec[a][b]);
            if (s->graph[a][b]) cudaGraphDestroy(s->graph[a][b]);
        }
    s->comm.destroy();
    for (int i = 0; i < 2; ++i) {
        cudaFree(s->vel[i]);
        cudaFree(s->col[i]);
        cudaFree(s->p[i]);
    }
    cudaFree(s->rhs);
    cudaFree(s->active);
    cudaFree(s->jac.mask[0]);
    cudaFree(s->jac.mask[1]);
    cudaFree(s->jac.work_list[0]);
    cudaFree(s->jac.work_list[1]);
    cudaFree(s->jac.work_count);
    cudaFree(s->jac.brick_flag);
    cudaFree(s->light_map);
    cudaFree(s->cube_map);
    if (s->stats_ring) cudaFreeHost(s->stats_ring);
    for (int i = 0; i < fxb_sim::kStatsSlots; ++i)
        if (s->stats_event[i]) cudaEventDestroy(s->stats_event[i]);
    cudaFree(s->whole_colour);
    cudaFree(s->whole_light_map);
    cudaFree(s->light_density);
    cudaFree(s->emitter_basis);
    cudaFree(s->axis_tables);
    cudaFree(s->d_frame);
    cudaFree(s->d_state);
    for (int i = 0; i < 8; ++i)
        if (s->ev[i]) cudaEventDestroy(s->ev[i]);
    if (s->own_stream) cudaStreamDestroy(s->own_stream);
    if (s->side_stream) cudaStreamDestroy(s->side_stream);
    delete s;
}

int fxb_update_frame(fxb_sim* s, float dt) {
    if (!s) return fail(FXB_ERR_INVALID, "fxb_update_frame: null handle");
    s->dt = dt;                      // Fluid.cpp:344
    if (dt > 0.0f) s->parity ^= 1;   // Fluid.cpp:345
    return FXB_OK;
}

int fxb_simulate(fxb_sim* s, void* cuda_stream) {
    if (!s) return fail(FXB_ERR_INVALID, "fxb_simulate: null handle");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    FXB_CUDA(cudaSetDevice(s->cfg.device));
    set_frame_kernel<<<1, 1, 0, st>>>(s->d_frame, s->d_state, s->dt, s->parity, s->steps);  // the CBSimulation upload (Fluid.cpp:288-290)
    const int npass = flips_per_step(s);
    const bool fused_halos = s->multi() && s->cfg.halo_backend == FXB_HALO_FUSED;
    if (fused_halos && s->halo_stale) {
        // fxb_set_field changed own planes behind the neighbours' back: one plain exchange of everything a step reads
        // from its halos before the step's own stores refresh them
        const fxb::HaloField f[5] = {{s->vel[0], s->plane_voxels() * 8, s->halo}, {s->col[0], s->plane_voxels() * 8, s->halo},
                                     {s->col[1], s->plane_voxels() * 8, s->halo}, {s->p[0], s->plane_voxels() * 4, s->halo},
                                     {s->p[1], s->plane_voxels() * 4, s->halo}};
        if (!s->comm.exchange(s->dom, f, 3, st) || !s->comm.exchange(s->dom, f + 3, 2, st))
            return fail(FXB_ERR_NCCL, "fxb_simulate: halo refresh: " + fxb::halo_last_error());
        s->halo_stale = false;
    }
    if (s->multi() && !fused_halos) {
        // the exchanges depend on dt > 0, the frame parity and the pressure parity: one graph per key, captured on
        // first use; a paused frame (dt <= 0) is enqueued directly
        const int a = s->parity, b = s->p_cur_host;
        if (s->cfg.use_graph && s->dt > 0.0f) {
            if (!s->graph_exec[a][b]) {
                const int rc = capture_graph(s, a, b);
                if (rc != FXB_OK) return rc;
            }
            FXB_CUDA(cudaGraphLaunch(s->graph_exec[a][b], st));
        } else {
            const int rc = enqueue_step(s, st, nullptr);
            if (rc != FXB_OK) return rc;
        }
        if (s->dt > 0.0f) s->p_cur_host = (s->p_cur_host + npass) & 1;
    } else if (s->graph_exec[0][0]) {
        FXB_CUDA(cudaGraphLaunch(s->graph_exec[0][0], st));
    } else {
        const int rc = enqueue_step(s, st, nullptr);
        if (rc != FXB_OK) return rc;
    }
    s->last_stream = st;
    ++s->steps;
    return FXB_OK;
}

int fxb_sync(fxb_sim* s) {
    if (!s) return fail(FXB_ERR_INVALID, "fxb_sync: null handle");
    FXB_CUDA(cudaSetDevice(s->cfg.device));
    FXB_CUDA(cudaDeviceSynchronize());
    if (s->multi() && s->comm.p2p_timed_out())
        return fail(FXB_ERR_NCCL, "fxb_sync: a peer-memory halo exchange timed out waiting for a neighbour (results are invalid)");
    return FXB_OK;
}

int fxb_get_slab(const fxb_sim* s, uint32_t* z0, uint32_t* count) {
    if (!s || !z0 || !count) return fail(FXB_ERR_INVALID, "fxb_get_slab: null argument");
    *z0 = (uint32_t)s->dom.z_own0;
    *count = (uint32_t)(s->dom.z_own1 - s->dom.z_own0);
    return FXB_OK;
}

int fxb_get_field(fxb_sim* s, int field, void* host, size_t bytes) {
    if (!s || !host) return fail(FXB_ERR_INVALID, "fxb_get_field: null argument");
    FXB_CUDA(cudaSetDevice(s->cfg.device));
    FXB_CUDA(cudaDeviceSynchronize());
    size_t eb; int err;
    const char* src = (const char*)field_device_ptr(s, field, &eb, &err);
    if (!src) return fail(err, "fxb_get_field: bad field");
    if (bytes != s->own_voxels() * eb) return fail(FXB_ERR_SIZE, "fxb_get_field: size mismatch");
    if (field == FXB_FIELD_PRESSURE && s->dom.pitch != s->dom.nx) {  // pitched rows -> the dense host layout
        const size_t rows = (size_t)s->dom.ny * (s->dom.z_own1 - s->dom.z_own0);
        const size_t first = (size_t)s->dom.ny * (s->dom.z_own0 - s->dom.z_first) * s->dom.pitch;
        FXB_CUDA(cudaMemcpy2D(host, (size_t)s->dom.nx * 4, src + first * 4, (size_t)s->dom.pitch * 4, (size_t)s->dom.nx * 4, rows,
                              cudaMemcpyDeviceToHost));
        return FXB_OK;
    }
    FXB_CUDA(cudaMemcpy(host, src + s->own_offset() * eb, bytes, cudaMemcpyDeviceToHost));
    return FXB_OK;
}

int fxb_get_field_async(fxb_sim* s, int field, void* host, size_t bytes, void* cuda_stream) {
    if (!s || !host) return fail(FXB_ERR_INVALID, "fxb_get_field_async: null argument");
    if (field == FXB_FIELD_PRESSURE)
        return fail(FXB_ERR_INVALID, "fxb_get_field_async: pressure needs the synchronous call");
    size_t eb; int err;
    const char* src = (const char*)field_device_ptr(s, field, &eb, &err);
    if (!src) return fail(err, "fxb_get_field_async: bad field");
    if (bytes != s->own_voxels() * eb) return fail(FXB_ERR_SIZE, "fxb_get_field_async: size mismatch");
    FXB_CUDA(cudaMemcpyAsync(host, src + s->own_offset() * eb, bytes, cudaMemcpyDeviceToHost,
                             (cudaStream_t)cuda_stream));
    return FXB_OK;
}

int fxb_set_field(fxb_sim* s, int field, const void* host, size_t bytes) {
    if (!s || !host) return fail(FXB_ERR_INVALID, "fxb_set_field: null argument");
    FXB_CUDA(cudaSetDevice(s->cfg.device));
    FXB_CUDA(cudaDeviceSynchronize());
    size_t eb; int err;
    char* dst = (char*)field_device_ptr(s, field, &eb, &err);
    if (!dst) return fail(err, "fxb_set_field: bad field");
    if (bytes != s->own_voxels() * eb) return fail(FXB_ERR_SIZE, "fxb_set_field: size mismatch");
    if (field == FXB_FIELD_PRESSURE && s->dom.pitch != s->dom.nx) {  // the dense host layout -> pitched rows
        const size_t rows = (size_t)s->dom.ny * (s->dom.z_own1 - s->dom.z_own0);
        const size_t first = (size_t)s->dom.ny * (s->dom.z_own0 - s->dom.z_first) * s->dom.pitch;
        FXB_CUDA(cudaMemcpy2D(dst + first * 4, (size_t)s->dom.pitch * 4, host, (size_t)s->dom.nx * 4, (size_t)s->dom.nx * 4, rows,
                              cudaMemcpyHostToDevice));
        s->halo_stale = true;
        return FXB_OK;
    }
    FXB_CUDA(cudaMemcpy(dst + s->own_offset() * eb, host, bytes, cudaMemcpyHostToDevice));
    s->halo_stale = true;
    return FXB_OK;
}

}  // extern "C"

namespace {
// The public record of a step from a snapshot of the device-side counters.
void fill_stats(const fxb_sim* s, const fxb::StepState& st, int parity, uint64_t steps, fxb_stats* out) {
    std::memset(out, 0, sizeof(*out));
    out->s_exec = st.s_exec;
    out->jacobi_passes = st.passes;
    out->fuse_t = s->fuse_t;
    out->halo_overflow = st.halo_overflow;
    out->frame_parity = parity;
    out->kernels_per_step = s->kernels_per_step;
    out->steps = steps;
    out->active_after_first_sweep = st.active_after[0];
    out->total_sweeps = st.total_sweeps;
    out->total_passes = st.total_passes;
    out->bricks_processed = st.bricks_processed;
    out->bricks_copied = st.bricks_copied;
    out->jacobi_fused = s->fused ? 1 : 0;
    out->tail_from = s->fused && s->jac.tail_from <= fxb::FusedJacobi::kMaxPasses ? s->jac.tail_from : 0;
    if (s->fused) {
        out->brick_cells = fxb::fused_jacobi_brick_cells(s->jac);
        out->bricks_per_pass = fxb::fused_jacobi_bricks(s->jac);
    }
}
}  // namespace

extern "C" {

int fxb_get_stats(fxb_sim* s, fxb_stats* out) {
    if (!s || !out) return fail(FXB_ERR_INVALID, "fxb_get_stats: null argument");
    FXB_CUDA(cudaSetDevice(s->cfg.device));
    fxb::StepState st;
    cudaStream_t stream = s->last_stream;
    FXB_CUDA(cudaMemcpyAsync(&st, s->d_state, sizeof(st), cudaMemcpyDeviceToHost, stream));
    FXB_CUDA(cudaStreamSynchronize(stream));
    fill_stats(s, st, s->parity, s->steps, out);
    return st.halo_overflow ? fail(FXB_ERR_HALO_OVERFLOW, "advection back-trace left the z-halo") : FXB_OK;
}

int fxb_post_stats(fxb_sim* s, int slot) {
    if (!s || slot < 0 || slot >= fxb_sim::kStatsSlots) return fail(FXB_ERR_INVALID, "fxb_post_stats: bad argument");
    FXB_CUDA(cudaSetDevice(s->cfg.device));
    if (!s->stats_ring) {
        FXB_CUDA(cudaHostAlloc((void**)&s->stats_ring, fxb_sim::kStatsSlots * sizeof(fxb::StepState), cudaHostAllocDefault));
        for (int i = 0; i < fxb_sim::kStatsSlots; ++i)
            FXB_CUDA(cudaEventCreateWithFlags(&s->stats_event[i], cudaEventDisableTiming));
    }
    FXB_CUDA(cudaMemcpyAsync(&s->stats_ring[slot], s->d_state, sizeof(fxb::StepState), cudaMemcpyDeviceToHost, s->last_stream));
    FXB_CUDA(cudaEventRecord(s->stats_event[slot], s->last_stream));
    s->stats_steps[slot] = s->steps;
    s->stats_parity[slot] = s->parity;
    return FXB_OK;
}

int fxb_wait_stats(fxb_sim* s, int slot, fxb_stats* out) {
    if (!s || !out || slot < 0 || slot >= fxb_sim::kStatsSlots) return fail(FXB_ERR_INVALID, "fxb_wait_stats: bad argument");
    if (!s->stats_ring || s->stats_steps[slot] == 0) return fail(FXB_ERR_INVALID, "fxb_wait_stats: nothing was posted to this slot");
    FXB_CUDA(cudaSetDevice(s->cfg.device));
    FXB_CUDA(cudaEventSynchronize(s->stats_event[slot]));
    const fxb::StepState& st = s->stats_ring[slot];
    fill_stats(s, st, s->stats_parity[slot], s->stats_steps[slot], out);
    return st.halo_overflow ? fail(FXB_ERR_HALO_OVERFLOW, "advection back-trace left the z-halo") : FXB_OK;
}

int fxb_p2p_plan(int32_t nz, int32_t nranks, int32_t rank, int32_t halo, int32_t depth, int64_t* out4) {
    if (!out4 || nz < 1 || nranks < 1 || rank < 0 || rank >= nranks || nranks > nz || halo < 0 || depth < 1)
        return fail(FXB_ERR_INVALID, "fxb_p2p_plan: bad argument");
    auto z_first = [&](int r) { return std::max(r * nz / nranks - halo, 0); };
    fxb::Domain d{};
    d.nz = nz;
    d.z_own0 = rank * nz / nranks;
    d.z_own1 = (rank + 1) * nz / nranks;
    d.z_first = z_first(rank);
    d.nz_alloc = std::min(d.z_own1 + halo, nz) - d.z_first;
    const fxb::P2PPlanes q = fxb::p2p_planes(d, depth, rank > 0 ? z_first(rank - 1) : 0, rank < nranks - 1 ? z_first(rank + 1) : 0);
    out4[0] = q.send_lo; out4[1] = q.dst_lo; out4[2] = q.send_hi; out4[3] = q.dst_hi;
    return FXB_OK;
}

int fxb_jacobi_schedule(int32_t iters, int32_t fuse_t, int32_t tail_from, int32_t* npass, int32_t* s0, int32_t n) {
    if (!npass || iters < 0 || fuse_t < 1 || fuse_t > 4 || tail_from < 0 || n < 0 || (n > 0 && !s0))
        return fail(FXB_ERR_INVALID, "fxb_jacobi_schedule: bad argument");
    fxb::FusedJacobi J;
    J.T = fuse_t;
    J.tail_from = tail_from > 0 ? tail_from : fxb::FusedJacobi::kMaxPasses + 1;
    *npass = fxb::fused_jacobi_passes(J, iters);
    for (int k = 0; k < n && k < *npass; ++k) s0[k] = fxb::fused_jacobi_s0(J, k);
    return FXB_OK;
}

int fxb_face_last_order(int32_t n, int32_t chunk, int32_t reach, int32_t has_lo, int32_t has_hi, int32_t* out, int32_t nchunks) {
    if (!out || n < 1 || chunk < 1 || reach < 0 || nchunks != (n + chunk - 1) / chunk)
        return fail(FXB_ERR_INVALID, "fxb_face_last_order: bad argument");
    for (int b = 0; b < nchunks; ++b) out[b] = fxb::face_last_chunk_of(has_lo, has_hi, b, nchunks, chunk, reach, n);
    return FXB_OK;
}

int fxb_emitter_box(uint32_t nx, uint32_t ny, uint32_t nz, int32_t* out6) {
    if (!out6 || nx == 0 || ny == 0 || nz == 0) return fail(FXB_ERR_INVALID, "fxb_emitter_box: bad argument");
    int lo[3], hi[3];
    emitter_box((int)nx, (int)ny, (int)nz, lo, hi);
    for (int a = 0; a < 3; ++a) { out6[a] = lo[a]; out6[3 + a] = hi[a]; }
    return FXB_OK;
}

int fxb_get_freeze_histogram(fxb_sim* s, uint64_t* out, int n) {
    if (!s || !out || n < 0 || n > 128) return fail(FXB_ERR_INVALID, "fxb_get_freeze_histogram: bad argument");
    FXB_CUDA(cudaSetDevice(s->cfg.device));
    FXB_CUDA(cudaDeviceSynchronize());
    FXB_CUDA(cudaMemcpy(out, s->d_state->active_after, (size_t)n * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    return FXB_OK;
}

int fxb_profile_step(fxb_sim* s, float* ms, int n) {
    if (!s || !ms || n < 6) return fail(FXB_ERR_INVALID, "fxb_profile_step: need ms[6]");
    FXB_CUDA(cudaSetDevice(s->cfg.device));
    cudaStream_t st = s->own_stream;
    FXB_CUDA(cudaDeviceSynchronize());
    set_frame_kernel<<<1, 1, 0, st>>>(s->d_frame, s->d_state, s->dt, s->parity, s->steps);
    FXB_CUDA(cudaEventRecord(s->ev[0], st));
    Enqueue q{s, st};
    for (int ph = 0; ph < PH_COUNT; ++ph) {
        enqueue_phase(q, ph);
        FXB_CUDA(cudaEventRecord(s->ev[ph + 1], st));
    }
    FXB_CUDA(cudaStreamSynchronize(st));
    if (!q.ok) return fail(FXB_ERR_CUDA, "fxb_profile_step: " + q.err);
    FXB_CUDA(cudaGetLastError());
    for (int ph = 0; ph < PH_COUNT; ++ph) FXB_CUDA(cudaEventElapsedTime(&ms[ph], s->ev[ph], s->ev[ph + 1]));
    ms[4] = 0.0f;
    FXB_CUDA(cudaEventElapsedTime(&ms[5], s->ev[0], s->ev[PH_COUNT]));
    if (s->multi() && s->fused && s->dt > 0.0f) s->p_cur_host = (s->p_cur_host + flips_per_step(s)) & 1;
    s->last_stream = st;
    ++s->steps;
    return FXB_OK;
}

#ifdef FXB_TIMING
int fxb_debug_stamps(fxb_sim* s, long long* out, int n) {  // debug build only (not in the header)
    if (!s || !out || n > 128) return FXB_ERR_INVALID;
    cudaDeviceSynchronize();
    return cudaMemcpy(out, s->d_state->dbg, (size_t)n * sizeof(long long), cudaMemcpyDeviceToHost) == cudaSuccess ? FXB_OK : FXB_ERR_CUDA;
}
#endif

int fxb_state_checksum(fxb_sim* s, uint64_t* out3) {
    if (!s || !out3) return fail(FXB_ERR_INVALID, "fxb_state_checksum: null argument");
    FXB_CUDA(cudaSetDevice(s->cfg.device));
    FXB_CUDA(cudaDeviceSynchronize());
    int p_cur = 0;
    FXB_CUDA(cudaMemcpy(&p_cur, &s->d_state->p_cur, sizeof(int), cudaMemcpyDeviceToHost));
    unsigned long long* d_out = nullptr;
    FXB_CUDA(cudaMalloc((void**)&d_out, 3 * sizeof(unsigned long long)));
    cudaError_t e = cudaMemset(d_out, 0, 3 * sizeof(unsigned long long));
    if (e == cudaSuccess) {
        checksum_kernel<<<1184, 256>>>(s->dom, (const uint2*)s->vel[0], (const uint2*)s->col[s->parity], s->p[p_cur & 1], d_out);
        e = cudaGetLastError();
    }
    unsigned long long h[3] = {0, 0, 0};
    if (e == cudaSuccess) e = cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    cudaFree(d_out);
    if (e != cudaSuccess) return fail(FXB_ERR_CUDA, std::string("fxb_state_checksum: ") + cudaGetErrorString(e));
    for (int i = 0; i < 3; ++i) out3[i] = h[i];
    return FXB_OK;
}

int fxb_get_phase_times(fxb_sim* s, double* ms, int n, int reset) {
    if (!s || !ms || n < 4) return fail(FXB_ERR_INVALID, "fxb_get_phase_times: need ms[4]");
    if (!s->cfg.phase_timing) return fail(FXB_ERR_INVALID, "fxb_get_phase_times: the handle was created without phase_timing");
    FXB_CUDA(cudaSetDevice(s->cfg.device));
    FXB_CUDA(cudaDeviceSynchronize());
    unsigned long long ns[8];
    FXB_CUDA(cudaMemcpy(ns, s->d_state->phase_ns, sizeof(ns), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 4; ++i) ms[i] = (double)ns[i] * 1e-6;
    if (reset) FXB_CUDA(cudaMemset(s->d_state->phase_ns, 0, sizeof(ns)));
    return FXB_OK;
}

int fxb_nccl_unique_id(void* out128) {
    if (!out128) return fail(FXB_ERR_INVALID, "fxb_nccl_unique_id: null argument");
    if (!fxb::halo_unique_id(out128)) return fail(FXB_ERR_NCCL, "fxb_nccl_unique_id: " + fxb::halo_last_error());
    return FXB_OK;
}

}  // extern "C"
