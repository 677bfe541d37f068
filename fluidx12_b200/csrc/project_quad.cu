// project_quad.cu — divergence and gradient-subtract of CSProject3D, one thread per 4 x-adjacent cells.
//
// Replaces GetDivergence (FluidX12/Content/Shaders/CSProject3D.hlsl:39-50) and Project + the soft-wall
// damping (CSProject3D.hlsl:55-63, :106-112) for 3D grids whose width is a multiple of 8; other grids use
// the one-thread-per-voxel kernels of project_simple.cu.  Both kernels are pure streaming stencils:
// 128-bit loads of whole quads, clamp-to-edge neighbours by index (CSProject3D.hlsl:76-83), per-axis
// tables instead of per-voxel divisions, 32-bit offsets.  Operation order as in SURVEY.md App. A.2.
// Algorithmic traffic: divergence 12 B/voxel (velocity in 8 + rhs out 4); gradient-subtract 20 B/voxel
// (pressure in 4 + velocity in 8 + velocity out 8).
#include "common.cuh"
#include "kernels.h"

namespace fxb {

namespace {

struct Quad8 {  // four RGBA16F texels
    uint4 a, b;
};

__device__ __forceinline__ Quad8 load_quad8(const uint2* __restrict__ f, unsigned i) {
    const uint4* p = reinterpret_cast<const uint4*>(f + i);
    Quad8 q;
    q.a = __ldg(p);
    q.b = __ldg(p + 1);
    return q;
}

__device__ __forceinline__ float h_lo(unsigned w) { return __half2float(__ushort_as_half((unsigned short)(w & 0xffffu))); }
__device__ __forceinline__ float h_hi(unsigned w) { return __half2float(__ushort_as_half((unsigned short)(w >> 16))); }

struct Rows {
    unsigned c, u, d, f, b;  // offsets of the quad at (x0, y, z) and of its y / z neighbours (clamped)
    int xl, xr;              // clamped x of the left / right neighbour cell
};

__device__ __forceinline__ Rows rows_of(const Domain& d, int x0, int y, int z) {
    Rows r;
    const unsigned plane = (unsigned)d.nx * d.ny;
    const unsigned zc = (unsigned)(z - d.z_first) * plane;
    r.c = zc + (unsigned)y * d.nx + x0;
    r.u = zc + (unsigned)(max(y, 1) - 1) * d.nx + x0;
    r.d = zc + (unsigned)min(y + 1, d.ny - 1) * d.nx + x0;
    r.f = (unsigned)(max(z, 1) - 1 - d.z_first) * plane + (unsigned)y * d.nx + x0;
    r.b = (unsigned)(min(z + 1, d.nz - 1) - d.z_first) * plane + (unsigned)y * d.nx + x0;
    r.xl = max(x0, 1) - 1;
    r.xr = min(x0 + 4, d.nx - 1);
    return r;
}

// z-marching: a thread owns the quad column (x0..x0+3, y) over kDivPlanes planes and keeps the z component of the
// plane below and the whole texels of the centre plane in registers, so every texel is fetched from L2/HBM once
// (plus the chunk's two end planes); only the y neighbours (rows y-1, y+1 of the centre plane, loaded a moment
// earlier by the neighbouring threads of the CTA) and the two x-edge texels come through L1 again.
constexpr int kDivPlanes = 8;

template <bool FUSED>  // FUSED: multi-GPU with fused halos (common.cuh PeerView)
__global__ void __launch_bounds__(256) divergence_quad_kernel(Domain d, const __grid_constant__ PeerView pv,
                                                              const FrameParams* __restrict__ frame,
                                                              const uint2* __restrict__ vel, float* __restrict__ rhs,
                                                              float* rhs_lo, float* rhs_hi, int push_depth) {
    if (!(0.0f < frame->dt)) return;
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int zc = FUSED ? face_last_chunk(pv, blockIdx.z, gridDim.z, kDivPlanes, push_depth, d.z_own1 - d.z_own0) : (int)blockIdx.z;
    const int z_begin = d.z_own0 + zc * kDivPlanes;
    const int z_end = min(z_begin + kDivPlanes, d.z_own1);
    // fused halos (common.cuh PeerView): the chunks at an interior face read the neighbour's first / last plane of the
    // advected velocity and store their right-hand side into the neighbour's halo as well (event m = 1)
    const bool near_lo = FUSED && pv.has_lo && z_begin < d.z_own0 + push_depth;
    const bool near_hi = FUSED && pv.has_hi && z_end > d.z_own1 - push_depth;
    if (FUSED && (near_lo || near_hi)) {
        if (threadIdx.x == 0 && threadIdx.y == 0) peer_wait(pv, frame->epoch_base + 1, near_lo, near_hi);
        __syncthreads();
    }
    // Threads beyond the grid's x edge stay (they take part in the barriers and shuffles below) on clamped coordinates
    // and store nothing.  (ny is a multiple of 8 on this path: nx == ny, nx % 8 == 0.)
    const bool valid = x0 < d.nx;
    const int xq = valid ? x0 : d.nx - 4;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const unsigned plane = (unsigned)d.nx * d.ny;
    const unsigned row_c = (unsigned)y * d.nx + xq;
    const unsigned row_u = (unsigned)(max(y, 1) - 1) * d.nx + xq;
    const unsigned row_d = (unsigned)min(y + 1, d.ny - 1) * d.nx + xq;
    const unsigned row0 = (unsigned)y * d.nx;
    const int xl = max(xq, 1) - 1, xr = min(xq + 4, d.nx - 1);
    const unsigned short* vs = reinterpret_cast<const unsigned short*>(vel);
    // the y components of the CTA's rows (plus one row above and below) of the plane being processed, double-buffered
    // by plane parity: [buffer][row + 1][quad] = the four y halves of a quad
    __shared__ uint2 s_y[2][10][32];
    auto y_of = [](const Quad8& q) {  // high halves of the texels' first words
        return make_uint2(__byte_perm(q.a.x, q.a.z, 0x7632), __byte_perm(q.b.x, q.b.z, 0x7632));
    };

    auto zoff = [&](int z) { return (unsigned)(z - d.z_first) * plane; };
    // plane below the first one (clamped at the grid face) and the first two planes of the chunk
    const Quad8 below = load_quad8(vel, zoff(max(z_begin, 1) - 1) + row_c);
    float fz[4] = {h_lo(below.a.y), h_lo(below.a.w), h_lo(below.b.y), h_lo(below.b.w)};
    Quad8 c = load_quad8(vel, zoff(z_begin) + row_c);
    Quad8 above = load_quad8(vel, zoff(min(z_begin + 1, d.nz - 1)) + row_c);
    // Software pipeline: the plane two above (streamed from HBM), the rows just outside the CTA (one quad for the
    // threads of the first / last row) and the x-edge texels of the warp's first / last lane are fetched one iteration
    // ahead.  Everything else a plane needs comes from registers (z), shared memory (y) or shuffles (x): each texel is
    // loaded from HBM once, and only the CTA's halo comes through the caches a second time.
    struct Fetch {
        Quad8 above2, edge_row;
        unsigned short el, er;
    };
    auto fetch = [&](int z) {  // for plane z: its edge data, and the plane z + 2
        Fetch f;
        const unsigned zc = zoff(z);
        f.above2 = load_quad8(vel, zoff(min(z + 2, d.nz - 1)) + row_c);
        f.edge_row.a = f.edge_row.b = make_uint4(0u, 0u, 0u, 0u);
        if (ty == 0) f.edge_row = load_quad8(vel, zc + row_u);
        if (ty == 7) f.edge_row = load_quad8(vel, zc + row_d);
        f.el = f.er = 0;
        if (tx == 0) f.el = __ldg(vs + 4 * (size_t)(zc + row0 + xl));
        if (tx == 31 || !valid || xq + 4 >= d.nx) f.er = __ldg(vs + 4 * (size_t)(zc + row0 + xr));
        return f;
    };
    Fetch cur = fetch(z_begin);
    for (int z = z_begin; z < z_end; ++z) {
        const unsigned zc = zoff(z);
        const int buf = z & 1;
        // publish this thread's y components (and, from the first / last row, the row outside the CTA)
        s_y[buf][ty + 1][tx] = y_of(c);
        if (ty == 0) s_y[buf][0][tx] = y_of(cur.edge_row);
        if (ty == 7) s_y[buf][9][tx] = y_of(cur.edge_row);
        Fetch nxt = cur;
        if (z + 1 < z_end) nxt = fetch(z + 1);
        __syncthreads();
        const uint2 u = s_y[buf][ty][tx], dn = s_y[buf][ty + 2][tx];
        // x neighbours of the quad's end cells: the neighbouring lanes' texels, or the fetched edge texel
        const unsigned own_first = c.a.x & 0xffffu, own_last = c.b.z & 0xffffu;
        unsigned left = __shfl_up_sync(0xffffffffu, own_last, 1), right = __shfl_down_sync(0xffffffffu, own_first, 1);
        if (tx == 0) left = cur.el;
        if (tx == 31 || xq + 4 >= d.nx) right = cur.er;
        const float vx[6] = {half_bits_to_float((unsigned short)left), h_lo(c.a.x), h_lo(c.a.z),
                             h_lo(c.b.x), h_lo(c.b.z), half_bits_to_float((unsigned short)right)};
        const float uy[4] = {h_lo(u.x), h_hi(u.x), h_lo(u.y), h_hi(u.y)};
        const float dy[4] = {h_lo(dn.x), h_hi(dn.x), h_lo(dn.y), h_hi(dn.y)};
        const float bz[4] = {h_lo(above.a.y), h_lo(above.a.w), h_lo(above.b.y), h_lo(above.b.w)};
        float out[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float a = -vx[j] + vx[j + 2];
            float s = -uy[j] + dy[j];
            s = s + a;
            const float cz = -fz[j] + bz[j];
            s = cz + s;
            out[j] = -0.5f * s;
        }
        const float4 res = make_float4(out[0], out[1], out[2], out[3]);
        if (valid) {
            *reinterpret_cast<float4*>(rhs + zc + row_c) = res;
            if (FUSED && near_lo && z < d.z_own0 + push_depth)
                *reinterpret_cast<float4*>(rhs_lo + (long long)zc + (long long)pv.dz_lo * plane + row_c) = res;
            if (FUSED && near_hi && z >= d.z_own1 - push_depth)
                *reinterpret_cast<float4*>(rhs_hi + (long long)zc + (long long)pv.dz_hi * plane + row_c) = res;
        }
        // march: the centre plane becomes the plane below, the plane above the centre, the plane two above the plane above
        fz[0] = h_lo(c.a.y); fz[1] = h_lo(c.a.w); fz[2] = h_lo(c.b.y); fz[3] = h_lo(c.b.w);
        c = above;
        above = cur.above2;
        cur = nxt;
    }
}

template <bool FUSED>
__global__ void __launch_bounds__(256) gradient_quad_kernel(Domain d, AxisTables tab,
                                                            const __grid_constant__ PeerView pv,
                                                            const FrameParams* __restrict__ frame,
                                                            const uint2* __restrict__ vel_in, const float* p0,
                                                            const float* p1, uint2* __restrict__ vel_out,
                                                            const StepState* __restrict__ state, uint2* vel_out_lo,
                                                            uint2* vel_out_hi, int reach, int event) {
    const int x0 = (blockIdx.x * 32 + threadIdx.x) * 4;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int z = d.z_own0 + (FUSED ? face_last_chunk(pv, blockIdx.z, gridDim.z, 1, reach, d.z_own1 - d.z_own0) : (int)blockIdx.z);
    // fused halos (common.cuh PeerView): the first / last plane reads the neighbour's pressure plane, and the planes the
    // neighbour's next advection can reach are stored into its halo as well
    const bool near_lo = FUSED && pv.has_lo && z < d.z_own0 + reach, near_hi = FUSED && pv.has_hi && z >= d.z_own1 - reach;
    if (FUSED && (near_lo || near_hi)) {
        if (threadIdx.x == 0 && threadIdx.y == 0) peer_wait(pv, frame->epoch_base + event, near_lo, near_hi);
        __syncthreads();
    }
    if (x0 >= d.nx || y >= d.ny) return;
    const Rows r = rows_of(d, x0, y, z);
    const Quad8 v = load_quad8(vel_in, r.c);
    float ux[4] = {h_lo(v.a.x), h_lo(v.a.z), h_lo(v.b.x), h_lo(v.b.z)};
    float uy[4] = {h_hi(v.a.x), h_hi(v.a.z), h_hi(v.b.x), h_hi(v.b.z)};
    float uz[4] = {h_lo(v.a.y), h_lo(v.a.w), h_lo(v.b.y), h_lo(v.b.w)};
    if (0.0f < frame->dt) {
        const float* __restrict__ p = state->p_cur ? p1 : p0;
        const float4 pc = __ldg(reinterpret_cast<const float4*>(p + r.c));
        const float4 pu = __ldg(reinterpret_cast<const float4*>(p + r.u));
        const float4 pd = __ldg(reinterpret_cast<const float4*>(p + r.d));
        // (fused halos: these may be halo planes a neighbour wrote during this frame: no non-coherent load)
        const float4 pf = FUSED ? __ldcg(reinterpret_cast<const float4*>(p + r.f)) : __ldg(reinterpret_cast<const float4*>(p + r.f));
        const float4 pb = FUSED ? __ldcg(reinterpret_cast<const float4*>(p + r.b)) : __ldg(reinterpret_cast<const float4*>(p + r.b));
        const unsigned row = r.c - x0;
        const float px[6] = {__ldg(p + row + r.xl), pc.x, pc.y, pc.z, pc.w, __ldg(p + row + r.xr)};
        const float pU[4] = {pu.x, pu.y, pu.z, pu.w}, pD[4] = {pd.x, pd.y, pd.z, pd.w};
        const float pF[4] = {pf.x, pf.y, pf.z, pf.w}, pB[4] = {pb.x, pb.y, pb.z, pb.w};
        const float4 bpx4 = __ldg(reinterpret_cast<const float4*>(tab.bp[0] + x0));
        const float4 wx4 = __ldg(reinterpret_cast<const float4*>(tab.wall[0] + x0));
        const float bpx[4] = {bpx4.x, bpx4.y, bpx4.z, bpx4.w}, wx[4] = {wx4.x, wx4.y, wx4.z, wx4.w};
        const float bpy = __ldg(tab.bp[1] + y), wy = __ldg(tab.wall[1] + y);
        const float bpz = __ldg(tab.bp[2] + z), wz = __ldg(tab.wall[2] + z);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float gx = -px[j] + px[j + 2];
            const float gy = -pU[j] + pD[j];
            const float gz = -pF[j] + pB[j];
            float a = __fmaf_rn(-gx, 1.04166675f, ux[j]);
            float b = __fmaf_rn(-gy, 1.04166675f, uy[j]);
            float c = __fmaf_rn(-gz, 1.04166675f, uz[j]);
            a = a * ((0.0f < a * bpx[j]) ? wx[j] : 1.0f);
            b = b * ((0.0f < b * bpy) ? wy : 1.0f);
            c = c * ((0.0f < c * bpz) ? wz : 1.0f);
            ux[j] = a; uy[j] = b; uz[j] = c;
        }
    }
    uint4 oa, ob;
    uint2 t;
    t = pack_texel4(ux[0], uy[0], uz[0], 0.0f); oa.x = t.x; oa.y = t.y;
    t = pack_texel4(ux[1], uy[1], uz[1], 0.0f); oa.z = t.x; oa.w = t.y;
    t = pack_texel4(ux[2], uy[2], uz[2], 0.0f); ob.x = t.x; ob.y = t.y;
    t = pack_texel4(ux[3], uy[3], uz[3], 0.0f); ob.z = t.x; ob.w = t.y;
    uint4* o = reinterpret_cast<uint4*>(vel_out + r.c);
    o[0] = oa;
    o[1] = ob;
    const long long plane = (long long)d.nx * d.ny;
    if (FUSED && near_lo) {
        uint4* q = reinterpret_cast<uint4*>(vel_out_lo + (long long)r.c + pv.dz_lo * plane);
        q[0] = oa;
        q[1] = ob;
    }
    if (FUSED && near_hi) {
        uint4* q = reinterpret_cast<uint4*>(vel_out_hi + (long long)r.c + pv.dz_hi * plane);
        q[0] = oa;
        q[1] = ob;
    }
}

inline dim3 quad_grid(const Domain& d) { return dim3((d.nx / 4 + 31) / 32, (d.ny + 7) / 8, d.z_own1 - d.z_own0); }

}  // namespace

bool quad_kernels_supported(const Domain& d) { return d.nz > 1 && (d.nx % 8) == 0; }

void launch_divergence_quad(const Domain& d, const FrameParams* frame, const void* vel, float* rhs, const PeerView& pv,
                            float* rhs_lo, float* rhs_hi, int push_depth, cudaStream_t stream) {
    const dim3 grid((d.nx / 4 + 31) / 32, (d.ny + 7) / 8, (d.z_own1 - d.z_own0 + kDivPlanes - 1) / kDivPlanes);
    if (pv.has_lo || pv.has_hi)
        divergence_quad_kernel<true><<<grid, dim3(32, 8), 0, stream>>>(d, pv, frame, (const uint2*)vel, rhs, rhs_lo, rhs_hi,
                                                                       push_depth);
    else
        divergence_quad_kernel<false><<<grid, dim3(32, 8), 0, stream>>>(d, pv, frame, (const uint2*)vel, rhs, rhs_lo,
                                                                        rhs_hi, push_depth);
}

void launch_gradient_quad(const Domain& d, const AxisTables& tab, const FrameParams* frame, const void* vel_in,
                          const float* p0, const float* p1, void* vel_out, const StepState* state, const PeerView& pv,
                          void* vel_out_lo, void* vel_out_hi, int reach, int event, cudaStream_t stream) {
    if (pv.has_lo || pv.has_hi)
        gradient_quad_kernel<true><<<quad_grid(d), dim3(32, 8), 0, stream>>>(d, tab, pv, frame, (const uint2*)vel_in, p0, p1,
                                                                             (uint2*)vel_out, state, (uint2*)vel_out_lo,
                                                                             (uint2*)vel_out_hi, reach, event);
    else
        gradient_quad_kernel<false><<<quad_grid(d), dim3(32, 8), 0, stream>>>(d, tab, pv, frame, (const uint2*)vel_in, p0,
                                                                              p1, (uint2*)vel_out, state, (uint2*)vel_out_lo,
                                                                              (uint2*)vel_out_hi, reach, event);
}

}  // namespace fxb
