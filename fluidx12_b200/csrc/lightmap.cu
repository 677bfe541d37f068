// lightmap.cu — the light-map pass that follows the simulation step in the reference's default render mode
// (SURVEY.md §8 f1): Fluid::rayMarchL (Fluid.cpp:857-878) dispatching CSRayMarchL.hlsl:15-80 over the colour field,
// one R11G11B10_FLOAT texel per voxel (Fluid.cpp:223-227).
//
// Two kernels:
//   extract_density_kernel  colour.w of every voxel into a compact half array (the only channel the pass samples):
//                           8 B in + 2 B out per voxel, pure streaming, 128-bit loads;
//   light_map_kernel        one thread per voxel, CTA 32 x 4 x 4 like the advection kernel so that neighbouring rays
//                           share their taps through L1; per voxel one fetch, and where there is smoke an adaptive
//                           march of up to num_samples fetches towards the light, six offset fetches for the density
//                           gradient, the SH irradiance and a second march along the gradient (lightmap_body.cuh).
// Algorithmic traffic: 12 B per voxel (colour in 8, light map out 4) + 4 B for the density scratch; the marches hit
// L1/L2 (the half array of a 256^3 grid is 32 MiB).  Work per voxel is data dependent (empty voxels leave after one
// fetch), so this pass is latency/issue bound where the plume is and streaming elsewhere.
#include "lightmap_body.cuh"
#include "kernels.h"

namespace fxb {
namespace {

__global__ void __launch_bounds__(256) extract_density_kernel(const uint2* __restrict__ colour,
                                                              unsigned short* __restrict__ dens, const size_t n) {
    const size_t quads = n / 4;
    const size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (q < quads) {
        const uint4* src = reinterpret_cast<const uint4*>(colour) + 2 * q;  // four 8-byte texels
        const uint4 a = __ldg(src), b = __ldg(src + 1);
        uint2 w;
        w.x = (a.y >> 16) | (a.w & 0xFFFF0000u);
        w.y = (b.y >> 16) | (b.w & 0xFFFF0000u);
        reinterpret_cast<uint2*>(dens)[q] = w;
    } else if (q == quads) {
        for (size_t i = quads * 4; i < n; ++i) dens[i] = (unsigned short)(__ldg(&colour[i].y) >> 16);
    }
}

__global__ void __launch_bounds__(512) light_map_kernel(const unsigned short* __restrict__ dens,
                                                        unsigned* __restrict__ out, const LightGeom g,
                                                        const __grid_constant__ LightConsts P) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int z = blockIdx.z * 4 + threadIdx.z;
    if (x >= g.nx || y >= g.ny || z >= g.nz) return;
    out[((size_t)z * g.ny + y) * g.nx + x] = light_map_voxel(dens, g, P, x, y, z);
}

}  // namespace

cudaError_t launch_light_map(const Domain& d, const void* colour, unsigned short* dens, unsigned* out,
                             const void* consts, cudaStream_t stream) {
    const size_t n = (size_t)d.nx * d.ny * d.nz;
    const size_t threads = n / 4 + 1;
    extract_density_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(static_cast<const uint2*>(colour), dens, n);
    const LightGeom g{d.nx, d.ny, d.nz};
    const dim3 block(32, 4, 4);
    const dim3 grid((d.nx + 31) / 32, (d.ny + 3) / 4, (d.nz + 3) / 4);
    light_map_kernel<<<grid, block, 0, stream>>>(dens, out, g, *static_cast<const LightConsts*>(consts));
    return cudaGetLastError();
}

}  // namespace fxb
