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
#include "halo.h"
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

// dens: the density of the WHOLE grid; this launch writes the voxels of global planes [z0, z1) (the rank's slab),
// out plane 0 = plane z0
__global__ void __launch_bounds__(512) light_map_kernel(const unsigned short* __restrict__ dens,
                                                        unsigned* __restrict__ out, const LightGeom g, const int z0,
                                                        const int z1, const __grid_constant__ LightConsts P) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 4 + threadIdx.y;
    const int z = z0 + blockIdx.z * 4 + threadIdx.z;
    if (x >= g.nx || y >= g.ny || z >= z1) return;
    out[((size_t)(z - z0) * g.ny + y) * g.nx + x] = light_map_voxel(dens, g, P, x, y, z);
}

}  // namespace

// colour.w of n voxels into the compact half array (also used by the non-separated ray march, raymarch.cu)
cudaError_t launch_extract_density(const void* colour, unsigned short* dens, size_t n, cudaStream_t stream) {
    const size_t threads = n / 4 + 1;
    extract_density_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(static_cast<const uint2*>(colour), dens, n);
    return cudaGetLastError();
}

// colour_own: the rank's owned planes of the colour field; dens: density array of the whole grid (plane 0 = global
// plane 0).  With several ranks the owned planes are extracted in place and every other slab arrives over NCCL
// (2 bytes per voxel of the grid per rank: a light ray crosses every slab).
cudaError_t launch_light_map(const Domain& d, const void* colour_own, unsigned short* dens, unsigned* out,
                             const void* consts, HaloComm* comm, cudaStream_t stream) {
    const size_t plane = (size_t)d.nx * d.ny;
    const size_t n = plane * (d.z_own1 - d.z_own0);
    const size_t threads = n / 4 + 1;
    // the 16-byte loads and 8-byte stores of the extraction start at a plane boundary of the allocations: fxb_light_map
    // requires nx * ny % 4 == 0 when nranks > 1
    extract_density_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>(static_cast<const uint2*>(colour_own),
                                                                                 dens + plane * d.z_own0, n);
    if (comm && comm->nranks > 1 && !comm->all_gather_slabs(dens, plane * sizeof(unsigned short), d.nz, stream))
        return cudaErrorUnknown;
    const LightGeom g{d.nx, d.ny, d.nz};
    const dim3 block(32, 4, 4);
    const dim3 grid((d.nx + 31) / 32, (d.ny + 3) / 4, (d.z_own1 - d.z_own0 + 3) / 4);
    light_map_kernel<<<grid, block, 0, stream>>>(dens, out, g, d.z_own0, d.z_own1, *static_cast<const LightConsts*>(consts));
    return cudaGetLastError();
}

}  // namespace fxb
