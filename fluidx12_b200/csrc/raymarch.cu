// raymarch.cu — cube-map-space view-ray march with the separate light pass: Fluid::rayMarchV
// (Fluid.cpp:880-908) dispatching CSRayMarchV over the six faces of mip m_cubeMapLOD of the R8G8B8A8_UNORM cube map.
// One thread per cube texel, CTA 8 x 8 x 1 as the reference's thread group (neighbouring rays of a face stay in one
// CTA, so their taps share L1 lines); a ray fetches the colour field (eight 8-byte texels per sample) and, where there
// is smoke, the light map written by lightmap.cu (eight 4-byte R11G11B10_FLOAT words).  6 S^2 rays of at most
// num_samples samples: small next to the simulation step (S = 256, 192 samples: < 0.1 G fetches), latency-bound.
// A second kernel serves the non-separated mode (CSRayMarch, Fluid::rayMarch, Fluid.cpp:825-855), which casts the light
// ray — and with light probes the occlusion ray — at every view sample over the compact density array of lightmap.cu.
#include "raymarch_body.cuh"
#include "kernels.h"

namespace fxb {
namespace {

__global__ void __launch_bounds__(64) ray_march_v_kernel(const uint2* __restrict__ colour,
                                                         const unsigned* __restrict__ light_map,
                                                         unsigned* __restrict__ cube, const LightGeom g,
                                                         const __grid_constant__ ViewConsts P) {
    const int x = blockIdx.x * 8 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int face = blockIdx.z;
    const int S = (int)P.cube_size;
    if (x >= S || y >= S) return;
    unsigned w;
    if (ray_march_texel<true>(colour, light_map, nullptr, g, P, nullptr, x, y, face, &w))
        cube[((size_t)face * S + y) * S + x] = w;
}

// The non-separated march (CSRayMarch, Fluid::rayMarch): light, occlusion ray and SH irradiance at every view sample.
__global__ void __launch_bounds__(64) ray_march_kernel(const uint2* __restrict__ colour,
                                                       const unsigned short* __restrict__ dens,
                                                       unsigned* __restrict__ cube, const LightGeom g,
                                                       const __grid_constant__ ViewConsts P,
                                                       const __grid_constant__ LightConsts LP) {
    const int x = blockIdx.x * 8 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int face = blockIdx.z;
    const int S = (int)P.cube_size;
    if (x >= S || y >= S) return;
    unsigned w;
    if (ray_march_texel<false>(colour, nullptr, dens, g, P, &LP, x, y, face, &w)) cube[((size_t)face * S + y) * S + x] = w;
}

}  // namespace

cudaError_t launch_ray_march_v(const Domain& d, const void* colour, const unsigned* light_map, unsigned* cube,
                               const void* consts, cudaStream_t stream) {
    const ViewConsts& P = *static_cast<const ViewConsts*>(consts);
    const LightGeom g{d.nx, d.ny, d.nz};
    const dim3 block(8, 8, 1);
    const dim3 grid((P.cube_size + 7) / 8, (P.cube_size + 7) / 8, 6);
    ray_march_v_kernel<<<grid, block, 0, stream>>>(static_cast<const uint2*>(colour), light_map, cube, g, P);
    return cudaGetLastError();
}

cudaError_t launch_ray_march(const Domain& d, const void* colour, const unsigned short* dens, unsigned* cube,
                             const void* view, const void* light, cudaStream_t stream) {
    const ViewConsts& P = *static_cast<const ViewConsts*>(view);
    const LightGeom g{d.nx, d.ny, d.nz};
    const dim3 block(8, 8, 1);
    const dim3 grid((P.cube_size + 7) / 8, (P.cube_size + 7) / 8, 6);
    ray_march_kernel<<<grid, block, 0, stream>>>(static_cast<const uint2*>(colour), dens, cube, g, P,
                                                 *static_cast<const LightConsts*>(light));
    return cudaGetLastError();
}

}  // namespace fxb
