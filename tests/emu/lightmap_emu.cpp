// lightmap_emu.cpp — CPU run of the light-map kernel's per-voxel body (lightmap_body.cuh).  TEST INFRASTRUCTURE.
//
// The CUDA kernel calls light_map_voxel once per thread on the compact density array its first kernel extracts from
// the colour field; here both steps run on the CPU and tests/test_lightmap_emu.py compares the result bit for bit with
// the oracle and with the golden vectors made from the reference's compiled shader.
#include <cstdint>

#include "../../fluidx12_b200/csrc/lightmap_body.cuh"

extern "C" {

// colour: [nz][ny][nx][4] half bits; params: fxb::LightConsts; scratch: nx*ny*nz half bits; out: nx*ny*nz words
void lightmap_emu_run(int nx, int ny, int nz, const uint16_t* colour, const void* params, uint16_t* scratch,
                      uint32_t* out) {
    using namespace fxb;
    const size_t n = (size_t)nx * ny * nz;
    for (size_t i = 0; i < n; ++i) scratch[i] = colour[4 * i + 3];  // extract_density_kernel
    const LightGeom g{nx, ny, nz};
    const LightConsts& P = *static_cast<const LightConsts*>(params);
    for (int z = 0; z < nz; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) out[((size_t)z * ny + y) * nx + x] = light_map_voxel(scratch, g, P, x, y, z);
}

// The z-slab form of the pass: `density` is the half array of the WHOLE grid (as gathered from all ranks), the voxels
// of global planes [z0, z1) are written, out plane 0 = plane z0 — what light_map_kernel does on one rank.
void lightmap_emu_run_slab(int nx, int ny, int nz, const uint16_t* density, const void* params, int z0, int z1,
                           uint32_t* out) {
    using namespace fxb;
    const LightGeom g{nx, ny, nz};
    const LightConsts& P = *static_cast<const LightConsts*>(params);
    for (int z = z0; z < z1; ++z)
        for (int y = 0; y < ny; ++y)
            for (int x = 0; x < nx; ++x) out[((size_t)(z - z0) * ny + y) * nx + x] = light_map_voxel(density, g, P, x, y, z);
}

}  // extern "C"
