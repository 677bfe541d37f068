// raymarch_emu.cpp — CPU run of the cube-map ray-march kernel's per-texel body (raymarch_body.cuh).  TEST
// INFRASTRUCTURE: compared bit for bit with the oracle and the golden vectors by tests/test_raymarch_emu.py.
#include <cstdint>

#include "../../fluidx12_b200/csrc/raymarch_body.cuh"

extern "C" {

// colour: [nz][ny][nx] 8-byte texels; light_map: [nz][ny][nx] words; params: fxb::ViewConsts; cube: [6][S][S] words
void raymarch_emu_run(int nx, int ny, int nz, const uint64_t* colour, const uint32_t* light_map, const void* params,
                      uint32_t* cube) {
    using namespace fxb;
    const LightGeom g{nx, ny, nz};
    const ViewConsts& P = *static_cast<const ViewConsts*>(params);
    const int S = (int)P.cube_size;
    for (int face = 0; face < 6; ++face)
        for (int y = 0; y < S; ++y)
            for (int x = 0; x < S; ++x) {
                unsigned w;
                if (ray_march_texel<true>(reinterpret_cast<const RU2*>(colour), light_map, nullptr, g, P, nullptr, x, y, face, &w))
                    cube[((size_t)face * S + y) * S + x] = w;
            }
}

// The non-separated march (CSRayMarch): density = colour.w of every voxel as half bits (what extract_density_kernel
// writes), light = fxb::LightConsts with num_samples = light-ray samples.
void raymarch_emu_run_full(int nx, int ny, int nz, const uint64_t* colour, const uint16_t* density, const void* view,
                           const void* light, uint32_t* cube) {
    using namespace fxb;
    const LightGeom g{nx, ny, nz};
    const ViewConsts& P = *static_cast<const ViewConsts*>(view);
    const LightConsts& LP = *static_cast<const LightConsts*>(light);
    const int S = (int)P.cube_size;
    for (int face = 0; face < 6; ++face)
        for (int y = 0; y < S; ++y)
            for (int x = 0; x < S; ++x) {
                unsigned w;
                if (ray_march_texel<false>(reinterpret_cast<const RU2*>(colour), nullptr, density, g, P, &LP, x, y, face, &w))
                    cube[((size_t)face * S + y) * S + x] = w;
            }
}

}  // extern "C"
