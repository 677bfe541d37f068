// advect_emu.cpp — CPU run of the interior path of the second advection kernel (advect_body.cuh).  TEST INFRASTRUCTURE.
//
// The CUDA kernel calls advect_interior_voxel once per thread; here the same function is called once per voxel and
// the result is compared bit for bit with the oracle by tests/test_advect_emu.py.  Voxels the function declines
// (a tap outside the grid or the slab) are reported in `handled` and left to the first kernel's general code.
#include <cstdint>

#include "../../fluidx12_b200/csrc/advect_body.cuh"

extern "C" {

// geom = {nx, ny, nz, z_first, nz_alloc, z_own0, z_own1, ex0, ey0, ez0, ex1, ey1, ez1}; pos = three tables indexed by
// the global coordinate; arrays hold nz_alloc planes.  Returns the number of voxels handled.
long long advect_emu_run(const int* geom, const float* pos_x, const float* pos_y, const float* pos_z, const float* basis,
                         float dt, const uint64_t* vel_in, const uint64_t* col_in, uint64_t* vel_out, uint64_t* col_out,
                         unsigned char* handled) {
    using namespace fxb;
    AdvectGeom g;
    g.nx = geom[0]; g.ny = geom[1]; g.nz = geom[2]; g.z_first = geom[3]; g.nz_alloc = geom[4];
    g.pos[0] = pos_x; g.pos[1] = pos_y; g.pos[2] = pos_z;
    g.ex0 = geom[7]; g.ey0 = geom[8]; g.ez0 = geom[9]; g.ex1 = geom[10]; g.ey1 = geom[11]; g.ez1 = geom[12];
    g.basis = basis;
    long long n = 0;
    for (int z = geom[5]; z < geom[6]; ++z)
        for (int y = 0; y < g.ny; ++y)
            for (int x = 0; x < g.nx; ++x) {
                const bool ok = advect_interior_voxel(g, dt, reinterpret_cast<const AU2*>(vel_in),
                                                      reinterpret_cast<const AU2*>(col_in), reinterpret_cast<AU2*>(vel_out),
                                                      reinterpret_cast<AU2*>(col_out), x, y, z);
                handled[((size_t)(z - g.z_first) * g.ny + y) * g.nx + x] = ok ? 1 : 0;
                n += ok ? 1 : 0;
            }
    return n;
}

}  // extern "C"
