// halo.h — z-slab halo exchange (see halo.cu).
#pragma once

#include <cuda_runtime.h>

#include <string>

#include "common.cuh"

namespace fxb {

struct HaloField {
    void* base;          // local array, plane 0 = global plane Domain::z_first
    size_t plane_bytes;  // bytes of one z plane
    int depth;           // planes exchanged on each interior face
};

struct HaloComm {
    void* comm = nullptr;  // ncclComm_t
    int rank = 0, nranks = 1;
    bool init(const void* unique_id128, int rank, int nranks);
    void destroy();
    bool exchange(const Domain& d, const HaloField* fields, int nfields, cudaStream_t stream);
    bool all_reduce_sum_u64(void* buf, size_t count, cudaStream_t stream);
};

bool halo_unique_id(void* out128);
const std::string& halo_last_error();

}  // namespace fxb
