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

// Peer-memory backend of the exchange (experimental, FXB_P2P=1): the neighbours' field buffers are mapped through
// CUDA IPC and one kernel per exchange stores this rank's face planes straight into their halo planes over NVLink,
// publishes an epoch flag in their memory and waits for theirs — no NCCL call on the step's critical path.
struct HaloP2P {
    static constexpr int kMaxBuffers = 12;
    bool enabled = false;
    int nbuf = 0;
    void* local[kMaxBuffers] = {};
    void* peer_lo[kMaxBuffers] = {};   // the same buffers of rank - 1 / rank + 1, mapped into this process
    void* peer_hi[kMaxBuffers] = {};
    unsigned long long* flags = nullptr;      // [8] in this rank's memory: [0] epoch published by rank - 1, [1] by
                                              // rank + 1, [2] this rank's exchange counter, [3] finished CTAs, [4] timeout
    unsigned long long* flags_lo = nullptr;   // the flag words of rank - 1 / rank + 1
    unsigned long long* flags_hi = nullptr;
    int z_first_lo = 0, z_first_hi = 0;       // Domain::z_first of the neighbours
};

// Where one face exchange of `depth` planes reads and writes, in LOCAL plane indices (pure arithmetic, no GPU):
// this rank sends its planes [send_lo, send_lo + depth) into rank - 1's array at [dst_lo, ...) and its planes
// [send_hi, send_hi + depth) into rank + 1's array at [dst_hi, ...).  z_first_*: global plane of local plane 0.
struct P2PPlanes {
    long long send_lo, dst_lo, send_hi, dst_hi;
};
P2PPlanes p2p_planes(const Domain& d, int depth, int z_first_lo, int z_first_hi);

struct HaloComm {
    void* comm = nullptr;  // ncclComm_t
    int rank = 0, nranks = 1;
    HaloP2P p2p;
    bool init(const void* unique_id128, int rank, int nranks);
    void destroy();
    bool exchange(const Domain& d, const HaloField* fields, int nfields, cudaStream_t stream);
    bool all_reduce_sum_u64(void* buf, size_t count, cudaStream_t stream);
    // Every rank's owned planes of a GLOBAL array (plane 0 = global plane 0; this rank's planes already in place) are
    // sent to every other rank: one ncclSend/ncclRecv pair per peer inside one group.  Slab r owns planes
    // [r nz / R, (r + 1) nz / R).  Used by the light-map pass, whose rays cross every slab.
    bool all_gather_slabs(void* global_base, size_t plane_bytes, int nz, cudaStream_t stream);
    // Maps the listed device buffers (cudaMalloc'ed, one per exchanged field) of both z neighbours; the handles travel
    // over the NCCL communicator.  z_first_lo / z_first_hi: global plane of local plane 0 on rank - 1 / rank + 1.
    bool p2p_init(void* const* buffers, int nbuffers, int z_first_lo, int z_first_hi, cudaStream_t stream);
    // 1 when a wait of the peer-memory exchange gave up (a neighbour never published its epoch); sticky.
    int p2p_timed_out() const;
    // Fused halos (common.cuh PeerView): the event words live next to the exchange's flag words ([8], [9]); the
    // neighbours' copy of a registered local buffer (side 0: rank - 1, side 1: rank + 1), nullptr at a grid face.
    PeerView peer_view(const Domain& d) const;
    void* peer_of(const void* local, int side) const;
};

bool halo_unique_id(void* out128);
const std::string& halo_last_error();

}  // namespace fxb
