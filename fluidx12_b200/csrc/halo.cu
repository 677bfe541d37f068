// halo.cu — z-slab face-halo exchange over NCCL send/recv (NVLink 5 / NVSwitch) for the multi-GPU step.
//
// The reference has no multi-GPU path (single adapter, nodeMask 0: FluidX12/FluidX12.cpp:139-141); the slab
// decomposition is this build's addition.  Fields are stored x-fastest, so the h planes next to
// a slab face are one contiguous range: an exchange is a plain pointer + count ncclSend / ncclRecv pair per
// neighbour and field, grouped so both directions of every field run concurrently.  NCCL is loaded at run time
// (dlopen of the libnccl.so.2 already in the process, e.g. the one bundled with PyTorch) and only when
// nranks > 1, so single-GPU use has no NCCL dependency.
#include "halo.h"

#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <vector>

namespace fxb {

namespace {

// Minimal NCCL ABI (stable since NCCL 2.x): opaque communicator, 128-byte unique id, enums as ints.
struct UniqueId { char internal[128]; };
typedef int (*GetUniqueIdFn)(UniqueId*);
typedef int (*CommInitRankFn)(void**, int, UniqueId, int);
typedef int (*CommDestroyFn)(void*);
typedef int (*SendFn)(const void*, size_t, int, int, void*, cudaStream_t);
typedef int (*RecvFn)(void*, size_t, int, int, void*, cudaStream_t);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*GroupFn)();
typedef const char* (*ErrStrFn)(int);

constexpr int kNcclInt8 = 0;    // ncclInt8 / ncclChar
constexpr int kNcclUint64 = 5;  // ncclUint64
constexpr int kNcclSum = 0;     // ncclSum

struct Api {
    void* lib = nullptr;
    GetUniqueIdFn get_unique_id = nullptr;
    CommInitRankFn comm_init_rank = nullptr;
    CommDestroyFn comm_destroy = nullptr;
    SendFn send = nullptr;
    RecvFn recv = nullptr;
    AllReduceFn all_reduce = nullptr;
    GroupFn group_start = nullptr, group_end = nullptr;
    ErrStrFn err_str = nullptr;
};

Api g_api;
std::string g_err;

bool load_api() {
    if (g_api.lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* lib = nullptr;
    for (const char* n : names) {
        lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (lib) break;
    }
    if (!lib) {
        g_err = std::string("cannot load libnccl.so.2: ") + dlerror();
        return false;
    }
    Api a;
    a.lib = lib;
    a.get_unique_id = (GetUniqueIdFn)dlsym(lib, "ncclGetUniqueId");
    a.comm_init_rank = (CommInitRankFn)dlsym(lib, "ncclCommInitRank");
    a.comm_destroy = (CommDestroyFn)dlsym(lib, "ncclCommDestroy");
    a.send = (SendFn)dlsym(lib, "ncclSend");
    a.recv = (RecvFn)dlsym(lib, "ncclRecv");
    a.all_reduce = (AllReduceFn)dlsym(lib, "ncclAllReduce");
    a.group_start = (GroupFn)dlsym(lib, "ncclGroupStart");
    a.group_end = (GroupFn)dlsym(lib, "ncclGroupEnd");
    a.err_str = (ErrStrFn)dlsym(lib, "ncclGetErrorString");
    if (!a.get_unique_id || !a.comm_init_rank || !a.comm_destroy || !a.send || !a.recv || !a.all_reduce ||
        !a.group_start || !a.group_end || !a.err_str) {
        g_err = "libnccl.so.2 lacks a required symbol";
        return false;
    }
    g_api = a;
    return true;
}

bool ok(int rc, const char* what) {
    if (rc == 0) return true;
    g_err = std::string(what) + ": " + (g_api.err_str ? g_api.err_str(rc) : "NCCL error");
    return false;
}

}  // namespace

const std::string& halo_last_error() { return g_err; }

bool halo_unique_id(void* out128) {
    if (!load_api()) return false;
    UniqueId id;
    if (!ok(g_api.get_unique_id(&id), "ncclGetUniqueId")) return false;
    std::memcpy(out128, &id, sizeof(id));
    return true;
}

bool HaloComm::init(const void* unique_id128, int rank_, int nranks_) {
    if (!load_api()) return false;
    UniqueId id;
    std::memcpy(&id, unique_id128, sizeof(id));
    rank = rank_;
    nranks = nranks_;
    return ok(g_api.comm_init_rank(&comm, nranks, id, rank), "ncclCommInitRank");
}

void HaloComm::destroy() {
    if (p2p.enabled) {
        for (int i = 0; i < p2p.nbuf; ++i) {
            if (p2p.peer_lo[i]) cudaIpcCloseMemHandle(p2p.peer_lo[i]);
            if (p2p.peer_hi[i]) cudaIpcCloseMemHandle(p2p.peer_hi[i]);
        }
        if (p2p.flags_lo) cudaIpcCloseMemHandle(p2p.flags_lo);
        if (p2p.flags_hi) cudaIpcCloseMemHandle(p2p.flags_hi);
        cudaFree(p2p.flags);
        p2p = HaloP2P();
    }
    if (comm && g_api.comm_destroy) g_api.comm_destroy(comm);
    comm = nullptr;
}

// ---- peer-memory backend -------------------------------------------------------------------------------------------
namespace {

struct P2PJob {
    const char* src;
    char* dst;
    unsigned long long bytes;
};

struct P2PArgs {
    P2PJob job[8];
    int njobs;
    int vec;                          // bytes per copy element: 16, 8 or 4 (what every job's alignment allows)
    unsigned long long* flags;        // this rank's flag words (HaloP2P::flags)
    unsigned long long* flags_lo;     // rank - 1's, or nullptr at the grid's lower face
    unsigned long long* flags_hi;     // rank + 1's, or nullptr
    long long timeout_cycles;
};

template <class V>
__device__ __forceinline__ void p2p_copy(const P2PJob& j) {
    const V* __restrict__ src = reinterpret_cast<const V*>(j.src);
    V* __restrict__ dst = reinterpret_cast<V*>(j.dst);
    const size_t n = j.bytes / sizeof(V), stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = src[i];
}

__device__ __forceinline__ void p2p_wait(volatile unsigned long long* flag, unsigned long long epoch, long long limit,
                                         unsigned long long* timed_out) {
    const long long t0 = clock64();
    while (*flag < epoch) {
        if (clock64() - t0 > limit) {  // never hang the device: record the failure and go on (results are then wrong)
            *timed_out = 1ull;
            break;
        }
        __nanosleep(64);
    }
}

// One exchange: store the face planes into the neighbours' halo planes, then (last CTA) publish this rank's epoch
// in the neighbours' flag words and wait until both neighbours have published theirs.  Ordering: every CTA fences
// at system scope before it reports done, the last CTA fences again before the flag stores, so a neighbour that sees
// the epoch also sees the planes; its own consumers start after its own exchange kernel has finished.
// A neighbour can be at most one exchange ahead (it waits for this rank's epoch at every exchange), which together
// with the ping-pong of the pressure / mask buffers rules out overwriting planes that are still being read.
__global__ void __launch_bounds__(256) halo_p2p_kernel(const __grid_constant__ P2PArgs a) {
    for (int j = 0; j < a.njobs; ++j) {
        if (a.vec == 16) p2p_copy<uint4>(a.job[j]);
        else if (a.vec == 8) p2p_copy<uint2>(a.job[j]);
        else p2p_copy<unsigned>(a.job[j]);
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x != 0) return;
    if (atomicAdd(&a.flags[3], 1ull) != (unsigned long long)gridDim.x - 1ull) return;
    a.flags[3] = 0ull;
    const unsigned long long epoch = a.flags[2] + 1ull;
    a.flags[2] = epoch;
    __threadfence_system();
    if (a.flags_lo) *reinterpret_cast<volatile unsigned long long*>(a.flags_lo + 1) = epoch;  // this rank is its upper neighbour
    if (a.flags_hi) *reinterpret_cast<volatile unsigned long long*>(a.flags_hi + 0) = epoch;
    __threadfence_system();
    if (a.flags_lo) p2p_wait(a.flags + 0, epoch, a.timeout_cycles, a.flags + 4);
    if (a.flags_hi) p2p_wait(a.flags + 1, epoch, a.timeout_cycles, a.flags + 4);
    __threadfence_system();
}

}  // namespace

P2PPlanes p2p_planes(const Domain& d, int depth, int z_first_lo, int z_first_hi) {
    P2PPlanes q;
    // my lowest planes are rank - 1's upper halo: global planes [z_own0, z_own0 + depth)
    q.send_lo = d.z_own0 - d.z_first;
    q.dst_lo = d.z_own0 - z_first_lo;
    // my highest planes are rank + 1's lower halo: global planes [z_own1 - depth, z_own1)
    q.send_hi = d.z_own1 - depth - d.z_first;
    q.dst_hi = d.z_own1 - depth - z_first_hi;
    return q;
}

bool HaloComm::p2p_init(void* const* buffers, int nbuffers, int z_first_lo, int z_first_hi, cudaStream_t stream) {
    if (nranks <= 1 || nbuffers <= 0 || nbuffers > HaloP2P::kMaxBuffers) return false;
    auto cuda_ok = [&](cudaError_t e, const char* what) {
        if (e == cudaSuccess) return true;
        g_err = std::string(what) + ": " + cudaGetErrorString(e);
        return false;
    };
    HaloP2P h;
    h.nbuf = nbuffers;
    h.z_first_lo = z_first_lo;
    h.z_first_hi = z_first_hi;
    // 2 MiB so that the flag words are an allocation of their own (an IPC handle maps a whole allocation)
    if (!cuda_ok(cudaMalloc((void**)&h.flags, 2u << 20), "cudaMalloc(flags)")) return false;
    if (!cuda_ok(cudaMemset(h.flags, 0, 2u << 20), "cudaMemset(flags)")) return false;
    // handles of this rank: the field buffers, then the flag words
    const int nh = nbuffers + 1;
    std::vector<cudaIpcMemHandle_t> mine(nh), from_lo(nh), from_hi(nh);
    for (int i = 0; i < nbuffers; ++i) {
        h.local[i] = buffers[i];
        if (!cuda_ok(cudaIpcGetMemHandle(&mine[i], buffers[i]), "cudaIpcGetMemHandle")) return false;
    }
    if (!cuda_ok(cudaIpcGetMemHandle(&mine[nbuffers], h.flags), "cudaIpcGetMemHandle(flags)")) return false;
    // ship them to both neighbours over the communicator that already exists (device staging buffers)
    const size_t bytes = (size_t)nh * sizeof(cudaIpcMemHandle_t);
    char* stage = nullptr;
    if (!cuda_ok(cudaMalloc((void**)&stage, 3 * bytes), "cudaMalloc(stage)")) return false;
    bool good = cuda_ok(cudaMemcpyAsync(stage, mine.data(), bytes, cudaMemcpyHostToDevice, stream), "cudaMemcpyAsync");
    good = good && ok(g_api.group_start(), "ncclGroupStart");
    if (good && rank > 0) {
        good = good && ok(g_api.send(stage, bytes, kNcclInt8, rank - 1, comm, stream), "ncclSend");
        good = good && ok(g_api.recv(stage + bytes, bytes, kNcclInt8, rank - 1, comm, stream), "ncclRecv");
    }
    if (good && rank < nranks - 1) {
        good = good && ok(g_api.send(stage, bytes, kNcclInt8, rank + 1, comm, stream), "ncclSend");
        good = good && ok(g_api.recv(stage + 2 * bytes, bytes, kNcclInt8, rank + 1, comm, stream), "ncclRecv");
    }
    good = ok(g_api.group_end(), "ncclGroupEnd") && good;
    good = good && cuda_ok(cudaMemcpyAsync(from_lo.data(), stage + bytes, bytes, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync");
    good = good && cuda_ok(cudaMemcpyAsync(from_hi.data(), stage + 2 * bytes, bytes, cudaMemcpyDeviceToHost, stream), "cudaMemcpyAsync");
    good = good && cuda_ok(cudaStreamSynchronize(stream), "cudaStreamSynchronize");
    cudaFree(stage);
    if (!good) return false;
    for (int i = 0; i < nh; ++i) {
        void** lo = i < nbuffers ? &h.peer_lo[i] : (void**)&h.flags_lo;
        void** hi = i < nbuffers ? &h.peer_hi[i] : (void**)&h.flags_hi;
        if (rank > 0 && !cuda_ok(cudaIpcOpenMemHandle(lo, from_lo[i], cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle"))
            return false;
        if (rank < nranks - 1 &&
            !cuda_ok(cudaIpcOpenMemHandle(hi, from_hi[i], cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle"))
            return false;
    }
    h.enabled = true;
    p2p = h;
    return true;
}

int HaloComm::p2p_timed_out() const {
    if (!p2p.enabled) return 0;
    unsigned long long v = 0;
    if (cudaMemcpy(&v, p2p.flags + 4, sizeof(v), cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
    return v != 0ull;
}

PeerView HaloComm::peer_view(const Domain& d) const {
    PeerView pv{};
    if (!p2p.enabled) return pv;
    pv.has_lo = rank > 0 ? 1 : 0;
    pv.has_hi = rank < nranks - 1 ? 1 : 0;
    pv.dz_lo = d.z_first - p2p.z_first_lo;
    pv.dz_hi = d.z_first - p2p.z_first_hi;
    pv.events = p2p.flags + 8;
    pv.ev_lo = pv.has_lo ? p2p.flags_lo + 8 : nullptr;
    pv.ev_hi = pv.has_hi ? p2p.flags_hi + 8 : nullptr;
    pv.error = p2p.flags + 4;
    static const long long timeout_s = getenv("FXB_P2P_TIMEOUT_S") ? atoll(getenv("FXB_P2P_TIMEOUT_S")) : 30;
    pv.timeout_cycles = timeout_s * 2000000000ll;
    return pv;
}

void* HaloComm::peer_of(const void* local, int side) const {
    if (!p2p.enabled) return nullptr;
    for (int k = 0; k < p2p.nbuf; ++k)
        if (p2p.local[k] == local) return side == 0 ? p2p.peer_lo[k] : p2p.peer_hi[k];
    return nullptr;
}

// Exchanges `depth` planes of every listed field with the z-1 and z+1 neighbours.
bool HaloComm::exchange(const Domain& d, const HaloField* fields, int nfields, cudaStream_t stream) {
    if (nranks <= 1) return true;
    if (p2p.enabled) {
        P2PArgs a;
        a.njobs = 0;
        a.vec = 16;
        a.flags = p2p.flags;
        a.flags_lo = rank > 0 ? p2p.flags_lo : nullptr;
        a.flags_hi = rank < nranks - 1 ? p2p.flags_hi : nullptr;
        // A neighbour may legitimately be seconds behind (host-side work between steps differs per rank), so the
        // wait is generous; it only exists so that a broken run ends with an error instead of a hung device.
        static const long long timeout_s = getenv("FXB_P2P_TIMEOUT_S") ? atoll(getenv("FXB_P2P_TIMEOUT_S")) : 30;
        a.timeout_cycles = timeout_s * 2000000000ll;
        size_t total = 0;
        for (int i = 0; i < nfields; ++i) {
            const HaloField& f = fields[i];
            int b = -1;
            for (int k = 0; k < p2p.nbuf; ++k)
                if (p2p.local[k] == f.base) b = k;
            if (b < 0 || a.njobs + 2 > 8) {
                g_err = "p2p exchange: field buffer was not registered";
                return false;
            }
            const char* base = static_cast<const char*>(f.base);
            const size_t pb = f.plane_bytes, n = pb * f.depth;
            const P2PPlanes q = p2p_planes(d, f.depth, p2p.z_first_lo, p2p.z_first_hi);
            if (rank > 0) {
                P2PJob& j = a.job[a.njobs++];
                j.src = base + q.send_lo * (long long)pb;
                j.dst = static_cast<char*>(p2p.peer_lo[b]) + q.dst_lo * (long long)pb;
                j.bytes = n;
            }
            if (rank < nranks - 1) {
                P2PJob& j = a.job[a.njobs++];
                j.src = base + q.send_hi * (long long)pb;
                j.dst = static_cast<char*>(p2p.peer_hi[b]) + q.dst_hi * (long long)pb;
                j.bytes = n;
            }
            total += 2 * n;
        }
        for (int j = 0; j < a.njobs; ++j)
            while (a.vec > 4 && (((size_t)a.job[j].src | (size_t)a.job[j].dst | (size_t)a.job[j].bytes) % a.vec) != 0) a.vec >>= 1;
        int grid = (int)(total / (256 * 64));  // about 64 copy elements' worth of bytes per thread
        grid = grid < 1 ? 1 : (grid > 592 ? 592 : grid);
        halo_p2p_kernel<<<grid, 256, 0, stream>>>(a);
        if (cudaGetLastError() != cudaSuccess) {
            g_err = "halo_p2p_kernel launch failed";
            return false;
        }
        return true;
    }
    if (!ok(g_api.group_start(), "ncclGroupStart")) return false;
    bool good = true;
    for (int i = 0; i < nfields && good; ++i) {
        const HaloField& f = fields[i];
        char* base = static_cast<char*>(f.base);
        const size_t pb = f.plane_bytes, n = pb * f.depth;
        const size_t own0 = (size_t)(d.z_own0 - d.z_first), own1 = (size_t)(d.z_own1 - d.z_first);
        if (rank > 0) {  // lower neighbour owns the planes below z_own0
            good = good && ok(g_api.send(base + own0 * pb, n, kNcclInt8, rank - 1, comm, stream), "ncclSend");
            good = good && ok(g_api.recv(base + (own0 - f.depth) * pb, n, kNcclInt8, rank - 1, comm, stream), "ncclRecv");
        }
        if (rank < nranks - 1) {
            good = good && ok(g_api.send(base + (own1 - f.depth) * pb, n, kNcclInt8, rank + 1, comm, stream), "ncclSend");
            good = good && ok(g_api.recv(base + own1 * pb, n, kNcclInt8, rank + 1, comm, stream), "ncclRecv");
        }
    }
    const bool ended = ok(g_api.group_end(), "ncclGroupEnd");
    return good && ended;
}

bool HaloComm::all_gather_slabs(void* global_base, size_t plane_bytes, int nz, cudaStream_t stream) {
    if (nranks <= 1) return true;
    char* base = static_cast<char*>(global_base);
    auto z_of = [&](int r) { return (long long)r * nz / nranks; };
    const size_t mine = (size_t)(z_of(rank + 1) - z_of(rank)) * plane_bytes;
    if (!ok(g_api.group_start(), "ncclGroupStart")) return false;
    bool good = true;
    for (int peer = 0; peer < nranks && good; ++peer) {
        if (peer == rank) continue;
        const size_t theirs = (size_t)(z_of(peer + 1) - z_of(peer)) * plane_bytes;
        good = ok(g_api.send(base + (size_t)z_of(rank) * plane_bytes, mine, kNcclInt8, peer, comm, stream), "ncclSend") &&
               ok(g_api.recv(base + (size_t)z_of(peer) * plane_bytes, theirs, kNcclInt8, peer, comm, stream), "ncclRecv");
    }
    return ok(g_api.group_end(), "ncclGroupEnd") && good;
}

bool HaloComm::all_reduce_sum_u64(void* buf, size_t count, cudaStream_t stream) {
    if (nranks <= 1) return true;
    return ok(g_api.all_reduce(buf, buf, count, kNcclUint64, kNcclSum, comm, stream), "ncclAllReduce");
}

}  // namespace fxb
