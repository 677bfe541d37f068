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

#include <cstring>

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
    if (comm && g_api.comm_destroy) g_api.comm_destroy(comm);
    comm = nullptr;
}

// Exchanges `depth` planes of every listed field with the z-1 and z+1 neighbours.
bool HaloComm::exchange(const Domain& d, const HaloField* fields, int nfields, cudaStream_t stream) {
    if (nranks <= 1) return true;
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

bool HaloComm::all_reduce_sum_u64(void* buf, size_t count, cudaStream_t stream) {
    if (nranks <= 1) return true;
    return ok(g_api.all_reduce(buf, buf, count, kNcclUint64, kNcclSum, comm, stream), "ncclAllReduce");
}

}  // namespace fxb
