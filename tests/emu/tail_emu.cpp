// tail_emu.cpp — CPU emulation of jacobi_tail_kernel (fluidx12_b200/csrc/jacobi_tail.cu).  TEST INFRASTRUCTURE ONLY.
//
// Compiles the kernel's real body (jacobi_tail_body.cuh, including the per-item sequence of phases and barriers,
// tail_run_item) with g++: the phases between two barriers are run thread after thread (each thread runs the whole
// barrier-free segment before the next one starts, in ascending or descending order — the most skewed interleaving
// a missing __syncthreads() would allow), shared memory becomes a heap block, atomics become plain
// read-modify-writes (CTAs run one after the other).  tests/test_tail_emu.py drives it against the oracle
// (oracle/fluid_oracle.cpp) bit for bit, so the indexing, clamp and freeze logic of the CUDA kernel is checked
// on the CPU; the product library never links or loads this file.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../fluidx12_b200/csrc/jacobi_tail_body.cuh"

namespace {

using namespace fxb;

long long g_paths[3] = {0, 0, 0};  // items that took the copy / sparse / dense path
long long g_smem_overruns = 0;     // work items that wrote outside the S::kBytes the CUDA launch requests
bool g_descending = false;         // thread order inside a barrier-free segment

template <class S>
void run_item(const TailParams& P, const TailWork& W, int brick, int sub, const float* p_in, float* p_out,
              const float* rhs, const unsigned char* m_in, unsigned char* m_out, unsigned long long* active_after_s0) {
    // the kernel's dynamic shared memory is exactly S::kBytes: guard bands on both sides catch any write outside it
    constexpr size_t kGuard = 4096;
    std::vector<unsigned char> block(S::kBytes + 2 * kGuard + 16, 0xA5);
    unsigned char* base = block.data() + kGuard;
    base += (16 - reinterpret_cast<uintptr_t>(base) % 16) % 16;
    std::fill(base, base + S::kBytes, (unsigned char)0);
    const TailShared<S> sh = tail_shared<S>(base);
    // poison the staged data so that a read of something never written shows up as a mismatch
    for (int i = 0; i < S::kPFloats + S::kRhsFloats; ++i) sh.p[i] = 1.0e30f;
    for (int i = 0; i < S::kCtrlWords; ++i) sh.ctrl[i] = 0xDEADBEEFu;
    TailEmu<S> emu;
    emu.descending = g_descending;
    TailTma tma;
    const int path = tail_run_item<S>(emu, sh, P, W, brick, sub, p_in, p_out, rhs, m_in, m_out, active_after_s0, nullptr, tma, P.levels);
    ++g_paths[path];
    bool clean = true;
    for (unsigned char* q = block.data(); q < base; ++q) clean = clean && *q == 0xA5;
    for (unsigned char* q = base + S::kBytes; q < block.data() + block.size(); ++q) clean = clean && *q == 0xA5;
    if (!clean) ++g_smem_overruns;
}

}  // namespace

extern "C" {

// geom = {nx, ny, nz_alloc, z_face_lo, z_face_hi, z_out0, z_out1, bx, by, bz}
// flags = {first, early_exit, levels, tt, sparse_cap (-1: the compiled capacity), cp_async, dense_mode}
// Returns the number of work items processed, or -1 for an unsupported shape.
int tail_emu_launch(const int* geom, const int* flags, const float* p_in, float* p_out, const float* rhs,
                    const unsigned char* m_in, unsigned char* m_out, const int* relax_in, int n_relax,
                    const int* copy_in, int n_copy, int* relax_out, int* relax_out_count, int* copy_out,
                    int* copy_out_count, int* brick_state, unsigned long long* active_after_s0) {
    TailParams P;
    P.nx = geom[0]; P.ny = geom[1]; P.nz_alloc = geom[2];
    P.z_face_lo = geom[3]; P.z_face_hi = geom[4]; P.z_out0 = geom[5]; P.z_out1 = geom[6];
    P.bx = geom[7]; P.by = geom[8]; P.bz = geom[9];
    P.ntx = (P.nx + P.bx - 1) / P.bx;
    P.nty = (P.ny + P.by - 1) / P.by;
    P.first = flags[0]; P.early_exit = flags[1]; P.levels = flags[2];
    const int tt = flags[3];
    using S4 = TailShape<4, 10, 12, 8>;
    using S2 = TailShape<2, 10, 12, 8>;
    const int cap = tt == 4 ? S4::kListCap : S2::kListCap;
    P.sparse_cap = flags[4] < 0 || flags[4] > cap ? cap : flags[4];
    P.cp_async = flags[5];
    P.dense_mode = flags[6];
    if (P.bx % S4::OX != 0 || P.by > S4::OY || P.bz > S4::OZ || P.nx % 8 != 0 || P.levels > tt) return -1;
    P.nsub = P.bx / S4::OX;
    TailWork W;
    W.relax_in = relax_in; W.copy_in = copy_in; W.n_relax = n_relax; W.n_copy = n_copy;
    W.relax_out = relax_out; W.copy_out = copy_out;
    W.relax_out_count = relax_out_count; W.copy_out_count = copy_out_count;
    W.brick_state = brick_state;
    const int items = n_copy + n_relax * P.nsub;
    for (int item = 0; item < items; ++item) {
        if (item < n_copy) {
            for (int tid = 0; tid < S4::kThreads; ++tid)
                tail_copy_brick(tid, S4::kThreads, P, copy_in[item], p_in, p_out, m_out);
            continue;
        }
        const int r = item - n_copy;
        if (tt == 4) run_item<S4>(P, W, relax_in[r / P.nsub], r % P.nsub, p_in, p_out, rhs, m_in, m_out, active_after_s0);
        else if (tt == 2) run_item<S2>(P, W, relax_in[r / P.nsub], r % P.nsub, p_in, p_out, rhs, m_in, m_out, active_after_s0);
        else return -1;
    }
    return items;
}

// Thread order inside a barrier-free segment: 0 ascending, 1 descending.
void tail_emu_thread_order(int descending) { g_descending = descending != 0; }

// Items that took the copy / sparse / dense path since the last call (and resets the counters).
void tail_emu_paths(long long* out3) {
    for (int i = 0; i < 3; ++i) { out3[i] = g_paths[i]; g_paths[i] = 0; }
}

// Work items that wrote outside the kernel's S::kBytes of shared memory since the last call (and resets the counter).
long long tail_emu_smem_overruns() {
    const long long n = g_smem_overruns;
    g_smem_overruns = 0;
    return n;
}

}  // extern "C"
