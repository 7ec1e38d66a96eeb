// common.cuh — complex helpers, bit twiddling and the handle type shared by all kernels.
// sm_100a only; built by __graft_entry__.build() / csrc/Makefile with
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "qsv.h"

// ---------------------------------------------------------------- complex math ----
template <typename R> struct CxT;
template <> struct CxT<float>  { using V = float2;  };
template <> struct CxT<double> { using V = double2; };

template <typename V> __device__ __forceinline__ V cx_make(decltype(V::x) re, decltype(V::x) im) {
    V r; r.x = re; r.y = im; return r;
}
// a*b
template <typename V> __device__ __forceinline__ V cx_mul(V a, V b) {
    return cx_make<V>(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x);
}
// acc + a*b  (4 FMAs)
template <typename V> __device__ __forceinline__ V cx_fma(V a, V b, V acc) {
    acc.x = fma(a.x, b.x, acc.x); acc.x = fma(-a.y, b.y, acc.x);
    acc.y = fma(a.x, b.y, acc.y); acc.y = fma(a.y, b.x, acc.y);
    return acc;
}

// 2x2 / 4x4 complex matrices passed by value as kernel parameters, in the compute type.
template <typename R> struct Mat2 { typename CxT<R>::V m[4]; };
template <typename R> struct Mat4 { typename CxT<R>::V m[16]; };

// --------------------------------------------------------------- bit twiddling ----
__host__ __device__ __forceinline__ uint64_t insert_zero_bit(uint64_t x, int pos) {
    const uint64_t low = x & ((1ull << pos) - 1ull);
    return ((x >> pos) << (pos + 1)) | low;
}

// ---------------------------------------------------------------------- handle ----
struct qsv_program {
    std::vector<qsv_pass> passes;
    std::vector<int>      op_offset;     // first op of each pass in d_ops
    std::vector<int>      fold_offset;   // first fold-table entry of each pass in d_tables
    double2 *d_tables = nullptr;
    qsv_op  *d_ops   = nullptr;          // device copy of all ops
    qsv_pass *d_passes = nullptr;        // device copy of pass descriptors
    cudaGraphExec_t graph = nullptr;
    // run-time specialised kernels (jit.cuh): one per pass, null = interpret the pass
    std::vector<cudaKernel_t> jit;
    std::vector<uint64_t> jit_keys;                   // entries of the kernel cache this program pins (jit.cuh)
    std::vector<std::vector<double>> jit_coefs;
    std::vector<std::vector<float>> jit_coefs_f;      // the same values for complex64 kernels
    // scatter passes (pass fused with the exchange after it): host copy of the ops for their
    // generator, kernels keyed by (pass index, swapped local bits)
    std::vector<qsv_op> h_ops;
    struct ScatterKernel { int pass_index; int n; int bits[3]; cudaKernel_t fn; };
    std::vector<ScatterKernel> scatter;
};

struct qsv_handle {
    int n_qubits = 0, n_local = 0, dtype = QSV_C128, device = 0, rank = 0, world = 1;
    int sm_count = 148;
    bool force_simple_pass = false;   // tests: run every pass on the one-CTA-per-tile kernel
    bool jit = true;                  // specialise the passes of a program at qsv_program_create
    size_t n_amps = 0;            // local amplitudes
    size_t amp_bytes = 16;
    void *d_state = nullptr;
    // the shard allocation carries a small TAIL (flags of the exchange kernels, csrc/xchg.cuh) that the
    // peers reach through the same mapping as the shard itself
    void *alloc_base = nullptr;                 // what cudaMalloc returned (d_state, unless a scatter pass swapped roles)
    unsigned long long *d_tail = nullptr;       // alloc_base + n_amps * amp_bytes
    // scatter passes write into a second buffer of the same size (allocated on demand); the two
    // exchange roles after every scatter pass.  scat_cur[r] / scat_other[r] = rank r's current and
    // second buffer as seen from this process (same-process pointers or CUDA IPC mappings).
    void *d_shadow = nullptr;
    std::vector<void *> scat_cur, scat_other;
    bool scat_ready = false, scat_ipc = false;
    cudaStream_t stream = nullptr;
    // asynchronous checkpoints (qsv_snapshot*): a device-side copy of the shard drained on a second stream
    cudaStream_t io_stream = nullptr;
    cudaEvent_t ev_snap = nullptr, ev_io = nullptr;
    bool io_pending = false;
    // scratch for reductions / one-shot passes
    double *d_partials = nullptr; size_t n_partials = 0;
    qsv_op *d_ops_scratch = nullptr; size_t ops_scratch_cap = 0;
    qsv_pass *d_pass_scratch = nullptr;
    // timing
    bool timing = false;
    struct TimedLaunch { cudaEvent_t a, b; int kind, pass_index; };
    std::vector<TimedLaunch> timed;
    cudaEvent_t t0 = nullptr, t1 = nullptr;   // qsv_timer_*
    cudaEvent_t swap_t0 = nullptr, swap_t1 = nullptr;
    bool use_peer_swap = true;        // swap through mapped peer memory when qsv_comm_set_peers succeeded
    // comm (NCCL) — opaque here, owned by exchange.cuh
    void *comm = nullptr;
    std::string err;
};
