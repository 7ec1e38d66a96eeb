// xchg.cuh — the exchange step of a global<->local qubit swap as ONE persistent TMA kernel per chunk,
// sized to a handful of SMs so that it runs BESIDE the pass kernels of the neighbouring chunks.
//
// Design source: HiSVSIM's bit redistribution (hisvsim_repo/mpi_redistributer.hpp:100-164: who sends
// which slots to whom) and the reference's reader / worker / writer overlap
// (wenbo_engine/runner/pipeline.py:50-82): here the "reader/writer" is this kernel, the "worker" the
// pass kernels of the chunks before and after it.
//
// What one launch does (chunk j of a swap of s rank bits with s local bits, group of P = 2^s ranks):
//   for every peer d != me:   my elements {swap bits = d, chunk bits = j}  <->  d's {swap bits = me, chunk bits = j}
// in place: of every pair the rank with the smaller group index owns the first half of the elements, the
// other rank the second half, so each element is moved by exactly one GPU.  A unit of work is 16 KB of
// the pair (one or more contiguous runs; a run is 2^lowest special position elements): one elected
// thread per CTA pulls the local and the remote unit into shared memory with cp.async.bulk (TMA, no
// registers, no L1), then pushes them back crosswise with cp.async.bulk shared -> global; a ring of NS
// stages keeps (NS-1) * 16 KB of remote reads in flight per SM.  Measured on 2 B200 (tools/xchg_bench.cu,
// profiles/r02/xchg_bench_2gpu.jsonl): throughput is in-flight bound, 43 GB/s per SM, and 16 SMs reach
// the same 686 GB/s per direction that the load/store kernel needs all 148 SMs for.
//
// Cross-GPU ordering without NCCL: two monotone 64-bit flags per source rank in the (peer-mapped)
// tail of every shard allocation.
//   kind 0  "the passes before this chunk's exchange are finished on rank r"  (written at kernel start:
//           the kernel is stream-ordered after those passes)
//   kind 1  "rank r has finished moving its halves of chunk j"
// A launch first signals kind 0 to its peers and waits for theirs; its last CTA signals kind 1 and
// waits for the peers' kind 1 before it exits, so completion of the kernel on rank r means: chunk j of
// r's shard is final.  Values are a per-handle sequence number (every rank executes the same swaps).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace qsvx {

constexpr int kXchgStages = 6;
constexpr int kXchgThreads = 128;
constexpr int kFlagSlots = 8;                 // world <= 8 on one box
constexpr size_t kTailBytes = 4096;           // flags + counters at the end of every shard allocation
// tail layout (uint64 words): [0..7] kind 0 by source rank, [8..15] kind 1 by source rank, [16] CTA counter
constexpr int kTailCounterWord = 16;

struct XchgArgs {
    char *mine;                               // this rank's shard
    char *peer[8];                            // shard of the rank with group index d (entry `me` unused)
    unsigned long long *my_flags;             // tail of this rank's allocation
    unsigned long long *peer_flags[8];        // tails of the peers, by group index
    int rank_of[8];                           // group index -> global rank (flag slot)
    int my_rank;
    int n_peers;                              // P = 2^s
    int me;                                   // this rank's group index
    int n_special;                            // s + c special index bits (swapped + chunk), ascending in pos[]
    int pos[8];
    int swap_pos[3];                          // position of the local bit exchanged with swapped rank bit i
    unsigned long long chunk_val;             // the chunk bits of this launch, already at their positions
    unsigned elem_log2;                       // log2(bytes per amplitude)
    unsigned run_log2;                        // log2(bytes of one contiguous run) <= stage_log2
    unsigned stage_log2;                      // log2(bytes of one unit of work = one ring stage per direction)
    unsigned long long units_per_half;        // units in one half of a pair
    unsigned long long half_elems;            // elements in one half of a pair (compacted index space)
    unsigned long long seq;                   // flag value of this launch
};

__device__ __forceinline__ uint32_t xs_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned long long xs_insert_zero(unsigned long long x, int pos) {
    const unsigned long long low = x & ((1ull << pos) - 1ull);
    return ((x >> pos) << (pos + 1)) | low;
}
__device__ __forceinline__ void xs_signal(unsigned long long *flag, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(flag), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long xs_peek(const unsigned long long *flag) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flag) : "memory");
    return v;
}
__device__ __forceinline__ void xs_wait(const unsigned long long *flag, unsigned long long v) {
    while (xs_peek(flag) < v) __nanosleep(64);
}

// byte offset inside a shard of compacted element offset e with the swapped bits set to `blk`
__device__ __forceinline__ unsigned long long xs_addr(const XchgArgs &A, unsigned long long e, int blk) {
#pragma unroll
    for (int i = 0; i < 8; ++i) if (i < A.n_special) e = xs_insert_zero(e, A.pos[i]);
    e |= A.chunk_val;
#pragma unroll
    for (int i = 0; i < 3; ++i) if ((1 << i) < A.n_peers) e |= (unsigned long long)((blk >> i) & 1) << A.swap_pos[i];
    return e << A.elem_log2;
}

__global__ void __launch_bounds__(kXchgThreads, 1) k_xchg_tma(const __grid_constant__ XchgArgs A) {
    extern __shared__ __align__(128) unsigned char xs_smem[];
    const unsigned R = 1u << A.stage_log2;                  // bytes per unit and direction
    const unsigned run = 1u << A.run_log2;                  // contiguous bytes per bulk copy
    const unsigned pieces = R >> A.run_log2;
    unsigned long long *bar = reinterpret_cast<unsigned long long *>(xs_smem + (size_t)kXchgStages * 2 * R);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < kXchgStages; ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xs_u32(&bar[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- barrier in: my passes of this chunk are done (stream order); wait for the peers' passes
    if (tid < A.n_peers - 1) {
        const int d = A.me ^ (tid + 1);
        if (blockIdx.x == 0) xs_signal(A.peer_flags[d] + A.my_rank, A.seq);
        xs_wait(A.my_flags + A.rank_of[d], A.seq);
    }
    __syncthreads();
    if (tid != 0) return;

    const unsigned long long total = (unsigned long long)(A.n_peers - 1) * A.units_per_half;
    const unsigned long long grid = gridDim.x, bid = blockIdx.x;
    const unsigned long long n_my = total > bid ? (total - bid + grid - 1) / grid : 0;
    const unsigned unit_elems = R >> A.elem_log2, run_elems = run >> A.elem_log2;

    auto issue_loads = [&](unsigned long long i) {
        const unsigned long long gi = bid + i * grid;
        const int k = (int)(gi % (unsigned long long)(A.n_peers - 1));
        const unsigned long long u = gi / (unsigned long long)(A.n_peers - 1);
        const int d = A.me ^ (k + 1);
        const unsigned long long e0 = (A.me < d ? 0ull : A.half_elems) + u * unit_elems;
        const int st = (int)(i % kXchgStages);
        unsigned char *sl = xs_smem + (size_t)st * 2 * R;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xs_u32(&bar[st])), "r"(2u * R) : "memory");
        for (unsigned q = 0; q < pieces; ++q) {            // remote first: longest latency
            const char *gr = A.peer[d] + xs_addr(A, e0 + (unsigned long long)q * run_elems, A.me);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(xs_u32(sl + R + q * run)), "l"(gr), "r"(run), "r"(xs_u32(&bar[st])) : "memory");
        }
        for (unsigned q = 0; q < pieces; ++q) {
            const char *gl = A.mine + xs_addr(A, e0 + (unsigned long long)q * run_elems, d);
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(xs_u32(sl + q * run)), "l"(gl), "r"(run), "r"(xs_u32(&bar[st])) : "memory");
        }
    };
    auto issue_stores = [&](unsigned long long i) {
        const unsigned long long gi = bid + i * grid;
        const int k = (int)(gi % (unsigned long long)(A.n_peers - 1));
        const unsigned long long u = gi / (unsigned long long)(A.n_peers - 1);
        const int d = A.me ^ (k + 1);
        const unsigned long long e0 = (A.me < d ? 0ull : A.half_elems) + u * unit_elems;
        const int st = (int)(i % kXchgStages);
        unsigned char *sl = xs_smem + (size_t)st * 2 * R;
        for (unsigned q = 0; q < pieces; ++q) {
            char *gr = A.peer[d] + xs_addr(A, e0 + (unsigned long long)q * run_elems, A.me);
            char *gl = A.mine + xs_addr(A, e0 + (unsigned long long)q * run_elems, d);
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gr), "r"(xs_u32(sl + q * run)), "r"(run) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gl), "r"(xs_u32(sl + R + q * run)), "r"(run) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    };

    for (unsigned long long i = 0; i < n_my && i < (unsigned long long)(kXchgStages - 1); ++i) issue_loads(i);
    for (unsigned long long k = 0; k < n_my; ++k) {
        const int st = (int)(k % kXchgStages);
        const uint32_t parity = (uint32_t)((k / kXchgStages) & 1ull);
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(xs_u32(&bar[st])), "r"(parity) : "memory");
        }
        issue_stores(k);
        const unsigned long long nxt = k + kXchgStages - 1;
        if (nxt < n_my) {
            // every store group but the newest has finished READING its stage: the stage of unit k-1 is free
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            issue_loads(nxt);
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // all my writes (local and remote) are complete
    asm volatile("fence.proxy.async;" ::: "memory");
    __threadfence_system();
    // ---- barrier out: the last CTA of this launch tells the peers and waits for them
    unsigned *ctr = reinterpret_cast<unsigned *>(A.my_flags + kTailCounterWord);
    const unsigned old = atomicInc(ctr, gridDim.x - 1);
    if (old == gridDim.x - 1) {
        __threadfence_system();
        for (int k = 0; k < A.n_peers - 1; ++k) xs_signal(A.peer_flags[A.me ^ (k + 1)] + kFlagSlots + A.my_rank, A.seq);
        for (int k = 0; k < A.n_peers - 1; ++k) xs_wait(A.my_flags + kFlagSlots + A.rank_of[A.me ^ (k + 1)], A.seq);
    }
}

// The same exchange with plain loads and stores (runs shorter than the TMA path wants, or QSV_XCHG=ldst):
// every thread moves 16-byte pieces; flags as above.
template <int U>
__global__ void __launch_bounds__(1024, 1) k_xchg_ldst(const __grid_constant__ XchgArgs A) {
    const int tid = threadIdx.x;
    if (tid < A.n_peers - 1) {
        const int d = A.me ^ (tid + 1);
        if (blockIdx.x == 0) xs_signal(A.peer_flags[d] + A.my_rank, A.seq);
        xs_wait(A.my_flags + A.rank_of[d], A.seq);
    }
    __syncthreads();
    const unsigned long long half16 = (A.half_elems << A.elem_log2) >> 4;          // 16-byte pieces in one half
    const unsigned long long total = (unsigned long long)(A.n_peers - 1) * half16;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned per16 = 4u - A.elem_log2;                                       // log2(elements per piece): 0 or 1
    for (unsigned long long p0 = (unsigned long long)blockIdx.x * blockDim.x + tid; p0 < total; p0 += stride * U) {
        int4 x[U], y[U];
        int4 *pl[U], *pr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const unsigned long long p = p0 + (unsigned long long)u * stride;
            const unsigned long long pp = p < total ? p : p0;
            const int k = (int)(pp / half16);
            const unsigned long long i = pp - (unsigned long long)k * half16;
            const int d = A.me ^ (k + 1);
            const unsigned long long e = (A.me < d ? 0ull : A.half_elems) + (i << per16);
            pl[u] = reinterpret_cast<int4 *>(A.mine + xs_addr(A, e, d));
            pr[u] = reinterpret_cast<int4 *>(A.peer[d] + xs_addr(A, e, A.me));
            y[u] = *pr[u];
            x[u] = *pl[u];
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (p0 + (unsigned long long)u * stride < total) { *pl[u] = y[u]; *pr[u] = x[u]; }
    }
    __threadfence_system();
    __syncthreads();
    if (tid != 0) return;
    unsigned *ctr = reinterpret_cast<unsigned *>(A.my_flags + kTailCounterWord);
    const unsigned old = atomicInc(ctr, gridDim.x - 1);
    if (old == gridDim.x - 1) {
        __threadfence_system();
        for (int k = 0; k < A.n_peers - 1; ++k) xs_signal(A.peer_flags[A.me ^ (k + 1)] + kFlagSlots + A.my_rank, A.seq);
        for (int k = 0; k < A.n_peers - 1; ++k) xs_wait(A.my_flags + kFlagSlots + A.rank_of[A.me ^ (k + 1)], A.seq);
    }
}

inline size_t xchg_smem_bytes(unsigned stage_log2) { return (size_t)kXchgStages * 2 * ((size_t)1 << stage_log2) + 8 * kXchgStages + 64; }

}  // namespace qsvx
