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
// registers, no L1), then pushes them back crosswise with cp.async.bulk shared -> global.  Two rings of
// 16 KB slots: a deep one for the remote halves (NVLink round trip) and a shallow one for the local
// halves.  Measured on 2 B200 (tools/xchg_bench.cu, profiles/r02/xchg_bench_2gpu.jsonl): throughput is
// in-flight bound (a 16 KB remote bulk read takes ~3.7 us alone, ~6 us beside the pass kernels); with
// 80 KB of remote reads in flight per SM 16 SMs reach the 686 GB/s per direction that the load/store
// kernel needs all 148 SMs for.
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

constexpr int kXchgThreads = 128;
constexpr int kFlagSlots = 8;                 // world <= 8 on one box
constexpr size_t kTailBytes = 4096;           // flags + counters at the end of every shard allocation
// tail layout (uint64 words): [0..7] kind 0 by source rank, [8..15] kind 1 by source rank, [16] CTA counter
constexpr int kTailCounterWord = 16;
constexpr int kTailErrorWord = 17;            // sticky: an exchange kernel gave up waiting for a peer (value = seq)

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
    unsigned stage_log2;                      // log2(bytes of one unit of work = one ring slot)
    unsigned n_warps, n_remote, n_local;      // TMA kernel: issuing threads per CTA, slots of each one's remote / local ring
    unsigned long long units_per_half;        // units in one half of a pair
    unsigned long long half_elems;            // elements in one half of a pair (compacted index space)
    unsigned long long seq;                   // flag value of this launch
    unsigned long long timeout_ns;            // a wait for a peer's flag gives up after this long (sticky error word)
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
__device__ __forceinline__ unsigned long long xs_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// Wait until *flag >= v.  Nothing guarantees that the peer's kernel is running (a crashed rank, shards of one
// process that the device could not co-schedule): after timeout_ns the wait gives up, records the launch in the
// sticky error word of this rank's tail and returns false — the host reports it at the next synchronisation.
__device__ __forceinline__ bool xs_wait(const XchgArgs &A, const unsigned long long *flag, unsigned long long v) {
    if (xs_peek(A.my_flags + kTailErrorWord) != 0ull) return false;
    const unsigned long long t0 = xs_now();
    unsigned spins = 0;
    while (xs_peek(flag) < v) {
        __nanosleep(64);
        if ((++spins & 1023u) == 0u && xs_now() - t0 > A.timeout_ns) {
            xs_signal(A.my_flags + kTailErrorWord, A.seq);
            return false;
        }
    }
    return true;
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

// W issuing threads per CTA (lane 0 of warps 0..W-1), each with its own interleaved share of the CTA's units
// and its own two rings of R-byte slots: NR slots for the REMOTE halves (NVLink round trip, ~3.3 us alone and
// more beside the pass kernels: the deep ring) and NL slots for the local halves (HBM latency only; loaded
// just in time).  What was measured on 2 B200 (profiles/r02/xchg_bench_2gpu*.jsonl): throughput is bound by
// the remote bytes in flight and small units complete sooner than large ones; the issue loop of one thread is
// instruction-latency bound (a version with 64-bit divisions in it ran at half the rate), hence several
// issuing threads, a division-free loop (cursors advance by precomputed steps) and the addresses of a unit
// computed once, at its remote load, and cached in shared memory for its local load and its stores.
struct XchgUnit { unsigned long long la, ra, e0; int d; int pad; };   // cached per remote slot

__global__ void __launch_bounds__(kXchgThreads, 1) k_xchg_tma(const __grid_constant__ XchgArgs A) {
    extern __shared__ __align__(128) unsigned char xs_smem[];
    const unsigned R = 1u << A.stage_log2;                  // bytes per unit and direction
    const unsigned run = 1u << A.run_log2;                  // contiguous bytes per bulk copy
    const unsigned pieces = R >> A.run_log2;
    const unsigned W = A.n_warps, NR = A.n_remote, NL = A.n_local;
    unsigned long long *bars = reinterpret_cast<unsigned long long *>(xs_smem + (size_t)W * (NR + NL) * R);
    XchgUnit *cache = reinterpret_cast<XchgUnit *>(bars + W * (NR + NL));
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (unsigned s = 0; s < W * (NR + NL); ++s)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(xs_u32(&bars[s])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // ---- barrier in: my passes of this chunk are done (stream order); wait for the peers' passes
    __shared__ int xs_ok;
    if (tid == 0) xs_ok = 1;
    __syncthreads();
    if (tid < A.n_peers - 1) {
        const int d = A.me ^ (tid + 1);
        if (blockIdx.x == 0) xs_signal(A.peer_flags[d] + A.my_rank, A.seq);
        if (!xs_wait(A, A.my_flags + A.rank_of[d], A.seq)) xs_ok = 0;
    }
    __syncthreads();
    if (!xs_ok) return;                                      // a peer never arrived: move nothing
    const unsigned warp = (unsigned)tid >> 5;
    if ((tid & 31) != 0 || warp >= W) return;

    unsigned char *rbuf = xs_smem + (size_t)warp * (NR + NL) * R, *lbuf = rbuf + (size_t)NR * R;
    unsigned long long *rbar = bars + warp * (NR + NL), *lbar = rbar + NR;
    XchgUnit *cch = cache + warp * NR;
    const unsigned P1 = (unsigned)A.n_peers - 1u;
    const unsigned long long total = (unsigned long long)P1 * A.units_per_half;
    const unsigned long long first = (unsigned long long)blockIdx.x + (unsigned long long)warp * gridDim.x;
    const unsigned long long step = (unsigned long long)W * gridDim.x;
    const unsigned long long n_my = total > first ? (total - first + step - 1) / step : 0;
    const unsigned unit_elems = R >> A.elem_log2, run_elems = run >> A.elem_log2;
    // remote-load cursor: unit index gi = first + i * step  ->  peer slot ck = gi % P1, unit in the pair cu = gi / P1
    unsigned ck = (unsigned)(first % P1);
    unsigned long long cu = first / P1;
    const unsigned dk = (unsigned)(step % P1);
    const unsigned long long du = step / P1;
    unsigned rs_load = 0;                                    // remote slot the next remote load goes into
    unsigned ls_load = 0, ls_unit = 0;                       // local slot of the next local load; remote slot of ITS unit
    unsigned long long issued_r = 0, issued_l = 0;

    auto load_remote = [&]() {
        XchgUnit u;
        u.d = A.me ^ (int)(ck + 1);
        u.e0 = (A.me < u.d ? 0ull : A.half_elems) + cu * unit_elems;
        u.la = (unsigned long long)A.mine + xs_addr(A, u.e0, u.d);
        u.ra = (unsigned long long)A.peer[u.d] + xs_addr(A, u.e0, A.me);
        u.pad = 0;
        cch[rs_load] = u;
        unsigned char *sl = rbuf + (size_t)rs_load * R;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xs_u32(&rbar[rs_load])), "r"(R) : "memory");
        if (pieces == 1) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(xs_u32(sl)), "l"(u.ra), "r"(R), "r"(xs_u32(&rbar[rs_load])) : "memory");
        } else {
            for (unsigned q = 0; q < pieces; ++q) {
                const char *gr = A.peer[u.d] + xs_addr(A, u.e0 + (unsigned long long)q * run_elems, A.me);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(xs_u32(sl + q * run)), "l"(gr), "r"(run), "r"(xs_u32(&rbar[rs_load])) : "memory");
            }
        }
        ck += dk; cu += du;
        if (ck >= P1) { ck -= P1; ++cu; }
        if (++rs_load == NR) rs_load = 0;
        ++issued_r;
    };
    auto load_local = [&]() {                                // the unit's remote load was issued earlier: its addresses are cached
        const XchgUnit u = cch[ls_unit];
        unsigned char *sl = lbuf + (size_t)ls_load * R;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xs_u32(&lbar[ls_load])), "r"(R) : "memory");
        if (pieces == 1) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                         ::"r"(xs_u32(sl)), "l"(u.la), "r"(R), "r"(xs_u32(&lbar[ls_load])) : "memory");
        } else {
            for (unsigned q = 0; q < pieces; ++q) {
                const char *gl = A.mine + xs_addr(A, u.e0 + (unsigned long long)q * run_elems, u.d);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(xs_u32(sl + q * run)), "l"(gl), "r"(run), "r"(xs_u32(&lbar[ls_load])) : "memory");
            }
        }
        if (++ls_load == NL) ls_load = 0;
        if (++ls_unit == NR) ls_unit = 0;
        ++issued_l;
    };
    auto wait_bar = [&](unsigned long long *bar, uint32_t parity) {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(xs_u32(bar)), "r"(parity) : "memory");
        }
    };

    while (issued_r < n_my && issued_r + 1 < NR) load_remote();
    while (issued_l < n_my && issued_l + 1 < NL) load_local();
    unsigned rs = 0, ls = 0, rpar = 0, lpar = 0;             // slots and phases of the unit whose stores are next
    for (unsigned long long k = 0; k < n_my; ++k) {
        wait_bar(&lbar[ls], lpar);
        wait_bar(&rbar[rs], rpar);
        const XchgUnit u = cch[rs];
        unsigned char *sr = rbuf + (size_t)rs * R, *sl = lbuf + (size_t)ls * R;
        if (pieces == 1) {                                   // crosswise: local half -> peer, remote half -> here
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(u.ra), "r"(xs_u32(sl)), "r"(R) : "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(u.la), "r"(xs_u32(sr)), "r"(R) : "memory");
        } else {
            for (unsigned q = 0; q < pieces; ++q) {
                char *gr = A.peer[u.d] + xs_addr(A, u.e0 + (unsigned long long)q * run_elems, A.me);
                char *gl = A.mine + xs_addr(A, u.e0 + (unsigned long long)q * run_elems, u.d);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gr), "r"(xs_u32(sl + q * run)), "r"(run) : "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gl), "r"(xs_u32(sr + q * run)), "r"(run) : "memory");
            }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (++rs == NR) { rs = 0; rpar ^= 1u; }
        if (++ls == NL) { ls = 0; lpar ^= 1u; }
        if (issued_r < n_my || issued_l < n_my) {
            // every store group but the newest has finished READING shared memory: the slots of unit k-1 (for
            // k = 0: the one slot of each ring that has not been used yet) are free — rs_load / ls_load point at them
            asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            if (issued_r < n_my) load_remote();
            if (issued_l < n_my) load_local();
        }
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // all my writes (local and remote) are complete
    asm volatile("fence.proxy.async;" ::: "memory");
    __threadfence_system();
    // ---- barrier out: the last issuing thread of this launch tells the peers and waits for them
    unsigned *ctr = reinterpret_cast<unsigned *>(A.my_flags + kTailCounterWord);
    const unsigned n_issuers = gridDim.x * W;
    const unsigned old = atomicInc(ctr, n_issuers - 1);
    if (old == n_issuers - 1) {
        __threadfence_system();
        for (int k = 0; k < A.n_peers - 1; ++k) xs_signal(A.peer_flags[A.me ^ (k + 1)] + kFlagSlots + A.my_rank, A.seq);
        for (int k = 0; k < A.n_peers - 1; ++k) xs_wait(A, A.my_flags + kFlagSlots + A.rank_of[A.me ^ (k + 1)], A.seq);
    }
}

// The same exchange with plain loads and stores (runs shorter than the TMA path wants, or QSV_XCHG=ldst):
// every thread moves 16-byte pieces; flags as above.
template <int U>
__global__ void __launch_bounds__(1024, 1) k_xchg_ldst(const __grid_constant__ XchgArgs A) {
    const int tid = threadIdx.x;
    __shared__ int xs_ok;
    if (tid == 0) xs_ok = 1;
    __syncthreads();
    if (tid < A.n_peers - 1) {
        const int d = A.me ^ (tid + 1);
        if (blockIdx.x == 0) xs_signal(A.peer_flags[d] + A.my_rank, A.seq);
        if (!xs_wait(A, A.my_flags + A.rank_of[d], A.seq)) xs_ok = 0;
    }
    __syncthreads();
    if (!xs_ok) return;
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
        for (int k = 0; k < A.n_peers - 1; ++k) xs_wait(A, A.my_flags + kFlagSlots + A.rank_of[A.me ^ (k + 1)], A.seq);
    }
}

inline size_t xchg_smem_bytes(unsigned stage_log2, unsigned n_warps, unsigned n_remote, unsigned n_local) {
    return (size_t)n_warps * ((size_t)(n_remote + n_local) * (((size_t)1 << stage_log2) + 8) + (size_t)n_remote * 32) + 64;
}
constexpr size_t kXchgMaxSmem = 227 * 1024 - 1024;     // dynamic part: the kernels also hold a few static words
// default shape: 4 issuing threads, the slots that fit split 10 : 4 between the remote and the local ring
// (4 KB slots: 4 x 14 x 4 KB = 224 KB, 144 KB of remote reads in flight per SM)
inline void xchg_ring_shape(unsigned stage_log2, unsigned &n_warps, unsigned &n_remote, unsigned &n_local) {
    n_warps = 4;
    unsigned slots = (unsigned)((kXchgMaxSmem - 4096) / ((size_t)n_warps << stage_log2));
    while (slots < 4 && n_warps > 1) { n_warps >>= 1; slots = (unsigned)((kXchgMaxSmem - 4096) / ((size_t)n_warps << stage_log2)); }
    if (slots > 14) slots = 14;
    if (slots < 4) slots = 4;
    n_local = slots >= 12 ? 4 : (slots >= 7 ? 3 : 2);
    n_remote = slots - n_local;
}

}  // namespace qsvx
