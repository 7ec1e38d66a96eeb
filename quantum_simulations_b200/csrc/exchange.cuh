// exchange.cuh — global<->local qubit swap across the shards of one box.
//
// Design source: HiSVSIM's bit redistribution (hisvsim_repo/mpi_redistributer.hpp:100-344,
// svsim-mpi.hpp:123-173): "make a different set of index bits the rank bits".  Here the set
// of local bits that leave is always the TOP n_swap local bits (the pass kernel relabels
// qubits inside a tile for free, so the planner parks the outgoing qubits there first);
// every peer block is then one contiguous slice of 2^(n_local - n_swap) amplitudes and the
// exchange is an all-to-all of equal contiguous blocks among groups of 2^n_swap ranks:
//      rank r, block d   <->   rank r' (= r with its swapped rank bits set to d), block me
// The block d == me stays where it is.
//
// Memory: shards of 128 GiB (n=36 on 8 GPUs, n=34 on 2) leave no room for a second copy, so
// the exchange is IN PLACE and chunked: chunk c of my block d goes out while the peer's chunk
// c of its block `me` lands in a small bounce buffer, which is then copied over the slot
// that was just sent.  NCCL (ncclSend/ncclRecv grouped) over NVLink 5 / NVSwitch carries the
// chunks; NCCL is dlopen()ed at qsv_comm_init so single-GPU use has no NCCL dependency.
#pragma once
#include <dlfcn.h>

#include "common.cuh"

namespace qsvx {

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
constexpr int ncclChar = 0;   // ncclInt8

struct Nccl {
    void *lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    std::string why;

    bool load() {
        if (lib) return true;
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { why = std::string("dlopen(libnccl.so.2): ") + dlerror(); return false; }
#define QSV_SYM(field, name) field = (decltype(field))dlsym(lib, name); if (!field) { why = std::string("missing symbol ") + name; return false; }
        QSV_SYM(GetUniqueId, "ncclGetUniqueId");
        QSV_SYM(CommInitRank, "ncclCommInitRank");
        QSV_SYM(CommDestroy, "ncclCommDestroy");
        QSV_SYM(Send, "ncclSend");
        QSV_SYM(Recv, "ncclRecv");
        QSV_SYM(GroupStart, "ncclGroupStart");
        QSV_SYM(GroupEnd, "ncclGroupEnd");
        QSV_SYM(AllReduce, "ncclAllReduce");
        QSV_SYM(GetErrorString, "ncclGetErrorString");
#undef QSV_SYM
        return true;
    }
};

inline Nccl &nccl() { static Nccl n; return n; }

struct Comm {
    ncclComm_t comm = nullptr;   // null: shards of ONE process (qsv_comm_init_local), ordering by the flags alone
    // peer-memory path: the other shards of the box mapped into this process (CUDA IPC)
    std::vector<void *> peer;    // peer[r] = rank r's shard, nullptr if not mapped
    std::vector<unsigned long long *> peer_tail;   // peer_tail[r] = flag words of rank r (xchg.cuh)
    bool peers_ready = false;
    bool ipc_mapped = false;     // peer[] entries were opened with cudaIpcOpenMemHandle
    unsigned long long seq = 0;  // sequence number of the exchange kernels (same on every rank)
    std::vector<cudaEvent_t> ev_pool;   // events of pipelined transitions, reused
    bool xchg_attr_set = false;
    void *bounce = nullptr;      // (peers) x chunk bytes, double buffered
    size_t bounce_bytes = 0;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_xfer[2] = {nullptr, nullptr};
    cudaEvent_t ev_copy[2] = {nullptr, nullptr};
};

}  // namespace qsvx

static inline void qsv_comm_teardown(qsv_handle *h) {
    auto *c = (qsvx::Comm *)h->comm;
    if (!c) return;
    if (c->ipc_mapped)
        for (size_t r = 0; r < c->peer.size(); ++r) if (c->peer[r] && (int)r != h->rank) cudaIpcCloseMemHandle(c->peer[r]);
    if (c->comm) qsvx::nccl().CommDestroy(c->comm);
    for (auto &e : c->ev_pool) cudaEventDestroy(e);
    if (c->bounce) cudaFree(c->bounce);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    for (auto &e : c->ev_xfer) if (e) cudaEventDestroy(e);
    for (auto &e : c->ev_copy) if (e) cudaEventDestroy(e);
    delete c;
    h->comm = nullptr;
}

#define QSVX_FAIL(h, code, ...)                      \
    do {                                             \
        char buf_[512];                              \
        snprintf(buf_, sizeof(buf_), __VA_ARGS__);   \
        (h)->err = buf_;                             \
        return (code);                               \
    } while (0)

#define QSVX_NCCL(h, expr)                                                                  \
    do {                                                                                    \
        int r_ = (expr);                                                                    \
        if (r_ != 0) QSVX_FAIL(h, QSV_ECOMM, "%s failed: %s", #expr, qsvx::nccl().GetErrorString(r_)); \
    } while (0)

#define QSVX_CUDA(h, expr)                                                                  \
    do {                                                                                    \
        cudaError_t e_ = (expr);                                                            \
        if (e_ != cudaSuccess) QSVX_FAIL(h, QSV_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

static inline void qsvx_timer_begin(qsv_handle *h) {
    if (!h->timing) return;
    cudaEventCreate(&h->swap_t0); cudaEventCreate(&h->swap_t1);
    cudaEventRecord(h->swap_t0, h->stream);
}
static inline void qsvx_timer_end(qsv_handle *h, int kind) {
    if (!h->timing) return;
    cudaEventRecord(h->swap_t1, h->stream);
    h->timed.push_back({h->swap_t0, h->swap_t1, kind, -1});
}

extern "C" int qsv_comm_unique_id(void *id128) {
    if (!id128) return QSV_EINVAL;
    if (!qsvx::nccl().load()) return QSV_ECOMM;
    qsvx::ncclUniqueId id;
    if (qsvx::nccl().GetUniqueId(&id) != 0) return QSV_ECOMM;
    memcpy(id128, &id, 128);
    return QSV_OK;
}

extern "C" int qsv_comm_init(qsv_handle *h, const void *id128) {
    if (!h || !id128) return QSV_EINVAL;
    if (!qsvx::nccl().load()) QSVX_FAIL(h, QSV_ECOMM, "NCCL unavailable: %s", qsvx::nccl().why.c_str());
    if (h->comm) qsv_comm_teardown(h);
    QSVX_CUDA(h, cudaSetDevice(h->device));
    auto *c = new qsvx::Comm();
    qsvx::ncclUniqueId id;
    memcpy(&id, id128, 128);
    int r = qsvx::nccl().CommInitRank(&c->comm, h->world, id, h->rank);
    if (r != 0) { delete c; QSVX_FAIL(h, QSV_ECOMM, "ncclCommInitRank: %s", qsvx::nccl().GetErrorString(r)); }
    cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (auto &e : c->ev_xfer) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto &e : c->ev_copy) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    h->comm = c;
    return QSV_OK;
}

// ---- peer-memory exchange -------------------------------------------------------------
// The block pair (rank a: block b) <-> (rank b: block a) is swapped IN PLACE by plain loads and
// stores through NVLink: rank a owns the first half of the pair's elements, rank b the second
// half (a < b in swapped-bit order), so every element is touched by exactly one thread of one
// GPU: no bounce buffer, no extra HBM pass, both NVLink directions carry one block each.
struct PeerTable { void *ptr[8]; };
// the swapped LOCAL bit positions: sorted[] ascending (for the zero-bit insertion), bit[i] = position
// of the local bit that is exchanged with the i-th swapped rank bit
struct SwapBits { int n; int sorted[3]; int bit[3]; };

__device__ __forceinline__ uint64_t swap_addr(const SwapBits &sb, uint64_t off, int blk) {
    uint64_t a = off;
#pragma unroll
    for (int i = 0; i < 3; ++i) if (i < sb.n) a = insert_zero_bit(a, sb.sorted[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) if (i < sb.n) a |= (uint64_t)((blk >> i) & 1) << sb.bit[i];
    return a;
}

template <typename V, int U>
__global__ void __launch_bounds__(512)
k_swap_peer(V *__restrict__ mine, PeerTable peers, const int n_peers, const int me, const uint64_t block_amps,
            const SwapBits sb, const int only_phase) {
    const uint64_t half = block_amps >> 1;
    const uint64_t total = (uint64_t)(only_phase >= 0 ? 1 : n_peers - 1) * half;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t e0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e0 < total; e0 += stride * U) {
        V x[U], y[U];
        V *pl[U], *pr[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint64_t e = e0 + (uint64_t)u * stride;
            const uint64_t ee = e < total ? e : e0;
            const int k0 = (int)(ee / half);                // phase 0 .. n_peers-2
            const uint64_t i = ee - (uint64_t)k0 * half;
            const int k = only_phase >= 0 ? only_phase : k0;
            const int d = me ^ (k + 1);                     // phase k pairs me with me^(k+1): a perfect
                                                            // matching of the group, so no GPU is the
                                                            // target of two others at the same time
            const uint64_t off = (me < d ? 0 : half) + i;   // my half of the pair
            pl[u] = mine + swap_addr(sb, off, d);                     // my "block d" (swapped local bits = d)
            pr[u] = (V *)peers.ptr[d] + swap_addr(sb, off, me);       // peer d's "block me"
            y[u] = *pr[u];
            x[u] = *pl[u];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (e0 + (uint64_t)u * stride < total) { *pl[u] = y[u]; *pr[u] = x[u]; }
        }
    }
}

extern "C" int qsv_comm_ipc_handle(qsv_handle *h, void *out64) {
    if (!h || !out64) return QSV_EINVAL;
    QSVX_CUDA(h, cudaSetDevice(h->device));
    cudaIpcMemHandle_t mh;
    static_assert(sizeof(mh) == 64, "cudaIpcMemHandle_t is 64 bytes");
    QSVX_CUDA(h, cudaIpcGetMemHandle(&mh, h->d_state));
    memcpy(out64, &mh, 64);
    return QSV_OK;
}

// handles: world x 64 bytes (entry r = qsv_comm_ipc_handle of rank r).  Maps every other shard.
extern "C" int qsv_comm_set_peers(qsv_handle *h, const void *handles) {
    if (!h || !handles) return QSV_EINVAL;
    auto *c = (qsvx::Comm *)h->comm;
    if (!c) QSVX_FAIL(h, QSV_ECOMM, "set_peers: qsv_comm_init was not called");
    if (h->world > 8) QSVX_FAIL(h, QSV_EINVAL, "set_peers: the peer path serves one box (world <= 8)");
    QSVX_CUDA(h, cudaSetDevice(h->device));
    c->peer.assign(h->world, nullptr);
    c->peers_ready = false;
    for (int r = 0; r < h->world; ++r) {
        if (r == h->rank) { c->peer[r] = h->d_state; continue; }
        cudaIpcMemHandle_t mh;
        memcpy(&mh, (const char *)handles + 64 * r, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&c->peer[r], mh, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            c->peer[r] = nullptr;
            QSVX_FAIL(h, QSV_ECOMM, "cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        }
    }
    c->peer_tail.assign(h->world, nullptr);
    for (int r = 0; r < h->world; ++r)
        c->peer_tail[r] = r == h->rank ? h->d_tail : (unsigned long long *)((char *)c->peer[r] + h->n_amps * h->amp_bytes);
    c->peers_ready = true;
    c->ipc_mapped = true;
    return QSV_OK;
}

// Shards of ONE process (several handles on the devices this process can address directly — tests on a
// single GPU, or one process driving a whole box): no NCCL communicator, no CUDA IPC; the exchange
// kernels order themselves through the flag words of xchg.cuh.
extern "C" int qsv_comm_init_local(qsv_handle *h) {
    if (!h) return QSV_EINVAL;
    if (h->comm) qsv_comm_teardown(h);
    QSVX_CUDA(h, cudaSetDevice(h->device));
    auto *c = new qsvx::Comm();
    cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (auto &e : c->ev_xfer) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    for (auto &e : c->ev_copy) cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    h->comm = c;
    return QSV_OK;
}

// shards[r] = device pointer of rank r's shard (qsv_device_ptr of that handle), valid in this process
extern "C" int qsv_comm_set_peers_local(qsv_handle *h, void *const *shards) {
    if (!h || !shards) return QSV_EINVAL;
    auto *c = (qsvx::Comm *)h->comm;
    if (!c) QSVX_FAIL(h, QSV_ECOMM, "set_peers_local: qsv_comm_init_local was not called");
    if (h->world > 8) QSVX_FAIL(h, QSV_EINVAL, "set_peers_local: world <= 8");
    if (shards[h->rank] != h->d_state) QSVX_FAIL(h, QSV_EINVAL, "set_peers_local: entry %d must be this handle's shard", h->rank);
    c->peer.assign(shards, shards + h->world);
    c->peer_tail.assign(h->world, nullptr);
    for (int r = 0; r < h->world; ++r) {
        if (!shards[r]) QSVX_FAIL(h, QSV_EINVAL, "set_peers_local: null shard of rank %d", r);
        c->peer_tail[r] = (unsigned long long *)((char *)shards[r] + h->n_amps * h->amp_bytes);
    }
    c->peers_ready = true;
    c->ipc_mapped = false;
    return QSV_OK;
}

// ---- exchange kernels of xchg.cuh: argument block of chunk `chunk_j` of a swap -------------------
// chunk_bits: c local positions (none for a whole-shard exchange), disjoint from local_bits
static int qsvx_xchg_args(qsv_handle *h, qsvx::Comm *c, int n_swap, const int *global_bits, const int *local_bits,
                          int n_chunk, const int *chunk_bits, unsigned chunk_j, qsvx::XchgArgs &A, bool &tma_ok) {
    memset(&A, 0, sizeof(A));
    const int peers = 1 << n_swap;
    int me = 0;
    for (int i = 0; i < n_swap; ++i) me |= ((h->rank >> (global_bits[i] - h->n_local)) & 1) << i;
    A.mine = (char *)h->d_state;
    A.my_flags = h->d_tail;
    A.my_rank = h->rank;
    A.n_peers = peers;
    A.me = me;
    for (int d = 0; d < peers; ++d) {
        int r = h->rank;
        for (int i = 0; i < n_swap; ++i) {
            const int rb = global_bits[i] - h->n_local;
            r = (r & ~(1 << rb)) | (((d >> i) & 1) << rb);
        }
        A.peer[d] = (char *)c->peer[r];
        A.peer_flags[d] = c->peer_tail[r];
        A.rank_of[d] = r;
        if (!A.peer[d] || !A.peer_flags[d]) QSVX_FAIL(h, QSV_ECOMM, "exchange: rank %d is not mapped", r);
    }
    int special[8], ns = 0;
    for (int i = 0; i < n_swap; ++i) { special[ns++] = local_bits[i]; A.swap_pos[i] = local_bits[i]; }
    for (int i = 0; i < n_chunk; ++i) {
        special[ns++] = chunk_bits[i];
        A.chunk_val |= (unsigned long long)((chunk_j >> i) & 1u) << chunk_bits[i];
    }
    std::sort(special, special + ns);
    for (int i = 0; i + 1 < ns; ++i) if (special[i] == special[i + 1]) QSVX_FAIL(h, QSV_EINVAL, "exchange: repeated index bit %d", special[i]);
    A.n_special = ns;
    for (int i = 0; i < ns; ++i) A.pos[i] = special[i];
    A.elem_log2 = h->amp_bytes == 16 ? 4u : 3u;
    const int free_bits = h->n_local - ns;                 // compacted index bits of one pair
    if (free_bits < 1) QSVX_FAIL(h, QSV_EINVAL, "exchange: shard too small for %d special bits", ns);
    A.half_elems = 1ull << (free_bits - 1);
    const unsigned half_log2 = (unsigned)(free_bits - 1) + A.elem_log2;          // log2(bytes of one half)
    A.stage_log2 = std::min(14u, half_log2);
    A.run_log2 = std::min(A.stage_log2, (unsigned)special[0] + A.elem_log2);
    A.units_per_half = 1ull << (half_log2 - A.stage_log2);
    A.seq = c->seq;
    {   // seconds a launch waits for a peer before it gives up (QSV_XCHG_TIMEOUT_S; shards of one process: short)
        const char *e = getenv("QSV_XCHG_TIMEOUT_S");
        const double sec = e ? atof(e) : (c->comm ? 60.0 : 10.0);
        A.timeout_ns = (unsigned long long)((sec > 0.01 ? sec : 0.01) * 1e9);
    }
    A.stage_log2 = std::min(12u, half_log2);                          // 4 KB units (small units complete sooner)
    if (const char *e = getenv("QSV_XCHG_STAGE_LOG2")) {             // experiment
        const unsigned v = (unsigned)atoi(e);
        if (v >= 10 && v <= 15 && v <= half_log2) A.stage_log2 = v;
    }
    A.run_log2 = std::min(A.stage_log2, (unsigned)special[0] + A.elem_log2);
    A.units_per_half = 1ull << (half_log2 - A.stage_log2);
    qsvx::xchg_ring_shape(A.stage_log2, A.n_warps, A.n_remote, A.n_local);
    if (const char *e = getenv("QSV_XCHG_RING")) {                   // "<issuing warps>,<remote slots>,<local slots>"
        unsigned w = 0, r = 0, l = 0;
        if (sscanf(e, "%u,%u,%u", &w, &r, &l) == 3 && w >= 1 && w <= 4 && r >= 2 && l >= 2 && r >= l &&
            qsvx::xchg_smem_bytes(A.stage_log2, w, r, l) <= qsvx::kXchgMaxSmem) { A.n_warps = w; A.n_remote = r; A.n_local = l; }
    }
    tma_ok = A.run_log2 >= 4 && A.stage_log2 >= 4 && (A.stage_log2 - A.run_log2) <= 6;   // <= 64 bulk copies per unit
    if (A.run_log2 < 4) QSVX_FAIL(h, QSV_EINVAL, "exchange: runs of %u bytes (lowest swapped / chunk position %d) are below the 16-byte pieces the kernels move",
                                  1u << A.run_log2, special[0]);
    return QSV_OK;
}

// The sticky error word of this rank's tail: non-zero = an exchange kernel gave up waiting for a peer.
static int qsvx_check_peers_alive(qsv_handle *h) {
    auto *c = (qsvx::Comm *)h->comm;
    if (!c || !c->peers_ready || c->seq == 0) return QSV_OK;
    unsigned long long w = 0;
    QSVX_CUDA(h, cudaMemcpy(&w, h->d_tail + qsvx::kTailErrorWord, sizeof(w), cudaMemcpyDeviceToHost));
    if (w) QSVX_FAIL(h, QSV_ECOMM, "exchange %llu: a peer did not arrive within the time limit (QSV_XCHG_TIMEOUT_S); the state of this shard is undefined", w);
    return QSV_OK;
}

// One exchange launch on `stream` with `sms` CTAs (one per SM: the TMA kernel claims the shared memory of its SM)
static int qsvx_xchg_launch(qsv_handle *h, qsvx::Comm *c, cudaStream_t stream, const qsvx::XchgArgs &A, bool tma_ok, int sms) {
    static const bool force_ldst = [] { const char *e = getenv("QSV_XCHG"); return e && !strcmp(e, "ldst"); }();
    if (!c->xchg_attr_set) {
        QSVX_CUDA(h, cudaFuncSetAttribute(qsvx::k_xchg_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qsvx::kXchgMaxSmem));
        c->xchg_attr_set = true;
    }
    if (sms < 1) sms = 1;
    if (tma_ok && !force_ldst) {
        unsigned long long total = (unsigned long long)(A.n_peers - 1) * A.units_per_half;
        const unsigned grid = (unsigned)std::min<unsigned long long>((unsigned long long)sms, std::max<unsigned long long>(total, 1ull));
        // full-size stages claim the whole SM (no second CTA beside it); small test shards take what they need
        qsvx::k_xchg_tma<<<grid, qsvx::kXchgThreads, qsvx::xchg_smem_bytes(A.stage_log2, A.n_warps, A.n_remote, A.n_local), stream>>>(A);
    } else {
        qsvx::k_xchg_ldst<4><<<(unsigned)sms, 1024, 0, stream>>>(A);
    }
    QSVX_CUDA(h, cudaGetLastError());
    return QSV_OK;
}

static int swap_barrier(qsv_handle *h, qsvx::Comm *c) {
    // stream-ordered barrier across the box: a 1-element all-reduce
    double *d = h->d_partials + h->n_partials - 2;
    QSVX_NCCL(h, qsvx::nccl().AllReduce(d, d, 1, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->comm, h->stream));
    return QSV_OK;
}

static int qsvx_swap_barrier_on(qsv_handle *h, qsvx::Comm *c, cudaStream_t stream) {
    double *d = h->d_partials + h->n_partials - 3;
    QSVX_NCCL(h, qsvx::nccl().AllReduce(d, d, 1, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->comm, stream));
    return QSV_OK;
}

// One phase of the peer-memory exchange on `stream` (used by qsv_pass_swap_overlapped): a barrier
// (every rank has finished the blocks of this phase), then the pair (me, me ^ (phase+1)) only.
static int qsvx_swap_phase(qsv_handle *h, qsvx::Comm *c, cudaStream_t stream, int n_swap, const int *global_bits,
                           const int *local_bits, int phase) {
    const int peers = 1 << n_swap;
    int me = 0;
    for (int i = 0; i < n_swap; ++i) me |= ((h->rank >> (global_bits[i] - h->n_local)) & 1) << i;
    PeerTable pt;
    for (int d = 0; d < 8; ++d) {
        pt.ptr[d] = nullptr;
        if (d < peers) {
            int r = h->rank;
            for (int i = 0; i < n_swap; ++i) {
                const int rb = global_bits[i] - h->n_local;
                r = (r & ~(1 << rb)) | (((d >> i) & 1) << rb);
            }
            pt.ptr[d] = c->peer[r];
        }
    }
    SwapBits sbits;
    sbits.n = n_swap;
    for (int i = 0; i < 3; ++i) { sbits.bit[i] = i < n_swap ? local_bits[i] : 0; sbits.sorted[i] = sbits.bit[i]; }
    std::sort(sbits.sorted, sbits.sorted + n_swap);
    int rc = qsvx_swap_barrier_on(h, c, stream);
    if (rc) return rc;
    const uint64_t block_amps = h->n_amps >> n_swap;
    const int grid = h->sm_count * 2;                       // leave SMs to the pass running beside it
    if (h->dtype == QSV_C128) k_swap_peer<double2, 4><<<grid, 512, 0, stream>>>((double2 *)h->d_state, pt, peers, me, block_amps, sbits, phase);
    else k_swap_peer<float2, 8><<<grid, 512, 0, stream>>>((float2 *)h->d_state, pt, peers, me, block_amps, sbits, phase);
    QSVX_CUDA(h, cudaGetLastError());
    return QSV_OK;
}

// Swap rank bits global_bits[i] (physical positions >= n_local) with local bits local_bits[i].
// Peer-memory path: any distinct local positions.  NCCL path: the TOP n_swap local bits in order
// (contiguous blocks), in place, chunked.  Collective over the group of 2^n_swap ranks.
extern "C" int qsv_swap_global_local(qsv_handle *h, int n_swap, const int *global_bits, const int *local_bits) {
    if (!h) return QSV_EINVAL;
    if (n_swap == 0) return QSV_OK;
    int g = h->n_qubits - h->n_local;
    if (n_swap < 0 || n_swap > g || !global_bits || !local_bits) QSVX_FAIL(h, QSV_EINVAL, "swap: n_swap=%d with %d rank bits", n_swap, g);
    for (int i = 0; i < n_swap; ++i) {
        if (global_bits[i] < h->n_local || global_bits[i] >= h->n_qubits) QSVX_FAIL(h, QSV_EINVAL, "swap: global bit %d is not a rank bit", global_bits[i]);
        if (local_bits[i] < 0 || local_bits[i] >= h->n_local) QSVX_FAIL(h, QSV_EINVAL, "swap: local bit %d outside the shard", local_bits[i]);
        for (int j = 0; j < i; ++j)
            if (global_bits[i] == global_bits[j] || local_bits[i] == local_bits[j]) QSVX_FAIL(h, QSV_EINVAL, "swap: repeated bit");
    }
    auto *c = (qsvx::Comm *)h->comm;
    if (!c) QSVX_FAIL(h, QSV_ECOMM, "swap: qsv_comm_init was not called");
    QSVX_CUDA(h, cudaSetDevice(h->device));
    auto &N = qsvx::nccl();
    bool top = true;                               // contiguous blocks: the top n_swap local bits in order
    for (int i = 0; i < n_swap; ++i) top = top && local_bits[i] == h->n_local - n_swap + i;
    if (n_swap > 3) QSVX_FAIL(h, QSV_EINVAL, "swap: at most 3 bits per exchange (one box)");

    const int peers = 1 << n_swap;
    int me = 0;                                    // my value of the swapped rank bits
    for (int i = 0; i < n_swap; ++i) me |= ((h->rank >> (global_bits[i] - h->n_local)) & 1) << i;
    auto peer_rank = [&](int d) {
        int r = h->rank;
        for (int i = 0; i < n_swap; ++i) {
            const int rb = global_bits[i] - h->n_local;
            r = (r & ~(1 << rb)) | (((d >> i) & 1) << rb);
        }
        return r;
    };
    const size_t block_bytes = (h->n_amps >> n_swap) * h->amp_bytes;
    static const bool xchg_default = [] { const char *e = getenv("QSV_SWAP_KERNEL"); return !(e && !strcmp(e, "pairs")); }();
    if (c->peers_ready && h->use_peer_swap && (!c->comm || xchg_default)) {
        // the whole exchange as ONE launch of the TMA kernel of xchg.cuh (its flag words are the barriers on
        // both sides: no NCCL on this path); QSV_SWAP_KERNEL=pairs keeps the load/store kernel below
        qsvx_timer_begin(h);
        qsvx::XchgArgs A;
        bool tma_ok = false;
        ++c->seq;
        int rc = qsvx_xchg_args(h, c, n_swap, global_bits, local_bits, 0, nullptr, 0u, A, tma_ok);
        if (rc) return rc;
        // one process per GPU: 32 SMs saturate NVLink twice over.  Shards of ONE process share a device, and
        // a launch spins until its peers' launches are resident: all of them together must fit on it
        const char *es = getenv("QSV_SWAP_SMS");
        const int sms = es ? atoi(es) : (c->comm ? 32 : std::max(1, h->sm_count / (2 * h->world)));
        rc = qsvx_xchg_launch(h, c, h->stream, A, tma_ok, sms);
        if (rc) return rc;
        qsvx_timer_end(h, 20 + n_swap);
        return QSV_OK;
    }
    if (c->peers_ready && h->use_peer_swap) {
        qsvx_timer_begin(h);
        PeerTable pt;
        for (int d = 0; d < 8; ++d) pt.ptr[d] = d < peers ? c->peer[peer_rank(d)] : nullptr;
        int rc = swap_barrier(h, c);                 // every peer has finished its previous pass
        if (rc) return rc;
        const uint64_t block_amps = h->n_amps >> n_swap;
        const char *eg = getenv("QSV_SWAP_GRID"), *eu = getenv("QSV_SWAP_UNROLL");
        const int grid = h->sm_count * (eg ? atoi(eg) : 4);
        const int unroll = eu ? atoi(eu) : 4;
        SwapBits sbits;
        sbits.n = n_swap;
        for (int i = 0; i < 3; ++i) { sbits.bit[i] = i < n_swap ? local_bits[i] : 0; sbits.sorted[i] = sbits.bit[i]; }
        std::sort(sbits.sorted, sbits.sorted + n_swap);
        if (h->dtype == QSV_C128) {
            if (unroll == 2) k_swap_peer<double2, 2><<<grid, 512, 0, h->stream>>>((double2 *)h->d_state, pt, peers, me, block_amps, sbits, -1);
            else if (unroll == 8) k_swap_peer<double2, 8><<<grid, 512, 0, h->stream>>>((double2 *)h->d_state, pt, peers, me, block_amps, sbits, -1);
            else k_swap_peer<double2, 4><<<grid, 512, 0, h->stream>>>((double2 *)h->d_state, pt, peers, me, block_amps, sbits, -1);
        } else k_swap_peer<float2, 8><<<grid, 512, 0, h->stream>>>((float2 *)h->d_state, pt, peers, me, block_amps, sbits, -1);
        QSVX_CUDA(h, cudaGetLastError());
        rc = swap_barrier(h, c);                     // nobody reads its shard before all exchanges landed
        if (rc) return rc;
        qsvx_timer_end(h, 20 + n_swap);
        return QSV_OK;
    }
    if (!top) QSVX_FAIL(h, QSV_EINVAL, "swap: the ncclSend/ncclRecv path needs the top %d local positions in order "
                                       "(arbitrary positions are served by the peer-memory kernel)", n_swap);
    size_t chunk = std::min(block_bytes, (size_t)256 << 20);
    const size_t need = 2 * (size_t)(peers - 1) * chunk;
    if (c->bounce_bytes < need) {
        if (c->bounce) cudaFree(c->bounce);
        c->bounce = nullptr; c->bounce_bytes = 0;
        QSVX_CUDA(h, cudaMalloc(&c->bounce, need));
        c->bounce_bytes = need;
    }
    char *state = (char *)h->d_state;
    const size_t n_chunks = (block_bytes + chunk - 1) / chunk;
    qsvx_timer_begin(h);
    // main stream: NCCL group of chunk k; copy stream: bounce -> vacated slots of chunk k.
    // Half (k&1) of the bounce buffer is reused by chunk k+2, which waits for copy k.
    for (size_t k = 0; k < n_chunks; ++k) {
        const size_t off = k * chunk;
        const size_t len = std::min(chunk, block_bytes - off);
        char *bb = (char *)c->bounce + (k & 1) * (size_t)(peers - 1) * chunk;
        if (k >= 2) QSVX_CUDA(h, cudaStreamWaitEvent(h->stream, c->ev_copy[k & 1], 0));
        QSVX_NCCL(h, N.GroupStart());
        int slot = 0;
        for (int d = 0; d < peers; ++d) {
            if (d == me) continue;
            const int pr = peer_rank(d);
            QSVX_NCCL(h, N.Send(state + (size_t)d * block_bytes + off, len, qsvx::ncclChar, pr, c->comm, h->stream));
            QSVX_NCCL(h, N.Recv(bb + (size_t)slot * chunk, len, qsvx::ncclChar, pr, c->comm, h->stream));
            ++slot;
        }
        QSVX_NCCL(h, N.GroupEnd());
        QSVX_CUDA(h, cudaEventRecord(c->ev_xfer[k & 1], h->stream));
        QSVX_CUDA(h, cudaStreamWaitEvent(c->copy_stream, c->ev_xfer[k & 1], 0));
        slot = 0;
        for (int d = 0; d < peers; ++d) {
            if (d == me) continue;
            QSVX_CUDA(h, cudaMemcpyAsync(state + (size_t)d * block_bytes + off, bb + (size_t)slot * chunk, len, cudaMemcpyDeviceToDevice, c->copy_stream));
            ++slot;
        }
        QSVX_CUDA(h, cudaEventRecord(c->ev_copy[k & 1], c->copy_stream));
    }
    QSVX_CUDA(h, cudaStreamWaitEvent(h->stream, c->ev_copy[0], 0));
    if (n_chunks > 1) QSVX_CUDA(h, cudaStreamWaitEvent(h->stream, c->ev_copy[1], 0));
    qsvx_timer_end(h, 20 + n_swap);                           // kind 20+s: swap of s bits
    return QSV_OK;
}

// All-reduce (sum) of one double across the shards: norm / sampling plumbing.
extern "C" int qsv_allreduce_sum(qsv_handle *h, double *value) {
    if (!h || !value) return QSV_EINVAL;
    if (h->world == 1) return QSV_OK;
    auto *c = (qsvx::Comm *)h->comm;
    if (!c) QSVX_FAIL(h, QSV_ECOMM, "allreduce: qsv_comm_init was not called");
    QSVX_CUDA(h, cudaSetDevice(h->device));
    double *d = h->d_partials + h->n_partials - 1;
    QSVX_CUDA(h, cudaMemcpyAsync(d, value, sizeof(double), cudaMemcpyHostToDevice, h->stream));
    QSVX_NCCL(h, qsvx::nccl().AllReduce(d, d, 1, /*ncclDouble*/ 8, /*ncclSum*/ 0, c->comm, h->stream));
    QSVX_CUDA(h, cudaMemcpyAsync(value, d, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    QSVX_CUDA(h, cudaStreamSynchronize(h->stream));
    return QSV_OK;
}
