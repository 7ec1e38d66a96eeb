// qsv.cu — C ABI (include/qsv.h) over the sm_100a kernels.  No torch, no CPU path:
// every entry point either launches CUDA work on the handle's stream or fails with a code.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>

#include "common.cuh"
#include "gate_kernels.cuh"
#include "pass_kernel.cuh"
#include "pass_ring.cuh"
#include "xchg.cuh"
#include "exchange.cuh"
#include "jit.cuh"
#include "sample.cuh"

static thread_local std::string g_create_error;

// One cached shard allocation per device: cudaFree / cudaMalloc of a multi-GiB buffer costs up to
// ~150 ms (page unmapping), which would dominate back-to-back simulations of the same size.  A
// destroyed handle parks its buffer here; the next qsv_create of the same size on that device takes
// it, any other size frees it first (so the cache never adds to the peak).  qsv_release_cached()
// empties it.
namespace {
struct CachedShard { void *ptr = nullptr; size_t bytes = 0; };
std::mutex g_cache_mu;
CachedShard g_cache[64];

void *cache_take(int device, size_t bytes) {
    std::lock_guard<std::mutex> g(g_cache_mu);
    CachedShard &c = g_cache[device & 63];
    if (!c.ptr) return nullptr;
    void *p = c.ptr;
    const bool fit = c.bytes == bytes;
    c = CachedShard();
    if (fit) return p;
    cudaFree(p);
    return nullptr;
}
void cache_put(int device, void *ptr, size_t bytes) {
    std::lock_guard<std::mutex> g(g_cache_mu);
    CachedShard &c = g_cache[device & 63];
    if (c.ptr) cudaFree(c.ptr);
    c.ptr = ptr; c.bytes = bytes;
}
}  // namespace

#define QSV_FAIL(h, code, ...)                                  \
    do {                                                        \
        char buf_[512];                                         \
        snprintf(buf_, sizeof(buf_), __VA_ARGS__);              \
        if (h) (h)->err = buf_; else g_create_error = buf_;     \
        return (code);                                          \
    } while (0)

#define QSV_CUDA(h, expr)                                                              \
    do {                                                                               \
        cudaError_t e_ = (expr);                                                       \
        if (e_ != cudaSuccess)                                                         \
            QSV_FAIL(h, QSV_ECUDA, "%s failed: %s", #expr, cudaGetErrorString(e_));    \
    } while (0)

#define QSV_CHECK_H(h) do { if (!(h)) return QSV_EINVAL; } while (0)

// Small device buffers come from the stream-ordered pool (cudaMallocAsync): no device-wide
// synchronisation and no page unmapping on free — a plain cudaFree was measured to stall for up to
// 0.7 s after a 16 GiB device-to-host copy.
static inline cudaError_t dev_alloc(qsv_handle *h, void **ptr, size_t bytes) { return cudaMallocAsync(ptr, bytes, h->stream); }
static inline void dev_free(qsv_handle *h, void *ptr) { if (ptr) cudaFreeAsync(ptr, h->stream); }
static void keep_pool_memory(int device) {
    static bool done[64] = {};
    if (done[device & 63]) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
        unsigned long long thr = ~0ull;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done[device & 63] = true;
}

static inline int grid_for(uint64_t items, int threads, int max_blocks = 148 * 16) {
    uint64_t b = (items + threads - 1) / threads;
    if (b < 1) b = 1;
    if (b > (uint64_t)max_blocks) b = max_blocks;
    return (int)b;
}

namespace {
struct ScopedTimer {
    qsv_handle *h; int kind, pass_index; cudaEvent_t a = nullptr, b = nullptr;
    ScopedTimer(qsv_handle *h_, int kind_, int pi = -1) : h(h_), kind(kind_), pass_index(pi) {
        if (h->timing) {
            cudaEventCreate(&a); cudaEventCreate(&b);
            cudaEventRecord(a, h->stream);
        }
    }
    ~ScopedTimer() {
        if (h->timing) {
            cudaEventRecord(b, h->stream);
            h->timed.push_back({a, b, kind, pass_index});
        }
    }
};
template <typename R> Mat2<R> to_mat2(const double *U) {
    Mat2<R> m;
    for (int i = 0; i < 4; ++i) { m.m[i].x = (R)U[2 * i]; m.m[i].y = (R)U[2 * i + 1]; }
    return m;
}
template <typename R> Mat4<R> to_mat4(const double *U) {
    Mat4<R> m;
    for (int i = 0; i < 16; ++i) { m.m[i].x = (R)U[2 * i]; m.m[i].y = (R)U[2 * i + 1]; }
    return m;
}
}  // namespace

extern "C" {

int qsv_abi_version(void) { return QSV_ABI_VERSION; }

int qsv_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *qsv_last_error(const qsv_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int qsv_create(qsv_handle **out, int n_qubits, int dtype, int device, int rank, int world) {
    qsv_handle *none = nullptr;
    if (!out) return QSV_EINVAL;
    *out = nullptr;
    if (n_qubits < 1 || n_qubits > 62) QSV_FAIL(none, QSV_EINVAL, "n_qubits=%d out of range", n_qubits);
    if (dtype != QSV_C64 && dtype != QSV_C128) QSV_FAIL(none, QSV_EINVAL, "bad dtype %d", dtype);
    if (world < 1 || (world & (world - 1))) QSV_FAIL(none, QSV_EINVAL, "world=%d must be a power of two", world);
    if (rank < 0 || rank >= world) QSV_FAIL(none, QSV_EINVAL, "rank=%d outside world=%d", rank, world);
    int g = 0;
    while ((1 << g) < world) ++g;
    if (g >= n_qubits) QSV_FAIL(none, QSV_EINVAL, "world=%d too large for %d qubits", world, n_qubits);
    int ndev = qsv_device_count();
    if (ndev == 0) QSV_FAIL(none, QSV_ECUDA, "no CUDA device visible: libqsv has no CPU path");
    if (device < 0 || device >= ndev) QSV_FAIL(none, QSV_EINVAL, "device=%d not in [0,%d)", device, ndev);

    qsv_handle *h = new (std::nothrow) qsv_handle();
    if (!h) return QSV_ENOMEM;
    h->n_qubits = n_qubits; h->n_local = n_qubits - g; h->dtype = dtype; h->device = device;
    h->rank = rank; h->world = world;
    h->n_amps = (size_t)1 << h->n_local;
    h->amp_bytes = dtype == QSV_C128 ? 16 : 8;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) {
        const size_t bytes = h->n_amps * h->amp_bytes + qsvx::kTailBytes;
        h->d_state = cache_take(device, bytes);
        if (!h->d_state) e = cudaMalloc(&h->d_state, bytes);
        h->alloc_base = h->d_state;
        h->d_tail = (unsigned long long *)((char *)h->d_state + h->n_amps * h->amp_bytes);
        if (e == cudaSuccess) e = cudaMemsetAsync(h->d_tail, 0, qsvx::kTailBytes, h->stream);
    }
    if (e == cudaSuccess) keep_pool_memory(device);
    if (e == cudaSuccess) { h->n_partials = kNormBlocks + 8; e = dev_alloc(h, (void **)&h->d_partials, h->n_partials * sizeof(double)); }
    if (e == cudaSuccess) e = dev_alloc(h, (void **)&h->d_pass_scratch, sizeof(qsv_pass));
    if (e != cudaSuccess) {
        g_create_error = std::string("qsv_create: ") + cudaGetErrorString(e);
        int code = (e == cudaErrorMemoryAllocation) ? QSV_ENOMEM : QSV_ECUDA;
        cudaGetLastError();
        if (h->d_state) cudaFree(h->d_state);
        if (h->stream) { dev_free(h, h->d_partials); dev_free(h, h->d_pass_scratch); }
        if (h->stream) cudaStreamDestroy(h->stream);
        delete h;
        return code;
    }
    *out = h;
    return QSV_OK;
}

int qsv_destroy(qsv_handle *h) {
    QSV_CHECK_H(h);
    const bool trace = getenv("QSV_TRACE") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count(); };
    auto t0 = now();
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    auto t1 = now();
    if (h->scat_ipc) {                              // every mapped peer buffer is in one of the two tables
        for (auto *tab : {&h->scat_cur, &h->scat_other})
            for (size_t r = 0; r < tab->size(); ++r) if ((*tab)[r] && (int)r != h->rank) cudaIpcCloseMemHandle((*tab)[r]);
        if (h->comm) ((qsvx::Comm *)h->comm)->peer.clear();
    }
    qsv_comm_teardown(h);
    if (h->io_stream) {
        cudaStreamSynchronize(h->io_stream);
        cudaStreamDestroy(h->io_stream);
        cudaEventDestroy(h->ev_snap); cudaEventDestroy(h->ev_io);
    }
    if (h->d_state != h->alloc_base) std::swap(h->d_state, h->d_shadow);     // a scatter pass left the roles exchanged
    if (h->d_shadow) cudaFree(h->d_shadow);
    for (auto &t : h->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
    if (h->t0) { cudaEventDestroy(h->t0); cudaEventDestroy(h->t1); }
    auto t2 = now();
    if (h->n_amps * h->amp_bytes >= ((size_t)64 << 20)) cache_put(h->device, h->d_state, h->n_amps * h->amp_bytes + qsvx::kTailBytes);
    else cudaFree(h->d_state);
    auto t3 = now();
    dev_free(h, h->d_partials); dev_free(h, h->d_pass_scratch); dev_free(h, h->d_ops_scratch);
    auto t4 = now();
    cudaStreamDestroy(h->stream);
    auto t5 = now();
    if (trace) fprintf(stderr, "[qsv] destroy: sync %.2f teardown %.2f state %.2f scratch %.2f stream %.2f ms\n",
                       ms(t0, t1), ms(t1, t2), ms(t2, t3), ms(t3, t4), ms(t4, t5));
    delete h;
    return QSV_OK;
}

int qsv_release_cached(void) {
    std::lock_guard<std::mutex> g(g_cache_mu);
    int dev0 = 0;
    cudaGetDevice(&dev0);
    for (int d = 0; d < 64; ++d)
        if (g_cache[d].ptr) { cudaSetDevice(d); cudaFree(g_cache[d].ptr); g_cache[d] = CachedShard(); }
    cudaSetDevice(dev0);
    return QSV_OK;
}

int qsv_sync(qsv_handle *h) {
    QSV_CHECK_H(h);
    QSV_CUDA(h, cudaSetDevice(h->device));
    QSV_CUDA(h, cudaStreamSynchronize(h->stream));
    return qsvx_check_peers_alive(h);
}

int qsv_device_ptr(qsv_handle *h, void **ptr, size_t *n_amps_local, void **stream) {
    QSV_CHECK_H(h);
    if (ptr) *ptr = h->d_state;
    if (n_amps_local) *n_amps_local = h->n_amps;
    if (stream) *stream = (void *)h->stream;
    return QSV_OK;
}

// ------------------------------------------------------------------ state I/O ----
int qsv_init_basis(qsv_handle *h, uint64_t index) {
    QSV_CHECK_H(h);
    if (h->n_qubits < 64 && (index >> h->n_qubits)) QSV_FAIL(h, QSV_EINVAL, "basis index out of range");
    QSV_CUDA(h, cudaSetDevice(h->device));
    QSV_CUDA(h, cudaMemsetAsync(h->d_state, 0, h->n_amps * h->amp_bytes, h->stream));
    if ((index >> h->n_local) == (uint64_t)h->rank) {
        const uint64_t li = index & (h->n_amps - 1);
        if (h->dtype == QSV_C128) k_set_amp<double><<<1, 1, 0, h->stream>>>((double2 *)h->d_state, li, 1.0, 0.0);
        else k_set_amp<float><<<1, 1, 0, h->stream>>>((float2 *)h->d_state, li, 1.0f, 0.0f);
        QSV_CUDA(h, cudaGetLastError());
    }
    return QSV_OK;
}

int qsv_init_zero(qsv_handle *h) { return qsv_init_basis(h, 0); }

static int copy_range(qsv_handle *h, void *host, size_t off, size_t n, bool to_device, bool sync) {
    QSV_CHECK_H(h);
    if (!host && n) QSV_FAIL(h, QSV_EINVAL, "null host buffer");
    if (off > h->n_amps || n > h->n_amps - off) QSV_FAIL(h, QSV_EINVAL, "range [%zu,+%zu) outside shard of %zu amps", off, n, h->n_amps);
    QSV_CUDA(h, cudaSetDevice(h->device));
    char *dev = (char *)h->d_state + off * h->amp_bytes;
    if (to_device) QSV_CUDA(h, cudaMemcpyAsync(dev, host, n * h->amp_bytes, cudaMemcpyHostToDevice, h->stream));
    else QSV_CUDA(h, cudaMemcpyAsync(host, dev, n * h->amp_bytes, cudaMemcpyDeviceToHost, h->stream));
    if (sync) QSV_CUDA(h, cudaStreamSynchronize(h->stream));
    return QSV_OK;
}
int qsv_upload(qsv_handle *h, const void *host, size_t off, size_t n) { return copy_range(h, (void *)host, off, n, true, true); }
int qsv_download(qsv_handle *h, void *host, size_t off, size_t n) { return copy_range(h, host, off, n, false, true); }
int qsv_upload_async(qsv_handle *h, const void *host, size_t off, size_t n) { return copy_range(h, (void *)host, off, n, true, false); }
int qsv_download_async(qsv_handle *h, void *host, size_t off, size_t n) { return copy_range(h, host, off, n, false, false); }

// ---- asynchronous checkpoints: snapshot on the device, drain beside the compute ----
// The reference's pipelined runner overlaps chunk I/O with gate application through reader / worker / writer
// threads (wenbo_engine/runner/pipeline.py:50-82).  With the state resident in HBM the only I/O is the
// checkpoint: qsv_snapshot copies the shard into the second buffer at HBM speed (stream ordered, ~10 ms per
// 16 GiB), the caller's writer thread then pulls the SNAPSHOT to pinned host memory on a second stream while
// the compute stream already runs the next steps.
int qsv_snapshot(qsv_handle *h) {
    QSV_CHECK_H(h);
    if (h->scat_ready) QSV_FAIL(h, QSV_EINVAL, "snapshot: the second buffer is in use by scatter passes");
    int rc = qsv_shadow_ptr(h, nullptr);
    if (rc) return rc;
    if (!h->io_stream) {
        QSV_CUDA(h, cudaStreamCreateWithFlags(&h->io_stream, cudaStreamNonBlocking));
        QSV_CUDA(h, cudaEventCreateWithFlags(&h->ev_snap, cudaEventDisableTiming));
        QSV_CUDA(h, cudaEventCreateWithFlags(&h->ev_io, cudaEventDisableTiming));
    }
    if (h->io_pending) QSV_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_io, 0));     // the previous snapshot is still being read
    QSV_CUDA(h, cudaMemcpyAsync(h->d_shadow, h->d_state, h->n_amps * h->amp_bytes, cudaMemcpyDeviceToDevice, h->stream));
    QSV_CUDA(h, cudaEventRecord(h->ev_snap, h->stream));
    QSV_CUDA(h, cudaStreamWaitEvent(h->io_stream, h->ev_snap, 0));
    return QSV_OK;
}

int qsv_snapshot_download_async(qsv_handle *h, void *host, size_t off, size_t n) {
    QSV_CHECK_H(h);
    if (!h->io_stream || !h->d_shadow) QSV_FAIL(h, QSV_EINVAL, "snapshot_download: no snapshot was taken");
    if (!host && n) QSV_FAIL(h, QSV_EINVAL, "null host buffer");
    if (off > h->n_amps || n > h->n_amps - off) QSV_FAIL(h, QSV_EINVAL, "range [%zu,+%zu) outside shard of %zu amps", off, n, h->n_amps);
    QSV_CUDA(h, cudaSetDevice(h->device));                  // called from the writer thread
    QSV_CUDA(h, cudaMemcpyAsync(host, (char *)h->d_shadow + off * h->amp_bytes, n * h->amp_bytes, cudaMemcpyDeviceToHost, h->io_stream));
    QSV_CUDA(h, cudaEventRecord(h->ev_io, h->io_stream));
    h->io_pending = true;
    return QSV_OK;
}

int qsv_snapshot_sync(qsv_handle *h) {
    QSV_CHECK_H(h);
    if (!h->io_stream) return QSV_OK;
    QSV_CUDA(h, cudaSetDevice(h->device));
    QSV_CUDA(h, cudaStreamSynchronize(h->io_stream));
    return QSV_OK;
}

int qsv_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return QSV_EINVAL;
    cudaError_t e = cudaHostAlloc(ptr, bytes, cudaHostAllocDefault);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); cudaGetLastError(); return QSV_ENOMEM; }
    return QSV_OK;
}
int qsv_host_free(void *ptr) { return cudaFreeHost(ptr) == cudaSuccess ? QSV_OK : QSV_ECUDA; }

// ---------------------------------------------------------- per-gate operators ----
static int check_local(qsv_handle *h, int q, const char *what) {
    if (q < 0 || q >= h->n_qubits) QSV_FAIL(h, QSV_EINVAL, "%s: qubit %d out of range [0,%d)", what, q, h->n_qubits);
    if (q >= h->n_local)
        QSV_FAIL(h, QSV_ENONLOCAL, "%s: qubit %d >= n_local=%d: non-local gate requires a remap step", what, q, h->n_local);
    return QSV_OK;
}

int qsv_apply_1q(qsv_handle *h, int q, const double U[8]) {
    QSV_CHECK_H(h);
    if (!U) QSV_FAIL(h, QSV_EINVAL, "null matrix");
    int rc = check_local(h, q, "apply_1q");
    if (rc) return rc;
    QSV_CUDA(h, cudaSetDevice(h->device));
    const uint64_t pairs = h->n_amps >> 1;
    ScopedTimer t(h, 1);
    if (h->dtype == QSV_C128)
        k_apply_1q<double, 2><<<grid_for(pairs, kGateThreads * 2), kGateThreads, 0, h->stream>>>((double2 *)h->d_state, pairs, q, to_mat2<double>(U));
    else
        k_apply_1q<float, 2><<<grid_for(pairs, kGateThreads * 2), kGateThreads, 0, h->stream>>>((float2 *)h->d_state, pairs, q, to_mat2<float>(U));
    QSV_CUDA(h, cudaGetLastError());
    return QSV_OK;
}

int qsv_apply_ctrl_1q(qsv_handle *h, int ctrl, int tgt, const double U[8]) {
    QSV_CHECK_H(h);
    if (!U) QSV_FAIL(h, QSV_EINVAL, "null matrix");
    if (ctrl == tgt) QSV_FAIL(h, QSV_EINVAL, "ctrl == tgt");
    if (ctrl < 0 || ctrl >= h->n_qubits) QSV_FAIL(h, QSV_EINVAL, "apply_ctrl_1q: control %d out of range", ctrl);
    int rc = check_local(h, tgt, "apply_ctrl_1q");
    if (rc) return rc;
    if (ctrl >= h->n_local) {           // control is a rank bit: whole shard or nothing
        if ((h->rank >> (ctrl - h->n_local)) & 1) return qsv_apply_1q(h, tgt, U);
        return QSV_OK;
    }
    QSV_CUDA(h, cudaSetDevice(h->device));
    const uint64_t quads = h->n_amps >> 2;
    ScopedTimer t(h, 2);
    if (h->dtype == QSV_C128)
        k_apply_ctrl_1q<double><<<grid_for(quads, kGateThreads), kGateThreads, 0, h->stream>>>((double2 *)h->d_state, quads, ctrl, tgt, to_mat2<double>(U));
    else
        k_apply_ctrl_1q<float><<<grid_for(quads, kGateThreads), kGateThreads, 0, h->stream>>>((float2 *)h->d_state, quads, ctrl, tgt, to_mat2<float>(U));
    QSV_CUDA(h, cudaGetLastError());
    return QSV_OK;
}

int qsv_apply_2q(qsv_handle *h, int qa, int qb, const double U[32]) {
    QSV_CHECK_H(h);
    if (!U) QSV_FAIL(h, QSV_EINVAL, "null matrix");
    if (qa == qb) QSV_FAIL(h, QSV_EINVAL, "apply_2q: qa == qb");
    int rc = check_local(h, qa, "apply_2q");
    if (rc) return rc;
    rc = check_local(h, qb, "apply_2q");
    if (rc) return rc;
    QSV_CUDA(h, cudaSetDevice(h->device));
    const uint64_t quads = h->n_amps >> 2;
    ScopedTimer t(h, 3);
    if (h->dtype == QSV_C128)
        k_apply_2q<double><<<grid_for(quads, kGateThreads), kGateThreads, 0, h->stream>>>((double2 *)h->d_state, quads, qa, qb, to_mat4<double>(U));
    else
        k_apply_2q<float><<<grid_for(quads, kGateThreads), kGateThreads, 0, h->stream>>>((float2 *)h->d_state, quads, qa, qb, to_mat4<float>(U));
    QSV_CUDA(h, cudaGetLastError());
    return QSV_OK;
}

}  // extern "C"
template <typename R>
static int launch_diag(qsv_handle *h, int nq, const int *qs, const double *phases) {
    DiagArgs<R> a;
    memset(&a, 0, sizeof(a));
    a.nq = nq;
    for (int i = 0; i < nq; ++i) a.qs[i] = qs[i];
    for (int i = 0; i < (1 << nq); ++i) { a.phase[i].x = (R)phases[2 * i]; a.phase[i].y = (R)phases[2 * i + 1]; }
    // args travel through the ops scratch buffer (stream ordered)
    if (h->ops_scratch_cap < sizeof(a)) {
        dev_free(h, h->d_ops_scratch);
        h->ops_scratch_cap = 1 << 20;
        QSV_CUDA(h, dev_alloc(h, (void **)&h->d_ops_scratch, h->ops_scratch_cap));
    }
    QSV_CUDA(h, cudaMemcpyAsync(h->d_ops_scratch, &a, sizeof(a), cudaMemcpyHostToDevice, h->stream));
    QSV_CUDA(h, cudaStreamSynchronize(h->stream));   // `a` is a stack object
    ScopedTimer t(h, 4);
    k_apply_diag<R><<<grid_for(h->n_amps, kGateThreads), kGateThreads, 0, h->stream>>>(
        (typename CxT<R>::V *)h->d_state, h->n_amps, (uint64_t)h->rank << h->n_local,
        (const DiagArgs<R> *)h->d_ops_scratch);
    QSV_CUDA(h, cudaGetLastError());
    return QSV_OK;
}

extern "C" {
int qsv_apply_diag(qsv_handle *h, int nq, const int *qs, const double *phases) {
    QSV_CHECK_H(h);
    if (nq < 1 || nq > 6 || !qs || !phases) QSV_FAIL(h, QSV_EINVAL, "apply_diag: need 1..6 qubits");
    for (int i = 0; i < nq; ++i) {
        if (qs[i] < 0 || qs[i] >= h->n_qubits) QSV_FAIL(h, QSV_EINVAL, "apply_diag: qubit %d out of range", qs[i]);
        for (int j = 0; j < i; ++j) if (qs[i] == qs[j]) QSV_FAIL(h, QSV_EINVAL, "apply_diag: repeated qubit %d", qs[i]);
    }
    QSV_CUDA(h, cudaSetDevice(h->device));
    return h->dtype == QSV_C128 ? launch_diag<double>(h, nq, qs, phases) : launch_diag<float>(h, nq, qs, phases);
}

}  // extern "C"
template <typename R, int K>
static int launch_kq(qsv_handle *h, const int *meta_dev, const void *U_dev) {
    using V = typename CxT<R>::V;
    const uint64_t groups = h->n_amps >> K;
    const size_t smem = sizeof(V) << (2 * K);
    QSV_CUDA(h, cudaFuncSetAttribute(k_apply_kq<R, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ScopedTimer t(h, 5);
    k_apply_kq<R, K><<<grid_for(groups, 128), 128, smem, h->stream>>>((V *)h->d_state, groups, meta_dev, (const V *)U_dev);
    QSV_CUDA(h, cudaGetLastError());
    return QSV_OK;
}

extern "C" {
int qsv_apply_kq(qsv_handle *h, int k, const int *qs, const double *U) {
    QSV_CHECK_H(h);
    if (k < 1 || k > 5 || !qs || !U) QSV_FAIL(h, QSV_EINVAL, "apply_kq: need 1 <= k <= 5");
    if (k > h->n_local) QSV_FAIL(h, QSV_EINVAL, "apply_kq: k=%d > n_local=%d", k, h->n_local);
    for (int i = 0; i < k; ++i) {
        int rc = check_local(h, qs[i], "apply_kq");
        if (rc) return rc;
        for (int j = 0; j < i; ++j) if (qs[i] == qs[j]) QSV_FAIL(h, QSV_EINVAL, "apply_kq: repeated qubit %d", qs[i]);
    }
    QSV_CUDA(h, cudaSetDevice(h->device));
    // meta: sorted qubits + the sub-space row bit each one drives (qs[0] = MSB)
    int meta[10];
    int order[5];
    for (int i = 0; i < k; ++i) order[i] = i;
    std::sort(order, order + k, [&](int a, int b) { return qs[a] < qs[b]; });
    for (int j = 0; j < k; ++j) { meta[j] = qs[order[j]]; meta[k + j] = k - 1 - order[j]; }
    const int D = 1 << k;
    const size_t mat_elems = (size_t)D * D;
    const size_t need = 64 + mat_elems * 16;
    if (h->ops_scratch_cap < need) {
        dev_free(h, h->d_ops_scratch);
        h->ops_scratch_cap = 1 << 20;
        QSV_CUDA(h, dev_alloc(h, (void **)&h->d_ops_scratch, h->ops_scratch_cap));
    }
    char *scratch = (char *)h->d_ops_scratch;
    std::vector<char> stage(need);
    memcpy(stage.data(), meta, sizeof(int) * 2 * k);
    if (h->dtype == QSV_C128) memcpy(stage.data() + 64, U, mat_elems * 16);
    else { float *f = (float *)(stage.data() + 64); for (size_t i = 0; i < mat_elems * 2; ++i) f[i] = (float)U[i]; }
    QSV_CUDA(h, cudaMemcpyAsync(scratch, stage.data(), need, cudaMemcpyHostToDevice, h->stream));
    QSV_CUDA(h, cudaStreamSynchronize(h->stream));
    const int *meta_dev = (const int *)scratch;
    const void *U_dev = scratch + 64;
#define KQ_CASE(K_) case K_: return h->dtype == QSV_C128 ? launch_kq<double, K_>(h, meta_dev, U_dev) : launch_kq<float, K_>(h, meta_dev, U_dev)
    switch (k) { KQ_CASE(1); KQ_CASE(2); KQ_CASE(3); KQ_CASE(4); KQ_CASE(5); }
#undef KQ_CASE
    return QSV_EINVAL;
}

// --------------------------------------------------------------- fused passes ----
static int validate_pass(qsv_handle *h, const qsv_pass *p, const qsv_op *ops) {
    const int T = p->n_tile;
    const int maxT = h->dtype == QSV_C128 ? 12 : 13;
    if (T < QSV_REG_BITS || T > maxT || T > h->n_local)
        QSV_FAIL(h, QSV_EINVAL, "pass: n_tile=%d outside [%d, min(%d, n_local=%d)]", T, QSV_REG_BITS, maxT, h->n_local);
    uint64_t tile_mask = 0, store_mask = 0;
    for (int i = 0; i < T; ++i) {
        const int b = p->load_bits[i], sb = p->store_bits[i];
        if (b < 0 || b >= h->n_local || (i && b <= p->load_bits[i - 1]))
            QSV_FAIL(h, QSV_EINVAL, "pass: load_bits must be ascending local bits (pos %d = %d)", i, b);
        if (sb < 0 || sb >= h->n_local || ((store_mask >> sb) & 1)) QSV_FAIL(h, QSV_EINVAL, "pass: bad store_bits[%d]=%d", i, sb);
        tile_mask |= 1ull << b; store_mask |= 1ull << sb;
    }
    if (tile_mask != store_mask) QSV_FAIL(h, QSV_EINVAL, "pass: store_bits is not a permutation of load_bits");
    if (p->store_flip & ~tile_mask) QSV_FAIL(h, QSV_EINVAL, "pass: store_flip names a bit outside the tile");
    if (p->n_active < -1 || p->n_active > QSV_MAX_ACTIVE_BITS || p->n_active > h->n_local - T)
        QSV_FAIL(h, QSV_EINVAL, "pass: n_active=%d", p->n_active);
    for (int k = 0; k < p->n_active; ++k) {
        const int b = p->active_bits[k];
        if (b < 0 || b >= h->n_local || ((tile_mask >> b) & 1) || (k && b <= p->active_bits[k - 1]))
            QSV_FAIL(h, QSV_EINVAL, "pass: active_bits must be ascending local non-tile positions (entry %d = %d)", k, b);
    }
    if (p->zero_input != 0 && p->zero_input != 1) QSV_FAIL(h, QSV_EINVAL, "pass: zero_input=%d", p->zero_input);
    if (p->n_rounds < 1 || p->n_rounds > QSV_MAX_ROUNDS) QSV_FAIL(h, QSV_EINVAL, "pass: n_rounds=%d", p->n_rounds);
    if (p->n_ops < 0 || p->n_fold < 0) QSV_FAIL(h, QSV_EINVAL, "pass: n_ops / n_fold < 0");
    for (int r = 0; r < p->n_rounds; ++r) {
        const qsv_round &rd = p->rounds[r];
        uint32_t seen = 0, regm = 0;
        for (int b = 0; b < QSV_REG_BITS; ++b) {
            if (rd.reg_pos[b] >= T || ((seen >> rd.reg_pos[b]) & 1)) QSV_FAIL(h, QSV_EINVAL, "pass: round %d bad reg_pos", r);
            seen |= 1u << rd.reg_pos[b];
        }
        regm = seen;
        for (int i = 0; i < T - QSV_REG_BITS; ++i) {
            if (rd.thr_pos[i] >= T || ((seen >> rd.thr_pos[i]) & 1)) QSV_FAIL(h, QSV_EINVAL, "pass: round %d bad thr_pos", r);
            seen |= 1u << rd.thr_pos[i];
        }
        if (rd.op_begin < 0 || rd.op_end < rd.op_begin || rd.op_end > p->n_ops) QSV_FAIL(h, QSV_EINVAL, "pass: round %d op slice", r);
        if (rd.fold_off < -1 || (rd.fold_off >= 0 && rd.fold_off + (1 << (T - QSV_REG_BITS)) > p->n_fold))
            QSV_FAIL(h, QSV_EINVAL, "pass: round %d fold table outside the pass's %d entries", r, p->n_fold);
        for (int o = rd.op_begin; o < rd.op_end; ++o) {
            const qsv_op &op = ops[o];
            if (op.kind >= QSV_OP_KINDS) QSV_FAIL(h, QSV_EINVAL, "pass: op %d bad kind %d", o, (int)op.kind);
            const bool has_target = op.kind == QSV_OP_HAD || op.kind == QSV_OP_ROT || op.kind == QSV_OP_XSWAP ||
                                    op.kind == QSV_OP_YSWAP || op.kind == QSV_OP_TPHASE;
            if (op.kind == QSV_OP_TPHASE) {
                const int toff = (int)op.m[0], goff = (int)op.m[1];
                const unsigned mask = (unsigned)op.m[2];
                if (op.reg_ctrl || op.tile_ctrl || op.glob_ctrl || op.flags) QSV_FAIL(h, QSV_EINVAL, "pass: op %d: TPHASE takes no controls", o);
                if (toff < -1 || (toff >= 0 && toff + (1 << (T - QSV_REG_BITS)) > p->n_fold)) QSV_FAIL(h, QSV_EINVAL, "pass: op %d: thread table outside the pass's tables", o);
                if (mask > 255u || goff < -1 || (goff >= 0 && goff + 256 * __builtin_popcount(mask) > p->n_fold) || (goff < 0 && mask))
                    QSV_FAIL(h, QSV_EINVAL, "pass: op %d: run tables outside the pass's tables", o);
            }
            if (has_target && op.target >= QSV_REG_BITS) QSV_FAIL(h, QSV_EINVAL, "pass: op %d target", o);
            const bool any_ctrl = op.reg_ctrl || op.tile_ctrl || op.glob_ctrl;
            if (op.flags & ~(QSV_OPF_PRESIGN | QSV_OPF_PRENEG | QSV_OPF_PREPHASE)) QSV_FAIL(h, QSV_EINVAL, "pass: op %d: unknown flags", o);
            if (op.flags) {
                if (op.kind != QSV_OP_HAD && op.kind != QSV_OP_ROT) QSV_FAIL(h, QSV_EINVAL, "pass: op %d: pre-ops need HAD or ROT", o);
                if (op.reg_ctrl) QSV_FAIL(h, QSV_EINVAL, "pass: op %d: pre-ops exclude register controls", o);
                if (!(op.flags & QSV_OPF_PRESIGN) && (op.tile_ctrl || op.glob_ctrl))
                    QSV_FAIL(h, QSV_EINVAL, "pass: op %d: parity masks without QSV_OPF_PRESIGN", o);
                if ((op.flags & QSV_OPF_PREPHASE) && !(op.m[2] >= -1.0000001 && op.m[2] <= 1.0000001))
                    QSV_FAIL(h, QSV_EINVAL, "pass: op %d: pre-phase |tan(phi/2)| must be <= 1", o);
            } else if ((op.kind == QSV_OP_HAD || op.kind == QSV_OP_SCALE) && any_ctrl)
                QSV_FAIL(h, QSV_EINVAL, "pass: op %d: HAD/SCALE cannot be controlled", o);
            if ((op.kind == QSV_OP_ROT || op.kind == QSV_OP_PHASE) && (!(op.m[0] >= -1.0000001 && op.m[0] <= 1.0000001)))
                QSV_FAIL(h, QSV_EINVAL, "pass: op %d: |tan(angle/2)| must be <= 1 (split larger angles)", o);
            if ((op.tile_ctrl >> T) || (op.tile_ctrl & regm)) QSV_FAIL(h, QSV_EINVAL, "pass: op %d tile_ctrl names a register position", o);
            if (op.glob_ctrl & tile_mask) QSV_FAIL(h, QSV_EINVAL, "pass: op %d glob_ctrl names a tile bit", o);
            if (h->n_qubits < 64 && (op.glob_ctrl >> h->n_qubits)) QSV_FAIL(h, QSV_EINVAL, "pass: op %d glob_ctrl out of range", o);
        }
    }
    return QSV_OK;
}

static int launch_pass(qsv_handle *h, const qsv_pass *host_pass, const qsv_pass *dev_pass, const qsv_op *dev_ops,
                       const double2 *dev_tables, int pass_index) {
    const int T = host_pass->n_tile;
    if (host_pass->zero_input) {          // the interpreting kernels read the shard: give them |0...0>
        int rc = qsv_init_zero(h);
        if (rc) return rc;
    }
    const uint64_t rank_bits = (uint64_t)h->rank << h->n_local;
    ScopedTimer t(h, 10, pass_index);
    if (h->dtype == QSV_C128 && T == kRingT && host_pass->n_ops <= kRingMaxOps && !h->force_simple_pass) {
        // persistent ring kernel: one CTA per SM, tiles streamed through shared memory
        const uint32_t n_tiles = (uint32_t)(h->n_amps >> T);
        const unsigned grid = n_tiles < (uint32_t)h->sm_count ? n_tiles : (unsigned)h->sm_count;
        const size_t smem = sizeof(RingSmem);
        QSV_CUDA(h, cudaFuncSetAttribute(k_pass_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_pass_ring<<<grid, kRingThreads, smem, h->stream>>>((double2 *)h->d_state, dev_pass, dev_ops, dev_tables, rank_bits, n_tiles);
        QSV_CUDA(h, cudaGetLastError());
        return QSV_OK;
    }
    const unsigned blocks = (unsigned)(h->n_amps >> T);
    const unsigned threads = 1u << (T - QSV_REG_BITS);
    const size_t smem = host_pass->n_rounds > 1 ? (h->amp_bytes << T) : 0;
    if (h->dtype == QSV_C128) {
        QSV_CUDA(h, cudaFuncSetAttribute(k_pass<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        k_pass<double><<<blocks, threads, smem, h->stream>>>((double2 *)h->d_state, dev_pass, dev_ops, dev_tables, rank_bits);
    } else {
        QSV_CUDA(h, cudaFuncSetAttribute(k_pass<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536));
        k_pass<float><<<blocks, threads, smem, h->stream>>>((float2 *)h->d_state, dev_pass, dev_ops, dev_tables, rank_bits);
    }
    QSV_CUDA(h, cudaGetLastError());
    return QSV_OK;
}

int qsv_apply_pass(qsv_handle *h, const qsv_pass *pass, const qsv_op *ops, const double *tables) {
    QSV_CHECK_H(h);
    if (!pass || (pass->n_ops > 0 && !ops) || (pass->n_fold > 0 && !tables)) QSV_FAIL(h, QSV_EINVAL, "apply_pass: null argument");
    int rc = validate_pass(h, pass, ops);
    if (rc) return rc;
    QSV_CUDA(h, cudaSetDevice(h->device));
    const size_t ops_bytes = sizeof(qsv_op) * (size_t)std::max(pass->n_ops, 1);
    const size_t tab_off = (ops_bytes + 255) & ~(size_t)255;
    const size_t need = tab_off + sizeof(double2) * (size_t)std::max(pass->n_fold, 1);
    if (h->ops_scratch_cap < need) {
        QSV_CUDA(h, cudaStreamSynchronize(h->stream));
        dev_free(h, h->d_ops_scratch);
        h->ops_scratch_cap = std::max(need, (size_t)1 << 20);
        QSV_CUDA(h, dev_alloc(h, (void **)&h->d_ops_scratch, h->ops_scratch_cap));
    }
    // stream order protects the scratch against the previous one-shot pass; the host buffers
    // are pageable, so cudaMemcpyAsync stages them before returning.
    char *scratch = (char *)h->d_ops_scratch;
    QSV_CUDA(h, cudaMemcpyAsync(h->d_pass_scratch, pass, sizeof(qsv_pass), cudaMemcpyHostToDevice, h->stream));
    if (pass->n_ops) QSV_CUDA(h, cudaMemcpyAsync(scratch, ops, sizeof(qsv_op) * pass->n_ops, cudaMemcpyHostToDevice, h->stream));
    if (pass->n_fold) QSV_CUDA(h, cudaMemcpyAsync(scratch + tab_off, tables, sizeof(double2) * pass->n_fold, cudaMemcpyHostToDevice, h->stream));
    return launch_pass(h, pass, h->d_pass_scratch, (const qsv_op *)scratch, (const double2 *)(scratch + tab_off), -1);
}

int qsv_program_create(qsv_handle *h, const qsv_pass *passes, int n_passes, const qsv_op *ops, const double *tables,
                       qsv_program **out) {
    QSV_CHECK_H(h);
    if (!out || n_passes < 0 || (n_passes && !passes)) QSV_FAIL(h, QSV_EINVAL, "program_create: bad arguments");
    *out = nullptr;
    QSV_CUDA(h, cudaSetDevice(h->device));
    qsv_program *p = new (std::nothrow) qsv_program();
    if (!p) return QSV_ENOMEM;
    int total_ops = 0, total_fold = 0;
    for (int i = 0; i < n_passes; ++i) {
        int rc = validate_pass(h, &passes[i], ops ? ops + total_ops : nullptr);
        if (rc) { delete p; return rc; }
        p->passes.push_back(passes[i]);
        p->op_offset.push_back(total_ops);
        p->fold_offset.push_back(total_fold);
        total_ops += passes[i].n_ops;
        total_fold += passes[i].n_fold;
    }
    if (total_fold > 0 && !tables) { delete p; QSV_FAIL(h, QSV_EINVAL, "program_create: fold tables missing"); }
    cudaError_t e = cudaSuccess;
    if (n_passes) e = dev_alloc(h, (void **)&p->d_passes, sizeof(qsv_pass) * n_passes);
    if (e == cudaSuccess) e = dev_alloc(h, (void **)&p->d_ops, sizeof(qsv_op) * std::max(total_ops, 1));
    if (e == cudaSuccess) e = dev_alloc(h, (void **)&p->d_tables, sizeof(double2) * std::max(total_fold, 1));
    if (e == cudaSuccess && total_fold) e = cudaMemcpyAsync(p->d_tables, tables, sizeof(double2) * total_fold, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess && n_passes) e = cudaMemcpyAsync(p->d_passes, passes, sizeof(qsv_pass) * n_passes, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess && total_ops) e = cudaMemcpyAsync(p->d_ops, ops, sizeof(qsv_op) * total_ops, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);          // the host arrays belong to the caller
    if (e != cudaSuccess) {
        h->err = std::string("program_create: ") + cudaGetErrorString(e);
        cudaGetLastError();
        dev_free(h, p->d_passes); dev_free(h, p->d_ops); dev_free(h, p->d_tables);
        delete p;
        return QSV_ECUDA;
    }
    if (ops && total_ops) p->h_ops.assign(ops, ops + total_ops);      // scatter kernels are generated on demand
    // specialise every eligible pass (complex128 ring tiles); the rest is interpreted
    p->jit.assign(n_passes, nullptr);
    p->jit_coefs.assign(n_passes, {});
    if (h->jit && qsvjit::enabled() && !h->force_simple_pass) {
        const bool f32 = h->dtype == QSV_C64;
        std::vector<std::string> srcs(n_passes);
        p->jit_coefs_f.assign(n_passes, {});
        for (int i = 0; i < n_passes; ++i) {
            if (!qsvjit::generate(passes[i], ops ? ops + p->op_offset[i] : nullptr, srcs[i], p->jit_coefs[i], f32)) srcs[i].clear();
            if (f32) p->jit_coefs_f[i].assign(p->jit_coefs[i].begin(), p->jit_coefs[i].end());
        }
        std::vector<qsvjit::Kernel> ks;
        std::string err;
        qsvjit::resolve(h->device, srcs, ks, err);
        for (int i = 0; i < n_passes; ++i) { p->jit[i] = ks[i].fn; if (ks[i].key) p->jit_keys.push_back(ks[i].key); }
        if (!err.empty()) h->err = "jit (falling back to the interpreting kernel): " + err;
    }
    *out = p;
    return QSV_OK;
}

// tiles [tile_begin, tile_end) of pass i on `stream` (the whole pass: 0 .. n_amps >> 11)
struct JitFixArg { unsigned n; unsigned pos[4]; unsigned blk; unsigned long long val; };   // = JitFix of jit_prelude.cuh

// QSV_JIT_TILE_BLOCK=k (default 1): a CTA is dealt 2^k consecutive tiles at a time instead of every grid-th tile
// (jit_seq in jit_prelude.cuh) — what makes the two tiles of a paired load (QSV_JIT_PAIR) neighbours in memory.
// Clipped so that every CTA still gets several blocks.  Alone it changes nothing (k = 4: 40.80 vs 40.82 ms).
constexpr int kDefaultTileBlock = 1;
static unsigned tile_block_log2(uint32_t count, unsigned grid) {
    static const int want = [] { const char *e = getenv("QSV_JIT_TILE_BLOCK"); int k = (e && e[0]) ? atoi(e) : kDefaultTileBlock; return k < 0 ? 0 : (k > 8 ? 8 : k); }();
    unsigned k = (unsigned)want;
    while (k > 0 && (((uint64_t)grid << k) * 4u) > count) --k;
    return k;
}

static int launch_pass_jit_range(qsv_handle *h, qsv_program *p, int i, uint32_t tile_begin, uint32_t tile_end, cudaStream_t stream,
                                 const JitFixArg *fix_in = nullptr, unsigned grid_cap = 0) {
    const uint32_t count = tile_end - tile_begin;
    const unsigned cap = grid_cap ? grid_cap : (unsigned)h->sm_count;
    const unsigned grid = count < cap ? count : cap;
    JitFixArg fix = {};
    if (fix_in) fix = *fix_in;
    fix.blk = tile_block_log2(count, grid);
    void *state = h->d_state;
    const double2 *tables = p->d_tables + p->fold_offset[i];
    unsigned long long rank_bits = (unsigned long long)h->rank << h->n_local;
    unsigned tb = tile_begin, te = tile_end;
    static const double zero = 0.0;
    void *coefs = p->jit_coefs[i].empty() ? (void *)&zero
                : (h->dtype == QSV_C64 ? (void *)p->jit_coefs_f[i].data() : (void *)p->jit_coefs[i].data());
    void *args[] = {&state, &tables, &rank_bits, &tb, &te, coefs, &fix};
    QSV_CUDA(h, cudaLaunchKernel((const void *)p->jit[i], dim3(grid), dim3(128 * (qsvjit::groups() + 1)), args, qsvjit::kSmemBytes, stream));
    return QSV_OK;
}

// pass i restricted to the chunk whose index bits chunk_bits[] (local positions outside the tile) equal chunk_j
static int launch_pass_jit_chunk(qsv_handle *h, qsv_program *p, int i, int n_chunk, const int *chunk_bits, unsigned chunk_j,
                                 unsigned grid_cap, cudaStream_t stream) {
    const qsv_pass &P = p->passes[i];
    JitFixArg fix = {};
    fix.n = (unsigned)n_chunk;
    for (int k = 0; k < n_chunk; ++k) {
        int below = 0;
        for (int t = 0; t < P.n_tile; ++t) below += P.load_bits[t] < chunk_bits[k];
        fix.pos[k] = (unsigned)(chunk_bits[k] - below);                  // position in tile-index space
        fix.val |= (unsigned long long)((chunk_j >> k) & 1u) << fix.pos[k];
    }
    const uint32_t tiles = (uint32_t)((h->n_amps >> qsvjit::kT) >> n_chunk);
    ScopedTimer t(h, 11, i);                                             // kind 11: one chunk of a pass
    return launch_pass_jit_range(h, p, i, 0u, tiles, stream, &fix, grid_cap);
}

static int launch_pass_jit(qsv_handle *h, qsv_program *p, int i) {
    ScopedTimer t(h, 10, i);
    if (p->passes[i].zero_input && !getenv("QSV_INIT_PASS_FULL")) {
        // Fused |0...0> initialisation: the pass reads nothing, and every tile but the one that holds
        // amplitude 0 (tile 0 of rank 0) is identically zero before AND after it (gates are linear;
        // the store flips only move whole tiles).  So: zero-fill the shard at write bandwidth and run
        // the arithmetic for that one tile only.  QSV_INIT_PASS_FULL=1 runs every tile as before.
        QSV_CUDA(h, cudaMemsetAsync(h->d_state, 0, h->n_amps * h->amp_bytes, h->stream));
        if (h->rank != 0) return QSV_OK;
        return launch_pass_jit_range(h, p, i, 0u, 1u, h->stream);
    }
    const int na = p->passes[i].n_active;                      // zero-support skipping: 2^n_active tiles
    const uint32_t tiles = na >= 0 ? (uint32_t)1 << na : (uint32_t)(h->n_amps >> qsvjit::kT);
    return launch_pass_jit_range(h, p, i, 0u, tiles, h->stream);
}

int qsv_program_run(qsv_handle *h, qsv_program *p) {
    QSV_CHECK_H(h);
    if (!p) QSV_FAIL(h, QSV_EINVAL, "program_run: null program");
    QSV_CUDA(h, cudaSetDevice(h->device));
    for (size_t i = 0; i < p->passes.size(); ++i) {
        int rc = p->jit[i] ? launch_pass_jit(h, p, (int)i)
                           : launch_pass(h, &p->passes[i], p->d_passes + i, p->d_ops + p->op_offset[i],
                                         p->d_tables + p->fold_offset[i], (int)i);
        if (rc) return rc;
    }
    return QSV_OK;
}

int qsv_program_run_range(qsv_handle *h, qsv_program *p, int first, int count) {
    QSV_CHECK_H(h);
    if (!p || first < 0 || count < 0 || (size_t)(first + count) > p->passes.size()) QSV_FAIL(h, QSV_EINVAL, "program_run_range: bad slice");
    QSV_CUDA(h, cudaSetDevice(h->device));
    for (int i = first; i < first + count; ++i) {
        int rc = p->jit[i] ? launch_pass_jit(h, p, i)
                           : launch_pass(h, &p->passes[i], p->d_passes + i, p->d_ops + p->op_offset[i],
                                         p->d_tables + p->fold_offset[i], i);
        if (rc) return rc;
    }
    return QSV_OK;
}

// ------------------------------------------------------- pipelined stage transition ----
// A swap of rank bits with local bits, PIPELINED with the passes around it.  The shard is cut into
// 2^n_chunk CHUNKS by index bits that none of those passes has in its tile (and that are not swapped):
// a pass restricted to a chunk is an independent launch (tiles never cross chunks), and so is the
// exchange of a chunk.  For chunk j:  passes a_first .. a_first+a_count-1 of `pa`  ->  exchange  ->
// passes 0 .. b_count-1 of `pb`, with the pass launches on the handle's stream (sm_count - xchg_sms
// CTAs) and the exchange kernels (csrc/xchg.cuh, xchg_sms CTAs of one SM each) on the copy stream, so
// the NVLink transfer of chunk j runs beside the arithmetic of chunks j+1, j+2 (before the swap) and
// j-1, j-2 (after it).  Cross-GPU ordering is inside the exchange kernels (flag words in peer memory);
// nothing here blocks the host.  Every rank of the group must make the same call (same programs, same
// bits): the CALLER agrees on that collectively (qsv_program_specialised over all ranks) — the library
// never decides per rank.  Conditions that do not hold are an error, not a silent fallback.
int qsv_program_specialised(qsv_handle *h, qsv_program *p, int *flags, int n) {
    QSV_CHECK_H(h);
    if (!p || !flags || n < (int)p->passes.size()) QSV_FAIL(h, QSV_EINVAL, "program_specialised: bad arguments");
    for (size_t i = 0; i < p->passes.size(); ++i) flags[i] = p->jit[i] != nullptr;
    return QSV_OK;
}

int qsv_swap_pipelined(qsv_handle *h, qsv_program *pa, int a_first, int a_count, qsv_program *pb, int b_count,
                       int n_swap, const int *global_bits, const int *local_bits, int n_chunk, const int *chunk_bits,
                       int xchg_sms) {
    QSV_CHECK_H(h);
    auto *c = (qsvx::Comm *)h->comm;
    if (!c || !c->peers_ready || !h->use_peer_swap) QSV_FAIL(h, QSV_ECOMM, "swap_pipelined: peers are not mapped");
    const int g = h->n_qubits - h->n_local;
    if (n_swap < 1 || n_swap > 3 || n_swap > g || !global_bits || !local_bits) QSV_FAIL(h, QSV_EINVAL, "swap_pipelined: bad swap bits");
    if (n_chunk < 1 || n_chunk > 4 || !chunk_bits) QSV_FAIL(h, QSV_EINVAL, "swap_pipelined: 1..4 chunk bits");
    if (a_count < 0 || b_count < 0 || (a_count && !pa) || (b_count && !pb)) QSV_FAIL(h, QSV_EINVAL, "swap_pipelined: bad pass ranges");
    if (a_count && (a_first < 0 || (size_t)(a_first + a_count) > pa->passes.size())) QSV_FAIL(h, QSV_EINVAL, "swap_pipelined: pass range A");
    if (b_count && (size_t)b_count > pb->passes.size()) QSV_FAIL(h, QSV_EINVAL, "swap_pipelined: pass range B");
    uint64_t special = 0;
    for (int i = 0; i < n_swap; ++i) {
        if (global_bits[i] < h->n_local || global_bits[i] >= h->n_qubits) QSV_FAIL(h, QSV_EINVAL, "swap_pipelined: global bit %d is not a rank bit", global_bits[i]);
        if (local_bits[i] < 0 || local_bits[i] >= h->n_local || ((special >> local_bits[i]) & 1)) QSV_FAIL(h, QSV_EINVAL, "swap_pipelined: local bit %d", local_bits[i]);
        for (int j = 0; j < i; ++j) if (global_bits[i] == global_bits[j]) QSV_FAIL(h, QSV_EINVAL, "swap_pipelined: repeated rank bit");
        special |= 1ull << local_bits[i];
    }
    uint64_t chunk_mask = 0;
    for (int i = 0; i < n_chunk; ++i) {
        if (chunk_bits[i] < 0 || chunk_bits[i] >= h->n_local || ((special >> chunk_bits[i]) & 1) || (i && chunk_bits[i] <= chunk_bits[i - 1]))
            QSV_FAIL(h, QSV_EINVAL, "swap_pipelined: chunk bits must be ascending local positions that are not swapped (entry %d = %d)", i, chunk_bits[i]);
        special |= 1ull << chunk_bits[i];
        chunk_mask |= 1ull << chunk_bits[i];
    }
    auto check_pass = [&](qsv_program *p, int i) -> const char * {
        if (!p->jit[i]) return "is not specialised";
        const qsv_pass &P = p->passes[i];
        if (P.n_active >= 0 || P.zero_input) return "skips tiles / creates the state itself";
        for (int k = 0; k < P.n_tile; ++k) if ((chunk_mask >> P.load_bits[k]) & 1) return "has a chunk bit in its tile";
        return nullptr;
    };
    for (int i = 0; i < a_count; ++i) if (const char *why = check_pass(pa, a_first + i)) QSV_FAIL(h, QSV_EINVAL, "swap_pipelined: pass %d before the swap %s", a_first + i, why);
    for (int i = 0; i < b_count; ++i) if (const char *why = check_pass(pb, i)) QSV_FAIL(h, QSV_EINVAL, "swap_pipelined: pass %d after the swap %s", i, why);
    if (h->n_local - n_chunk < qsvjit::kT) QSV_FAIL(h, QSV_EINVAL, "swap_pipelined: chunks smaller than a tile");
    QSV_CUDA(h, cudaSetDevice(h->device));
    if (xchg_sms < 1) xchg_sms = 1;
    if (xchg_sms > h->sm_count / 2) xchg_sms = h->sm_count / 2;
    const unsigned pass_grid = (unsigned)(h->sm_count - xchg_sms);
    const int J = 1 << n_chunk;
    // a spinning exchange kernel must never wait for a kernel whose FIRST launch still has to load its
    // module (lazy loading synchronises the context): touch every function of the pipeline now
    {
        cudaFuncAttributes fa;
        for (int i = 0; i < a_count; ++i) QSV_CUDA(h, cudaFuncGetAttributes(&fa, (const void *)pa->jit[a_first + i]));
        for (int i = 0; i < b_count; ++i) QSV_CUDA(h, cudaFuncGetAttributes(&fa, (const void *)pb->jit[i]));
        QSV_CUDA(h, cudaFuncGetAttributes(&fa, (const void *)qsvx::k_xchg_tma));
        QSV_CUDA(h, cudaFuncGetAttributes(&fa, (const void *)qsvx::k_xchg_ldst<4>));
    }
    while ((int)c->ev_pool.size() < 2 * J) {
        cudaEvent_t e;
        QSV_CUDA(h, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->ev_pool.push_back(e);
    }
    cudaStream_t S = h->stream, X = c->copy_stream;
    cudaEvent_t *evA = c->ev_pool.data(), *evX = c->ev_pool.data() + J;
    ScopedTimer timer(h, 50 + n_swap, a_count * 16 + b_count);
    const int lag = 2;
    auto after = [&](int j) -> int {                     // passes after the swap on chunk j
        QSV_CUDA(h, cudaStreamWaitEvent(S, evX[j], 0));
        for (int i = 0; i < b_count; ++i) {
            int rc = launch_pass_jit_chunk(h, pb, i, n_chunk, chunk_bits, (unsigned)j, pass_grid, S);
            if (rc) return rc;
        }
        return QSV_OK;
    };
    for (int j = 0; j < J; ++j) {
        for (int i = 0; i < a_count; ++i) {
            int rc = launch_pass_jit_chunk(h, pa, a_first + i, n_chunk, chunk_bits, (unsigned)j, pass_grid, S);
            if (rc) return rc;
        }
        QSV_CUDA(h, cudaEventRecord(evA[j], S));
        QSV_CUDA(h, cudaStreamWaitEvent(X, evA[j], 0));
        qsvx::XchgArgs A;
        bool tma_ok = false;
        ++c->seq;
        int rc = qsvx_xchg_args(h, c, n_swap, global_bits, local_bits, n_chunk, chunk_bits, (unsigned)j, A, tma_ok);
        if (rc) return rc;
        rc = qsvx_xchg_launch(h, c, X, A, tma_ok, xchg_sms);
        if (rc) return rc;
        QSV_CUDA(h, cudaEventRecord(evX[j], X));
        if (j >= lag) { rc = after(j - lag); if (rc) return rc; }
    }
    for (int j = J - lag < 0 ? 0 : J - lag; j < J; ++j) { int rc = after(j); if (rc) return rc; }
    return QSV_OK;
}

// ------------------------------------------------------------------ scatter pass ----
int qsv_shadow_ptr(qsv_handle *h, void **ptr) {
    QSV_CHECK_H(h);
    QSV_CUDA(h, cudaSetDevice(h->device));
    if (!h->d_shadow) {
        cudaError_t e = cudaMalloc(&h->d_shadow, h->n_amps * h->amp_bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            h->d_shadow = nullptr;
            QSV_FAIL(h, e == cudaErrorMemoryAllocation ? QSV_ENOMEM : QSV_ECUDA, "shadow buffer (%zu bytes): %s",
                     h->n_amps * h->amp_bytes, cudaGetErrorString(e));
        }
    }
    if (ptr) *ptr = h->d_shadow;
    return QSV_OK;
}

int qsv_comm_shadow_ipc_handle(qsv_handle *h, void *out64) {
    QSV_CHECK_H(h);
    if (!out64) QSV_FAIL(h, QSV_EINVAL, "shadow_ipc_handle: null output");
    int rc = qsv_shadow_ptr(h, nullptr);
    if (rc) return rc;
    cudaIpcMemHandle_t mh;
    QSV_CUDA(h, cudaIpcGetMemHandle(&mh, h->d_shadow));
    memcpy(out64, &mh, 64);
    return QSV_OK;
}

int qsv_comm_set_shadow_peers(qsv_handle *h, const void *handles) {
    QSV_CHECK_H(h);
    auto *c = (qsvx::Comm *)h->comm;
    if (!handles || !c || !c->peers_ready) QSV_FAIL(h, QSV_ECOMM, "set_shadow_peers: call qsv_comm_set_peers first");
    if (h->scat_ready) QSV_FAIL(h, QSV_EINVAL, "set_shadow_peers: targets are already wired");
    int rc = qsv_shadow_ptr(h, nullptr);
    if (rc) return rc;
    std::vector<void *> other(h->world, nullptr);
    for (int r = 0; r < h->world; ++r) {
        if (r == h->rank) { other[r] = h->d_shadow; continue; }
        cudaIpcMemHandle_t mh;
        memcpy(&mh, (const char *)handles + 64 * r, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&other[r], mh, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < r; ++q) if (q != h->rank && other[q]) cudaIpcCloseMemHandle(other[q]);
            QSV_FAIL(h, QSV_ECOMM, "cudaIpcOpenMemHandle(shadow of rank %d): %s", r, cudaGetErrorString(e));
        }
    }
    h->scat_cur = c->peer;
    h->scat_other = other;
    h->scat_ready = h->scat_ipc = true;
    return QSV_OK;
}

int qsv_scatter_set_targets(qsv_handle *h, void *const *current, void *const *shadow) {
    QSV_CHECK_H(h);
    if (!current || !shadow) QSV_FAIL(h, QSV_EINVAL, "scatter_set_targets: null tables");
    if (h->scat_ipc) QSV_FAIL(h, QSV_EINVAL, "scatter_set_targets: targets are already wired through CUDA IPC");
    int rc = qsv_shadow_ptr(h, nullptr);
    if (rc) return rc;
    if (current[h->rank] != h->d_state || shadow[h->rank] != h->d_shadow)
        QSV_FAIL(h, QSV_EINVAL, "scatter_set_targets: entry %d must be this handle's own buffers", h->rank);
    h->scat_cur.assign(current, current + h->world);
    h->scat_other.assign(shadow, shadow + h->world);
    h->scat_ready = true;
    return QSV_OK;
}

static bool scatter_bits_ok(qsv_handle *h, int n_swap, const int *local_bits) {
    if (n_swap < 1 || n_swap > 3 || !local_bits) return false;
    for (int i = 0; i < n_swap; ++i) {
        if (local_bits[i] < 0 || local_bits[i] >= h->n_local) return false;
        for (int j = 0; j < i; ++j) if (local_bits[i] == local_bits[j]) return false;
    }
    return true;
}

// kernel of (pass, swapped local bits): program cache, then the cubin caches / NVRTC of jit.cuh
static cudaKernel_t scatter_kernel(qsv_handle *h, qsv_program *p, int i, int n_swap, const int *local_bits) {
    for (auto &k : p->scatter)
        if (k.pass_index == i && k.n == n_swap && !memcmp(k.bits, local_bits, sizeof(int) * n_swap)) return k.fn;
    if (!h->jit || !qsvjit::enabled() || h->force_simple_pass || !p->jit[i]) return nullptr;
    qsvjit::Scatter sc;
    sc.n = n_swap;
    for (int k = 0; k < n_swap; ++k) sc.local_bits[k] = local_bits[k];
    std::vector<std::string> srcs(1);
    std::vector<double> coefs;                      // same op order as the plain kernel: p->jit_coefs[i] is reused
    const qsv_op *ops = p->h_ops.empty() ? nullptr : p->h_ops.data() + p->op_offset[i];
    if (!qsvjit::generate(p->passes[i], ops, srcs[0], coefs, h->dtype == QSV_C64, &sc)) return nullptr;
    std::vector<qsvjit::Kernel> ks;
    std::string err;
    qsvjit::resolve(h->device, srcs, ks, err);
    if (!err.empty()) h->err = "jit (scatter pass): " + err;
    if (ks[0].key) p->jit_keys.push_back(ks[0].key);
    qsv_program::ScatterKernel e{i, n_swap, {0, 0, 0}, ks[0].fn};
    for (int k = 0; k < n_swap; ++k) e.bits[k] = local_bits[k];
    p->scatter.push_back(e);                        // a failed build is remembered as well (fn == nullptr)
    return ks[0].fn;
}

int qsv_pass_scatter_prepare(qsv_handle *h, qsv_program *p, int pass_index, int n_swap, const int *local_bits) {
    QSV_CHECK_H(h);
    if (!p || pass_index < 0 || (size_t)pass_index >= p->passes.size() || !scatter_bits_ok(h, n_swap, local_bits))
        QSV_FAIL(h, QSV_EINVAL, "pass_scatter_prepare: bad arguments");
    QSV_CUDA(h, cudaSetDevice(h->device));
    return scatter_kernel(h, p, pass_index, n_swap, local_bits) ? QSV_OK : QSV_EINVAL;
}

int qsv_pass_scatter(qsv_handle *h, qsv_program *p, int pass_index, int n_swap, const int *global_bits,
                     const int *local_bits, int *fused) {
    QSV_CHECK_H(h);
    if (fused) *fused = 0;
    if (!p || pass_index < 0 || (size_t)pass_index >= p->passes.size()) QSV_FAIL(h, QSV_EINVAL, "pass_scatter: bad pass");
    const int g = h->n_qubits - h->n_local;
    if (!scatter_bits_ok(h, n_swap, local_bits) || !global_bits || n_swap > g) QSV_FAIL(h, QSV_EINVAL, "pass_scatter: bad swap bits");
    for (int i = 0; i < n_swap; ++i) {
        if (global_bits[i] < h->n_local || global_bits[i] >= h->n_qubits) QSV_FAIL(h, QSV_EINVAL, "pass_scatter: global bit %d is not a rank bit", global_bits[i]);
        for (int j = 0; j < i; ++j) if (global_bits[i] == global_bits[j]) QSV_FAIL(h, QSV_EINVAL, "pass_scatter: repeated bit");
    }
    QSV_CUDA(h, cudaSetDevice(h->device));
    auto *c = (qsvx::Comm *)h->comm;
    cudaKernel_t fn = nullptr;
    if (h->scat_ready && p->passes[pass_index].n_active < 0) fn = scatter_kernel(h, p, pass_index, n_swap, local_bits);
    if (!fn) {                                       // every rank takes the same decision: same program, same wiring
        int rc = qsv_program_run_range(h, p, pass_index, 1);
        if (rc) return rc;
        return qsv_swap_global_local(h, n_swap, global_bits, local_bits);
    }
    struct { void *p[8]; unsigned long long keep; } dst;
    dst.keep = 0;
    for (int i = 0; i < n_swap; ++i)
        dst.keep |= (unsigned long long)((h->rank >> (global_bits[i] - h->n_local)) & 1) << local_bits[i];
    for (int x = 0; x < 8; ++x) {
        dst.p[x] = nullptr;
        if (x >= (1 << n_swap)) continue;
        int r = h->rank;
        for (int i = 0; i < n_swap; ++i) {
            const int rb = global_bits[i] - h->n_local;
            r = (r & ~(1 << rb)) | (((x >> i) & 1) << rb);
        }
        dst.p[x] = h->scat_other[r];
        if (!dst.p[x]) QSV_FAIL(h, QSV_ECOMM, "pass_scatter: the second buffer of rank %d is not mapped", r);
    }
    {
        ScopedTimer t(h, 40 + n_swap, pass_index);
        const uint32_t tiles = (uint32_t)(h->n_amps >> qsvjit::kT);
        const unsigned grid = tiles < (uint32_t)h->sm_count ? tiles : (unsigned)h->sm_count;
        void *state = h->d_state;
        const double2 *tables = p->d_tables + p->fold_offset[pass_index];
        unsigned long long rank_bits = (unsigned long long)h->rank << h->n_local;
        unsigned tb = 0, te = tiles;
        static const double zero = 0.0;
        void *coefs = p->jit_coefs[pass_index].empty() ? (void *)&zero
                    : (h->dtype == QSV_C64 ? (void *)p->jit_coefs_f[pass_index].data() : (void *)p->jit_coefs[pass_index].data());
        JitFixArg fix = {};
        void *args[] = {&state, &tables, &rank_bits, &tb, &te, coefs, &fix, &dst};
        QSV_CUDA(h, cudaLaunchKernel((const void *)fn, dim3(grid), dim3(128 * (qsvjit::groups() + 1)), args, qsvjit::kSmemBytes, h->stream));
        if (c && c->comm) {                          // all stores of all ranks have landed before anyone reads
            int rc = swap_barrier(h, c);
            if (rc) return rc;
        }
    }
    std::swap(h->d_state, h->d_shadow);
    std::swap(h->scat_cur, h->scat_other);
    if (c && c->peers_ready) c->peer = h->scat_cur;
    if (fused) *fused = 1;
    return QSV_OK;
}

int qsv_jit_source_scatter(const qsv_pass *pass, const qsv_op *ops, int dtype, int n_swap, const int *local_bits,
                           char *out, size_t cap, size_t *needed) {
    if (!pass || (pass->n_ops > 0 && !ops) || n_swap < 1 || n_swap > 3 || !local_bits) return QSV_EINVAL;
    qsvjit::Scatter sc;
    sc.n = n_swap;
    for (int k = 0; k < n_swap; ++k) sc.local_bits[k] = local_bits[k];
    std::string src;
    std::vector<double> coefs;
    if (!qsvjit::generate(*pass, ops, src, coefs, dtype == QSV_C64, &sc)) return QSV_EINVAL;
    if (needed) *needed = src.size() + 1;
    if (out && cap) snprintf(out, cap, "%s", src.c_str());
    return QSV_OK;
}

int qsv_jit_build_scatter(const qsv_pass *pass, const qsv_op *ops, int dtype, int n_swap, const int *local_bits,
                          size_t *cubin_bytes, char *log, size_t log_cap) {
    if (!pass || (pass->n_ops > 0 && !ops) || n_swap < 1 || n_swap > 3 || !local_bits) return QSV_EINVAL;
    qsvjit::Scatter sc;
    sc.n = n_swap;
    for (int k = 0; k < n_swap; ++k) sc.local_bits[k] = local_bits[k];
    std::string src, msg;
    std::vector<double> coefs;
    std::vector<char> cubin;
    int rc = QSV_OK;
    if (!qsvjit::generate(*pass, ops, src, coefs, dtype == QSV_C64, &sc)) { msg = "pass is not eligible for a scatter kernel"; rc = QSV_EINVAL; }
    else if (!qsvjit::nvrtc().load()) { msg = "NVRTC unavailable: " + qsvjit::nvrtc().why; rc = QSV_EIO; }
    else if (!qsvjit::compile(src, cubin, msg)) rc = QSV_ECUDA;
    if (cubin_bytes) *cubin_bytes = cubin.size();
    if (log && log_cap) { snprintf(log, log_cap, "%s", msg.c_str()); }
    return rc;
}

int qsv_set_option(qsv_handle *h, int option, long long value) {
    QSV_CHECK_H(h);
    switch (option) {
        case QSV_OPT_JIT: h->jit = value != 0; return QSV_OK;
        case QSV_OPT_SIMPLE_PASS: h->force_simple_pass = value != 0; return QSV_OK;
        case QSV_OPT_PEER_SWAP: h->use_peer_swap = value != 0; return QSV_OK;
        default: QSV_FAIL(h, QSV_EINVAL, "set_option: unknown option %d", option);
    }
}

int qsv_jit_build_pass(const qsv_pass *pass, const qsv_op *ops, int dtype, size_t *cubin_bytes, char *log, size_t log_cap) {
    if (!pass || (pass->n_ops > 0 && !ops)) return QSV_EINVAL;
    std::string src, msg;
    std::vector<double> coefs;
    std::vector<char> cubin;
    int rc = QSV_OK;
    if (!qsvjit::generate(*pass, ops, src, coefs, dtype == QSV_C64)) { msg = "pass is not eligible for specialisation"; rc = QSV_EINVAL; }
    else if (!qsvjit::nvrtc().load()) { msg = "NVRTC unavailable: " + qsvjit::nvrtc().why; rc = QSV_EIO; }
    else if (!qsvjit::compile(src, cubin, msg)) rc = QSV_ECUDA;
    if (rc == QSV_OK && getenv("QSV_JIT_WARM")) {            // build-time cache warm-up (no device needed)
        const std::string dir = qsvjit::cache_dir();
        if (!dir.empty()) {
            mkdir(dir.c_str(), 0755);
            char name[64];
            snprintf(name, sizeof(name), "/%016llx.cubin", (unsigned long long)qsvjit::source_key(src));
            qsvjit::write_file_atomic(dir + name, qsvjit::cache_pack(src, cubin));
        }
    }
    if (cubin_bytes) *cubin_bytes = cubin.size();
    if (log && log_cap) { snprintf(log, log_cap, "%s", msg.c_str()); }
    return rc;
}

int qsv_jit_source(const qsv_pass *pass, const qsv_op *ops, int dtype, char *out, size_t cap, size_t *needed) {
    if (!pass || (pass->n_ops > 0 && !ops)) return QSV_EINVAL;
    std::string src;
    std::vector<double> coefs;
    if (!qsvjit::generate(*pass, ops, src, coefs, dtype == QSV_C64)) return QSV_EINVAL;
    if (needed) *needed = src.size() + 1;
    if (out && cap) snprintf(out, cap, "%s", src.c_str());
    return QSV_OK;
}

int qsv_jit_coefs(const qsv_pass *pass, const qsv_op *ops, int dtype, double *out, size_t cap, size_t *needed) {
    if (!pass || (pass->n_ops > 0 && !ops)) return QSV_EINVAL;
    std::string src;
    std::vector<double> coefs;
    if (!qsvjit::generate(*pass, ops, src, coefs, dtype == QSV_C64)) return QSV_EINVAL;
    if (needed) *needed = coefs.size();
    if (out) for (size_t i = 0; i < coefs.size() && i < cap; ++i) out[i] = coefs[i];
    return QSV_OK;
}

int qsv_jit_stats(int *compiled, int *disk_hits, int *mem_hits, int *failed, double *compile_seconds) {
    const qsvjit::Stats &s = qsvjit::stats();
    if (compiled) *compiled = s.compiled;
    if (disk_hits) *disk_hits = s.disk_hits;
    if (mem_hits) *mem_hits = s.mem_hits;
    if (failed) *failed = s.failed;
    if (compile_seconds) *compile_seconds = s.compile_s;
    return QSV_OK;
}

int qsv_program_destroy(qsv_handle *h, qsv_program *p) {
    QSV_CHECK_H(h);
    if (!p) return QSV_OK;
    cudaSetDevice(h->device);
    if (p->graph) cudaGraphExecDestroy(p->graph);
    dev_free(h, p->d_passes); dev_free(h, p->d_ops); dev_free(h, p->d_tables);   // stream-ordered after the last run
    if (!p->jit_keys.empty()) {
        cudaStreamSynchronize(h->stream);                 // its kernels may be unloaded once nobody holds them
        qsvjit::release(p->jit_keys);
    }
    delete p;
    return QSV_OK;
}

// ----------------------------------------------------------------- reductions ----
int qsv_norm2(qsv_handle *h, double *out) {
    QSV_CHECK_H(h);
    if (!out) QSV_FAIL(h, QSV_EINVAL, "norm2: null out");
    QSV_CUDA(h, cudaSetDevice(h->device));
    if (h->dtype == QSV_C128) k_norm2_partial<double><<<kNormBlocks, 256, 0, h->stream>>>((const double2 *)h->d_state, h->n_amps, h->d_partials);
    else k_norm2_partial<float><<<kNormBlocks, 256, 0, h->stream>>>((const float2 *)h->d_state, h->n_amps, h->d_partials);
    k_sum_partials<<<1, 256, 0, h->stream>>>(h->d_partials, kNormBlocks, h->d_partials + kNormBlocks);
    QSV_CUDA(h, cudaGetLastError());
    QSV_CUDA(h, cudaMemcpyAsync(out, h->d_partials + kNormBlocks, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    QSV_CUDA(h, cudaStreamSynchronize(h->stream));
    return QSV_OK;
}

// ------------------------------------------------------------------- sampling ----
int qsv_leaf_sums(qsv_handle *h, double *out_host) {
    QSV_CHECK_H(h);
    if (!out_host) QSV_FAIL(h, QSV_EINVAL, "leaf_sums: null out");
    QSV_CUDA(h, cudaSetDevice(h->device));
    const int leaf_log2 = h->n_local < kSampleLeafLog2 ? h->n_local : kSampleLeafLog2;
    const int leaf = 1 << leaf_log2;
    const uint64_t n_leaves = h->n_amps >> leaf_log2;
    double *d = nullptr;
    QSV_CUDA(h, dev_alloc(h, (void **)&d, n_leaves * sizeof(double)));
    const unsigned grid = (unsigned)((n_leaves + 127) / 128);
    if (h->dtype == QSV_C128) k_leaf_sums<double><<<grid, 128, 0, h->stream>>>((const double2 *)h->d_state, n_leaves, leaf, d);
    else k_leaf_sums<float><<<grid, 128, 0, h->stream>>>((const float2 *)h->d_state, n_leaves, leaf, d);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d, n_leaves * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    dev_free(h, d);
    QSV_CUDA(h, e);
    return QSV_OK;
}

int qsv_sample_in_leaves(qsv_handle *h, int shots, const uint64_t *leaf_idx, const double *leaf_off, const double *x,
                         uint64_t *out_local_index) {
    QSV_CHECK_H(h);
    if (shots < 0 || (shots && (!leaf_idx || !leaf_off || !x || !out_local_index))) QSV_FAIL(h, QSV_EINVAL, "sample_in_leaves: bad arguments");
    if (shots == 0) return QSV_OK;
    QSV_CUDA(h, cudaSetDevice(h->device));
    const int leaf_log2 = h->n_local < kSampleLeafLog2 ? h->n_local : kSampleLeafLog2;
    const uint64_t n_leaves = h->n_amps >> leaf_log2;
    for (int i = 0; i < shots; ++i) if (leaf_idx[i] >= n_leaves) QSV_FAIL(h, QSV_EINVAL, "sample_in_leaves: leaf %llu outside the shard", (unsigned long long)leaf_idx[i]);
    char *d = nullptr;
    const size_t n8 = (size_t)shots * 8;
    QSV_CUDA(h, dev_alloc(h, (void **)&d, 4 * n8));
    cudaError_t e = cudaMemcpyAsync(d, leaf_idx, n8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + n8, leaf_off, n8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + 2 * n8, x, n8, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) {
        const unsigned grid = (unsigned)((shots + 127) / 128);
        if (h->dtype == QSV_C128)
            k_sample_walk<double><<<grid, 128, 0, h->stream>>>((const double2 *)h->d_state, 1 << leaf_log2, shots, (const uint64_t *)d,
                                                              (const double *)(d + n8), (const double *)(d + 2 * n8), (uint64_t *)(d + 3 * n8));
        else
            k_sample_walk<float><<<grid, 128, 0, h->stream>>>((const float2 *)h->d_state, 1 << leaf_log2, shots, (const uint64_t *)d,
                                                             (const double *)(d + n8), (const double *)(d + 2 * n8), (uint64_t *)(d + 3 * n8));
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_local_index, d + 3 * n8, n8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    dev_free(h, d);
    QSV_CUDA(h, e);
    return QSV_OK;
}

int qsv_sample(qsv_handle *h, uint64_t, int shots, const double *sorted_u, uint64_t *out_indices) {
    QSV_CHECK_H(h);
    if (h->world != 1) QSV_FAIL(h, QSV_EINVAL, "qsv_sample serves one shard; chain qsv_leaf_sums / qsv_sample_in_leaves across ranks");
    if (shots < 0 || (shots && (!sorted_u || !out_indices))) QSV_FAIL(h, QSV_EINVAL, "sample: bad arguments");
    const int leaf_log2 = h->n_local < kSampleLeafLog2 ? h->n_local : kSampleLeafLog2;
    const uint64_t n_leaves = h->n_amps >> leaf_log2;
    std::vector<double> sums(n_leaves), offs(n_leaves + 1);
    int rc = qsv_leaf_sums(h, sums.data());
    if (rc) return rc;
    offs[0] = 0.0;
    for (uint64_t b = 0; b < n_leaves; ++b) offs[b + 1] = offs[b] + sums[b];     // sequential scan (the definition)
    const double total = offs[n_leaves];
    std::vector<uint64_t> lidx;
    std::vector<double> loff, xs;
    std::vector<int> where;
    for (int s = 0; s < shots; ++s) {
        const double x = sorted_u[s] * total;
        // first b with offs[b+1] > x  (np.searchsorted(offs[1:], x, side="right"))
        const uint64_t b = (uint64_t)(std::upper_bound(offs.begin() + 1, offs.end(), x) - (offs.begin() + 1));
        if (b >= n_leaves) { out_indices[s] = h->n_amps - 1; continue; }
        lidx.push_back(b); loff.push_back(offs[b]); xs.push_back(x); where.push_back(s);
    }
    std::vector<uint64_t> got(lidx.size());
    rc = qsv_sample_in_leaves(h, (int)lidx.size(), lidx.data(), loff.data(), xs.data(), got.data());
    if (rc) return rc;
    for (size_t i = 0; i < got.size(); ++i) out_indices[where[i]] = got[i];
    return QSV_OK;
}

// ---------------------------------------------------------------- observables ----
int qsv_probabilities(qsv_handle *h, int nq, const int *qubits, double *out_host) {
    QSV_CHECK_H(h);
    if (nq < 0 || nq > 20 || (nq && !qubits) || !out_host) QSV_FAIL(h, QSV_EINVAL, "probabilities: need 0..20 qubits");
    MarginalArgs a;
    memset(&a, 0, sizeof(a));
    a.nq = nq;
    for (int i = 0; i < nq; ++i) {
        if (qubits[i] < 0 || qubits[i] >= h->n_qubits) QSV_FAIL(h, QSV_EINVAL, "probabilities: qubit %d out of range", qubits[i]);
        for (int j = 0; j < i; ++j) if (qubits[i] == qubits[j]) QSV_FAIL(h, QSV_EINVAL, "probabilities: repeated qubit %d", qubits[i]);
        a.qs[i] = qubits[i];
    }
    QSV_CUDA(h, cudaSetDevice(h->device));
    const size_t bins = (size_t)1 << nq;
    double *d = nullptr;
    QSV_CUDA(h, dev_alloc(h, (void **)&d, bins * sizeof(double)));
    cudaError_t e = cudaMemsetAsync(d, 0, bins * sizeof(double), h->stream);
    if (e == cudaSuccess) {
        const size_t smem = nq <= 10 ? bins * sizeof(double) : 0;
        const uint64_t rank_bits = (uint64_t)h->rank << h->n_local;
        const int grid = grid_for(h->n_amps, 256, h->sm_count * 8);
        if (h->dtype == QSV_C128) k_marginal<double><<<grid, 256, smem, h->stream>>>((const double2 *)h->d_state, h->n_amps, rank_bits, a, d);
        else k_marginal<float><<<grid, 256, smem, h->stream>>>((const float2 *)h->d_state, h->n_amps, rank_bits, a, d);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_host, d, bins * sizeof(double), cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    dev_free(h, d);
    QSV_CUDA(h, e);
    return QSV_OK;
}

int qsv_expect_z(qsv_handle *h, uint64_t mask, double *out) {
    QSV_CHECK_H(h);
    if (!out) QSV_FAIL(h, QSV_EINVAL, "expect_z: null out");
    if (h->n_qubits < 64 && (mask >> h->n_qubits)) QSV_FAIL(h, QSV_EINVAL, "expect_z: mask names a qubit >= %d", h->n_qubits);
    QSV_CUDA(h, cudaSetDevice(h->device));
    const uint64_t rank_bits = (uint64_t)h->rank << h->n_local;
    if (h->dtype == QSV_C128) k_expect_z_partial<double><<<kNormBlocks, 256, 0, h->stream>>>((const double2 *)h->d_state, h->n_amps, rank_bits, mask, h->d_partials);
    else k_expect_z_partial<float><<<kNormBlocks, 256, 0, h->stream>>>((const float2 *)h->d_state, h->n_amps, rank_bits, mask, h->d_partials);
    k_sum_partials<<<1, 256, 0, h->stream>>>(h->d_partials, kNormBlocks, h->d_partials + kNormBlocks);
    QSV_CUDA(h, cudaGetLastError());
    QSV_CUDA(h, cudaMemcpyAsync(out, h->d_partials + kNormBlocks, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    QSV_CUDA(h, cudaStreamSynchronize(h->stream));
    return QSV_OK;
}

// --------------------------------------------------------------------- timing ----
int qsv_timing_enable(qsv_handle *h, int on) {
    QSV_CHECK_H(h);
    h->timing = on != 0;
    if (!on) {
        for (auto &t : h->timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
        h->timed.clear();
    }
    return QSV_OK;
}

int qsv_get_timings(qsv_handle *h, qsv_timing *out, int max, int *n_out) {
    QSV_CHECK_H(h);
    QSV_CUDA(h, cudaSetDevice(h->device));
    QSV_CUDA(h, cudaStreamSynchronize(h->stream));
    int n = 0;
    for (auto &t : h->timed) {
        if (out && n < max) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, t.a, t.b);
            out[n].ms = ms; out[n].kind = t.kind; out[n].pass_index = t.pass_index;
        }
        ++n;
        cudaEventDestroy(t.a); cudaEventDestroy(t.b);
    }
    h->timed.clear();
    if (n_out) *n_out = n;
    return QSV_OK;
}

int qsv_timer_start(qsv_handle *h) {
    QSV_CHECK_H(h);
    QSV_CUDA(h, cudaSetDevice(h->device));
    if (!h->t0) { QSV_CUDA(h, cudaEventCreate(&h->t0)); QSV_CUDA(h, cudaEventCreate(&h->t1)); }
    QSV_CUDA(h, cudaEventRecord(h->t0, h->stream));
    return QSV_OK;
}

int qsv_timer_stop(qsv_handle *h, float *elapsed_ms) {
    QSV_CHECK_H(h);
    if (!h->t0 || !elapsed_ms) QSV_FAIL(h, QSV_EINVAL, "timer_stop without timer_start");
    QSV_CUDA(h, cudaSetDevice(h->device));
    QSV_CUDA(h, cudaEventRecord(h->t1, h->stream));
    QSV_CUDA(h, cudaEventSynchronize(h->t1));
    QSV_CUDA(h, cudaEventElapsedTime(elapsed_ms, h->t0, h->t1));
    return QSV_OK;
}

}  // extern "C"
