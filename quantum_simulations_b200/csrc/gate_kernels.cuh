// gate_kernels.cuh — one-gate-per-sweep streaming kernels (the a1/a2 operator face).
//
// These are the direct sm_100a restatement of the reference's local kernels
//   apply_1q   wenbo_engine/kernel/cpu_scalar.py:21-32   (pairs at stride 2^q)
//   apply_2q   wenbo_engine/kernel/cpu_scalar.py:35-47   (quads, row = 2*bit(qa)+bit(qb))
// and of the butterfly variants in cpu_nonlocal.py:22-67 (on a GPU a "non-local" qubit that
// is still on the device is just a larger stride).  Every amplitude is read once and written
// once: algorithmic traffic = 2 * sizeof(amp) * 2^n per launch.  The fused path
// (pass_kernel.cuh) is what the runner uses; these exist for the per-gate ABI and as an
// independent cross-check of the fused path in the GPU tests.
#pragma once
#include "common.cuh"

constexpr int kGateThreads = 256;

// ---- 1 qubit: each thread updates kU independent pairs --------------------------------
template <typename R, int kU>
__global__ void __launch_bounds__(kGateThreads)
k_apply_1q(typename CxT<R>::V *__restrict__ s, uint64_t n_pairs, int q, Mat2<R> U) {
    using V = typename CxT<R>::V;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t bit = 1ull << q;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_pairs; p += stride * kU) {
        V a[kU], b[kU];
        uint64_t i0[kU];
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            const uint64_t pp = p + u * stride;
            i0[u] = insert_zero_bit(pp < n_pairs ? pp : p, q);
            a[u] = s[i0[u]];
            b[u] = s[i0[u] | bit];
        }
#pragma unroll
        for (int u = 0; u < kU; ++u) {
            if (p + u * stride >= n_pairs) break;
            V na = cx_fma(U.m[1], b[u], cx_mul(U.m[0], a[u]));
            V nb = cx_fma(U.m[3], b[u], cx_mul(U.m[2], a[u]));
            s[i0[u]] = na;
            s[i0[u] | bit] = nb;
        }
    }
}

// ---- controlled 1 qubit (ctrl local): only the ctrl=1 half is touched ------------------
template <typename R>
__global__ void __launch_bounds__(kGateThreads)
k_apply_ctrl_1q(typename CxT<R>::V *__restrict__ s, uint64_t n_quads, int ctrl, int tgt, Mat2<R> U) {
    using V = typename CxT<R>::V;
    const int lo = ctrl < tgt ? ctrl : tgt, hi = ctrl < tgt ? tgt : ctrl;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_quads; p += stride) {
        const uint64_t i0 = insert_zero_bit(insert_zero_bit(p, lo), hi) | (1ull << ctrl);
        const uint64_t i1 = i0 | (1ull << tgt);
        V a = s[i0], b = s[i1];
        s[i0] = cx_fma(U.m[1], b, cx_mul(U.m[0], a));
        s[i1] = cx_fma(U.m[3], b, cx_mul(U.m[2], a));
    }
}

// ---- 2 qubits: dense 4x4 on quads -----------------------------------------------------
template <typename R>
__global__ void __launch_bounds__(kGateThreads)
k_apply_2q(typename CxT<R>::V *__restrict__ s, uint64_t n_quads, int qa, int qb, Mat4<R> U) {
    using V = typename CxT<R>::V;
    const int lo = qa < qb ? qa : qb, hi = qa < qb ? qb : qa;
    const uint64_t ma = 1ull << qa, mb = 1ull << qb;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_quads; p += stride) {
        const uint64_t base = insert_zero_bit(insert_zero_bit(p, lo), hi);
        const uint64_t idx[4] = {base, base | mb, base | ma, base | ma | mb};
        V v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) v[k] = s[idx[k]];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            V acc = cx_mul(U.m[4 * r], v[0]);
#pragma unroll
            for (int c = 1; c < 4; ++c) acc = cx_fma(U.m[4 * r + c], v[c], acc);
            s[idx[r]] = acc;
        }
    }
}

// ---- diagonal gate on up to 6 qubits (any of them may be rank bits) -------------------
template <typename R> struct DiagArgs {
    int nq;
    int qs[6];                       // physical bits, qs[0] = most significant table bit
    typename CxT<R>::V phase[64];
};

template <typename R>
__global__ void __launch_bounds__(kGateThreads)
k_apply_diag(typename CxT<R>::V *__restrict__ s, uint64_t n_amps, uint64_t rank_offset,
             const DiagArgs<R> *__restrict__ args) {
    using V = typename CxT<R>::V;
    __shared__ DiagArgs<R> A;
    for (int i = threadIdx.x; i < (int)(sizeof(A) / 4); i += blockDim.x)
        reinterpret_cast<uint32_t *>(&A)[i] = reinterpret_cast<const uint32_t *>(args)[i];
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
        const uint64_t g = rank_offset | i;
        int t = 0;
#pragma unroll 1
        for (int k = 0; k < A.nq; ++k) t = (t << 1) | (int)((g >> A.qs[k]) & 1ull);
        s[i] = cx_mul(A.phase[t], s[i]);
    }
}

// ---- dense k-qubit unitary (k <= 5), matrix in shared memory ---------------------------
template <typename R, int K>
__global__ void __launch_bounds__(128)
k_apply_kq(typename CxT<R>::V *__restrict__ s, uint64_t n_groups,
           const int *__restrict__ qs_sorted_and_rowbit,   // [0..K) sorted ascending, [K..2K) row bit of each
           const typename CxT<R>::V *__restrict__ U) {
    using V = typename CxT<R>::V;
    constexpr int D = 1 << K;
    extern __shared__ unsigned char smem_raw[];
    V *sU = reinterpret_cast<V *>(smem_raw);
    __shared__ uint64_t off[D];
    __shared__ int sq[K];
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) sU[i] = U[i];
    if (threadIdx.x < K) sq[threadIdx.x] = qs_sorted_and_rowbit[threadIdx.x];
    if (threadIdx.x < D) {
        // offset of sub-space row r: bit (rowbit[j]) of r selects physical bit qs_sorted[j]
        uint64_t o = 0;
        for (int j = 0; j < K; ++j)
            if ((threadIdx.x >> qs_sorted_and_rowbit[K + j]) & 1) o |= 1ull << qs_sorted_and_rowbit[j];
        off[threadIdx.x] = o;
    }
    __syncthreads();
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n_groups; p += stride) {
        uint64_t base = p;
#pragma unroll
        for (int j = 0; j < K; ++j) base = insert_zero_bit(base, sq[j]);
        V v[D];
#pragma unroll
        for (int c = 0; c < D; ++c) v[c] = s[base | off[c]];
#pragma unroll 2
        for (int r = 0; r < D; ++r) {
            V acc = cx_mul(sU[r * D], v[0]);
#pragma unroll
            for (int c = 1; c < D; ++c) acc = cx_fma(sU[r * D + c], v[c], acc);
            s[base | off[r]] = acc;
        }
    }
}

// ---- init / reductions -----------------------------------------------------------------
template <typename R>
__global__ void k_fill_zero(typename CxT<R>::V *__restrict__ s, uint64_t n_amps) {
    using V = typename CxT<R>::V;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride)
        s[i] = cx_make<V>(0, 0);
}

template <typename R>
__global__ void k_set_amp(typename CxT<R>::V *s, uint64_t idx, R re, R im) {
    s[idx] = cx_make<typename CxT<R>::V>(re, im);
}

// deterministic two-stage sum of |amp|^2: fixed grid, fixed per-thread order, fixed tree.
constexpr int kNormBlocks = 1184;  // 148 SMs x 8
template <typename R>
__global__ void __launch_bounds__(256)
k_norm2_partial(const typename CxT<R>::V *__restrict__ s, uint64_t n_amps, double *__restrict__ partial) {
    __shared__ double sh[256];
    double acc = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
        const auto v = s[i];
        acc += (double)v.x * (double)v.x + (double)v.y * (double)v.y;
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void k_sum_partials(const double *__restrict__ partial, int n, double *out) {
    __shared__ double sh[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) acc += partial[i];
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0];
}
