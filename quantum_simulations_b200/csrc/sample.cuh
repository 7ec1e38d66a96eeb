// sample.cuh — deterministic measurement sampling on a shard.
//
// The reference has no sampler ("parity unpinned"); the definition is
// frozen in oracle/ref_dense.py::sample_indices and restated here operation for operation so that
// the sample indices are BIT-EXACT with the oracle under a fixed seed:
//   p_i      = re*re + im*im   in float64, each product and the sum rounded separately (no FMA)
//   leaf_b   = sequential left-to-right sum of the 1024 p_i of block b
//   offs     = sequential exclusive scan of the leaf sums (on the host: one chain over all shards)
//   sample s = first index i with (offs[b] + running in-leaf sum) > u_s * total, clamped to the end
#pragma once
#include "common.cuh"

constexpr int kSampleLeafLog2 = 10;

template <typename R>
__device__ __forceinline__ double prob_of(const typename CxT<R>::V v) {
    const double re = (double)v.x, im = (double)v.y;
    return __dadd_rn(__dmul_rn(re, re), __dmul_rn(im, im));
}

// one thread per leaf: the sequential order is part of the definition
template <typename R>
__global__ void __launch_bounds__(128)
k_leaf_sums(const typename CxT<R>::V *__restrict__ s, uint64_t n_leaves, int leaf, double *__restrict__ out) {
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n_leaves) return;
    const typename CxT<R>::V *p = s + b * (uint64_t)leaf;
    double acc = 0.0;
    for (int j = 0; j < leaf; ++j) acc = __dadd_rn(acc, prob_of<R>(p[j]));
    out[b] = acc;
}

// one thread per shot: walk leaf leaf_idx[s] from its exclusive prefix leaf_off[s]
template <typename R>
__global__ void __launch_bounds__(128)
k_sample_walk(const typename CxT<R>::V *__restrict__ s, int leaf, int shots, const uint64_t *__restrict__ leaf_idx,
              const double *__restrict__ leaf_off, const double *__restrict__ x, uint64_t *__restrict__ out) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= shots) return;
    const uint64_t b = leaf_idx[t];
    const typename CxT<R>::V *p = s + b * (uint64_t)leaf;
    double run = leaf_off[t];
    const double target = x[t];
    int hit = leaf - 1;
    for (int j = 0; j < leaf; ++j) {
        run = __dadd_rn(run, prob_of<R>(p[j]));
        if (run > target) { hit = j; break; }
    }
    out[t] = b * (uint64_t)leaf + (uint64_t)hit;
}

// ---- observables after the path ------------------------------------------
// Marginal distribution over nq qubits: out[o] += |amp_i|^2 with bit k of o = index bit qs[k]
// (rank bits allowed: constant per shard).  Shared-memory histogram for nq <= 10, global
// atomics beyond; the caller sums the shards' contributions.
struct MarginalArgs { int nq; int qs[20]; };

template <typename R>
__global__ void __launch_bounds__(256)
k_marginal(const typename CxT<R>::V *__restrict__ s, uint64_t n_amps, uint64_t rank_bits, MarginalArgs a,
           double *__restrict__ out) {
    extern __shared__ double hist[];
    const int bins = 1 << a.nq;
    const bool use_smem = a.nq <= 10;
    if (use_smem) {
        for (int i = threadIdx.x; i < bins; i += blockDim.x) hist[i] = 0.0;
        __syncthreads();
    }
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
        const double p = prob_of<R>(s[i]);
        const uint64_t idx = rank_bits | i;
        int o = 0;
        for (int k = 0; k < a.nq; ++k) o |= (int)((idx >> a.qs[k]) & 1ull) << k;
        if (p != 0.0) atomicAdd(use_smem ? &hist[o] : &out[o], p);
    }
    if (use_smem) {
        __syncthreads();
        for (int i = threadIdx.x; i < bins; i += blockDim.x) if (hist[i] != 0.0) atomicAdd(&out[i], hist[i]);
    }
}

// <Z_mask> contribution of the shard: sum_i |amp_i|^2 (-1)^parity(index & mask); fixed reduction tree
template <typename R>
__global__ void __launch_bounds__(256)
k_expect_z_partial(const typename CxT<R>::V *__restrict__ s, uint64_t n_amps, uint64_t rank_bits, uint64_t mask,
                   double *__restrict__ partial) {
    __shared__ double sh[256];
    double acc = 0.0;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_amps; i += stride) {
        const double p = prob_of<R>(s[i]);
        acc += (__popcll((rank_bits | i) & mask) & 1) ? -p : p;
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) sh[threadIdx.x] += sh[threadIdx.x + w];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
