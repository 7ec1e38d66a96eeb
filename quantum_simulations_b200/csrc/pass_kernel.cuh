// pass_kernel.cuh — the fused gate-application pass: ONE read + ONE write of the shard
// applies a whole list of gates (what batch_levels / _process_local_chunk do per chunk
// file in the reference, wenbo_engine/circuit/fusion.py:86-142 and
// runner/single_node.py:208-216, but with the "chunk" living on chip).
//
// Data layout.  A CTA owns a TILE of 2^T amplitudes: the amplitudes whose T tile bits
// (physical index bits load_bits[0..T), ascending, the lowest ones contiguous) vary while all
// other index bits are fixed by blockIdx.  Each thread keeps 2^4 = 16 amplitudes in
// registers.  A pass is a sequence of ROUNDS; round r names 4 tile positions (reg_pos) that
// vary inside a thread, the remaining T-4 positions are spread over the 2^(T-4) threads
// (thr_pos).  Gates whose target is one of the 4 register positions are applied entirely in
// registers; controls and diagonal gates may look at ANY index bit (register slot,
// thread-fixed tile bit, or a bit outside the tile incl. rank bits) because the thread
// knows its full global index.
//   round 0      : amplitudes come straight from HBM into registers (coalesced: the
//                  low thread bits map to the low physical bits),
//   between rounds: registers -> shared tile -> registers with a different (reg_pos, thr_pos);
//                  the tile index is XOR-swizzled so that every quarter-warp of a 128-bit
//                  LDS/STS touches 8 distinct 16-byte bank groups,
//   last round   : registers -> HBM through store_bits (a permutation of load_bits, i.e. a
//                  free relabelling of the tile's qubits).
// HBM traffic is exactly 2 * sizeof(amp) * 2^n per pass, whatever the number of gates.
#pragma once
#include "common.cuh"

constexpr int kRegBits = QSV_REG_BITS;          // 4
constexpr int kRegAmps = 1 << kRegBits;         // 16 amplitudes per thread

// GF(2)-linear swizzle of a tile index: fold every W-bit group above the lowest into the
// lowest W bits (W = log2(128 B / sizeof(amp)): 3 for complex128, 4 for complex64).
template <int W> __host__ __device__ __forceinline__ uint32_t tile_swizzle(uint32_t x) {
    constexpr uint32_t M = (1u << W) - 1u;
    uint32_t f = x;
#pragma unroll
    for (int s = W; s < 16; s += W) f ^= (x >> s) & M;
    return f;
}

template <typename V> __device__ __forceinline__ V cx_neg(V a) { a.x = -a.x; a.y = -a.y; return a; }

// ---- op bodies: everything is unrolled over the 16 register slots ----------------------
// general complex 2x2 on register slot TB
template <typename V, int TB>
__device__ __forceinline__ void op_mat(V (&v)[kRegAmps], const V u00, const V u01, const V u10,
                                       const V u11, const uint32_t rc) {
#pragma unroll
    for (int j = 0; j < kRegAmps; ++j) {
        if (j & (1 << TB)) continue;
        if ((j & rc) != rc) continue;              // rc is warp-uniform: no divergence
        const V a = v[j], b = v[j | (1 << TB)];
        v[j] = cx_fma(u01, b, cx_mul(u00, a));
        v[j | (1 << TB)] = cx_fma(u11, b, cx_mul(u10, a));
    }
}

// real 2x2 on register slot TB: half the flops of op_mat
template <typename V, typename R, int TB>
__device__ __forceinline__ void op_real(V (&v)[kRegAmps], const R m00, const R m01, const R m10,
                                        const R m11, const uint32_t rc) {
#pragma unroll
    for (int j = 0; j < kRegAmps; ++j) {
        if (j & (1 << TB)) continue;
        if ((j & rc) != rc) continue;
        const V a = v[j], b = v[j | (1 << TB)];
        v[j].x = fma(m01, b.x, m00 * a.x);
        v[j].y = fma(m01, b.y, m00 * a.y);
        v[j | (1 << TB)].x = fma(m11, b.x, m10 * a.x);
        v[j | (1 << TB)].y = fma(m11, b.y, m10 * a.y);
    }
}

template <typename V>
__device__ __forceinline__ void op_phase(V (&v)[kRegAmps], const V ph, const uint32_t rc) {
#pragma unroll
    for (int j = 0; j < kRegAmps; ++j) {
        if ((j & rc) != rc) continue;
        v[j] = cx_mul(ph, v[j]);
    }
}

template <typename V>
__device__ __forceinline__ void op_sign(V (&v)[kRegAmps], const uint32_t rc) {
#pragma unroll
    for (int j = 0; j < kRegAmps; ++j) {
        if ((j & rc) != rc) continue;
        v[j] = cx_neg(v[j]);
    }
}

#define QSV_DISPATCH_TB(CALL)                                   \
    switch (tb) {                                               \
        case 0: { constexpr int TB = 0; CALL; } break;          \
        case 1: { constexpr int TB = 1; CALL; } break;          \
        case 2: { constexpr int TB = 2; CALL; } break;          \
        default: { constexpr int TB = 3; CALL; } break;         \
    }

template <typename R>
__global__ void __launch_bounds__(512)
k_pass(typename CxT<R>::V *__restrict__ state, const qsv_pass *__restrict__ pass_ptr,
       const qsv_op *__restrict__ ops, const uint64_t rank_bits) {
    using V = typename CxT<R>::V;
    constexpr int W = (sizeof(V) == 16) ? 3 : 4;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    V *tile = reinterpret_cast<V *>(smem_raw);

    const qsv_pass &P = *pass_ptr;
    const int T = P.n_tile;
    const int n_thr_bits = T - kRegBits;
    const uint32_t tid = threadIdx.x;

    // tile base: spread blockIdx over the non-tile index bits (load_bits ascending)
    uint64_t base = blockIdx.x;
    for (int i = 0; i < T; ++i) base = insert_zero_bit(base, P.load_bits[i]);
    const uint64_t glob = rank_bits | base;

    V v[kRegAmps];
    const int n_rounds = P.n_rounds;

    for (int r = 0; r < n_rounds; ++r) {
        const qsv_round &rd = P.rounds[r];
        // thread-fixed part of the tile index and the 4 register strides
        uint32_t xb = 0;
        for (int i = 0; i < n_thr_bits; ++i) xb |= ((tid >> i) & 1u) << rd.thr_pos[i];
        uint32_t xr[kRegBits];
#pragma unroll
        for (int b = 0; b < kRegBits; ++b) xr[b] = 1u << rd.reg_pos[b];

        if (r == 0) {
            // HBM -> registers
            uint64_t gb = base;
            for (int i = 0; i < n_thr_bits; ++i)
                gb |= (uint64_t)((tid >> i) & 1u) << P.load_bits[rd.thr_pos[i]];
            uint64_t gr[kRegBits];
#pragma unroll
            for (int b = 0; b < kRegBits; ++b) gr[b] = 1ull << P.load_bits[rd.reg_pos[b]];
#pragma unroll
            for (int j = 0; j < kRegAmps; ++j) {
                uint64_t a = gb;
#pragma unroll
                for (int b = 0; b < kRegBits; ++b) if (j & (1 << b)) a |= gr[b];
                v[j] = state[a];
            }
        } else {
            // shared tile -> registers (the tile was written by the previous round)
            const uint32_t sb = tile_swizzle<W>(xb);
            uint32_t sr[kRegBits];
#pragma unroll
            for (int b = 0; b < kRegBits; ++b) sr[b] = tile_swizzle<W>(xr[b]);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < kRegAmps; ++j) {
                uint32_t a = sb;
#pragma unroll
                for (int b = 0; b < kRegBits; ++b) if (j & (1 << b)) a ^= sr[b];
                v[j] = tile[a];
            }
        }

        // ---- the round's gates, all in registers ----
        for (int o = rd.op_begin; o < rd.op_end; ++o) {
            const qsv_op &op = ops[o];
            if ((glob & op.glob_ctrl) != op.glob_ctrl) continue;        // CTA-uniform
            if ((xb & op.tile_ctrl) != op.tile_ctrl) continue;          // per thread
            const uint32_t rc = op.reg_ctrl;
            const int tb = op.target;
            switch (op.kind) {
                case QSV_OP_MAT: {
                    const V u00 = cx_make<V>((R)op.m[0], (R)op.m[1]), u01 = cx_make<V>((R)op.m[2], (R)op.m[3]);
                    const V u10 = cx_make<V>((R)op.m[4], (R)op.m[5]), u11 = cx_make<V>((R)op.m[6], (R)op.m[7]);
                    QSV_DISPATCH_TB((op_mat<V, TB>(v, u00, u01, u10, u11, rc)));
                } break;
                case QSV_OP_REAL: {
                    const R m00 = (R)op.m[0], m01 = (R)op.m[2], m10 = (R)op.m[4], m11 = (R)op.m[6];
                    QSV_DISPATCH_TB((op_real<V, R, TB>(v, m00, m01, m10, m11, rc)));
                } break;
                case QSV_OP_PHASE:
                    op_phase<V>(v, cx_make<V>((R)op.m[0], (R)op.m[1]), rc);
                    break;
                case QSV_OP_SIGN:
                    op_sign<V>(v, rc);
                    break;
                default: break;
            }
        }

        if (r == n_rounds - 1) {
            // registers -> HBM through store_bits
            uint64_t gb = base;
            for (int i = 0; i < n_thr_bits; ++i)
                gb |= (uint64_t)((tid >> i) & 1u) << P.store_bits[rd.thr_pos[i]];
            uint64_t gr[kRegBits];
#pragma unroll
            for (int b = 0; b < kRegBits; ++b) gr[b] = 1ull << P.store_bits[rd.reg_pos[b]];
#pragma unroll
            for (int j = 0; j < kRegAmps; ++j) {
                uint64_t a = gb;
#pragma unroll
                for (int b = 0; b < kRegBits; ++b) if (j & (1 << b)) a |= gr[b];
                state[a] = v[j];
            }
        } else {
            const uint32_t sb = tile_swizzle<W>(xb);
            uint32_t sr[kRegBits];
#pragma unroll
            for (int b = 0; b < kRegBits; ++b) sr[b] = tile_swizzle<W>(xr[b]);
            // (a thread writes exactly the slots it read in this round: no barrier needed here)
#pragma unroll
            for (int j = 0; j < kRegAmps; ++j) {
                uint32_t a = sb;
#pragma unroll
                for (int b = 0; b < kRegBits; ++b) if (j & (1 << b)) a ^= sr[b];
                tile[a] = v[j];
            }
        }
    }
}
