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
#include "pass_ops.cuh"

template <typename R>
__global__ void __launch_bounds__(sizeof(R) == 8 ? 256 : 512, sizeof(R) == 8 ? 2 : 1)
k_pass(typename CxT<R>::V *__restrict__ state, const qsv_pass *__restrict__ pass_ptr,
       const qsv_op *__restrict__ ops, const double2 *__restrict__ tables, const uint64_t rank_bits) {
    using V = typename CxT<R>::V;
    constexpr int W = (sizeof(V) == 16) ? 3 : 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
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
        if (rd.fold_off >= 0) {                         // fold table: one phase per thread
            const double2 f = tables[rd.fold_off + tid];
            R pr = (R)f.x, pi = (R)f.y;
            const bool neg = pr < (R)0;
            if (neg) { pr = -pr; pi = -pi; }
            op_phase_mask<V, R>(v, pi / ((R)1 + pr), pi, 0u);
            if (neg) op_sign_mask<V>(v, 0u);
        }
        for (int o = rd.op_begin; o < rd.op_end; ++o) {
            const qsv_op &op = ops[o];
            if (op.kind == QSV_OP_TPHASE) {
                const double2 f = tphase_factor(tables, op.m, tid, glob);
                QSV_DISPATCH_TB(op.target, (op_cmul_slot<V, R, TB>(v, (R)f.x, (R)f.y)));
                continue;
            }
            if (op.flags) {     // HAD / ROT with pre-ops: the control fields are parity masks
                const int sm = (int)(((uint32_t)__popc(xb & op.tile_ctrl) + (uint32_t)__popcll(glob & op.glob_ctrl) +
                                      ((uint32_t)op.flags >> 1)) << 31);
                apply_reg_op<V, R>(v, op.kind, op.target, 0u, op.m, op.flags, sm);
                continue;
            }
            if ((glob & op.glob_ctrl) != op.glob_ctrl) continue;        // CTA-uniform
            if ((xb & op.tile_ctrl) != op.tile_ctrl) continue;          // per thread
            apply_reg_op<V, R>(v, op.kind, op.target, op.reg_ctrl, op.m);
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
                state[a ^ P.store_flip] = v[j];
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
