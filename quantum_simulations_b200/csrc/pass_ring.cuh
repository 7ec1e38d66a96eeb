// pass_ring.cuh — the persistent, asynchronously fed pass kernel (complex128, 2^11-amp tiles).
//
// Why (profiles/r01): the one-CTA-per-tile kernel (pass_kernel.cuh) streams at the HBM
// roofline only while every resident warp has loads in flight; a tile that sits in registers
// waiting for shared-memory rounds or FP64 work takes its in-flight bytes with it, so exchange
// rounds and arithmetic ADD to the HBM time instead of hiding under it.  Here the HBM traffic
// is decoupled from the math:
//
//   * one persistent CTA per SM, 512 threads = 3 consumer GROUPS of 128 + a PRODUCER warpgroup;
//     setmaxnreg moves registers from the producers (40) to the consumers (152);
//   * a ring of 6 shared-memory buffers of 32 KB (one 2^11-amplitude tile each); buffer b
//     always belongs to group b % 3, so each group double-buffers privately;
//   * producers stream tiles HBM -> shared with cp.async (LDGSTS, 16 B per lane, no register
//     staging), writing directly in the XOR-swizzled layout, and publish a tile through an
//     mbarrier (cp.async.mbarrier.arrive.noinc);
//   * a group pulls its tile into registers (16 amplitudes per thread), runs the pass's
//     rounds (gates in registers; in-place exchange through its buffer between rounds, one
//     named barrier per exchange), releases the buffer to the producers after the last
//     shared read, and stores registers -> HBM with coalesced 128-bit stores.
//   Shared-memory rounds, FP64 math and HBM streaming of different tiles overlap on the SM.
//
// Per-thread phase folding: PHASE / SIGN ops whose controls are all thread-fixed (no register
// slot involved) do not touch the 16 amplitudes; they multiply a per-thread unit scalar P that
// is applied once at the end of the round (sign-bit flips when every P of the warp is real).
#pragma once
#include "common.cuh"
#include "pass_ops.cuh"

constexpr int kRingT = 11;                         // tile bits
constexpr int kRingTileAmps = 1 << kRingT;         // 2048 amplitudes = 32 KB
#ifndef QSV_RING_GROUPS
#define QSV_RING_GROUPS 3
#endif
#ifndef QSV_RING_CONSUMER_REGS
#define QSV_RING_CONSUMER_REGS 152
#endif
constexpr int kRingGroups = QSV_RING_GROUPS;
constexpr int kRingBufs = 6;
constexpr int kRingGroupThreads = kRingTileAmps / kRegAmps;   // 128
constexpr int kRingProducers = 128;                // 4 producer warps = one warpgroup
constexpr int kRingThreads = kRingGroups * kRingGroupThreads + kRingProducers;   // 512
constexpr int kRingMaxOps = 400;                   // ops of one pass kept in shared memory
constexpr int kRingPrefetchAhead = 0;              // >0: bulk-prefetch tiles beyond the ring into L2 (measured: hurts, TMA issue rate)

struct RingSmem {
    double2 buf[kRingBufs][kRingTileAmps];         // 196608 B
    qsv_op ops[kRingMaxOps];                       //  19200 B
    qsv_pass pass;                                 //    572 B
    unsigned long long p_off[16];                  // producer tables: tile index bits 7..10 -> HBM offset
    uint32_t p_sw[16];                             //                                      -> swizzled slot
    unsigned long long full[kRingBufs];
    unsigned long long empty[kRingBufs];
};

// dense dispatch codes (kept in bits 2..7 of qsv_op.target while the ops are staged in shared memory)
enum : int { RC_HAD = 0, RC_ROT = 4, RC_MIX_END = 8, RC_PHASE1 = 8, RC_SIGN1 = 12, RC_SIGN2 = 16,
             RC_FOLD_SIGN = 22, RC_FOLD_PHASE = 23, RC_SCALE = 24, RC_XSWAP = 25, RC_YSWAP = 29, RC_TPHASE = 33, RC_GENERIC = 37 };

__device__ __forceinline__ int ring_dispatch_code(int kind, int tb, uint32_t rc, uint32_t fl, bool other_ctrl) {
    const bool one = rc != 0 && (rc & (rc - 1)) == 0;
    const int slot = 31 - __clz((int)(rc | 1));
    switch (kind) {
        case QSV_OP_HAD: return RC_HAD + tb;
        case QSV_OP_ROT: return (rc || (!fl && other_ctrl)) ? RC_GENERIC : RC_ROT + tb;
        case QSV_OP_XSWAP: return rc ? RC_GENERIC : RC_XSWAP + tb;
        case QSV_OP_YSWAP: return rc ? RC_GENERIC : RC_YSWAP + tb;
        case QSV_OP_PHASE: return rc == 0 ? RC_FOLD_PHASE : (one ? RC_PHASE1 + slot : RC_GENERIC);
        case QSV_OP_SIGN:
            if (rc == 0) return RC_FOLD_SIGN;
            if (one) return RC_SIGN1 + slot;
            if (__popc(rc) == 2) {                       // pair (lo, hi) -> 0..5
                const int lo = __ffs((int)rc) - 1, hi = slot;
                return RC_SIGN2 + (lo == 0 ? hi - 1 : (lo == 1 ? hi + 1 : 5));
            }
            return RC_GENERIC;
        case QSV_OP_SCALE: return RC_SCALE;
        case QSV_OP_TPHASE: return RC_TPHASE + tb;
        default: return RC_GENERIC;
    }
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    } while (!done);
}
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
// the mbarrier receives one arrival when all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive(unsigned long long *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void l2_prefetch_bulk(const void *gmem, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gmem), "r"(bytes) : "memory");
}

__device__ __forceinline__ uint64_t ring_tile_base(const qsv_pass &P, uint32_t tile) {
    uint64_t base = tile;
#pragma unroll 1
    for (int i = 0; i < kRingT; ++i) base = insert_zero_bit(base, P.load_bits[i]);
    return base;
}

__global__ void __launch_bounds__(kRingThreads, 1)
k_pass_ring(double2 *__restrict__ state, const qsv_pass *__restrict__ pass_ptr,
            const qsv_op *__restrict__ ops_ptr, const double2 *__restrict__ tables,
            const uint64_t rank_bits, const uint32_t n_tiles) {
    using V = double2;
    using R = double;
    constexpr int W = 3, T = kRingT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    RingSmem &S = *reinterpret_cast<RingSmem *>(smem_raw);
    const int tid = threadIdx.x;

    // ---- one-time setup: pass descriptor + ops into shared memory, tables, barriers ----
    {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(pass_ptr);
        uint32_t *dst = reinterpret_cast<uint32_t *>(&S.pass);
        for (int i = tid; i < (int)(sizeof(qsv_pass) / 4); i += kRingThreads) dst[i] = src[i];
        const int n_ops = pass_ptr->n_ops;
        const uint4 *so = reinterpret_cast<const uint4 *>(ops_ptr);
        uint4 *sd = reinterpret_cast<uint4 *>(S.ops);
        for (int i = tid; i < n_ops * (int)(sizeof(qsv_op) / 16); i += kRingThreads) sd[i] = so[i];
        __syncthreads();
        for (int i = tid; i < n_ops; i += kRingThreads) {
            const qsv_op &q = S.ops[i];
            const int code = ring_dispatch_code(q.kind, q.target, q.reg_ctrl, q.flags, q.tile_ctrl != 0 || q.glob_ctrl != 0);
            S.ops[i].target = (uint8_t)(q.target | (code << 2));
        }
        if (tid < 16) {
            unsigned long long o = 0;
            for (int i = 0; i < 4; ++i) if (tid & (1 << i)) o |= 1ull << S.pass.load_bits[7 + i];
            S.p_off[tid] = o;
            S.p_sw[tid] = tile_swizzle<W>((uint32_t)tid << 7);
        }
        if (tid == 0) {
            for (int b = 0; b < kRingBufs; ++b) {
                mbar_init(&S.full[b], kRingProducers);
                mbar_init(&S.empty[b], kRingGroupThreads);
            }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
    }
    __syncthreads();
    const qsv_pass &P = S.pass;

    if (tid >= kRingGroups * kRingGroupThreads) {
        // =========================== producers ===========================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        const uint32_t pt = tid - kRingGroups * kRingGroupThreads;          // 0..127 = tile index bits 0..6
        uint64_t off_lo = 0;
#pragma unroll
        for (int i = 0; i < 7; ++i) off_lo |= (uint64_t)((pt >> i) & 1u) << P.load_bits[i];
        const uint32_t sw_lo = tile_swizzle<W>(pt);
        // L2 prefetch geometry: the tile is 2^(T-run) contiguous runs of 2^run amplitudes
        int run = 0;
        while (run < T && P.load_bits[run] == run) ++run;
        const uint32_t run_bytes = 16u << run;
        const uint32_t n_runs = 1u << (T - run);
        uint32_t s = 0;
        for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++s) {
            const int b = s % kRingBufs;
            const uint32_t use = s / kRingBufs;
            // deepen the HBM queue: pull a tile that is still outside the ring into L2
            const uint64_t ahead = (uint64_t)tile + (uint64_t)gridDim.x * (kRingBufs - kRingGroups + kRingPrefetchAhead);
            if (kRingPrefetchAhead > 0 && ahead < n_tiles) {
                const uint64_t abase = ring_tile_base(P, (uint32_t)ahead);
                for (uint32_t r = pt; r < n_runs; r += kRingProducers) {
                    uint64_t o = 0;
                    for (int i = run; i < T; ++i) o |= (uint64_t)((r >> (i - run)) & 1u) << P.load_bits[i];
                    l2_prefetch_bulk(state + abase + o, run_bytes);
                }
            }
            const double2 *g = state + ring_tile_base(P, tile) + off_lo;
            double2 *d = S.buf[b];
            if (use > 0) mbar_wait(&S.empty[b], (use - 1) & 1);
#pragma unroll
            for (int k = 0; k < 16; ++k) cp_async16(d + (sw_lo ^ S.p_sw[k]), g + S.p_off[k]);
            cp_async_arrive(&S.full[b]);
        }
        return;
    }

    // =========================== consumer groups ===========================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(QSV_RING_CONSUMER_REGS));
    const int grp = tid / kRingGroupThreads;
    const uint32_t gt = tid % kRingGroupThreads;
    const int n_rounds = P.n_rounds;
    constexpr int n_thr_bits = T - kRegBits;      // 7

    uint32_t s = grp;
    for (uint32_t tile = blockIdx.x + (uint32_t)grp * gridDim.x; tile < n_tiles;
         tile += gridDim.x * kRingGroups, s += kRingGroups) {
        const int b = s % kRingBufs;
        const uint32_t use = s / kRingBufs;
        V *buf = S.buf[b];
        const uint64_t base = ring_tile_base(P, tile);
        const uint64_t glob = rank_bits | base;

        mbar_wait(&S.full[b], use & 1);

        V v[kRegAmps];
        for (int r = 0; r < n_rounds; ++r) {
            const qsv_round &rd = P.rounds[r];
            uint32_t xb = 0;
#pragma unroll
            for (int i = 0; i < n_thr_bits; ++i) xb |= ((gt >> i) & 1u) << rd.thr_pos[i];
            const uint32_t sb = tile_swizzle<W>(xb);
            uint32_t sr[kRegBits];
#pragma unroll
            for (int q = 0; q < kRegBits; ++q) sr[q] = tile_swizzle<W>(1u << rd.reg_pos[q]);
            // this thread's entry of the round's fold table (issued early: L2 latency hides
            // behind the shared-memory reads)
            const int fold_off = rd.fold_off;
            double2 fold = make_double2(1.0, 0.0);
            if (fold_off >= 0) fold = __ldg(&tables[fold_off + gt]);

            // shared tile -> registers
#pragma unroll
            for (int j = 0; j < kRegAmps; ++j) {
                uint32_t a = sb;
#pragma unroll
                for (int q = 0; q < kRegBits; ++q) if (j & (1 << q)) a ^= sr[q];
                v[j] = buf[a];
            }
            const bool last = (r == n_rounds - 1);
            if (last) mbar_arrive(&S.empty[b]);    // buffer no longer needed: back to the producers

            // ---- the round's gates, in registers ----
            R pr = fold.x, pi = fold.y;           // folded per-thread phase (unit modulus)
            bool dirty = fold_off >= 0;
            const int o_end = rd.op_end;
            int o = rd.op_begin;
            uint4 nxt = make_uint4(0, 0, 0, 0);
            double2 nxt_c = make_double2(0.0, 0.0);
            if (o < o_end) {
                nxt = *reinterpret_cast<const uint4 *>(&S.ops[o]);
                nxt_c = *reinterpret_cast<const double2 *>(S.ops[o].m);
            }
            for (; o < o_end; ++o) {
                const uint4 hd = nxt;                                      // header of op o
                const double2 c = nxt_c;                                   // m[0], m[1]
                if (o + 1 < o_end) {                                       // prefetch op o+1
                    nxt = *reinterpret_cast<const uint4 *>(&S.ops[o + 1]);
                    nxt_c = *reinterpret_cast<const double2 *>(S.ops[o + 1].m);
                }
                const uint32_t code = (hd.x >> 10) & 63u;
                if (code < RC_MIX_END) {
                    // HAD / ROT without controls: the bulk of the FP64 work.  Pre-ops (pending Z / CZ
                    // signs and the ZYZ pre-phase on the target) ride along in the same record.
                    const uint32_t fl = hd.x >> 24;
                    int sm = 0;
                    double2 pc = make_double2(0.0, 0.0);
                    if (fl) {
                        const uint64_t gc = ((uint64_t)hd.w << 32) | hd.z;
                        sm = (int)(((uint32_t)__popc(xb & hd.y) + (uint32_t)__popcll(glob & gc) + (fl >> 1)) << 31);
                        pc = *reinterpret_cast<const double2 *>(S.ops[o].m + 2);
                    }
                    switch (code) {
#define RING_CASE4(BASE, CALL)                                              \
                    case BASE + 0: { constexpr int TB = 0; CALL; } break;   \
                    case BASE + 1: { constexpr int TB = 1; CALL; } break;   \
                    case BASE + 2: { constexpr int TB = 2; CALL; } break;   \
                    case BASE + 3: { constexpr int TB = 3; CALL; } break;
                        RING_CASE4(RC_HAD, (op_pre<V, R, TB>(v, fl, sm, pc.x, pc.y), op_had<V, TB>(v)))
                        RING_CASE4(RC_ROT, (op_pre<V, R, TB>(v, fl, sm, pc.x, pc.y), op_rot<V, R, TB, false>(v, c.x, c.y, 0u)))
                        default: __builtin_unreachable();
                    }
                    continue;
                }
                double2 tf = make_double2(1.0, 0.0);
                if (code >= RC_TPHASE && code < RC_TPHASE + 4) tf = tphase_factor(tables, S.ops[o].m, gt, glob);
                if (hd.y | hd.z | hd.w) {
                    const uint64_t gc = ((uint64_t)hd.w << 32) | hd.z;
                    if ((glob & gc) != gc) continue;                       // group-uniform
                    if ((xb & hd.y) != hd.y) continue;                     // per thread
                }
                switch (code) {
                    RING_CASE4(RC_PHASE1, (op_phase_slot<V, R, TB>(v, c.x, c.y)))
                    RING_CASE4(RC_SIGN1, (op_sign_slot<V, TB>(v)))
                    RING_CASE4(RC_TPHASE, (op_cmul_slot<V, R, TB>(v, tf.x, tf.y)))
                    RING_CASE4(RC_XSWAP, (op_xswap<V, TB, false>(v, 0u)))
                    RING_CASE4(RC_YSWAP, (op_yswap<V, TB, false>(v, 0u)))
#undef RING_CASE4
                    case RC_SIGN2 + 0: op_sign_slot2<V, 0, 1>(v); break;
                    case RC_SIGN2 + 1: op_sign_slot2<V, 0, 2>(v); break;
                    case RC_SIGN2 + 2: op_sign_slot2<V, 0, 3>(v); break;
                    case RC_SIGN2 + 3: op_sign_slot2<V, 1, 2>(v); break;
                    case RC_SIGN2 + 4: op_sign_slot2<V, 1, 3>(v); break;
                    case RC_SIGN2 + 5: op_sign_slot2<V, 2, 3>(v); break;
                    case RC_FOLD_SIGN: pr = flip_sign(pr); pi = flip_sign(pi); dirty = true; break;
                    case RC_FOLD_PHASE: {
                        const double2 e = *reinterpret_cast<const double2 *>(S.ops[o].m + 2);   // cos, sin
                        const R nr = pr * e.x - pi * e.y;
                        pi = pr * e.y + pi * e.x; pr = nr; dirty = true;
                    } break;
                    case RC_SCALE: op_scale<V, R>(v, c.x); break;
                    case RC_GENERIC:
                        apply_reg_op<V, R>(v, hd.x & 0xff, (hd.x >> 8) & 3, (hd.x >> 16) & 0xff, S.ops[o].m);
                        break;
                    default: __builtin_unreachable();
                }
            }
            // apply the folded phase P in place; sign-bit flips only if every P of the warp is real
            if (__any_sync(0xffffffffu, dirty)) {
                const bool neg = pr < 0.0;
                if (__any_sync(0xffffffffu, pi != 0.0)) {
                    if (neg) { pr = -pr; pi = -pi; }
                    const R t = pi / (1.0 + pr);                   // tan(phi/2), |phi| <= pi/2 now
                    op_phase_mask<V, R>(v, t, pi, 0u);
                }
                const int mask = neg ? (int)0x80000000 : 0;
#pragma unroll
                for (int j = 0; j < kRegAmps; ++j) {
                    v[j].x = __hiloint2double(__double2hiint(v[j].x) ^ mask, __double2loint(v[j].x));
                    v[j].y = __hiloint2double(__double2hiint(v[j].y) ^ mask, __double2loint(v[j].y));
                }
            }

            if (last) {
                // registers -> HBM through store_bits (lanes drive the lowest store positions)
                uint64_t gb = base;
#pragma unroll
                for (int i = 0; i < n_thr_bits; ++i)
                    gb |= (uint64_t)((gt >> i) & 1u) << P.store_bits[rd.thr_pos[i]];
                uint64_t gr[kRegBits];
#pragma unroll
                for (int q = 0; q < kRegBits; ++q) gr[q] = 1ull << P.store_bits[rd.reg_pos[q]];
#pragma unroll
                for (int j = 0; j < kRegAmps; ++j) {
                    uint64_t a = gb;
#pragma unroll
                    for (int q = 0; q < kRegBits; ++q) if (j & (1 << q)) a |= gr[q];
                    state[a ^ P.store_flip] = v[j];
                }
            } else {
                // registers -> the same shared slots, then the group re-partitions the tile
#pragma unroll
                for (int j = 0; j < kRegAmps; ++j) {
                    uint32_t a = sb;
#pragma unroll
                    for (int q = 0; q < kRegBits; ++q) if (j & (1 << q)) a ^= sr[q];
                    buf[a] = v[j];
                }
                asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(kRingGroupThreads) : "memory");
            }
        }
    }
}
