// jit_prelude.cuh — everything a run-time specialised pass kernel needs, NVRTC-clean (no system
// headers): the register-level gate bodies of pass_ops.cuh plus the ring plumbing of
// pass_ring.cuh (shared-memory ring, mbarriers, cp.async).
#pragma once
#define QSV_JIT 1
typedef unsigned char uint8_t;
typedef unsigned int uint32_t;
typedef int int32_t;
typedef unsigned long long uint64_t;
#define QSV_REG_BITS 4
#define QSV_OP_HAD 0
#define QSV_OP_ROT 1
#define QSV_OP_XSWAP 2
#define QSV_OP_YSWAP 3
#define QSV_OP_PHASE 4
#define QSV_OP_SIGN 5
#define QSV_OP_SCALE 6
#define QSV_OPF_PRESIGN 1
#define QSV_OPF_PRENEG 2
#define QSV_OPF_PREPHASE 4
__device__ __forceinline__ uint64_t insert_zero_bit(uint64_t x, int pos) {
    const uint64_t low = x & ((1ull << pos) - 1ull);
    return ((x >> pos) << (pos + 1)) | low;
}
#include "pass_ops.cuh"

// element type of the pass: complex128 (default) or complex64 (-> #define JIT_F32 before the include)
#ifdef JIT_F32
typedef float2 JV;
typedef float JR;
#ifndef JIT_NBUF
#define JIT_NBUF 12          // 12 x 16 KB: four buffers per consumer group
#endif
// shared slot of tile index x: the 16-byte PAIR index is swizzled, the pair stays in order, so a
// 16-byte cp.async still lands two consecutive amplitudes correctly
__device__ __forceinline__ uint32_t jit_slot(uint32_t x) { return (tile_swizzle<3>(x >> 1) << 1) | (x & 1u); }
#else
typedef double2 JV;
typedef double JR;
#ifndef JIT_NBUF
#define JIT_NBUF 6           //  6 x 32 KB: two buffers per consumer group
#endif
__device__ __forceinline__ uint32_t jit_slot(uint32_t x) { return tile_swizzle<3>(x); }
#endif

// targets of a scatter pass (pass fused with the exchange that follows it): p[x] = the second buffer
// of the rank whose swapped rank bits equal x, keep = this rank's swapped bits at the local positions
struct JitDst { JV *p[8]; unsigned long long keep; };

// index bits fixed from outside (a launch over ONE CHUNK of the shard: pipelined stage transitions run the
// passes next to a swap chunk by chunk): the tile counter runs over the remaining non-tile bits; pos[] are
// positions in TILE-INDEX space (tile bits removed), ascending; val = the fixed bits at those positions
// blk: the CTAs are dealt blocks of 2^blk CONSECUTIVE tiles round-robin (0 = tile by tile), see jit_seq
struct JitFix { unsigned n; unsigned pos[4]; unsigned blk; unsigned long long val; };
__device__ __forceinline__ unsigned long long jit_fix(unsigned long long t, const JitFix &F) {
#pragma unroll
    for (int i = 0; i < 4; ++i) if (i < (int)F.n) t = insert_zero_bit(t, (int)F.pos[i]);
    return t | F.val;
}
// the s-th tile (relative to tile_begin) of this CTA.  Strictly increasing in s for every CTA, and (cta, s) -> tile is
// a bijection onto 0, 1, 2, ...: producers and consumer groups of a CTA walk the same sequence and stop at the
// first tile beyond the range.
__device__ __forceinline__ unsigned jit_seq(unsigned s, unsigned blk) {
    return ((((s >> blk) * gridDim.x) + blockIdx.x) << blk) + (s & ((1u << blk) - 1u));
}

struct JitRingSmem {
    JV buf[JIT_NBUF][2048];
    unsigned long long full[JIT_NBUF];
    unsigned long long empty[JIT_NBUF];
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try(unsigned long long *bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return done != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, uint32_t parity) {
    while (!mbar_try(bar, parity)) {}
}
__device__ __forceinline__ void mbar_wait_sleep(unsigned long long *bar, uint32_t parity) {
    while (!mbar_try(bar, parity)) __nanosleep(100);
}
__device__ __forceinline__ void cp_async16(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_arrive(unsigned long long *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// named barrier of one consumer group (128 threads; barrier 0 is __syncthreads)
__device__ __forceinline__ void group_bar(int grp) {
    asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(128) : "memory");
}

// the round's fold-table entry (unit complex) applied to all 16 amplitudes
__device__ __forceinline__ void apply_fold(JV (&v)[16], const double2 f) {
    JR pr = (JR)f.x, pi = (JR)f.y;
    const bool neg = pr < (JR)0;
    if (neg) { pr = -pr; pi = -pi; }
    if (pi != (JR)0) op_phase_mask<JV, JR>(v, pi / ((JR)1 + pr), pi, 0u);
    const int mask = neg ? (int)0x80000000 : 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) { v[j].x = xor_sign(v[j].x, mask); v[j].y = xor_sign(v[j].y, mask); }
}

#define JIT_RING_PROLOGUE                                                              \
    extern __shared__ __align__(128) unsigned char smem_raw[];                         \
    JitRingSmem &S = *reinterpret_cast<JitRingSmem *>(smem_raw);                       \
    const int tid = threadIdx.x;                                                       \
    if (tid == 0) {                                                                    \
        for (int b = 0; b < JIT_NBUF; ++b) { mbar_init(&S.full[b], 128); mbar_init(&S.empty[b], 128); } \
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");             \
    }                                                                                  \
    __syncthreads();
