// HOST stand-in for csrc/jit_prelude.cuh (test infrastructure, tests/jit_host_run.py).
//
// The source qsvjit::generate() emits is CUDA: a persistent CTA of 128 producer threads (cp.async +
// mbarrier ring) and 3-4 consumer groups of 128 threads.  Compiled against THIS header with g++, the
// same source runs on the CPU with one OS thread per CUDA thread: cp.async is a 16-byte copy,
// mbarriers are arrival counters with a phase, the named barrier of a consumer group is a real
// barrier.  The gate bodies are the real ones (csrc/pass_ops.cuh is included unchanged), so what is
// checked is everything the generator decides: slot codes, round exchanges, fold tables, op order,
// coefficient indices, load / store / scatter addressing, the zero-input form.
#pragma once
#define QSV_JIT 1
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __restrict__
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(x) alignas(x)

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

struct double2 { double x, y; };
struct float2 { float x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }

// ---- thread identity (set by the launcher in jit_host_main.inc)
struct Dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local Dim3 threadIdx, blockIdx;
static Dim3 gridDim, blockDim;

// ---- intrinsics
static inline int __double2hiint(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b >> 32); }
static inline int __double2loint(double x) { int64_t b; memcpy(&b, &x, 8); return (int)(b & 0xffffffffll); }
static inline double __hiloint2double(int hi, int lo) {
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double x; memcpy(&x, &b, 8); return x;
}
static inline int __float_as_int(float x) { int b; memcpy(&b, &x, 4); return b; }
static inline float __int_as_float(int b) { float x; memcpy(&x, &b, 4); return x; }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
template <typename T> static inline T __ldg(const T *p) { return *p; }
template <typename T> static inline void __stcs(T *p, const T &v) { *p = v; }
using std::fma;

static inline uint64_t insert_zero_bit(uint64_t x, int pos) {
    const uint64_t low = x & ((1ull << pos) - 1ull);
    return ((x >> pos) << (pos + 1)) | low;
}
#include "pass_ops.cuh"

#ifdef JIT_F32
typedef float2 JV;
typedef float JR;
#ifndef JIT_NBUF
#define JIT_NBUF 12
#endif
static inline uint32_t jit_slot(uint32_t x) { return (tile_swizzle<3>(x >> 1) << 1) | (x & 1u); }
#else
typedef double2 JV;
typedef double JR;
#ifndef JIT_NBUF
#define JIT_NBUF 6
#endif
static inline uint32_t jit_slot(uint32_t x) { return tile_swizzle<3>(x); }
#endif

struct JitDst { JV *p[8]; unsigned long long keep; };
struct JitFix { unsigned n; unsigned pos[4]; unsigned blk; unsigned long long val; };
static inline unsigned long long jit_fix(unsigned long long t, const JitFix &F) {
    for (int i = 0; i < 4; ++i) if (i < (int)F.n) t = insert_zero_bit(t, (int)F.pos[i]);
    return t | F.val;
}
static inline unsigned jit_seq(unsigned s, unsigned blk) {
    return ((((s >> blk) * gridDim.x) + blockIdx.x) << blk) + (s & ((1u << blk) - 1u));
}

// ---- mbarrier: `count` arrivals complete a phase; wait(parity) returns once the phase of that parity is over
struct HostBar {
    std::atomic<uint32_t> phase{0};
    std::atomic<int> pending{0};
    int count = 0;
};
struct JitRingSmem {
    JV buf[JIT_NBUF][2048];
    HostBar full[JIT_NBUF];
    HostBar empty[JIT_NBUF];
};
static inline void mbar_init(HostBar *b, uint32_t count) { b->count = (int)count; b->pending = (int)count; b->phase = 0; }
static inline void mbar_arrive(HostBar *b) {
    if (b->pending.fetch_sub(1, std::memory_order_acq_rel) == 1) {
        b->pending.store(b->count, std::memory_order_relaxed);
        b->phase.fetch_add(1, std::memory_order_release);
    }
}
static inline bool mbar_try(HostBar *b, uint32_t parity) { return (b->phase.load(std::memory_order_acquire) & 1u) != (parity & 1u); }
static inline void mbar_wait(HostBar *b, uint32_t parity) { while (!mbar_try(b, parity)) std::this_thread::yield(); }
static inline void mbar_wait_sleep(HostBar *b, uint32_t parity) { mbar_wait(b, parity); }
static inline void cp_async16(void *dst, const void *src) { memcpy(dst, src, 16); }
static inline void cp_async_arrive(HostBar *b) { mbar_arrive(b); }     // the copies above are already done

// ---- barriers among the threads of the (single) running CTA
struct SpinBarrier {
    std::atomic<int> waiting{0};
    std::atomic<uint32_t> gen{0};
    int n = 0;
    void wait() {
        const uint32_t g = gen.load(std::memory_order_acquire);
        if (waiting.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
            waiting.store(0, std::memory_order_relaxed);
            gen.fetch_add(1, std::memory_order_release);
        } else {
            while (gen.load(std::memory_order_acquire) == g) std::this_thread::yield();
        }
    }
};
static SpinBarrier g_cta_bar, g_group_bar[4], g_warp_bar[32];
static JitRingSmem *g_smem = nullptr;
static inline void __syncthreads() { g_cta_bar.wait(); }
static inline void group_bar(int grp) { g_group_bar[grp].wait(); }
static inline void __syncwarp() { g_warp_bar[threadIdx.x >> 5].wait(); }

static inline void apply_fold(JV (&v)[16], const double2 f) {
    JR pr = (JR)f.x, pi = (JR)f.y;
    const bool neg = pr < (JR)0;
    if (neg) { pr = -pr; pi = -pi; }
    if (pi != (JR)0) op_phase_mask<JV, JR>(v, pi / ((JR)1 + pr), pi, 0u);
    const int mask = neg ? (int)0x80000000 : 0;
    for (int j = 0; j < 16; ++j) { v[j].x = xor_sign(v[j].x, mask); v[j].y = xor_sign(v[j].y, mask); }
}

#define JIT_RING_PROLOGUE                                                                   \
    JitRingSmem &S = *g_smem;                                                               \
    const int tid = (int)threadIdx.x;                                                       \
    if (tid == 0) {                                                                         \
        for (int b = 0; b < JIT_NBUF; ++b) { mbar_init(&S.full[b], 128); mbar_init(&S.empty[b], 128); } \
    }                                                                                       \
    __syncthreads();
