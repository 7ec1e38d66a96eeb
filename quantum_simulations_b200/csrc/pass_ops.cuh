// pass_ops.cuh — register-level gate bodies shared by the pass kernels.
//
// A thread holds 16 amplitudes v[0..16): register slot b (0..3) is bit b of the index j.
// Every body is fully unrolled so v[] stays in registers; the target slot TB is a template
// parameter (runtime dispatch happens once per op, outside).
//
// ALL bodies are IN PLACE: every arithmetic statement overwrites one of its own operands
// ("lifting" steps).  This matters more than the flop count: profiles/r01 shows that bodies
// which compute into temporaries make ptxas shuffle the 64 data registers at the dispatch
// join (>50 % of executed instructions were IMAD.MOV); the in-place forms compile to zero
// moves, ~80 registers and a jump-table dispatch.
//   HAD    b <- a - b ; a <- 2a - b                         2 FP64 / real component pair
//   ROT    a -= t b ; b += s a ; a -= t b                   3   (t = tan(theta/2), s = sin(theta))
//   PHASE  re -= t im ; im += s re ; re -= t im             3 / amplitude it touches
//   SIGN / XSWAP / YSWAP                                    0   (integer pipe: xor on sign / words)
//   SCALE  re *= s ; im *= s                                2 / amplitude
// FP64 (B200: 64 lanes/SM, 34 TFLOP/s measured) is the scarce resource next to HBM.
#pragma once
#ifndef QSV_JIT          // the NVRTC build (jit_prelude.cuh) supplies the few names needed instead
#include "common.cuh"
#endif

constexpr int kRegBits = QSV_REG_BITS;          // 4
constexpr int kRegAmps = 1 << kRegBits;         // 16 amplitudes per thread

__device__ __forceinline__ double flip_sign(double x) {
    return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x));
}
__device__ __forceinline__ float flip_sign(float x) { return __int_as_float(__float_as_int(x) ^ (int)0x80000000); }
// x with its sign bit xor-ed by m (m = 0 or 0x80000000)
__device__ __forceinline__ double xor_sign(double x, int m) { return __hiloint2double(__double2hiint(x) ^ m, __double2loint(x)); }
__device__ __forceinline__ float xor_sign(float x, int m) { return __int_as_float(__float_as_int(x) ^ m); }

// swap two scalars in place with three xors per 32-bit word (no temporaries => no shuffles)
__device__ __forceinline__ void xor_swap(double &a, double &b) {
    int ah = __double2hiint(a), al = __double2loint(a), bh = __double2hiint(b), bl = __double2loint(b);
    ah ^= bh; bh ^= ah; ah ^= bh;
    al ^= bl; bl ^= al; al ^= bl;
    a = __hiloint2double(ah, al); b = __hiloint2double(bh, bl);
}
__device__ __forceinline__ void xor_swap(float &a, float &b) {
    int x = __float_as_int(a), y = __float_as_int(b);
    x ^= y; y ^= x; x ^= y;
    a = __int_as_float(x); b = __int_as_float(y);
}

// CHECK: honour register-slot controls rc (warp-uniform value)
#define QSV_PAIR_LOOP(TB, CHECK, rc)                                   \
    _Pragma("unroll") for (int j = 0; j < kRegAmps; ++j)               \
        if (!(j & (1 << TB)) && (!(CHECK) || (j & (rc)) == (rc)))

// PRE-OPS of a mixing op (qsv.h QSV_OPF_*): diagonal work on the b half (target bit = 1) that the
// compiler folded into the HAD / ROT that follows it, so it costs no dispatch of its own:
//   sign   b.hi ^= sm      sm = 0 or 0x80000000 per THREAD (parity of the partner bits of the CZ / Z
//                          gates pending on the target) — 2 LOP3 per b on the integer pipe
//   phase  b *= e^{i phi}  three shears per b (the diag(1, e^{il}) factor of the ZYZ form)
template <typename V, typename R, int TB>
__device__ __forceinline__ void op_pre(V (&v)[kRegAmps], const uint32_t fl, const int sm, const R tp, const R sp) {
    if (fl & (QSV_OPF_PRESIGN | QSV_OPF_PRENEG)) {
#pragma unroll
        for (int j = 0; j < kRegAmps; ++j) if (j & (1 << TB)) {
            v[j].x = xor_sign(v[j].x, sm); v[j].y = xor_sign(v[j].y, sm);
        }
    }
    if (fl & QSV_OPF_PREPHASE) {
#pragma unroll
        for (int j = 0; j < kRegAmps; ++j) if (j & (1 << TB)) {
            v[j].x = fma(-tp, v[j].y, v[j].x);
            v[j].y = fma(sp, v[j].x, v[j].y);
            v[j].x = fma(-tp, v[j].y, v[j].x);
        }
    }
}

template <typename V, int TB>
__device__ __forceinline__ void op_had(V (&v)[kRegAmps]) {
    using R = decltype(V::x);
    QSV_PAIR_LOOP(TB, false, 0u) {
        V &a = v[j], &b = v[j | (1 << TB)];
        b.x = a.x - b.x; a.x = fma((R)2, a.x, -b.x);
        b.y = a.y - b.y; a.y = fma((R)2, a.y, -b.y);
    }
}

template <typename V, typename R, int TB, bool CHECK>
__device__ __forceinline__ void op_rot(V (&v)[kRegAmps], const R t, const R s, const uint32_t rc) {
    QSV_PAIR_LOOP(TB, CHECK, rc) {
        V &a = v[j], &b = v[j | (1 << TB)];
        a.x = fma(-t, b.x, a.x); a.y = fma(-t, b.y, a.y);
        b.x = fma(s, a.x, b.x);  b.y = fma(s, a.y, b.y);
        a.x = fma(-t, b.x, a.x); a.y = fma(-t, b.y, a.y);
    }
}

// ---- sign-deferring forms (used by the run-time specialised kernels, csrc/jit.cuh) --------------------------
// The specialised kernels do not execute sign flips where the circuit has them (a quarter of the instructions
// of a pass were LOP3 on sign bits): the generator carries them as PENDING signs — per register a compile-time
// bit (NEG), per register slot a per-thread mask (the parity of CZ partners) — and folds them into the next
// mixing op of that slot, where they are free: operand negation, or a +-1 / sign-carrying coefficient.
__device__ __forceinline__ double sign_factor(double, int m) { return __hiloint2double(0x3ff00000 ^ m, 0); }   // +-1.0
__device__ __forceinline__ float sign_factor(float, int m) { return __int_as_float(0x3f800000 ^ m); }

// HAD on (A, sg * B) with A = +-a, B = +-b (compile-time signs NEG, bit j = register j); sg = +-1 per thread.
// Outputs carry no pending sign:  b <- A - sg B ;  a <- 2A - b
template <typename V, typename R, int TB, unsigned NEG>
__device__ __forceinline__ void op_had_sg(V (&v)[kRegAmps], const R sg) {
#pragma unroll
    for (int j = 0; j < kRegAmps; ++j) if (!(j & (1 << TB))) {
        V &a = v[j], &b = v[j | (1 << TB)];
        const bool na = (NEG >> j) & 1u, nb = (NEG >> (j | (1 << TB))) & 1u;
        b.x = fma(nb ? sg : -sg, b.x, na ? -a.x : a.x); a.x = fma(na ? (R)-2 : (R)2, a.x, -b.x);
        b.y = fma(nb ? sg : -sg, b.y, na ? -a.y : a.y); a.y = fma(na ? (R)-2 : (R)2, a.y, -b.y);
    }
}
// ROT on (A, sigma * B), ts = t * sigma, ss = s * sigma (sigma = +-1 per thread, folded into the coefficients by
// the caller).  The a outputs are clean; the b registers hold sigma * (true b): sigma STAYS PENDING on the b half.
template <typename V, typename R, int TB, unsigned NEG>
__device__ __forceinline__ void op_rot_sg(V (&v)[kRegAmps], const R ts, const R ss) {
#pragma unroll
    for (int j = 0; j < kRegAmps; ++j) if (!(j & (1 << TB))) {
        V &a = v[j], &b = v[j | (1 << TB)];
        const bool na = (NEG >> j) & 1u, nb = (NEG >> (j | (1 << TB))) & 1u;
        a.x = fma(nb ? ts : -ts, b.x, na ? -a.x : a.x); a.y = fma(nb ? ts : -ts, b.y, na ? -a.y : a.y);
        b.x = fma(ss, a.x, nb ? -b.x : b.x);            b.y = fma(ss, a.y, nb ? -b.y : b.y);
        a.x = fma(-ts, b.x, a.x);                       a.y = fma(-ts, b.y, a.y);
    }
}
// SCALE that also settles the compile-time pending signs
template <typename V, typename R, unsigned NEG>
__device__ __forceinline__ void op_scale_neg(V (&v)[kRegAmps], const R s) {
#pragma unroll
    for (int j = 0; j < kRegAmps; ++j) { const R f = ((NEG >> j) & 1u) ? -s : s; v[j].x *= f; v[j].y *= f; }
}

template <typename V, int TB, bool CHECK>
__device__ __forceinline__ void op_xswap(V (&v)[kRegAmps], const uint32_t rc) {
    QSV_PAIR_LOOP(TB, CHECK, rc) {
        xor_swap(v[j].x, v[j | (1 << TB)].x);
        xor_swap(v[j].y, v[j | (1 << TB)].y);
    }
}

// Y: (a, b) -> (-i b, i a):  a.x <-> b.y, a.y <-> b.x, then negate a.y and b.x
template <typename V, int TB, bool CHECK>
__device__ __forceinline__ void op_yswap(V (&v)[kRegAmps], const uint32_t rc) {
    QSV_PAIR_LOOP(TB, CHECK, rc) {
        V &a = v[j], &b = v[j | (1 << TB)];
        xor_swap(a.x, b.y);
        xor_swap(a.y, b.x);
        a.y = flip_sign(a.y); b.x = flip_sign(b.x);
    }
}

template <typename V, typename R>
__device__ __forceinline__ void phase_inplace(V &a, const R t, const R s) {
    a.x = fma(-t, a.y, a.x);
    a.y = fma(s, a.x, a.y);
    a.x = fma(-t, a.y, a.x);
}

// diagonal ops controlled by exactly ONE register slot TB: the 8 amplitudes with bit TB set
template <typename V, typename R, int TB>
__device__ __forceinline__ void op_phase_slot(V (&v)[kRegAmps], const R t, const R s) {
#pragma unroll
    for (int j = 0; j < kRegAmps; ++j) if (j & (1 << TB)) phase_inplace<V, R>(v[j], t, s);
}
template <typename V, int TB>
__device__ __forceinline__ void op_sign_slot(V (&v)[kRegAmps]) {
#pragma unroll
    for (int j = 0; j < kRegAmps; ++j) if (j & (1 << TB)) { v[j].x = flip_sign(v[j].x); v[j].y = flip_sign(v[j].y); }
}

// TABLE PHASE: the b half of slot TB times a unit complex number (general product: 4 FP64 per amplitude)
template <typename V, typename R, int TB>
__device__ __forceinline__ void op_cmul_slot(V (&v)[kRegAmps], const R fr, const R fi) {
#pragma unroll
    for (int j = 0; j < kRegAmps; ++j) if (j & (1 << TB)) {
        const R x = v[j].x, y = v[j].y;
        v[j].x = fma(-y, fi, x * fr);
        v[j].y = fma(y, fr, x * fi);
    }
}
// the factor of a QSV_OP_TPHASE op for this thread / tile
__device__ __forceinline__ double2 tphase_factor(const double2 *__restrict__ tables, const double *__restrict__ m,
                                                 const uint32_t thread, const uint64_t glob) {
    double2 f = make_double2(1.0, 0.0);
    const int toff = (int)m[0], goff = (int)m[1];
    uint32_t mask = (uint32_t)m[2];
    if (toff >= 0) f = tables[toff + thread];
    if (goff >= 0) {
        int k = 0;
        while (mask) {
            const int r = __ffs((int)mask) - 1;
            mask &= mask - 1;
            const double2 g = tables[goff + 256 * k + (int)((glob >> (8 * r)) & 255ull)];
            f = make_double2(f.x * g.x - f.y * g.y, f.x * g.y + f.y * g.x);
            ++k;
        }
    }
    return f;
}

// SIGN controlled by exactly TWO register slots (a CZ between two register-resident qubits)
template <typename V, int A, int B>
__device__ __forceinline__ void op_sign_slot2(V (&v)[kRegAmps]) {
#pragma unroll
    for (int j = 0; j < kRegAmps; ++j) if ((j & (1 << A)) && (j & (1 << B))) { v[j].x = flip_sign(v[j].x); v[j].y = flip_sign(v[j].y); }
}

// diagonal ops with an arbitrary register-slot control mask (rc == 0: all 16 amplitudes)
template <typename V, typename R>
__device__ __forceinline__ void op_phase_mask(V (&v)[kRegAmps], const R t, const R s, const uint32_t rc) {
#pragma unroll
    for (int j = 0; j < kRegAmps; ++j) if ((j & rc) == rc) phase_inplace<V, R>(v[j], t, s);
}
template <typename V>
__device__ __forceinline__ void op_sign_mask(V (&v)[kRegAmps], const uint32_t rc) {
#pragma unroll
    for (int j = 0; j < kRegAmps; ++j) if ((j & rc) == rc) { v[j].x = flip_sign(v[j].x); v[j].y = flip_sign(v[j].y); }
}

template <typename V, typename R>
__device__ __forceinline__ void op_scale(V (&v)[kRegAmps], const R s) {
#pragma unroll
    for (int j = 0; j < kRegAmps; ++j) { v[j].x *= s; v[j].y *= s; }
}

#define QSV_DISPATCH_TB(tb_, CALL)                              \
    switch (tb_) {                                              \
        case 0: { constexpr int TB = 0; CALL; } break;          \
        case 1: { constexpr int TB = 1; CALL; } break;          \
        case 2: { constexpr int TB = 2; CALL; } break;          \
        default: { constexpr int TB = 3; CALL; } break;         \
    }

// Straightforward interpreter of one op (used by the one-CTA-per-tile kernel and as the rare
// path of the ring kernel).  Thread-fixed controls were already checked by the caller.
// `fl` / `sm`: pre-op flags of a mixing op and this thread's pre-sign mask (see op_pre).
template <typename V, typename R>
__device__ __forceinline__ void apply_reg_op(V (&v)[kRegAmps], const int kind, const int tb,
                                             const uint32_t rc, const double *__restrict__ m,
                                             const uint32_t fl = 0u, const int sm = 0) {
    switch (kind) {
        case QSV_OP_HAD: { const R tp = (R)m[2], sp = (R)m[3];
                           QSV_DISPATCH_TB(tb, (op_pre<V, R, TB>(v, fl, sm, tp, sp), op_had<V, TB>(v))); } break;
        case QSV_OP_ROT: { const R t = (R)m[0], s = (R)m[1], tp = (R)m[2], sp = (R)m[3];
                           QSV_DISPATCH_TB(tb, (op_pre<V, R, TB>(v, fl, sm, tp, sp), op_rot<V, R, TB, true>(v, t, s, rc))); } break;
        case QSV_OP_XSWAP: QSV_DISPATCH_TB(tb, (op_xswap<V, TB, true>(v, rc))); break;
        case QSV_OP_YSWAP: QSV_DISPATCH_TB(tb, (op_yswap<V, TB, true>(v, rc))); break;
        case QSV_OP_PHASE: op_phase_mask<V, R>(v, (R)m[0], (R)m[1], rc); break;
        case QSV_OP_SIGN: op_sign_mask<V>(v, rc); break;
        case QSV_OP_SCALE: op_scale<V, R>(v, (R)m[0]); break;
        default: break;
    }
}

// GF(2)-linear swizzle of a tile index: fold every W-bit group above the lowest into the
// lowest W bits (W = log2(128 B / sizeof(element)): 3 for 16-byte, 4 for 8-byte elements), so
// a quarter-warp (half-warp) of 128-bit (64-bit) shared accesses hits distinct bank groups
// whenever its low thread bits drive tile positions with distinct residues mod W.
template <int W> __host__ __device__ __forceinline__ uint32_t tile_swizzle(uint32_t x) {
    constexpr uint32_t M = (1u << W) - 1u;
    uint32_t f = x;
#pragma unroll
    for (int s = W; s < 16; s += W) f ^= (x >> s) & M;
    return f;
}
