/* ORACLE — plain-C restatement of the reference's pair / quad update (complex128).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/ref_dense.py header for the import rule).
 * Parity status: PINNED through tests/test_oracle_golden.py (this library is compared
 * with oracle/ref_dense.py, which is itself pinned to reference outputs).
 *
 * Restates, loop by loop, what the reference does with NumPy index arrays:
 *   oracle_apply_1q : wenbo_engine/kernel/ref_dense.py:13-23 (= cpu_scalar.py:21-32)
 *                     a' = U00*a + U01*b ; b' = U10*a + U11*b over pairs at stride 2^q
 *   oracle_apply_2q : wenbo_engine/kernel/ref_dense.py:26-41 (= cpu_scalar.py:35-47)
 *                     v' = U v over quads, row index = 2*bit(qa) + bit(qb)
 * No index arrays are materialised, so n = 30 (16 GiB) fits a 64 GB host; OpenMP over
 * the outer loop uses all host cores (the "all the host threads" CPU baseline).
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared -fPIC).
 */
#include <complex.h>
#include <stddef.h>
#include <stdint.h>

typedef double complex c128;

static inline uint64_t insert_zero(uint64_t x, int pos) {
    uint64_t low = x & ((1ULL << pos) - 1);
    return ((x >> pos) << (pos + 1)) | low;
}

/* psi: 2^n amplitudes, interleaved re/im. U: 2x2 row-major, re/im interleaved. */
void oracle_apply_1q(double *psi_, int n, int q, const double *U_) {
    c128 *psi = (c128 *)psi_;
    const c128 u00 = U_[0] + U_[1] * I, u01 = U_[2] + U_[3] * I;
    const c128 u10 = U_[4] + U_[5] * I, u11 = U_[6] + U_[7] * I;
    const int64_t pairs = (int64_t)1 << (n - 1);
    const uint64_t step = 1ULL << q;
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < pairs; ++p) {
        uint64_t i0 = insert_zero((uint64_t)p, q);
        c128 a = psi[i0], b = psi[i0 | step];
        psi[i0] = u00 * a + u01 * b;
        psi[i0 | step] = u10 * a + u11 * b;
    }
}

void oracle_apply_2q(double *psi_, int n, int qa, int qb, const double *U_) {
    c128 *psi = (c128 *)psi_;
    c128 U[4][4];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c)
            U[r][c] = U_[2 * (4 * r + c)] + U_[2 * (4 * r + c) + 1] * I;
    const int lo = qa < qb ? qa : qb, hi = qa < qb ? qb : qa;
    const uint64_t ma = 1ULL << qa, mb = 1ULL << qb;
    const int64_t quads = (int64_t)1 << (n - 2);
#pragma omp parallel for schedule(static)
    for (int64_t p = 0; p < quads; ++p) {
        uint64_t base = insert_zero(insert_zero((uint64_t)p, lo), hi);
        uint64_t idx[4] = {base, base | mb, base | ma, base | ma | mb};
        c128 v[4], w[4];
        for (int k = 0; k < 4; ++k) v[k] = psi[idx[k]];
        for (int r = 0; r < 4; ++r) {
            c128 acc = U[r][0] * v[0];
            for (int c = 1; c < 4; ++c) acc += U[r][c] * v[c];
            w[r] = acc;
        }
        for (int k = 0; k < 4; ++k) psi[idx[k]] = w[k];
    }
}

void oracle_init_zero(double *psi, int n) {
    const int64_t len = (int64_t)2 << n;
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < len; ++i) psi[i] = 0.0;
    psi[0] = 1.0;
}

double oracle_norm2(const double *psi, int n) {
    const int64_t len = (int64_t)2 << n;
    double acc = 0.0;
#pragma omp parallel for reduction(+ : acc) schedule(static)
    for (int64_t i = 0; i < len; ++i) acc += psi[i] * psi[i];
    return acc;
}
