/*
 * lsr_oracle_quotient.c -- CPU ORACLE for SURVEY row N1 at sizes the Python restatement
 * (oracle/quotient.py) cannot reach.  TEST INFRASTRUCTURE ONLY (see lsr_oracle.h).
 *
 * Restates, from /root/reference:
 *   rust-api/lambda-snark/src/ntt.rs:117-160   ntt_forward (bit-reversal + radix-2 DIT, running twiddle)
 *   rust-api/lambda-snark/src/ntt.rs:181-201   ntt_inverse (forward with omega^-1, then * n^-1)
 *   rust-api/lambda-snark/src/sparse_matrix.rs:259-289  mul_vec (acc += (val % q) * (v[col] % q) mod q)
 *   rust-api/lambda-snark/src/r1cs.rs:474-503  compute_quotient_poly
 *   rust-api/lambda-snark/src/r1cs.rs:995-1065 poly_div_vanishing, use_ntt = true (divide by X^m - 1)
 * One deliberate difference: A_z * B_z is formed with size-2m transforms instead of the reference's
 * O(m^2) poly_mul (r1cs.rs:846-863).  The product of two polynomials is unique, so the numerator --
 * and with it the quotient -- is the same coefficient vector; tests/test_oracle_quotient.py pins this
 * file against the schoolbook Python restatement for every m the latter finishes.
 *
 * Parity pinning: the cyclic transform is pinned by the reference's own unit tests
 * (ntt.rs:284-347: [1,2] -> [3, q-1], sums, round trips n = 2..1024).
 */
#include <stdlib.h>
#include <string.h>

#include "lsr_oracle.h"

typedef unsigned __int128 u128;

static inline uint64_t mulm(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128)a * b) % q); }
static inline uint64_t addm(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128)a + b) % q); }
static inline uint64_t subm(uint64_t a, uint64_t b, uint64_t q) { return a >= b ? a - b : (uint64_t)((u128)a + q - b); }

static uint64_t powm(uint64_t a, uint64_t e, uint64_t q) {
    uint64_t r = 1 % q;
    a %= q;
    while (e) {
        if (e & 1) r = mulm(r, a, q);
        a = mulm(a, a, q);
        e >>= 1;
    }
    return r;
}

/* ntt.rs:82-96 bit_reverse_permutation */
static void bit_reverse(uint64_t *d, size_t n) {
    unsigned bits = 0;
    while (((size_t)1 << bits) < n) ++bits;
    for (size_t i = 0; i < n; i++) {
        size_t j = 0;
        for (unsigned b = 0; b < bits; b++) j |= ((i >> b) & 1) << (bits - 1 - b);
        if (i < j) { uint64_t t = d[i]; d[i] = d[j]; d[j] = t; }
    }
}

/* ntt.rs:117-160.  Inputs are reduced first (the device path reduces raw words on load). */
int lsro_cyclic_ntt_forward(uint64_t *data, size_t n, uint64_t q, uint64_t omega) {
    if (!data || n == 0 || (n & (n - 1))) return -1;
    for (size_t i = 0; i < n; i++) data[i] %= q;
    if (n == 1) return 0;
    bit_reverse(data, n);
    for (size_t m = 2; m <= n; m <<= 1) {
        const size_t half = m >> 1;
        const uint64_t omega_m = powm(omega, (uint64_t)(n / m), q);
        for (size_t k = 0; k < n; k += m) {
            uint64_t w = 1;
            for (size_t j = 0; j < half; j++) {
                const uint64_t t = mulm(data[k + j + half], w, q);
                const uint64_t u = data[k + j];
                data[k + j] = addm(u, t, q);
                data[k + j + half] = subm(u, t, q);
                w = mulm(w, omega_m, q);
            }
        }
    }
    return 0;
}

/* ntt.rs:181-201 */
int lsro_cyclic_ntt_inverse(uint64_t *data, size_t n, uint64_t q, uint64_t omega) {
    if (!data || n == 0 || (n & (n - 1))) return -1;
    if (n == 1) { data[0] %= q; return 0; }
    const uint64_t omega_inv = powm(omega, q - 2, q);
    if (lsro_cyclic_ntt_forward(data, n, q, omega_inv)) return -1;
    const uint64_t n_inv = powm((uint64_t)n % q, q - 2, q);
    for (size_t i = 0; i < n; i++) data[i] = mulm(data[i], n_inv, q);
    return 0;
}

/* sparse_matrix.rs:259-289; entries (row, col, value) in any order */
static void mul_vec(uint64_t *out, size_t rows, const uint32_t *row, const uint32_t *col, const uint64_t *val,
                    size_t nnz, const uint64_t *v, uint64_t q) {
    memset(out, 0, rows * sizeof(uint64_t));
    for (size_t e = 0; e < nnz; e++)
        out[row[e]] = addm(out[row[e]], mulm(val[e] % q, v[col[e]] % q, q), q);
}

/*
 * r1cs.rs:474-503 (NTT path).  m = rows (power of two), omega a primitive m-th root, omega2 a primitive
 * 2m-th root (any: it only carries the product).  out[m] receives Q zero-padded.
 * Returns 0, 1 when the witness does not satisfy the constraints (r1cs.rs:477-481 / non-zero
 * remainder :1054-1060), -1 on bad arguments.
 */
int lsro_r1cs_quotient(size_t m, size_t cols, const uint32_t *rows_idx[3], const uint32_t *cols_idx[3],
                       const uint64_t *vals[3], const size_t nnz[3], const uint64_t *witness, uint64_t q,
                       uint64_t omega, uint64_t omega2, uint64_t *out) {
    (void)cols;
    if (m == 0 || (m & (m - 1)) || !out) return -1;
    uint64_t *ev = (uint64_t *)calloc(3 * m, sizeof(uint64_t));
    uint64_t *big = (uint64_t *)calloc(3 * 2 * m, sizeof(uint64_t));
    if (!ev || !big) { free(ev); free(big); return -1; }
    int status = 0;
    for (int k = 0; k < 3; k++) mul_vec(ev + k * m, m, rows_idx[k], cols_idx[k], vals[k], nnz[k], witness, q);
    for (size_t i = 0; i < m; i++)
        if (mulm(ev[i], ev[m + i], q) != ev[2 * m + i]) status = 1;               /* is_satisfied */
    if (m == 1) {                                                                   /* degree-0 numerator */
        out[0] = 0;
        free(ev); free(big);
        return status;
    }
    for (int k = 0; k < 3; k++) {                                                   /* lagrange_interpolate_ntt */
        lsro_cyclic_ntt_inverse(ev + k * m, m, q, omega);
        memcpy(big + (size_t)k * 2 * m, ev + k * m, m * sizeof(uint64_t));
        lsro_cyclic_ntt_forward(big + (size_t)k * 2 * m, 2 * m, q, omega2);
    }
    for (size_t i = 0; i < 2 * m; i++)                                              /* A*B - C on 2m points */
        big[i] = subm(mulm(big[i], big[2 * m + i], q), big[4 * m + i], q);
    lsro_cyclic_ntt_inverse(big, 2 * m, q, omega2);                                 /* numerator, degree <= 2m-2 */
    /* poly_div_vanishing: leading coefficient 1, divisor X^m - 1: q_i = rem[i+m]; rem[i] += q_i */
    for (size_t i = m; i-- > 0;) {
        const uint64_t qc = big[i + m];
        out[i] = qc;
        big[i + m] = 0;
        big[i] = addm(big[i], qc, q);
    }
    for (size_t i = 0; i < m; i++)
        if (big[i] != 0) status = 1;
    free(ev); free(big);
    return status;
}
