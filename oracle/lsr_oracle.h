/*
 * lsr_oracle.h -- CPU ORACLE for the LambdaSNARK-R prover hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under lambda_snark_r_b200/ may include,
 * link or call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and only as the checker or as
 * the timed CPU baseline -- never as the product.
 *
 * What it restates (paths relative to /root/reference):
 *   - cpp-core/src/ntt.cpp:30-119     C-ABI semantics of ntt_context_create /
 *                                     ntt_forward / ntt_inverse / ntt_mul_pointwise
 *   - Microsoft SEAL 4.1 (vcpkg `seal >= 4.1.2`, baseline 74e6536..., NOT
 *     vendored under /root/reference): util/ntt.cpp (NTTTables::initialize),
 *     util/dwthandler.h (transform_to_rev / transform_from_rev),
 *     util/numth.cpp (try_minimal_primitive_root), util/uintarithsmallmod.h
 *     (multiply_uint_mod, MultiplyUIntModOperand).  Restated from the
 *     published algorithm.
 *   - cpp-core/src/utils.cpp:24-121   CDT table + (u1,u2) -> sample mapping
 *   - cpp-core/src/commitment.cpp:44-60,138-164,200-276  container layout,
 *     truncation, verify/linear-combine semantics (the commitment ARITHMETIC is
 *     the Module-LWE definition of DESIGN.md section 3 -- the reference's SEAL
 *     BFV ciphertext is randomised and cannot be reproduced).
 *
 * Parity pinning status:
 *   - sampler (CDT + mapping): PINNED against the reference's own utils.cpp
 *     compiled into oracle/_ref (tests/test_oracle_pinning.py).
 *   - NTT: pinned against the KATs SURVEY.md 8c derived from the SEAL spec,
 *     against the closed form out[i] = f(psi^(2*brv(i)+1)) and against the
 *     reference's own test assertions (round trip, 2*3=6).  No SEAL binary is
 *     available, so NTT byte parity with a SEAL build is "parity unpinned".
 *   - commitment bytes / proofs: "parity unpinned" (and unpinnable).
 */
#ifndef LSR_ORACLE_H
#define LSR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- arithmetic */
uint64_t lsro_mulmod(uint64_t a, uint64_t b, uint64_t q);
uint64_t lsro_powmod(uint64_t a, uint64_t e, uint64_t q);
int      lsro_is_prime(uint64_t q);
/* smallest primitive 2n-th root of unity mod q; 0 if none (SEAL numth.cpp) */
uint64_t lsro_min_primitive_root(uint64_t q, uint64_t two_n);

/* ---------------------------------------------------------------------- NTT */
typedef struct lsro_ntt lsro_ntt;
lsro_ntt *lsro_ntt_create(uint64_t q, uint32_t n);      /* NULL like ntt.cpp:30-70 */
void      lsro_ntt_free(lsro_ntt *c);
uint64_t  lsro_ntt_psi(const lsro_ntt *c);
uint64_t  lsro_ntt_modulus(const lsro_ntt *c);
uint32_t  lsro_ntt_degree(const lsro_ntt *c);
/* table access for cross-checks: which = 0 rp, 1 rp_shoup, 2 irp, 3 irp_shoup */
const uint64_t *lsro_ntt_table(const lsro_ntt *c, int which);
int  lsro_ntt_forward(const lsro_ntt *c, uint64_t *x, uint32_t n);   /* 0 / -1 */
int  lsro_ntt_inverse(const lsro_ntt *c, uint64_t *x, uint32_t n);
void lsro_ntt_mul_pointwise(const lsro_ntt *c, uint64_t *r, const uint64_t *a,
                            const uint64_t *b, uint32_t n);
/* batched (OpenMP over polynomials) -- the CPU baseline legs */
int  lsro_ntt_forward_batch(const lsro_ntt *c, uint64_t *x, size_t batch, int threads);
int  lsro_ntt_inverse_batch(const lsro_ntt *c, uint64_t *x, size_t batch, int threads);
void lsro_ntt_mul_pointwise_batch(const lsro_ntt *c, uint64_t *r, const uint64_t *a,
                                  const uint64_t *b, size_t total, int threads);

/* ------------------------------------------------------------------ sampler */
/* utils.cpp:26-75; returns number of entries written (<= cap), 0 on bad sigma */
size_t  lsro_cdt_build(double sigma, uint64_t *cdf, size_t cap);
/* utils.cpp:95-121 with the two random_u64 draws supplied by the caller */
int64_t lsro_cdt_sample(const uint64_t *cdf, size_t count, uint64_t u1, uint64_t u2);

/* deterministic sample_gaussian: sample i uses 64-bit draws (2i, 2i+1) of the
 * ChaCha stream keyed by seed32 (domain 0x05), 8 draws per block */
int lsro_sample_gaussian_seeded(uint64_t *out, size_t len, double sigma, const uint8_t seed32[32]);

/* ChaCha block (rounds = LSRO_CHACHA_ROUNDS), RFC 7539 quarter round */
#define LSRO_CHACHA_ROUNDS 8
void lsro_chacha_block(const uint32_t key[8], uint32_t w12, uint32_t w13,
                       uint32_t w14, uint32_t w15, uint32_t out[16]);

/* --------------------------------------------------------------- commitment */
typedef struct lsro_lwe lsro_lwe;
/* modulus_req/n/k/sigma as PublicParams; seed32 = 32-byte context seed */
lsro_lwe *lsro_lwe_create(uint64_t modulus_req, uint32_t n, uint32_t k, double sigma,
                          const uint8_t seed32[32]);
void      lsro_lwe_free(lsro_lwe *c);
uint64_t  lsro_lwe_modulus(const lsro_lwe *c);      /* ring modulus actually used */
uint64_t  lsro_lwe_plain_modulus(const lsro_lwe *c);
uint64_t  lsro_lwe_delta(const lsro_lwe *c);
uint32_t  lsro_lwe_words(const lsro_lwe *c);        /* 1 + k*n */
/* A-hat (NTT domain, [k][k][n]) and trapdoor z-hat ([k-1][n]) for cross-checks */
const uint64_t *lsro_lwe_matrix(const lsro_lwe *c);
const uint64_t *lsro_lwe_trapdoor(const lsro_lwe *c);
/* sampled s,e of a commitment (two's complement), each [k][n]; for tests */
void lsro_lwe_sample_se(const lsro_lwe *c, uint64_t seed, int64_t *s, int64_t *e);
/* out = [byte_len, t row-major]; message truncated to n (commitment.cpp:146-149) */
int  lsro_lwe_commit(const lsro_lwe *c, const uint64_t *msg, size_t msg_len,
                     uint64_t seed, uint64_t *out_words);
/* explicit mode: s, e supplied ([k][n] two's complement, any int64, reduced mod q) */
int  lsro_lwe_commit_explicit(const lsro_lwe *c, const uint64_t *msg, size_t msg_len,
                              const int64_t *s, const int64_t *e, uint64_t *out_words);
int  lsro_lwe_commit_batch(const lsro_lwe *c, const uint64_t *msgs, size_t msg_len,
                           const uint64_t *seeds, size_t count, uint64_t *out_words,
                           int threads);
/* 1 / 0 / -1 exactly as commitment.cpp:200-232 (opening ignored) */
int  lsro_lwe_verify(const lsro_lwe *c, const uint64_t *comm_words, size_t comm_len,
                     const uint64_t *msg, size_t msg_len);
/* sum coeffs[i]*comms[i]; comms[i]==NULL skipped; 0 ok / -1 (commitment.cpp:234-276) */
int  lsro_lwe_linear_combine(const lsro_lwe *c, const uint64_t *const *comms,
                             const size_t *comm_lens, const uint64_t *coeffs,
                             size_t count, uint64_t *out_words);

int lsro_max_threads(void);

/* ------------------------------------------------- quotient pipeline
 * lsr_oracle_quotient.c: rust-api/lambda-snark/src/ntt.rs:117-201 (cyclic transform, natural order)
 * and r1cs.rs:474-503, 995-1065 (quotient by X^m - 1), for sizes oracle/quotient.py cannot reach.  */
int lsro_cyclic_ntt_forward(uint64_t *data, size_t n, uint64_t q, uint64_t omega);
int lsro_cyclic_ntt_inverse(uint64_t *data, size_t n, uint64_t q, uint64_t omega);
/* rows/cols/vals[3]: the entries of A, B, C; out[m]; returns 0, 1 (witness does not satisfy), -1 */
int lsro_r1cs_quotient(size_t m, size_t cols, const uint32_t *rows_idx[3], const uint32_t *cols_idx[3],
                       const uint64_t *vals[3], const size_t nnz[3], const uint64_t *witness, uint64_t q,
                       uint64_t omega, uint64_t omega2, uint64_t *out);

#ifdef __cplusplus
}
#endif
#endif
