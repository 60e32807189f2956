/*
 * lambda_snark_b200.h -- C ABI of the B200-native prover hot path.
 *
 * Part 1 is the drop-in surface: every symbol rust-api/lambda-snark-sys binds
 * today (bindgen allowlist `lwe_.*`, `ntt_.*`, `lambda_snark_r1cs_.*`,
 * rust-api/lambda-snark-sys/build.rs:197-199) plus `sample_gaussian`, with the
 * reference's exact signatures, struct layouts, ownership and error rules.
 * Each declaration cites the reference interface it replaces (paths relative
 * to the reference repo).  Part 2 holds the batched / device-pointer
 * extensions a GPU back-end needs to amortise launch and PCIe cost.
 *
 * Plain C: pointers and sizes only, no CUDA or torch types.  `void* stream`
 * is a cudaStream_t passed opaquely (NULL = the context's own stream).
 *
 * The legacy include paths <lambda_snark/{types,ntt,commitment,utils,r1cs}.h>
 * are thin forwarding headers onto this file, so reference-side C++ (e.g.
 * cpp-core/tests/test_ntt.cpp) compiles unchanged.
 */
#ifndef LAMBDA_SNARK_B200_H
#define LAMBDA_SNARK_B200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
#define LSR_NOEXCEPT noexcept
extern "C" {
#else
#define LSR_NOEXCEPT
#endif

/* ===================================================================== */
/* Part 1a -- FFI types (cpp-core/include/lambda_snark/types.h:27-78)     */
/* ===================================================================== */

typedef struct LweContext LweContext;            /* types.h:27, opaque */
typedef struct NttContext NttContext;            /* ntt.h:24,   opaque */

/* types.h:36-39.  data[0] = payload byte length, data[1..] = payload.
 * Here the payload is t = A*s + e + enc(m), k*n little-endian u64 words,
 * row-major, natural coefficient order, every word in [0, q).  `data` is
 * host memory owned by the library until lwe_commitment_free.            */
typedef struct {
    uint64_t* data;
    size_t    len;
} LweCommitment;

/* types.h:44-47 */
typedef struct {
    uint64_t* randomness;
    size_t    rand_len;
} LweOpening;

/* types.h:52-55 */
typedef enum {
    PROFILE_SCALAR_A = 0,
    PROFILE_RING_B   = 1
} ProfileType;

/* types.h:60-67 (32 bytes).  Unlike the reference, which stores but ignores
 * modulus / module_rank / sigma (commitment.cpp:108-111), all four numeric
 * fields are honoured; see lwe_context_create.                            */
typedef struct {
    ProfileType profile;
    uint32_t    security_level;
    uint64_t    modulus;
    uint32_t    ring_degree;
    uint32_t    module_rank;
    double      sigma;
} PublicParams;

/* types.h:72-78 */
typedef enum {
    LAMBDA_SNARK_OK                 = 0,
    LAMBDA_SNARK_ERR_NULL_PTR       = 1,
    LAMBDA_SNARK_ERR_INVALID_PARAMS = 2,
    LAMBDA_SNARK_ERR_ALLOC_FAILED   = 3,
    LAMBDA_SNARK_ERR_CRYPTO_FAILED  = 4
} LambdaSnarkError;

/* ===================================================================== */
/* Part 1b -- NTT (cpp-core/include/lambda_snark/ntt.h, src/ntt.cpp)      */
/* ===================================================================== */

/* ntt.h:34 / ntt.cpp:30-70.  NULL when n == 0, n not a power of two, n
 * outside [2, 2^17], q outside [2, 2^61), or no primitive 2n-th root of
 * unity exists mod q (SEAL NTTTables throws -> NULL).  Tables use SEAL's
 * minimal primitive root.  Also NULL if no CUDA device is usable.         */
NttContext* ntt_context_create(uint64_t q, uint32_t n) LSR_NOEXCEPT;

/* ntt.h:41 / ntt.cpp:72-74.  NULL-safe. */
void ntt_context_free(NttContext* ctx) LSR_NOEXCEPT;

/* ntt.h:55-59 / ntt.cpp:76-89.  In-place negacyclic forward NTT, natural
 * order in, bit-reversed order out, output in [0,q).  0 on success, -1 on
 * NULL ctx / NULL coeffs / n != degree (or a CUDA failure).                */
int ntt_forward(const NttContext* ctx, uint64_t* coeffs, uint32_t n) LSR_NOEXCEPT;

/* ntt.h:69-73 / ntt.cpp:91-104.  Exact inverse of ntt_forward (n^-1 folded in). */
int ntt_inverse(const NttContext* ctx, uint64_t* evals, uint32_t n) LSR_NOEXCEPT;

/* ntt.h:86-92 / ntt.cpp:106-119.  result[i] = a[i]*b[i] mod q, exact for any
 * u64 inputs; result may alias a or b; NO check of n against the context;
 * returns silently on any NULL argument.                                   */
void ntt_mul_pointwise(const NttContext* ctx, uint64_t* result, const uint64_t* a,
                       const uint64_t* b, uint32_t n) LSR_NOEXCEPT;

/* ===================================================================== */
/* Part 1c -- commitment (include/lambda_snark/commitment.h, src/commitment.cpp) */
/* ===================================================================== */

/* commitment.h:31 / commitment.cpp:102-132.  NULL on NULL params or
 * unusable parameters.  Context key material (matrix A, trapdoor) is drawn
 * from the OS entropy source, as the reference's SEAL keygen is.  The ring
 * modulus is params->modulus when it is an NTT-friendly prime for
 * ring_degree, else the built-in 17592169062401 (n <= 4096) /
 * 17592180539393 (n <= 131072) -- the reference ignores the field entirely.*/
LweContext* lwe_context_create(const PublicParams* params) LSR_NOEXCEPT;

/* commitment.h:38 / commitment.cpp:134-136.  NULL-safe; zeroises key material. */
void lwe_context_free(LweContext* ctx) LSR_NOEXCEPT;

/* commitment.h:58-63 / commitment.cpp:138-164.  Message is padded with zeros
 * or silently truncated to ring_degree words (commitment.cpp:146-149).
 *
 * MESSAGE RANGE.  The commitment binds every message word MODULO THE PLAINTEXT
 * MODULUS p = lsr_lwe_plain_modulus(ctx) (204800 for the default q): the slot
 * holds Delta * (word mod p).  A word >= p is NOT rejected by default, because
 * the reference's callers pass field elements (44 .. 64 bits) and SEAL's
 * BatchEncoder range check (commitment.cpp:152) exists in debug builds only;
 * lwe_verify_opening accordingly compares with (word mod p).  Callers that
 * need the whole word bound either commit its base-p digits
 * (lsr_lwe_commit_digits_batch_device; lsr_prover_commit_quotient does so by
 * itself) or switch the context to strict mode (lsr_lwe_set_strict_messages),
 * in which a word >= p makes lwe_commit return NULL, as SEAL's debug build does.
 *
 * SEED.  seed == 0 draws fresh randomness (commitment.h:52 "0 = random"); any
 * other seed makes the commitment a deterministic function of (context,
 * message, seed): s and e depend on (context key, seed) ONLY, so two messages
 * committed under one non-zero seed differ by exactly Delta * (m1 - m2) in the
 * last row -- never reuse a non-zero seed for different messages (the
 * reference ignores the argument and always randomises).
 * NULL on NULL ctx / NULL message / failure.                                */
LweCommitment* lwe_commit(LweContext* ctx, const uint64_t* message, size_t msg_len,
                          uint64_t seed) LSR_NOEXCEPT;

/* commitment.h:70 / commitment.cpp:166-177.  NULL-safe; zeroises then frees. */
void lwe_commitment_free(LweCommitment* comm) LSR_NOEXCEPT;

/* commitment.h:78 / commitment.cpp:179-198.  Deep copy; NULL on NULL/empty. */
LweCommitment* lwe_commitment_clone(const LweCommitment* comm) LSR_NOEXCEPT;

/* commitment.h:94-100 / commitment.cpp:200-232.  Opens the commitment with the
 * context's trapdoor (the reference decrypts with the SEAL secret key) and
 * compares the first msg_len decoded words with `message` mod p (see lwe_commit,
 * MESSAGE RANGE); the differences are OR-folded without a branch on the data,
 * as commitment.cpp:223-226 does.
 * `opening` is ignored, as in the reference (commitment.cpp:205).
 * 1 = match, 0 = mismatch (also msg_len > ring_degree), -1 = NULL argument
 * or malformed container.                                                  */
int lwe_verify_opening(const LweContext* ctx, const LweCommitment* commitment,
                       const uint64_t* message, size_t msg_len,
                       const LweOpening* opening) LSR_NOEXCEPT;

/* commitment.h:113-118 / commitment.cpp:234-276.  sum_i c'_i * C_i with c'_i the
 * centred representative of coeffs[i] mod p, p = plaintext modulus
 * (commitment.cpp:90 reduces coefficients mod the plain modulus; the result
 * commits to sum_i coeffs[i] * m_i mod p either way).
 * COEFFICIENT BOUND.  With a 44-bit ring modulus the noise budget is small: the
 * result opens correctly while sum_i |c'_i| <= lsr_lwe_lincomb_budget(ctx)
 * (about 3 * 10^3 for n = 4096, k = 2, sigma = 3.19; inputs assumed fresh).
 * Beyond that the call returns NULL instead of an undecodable commitment
 * (the reference, with a 109-bit ciphertext modulus, accepts any coefficient).
 * NULL entries of `commitments` are skipped; NULL result on NULL arguments,
 * count == 0, a malformed container, or when every entry was NULL.          */
LweCommitment* lwe_linear_combine(const LweContext* ctx, const LweCommitment** commitments,
                                  const uint64_t* coeffs, size_t count) LSR_NOEXCEPT;

/* ===================================================================== */
/* Part 1d -- sampler (include/lambda_snark/utils.h:27, src/utils.cpp:132-146) */
/* ===================================================================== */

/* CDT discrete Gaussian; samples are two's-complement int64 stored in u64.
 * -1 on NULL output, len == 0, sigma <= 0 or non-finite.  Entropy: a fresh
 * 256-bit key from the OS per call, expanded on the device.                */
int sample_gaussian(uint64_t* output, size_t len, double sigma) LSR_NOEXCEPT;

/* ===================================================================== */
/* Part 1e -- R1CS handle (include/lambda_snark/r1cs.h:38-79, src/ffi.cpp:27-105),
 * declared by hand in rust-api/lambda-snark-core/src/r1cs.rs:121-141.
 * Host-only (a handful of sparse mat-vecs), NTL-free.                     */
/* ===================================================================== */

typedef struct {
    uint32_t row;
    uint32_t col;
    uint64_t value;
} SparseEntry;

typedef struct {
    SparseEntry* entries;
    size_t       n_entries;
    uint32_t     n_rows;
    uint32_t     n_cols;
} SparseMatrix;

/* r1cs.h:62-69 -- named in lambda-snark-sys's bindgen allowlist (build.rs:201) */
typedef struct {
    SparseMatrix A;
    SparseMatrix B;
    SparseMatrix C;
    uint32_t n_vars;
    uint32_t n_public_inputs;
    uint32_t n_constraints;
} R1CSConstraintSystem;

typedef struct {
    uint64_t* values;
    size_t    len;
} R1CSWitness;

LambdaSnarkError lambda_snark_r1cs_create(const SparseMatrix* A, const SparseMatrix* B,
                                          const SparseMatrix* C, uint64_t modulus,
                                          void** out_r1cs);
LambdaSnarkError lambda_snark_r1cs_validate_witness(void* r1cs, const R1CSWitness* witness,
                                                    bool* out_valid);
void     lambda_snark_r1cs_free(void* r1cs);
uint32_t lambda_snark_r1cs_num_constraints(void* r1cs);
uint32_t lambda_snark_r1cs_num_variables(void* r1cs);

/* ---- Lean 4 export (cpp-core/src/lean_ffi.cpp:152-314; declared by hand in the reference, no header).
 * export_vk_to_lean:     "<nCons, nVars, nPublic, q, SparseMatrix.mk r c [(row, col, value), ...], B, C>" in
 *                        U+27E8/U+27E9 brackets; q is params->modulus.  Bytes written (without the NUL) or -1.
 * export_params_to_lean: "{ n := 4096, k := 2, q := 12289, sigma := 3.2, lambda := 128 }" with the Greek letters
 *                        as UTF-8 and sigma printed with one decimal (std::fixed, setprecision(1)).
 * export_seal_*:         SEAL-specific; this library holds no SEAL object, so both fail with -1 exactly as the
 *                        reference does when its SEAL context is absent (lean_ffi.cpp:251-254, 293-296).       */
int export_vk_to_lean(const R1CSConstraintSystem* r1cs, const PublicParams* params, char* out_buffer,
                      size_t buffer_size) LSR_NOEXCEPT;
int export_params_to_lean(const PublicParams* params, char* out_buffer, size_t buffer_size) LSR_NOEXCEPT;
int export_seal_context_to_lean(const LweContext* ctx, char* out_buffer, size_t buffer_size) LSR_NOEXCEPT;
int export_seal_pubkey_to_lean(const LweContext* ctx, char* out_buffer, size_t buffer_size) LSR_NOEXCEPT;

/* ===================================================================== */
/* Part 2 -- batched and device-pointer extensions (new; no reference      */
/* counterpart: the reference is one polynomial / one commitment per call) */
/* ===================================================================== */

/* Library / device management.  All return 0 on success, -1 on failure.   */
int         lsr_device_count(void) LSR_NOEXCEPT;
int         lsr_set_device(int device) LSR_NOEXCEPT;   /* device new contexts bind to (per thread) */
const char* lsr_version(void) LSR_NOEXCEPT;
const char* lsr_last_error(void) LSR_NOEXCEPT;         /* thread-local diagnostic string */

/* Page-locked host buffers for the batched host-pointer entry points.  They accept any host memory: from ordinary
 * (pageable) memory -- a Rust Vec<u64>, a numpy array -- lwe_commit_batch stages chunks through its own page-locked
 * buffers with a few copy threads (0.38 M commitments/s; the driver's own pageable staging gives 0.10 M/s); from
 * these buffers there is no staging at all: 0.78 M/s, PCIe-bound (tools/latency.py, bench.py e2e).  NULL on failure. */
void* lsr_host_alloc(size_t bytes) LSR_NOEXCEPT;
void  lsr_host_free(void* p) LSR_NOEXCEPT;

/* ---- Final gather over NVLink peer memory (collective C1).  One process per GPU: the gathering rank
 * allocates the destination with lsr_device_alloc and exports it (lsr_peer_export, a 64-byte handle any transport can
 * carry); every other rank maps it (lsr_peer_open) and passes `mapped + its slice offset` as d_out_words of
 * lsr_lwe_commit_batch_device / lsr_prover_commit_quotient_device.  The fused commitment kernel then stores its
 * container rows straight into the peer's HBM through NVLink / NVSwitch: compute and gather are ONE kernel, there is no
 * separate collective and no second pass over the containers.  Buffers come from plain cudaMalloc (a handle names a
 * whole allocation).  lsr_peer_open returns NULL when the two devices have no peer path.                          */
void* lsr_device_alloc(size_t bytes) LSR_NOEXCEPT;
void  lsr_device_free(void* d_ptr) LSR_NOEXCEPT;
int   lsr_peer_export(void* d_ptr, uint8_t handle[64]) LSR_NOEXCEPT;
void* lsr_peer_open(const uint8_t handle[64]) LSR_NOEXCEPT;
int   lsr_peer_close(void* d_mapped) LSR_NOEXCEPT;

/* Integer-multiply roofline denominator: dependency-free mad.wide.u32 (wide=1)
 * or mad.lo.u32 (wide=0) on every SM, timed with CUDA events; result in
 * 10^9 IMAD per second.  sm_mhz_effective (optional) = the SM clock implied by
 * 64 lanes/clk/SM.                                                          */
int lsr_measure_imad_peak(int wide, double* gimad_per_s, double* sm_mhz_effective) LSR_NOEXCEPT;
/* Same for the FP64 pipe (DFMA / DADD / DMUL issue rate, 10^9 thread-instructions/s): the
 * denominator of the FP64-butterfly roofline.                                            */
int lsr_measure_fp64_peak(double* ginst_per_s) LSR_NOEXCEPT;

/* Introspection used by the tests and the host wrappers. */
uint64_t lsr_ntt_modulus(const NttContext* ctx) LSR_NOEXCEPT;
uint32_t lsr_ntt_degree(const NttContext* ctx) LSR_NOEXCEPT;
uint64_t lsr_ntt_root(const NttContext* ctx) LSR_NOEXCEPT;      /* minimal primitive 2n-th root */
int      lsr_ntt_device(const NttContext* ctx) LSR_NOEXCEPT;

/* batch polynomials, contiguous [batch][n], HOST memory, in place */
int ntt_forward_batch(const NttContext* ctx, uint64_t* coeffs, size_t batch) LSR_NOEXCEPT;
int ntt_inverse_batch(const NttContext* ctx, uint64_t* evals, size_t batch) LSR_NOEXCEPT;
int ntt_mul_pointwise_batch(const NttContext* ctx, uint64_t* result, const uint64_t* a,
                            const uint64_t* b, size_t total) LSR_NOEXCEPT;

/* same, DEVICE memory; asynchronous on `stream` */
int lsr_ntt_forward_device(const NttContext* ctx, uint64_t* d_coeffs, size_t batch,
                           void* stream) LSR_NOEXCEPT;
int lsr_ntt_inverse_device(const NttContext* ctx, uint64_t* d_evals, size_t batch,
                           void* stream) LSR_NOEXCEPT;
int lsr_ntt_mul_pointwise_device(const NttContext* ctx, uint64_t* d_result, const uint64_t* d_a,
                                 const uint64_t* d_b, size_t total, void* stream) LSR_NOEXCEPT;

/* Reproducible context: same (params, seed32) -> same matrix A / trapdoor on
 * every rank and on the CPU oracle.                                         */
LweContext* lwe_context_create_seeded(const PublicParams* params,
                                      const uint8_t seed32[32]) LSR_NOEXCEPT;

uint64_t lsr_lwe_modulus(const LweContext* ctx) LSR_NOEXCEPT;          /* ring modulus q in use */
uint64_t lsr_lwe_plain_modulus(const LweContext* ctx) LSR_NOEXCEPT;    /* p */
uint64_t lsr_lwe_delta(const LweContext* ctx) LSR_NOEXCEPT;            /* (q-1)/p */
size_t   lsr_lwe_commitment_words(const LweContext* ctx) LSR_NOEXCEPT; /* 1 + k*n */
/* copies A-hat ([k][k][n], NTT domain) to host memory; for cross-checks */
int      lsr_lwe_copy_matrix(const LweContext* ctx, uint64_t* out) LSR_NOEXCEPT;

/* ---- Quotient pipeline of the prover.  These have no counterpart in the reference's C
 * ABI: they replace Rust code (rust-api/lambda-snark/src/ntt.rs, r1cs.rs) that a maintainer would
 * bind through lambda-snark-sys.
 *
 * Cyclic transform over X^n - 1, natural order in and out: ntt_forward / ntt_inverse of ntt.rs:117-201
 * (coefficients -> [f(omega^0), ..., f(omega^(n-1))]).  q: a prime < 2^61 with n | q-1, or Goldilocks
 * 2^64 - 2^32 + 1 (lambda-snark-core/src/lib.rs:58 NTT_MODULUS).  omega = 0 picks the reference's root:
 * NTT_PRIMITIVE_ROOT^(2^32/n) (lib.rs:78, ntt.rs:214-221), 3^((q-1)/n) for 17592169062401
 * (r1cs.rs:534-547), the minimal primitive n-th root otherwise.  The context is an ordinary NttContext
 * (ntt_context_free releases it; ntt_forward_batch on it returns the bit-reversed order).               */
uint64_t    lsr_reference_root_of_unity(uint64_t q, uint32_t n) LSR_NOEXCEPT;
NttContext* lsr_cyclic_ntt_context_create(uint64_t q, uint32_t n, uint64_t omega) LSR_NOEXCEPT;
int lsr_cyclic_ntt_forward(const NttContext* ctx, uint64_t* coeffs, size_t batch) LSR_NOEXCEPT;
int lsr_cyclic_ntt_inverse(const NttContext* ctx, uint64_t* evals, size_t batch) LSR_NOEXCEPT;

/* Quotient polynomial Q(X) = (A_z(X) B_z(X) - C_z(X)) / (X^m - 1) of the R1CS instance behind `r1cs`
 * (lambda_snark_r1cs_create), A_z, B_z, C_z interpolated over H = {omega^j}: compute_quotient_poly on
 * its NTT path (r1cs.rs:474-503, :746-793, :995-1065).  m = number of constraints, a power of two
 * <= 2^23; matrix values and witness words are reduced as unsigned words (sparse_matrix.rs:279).
 * One witness: `out` receives the coefficients with trailing zeros removed (*out_len >= 1, out_cap >= m);
 * LAMBDA_SNARK_ERR_CRYPTO_FAILED when the witness does not satisfy the constraints.
 * Batch: witnesses [count][witness_len] -> out [count][m] zero-padded, status[count] (0 ok, 1 unsatisfied). */
LambdaSnarkError lsr_r1cs_quotient(void* r1cs, const uint64_t* witness, size_t witness_len, uint64_t omega,
                                   uint64_t* out, size_t out_cap, size_t* out_len) LSR_NOEXCEPT;
int lsr_r1cs_quotient_batch(void* r1cs, const uint64_t* witnesses, size_t witness_len, size_t count,
                            uint64_t omega, uint64_t* out, int* status) LSR_NOEXCEPT;

/* Commitment phase of the prover for a circuit larger than one ring element (BASELINE configs[4]; replaces
 * compute_quotient_poly + Commitment::new of prove_r1cs, rust-api/lambda-snark/src/lib.rs:747-757, where the
 * reference silently truncates Q to ring_degree coefficients -- cpp-core/src/commitment.cpp:146-149).
 * The quotient of every witness is cut into max(1, m / ring_degree) ring elements of ring_degree coefficients
 * (one of m coefficients when m < ring_degree), and every ring element into L = lsr_prover_quotient_planes()
 * DIGIT PLANES: a commitment binds its words modulo p only (lwe_commit, MESSAGE RANGE), so the field elements are
 * committed as their L = ceil(log_p(field modulus)) base-p digits (4 for Goldilocks, 3 for a 44-bit field), which
 * together bind the whole quotient.  A unit is a (ring element j, plane l) pair, unit index u = j * L + l;
 * lsr_prover_quotient_chunks() returns the number of units per witness (called `chunks` below); unit (w, u) is
 * committed with seeds[w * chunks + u] exactly as lwe_commit_batch would commit the digit row
 * (Q_j[x] / p^l) mod p.  Only unit indices [chunk_lo, chunk_hi) are committed -- a rank's
 * slice of a job sharded over GPUs; seeds is always indexed globally ([count][chunks]) so the containers do not
 * depend on the sharding.  Quotient and messages never leave the device.
 * out: [count][chunk_hi - chunk_lo][1 + k n] words; status[count]: 0 ok, 1 witness does not satisfy the
 * constraints (its containers are then commitments to a meaningless quotient and must be discarded).
 * Returns a LambdaSnarkError code.  _device: witnesses, seeds and out are DEVICE pointers on the context's
 * device (status stays on the host); the call synchronises before returning.                                  */
size_t   lsr_prover_quotient_chunks(void* r1cs, const LweContext* ctx) LSR_NOEXCEPT;
uint32_t lsr_prover_quotient_planes(void* r1cs, const LweContext* ctx) LSR_NOEXCEPT;   /* 0: more than 4 would be needed */
int lsr_prover_commit_quotient(void* r1cs, LweContext* ctx, const uint64_t* witnesses, size_t witness_len,
                               size_t count, uint64_t omega, const uint64_t* seeds, size_t chunk_lo,
                               size_t chunk_hi, uint64_t* out, int* status) LSR_NOEXCEPT;
int lsr_prover_commit_quotient_device(void* r1cs, LweContext* ctx, const uint64_t* d_witnesses,
                                      size_t witness_len, size_t count, uint64_t omega, const uint64_t* d_seeds,
                                      size_t chunk_lo, size_t chunk_hi, uint64_t* d_out, int* status) LSR_NOEXCEPT;

/* ---- Fiat-Shamir transcript and polynomial evaluations of the prover on the device.  No counterpart
 * in the reference's C ABI: they replace Rust code (challenge.rs:102-134 Challenge::derive, r1cs.rs:362-373
 * eval_poly, lib.rs:747-809 prove_r1cs) that a maintainer would bind through lambda-snark-sys.
 *
 * lsr_fs_challenge_batch: for i < count,
 *     hash  = SHA3-256("LAMBDA-SNARK-R-FS-v1" || le64(n_public) || public_inputs[i][..] || le64(words) || containers[i][..])
 *     alpha = le64(hash[0..8]) mod modulus                                        -> challenges[i][0], hashes[i][0][0..32)
 *   and, when chain != 0, beta = the same with the single public input alpha (lib.rs:764-767)
 *                                                                                 -> challenges[i][1], hashes[i][1][0..32)
 *   (chain == 0 leaves those zero).  hashes is [count][2][4] 64-bit little-endian lanes = [count][2][32] bytes.
 * lsr_poly_eval_batch: out[p][j] = sum_i coeffs[p][i] * points[p][j]^i mod modulus (modulus < 2^61 or Goldilocks).
 * lsr_prove_r1cs_batch: prove_r1cs for `count` witnesses of one circuit with m <= ring_degree constraints,
 *   interpolation over the roots of unity (the NTT path of the reference).
 *   DOMAIN RESTRICTION: H is always the m-th roots of unity, Z_H = X^m - 1, m a power of two with 2m | modulus - 1.
 *   The reference takes that path only when should_use_ntt() holds (modulus == Goldilocks NTT_MODULUS and m a power
 *   of two, r1cs.rs:386-389); for every other modulus it interpolates over H = {0..m-1} with Z_H = prod (X - i), and
 *   proofs / verdicts of the two domains do not interoperate.  So: with the Goldilocks modulus this is the reference's
 *   proof; with any other NTT-friendly modulus it is the same protocol over a different evaluation domain (prover and
 *   verifier of THIS library agree with each other, not with the reference's baseline-domain verifier).  The same holds
 *   for lsr_verify_r1cs_batch and lsr_r1cs_quotient*.
 *   The single commitment of the reference's proof format (ProofR1CS::commitment_q) binds Q modulo p, like the
 *   reference's own 20-bit BFV plaintext; lsr_prover_commit_quotient commits digit planes instead.
 *   Per witness the commitment container
 *   ([1 + k n] words, seed seeds[w]), challenges (alpha, beta), the two transcript hashes, and evals[w][8] =
 *   {Q(alpha), Q(beta), A_z(alpha), B_z(alpha), C_z(alpha), A_z(beta), B_z(beta), C_z(beta)} -- the field order of
 *   ProofR1CS::new (lib.rs:795-809; the two openings are Q(alpha), Q(beta) again).  status[w] = 1 marks a witness
 *   that does not satisfy the constraints (prove_r1cs returns Err: discard that row).  HOST pointers.             */
int lsr_fs_challenge_batch(const uint64_t* public_inputs, size_t n_public, const uint64_t* containers, size_t words,
                           size_t count, uint64_t modulus, int chain, uint64_t* challenges,
                           uint64_t* hashes) LSR_NOEXCEPT;
int lsr_fs_challenge_batch_device(const uint64_t* d_public_inputs, size_t n_public, const uint64_t* d_containers,
                                  size_t words, size_t count, uint64_t modulus, int chain, uint64_t* d_challenges,
                                  uint64_t* d_hashes, void* stream) LSR_NOEXCEPT;
int lsr_poly_eval_batch(uint64_t modulus, const uint64_t* coeffs, size_t len, size_t polys, const uint64_t* points,
                        size_t npts, uint64_t* out) LSR_NOEXCEPT;
int lsr_prove_r1cs_batch(void* r1cs, LweContext* ctx, const uint64_t* witnesses, size_t witness_len, size_t count,
                         size_t n_public, uint64_t omega, const uint64_t* seeds, uint64_t* containers,
                         uint64_t* challenges, uint64_t* hashes, uint64_t* evals, int* status) LSR_NOEXCEPT;

/* verify_r1cs (lib.rs:1016-1082) for `count` proofs of one circuit with n_constraints = m (a power of two) on the
 * NTT path, Z_H(X) = X^m - 1 (r1cs.rs:424-430): recomputes alpha from (public_inputs[i], containers[i]) and beta from
 * (alpha, containers[i]) on the device, compares them with challenges[i][0..2), and checks
 * Q(x) Z_H(x) = A_z(x) B_z(x) - C_z(x) at x = alpha, beta with evals[i][8] in the order lsr_prove_r1cs_batch writes.
 * results[i] = 1 accept / 0 reject; returns 0, or -1 on bad arguments / device failure.  HOST pointers.           */
int lsr_verify_r1cs_batch(uint64_t n_constraints, uint64_t modulus, const uint64_t* public_inputs, size_t n_public,
                          const uint64_t* containers, size_t words, const uint64_t* challenges, const uint64_t* evals,
                          size_t count, int* results) LSR_NOEXCEPT;

/* Arithmetic of the NTT butterflies (NttContext, and the NttContext inside an
 * LweContext): 0 auto -- FP64-pipe butterflies (exact modular products by
 * error-free fma multiplication) when q < 2^45, else u64 Shoup butterflies;
 * 1 u64 arithmetic only; 2 require FP64 (-1 if q >= 2^45).  Results are
 * bit-identical either way; the switch exists for cross-checks and benchmarks.
 * lsr_ntt_arith returns the policy in effect (1 or 2).                        */
int lsr_ntt_set_arith(NttContext* ctx, int arith) LSR_NOEXCEPT;
int lsr_ntt_arith(const NttContext* ctx) LSR_NOEXCEPT;
int lsr_lwe_set_arith(LweContext* ctx, int arith) LSR_NOEXCEPT;

/* Path selection for lwe_commit_batch*: 0 auto, 1 force the generic
 * multi-kernel path, 2 force the fused kernel (-1 at call time if it does
 * not support the context's (n, k, sigma)).                                 */
int lsr_lwe_set_commit_path(LweContext* ctx, int path) LSR_NOEXCEPT;

/* `count` commitments in one call.  messages: [count][msg_len] (each padded /
 * truncated to n as lwe_commit does); seeds: [count], all non-zero use is the
 * caller's business (0 is NOT replaced here); out: [count][1 + k*n] words in
 * the LweCommitment container layout.  HOST memory (pinned memory makes the
 * copies overlap with the kernels).                                         */
int lwe_commit_batch(LweContext* ctx, const uint64_t* messages, size_t msg_len,
                     const uint64_t* seeds, size_t count, uint64_t* out_words) LSR_NOEXCEPT;

/* same, DEVICE memory; asynchronous on `stream` */
int lsr_lwe_commit_batch_device(LweContext* ctx, const uint64_t* d_messages, size_t msg_len,
                                const uint64_t* d_seeds, size_t count, uint64_t* d_out_words,
                                void* stream) LSR_NOEXCEPT;

/* Digit planes: binds whole 64-bit message words.  Every message row is committed as `planes` (1..4) commitments,
 * unit (i, l) = i * planes + l holding the base-p digits (messages[i][x] / p^l) mod p under d_seeds[i * planes + l];
 * d_out_words: [count * planes][1 + k*n].  lsr_lwe_message_planes(ctx, modulus) = smallest L with p^L >= modulus
 * (0 if more than 4).  DEVICE memory, asynchronous on `stream`.                                                   */
int lsr_lwe_commit_digits_batch_device(LweContext* ctx, const uint64_t* d_messages, size_t msg_len,
                                       const uint64_t* d_seeds, size_t count, uint32_t planes,
                                       uint64_t* d_out_words, void* stream) LSR_NOEXCEPT;
uint32_t lsr_lwe_message_planes(const LweContext* ctx, uint64_t modulus) LSR_NOEXCEPT;

/* Strict message range for the HOST-pointer entry points (lwe_commit, lwe_commit_batch, lwe_verify_opening*):
 * a message word >= p makes the commit calls fail (NULL / -1) and a verification answer 0 -- SEAL's debug-build
 * behaviour (BatchEncoder::encode throws, commitment.cpp:152,158).  Default 0 (words are reduced mod p).        */
int lsr_lwe_set_strict_messages(LweContext* ctx, int strict) LSR_NOEXCEPT;

/* Largest sum of |centred coefficients| lwe_linear_combine accepts (see there). */
uint64_t lsr_lwe_lincomb_budget(const LweContext* ctx) LSR_NOEXCEPT;

/* Explicit mode: t = A*s + e + Delta*m for CALLER-SUPPLIED s, e instead of
 * sampled ones -- the parity entry (any s, e can be put through the kernels and compared with the
 * oracle) and the way to commit under externally generated randomness.  s, e: [count][k][n]
 * two's-complement words, any int64 (reduced mod q); messages / out_words as lwe_commit_batch.
 * Runs the un-fused kernels (transforms, mat-vec, finalize); lsr_lwe_sample_se(seed) fed back
 * through it reproduces lwe_commit(seed) bit for bit.  HOST memory.          */
int lsr_lwe_commit_explicit(LweContext* ctx, const uint64_t* messages, size_t msg_len,
                            const int64_t* s, const int64_t* e, size_t count,
                            uint64_t* out_words) LSR_NOEXCEPT;

/* same, DEVICE memory; asynchronous on `stream` */
int lsr_lwe_commit_explicit_device(LweContext* ctx, const uint64_t* d_messages, size_t msg_len,
                                   const int64_t* d_s, const int64_t* d_e, size_t count,
                                   uint64_t* d_out_words, void* stream) LSR_NOEXCEPT;

/* results[i] in {1, 0, -1} as lwe_verify_opening; commitments: [count][1+k*n]
 * container words, messages: [count][msg_len].  HOST memory.                */
int lwe_verify_opening_batch(const LweContext* ctx, const uint64_t* comm_words,
                             const uint64_t* messages, size_t msg_len, size_t count,
                             int* results) LSR_NOEXCEPT;

/* same for DEVICE memory, asynchronous on `stream`: d_diff[i] receives the OR of (decoded slot XOR message word mod p) over
 * the first msg_len slots and d_invalid[i] a non-zero value for a malformed container, i.e. the opening verifies iff both are
 * zero (1 / 0 / -1 of lwe_verify_opening = !invalid && !diff / !invalid && diff / invalid).  Both arrays are zeroed by the
 * call.  msg_len <= ring_degree.  Served by the one-kernel verification (ring_degree 4096, module_rank 2..4); -1 for any
 * other shape (use the host-pointer entry point).                                                                      */
int lsr_lwe_verify_opening_batch_device(const LweContext* ctx, const uint64_t* d_comm_words, const uint64_t* d_messages,
                                        size_t msg_len, size_t count, uint64_t* d_diff, int* d_invalid,
                                        void* stream) LSR_NOEXCEPT;

/* Deterministic sampler: samples a pure function of (seed32, sigma, index);
 * the oracle implements the same stream (lsro_sample_gaussian_seeded).      */
int lsr_sample_gaussian_seeded(uint64_t* output, size_t len, double sigma,
                               const uint8_t seed32[32]) LSR_NOEXCEPT;

/* Test hook: CDT magnitude #{k : cdf[k] < u[i]} evaluated on the device for caller-chosen
 * u (boundary cases the keystream never reaches).  variant 0 = linear scan of the table in
 * global memory (generic path), 1 = unrolled scan of the by-value table, 2 = warp-shuffle
 * binary search over the first 31 entries + tail scan, 3 = the compact carry-chain search
 * over the distinct table values when there are at most 31 of them, else variant 2,
 * 4 = the fused kernel's decision: 25-bit prefix search, and variant 3 for the whole warp
 * when some lane's prefix ties with a table prefix.  HOST memory.                      */
int lsr_cdt_magnitude_device(double sigma, const uint64_t* u, size_t count, uint32_t* out,
                             int variant) LSR_NOEXCEPT;
/* Timing hook, the device analogue of the reference's dudect harness (cpp-core/tools/dudect_sampler.cpp:105-141): the
 * same searches, and for every warp (32 consecutive u) the clock64() ticks it spent inside the search ->
 * cycles_per_warp[ceil(count / 32)].  A caller feeds two input classes (fixed / random) and compares the two timing
 * distributions with Welch's t (tests/test_gpu_timing.py, threshold |t| < 4.5 as in the reference).  HOST memory. */
int lsr_cdt_timing_device(double sigma, const uint64_t* u, size_t count, int variant, uint32_t* out,
                          uint64_t* cycles_per_warp) LSR_NOEXCEPT;

/* Test hook for the Goldilocks (q = 2^64 - 2^32 + 1) primitives of the quotient pipeline's transforms, on any
 * 64-bit operands: out[4i..4i+3] = { a*b mod q, (a + (b mod q)) mod q through the lazy sum of the forward
 * butterflies, (a - (b mod q)) mod q, ((a mod q) + (b mod q)) mod q through the canonical sum }.  HOST memory. */
int lsr_goldilocks_probe_device(const uint64_t* a, const uint64_t* b, size_t count, uint64_t* out) LSR_NOEXCEPT;

/* s, e of the commitment (context, seed): two's-complement, each [k][n];
 * test hook for the fused sampler.  HOST memory.                            */
int lsr_lwe_sample_se(LweContext* ctx, uint64_t seed, int64_t* s, int64_t* e) LSR_NOEXCEPT;

#ifdef __cplusplus
}  /* extern "C" */
#endif
#endif /* LAMBDA_SNARK_B200_H */
