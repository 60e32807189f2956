// lsr_abi.cpp -- the extern "C" surface declared in include/lambda_snark_b200.h.
//
// Mirrors the error and ownership conventions of the reference's C ABI
// (cpp-core/src/ntt.cpp, commitment.cpp, utils.cpp): no exception crosses the
// boundary, failures are NULL / -1, every returned object is freed by its
// matching *_free, NULL frees are no-ops, commitments are zeroised on free.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>

#include "lambda_snark_b200.h"
#include "lsr_engine.h"
#include "lsr_r1cs.h"

namespace lsr {
u64 reference_root_of_unity(u64 q, uint32_t n);
int r1cs_quotient_batch(R1csHandle* h, const u64* witnesses, size_t count, u64 omega, u64* out, int* status);
bool fs_challenge_launch(const u64* d_pub, size_t n_pub, const u64* d_containers, size_t words, size_t count,
                         u64 modulus, bool chain, u64* d_ab, u64* d_hashes, cudaStream_t s);
bool fs_challenge_host(const u64* pub, size_t n_pub, const u64* containers, size_t words, size_t count, u64 modulus,
                       bool chain, u64* ab, u64* hashes);
bool verify_r1cs_host(u64 m, u64 modulus, const u64* pub, size_t n_pub, const u64* containers, size_t words,
                      const u64* challenges, const u64* evals, size_t count, int* results);
bool poly_eval_host(u64 modulus, const u64* coeffs, size_t len, size_t polys, const u64* points, size_t npts, u64* out);
int prove_r1cs_batch(R1csHandle* h, const LweContext* lwe, const u64* witnesses, size_t count, size_t n_public,
                     u64 omega, const u64* seeds, u64* containers, u64* challenges, u64* hashes, u64* evals, int* status);
bool gold_probe_host(const u64* a, const u64* b, size_t count, u64* out);
int prover_commit_quotient(R1csHandle* h, const LweContext* lwe, const u64* witnesses, size_t count, u64 omega,
                           const u64* seeds, size_t chunk_lo, size_t chunk_hi, u64* out, bool io_on_device, int* status);
}

using lsr::u64;

#define LSR_TRY try {
#define LSR_CATCH(ret)                                                        \
    } catch (const std::exception& e) {                                       \
        std::fprintf(stderr, "%s error: %s\n", __func__, e.what());           \
        lsr::set_error(e.what());                                             \
        return ret;                                                           \
    } catch (...) {                                                           \
        std::fprintf(stderr, "%s error: unknown exception\n", __func__);      \
        return ret;                                                           \
    }

static LweCommitment* new_commitment(size_t words) {
    LweCommitment* c = new (std::nothrow) LweCommitment;
    if (!c) return nullptr;
    c->len = words;
    c->data = new (std::nothrow) uint64_t[words];
    if (!c->data) { delete c; return nullptr; }
    return c;
}

// commitment.cpp:62-86 container check (+ our fixed payload size and range)
static bool container_ok(const LweContext* ctx, const LweCommitment* c) {
    if (!c || !c->data || c->len < 1) return false;
    const uint64_t byte_len = c->data[0];
    if (byte_len == 0 || byte_len > (c->len - 1) * sizeof(uint64_t)) return false;
    const size_t kn = (size_t)ctx->k * ctx->n;
    if (byte_len != kn * sizeof(uint64_t)) return false;
    for (size_t i = 0; i < kn; i++) if (c->data[1 + i] >= ctx->q) return false;
    return true;
}

extern "C" {

/* ------------------------------------------------------------------ misc */
const char* lsr_version(void) LSR_NOEXCEPT { return "lambda-snark-r_b200 0.1 (sm_100a)"; }
const char* lsr_last_error(void) LSR_NOEXCEPT { return lsr::last_error(); }

int lsr_device_count(void) LSR_NOEXCEPT {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int lsr_set_device(int device) LSR_NOEXCEPT {
    int n = lsr_device_count();
    if (device < 0 || device >= n) { lsr::set_error("lsr_set_device: no such device"); return -1; }
    lsr::set_device_choice(device);
    return lsr::cuda_ok(cudaSetDevice(device), "cudaSetDevice") ? 0 : -1;
}

/* ------------------------------------------------------------------- NTT */
NttContext* ntt_context_create(uint64_t q, uint32_t n) LSR_NOEXCEPT {
    LSR_TRY
    if (n == 0) return nullptr;                       // ntt.cpp:31
    return lsr::ntt_create(q, n);
    LSR_CATCH(nullptr)
}

void ntt_context_free(NttContext* ctx) LSR_NOEXCEPT {
    try { lsr::ntt_destroy(ctx); } catch (...) {}
}

int ntt_forward(const NttContext* ctx, uint64_t* coeffs, uint32_t n) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !coeffs || n != ctx->degree) return -1;   // ntt.cpp:81
    return lsr::ntt_transform_host(ctx, reinterpret_cast<u64*>(coeffs), 1, false) ? 0 : -1;
    LSR_CATCH(-1)
}

int ntt_inverse(const NttContext* ctx, uint64_t* evals, uint32_t n) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !evals || n != ctx->degree) return -1;    // ntt.cpp:96
    return lsr::ntt_transform_host(ctx, reinterpret_cast<u64*>(evals), 1, true) ? 0 : -1;
    LSR_CATCH(-1)
}

void ntt_mul_pointwise(const NttContext* ctx, uint64_t* result, const uint64_t* a, const uint64_t* b,
                       uint32_t n) LSR_NOEXCEPT {
    try {
        if (!ctx || !result || !a || !b) return;          // ntt.cpp:113 (silent)
        lsr::pointwise_host(ctx, reinterpret_cast<u64*>(result), reinterpret_cast<const u64*>(a),
                            reinterpret_cast<const u64*>(b), n);
    } catch (...) {}
}

int ntt_forward_batch(const NttContext* ctx, uint64_t* coeffs, size_t batch) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !coeffs) return -1;
    return lsr::ntt_transform_host(ctx, reinterpret_cast<u64*>(coeffs), batch, false) ? 0 : -1;
    LSR_CATCH(-1)
}

int ntt_inverse_batch(const NttContext* ctx, uint64_t* evals, size_t batch) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !evals) return -1;
    return lsr::ntt_transform_host(ctx, reinterpret_cast<u64*>(evals), batch, true) ? 0 : -1;
    LSR_CATCH(-1)
}

/* ------------------------------------------- cyclic transforms + quotient (N1) */
uint64_t lsr_reference_root_of_unity(uint64_t q, uint32_t n) LSR_NOEXCEPT {
    LSR_TRY
    return lsr::reference_root_of_unity(q, n);
    LSR_CATCH(0)
}

NttContext* lsr_cyclic_ntt_context_create(uint64_t q, uint32_t n, uint64_t omega) LSR_NOEXCEPT {
    LSR_TRY
    if (omega == 0) omega = lsr::reference_root_of_unity(q, n);
    return lsr::ntt_create_cyclic(q, n, omega);
    LSR_CATCH(nullptr)
}

int lsr_cyclic_ntt_forward(const NttContext* ctx, uint64_t* coeffs, size_t batch) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !coeffs || !ctx->cyclic) return -1;
    return lsr::ntt_transform_host(ctx, reinterpret_cast<u64*>(coeffs), batch, false, true) ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_cyclic_ntt_inverse(const NttContext* ctx, uint64_t* evals, size_t batch) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !evals || !ctx->cyclic) return -1;
    return lsr::ntt_transform_host(ctx, reinterpret_cast<u64*>(evals), batch, true, true) ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_r1cs_quotient_batch(void* r1cs, const uint64_t* witnesses, size_t witness_len, size_t count, uint64_t omega,
                            uint64_t* out, int* status) LSR_NOEXCEPT {
    LSR_TRY
    if (!r1cs || !witnesses || !out || !status) return LAMBDA_SNARK_ERR_NULL_PTR;
    lsr::R1csHandle* h = static_cast<lsr::R1csHandle*>(r1cs);
    if (witness_len != h->cols) return LAMBDA_SNARK_ERR_INVALID_PARAMS;
    return lsr::r1cs_quotient_batch(h, reinterpret_cast<const u64*>(witnesses), count, omega, reinterpret_cast<u64*>(out), status);
    LSR_CATCH(LAMBDA_SNARK_ERR_CRYPTO_FAILED)
}

LambdaSnarkError lsr_r1cs_quotient(void* r1cs, const uint64_t* witness, size_t witness_len, uint64_t omega,
                                   uint64_t* out, size_t out_cap, size_t* out_len) LSR_NOEXCEPT {
    LSR_TRY
    if (!r1cs || !witness || !out || !out_len) return LAMBDA_SNARK_ERR_NULL_PTR;
    lsr::R1csHandle* h = static_cast<lsr::R1csHandle*>(r1cs);
    if (witness_len != h->cols || out_cap < h->rows) return LAMBDA_SNARK_ERR_INVALID_PARAMS;
    int status = 0;
    const int rc = lsr::r1cs_quotient_batch(h, reinterpret_cast<const u64*>(witness), 1, omega, reinterpret_cast<u64*>(out), &status);
    if (rc != 0) return static_cast<LambdaSnarkError>(rc);
    if (status != 0) return LAMBDA_SNARK_ERR_CRYPTO_FAILED;          // r1cs.rs:477-481, :1055-1059: witness invalid
    size_t len = h->rows;                                            // r1cs.rs:1062-1064: trailing zeros removed
    while (len > 1 && out[len - 1] == 0) --len;
    *out_len = len;
    return LAMBDA_SNARK_OK;
    LSR_CATCH(LAMBDA_SNARK_ERR_CRYPTO_FAILED)
}

uint32_t lsr_prover_quotient_planes(void* r1cs, const LweContext* ctx) LSR_NOEXCEPT {
    if (!r1cs || !ctx) return 0;
    return lsr::message_planes(ctx->p, static_cast<const lsr::R1csHandle*>(r1cs)->q);
}

size_t lsr_prover_quotient_chunks(void* r1cs, const LweContext* ctx) LSR_NOEXCEPT {
    if (!r1cs || !ctx) return 0;
    const lsr::R1csHandle* h = static_cast<const lsr::R1csHandle*>(r1cs);
    return (h->rows <= ctx->n ? 1 : (size_t)h->rows / ctx->n) * lsr::message_planes(ctx->p, h->q);
}

static int prover_commit_impl(void* r1cs, LweContext* ctx, const uint64_t* witnesses, size_t witness_len, size_t count,
                              uint64_t omega, const uint64_t* seeds, size_t chunk_lo, size_t chunk_hi, uint64_t* out,
                              int* status, bool on_device) {
    if (!r1cs || !ctx || !witnesses || !seeds || !out || !status) return LAMBDA_SNARK_ERR_NULL_PTR;
    lsr::R1csHandle* h = static_cast<lsr::R1csHandle*>(r1cs);
    if (witness_len != h->cols) return LAMBDA_SNARK_ERR_INVALID_PARAMS;
    return lsr::prover_commit_quotient(h, ctx, reinterpret_cast<const u64*>(witnesses), count, omega,
                                       reinterpret_cast<const u64*>(seeds), chunk_lo, chunk_hi,
                                       reinterpret_cast<u64*>(out), on_device, status);
}

int lsr_prover_commit_quotient(void* r1cs, LweContext* ctx, const uint64_t* witnesses, size_t witness_len, size_t count,
                               uint64_t omega, const uint64_t* seeds, size_t chunk_lo, size_t chunk_hi, uint64_t* out,
                               int* status) LSR_NOEXCEPT {
    LSR_TRY
    return prover_commit_impl(r1cs, ctx, witnesses, witness_len, count, omega, seeds, chunk_lo, chunk_hi, out, status, false);
    LSR_CATCH(LAMBDA_SNARK_ERR_CRYPTO_FAILED)
}

int lsr_prover_commit_quotient_device(void* r1cs, LweContext* ctx, const uint64_t* d_witnesses, size_t witness_len,
                                      size_t count, uint64_t omega, const uint64_t* d_seeds, size_t chunk_lo,
                                      size_t chunk_hi, uint64_t* d_out, int* status) LSR_NOEXCEPT {
    LSR_TRY
    return prover_commit_impl(r1cs, ctx, d_witnesses, witness_len, count, omega, d_seeds, chunk_lo, chunk_hi, d_out, status, true);
    LSR_CATCH(LAMBDA_SNARK_ERR_CRYPTO_FAILED)
}

/* ------------------------------------------- Fiat-Shamir + evaluations (N2) */
int lsr_fs_challenge_batch(const uint64_t* public_inputs, size_t n_public, const uint64_t* containers, size_t words,
                           size_t count, uint64_t modulus, int chain, uint64_t* challenges, uint64_t* hashes) LSR_NOEXCEPT {
    LSR_TRY
    if ((!public_inputs && n_public && count) || (!containers && words && count) || !challenges || !hashes || modulus == 0) return -1;
    return lsr::fs_challenge_host(reinterpret_cast<const u64*>(public_inputs), n_public, reinterpret_cast<const u64*>(containers),
                                  words, count, modulus, chain != 0, reinterpret_cast<u64*>(challenges),
                                  reinterpret_cast<u64*>(hashes)) ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_fs_challenge_batch_device(const uint64_t* d_public_inputs, size_t n_public, const uint64_t* d_containers,
                                  size_t words, size_t count, uint64_t modulus, int chain, uint64_t* d_challenges,
                                  uint64_t* d_hashes, void* stream) LSR_NOEXCEPT {
    LSR_TRY
    if (!d_challenges || !d_hashes || modulus == 0) return -1;
    return lsr::fs_challenge_launch(reinterpret_cast<const u64*>(d_public_inputs), n_public,
                                    reinterpret_cast<const u64*>(d_containers), words, count, modulus, chain != 0,
                                    reinterpret_cast<u64*>(d_challenges), reinterpret_cast<u64*>(d_hashes),
                                    static_cast<cudaStream_t>(stream)) ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_poly_eval_batch(uint64_t modulus, const uint64_t* coeffs, size_t len, size_t polys, const uint64_t* points,
                        size_t npts, uint64_t* out) LSR_NOEXCEPT {
    LSR_TRY
    if (!coeffs || !points || !out) return -1;
    return lsr::poly_eval_host(modulus, reinterpret_cast<const u64*>(coeffs), len, polys, reinterpret_cast<const u64*>(points),
                               npts, reinterpret_cast<u64*>(out)) ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_prove_r1cs_batch(void* r1cs, LweContext* ctx, const uint64_t* witnesses, size_t witness_len, size_t count,
                         size_t n_public, uint64_t omega, const uint64_t* seeds, uint64_t* containers,
                         uint64_t* challenges, uint64_t* hashes, uint64_t* evals, int* status) LSR_NOEXCEPT {
    LSR_TRY
    if (!r1cs || !ctx || !witnesses || !seeds || !containers || !challenges || !hashes || !evals || !status)
        return LAMBDA_SNARK_ERR_NULL_PTR;
    lsr::R1csHandle* h = static_cast<lsr::R1csHandle*>(r1cs);
    if (witness_len != h->cols) return LAMBDA_SNARK_ERR_INVALID_PARAMS;
    return lsr::prove_r1cs_batch(h, ctx, reinterpret_cast<const u64*>(witnesses), count, n_public, omega,
                                 reinterpret_cast<const u64*>(seeds), reinterpret_cast<u64*>(containers),
                                 reinterpret_cast<u64*>(challenges), reinterpret_cast<u64*>(hashes),
                                 reinterpret_cast<u64*>(evals), status);
    LSR_CATCH(LAMBDA_SNARK_ERR_CRYPTO_FAILED)
}

int lsr_verify_r1cs_batch(uint64_t n_constraints, uint64_t modulus, const uint64_t* public_inputs, size_t n_public,
                          const uint64_t* containers, size_t words, const uint64_t* challenges, const uint64_t* evals,
                          size_t count, int* results) LSR_NOEXCEPT {
    LSR_TRY
    if ((!public_inputs && n_public && count) || !containers || !challenges || !evals || !results) return -1;
    return lsr::verify_r1cs_host(n_constraints, modulus, reinterpret_cast<const u64*>(public_inputs), n_public,
                                 reinterpret_cast<const u64*>(containers), words, reinterpret_cast<const u64*>(challenges),
                                 reinterpret_cast<const u64*>(evals), count, results) ? 0 : -1;
    LSR_CATCH(-1)
}

void* lsr_host_alloc(size_t bytes) LSR_NOEXCEPT {
    LSR_TRY
    void* p = nullptr;
    if (bytes == 0 || !lsr::cuda_ok(cudaHostAlloc(&p, bytes, cudaHostAllocPortable), "cudaHostAlloc")) return nullptr;
    return p;
    LSR_CATCH(nullptr)
}

void lsr_host_free(void* p) LSR_NOEXCEPT {
    if (p) cudaFreeHost(p);
}

/* ------------------------------------------------ peer-memory gather (C1) */
void* lsr_device_alloc(size_t bytes) LSR_NOEXCEPT {
    LSR_TRY
    void* p = nullptr;
    if (bytes == 0 || !lsr::cuda_ok(cudaSetDevice(lsr::current_device_choice()), "cudaSetDevice") ||
        !lsr::cuda_ok(cudaMalloc(&p, bytes), "cudaMalloc")) return nullptr;
    return p;
    LSR_CATCH(nullptr)
}

void lsr_device_free(void* d_ptr) LSR_NOEXCEPT {
    if (d_ptr) cudaFree(d_ptr);
}

static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size is part of the ABI");

int lsr_peer_export(void* d_ptr, uint8_t handle[64]) LSR_NOEXCEPT {
    LSR_TRY
    if (!d_ptr || !handle) return -1;
    cudaIpcMemHandle_t h;
    if (!lsr::cuda_ok(cudaIpcGetMemHandle(&h, d_ptr), "cudaIpcGetMemHandle")) return -1;
    std::memcpy(handle, &h, sizeof h);
    return 0;
    LSR_CATCH(-1)
}

void* lsr_peer_open(const uint8_t handle[64]) LSR_NOEXCEPT {
    LSR_TRY
    if (!handle) return nullptr;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, sizeof h);
    void* p = nullptr;
    if (!lsr::cuda_ok(cudaSetDevice(lsr::current_device_choice()), "cudaSetDevice") ||
        !lsr::cuda_ok(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle")) return nullptr;
    return p;
    LSR_CATCH(nullptr)
}

int lsr_peer_close(void* d_mapped) LSR_NOEXCEPT {
    if (!d_mapped) return 0;
    return lsr::cuda_ok(cudaIpcCloseMemHandle(d_mapped), "cudaIpcCloseMemHandle") ? 0 : -1;
}

int ntt_mul_pointwise_batch(const NttContext* ctx, uint64_t* result, const uint64_t* a, const uint64_t* b,
                            size_t total) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !result || !a || !b) return -1;
    return lsr::pointwise_host(ctx, reinterpret_cast<u64*>(result), reinterpret_cast<const u64*>(a),
                               reinterpret_cast<const u64*>(b), total) ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_ntt_forward_device(const NttContext* ctx, uint64_t* d, size_t batch, void* stream) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !d) return -1;
    return lsr::ntt_forward_launch(ctx, reinterpret_cast<u64*>(d), batch, static_cast<cudaStream_t>(stream)) ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_ntt_inverse_device(const NttContext* ctx, uint64_t* d, size_t batch, void* stream) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !d) return -1;
    return lsr::ntt_inverse_launch(ctx, reinterpret_cast<u64*>(d), batch, static_cast<cudaStream_t>(stream)) ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_ntt_mul_pointwise_device(const NttContext* ctx, uint64_t* r, const uint64_t* a, const uint64_t* b,
                                 size_t total, void* stream) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !r || !a || !b) return -1;
    return lsr::pointwise_launch(ctx, reinterpret_cast<u64*>(r), reinterpret_cast<const u64*>(a),
                                 reinterpret_cast<const u64*>(b), total, static_cast<cudaStream_t>(stream)) ? 0 : -1;
    LSR_CATCH(-1)
}

uint64_t lsr_ntt_modulus(const NttContext* ctx) LSR_NOEXCEPT { return ctx ? ctx->modulus : 0; }
uint32_t lsr_ntt_degree(const NttContext* ctx) LSR_NOEXCEPT { return ctx ? ctx->degree : 0; }
uint64_t lsr_ntt_root(const NttContext* ctx) LSR_NOEXCEPT { return ctx ? ctx->psi : 0; }
int      lsr_ntt_device(const NttContext* ctx) LSR_NOEXCEPT { return ctx ? ctx->device : -1; }

/* ------------------------------------------------------------ commitment */
LweContext* lwe_context_create_seeded(const PublicParams* params, const uint8_t seed32[32]) LSR_NOEXCEPT {
    LSR_TRY
    if (!params || !seed32) return nullptr;               // commitment.cpp:103
    return lsr::lwe_create(params->modulus, params->ring_degree, params->module_rank, params->sigma, seed32);
    LSR_CATCH(nullptr)
}

LweContext* lwe_context_create(const PublicParams* params) LSR_NOEXCEPT {
    LSR_TRY
    if (!params) return nullptr;
    uint8_t seed[32];
    if (!lsr::host::os_entropy(seed, sizeof(seed))) { lsr::set_error("no entropy source"); return nullptr; }
    LweContext* c = lsr::lwe_create(params->modulus, params->ring_degree, params->module_rank, params->sigma, seed);
    volatile uint8_t* sp = seed;
    for (size_t i = 0; i < sizeof(seed); i++) sp[i] = 0;
    return c;
    LSR_CATCH(nullptr)
}

void lwe_context_free(LweContext* ctx) LSR_NOEXCEPT {
    try { lsr::lwe_destroy(ctx); } catch (...) {}
}

uint64_t lsr_lwe_modulus(const LweContext* ctx) LSR_NOEXCEPT { return ctx ? ctx->q : 0; }
uint64_t lsr_lwe_plain_modulus(const LweContext* ctx) LSR_NOEXCEPT { return ctx ? ctx->p : 0; }
uint64_t lsr_lwe_delta(const LweContext* ctx) LSR_NOEXCEPT { return ctx ? ctx->delta : 0; }
size_t   lsr_lwe_commitment_words(const LweContext* ctx) LSR_NOEXCEPT { return ctx ? lsr::lwe_words(ctx) : 0; }

int lsr_lwe_copy_matrix(const LweContext* ctx, uint64_t* out) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !out) return -1;
    if (!lsr::cuda_ok(cudaSetDevice(ctx->device), "cudaSetDevice")) return -1;
    const size_t bytes = (size_t)ctx->k * ctx->k * ctx->n * sizeof(uint64_t);
    return lsr::cuda_ok(cudaMemcpy(out, ctx->d_A, bytes, cudaMemcpyDeviceToHost), "copy matrix") ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_lwe_set_strict_messages(LweContext* ctx, int strict) LSR_NOEXCEPT {
    if (!ctx) return -1;
    ctx->strict_messages = strict ? 1 : 0;
    return 0;
}

uint64_t lsr_lwe_lincomb_budget(const LweContext* ctx) LSR_NOEXCEPT { return ctx ? lsr::lincomb_budget(ctx) : 0; }

uint32_t lsr_lwe_message_planes(const LweContext* ctx, uint64_t modulus) LSR_NOEXCEPT {
    return ctx ? lsr::message_planes(ctx->p, modulus) : 0;
}

// strict mode: the first min(msg_len, n) words of every message must be below p
static bool messages_in_range(const LweContext* ctx, const uint64_t* messages, size_t msg_len, size_t count) {
    const size_t used = msg_len < ctx->n ? msg_len : ctx->n;
    for (size_t i = 0; i < count; i++)
        for (size_t x = 0; x < used; x++)
            if (messages[i * msg_len + x] >= ctx->p) return false;
    return true;
}

int lsr_lwe_set_commit_path(LweContext* ctx, int path) LSR_NOEXCEPT {
    if (!ctx || path < 0 || path > 2) return -1;
    ctx->commit_path = path;
    return 0;
}

int lsr_ntt_set_arith(NttContext* ctx, int arith) LSR_NOEXCEPT {
    if (!ctx || arith < 0 || arith > 2) return -1;
    if (arith == 2 && !ctx->mp.f64_ok) return -1;         // FP64 butterflies are exact only for q < 2^45
    ctx->arith = arith == 2 ? 0 : arith;                  // 0 (auto) already picks FP64 whenever it is exact
    return 0;
}

int lsr_ntt_arith(const NttContext* ctx) LSR_NOEXCEPT {
    if (!ctx) return -1;
    return (ctx->arith != 1 && ctx->mp.f64_ok) ? 2 : 1;
}

int lsr_lwe_set_arith(LweContext* ctx, int arith) LSR_NOEXCEPT {
    if (!ctx) return -1;
    return lsr_ntt_set_arith(ctx->ntt, arith);
}

LweCommitment* lwe_commit(LweContext* ctx, const uint64_t* message, size_t msg_len, uint64_t seed) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !message) return nullptr;                 // commitment.cpp:144
    if (ctx->strict_messages && !messages_in_range(ctx, message, msg_len, 1)) {
        // BatchEncoder::encode rejects values >= plain_modulus (SEAL debug builds) -> exception -> NULL (commitment.cpp:158)
        lsr::set_error("lwe_commit: message word >= plaintext modulus (strict mode)");
        return nullptr;
    }
    if (seed == 0) {                                      // commitment.h:52 "0 = random"
        uint8_t buf[8];
        do {
            if (!lsr::host::os_entropy(buf, sizeof(buf))) return nullptr;
            std::memcpy(&seed, buf, sizeof(seed));
        } while (seed == 0);
    }
    const size_t words = lsr::lwe_words(ctx);
    LweCommitment* c = new_commitment(words);
    if (!c) return nullptr;
    const size_t used = msg_len < ctx->n ? msg_len : ctx->n;   // commitment.cpp:146-149
    if (!lsr::lwe_commit_host(ctx, reinterpret_cast<const u64*>(message), used, reinterpret_cast<const u64*>(&seed), 1,
                              reinterpret_cast<u64*>(c->data))) {
        lwe_commitment_free(c);
        return nullptr;
    }
    return c;
    LSR_CATCH(nullptr)
}

int lwe_commit_batch(LweContext* ctx, const uint64_t* messages, size_t msg_len, const uint64_t* seeds,
                     size_t count, uint64_t* out_words) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !seeds || !out_words || (!messages && msg_len)) return -1;
    if (ctx->strict_messages && msg_len && !messages_in_range(ctx, messages, msg_len, count)) {
        lsr::set_error("lwe_commit_batch: message word >= plaintext modulus (strict mode)");
        return -1;
    }
    return lsr::lwe_commit_host(ctx, reinterpret_cast<const u64*>(messages), msg_len,
                                reinterpret_cast<const u64*>(seeds), count, reinterpret_cast<u64*>(out_words)) ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_lwe_commit_batch_device(LweContext* ctx, const uint64_t* d_messages, size_t msg_len,
                                const uint64_t* d_seeds, size_t count, uint64_t* d_out_words,
                                void* stream) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !d_seeds || !d_out_words || (!d_messages && msg_len)) return -1;
    return lsr::lwe_commit_launch(ctx, reinterpret_cast<const u64*>(d_messages), msg_len,
                                  reinterpret_cast<const u64*>(d_seeds), count,
                                  reinterpret_cast<u64*>(d_out_words), static_cast<cudaStream_t>(stream)) ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_lwe_commit_digits_batch_device(LweContext* ctx, const uint64_t* d_messages, size_t msg_len,
                                       const uint64_t* d_seeds, size_t count, uint32_t planes, uint64_t* d_out_words,
                                       void* stream) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !d_seeds || !d_out_words || (!d_messages && msg_len)) return -1;
    if (planes < 1 || planes > 4 || count > (SIZE_MAX / planes)) { lsr::set_error("digit planes must be 1..4"); return -1; }
    return lsr::lwe_commit_launch(ctx, reinterpret_cast<const u64*>(d_messages), msg_len,
                                  reinterpret_cast<const u64*>(d_seeds), count * planes,
                                  reinterpret_cast<u64*>(d_out_words), static_cast<cudaStream_t>(stream), planes, 0) ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_lwe_commit_explicit(LweContext* ctx, const uint64_t* messages, size_t msg_len, const int64_t* s,
                            const int64_t* e, size_t count, uint64_t* out_words) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !s || !e || !out_words || (!messages && msg_len)) return -1;
    return lsr::lwe_commit_explicit_host(ctx, reinterpret_cast<const u64*>(messages), msg_len, s, e, count,
                                         reinterpret_cast<u64*>(out_words)) ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_lwe_commit_explicit_device(LweContext* ctx, const uint64_t* d_messages, size_t msg_len, const int64_t* d_s,
                                   const int64_t* d_e, size_t count, uint64_t* d_out_words, void* stream) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !d_s || !d_e || !d_out_words || (!d_messages && msg_len)) return -1;
    return lsr::lwe_commit_explicit_launch(ctx, reinterpret_cast<const u64*>(d_messages), msg_len, d_s, d_e, count,
                                           reinterpret_cast<u64*>(d_out_words), static_cast<cudaStream_t>(stream)) ? 0 : -1;
    LSR_CATCH(-1)
}

void lwe_commitment_free(LweCommitment* comm) LSR_NOEXCEPT {
    if (!comm) return;                                    // commitment.cpp:167
    if (comm->data) {
        volatile uint64_t* p = comm->data;                // zeroise (commitment.cpp:169-173)
        for (size_t i = 0; i < comm->len; i++) p[i] = 0;
        delete[] comm->data;
    }
    delete comm;
}

LweCommitment* lwe_commitment_clone(const LweCommitment* comm) LSR_NOEXCEPT {
    if (!comm || comm->len == 0 || !comm->data) return nullptr;   // commitment.cpp:180-182
    LweCommitment* c = new_commitment(comm->len);
    if (!c) return nullptr;
    std::memcpy(c->data, comm->data, comm->len * sizeof(uint64_t));
    return c;
}

int lwe_verify_opening(const LweContext* ctx, const LweCommitment* commitment, const uint64_t* message,
                       size_t msg_len, const LweOpening* /*opening*/) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !commitment || !message) return -1;       // commitment.cpp:207
    if (!container_ok(ctx, commitment)) return -1;        // :210-212
    if (msg_len > ctx->n) return 0;                       // :219-221
    if (ctx->strict_messages && !messages_in_range(ctx, message, msg_len, 1)) return 0;   // no commitment to such a word exists
    int result = -1;
    if (!lsr::lwe_verify_host(ctx, reinterpret_cast<const u64*>(commitment->data),
                              reinterpret_cast<const u64*>(message), msg_len, 1, &result)) return -1;
    return result;
    LSR_CATCH(-1)
}

int lwe_verify_opening_batch(const LweContext* ctx, const uint64_t* comm_words, const uint64_t* messages,
                             size_t msg_len, size_t count, int* results) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !comm_words || !results || (!messages && msg_len)) return -1;
    if (!lsr::lwe_verify_host(ctx, reinterpret_cast<const u64*>(comm_words),
                              reinterpret_cast<const u64*>(messages), msg_len, count, results)) return -1;
    if (ctx->strict_messages && msg_len)
        for (size_t i = 0; i < count; i++)
            if (results[i] == 1 && !messages_in_range(ctx, messages + i * msg_len, msg_len, 1)) results[i] = 0;
    return 0;
    LSR_CATCH(-1)
}

LweCommitment* lwe_linear_combine(const LweContext* ctx, const LweCommitment** commitments,
                                  const uint64_t* coeffs, size_t count) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !commitments || !coeffs || count == 0) return nullptr;   // commitment.cpp:240-242
    const size_t kn = (size_t)ctx->k * ctx->n;
    std::vector<u64> payloads;
    std::vector<u64> cf;
    for (size_t i = 0; i < count; i++) {
        if (!commitments[i]) continue;                    // :248-250
        if (!container_ok(ctx, commitments[i])) return nullptr;   // :253-255
        payloads.insert(payloads.end(), commitments[i]->data + 1, commitments[i]->data + 1 + kn);
        cf.push_back(coeffs[i]);
    }
    if (cf.empty()) return nullptr;                       // :268-270
    LweCommitment* out = new_commitment(1 + kn);
    if (!out) return nullptr;
    out->data[0] = kn * sizeof(uint64_t);
    if (!lsr::lwe_lincomb_host(ctx, payloads.data(), cf.data(), cf.size(), reinterpret_cast<u64*>(out->data + 1))) {
        lwe_commitment_free(out);
        return nullptr;
    }
    return out;
    LSR_CATCH(nullptr)
}

int lsr_lwe_sample_se(LweContext* ctx, uint64_t seed, int64_t* s, int64_t* e) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !s || !e) return -1;
    return lsr::lwe_sample_se_host(ctx, seed, s, e) ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_goldilocks_probe_device(const uint64_t* a, const uint64_t* b, size_t count, uint64_t* out) LSR_NOEXCEPT {
    LSR_TRY
    if (!a || !b || !out) return -1;
    return lsr::gold_probe_host(reinterpret_cast<const u64*>(a), reinterpret_cast<const u64*>(b), count,
                                reinterpret_cast<u64*>(out)) ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_cdt_magnitude_device(double sigma, const uint64_t* u, size_t count, uint32_t* out, int variant) LSR_NOEXCEPT {
    LSR_TRY
    if (!u || !out || count == 0 || variant < 0 || variant > 4) return -1;
    return lsr::cdt_probe_host(sigma, reinterpret_cast<const u64*>(u), count, out, variant) ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_lwe_verify_opening_batch_device(const LweContext* ctx, const uint64_t* d_comm_words, const uint64_t* d_messages,
                                        size_t msg_len, size_t count, uint64_t* d_diff, int* d_invalid,
                                        void* stream) LSR_NOEXCEPT {
    LSR_TRY
    if (!ctx || !d_comm_words || (!d_messages && msg_len) || !d_diff || !d_invalid || msg_len > ctx->n) return -1;
    if (count == 0) return 0;
    if (!lsr::fused_verify_supported(ctx)) { lsr::set_error("device-pointer verification: unsupported (ring degree, module rank)"); return -1; }
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (!lsr::cuda_ok(cudaSetDevice(ctx->device), "cudaSetDevice") ||
        !lsr::cuda_ok(cudaMemsetAsync(d_diff, 0, count * sizeof(uint64_t), s), "memset") ||
        !lsr::cuda_ok(cudaMemsetAsync(d_invalid, 0, count * sizeof(int), s), "memset")) return -1;
    return lsr::fused_verify_launch(ctx, reinterpret_cast<const u64*>(d_comm_words), lsr::lwe_words(ctx),
                                    reinterpret_cast<const u64*>(d_messages), msg_len, count,
                                    reinterpret_cast<unsigned long long*>(d_diff), d_invalid, s) ? 0 : -1;
    LSR_CATCH(-1)
}

int lsr_cdt_timing_device(double sigma, const uint64_t* u, size_t count, int variant, uint32_t* out,
                          uint64_t* cycles_per_warp) LSR_NOEXCEPT {
    LSR_TRY
    if (!u || !out || !cycles_per_warp || count == 0 || variant < 0 || variant > 4) return -1;
    return lsr::cdt_probe_host(sigma, reinterpret_cast<const u64*>(u), count, out, variant,
                               reinterpret_cast<u64*>(cycles_per_warp)) ? 0 : -1;
    LSR_CATCH(-1)
}

/* --------------------------------------------------------------- sampler */
int lsr_sample_gaussian_seeded(uint64_t* output, size_t len, double sigma, const uint8_t seed32[32]) LSR_NOEXCEPT {
    LSR_TRY
    if (!output || len == 0 || !(sigma > 0.0) || !std::isfinite(sigma) || !seed32) return -1;   // utils.cpp:133-135
    return lsr::sample_gaussian_host(reinterpret_cast<u64*>(output), len, sigma, seed32) ? 0 : -1;
    LSR_CATCH(-1)
}

int sample_gaussian(uint64_t* output, size_t len, double sigma) LSR_NOEXCEPT {
    LSR_TRY
    if (!output || len == 0 || !(sigma > 0.0) || !std::isfinite(sigma)) return -1;               // utils.cpp:133-135
    uint8_t seed[32];
    if (!lsr::host::os_entropy(seed, sizeof(seed))) return -1;
    const bool ok = lsr::sample_gaussian_host(reinterpret_cast<u64*>(output), len, sigma, seed);
    volatile uint8_t* sp = seed;
    for (size_t i = 0; i < sizeof(seed); i++) sp[i] = 0;
    return ok ? 0 : -1;
    LSR_CATCH(-1)
}

}  // extern "C"
