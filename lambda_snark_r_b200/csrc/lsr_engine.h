// lsr_engine.h -- host runtime objects behind the C ABI (NttContext, LweContext)
// and the launchers of the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>

#include <mutex>
#include <string>
#include <vector>

#include "lsr_common.h"
#include "lsr_host.h"

namespace lsr {

// thread-local diagnostic (lsr_last_error)
void set_error(const std::string& msg);
const char* last_error();
bool cuda_ok(cudaError_t e, const char* what);
int current_device_choice();          // device chosen by lsr_set_device on this thread (default 0)
void set_device_choice(int dev);

// RAII device buffer that only grows
struct DeviceScratch {
    void* ptr = nullptr;
    size_t bytes = 0;
    bool reserve(size_t need);
    void release();
};

struct PinnedScratch {
    void* ptr = nullptr;
    size_t bytes = 0;
    bool reserve(size_t need);
    void release();
};

}  // namespace lsr

// The opaque handle of ntt.h:24 (reference layout: cpp-core/src/ntt.cpp:21-26)
struct NttContext {
    lsr::u64 modulus = 0;
    uint32_t degree = 0;
    uint32_t logn = 0;
    lsr::u64 psi = 0;
    int device = 0;
    lsr::ModParams mp{};
    lsr::NttTables tables{};          // device pointers
    ulonglong2* d_fwd = nullptr;
    ulonglong2* d_inv = nullptr;
    ulonglong2* d_fwd_last = nullptr;
    ulonglong2* d_inv_last = nullptr;
    lsr::NttTables tables_f{};        // POL_F64 tables (doubles), valid iff mp.f64_ok
    ulonglong2* d_f_fwd = nullptr;
    ulonglong2* d_f_inv = nullptr;
    ulonglong2* d_f_fwd_last = nullptr;
    ulonglong2* d_f_inv_last = nullptr;
    int arith = 0;                    // 0 auto (FP64 butterflies when exact), 1 integer only
    bool cyclic = false;              // tables of the cyclic transform (X^n - 1); psi then holds omega
    cudaStream_t stream = nullptr;    // used by the host-pointer entry points
    cudaStream_t copy_streams[2] = {nullptr, nullptr};
    cudaEvent_t events[4] = {nullptr, nullptr, nullptr, nullptr};
    mutable std::mutex mu;            // serialises use of the scratch buffers
    mutable lsr::DeviceScratch scratch[3];
    mutable lsr::PinnedScratch pin;   // small single calls: the kernel works on this page-locked buffer in place (zero copy)
};

// The opaque handle of types.h:27 (reference layout: cpp-core/src/commitment.cpp:31-40).
// Holds the Module-LWE public matrix A-hat (NTT domain), the trapdoor z-hat
// that plays the role of the reference's SEAL secret key in
// lwe_verify_opening, the CDT, and the PRF key of the commitment randomness.
struct LweContext {
    lsr::u64 q = 0, p = 0, delta = 0;
    uint32_t n = 0, k = 0, logn = 0;
    double sigma = 0.0;
    int device = 0;
    uint32_t key[8] = {0};
    std::vector<lsr::u64> cdf;        // host copy (full table)
    NttContext* ntt = nullptr;        // owned
    lsr::u64* d_A = nullptr;          // [k][k][n] residues, NTT domain
    ulonglong2* d_A2 = nullptr;       // same with Shoup quotients, fused-kernel layout
    ulonglong2* d_A2f = nullptr;      // same as doubles (a, a/q), fused kernel with FP64 butterflies
    lsr::u64* d_zh = nullptr;         // [k-1][n]
    lsr::u64* d_cdf = nullptr;        // [cdf.size()]
    int commit_path = 0;              // 0 auto, 1 generic, 2 fused
    int strict_messages = 0;          // 1: host-pointer entry points reject message words >= p (lsr_lwe_set_strict_messages)
    mutable std::mutex mu;
    mutable lsr::DeviceScratch scratch[10];
    mutable lsr::PinnedScratch staging[6];   // pageable callers: 3 slots x (messages + seeds | containers)
};

namespace lsr {

LweContext* lwe_create(u64 modulus_req, uint32_t n, uint32_t k, double sigma, const uint8_t seed32[32]);
void lwe_destroy(LweContext* ctx);
size_t lwe_words(const LweContext* ctx);

// device-pointer commit: msgs [count][msg_len], seeds [count], out [count][1+k*n].
// Digit planes: with planes = L > 1 commitment b is unit g = unit0 + b and commits the base-p digit
// g % L of message row g / L (msgs then holds ceil((unit0 + count) / L) rows); planes = 1 commits msgs[b] mod p.
bool lwe_commit_launch(const LweContext* ctx, const u64* d_msgs, size_t msg_len, const u64* d_seeds,
                       size_t count, u64* d_out, cudaStream_t stream, uint32_t planes = 1, size_t unit0 = 0);
// p^l and floor((2^64-1) / p^l) for l = 0..3 (entries whose power overflows 2^63 are 0: such a digit is always 0)
void plane_divisors(u64 p, u64 pdiv[4], u64 pdinv[4]);
// smallest L with p^L >= modulus: digits that bind a whole field element (0 if more than 4 would be needed)
uint32_t message_planes(u64 p, u64 modulus);
// sum of |centred coefficient| a linear combination of fresh commitments may carry and still decode
u64 lincomb_budget(const LweContext* ctx);
bool lwe_commit_host(const LweContext* ctx, const u64* msgs, size_t msg_len, const u64* seeds,
                     size_t count, u64* out);
// results[i] in {1,0,-1}
bool lwe_verify_host(const LweContext* ctx, const u64* comm_words, const u64* msgs, size_t msg_len,
                     size_t count, int* results);
// payloads: [count][k*n] host words (already validated), coeffs [count]; out payload [k*n]
bool lwe_lincomb_host(const LweContext* ctx, const u64* payloads, const u64* coeffs, size_t count,
                      u64* out_payload);
bool lwe_sample_se_host(const LweContext* ctx, u64 seed, int64_t* s, int64_t* e);
bool lwe_commit_explicit_host(const LweContext* ctx, const u64* msgs, size_t msg_len, const int64_t* s, const int64_t* e,
                              size_t count, u64* out);
bool lwe_commit_explicit_launch(const LweContext* ctx, const u64* d_msgs, size_t msg_len, const int64_t* d_s,
                                const int64_t* d_e, size_t count, u64* d_out, cudaStream_t stream);
bool sample_gaussian_host(u64* out, size_t len, double sigma, const uint8_t seed32[32]);
bool fused_commit_supported(const LweContext* ctx);
bool fused_verify_supported(const LweContext* ctx);
bool fused_verify_launch(const LweContext* ctx, const u64* d_comm, size_t stride, const u64* d_msgs, size_t cmp_len,
                         size_t count, unsigned long long* d_diff, int* d_invalid, cudaStream_t s);
bool cdt_probe_host(double sigma, const u64* u, size_t count, uint32_t* out, int variant, u64* cycles_per_warp = nullptr);

NttContext* ntt_create(u64 q, uint32_t n);
// cyclic transform over X^n - 1 with the given primitive n-th root (rust-api/lambda-snark/src/ntt.rs);
// q < 2^61 prime, or Goldilocks 2^64 - 2^32 + 1
NttContext* ntt_create_cyclic(u64 q, uint32_t n, u64 omega);
// negacyclic transform (bit-reversed evaluations at psi^(2 brv(i) + 1)) for a caller-chosen primitive 2n-th
// root psi: the coset evaluations of the quotient pipeline; q < 2^61 prime or Goldilocks, n <= 2^24
NttContext* ntt_create_negacyclic(u64 q, uint32_t n, u64 psi);
// in-place bit-reversal permutation of each polynomial (natural-order views of the transforms)
bool ntt_bitrev_launch(const NttContext* ctx, u64* d_data, size_t batch, cudaStream_t stream);
void ntt_destroy(NttContext* ctx);

// asynchronous launches on `stream`; data on the context's device
bool ntt_forward_launch(const NttContext* ctx, u64* d_data, size_t batch, cudaStream_t stream);
bool ntt_inverse_launch(const NttContext* ctx, u64* d_data, size_t batch, cudaStream_t stream);
// inverse transform with the quotient pipeline's neighbours fused in (InvFusion, lsr_common.h): input data * mul, first
// kernel writing to dst, last kernel storing (fin_c - x) * fin_scale.  Not for n <= 16 (ntt_inverse_fused_supported).
bool ntt_inverse_fused_supported(const NttContext* ctx);
bool ntt_inverse_fused_launch(const NttContext* ctx, u64* d_data, size_t batch, cudaStream_t stream, const InvFusion& fz);
bool pointwise_launch(const NttContext* ctx, u64* d_r, const u64* d_a, const u64* d_b, size_t total,
                      cudaStream_t stream);

// host-pointer paths (H2D, kernel, D2H, synchronised), chunked + double buffered
// natural = true: forward output / inverse input in natural order (bit-reversal permutation added)
bool ntt_transform_host(const NttContext* ctx, u64* host, size_t batch, bool inverse, bool natural = false);
bool pointwise_host(const NttContext* ctx, u64* r, const u64* a, const u64* b, size_t total);

}  // namespace lsr
