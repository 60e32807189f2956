// lsr_commit.cu -- Module-LWE commitment engine behind lwe_* (K4-K7).
//
// Replaces cpp-core/src/commitment.cpp (SEAL BFV symmetric encryption + zstd
// container) with t = A*s + e + Delta*m over R_q = Z_q[X]/(X^n+1), the
// definition north_star asks for.  This file holds the
// context life cycle and the GENERIC multi-kernel path, valid for every (n, k)
// the NTT supports:
//     sample_se -> forward NTT (batch*k) -> matvec -> inverse NTT -> finalize
// lsr_commit_fused.cu holds the single-kernel fused path used for the
// headline configuration.
#include <algorithm>
#include <mutex>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>

#include "lsr_arith.cuh"
#include "lsr_copy_pool.h"
#include "lsr_engine.h"
#include "lsr_sampler.cuh"

namespace lsr {

bool fused_commit_launch(const LweContext* ctx, const u64* d_msgs, size_t msg_len, const u64* d_seeds,
                         size_t count, u64* d_out, cudaStream_t stream, uint32_t planes, size_t unit0);   // lsr_commit_fused.cu
bool fused_prepare(LweContext* ctx, cudaStream_t stream);                  // lsr_commit_fused.cu
bool fused_verify_supported(const LweContext* ctx);                         // lsr_commit_fused.cu
bool fused_verify_launch(const LweContext* ctx, const u64* d_comm, size_t stride, const u64* d_msgs, size_t cmp_len,
                         size_t count, unsigned long long* d_diff, int* d_invalid, cudaStream_t s);

// commitments per pipeline slot of the host-pointer path: 444 = three fused CTAs on each of the 148 SMs, one full
// wave per launch (28 MiB of containers at n = 4096, k = 2)
constexpr size_t kCommitHostChunk = 444;

static ChaChaKey make_key(const LweContext* c) {
    ChaChaKey k;
    for (int i = 0; i < 8; i++) k.k[i] = c->key[i];
    return k;
}

// ---------------------------------------------------------------------------
// generic K7-in-K4: one thread per (commitment, 16-coefficient chunk).
// Randomness layout of DESIGN.md 3.3: word j of block P is the sign draw (bit 0) and the top 31 bits of the
// magnitude draw of coefficient j of polynomial P; the low 33 bits come from the refinement blocks
// 0x100 | (2P + (j>>3)).  This path always forms the full 64-bit draw (the fused kernel fetches the refinement
// only when the top bits tie with a table entry).
// s -> S[b][P][n] as residues, e -> out payload (finalize adds the rest).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
sample_se_kernel(const ChaChaKey key, const u64* __restrict__ cdf, u32 cdf_n, u64 q, u32 n, u32 k,
                 const u64* __restrict__ seeds, size_t count, u64* __restrict__ S,
                 u64* __restrict__ out, size_t out_stride) {
    const size_t chunks = n >> 4;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count * chunks) return;
    const size_t b = idx / chunks;
    const u32 tau = (u32)(idx % chunks);
    const u64 seed = seeds[b];
    const u32 s_lo = (u32)seed, s_hi = (u32)(seed >> 32);
    for (u32 P = 0; P < 2 * k; P++) {
        // chunk tau = the 16 coefficients tau + (n/16) j
        u64* dst = (P < k) ? S + ((b * k + P) * (size_t)n) + tau
                           : out + b * out_stride + 1 + (size_t)(P - k) * n + tau;
        u32 x[16];
        chacha_block(key, s_lo, s_hi, tau, kDomCommit | P, x);
#pragma unroll
        for (u32 h = 0; h < 2; h++) {
            u32 f[16];
            chacha_block(key, s_lo, s_hi, tau, kDomCommit | kDomCommitFine | (2u * P + h), f);
#pragma unroll
            for (u32 w = 0; w < 8; w++) {
                const u32 j = 8 * h + w;
                const u64 u = commit_draw(x[j], f[2 * w], f[2 * w + 1]);
                const u32 mag = cdt_magnitude_global(cdf, cdf_n, u);
                dst[(size_t)j * chunks] = signed_residue(mag, x[j] & 1u, q);
            }
        }
    }
}

// explicit mode: caller-supplied s, e (two's complement, any int64) -> residues;
// s -> S[b][P][n], e -> out payload, exactly where sample_se_kernel leaves its samples
__global__ void __launch_bounds__(256)
load_se_kernel(u64 q, size_t kn, const long long* __restrict__ s, const long long* __restrict__ e, size_t count,
               u64* __restrict__ S, u64* __restrict__ out, size_t out_stride) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count * kn) return;
    const size_t b = idx / kn, r = idx % kn;
    auto residue = [q](long long v) -> u64 {
        const u64 mag = v < 0 ? 0ull - (u64)v : (u64)v;
        const u64 m = mag % q;
        return (v < 0 && m) ? q - m : m;
    };
    S[idx] = residue(s[idx]);
    out[b * out_stride + 1 + r] = residue(e[idx]);
}

// S[b][.][x] <- A[.][.][x] * S[b][.][x]   (NTT-domain mat-vec, in place)
template <int KMAX>
__global__ void __launch_bounds__(256)
matvec_kernel(const ModParams mp, const u64* __restrict__ A, u32 n, u32 k, size_t count,
              u64* __restrict__ S) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count * n) return;
    const size_t b = idx / n;
    const u32 x = (u32)(idx % n);
    u64 sv[KMAX];
    u64* base = S + b * k * (size_t)n + x;
#pragma unroll
    for (u32 j = 0; j < KMAX; j++) if (j < k) sv[j] = base[(size_t)j * n];
    for (u32 i = 0; i < k; i++) {
        u64 acc = 0;
#pragma unroll
        for (u32 j = 0; j < KMAX; j++) {
            if (j < k) acc = addmod(acc, mulmod_exact(A[((size_t)i * k + j) * n + x], sv[j], mp), mp.q);
        }
        base[(size_t)i * n] = acc;
    }
}

// out[b] = [k*n*8, S[b] + e (already in out) + [i==k-1] Delta*(digit of m)], digit = (m / p^plane) mod p for
// unit g = unit0 + b: message row g / planes, plane g % planes (planes = 1: m mod p)
__global__ void __launch_bounds__(256)
finalize_kernel(const ModParams mp, u64 delta, u64 p, u32 n, u32 k, const u64* __restrict__ S,
                const u64* __restrict__ msgs, size_t msg_len, size_t count, u64* __restrict__ out,
                size_t out_stride, u32 planes, size_t unit0) {
    const size_t kn = (size_t)k * n;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count * kn) return;
    const size_t b = idx / kn;
    const size_t r = idx % kn;
    u64* o = out + b * out_stride;
    if (r == 0) o[0] = kn * 8;
    u64 v = addmod(S[idx], o[1 + r], mp.q);
    const size_t x = r - (size_t)(k - 1) * n;          // coefficient index if last row
    if (r >= (size_t)(k - 1) * n && x < msg_len) {     // msg_len already clamped to n
        const size_t unit = unit0 + b;
        u64 word = msgs[(unit / planes) * msg_len + x];
        for (u32 l = (u32)(unit % planes); l > 0; l--) word /= p;
        const u64 m = word % p;
        v = addmod(v, mulmod_exact(delta, m, mp), mp.q);
    }
    o[1 + r] = v;
}

// A[k-1][j] = f_j - sum_i z_i * A[i][j]   (context set-up; F holds f-hat)
__global__ void trapdoor_row_kernel(const ModParams mp, u32 n, u32 k, const u64* __restrict__ zh,
                                    const u64* __restrict__ F, u64* __restrict__ A) {
    const u32 idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= k * n) return;
    const u32 j = idx / n, x = idx % n;
    u64 acc = F[(size_t)j * n + x];
    for (u32 i = 0; i + 1 < k; i++) {
        const u64 prod = mulmod_exact(zh[(size_t)i * n + x], A[((size_t)i * k + j) * n + x], mp);
        acc = submod(acc, prod, mp.q);
    }
    A[((size_t)(k - 1) * k + j) * n + x] = acc;
}

// ---------------------------------------------------------------------------
// K6: verification.  check -> copy rows 0..k-2 -> forward NTT -> U = sum z*t
// -> inverse NTT -> decode + compare.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
verify_prepare_kernel(u64 q, u32 n, u32 k, const u64* __restrict__ comm, size_t stride, size_t count,
                      u64* __restrict__ T, int* __restrict__ invalid) {
    const size_t kn = (size_t)k * n;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count * kn) return;
    const size_t b = idx / kn, r = idx % kn;
    const u64* c = comm + b * stride;
    if (r == 0 && c[0] != kn * 8) invalid[b] = 1;
    const u64 v = c[1 + r];
    if (v >= q) invalid[b] = 1;
    if (r < (size_t)(k - 1) * n) T[b * (size_t)(k - 1) * n + r] = v < q ? v : 0;
}

__global__ void __launch_bounds__(256)
verify_dot_kernel(const ModParams mp, u32 n, u32 k, const u64* __restrict__ zh,
                  const u64* __restrict__ T, size_t count, u64* __restrict__ U) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count * n) return;
    const size_t b = idx / n;
    const u32 x = (u32)(idx % n);
    u64 acc = 0;
    for (u32 i = 0; i + 1 < k; i++)
        acc = addmod(acc, mulmod_exact(zh[(size_t)i * n + x], T[(b * (k - 1) + i) * (size_t)n + x], mp), mp.q);
    U[idx] = acc;
}

// diff[b] |= decode(U + t_last)[x] ^ msg[x]  for x < msg_len  (commitment.cpp:223-226)
__global__ void __launch_bounds__(256)
verify_decode_kernel(u64 q, u64 delta, u64 p, u32 n, u32 k, const u64* __restrict__ U,
                     const u64* __restrict__ comm, size_t stride, const u64* __restrict__ msgs,
                     size_t msg_len, size_t count, unsigned long long* __restrict__ diff) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= count * msg_len) return;
    const size_t b = idx / msg_len, x = idx % msg_len;
    u64 last = comm[b * stride + 1 + (size_t)(k - 1) * n + x];
    if (last >= q) last = 0;                                  // flagged invalid elsewhere
    const u64 v = addmod(U[b * (size_t)n + x], last, q);
    const u64 d = ((v + delta / 2) / delta) % p;
    // the commitment binds message words modulo p (lambda_snark_b200.h, lwe_commit): compare with m mod p.
    // No branch on the outcome (commitment.cpp:223-226 folds the differences the same way).
    const u64 df = d ^ (msgs[b * msg_len + x] % p);
    atomicOr(diff + b, (unsigned long long)df);
}

// K5: out[x] = sum_i c'_i * t_i[x] mod q, c'_i the CENTRED representative of c_i mod p (commitment.cpp:90 reduces the
// coefficient mod the plain modulus; the centred lift keeps the noise growth at |c'| <= p/2 instead of c < p, and
// Delta * p = -1 (mod q) makes c and c - p act identically on the message up to one unit of noise per message unit)
__global__ void __launch_bounds__(256)
lincomb_kernel(const ModParams mp, u64 p, size_t kn, const u64* __restrict__ payloads,
               const u64* __restrict__ coeffs, size_t count, u64* __restrict__ out) {
    const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= kn) return;
    u64 acc = 0;
    for (size_t i = 0; i < count; i++) {
        const u64 c = coeffs[i] % p;
        const u64 cr = c <= p / 2 ? c : mp.q - (p - c);
        acc = addmod(acc, mulmod_exact(cr, payloads[i * kn + x], mp), mp.q);
    }
    out[x] = acc;
}

// K7 standalone: sample i uses draws (2i, 2i+1) of stream (key, dom=kDomSample),
// i.e. one ChaCha block per 4 samples (utils.cpp:95-121 draw order)
__global__ void __launch_bounds__(128)
sample_gaussian_kernel(const ChaChaKey key, const u64* __restrict__ cdf, u32 cdf_n, size_t len,
                       u64* __restrict__ out) {
    const size_t blk = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (blk * 4 >= len) return;
    u32 x[16];
    chacha_block(key, (u32)blk, (u32)(blk >> 32), 0u, kDomSample, x);
#pragma unroll
    for (u32 w = 0; w < 4; w++) {
        const size_t i = blk * 4 + w;
        if (i < len) {
            const u64 u1 = (u64)x[4 * w] | ((u64)x[4 * w + 1] << 32);
            const u32 mag = cdt_magnitude_global(cdf, cdf_n, u1);
            out[i] = (u64)signed_value(mag, x[4 * w + 2] & 1u);
        }
    }
}

// ------------------------------------------------------------------ life cycle
static inline unsigned grid_for(size_t items, unsigned block) { return (unsigned)((items + block - 1) / block); }

size_t lwe_words(const LweContext* c) { return 1 + (size_t)c->k * c->n; }

LweContext* lwe_create(u64 modulus_req, uint32_t n, uint32_t k, double sigma, const uint8_t seed32[32]) {
    // SEAL BFV needs a power-of-two degree >= 1024 (reference test_commitment.cpp:16);
    // here any power of two in [16, 2^17] works (16 = one randomness chunk)
    if (n < 16 || (n & (n - 1)) || n > (1u << kMaxLogN)) { set_error("lwe_context_create: bad ring_degree"); return nullptr; }
    if (k < 1 || k > 16) { set_error("lwe_context_create: bad module_rank"); return nullptr; }
    std::vector<u64> cdf = host::build_cdt(sigma);
    if (cdf.empty()) { set_error("lwe_context_create: bad sigma"); return nullptr; }
    const u64 q = host::ntt_friendly_prime(modulus_req, n)
                      ? modulus_req
                      : (n <= 4096 ? kDefaultModulusSmall : kDefaultModulusLarge);

    LweContext* c = new (std::nothrow) LweContext;
    if (!c) return nullptr;
    c->q = q; c->n = n; c->k = k; c->sigma = sigma;
    c->p = host::plain_modulus(q);
    c->delta = (q - 1) / c->p;
    c->cdf = std::move(cdf);
    host::load_key(seed32, c->key);
    c->ntt = ntt_create(q, n);
    if (!c->ntt) { delete c; return nullptr; }
    c->device = c->ntt->device;
    c->logn = c->ntt->logn;

    const size_t poly = (size_t)n * sizeof(u64);
    bool ok = cuda_ok(cudaMalloc(&c->d_A, poly * k * k), "cudaMalloc(A)") &&
              cuda_ok(cudaMalloc(&c->d_zh, poly * std::max<uint32_t>(k - 1, 1)), "cudaMalloc(z)") &&
              cuda_ok(cudaMalloc(&c->d_cdf, c->cdf.size() * sizeof(u64)), "cudaMalloc(cdf)") &&
              cuda_ok(cudaMemcpy(c->d_cdf, c->cdf.data(), c->cdf.size() * sizeof(u64), cudaMemcpyHostToDevice), "upload cdf");
    u64* d_F = nullptr;
    ok = ok && cuda_ok(cudaMalloc(&d_F, poly * k), "cudaMalloc(f)");
    if (ok) {
        std::vector<u64> tmp(n);
        for (uint32_t i = 0; ok && i + 1 < k; i++) {
            for (uint32_t j = 0; ok && j < k; j++) {
                host::uniform_poly(c->key, i * k + j, q, n, tmp.data());
                ok = cuda_ok(cudaMemcpy(c->d_A + ((size_t)i * k + j) * n, tmp.data(), poly, cudaMemcpyHostToDevice), "upload A");
            }
        }
        for (uint32_t P = 0; ok && P < 2 * k - 1; P++) {
            host::gaussian_poly(c->key, P, c->cdf, q, n, tmp.data());
            u64* dst = P < k - 1 ? c->d_zh + (size_t)P * n : d_F + (size_t)(P - (k - 1)) * n;
            ok = cuda_ok(cudaMemcpy(dst, tmp.data(), poly, cudaMemcpyHostToDevice), "upload trapdoor");
        }
        std::fill(tmp.begin(), tmp.end(), 0);   // host copy of secret material
    }
    cudaStream_t s = c->ntt->stream;
    if (ok && k > 1) ok = ntt_forward_launch(c->ntt, c->d_zh, k - 1, s);
    if (ok) ok = ntt_forward_launch(c->ntt, d_F, k, s);
    if (ok) {
        trapdoor_row_kernel<<<grid_for((size_t)k * n, 256), 256, 0, s>>>(c->ntt->mp, n, k, c->d_zh, d_F, c->d_A);
        ok = cuda_ok(cudaGetLastError(), "trapdoor_row_kernel") && fused_prepare(c, s) &&
             cuda_ok(cudaStreamSynchronize(s), "context setup");
    }
    if (d_F) { cudaMemset(d_F, 0, poly * k); cudaFree(d_F); }
    if (!ok) { lwe_destroy(c); return nullptr; }
    return c;
}

void lwe_destroy(LweContext* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    // commitment.h:34 promises zeroisation of sensitive data
    if (c->d_zh) { cudaMemset(c->d_zh, 0, (size_t)std::max<uint32_t>(c->k - 1, 1) * c->n * sizeof(u64)); cudaFree(c->d_zh); }
    if (c->d_A) cudaFree(c->d_A);
    if (c->d_A2) cudaFree(c->d_A2);
    if (c->d_A2f) cudaFree(c->d_A2f);
    if (c->d_cdf) cudaFree(c->d_cdf);
    for (auto& s : c->scratch) s.release();
    for (auto& s : c->staging) s.release();
    ntt_destroy(c->ntt);
    volatile uint32_t* kp = c->key;
    for (int i = 0; i < 8; i++) kp[i] = 0;
    delete c;
}

// ------------------------------------------------------------------- generic
// Scratch S is allocated and freed in stream order (cudaMallocAsync), so concurrent device-pointer calls on one
// context -- which do not take ctx->mu, they only enqueue -- never share or resize a buffer under a running kernel.
static bool generic_commit_launch(const LweContext* c, const u64* d_msgs, size_t msg_len,
                                  const u64* d_seeds, size_t count, u64* d_out, cudaStream_t s,
                                  const long long* d_s = nullptr, const long long* d_e = nullptr,
                                  uint32_t planes = 1, size_t unit0 = 0) {
    const uint32_t n = c->n, k = c->k;
    const size_t words = lwe_words(c);
    // scratch S: chunk so it stays <= 1 GiB
    const size_t per = (size_t)k * n * sizeof(u64);
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(count, ((size_t)1 << 30) / per));
    void* Sv = nullptr;
    // The library's own stream-ordered pool (one per device), with a release threshold: in the device's default pool
    // (threshold 0) every synchronisation hands the scratch back to the driver and the next call pays a fresh allocation
    // (577 us per single commitment on this path against 115 with a cached block); the default pool is the host
    // application's and is left alone.
    static std::once_flag once[16];
    static cudaMemPool_t pools[16] = {};
    int dev = 0;
    if (!cuda_ok(cudaGetDevice(&dev), "cudaGetDevice") || dev < 0 || dev >= 16) return false;
    std::call_once(once[dev], [dev] {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t pool = nullptr;
        if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
            unsigned long long keep = 1ull << 31;                          // up to 2 GiB stay cached (a scratch is <= 1 GiB)
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            pools[dev] = pool;
        }
        cudaGetLastError();
    });
    if (pools[dev]) {
        if (!cuda_ok(cudaMallocFromPoolAsync(&Sv, chunk * per, pools[dev], s), "cudaMallocFromPoolAsync(S)")) return false;
    } else if (!cuda_ok(cudaMallocAsync(&Sv, chunk * per, s), "cudaMallocAsync(S)")) return false;
    u64* S = static_cast<u64*>(Sv);
    const ChaChaKey key = make_key(c);
    bool ok = true;
    for (size_t done = 0; ok && done < count; done += chunk) {
        const size_t cnt = std::min(chunk, count - done);
        u64* out = d_out + done * words;
        if (d_s) {
            load_se_kernel<<<grid_for(cnt * k * n, 256), 256, 0, s>>>(c->q, (size_t)k * n, d_s + done * k * n, d_e + done * k * n,
                                                                     cnt, S, out, words);
            ok = cuda_ok(cudaGetLastError(), "load_se_kernel");
        } else {
            sample_se_kernel<<<grid_for(cnt * (n >> 4), 128), 128, 0, s>>>(key, c->d_cdf, (u32)c->cdf.size(), c->q, n, k,
                                                                          d_seeds + done, cnt, S, out, words);
            ok = cuda_ok(cudaGetLastError(), "sample_se_kernel");
        }
        ok = ok && ntt_forward_launch(c->ntt, S, cnt * k, s);
        if (ok) {
            matvec_kernel<16><<<grid_for(cnt * n, 256), 256, 0, s>>>(c->ntt->mp, c->d_A, n, k, cnt, S);
            ok = cuda_ok(cudaGetLastError(), "matvec_kernel");
        }
        ok = ok && ntt_inverse_launch(c->ntt, S, cnt * k, s);
        if (ok) {
            // messages are addressed with their true stride msg_len; only the first min(msg_len, n) words are used.
            // With digit planes the message row of a commitment follows from its unit index, so the base stays put.
            const u64* mbase = planes > 1 ? d_msgs : d_msgs + done * msg_len;
            finalize_kernel<<<grid_for(cnt * k * n, 256), 256, 0, s>>>(c->ntt->mp, c->delta, c->p, n, k, S, mbase, msg_len, cnt,
                                                                      out, words, planes, planes > 1 ? unit0 + done : 0);
            ok = cuda_ok(cudaGetLastError(), "finalize_kernel");
        }
    }
    return cuda_ok(cudaFreeAsync(Sv, s), "cudaFreeAsync(S)") && ok;
}

// explicit mode: s, e supplied by the caller (device pointers, [count][k][n] two's complement)
bool lwe_commit_explicit_launch(const LweContext* c, const u64* d_msgs, size_t msg_len, const int64_t* d_s,
                                const int64_t* d_e, size_t count, u64* d_out, cudaStream_t stream) {
    if (count == 0) return true;
    return generic_commit_launch(c, d_msgs, msg_len, nullptr, count, d_out, stream,
                                 reinterpret_cast<const long long*>(d_s), reinterpret_cast<const long long*>(d_e));
}

// explicit mode from host memory: a parity / measurement entry, so plain synchronous staging per chunk
bool lwe_commit_explicit_host(const LweContext* c, const u64* msgs, size_t msg_len, const int64_t* s_in,
                              const int64_t* e_in, size_t count, u64* out) {
    if (count == 0) return true;
    std::lock_guard<std::mutex> lock(c->mu);
    if (!cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return false;
    const size_t words = lwe_words(c), kn = (size_t)c->k * c->n;
    const size_t chunk = std::min<size_t>(count, 64);
    cudaStream_t st = c->ntt->stream;
    if (!c->scratch[1].reserve(chunk * std::max<size_t>(msg_len, 1) * sizeof(u64))) return false;
    if (!c->scratch[2].reserve(chunk * kn * 2 * sizeof(int64_t))) return false;
    if (!c->scratch[3].reserve(chunk * words * sizeof(u64))) return false;
    u64* dm = static_cast<u64*>(c->scratch[1].ptr);
    int64_t* ds = static_cast<int64_t*>(c->scratch[2].ptr);
    int64_t* de = ds + chunk * kn;
    u64* dout = static_cast<u64*>(c->scratch[3].ptr);
    bool ok = true;
    for (size_t done = 0; ok && done < count; done += chunk) {
        const size_t cnt = std::min(chunk, count - done);
        if (msg_len) ok = cuda_ok(cudaMemcpyAsync(dm, msgs + done * msg_len, cnt * msg_len * sizeof(u64), cudaMemcpyHostToDevice, st), "H2D msgs");
        ok = ok && cuda_ok(cudaMemcpyAsync(ds, s_in + done * kn, cnt * kn * sizeof(int64_t), cudaMemcpyHostToDevice, st), "H2D s") &&
             cuda_ok(cudaMemcpyAsync(de, e_in + done * kn, cnt * kn * sizeof(int64_t), cudaMemcpyHostToDevice, st), "H2D e") &&
             generic_commit_launch(c, dm, msg_len, nullptr, cnt, dout, st, reinterpret_cast<const long long*>(ds),
                                   reinterpret_cast<const long long*>(de)) &&
             cuda_ok(cudaMemcpyAsync(out + done * words, dout, cnt * words * sizeof(u64), cudaMemcpyDeviceToHost, st), "D2H") &&
             cuda_ok(cudaStreamSynchronize(st), "sync");
    }
    return ok;
}

bool lwe_commit_launch(const LweContext* c, const u64* d_msgs, size_t msg_len, const u64* d_seeds,
                       size_t count, u64* d_out, cudaStream_t stream, uint32_t planes, size_t unit0) {
    if (count == 0) return true;
    if (planes < 1 || planes > 4) { set_error("lwe_commit: digit planes must be 1..4"); return false; }
    const bool fused_ok = fused_commit_supported(c);
    if (c->commit_path == 2 && !fused_ok) { set_error("fused commit path does not support this context"); return false; }
    if (fused_ok && c->commit_path != 1)
        return fused_commit_launch(c, d_msgs, msg_len, d_seeds, count, d_out, stream, planes, unit0);
    return generic_commit_launch(c, d_msgs, msg_len, d_seeds, count, d_out, stream, nullptr, nullptr, planes, unit0);
}

void plane_divisors(u64 p, u64 pdiv[4], u64 pdinv[4]) {
    unsigned __int128 d = 1;
    for (int l = 0; l < 4; l++) {
        const bool fits = d < ((unsigned __int128)1 << 63);
        pdiv[l] = fits ? (u64)d : 0;
        pdinv[l] = fits ? ~0ull / (u64)d : 0;
        d *= p;
    }
}

uint32_t message_planes(u64 p, u64 modulus) {
    unsigned __int128 d = 1;
    for (uint32_t L = 1; L <= 4; L++) {
        d *= p;
        if (d >= modulus) return L;
    }
    return 0;
}

// Noise of a fresh commitment seen by the trapdoor: f^T s + z^T e (+ the last error term), i.e. 2k - 1 ring products of
// sigma-Gaussians plus one sample: variance (2k - 1) n sigma^4 + sigma^2 per coefficient.  A combination with centred
// coefficients c'_i has noise at most sum |c'_i| times the single-commitment bound, taken at 12 standard deviations;
// a plaintext wrap-around costs one more unit each (Delta * p = -1 mod q), at most sum |c'_i| of them; the result
// decodes while the total stays below Delta / 2.
u64 lincomb_budget(const LweContext* c) {
    const double s2 = c->sigma * c->sigma;
    const double bound = 12.0 * std::sqrt((2.0 * c->k - 1.0) * c->n * s2 * s2 + s2) + 1.0;
    const double budget = std::floor((double)(c->delta / 2) / bound);
    return budget < 1.0 ? 1 : (u64)budget;
}

// Host path: chunks of 256 commitments rotate over three streams, each running
// H2D (messages + seeds) -> kernel(s) -> D2H (containers) in order, so the D2H
// copy engine -- the bound: 64 KiB out per commitment -- never idles.  The
// generic path shares one S scratch buffer and therefore stays on one stream.
static bool is_pageable(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}

// Pageable caller memory, fused path, batches of at least one chunk: three slots of page-locked staging, the copy pool
// moves chunk i in and chunk i-3 out while the GPU works on the chunks in between (the calling thread orchestrates and
// copies too).  ~4x the driver's own pageable staging.
static bool lwe_commit_host_staged(const LweContext* c, const u64* msgs, size_t msg_len, const u64* seeds,
                                   size_t count, u64* out) {
    constexpr int NB = 3;
    const size_t words = lwe_words(c);
    const size_t chunk = count >= 1024 ? 256 : 64;
    const size_t in_words = chunk * (msg_len + 1);                         // messages, then the seeds
    for (int b = 0; b < NB; b++) {
        if (!c->staging[2 * b].reserve(in_words * sizeof(u64))) return false;
        if (!c->staging[2 * b + 1].reserve(chunk * words * sizeof(u64))) return false;
        if (!c->scratch[1 + 3 * b].reserve(in_words * sizeof(u64))) return false;
        if (!c->scratch[3 + 3 * b].reserve(chunk * words * sizeof(u64))) return false;
    }
    cudaStream_t streams[NB] = {c->ntt->copy_streams[0], c->ntt->copy_streams[1], c->ntt->stream};
    CopyPool& pool = CopyPool::get();
    const size_t nchunks = (count + chunk - 1) / chunk;
    bool ok = true;
    for (size_t i = 0; ok && i < nchunks + NB; i++) {
        const int b = (int)(i % NB);
        cudaStream_t s = streams[b];
        u64* pin_in = static_cast<u64*>(c->staging[2 * b].ptr);
        u64* pin_out = static_cast<u64*>(c->staging[2 * b + 1].ptr);
        if (i >= NB) {                                                     // retire chunk i - NB from this slot
            const size_t j = i - NB, done = j * chunk, cnt = std::min(chunk, count - done);
            ok = cuda_ok(cudaStreamSynchronize(s), "sync");
            if (ok) pool.copy(out + done * words, pin_out, cnt * words * sizeof(u64));
        }
        if (ok && i < nchunks) {
            const size_t done = i * chunk, cnt = std::min(chunk, count - done);
            u64* din = static_cast<u64*>(c->scratch[1 + 3 * b].ptr);
            u64* dout = static_cast<u64*>(c->scratch[3 + 3 * b].ptr);
            if (msg_len) pool.copy(pin_in, msgs + done * msg_len, cnt * msg_len * sizeof(u64));
            std::memcpy(pin_in + cnt * msg_len, seeds + done, cnt * sizeof(u64));
            ok = cuda_ok(cudaMemcpyAsync(din, pin_in, cnt * (msg_len + 1) * sizeof(u64), cudaMemcpyHostToDevice, s), "H2D") &&
                 lwe_commit_launch(c, din, msg_len, din + cnt * msg_len, cnt, dout, s) &&
                 cuda_ok(cudaMemcpyAsync(pin_out, dout, cnt * words * sizeof(u64), cudaMemcpyDeviceToHost, s), "D2H");
        }
    }
    for (int b = 0; b < NB; b++) ok = cuda_ok(cudaStreamSynchronize(streams[b]), "sync") && ok;
    return ok;
}

bool lwe_commit_host(const LweContext* c, const u64* msgs, size_t msg_len, const u64* seeds,
                     size_t count, u64* out) {
    if (count == 0) return true;
    std::lock_guard<std::mutex> lock(c->mu);
    if (!cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return false;
    if (count >= 128 && fused_commit_supported(c) && c->commit_path != 1 && (is_pageable(out) || is_pageable(msgs)))
        return lwe_commit_host_staged(c, msgs, msg_len, seeds, count, out);
    const size_t words = lwe_words(c);
    const size_t eff_len = std::max<size_t>(msg_len, 1);
    const size_t chunk = std::min<size_t>(count, kCommitHostChunk);
    const bool fused = fused_commit_supported(c) && c->commit_path != 1;
    // A few commitments per call (the reference's lwe_commit is one): no device buffers and no copy commands.  Messages
    // and seeds go into a page-locked buffer, the fused kernel reads them and writes the containers through the device
    // mapping of page-locked memory, one synchronisation, one copy back.
    if (fused && count <= 4) {
        const size_t in_words = count * (eff_len + 1);
        if (!c->staging[0].reserve(in_words * sizeof(u64)) || !c->staging[1].reserve(count * words * sizeof(u64))) return false;
        u64* pin_in = static_cast<u64*>(c->staging[0].ptr);
        u64* pin_out = static_cast<u64*>(c->staging[1].ptr);
        if (msg_len) std::memcpy(pin_in, msgs, count * msg_len * sizeof(u64));
        std::memcpy(pin_in + count * eff_len, seeds, count * sizeof(u64));
        cudaStream_t s = c->ntt->stream;
        const bool ok = lwe_commit_launch(c, pin_in, msg_len, pin_in + count * eff_len, count, pin_out, s) &&
                        cuda_ok(cudaStreamSynchronize(s), "sync");
        if (ok) std::memcpy(out, pin_out, count * words * sizeof(u64));
        return ok;
    }
    const int nbuf = (fused && count > chunk) ? 3 : 1;
    cudaStream_t streams[3] = {c->ntt->copy_streams[0], c->ntt->copy_streams[1], c->ntt->stream};
    for (int b = 0; b < nbuf; b++) {
        if (!c->scratch[1 + 3 * b].reserve(chunk * eff_len * sizeof(u64))) return false;
        if (!c->scratch[2 + 3 * b].reserve(chunk * sizeof(u64))) return false;
        if (!c->scratch[3 + 3 * b].reserve(chunk * words * sizeof(u64))) return false;
    }
    bool ok = true;
    int b = 0;
    for (size_t done = 0; ok && done < count; done += chunk) {
        const size_t cnt = std::min(chunk, count - done);
        cudaStream_t s = nbuf == 1 ? c->ntt->stream : streams[b];
        u64* dm = static_cast<u64*>(c->scratch[1 + 3 * b].ptr);
        u64* ds = static_cast<u64*>(c->scratch[2 + 3 * b].ptr);
        u64* dout = static_cast<u64*>(c->scratch[3 + 3 * b].ptr);
        if (msg_len) ok = cuda_ok(cudaMemcpyAsync(dm, msgs + done * msg_len, cnt * msg_len * sizeof(u64), cudaMemcpyHostToDevice, s), "H2D msgs");
        ok = ok && cuda_ok(cudaMemcpyAsync(ds, seeds + done, cnt * sizeof(u64), cudaMemcpyHostToDevice, s), "H2D seeds") &&
             lwe_commit_launch(c, dm, msg_len, ds, cnt, dout, s) &&
             cuda_ok(cudaMemcpyAsync(out + done * words, dout, cnt * words * sizeof(u64), cudaMemcpyDeviceToHost, s), "D2H");
        if (nbuf > 1) b = (b + 1) % nbuf;
    }
    for (int i = 0; i < (nbuf == 1 ? 1 : 3); i++)
        ok = cuda_ok(cudaStreamSynchronize(nbuf == 1 ? c->ntt->stream : streams[i]), "sync") && ok;
    return ok;
}

bool lwe_verify_host(const LweContext* c, const u64* comm_words, const u64* msgs, size_t msg_len,
                     size_t count, int* results) {
    if (count == 0) return true;
    const uint32_t n = c->n, k = c->k;
    if (msg_len > n) {                       // commitment.cpp:219-221: decoded.size() < msg_len -> 0
        for (size_t i = 0; i < count; i++) results[i] = 0;
        // (container errors still win, as load() precedes decode in the reference)
    }
    std::lock_guard<std::mutex> lock(c->mu);
    if (!cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return false;
    const size_t words = lwe_words(c);
    const size_t cmp_len = std::min<size_t>(msg_len, n);
    const size_t chunk = std::min<size_t>(count, 1024);
    cudaStream_t s = c->ntt->stream;
    if (!c->scratch[0].reserve(chunk * words * sizeof(u64))) return false;                       // containers
    if (!c->scratch[1].reserve(chunk * std::max<size_t>(cmp_len, 1) * sizeof(u64))) return false; // messages
    if (!c->scratch[2].reserve(chunk * std::max<uint32_t>(k - 1, 1) * (size_t)n * sizeof(u64))) return false;  // T
    if (!c->scratch[3].reserve(chunk * (size_t)n * sizeof(u64))) return false;                    // U
    if (!c->scratch[4].reserve(chunk * (sizeof(int) + sizeof(unsigned long long)))) return false; // flags
    u64* d_comm = static_cast<u64*>(c->scratch[0].ptr);
    u64* d_msg = static_cast<u64*>(c->scratch[1].ptr);
    u64* d_T = static_cast<u64*>(c->scratch[2].ptr);
    u64* d_U = static_cast<u64*>(c->scratch[3].ptr);
    unsigned long long* d_diff = static_cast<unsigned long long*>(c->scratch[4].ptr);
    int* d_inv = reinterpret_cast<int*>(d_diff + chunk);
    std::vector<unsigned long long> h_diff(chunk);
    std::vector<int> h_inv(chunk);
    std::vector<u64> packed;
    bool ok = true;
    // A few openings per call (the reference's lwe_verify_opening is one): containers and messages are copied into a
    // page-locked buffer and the kernels read them through its device mapping -- no copy commands on the way in.  The fused
    // kernel requests the 16 words a thread owns of a row together, so a row costs about one PCIe round trip (while it walked
    // the container a word per step it was given two asynchronous copies instead: 78 us per call against 66 now).
    const bool small = count <= 4;
    const bool fused_v = fused_verify_supported(c);
    if (small) {
        if (!c->staging[2].reserve(count * (words + std::max<size_t>(cmp_len, 1)) * sizeof(u64))) return false;
        u64* pin = static_cast<u64*>(c->staging[2].ptr);
        std::memcpy(pin, comm_words, count * words * sizeof(u64));
        for (size_t i = 0; i < count && cmp_len; i++)
            std::memcpy(pin + count * words + i * cmp_len, msgs + i * msg_len, cmp_len * sizeof(u64));
        d_comm = pin;
        d_msg = pin + count * words;
    }
    for (size_t done = 0; ok && done < count; done += chunk) {
        const size_t cnt = std::min(chunk, count - done);
        if (!small)
            ok = cuda_ok(cudaMemcpyAsync(d_comm, comm_words + done * words, cnt * words * sizeof(u64), cudaMemcpyHostToDevice, s), "H2D comm");
        // small calls: the two flag words per commitment live in page-locked memory and the kernels OR into them through
        // the device mapping -- no memset command, no copies back
        unsigned long long* pin_flags = nullptr;
        if (small) {
            if (!c->staging[3].reserve(chunk * (sizeof(int) + sizeof(unsigned long long)))) return false;
            pin_flags = static_cast<unsigned long long*>(c->staging[3].ptr);
            std::memset(pin_flags, 0, chunk * (sizeof(int) + sizeof(unsigned long long)));
            d_diff = pin_flags;
            d_inv = reinterpret_cast<int*>(pin_flags + chunk);
        } else
        ok = ok && cuda_ok(cudaMemsetAsync(d_diff, 0, chunk * (sizeof(int) + sizeof(unsigned long long)), s), "memset");
        if (ok && cmp_len && !small) {
            // compare only the first cmp_len words of each message (stride msg_len on the host)
            const u64* src = msgs + done * msg_len;
            if (cmp_len != msg_len) {
                packed.resize(cnt * cmp_len);
                for (size_t i = 0; i < cnt; i++) std::memcpy(&packed[i * cmp_len], src + i * msg_len, cmp_len * sizeof(u64));
                src = packed.data();
            }
            ok = cuda_ok(cudaMemcpyAsync(d_msg, src, cnt * cmp_len * sizeof(u64), cudaMemcpyHostToDevice, s), "H2D msgs");
        }
        if (!ok) break;
        if (fused_v) {
            ok = fused_verify_launch(c, d_comm, words, d_msg, cmp_len, cnt, d_diff, d_inv, s);
        } else {
        verify_prepare_kernel<<<grid_for(cnt * k * n, 256), 256, 0, s>>>(c->q, n, k, d_comm, words, cnt, d_T, d_inv);
        ok = cuda_ok(cudaGetLastError(), "verify_prepare_kernel");
        if (ok && k > 1) ok = ntt_forward_launch(c->ntt, d_T, cnt * (k - 1), s);
        if (ok) {
            verify_dot_kernel<<<grid_for(cnt * n, 256), 256, 0, s>>>(c->ntt->mp, n, k, c->d_zh, d_T, cnt, d_U);
            ok = cuda_ok(cudaGetLastError(), "verify_dot_kernel");
        }
        if (ok && k > 1) ok = ntt_inverse_launch(c->ntt, d_U, cnt, s);
        if (ok && cmp_len) {
            verify_decode_kernel<<<grid_for(cnt * cmp_len, 256), 256, 0, s>>>(c->q, c->delta, c->p, n, k, d_U, d_comm, words,
                                                                            d_msg, cmp_len, cnt, d_diff);
            ok = cuda_ok(cudaGetLastError(), "verify_decode_kernel");
        }
        }
        if (small) {
            ok = ok && cuda_ok(cudaStreamSynchronize(s), "sync");
            if (ok) {
                std::memcpy(h_diff.data(), d_diff, cnt * sizeof(unsigned long long));
                std::memcpy(h_inv.data(), d_inv, cnt * sizeof(int));
            }
        } else
        ok = ok && cuda_ok(cudaMemcpyAsync(h_diff.data(), d_diff, cnt * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s), "D2H") &&
             cuda_ok(cudaMemcpyAsync(h_inv.data(), d_inv, cnt * sizeof(int), cudaMemcpyDeviceToHost, s), "D2H") &&
             cuda_ok(cudaStreamSynchronize(s), "sync");
        if (!ok) break;
        for (size_t i = 0; i < cnt; i++) {
            if (h_inv[i]) results[done + i] = -1;
            else if (msg_len > n) results[done + i] = 0;
            else results[done + i] = h_diff[i] == 0 ? 1 : 0;
        }
    }
    return ok;
}

bool lwe_lincomb_host(const LweContext* c, const u64* payloads, const u64* coeffs, size_t count,
                      u64* out_payload) {
    // decodability budget (lambda_snark_b200.h, lwe_linear_combine): beyond it lwe_verify_opening of the result would fail
    // for honest inputs, so the combination is refused instead of returned
    {
        const u64 budget = lincomb_budget(c);
        u64 sum = 0;
        for (size_t i = 0; i < count && sum <= budget; i++) {
            const u64 cm = coeffs[i] % c->p;
            sum += cm <= c->p / 2 ? cm : c->p - cm;
        }
        if (sum > budget) {
            set_error("lwe_linear_combine: sum of |centred coefficients| exceeds the decodable budget (lsr_lwe_lincomb_budget)");
            return false;
        }
    }
    std::lock_guard<std::mutex> lock(c->mu);
    if (!cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return false;
    const size_t kn = (size_t)c->k * c->n;
    cudaStream_t s = c->ntt->stream;
    if (!c->scratch[0].reserve(count * kn * sizeof(u64))) return false;
    if (!c->scratch[1].reserve(count * sizeof(u64))) return false;
    if (!c->scratch[2].reserve(kn * sizeof(u64))) return false;
    u64* d_p = static_cast<u64*>(c->scratch[0].ptr);
    u64* d_c = static_cast<u64*>(c->scratch[1].ptr);
    u64* d_o = static_cast<u64*>(c->scratch[2].ptr);
    bool ok = cuda_ok(cudaMemcpyAsync(d_p, payloads, count * kn * sizeof(u64), cudaMemcpyHostToDevice, s), "H2D") &&
              cuda_ok(cudaMemcpyAsync(d_c, coeffs, count * sizeof(u64), cudaMemcpyHostToDevice, s), "H2D");
    if (ok) {
        lincomb_kernel<<<grid_for(kn, 256), 256, 0, s>>>(c->ntt->mp, c->p, kn, d_p, d_c, count, d_o);
        ok = cuda_ok(cudaGetLastError(), "lincomb_kernel");
    }
    return ok && cuda_ok(cudaMemcpyAsync(out_payload, d_o, kn * sizeof(u64), cudaMemcpyDeviceToHost, s), "D2H") &&
           cuda_ok(cudaStreamSynchronize(s), "sync");
}

// test hook: s, e of one commitment through the generic sampler kernel
bool lwe_sample_se_host(const LweContext* c, u64 seed, int64_t* s_out, int64_t* e_out) {
    std::lock_guard<std::mutex> lock(c->mu);
    if (!cuda_ok(cudaSetDevice(c->device), "cudaSetDevice")) return false;
    const uint32_t n = c->n, k = c->k;
    const size_t kn = (size_t)k * n, words = 1 + kn;
    cudaStream_t st = c->ntt->stream;
    if (!c->scratch[0].reserve(kn * sizeof(u64))) return false;
    if (!c->scratch[1].reserve(words * sizeof(u64))) return false;
    if (!c->scratch[2].reserve(sizeof(u64))) return false;
    u64* S = static_cast<u64*>(c->scratch[0].ptr);
    u64* O = static_cast<u64*>(c->scratch[1].ptr);
    u64* dseed = static_cast<u64*>(c->scratch[2].ptr);
    bool ok = cuda_ok(cudaMemcpyAsync(dseed, &seed, sizeof(u64), cudaMemcpyHostToDevice, st), "H2D");
    if (!ok) return false;
    sample_se_kernel<<<grid_for(n >> 4, 128), 128, 0, st>>>(make_key(c), c->d_cdf, (u32)c->cdf.size(), c->q, n, k, dseed, 1, S, O, words);
    std::vector<u64> hs(kn), he(kn);
    ok = cuda_ok(cudaGetLastError(), "sample_se_kernel") &&
         cuda_ok(cudaMemcpyAsync(hs.data(), S, kn * sizeof(u64), cudaMemcpyDeviceToHost, st), "D2H") &&
         cuda_ok(cudaMemcpyAsync(he.data(), O + 1, kn * sizeof(u64), cudaMemcpyDeviceToHost, st), "D2H") &&
         cuda_ok(cudaStreamSynchronize(st), "sync");
    if (!ok) return false;
    const u64 half = c->q / 2;
    for (size_t i = 0; i < kn; i++) {
        s_out[i] = hs[i] > half ? -(int64_t)(c->q - hs[i]) : (int64_t)hs[i];
        e_out[i] = he[i] > half ? -(int64_t)(c->q - he[i]) : (int64_t)he[i];
    }
    return true;
}

bool sample_gaussian_host(u64* out, size_t len, double sigma, const uint8_t seed32[32]) {
    std::vector<u64> cdf = host::build_cdt(sigma);
    if (cdf.empty()) return false;
    if (!cuda_ok(cudaSetDevice(current_device_choice()), "cudaSetDevice")) return false;
    ChaChaKey key;
    host::load_key(seed32, key.k);
    u64 *d_cdf = nullptr, *d_out = nullptr;
    cudaStream_t s = nullptr;
    bool ok = cuda_ok(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate") &&
              cuda_ok(cudaMalloc(&d_cdf, cdf.size() * sizeof(u64)), "cudaMalloc") &&
              cuda_ok(cudaMalloc(&d_out, len * sizeof(u64)), "cudaMalloc") &&
              cuda_ok(cudaMemcpyAsync(d_cdf, cdf.data(), cdf.size() * sizeof(u64), cudaMemcpyHostToDevice, s), "H2D");
    if (ok) {
        sample_gaussian_kernel<<<grid_for((len + 3) / 4, 128), 128, 0, s>>>(key, d_cdf, (u32)cdf.size(), len, d_out);
        ok = cuda_ok(cudaGetLastError(), "sample_gaussian_kernel") &&
             cuda_ok(cudaMemcpyAsync(out, d_out, len * sizeof(u64), cudaMemcpyDeviceToHost, s), "D2H") &&
             cuda_ok(cudaStreamSynchronize(s), "sync");
    }
    if (d_cdf) cudaFree(d_cdf);
    if (d_out) { cudaMemsetAsync(d_out, 0, len * sizeof(u64), s); cudaStreamSynchronize(s); cudaFree(d_out); }
    if (s) cudaStreamDestroy(s);
    for (int i = 0; i < 8; i++) key.k[i] = 0;
    return ok;
}

}  // namespace lsr
