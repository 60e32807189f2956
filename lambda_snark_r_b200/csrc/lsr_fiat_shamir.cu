// lsr_fiat_shamir.cu -- SURVEY row N2: the Fiat-Shamir transcript and the polynomial evaluations of
// prove_r1cs on the device, so that a batch of proofs needs no commitment on the host before its
// challenges exist.
//
// Replaces, per statement of a batch:
//   Challenge::derive          rust-api/lambda-snark/src/challenge.rs:102-134  -> fs_challenge_kernel
//       alpha = LE64(SHA3-256("LAMBDA-SNARK-R-FS-v1" || #pub || pub || #words || words)[0..8]) mod q
//       beta  = the same with the single public input alpha (lib.rs:760-767)
//   R1CS::eval_poly            rust-api/lambda-snark/src/r1cs.rs:362-373       -> poly_eval_kernel
//       sum_i c_i x^i mod q (the reference's running-power loop; any evaluation order gives the same residue)
//
// SHA3 is sequential in the message (482 Keccak-f permutations for a 64 KiB commitment), so the
// parallelism is across statements: one thread per statement, state in registers (lsr_keccak.h; the same
// source is checked on the host against hashlib by the CPU tests).  Evaluations: one CTA per
// (polynomial, point), each thread a contiguous run of coefficients by Horner, scaled by x^start, tree sum.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "lsr_arith.cuh"
#include "lsr_engine.h"
#include "lsr_keccak.h"

namespace lsr {

constexpr int kFsThreads = 32;       // one warp per CTA: statements spread over as many SMs as possible
constexpr int kEvalThreads = 256;

// containers [count][words]; pub [count][n_pub]; ab [count][2] = (alpha, beta); hashes [count][2][4]
// chain = 0: only alpha (beta slot and second hash are written as zero)
__global__ void __launch_bounds__(kFsThreads)
fs_challenge_kernel(const u64* __restrict__ pub, size_t n_pub, const u64* __restrict__ containers, size_t words,
                    size_t count, u64 modulus, int chain, u64* __restrict__ ab, u64* __restrict__ hashes) {
    const size_t i = (size_t)blockIdx.x * kFsThreads + threadIdx.x;
    if (i >= count) return;
    const u64* w = containers + i * words;
    u64 h[4];
    fs_sha3_256(FsTranscript{pub + i * n_pub, (u64)n_pub, w, (u64)words}, h);
    const u64 alpha = h[0] % modulus;
    ab[2 * i] = alpha;
#pragma unroll
    for (int j = 0; j < 4; j++) hashes[8 * i + j] = h[j];
    u64 beta = 0;
    u64 g[4] = {0, 0, 0, 0};
    if (chain) {
        const u64 one_pub[1] = {alpha};
        fs_sha3_256(FsTranscript{one_pub, 1, w, (u64)words}, g);
        beta = g[0] % modulus;
    }
    ab[2 * i + 1] = beta;
#pragma unroll
    for (int j = 0; j < 4; j++) hashes[8 * i + 4 + j] = g[j];
}

// Small batches: one WARP per statement (lsr_keccak.h, lane-parallel form).  A statement then takes ~1/3 of the time of
// the one-thread form; at ~1 M statements/s the warps saturate the issue ports, so from kFsWarpMaxCount statements on the
// one-thread-per-statement kernel above is the faster one again.
constexpr int kFsWarpCta = 128;            // four statements per CTA
constexpr size_t kFsWarpMaxCount = 4096;
__global__ void __launch_bounds__(kFsWarpCta)
fs_challenge_warp_kernel(const u64* __restrict__ pub, size_t n_pub, const u64* __restrict__ containers, size_t words,
                         size_t count, u64 modulus, int chain, u64* __restrict__ ab, u64* __restrict__ hashes) {
    const size_t i = (size_t)blockIdx.x * (kFsWarpCta / 32) + (threadIdx.x >> 5);     // uniform over the warp
    if (i >= count) return;
    const unsigned lane = threadIdx.x & 31u;
    const u64* w = containers + i * words;
    u64 h[4];
    fs_sha3_256_warp(FsTranscript{pub + i * n_pub, (u64)n_pub, w, (u64)words}, h);
    const u64 alpha = h[0] % modulus;
    u64 beta = 0;
    u64 g[4] = {0, 0, 0, 0};
    if (chain) {
        const u64 one_pub[1] = {alpha};
        fs_sha3_256_warp(FsTranscript{one_pub, 1, w, (u64)words}, g);
        beta = g[0] % modulus;
    }
    if (lane == 0) {
        ab[2 * i] = alpha;
        ab[2 * i + 1] = beta;
#pragma unroll
        for (int j = 0; j < 4; j++) { hashes[8 * i + j] = h[j]; hashes[8 * i + 4 + j] = g[j]; }
    }
}

// out[p][j] = sum_i coeffs[p][i] * x^i, x = points[(p % point_rows)][j];  grid = (npts, polys of this launch);
// p0 = index of the launch's first polynomial (batches above the 65 535 limit of gridDim.y go in several launches)
__global__ void __launch_bounds__(kEvalThreads)
poly_eval_kernel(const ModParams mp, const u64* __restrict__ coeffs, size_t len, const u64* __restrict__ points,
                 size_t npts, size_t point_rows, u64* __restrict__ out, size_t p0) {
    __shared__ u64 part[kEvalThreads];
    const size_t p = p0 + blockIdx.y, j = blockIdx.x;
    const u64 x = reduce64(points[(p % point_rows) * npts + j], mp);
    const u64* __restrict__ c = coeffs + p * len;
    const size_t per = (len + kEvalThreads - 1) / kEvalThreads;
    const size_t lo = (size_t)threadIdx.x * per;
    const size_t hi = lo + per < len ? lo + per : len;
    u64 acc = 0;
    if (lo < hi) {
        for (size_t i = hi; i-- > lo;) acc = field_add(field_mul(acc, x, mp), reduce64(c[i], mp), mp);
        // * x^lo, square and multiply
        u64 base = x, pw = 1;
        for (size_t e = lo; e; e >>= 1) {
            if (e & 1) pw = field_mul(pw, base, mp);
            base = field_mul(base, base, mp);
        }
        acc = field_mul(acc, pw, mp);
    }
    part[threadIdx.x] = acc;
    __syncthreads();
    for (int s = kEvalThreads / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) part[threadIdx.x] = field_add(part[threadIdx.x], part[threadIdx.x + s], mp);
        __syncthreads();
    }
    if (threadIdx.x == 0) out[p * npts + j] = part[0];
}

bool fs_challenge_launch(const u64* d_pub, size_t n_pub, const u64* d_containers, size_t words, size_t count,
                         u64 modulus, bool chain, u64* d_ab, u64* d_hashes, cudaStream_t s) {
    if (count == 0) return true;
    if (modulus == 0) { set_error("fs_challenge: modulus 0"); return false; }
#ifdef LSR_PROFILING    // tools/ build only: 1 = thread per statement, 2 = warp per statement
    static const int force = [] { const char* e = std::getenv("LSR_FS_KERNEL"); return e ? std::atoi(e) : 0; }();
#else
    constexpr int force = 0;
#endif
    const bool warp = force == 2 || (force != 1 && count < kFsWarpMaxCount);
    if (warp) {
        const size_t blocks = (count + kFsWarpCta / 32 - 1) / (kFsWarpCta / 32);
        if (blocks > 0x7fffffffull) { set_error("batch too large for one launch"); return false; }
        fs_challenge_warp_kernel<<<(unsigned)blocks, kFsWarpCta, 0, s>>>(d_pub, n_pub, d_containers, words, count, modulus,
                                                                        chain ? 1 : 0, d_ab, d_hashes);
        return cuda_ok(cudaGetLastError(), "fs_challenge_warp_kernel");
    }
    const size_t blocks = (count + kFsThreads - 1) / kFsThreads;
    if (blocks > 0x7fffffffull) { set_error("batch too large for one launch"); return false; }
    fs_challenge_kernel<<<(unsigned)blocks, kFsThreads, 0, s>>>(d_pub, n_pub, d_containers, words, count, modulus,
                                                              chain ? 1 : 0, d_ab, d_hashes);
    return cuda_ok(cudaGetLastError(), "fs_challenge_kernel");
}

bool poly_eval_launch(u64 modulus, const u64* d_coeffs, size_t len, size_t polys, const u64* d_points, size_t npts,
                      size_t point_rows, u64* d_out, cudaStream_t s) {
    if (polys == 0 || npts == 0) return true;
    if (modulus < 2 || (modulus != kGoldilocks && (modulus >> 61))) { set_error("poly_eval: unsupported modulus"); return false; }
    if (len == 0 || point_rows == 0) { set_error("poly_eval: empty polynomial"); return false; }
    if (npts > 0x7fffffffull) { set_error("poly_eval: too many points"); return false; }
    const ModParams mp = host::make_mod_params(modulus, 1);
    for (size_t p0 = 0; p0 < polys; p0 += 65535) {
        const size_t cnt = std::min<size_t>(65535, polys - p0);
        poly_eval_kernel<<<dim3((unsigned)npts, (unsigned)cnt), kEvalThreads, 0, s>>>(mp, d_coeffs, len, d_points, npts,
                                                                                     point_rows, d_out, p0);
        if (!cuda_ok(cudaGetLastError(), "poly_eval_kernel")) return false;
    }
    return true;
}

// host-pointer forms (tests, small batches): stage through temporary device buffers on the chosen device
bool fs_challenge_host(const u64* pub, size_t n_pub, const u64* containers, size_t words, size_t count, u64 modulus,
                       bool chain, u64* ab, u64* hashes) {
    if (count == 0) return true;
    if (!cuda_ok(cudaSetDevice(current_device_choice()), "cudaSetDevice")) return false;
    u64 *d_pub = nullptr, *d_c = nullptr, *d_ab = nullptr, *d_h = nullptr;
    bool ok = cuda_ok(cudaMalloc(&d_pub, std::max<size_t>(count * n_pub, 1) * 8), "cudaMalloc") &&
              cuda_ok(cudaMalloc(&d_c, std::max<size_t>(count * words, 1) * 8), "cudaMalloc") &&
              cuda_ok(cudaMalloc(&d_ab, count * 16), "cudaMalloc") && cuda_ok(cudaMalloc(&d_h, count * 64), "cudaMalloc");
    if (ok && n_pub) ok = cuda_ok(cudaMemcpy(d_pub, pub, count * n_pub * 8, cudaMemcpyHostToDevice), "H2D");
    if (ok && words) ok = cuda_ok(cudaMemcpy(d_c, containers, count * words * 8, cudaMemcpyHostToDevice), "H2D");
    ok = ok && fs_challenge_launch(d_pub, n_pub, d_c, words, count, modulus, chain, d_ab, d_h, nullptr) &&
         cuda_ok(cudaMemcpy(ab, d_ab, count * 16, cudaMemcpyDeviceToHost), "D2H") &&
         cuda_ok(cudaMemcpy(hashes, d_h, count * 64, cudaMemcpyDeviceToHost), "D2H");
    cudaFree(d_pub); cudaFree(d_c); cudaFree(d_ab); cudaFree(d_h);
    return ok;
}

bool poly_eval_host(u64 modulus, const u64* coeffs, size_t len, size_t polys, const u64* points, size_t npts, u64* out) {
    if (polys == 0 || npts == 0) return true;
    if (!cuda_ok(cudaSetDevice(current_device_choice()), "cudaSetDevice")) return false;
    u64 *d_c = nullptr, *d_p = nullptr, *d_o = nullptr;
    bool ok = cuda_ok(cudaMalloc(&d_c, std::max<size_t>(polys * len, 1) * 8), "cudaMalloc") &&
              cuda_ok(cudaMalloc(&d_p, polys * npts * 8), "cudaMalloc") && cuda_ok(cudaMalloc(&d_o, polys * npts * 8), "cudaMalloc");
    if (ok && len) ok = cuda_ok(cudaMemcpy(d_c, coeffs, polys * len * 8, cudaMemcpyHostToDevice), "H2D");
    ok = ok && cuda_ok(cudaMemcpy(d_p, points, polys * npts * 8, cudaMemcpyHostToDevice), "H2D") &&
         poly_eval_launch(modulus, d_c, len, polys, d_p, npts, polys, d_o, nullptr) &&
         cuda_ok(cudaMemcpy(out, d_o, polys * npts * 8, cudaMemcpyDeviceToHost), "D2H");
    cudaFree(d_c); cudaFree(d_p); cudaFree(d_o);
    return ok;
}

// verify_r1cs (rust-api/lambda-snark/src/lib.rs:1016-1082) for a batch of proofs of one circuit on its NTT path
// (Z_H(X) = X^m - 1, r1cs.rs:424-430): the two transcript hashes run on the device (fs_challenge_kernel), the two
// field equations per proof are a dozen modular products on the host.  results[i] = 1 accept, 0 reject.
bool verify_r1cs_host(u64 m, u64 modulus, const u64* pub, size_t n_pub, const u64* containers, size_t words,
                      const u64* challenges, const u64* evals, size_t count, int* results) {
    if (count == 0) return true;
    if (modulus < 2 || m == 0) { set_error("verify_r1cs: bad modulus / constraint count"); return false; }
    std::vector<u64> ab(count * 2), hs(count * 8);
    if (!fs_challenge_host(pub, n_pub, containers, words, count, modulus, true, ab.data(), hs.data())) return false;
    typedef unsigned __int128 u128;
    auto mul = [&](u64 a, u64 b) { return (u64)(((u128)(a % modulus) * (b % modulus)) % modulus); };
    auto sub = [&](u64 a, u64 b) { a %= modulus; b %= modulus; return a >= b ? a - b : (u64)((u128)a + modulus - b); };
    auto zh = [&](u64 x) {                                     // x^m - 1
        u64 r = 1 % modulus, base = x % modulus;
        for (u64 e = m; e; e >>= 1) { if (e & 1) r = mul(r, base); base = mul(base, base); }
        return sub(r, 1);
    };
    for (size_t i = 0; i < count; i++) {
        const u64* e = evals + 8 * i;                          // {Q(a), Q(b), A(a), B(a), C(a), A(b), B(b), C(b)}
        const u64 alpha = ab[2 * i], beta = ab[2 * i + 1];
        bool ok = challenges[2 * i] == alpha && challenges[2 * i + 1] == beta;       // lib.rs:1020-1037
        ok = ok && mul(e[0], zh(alpha)) == sub(mul(e[2], e[3]), e[4]);                // :1047-1056
        ok = ok && mul(e[1], zh(beta)) == sub(mul(e[5], e[6]), e[7]);                 // :1058-1067
        results[i] = ok ? 1 : 0;                               // openings carry Q(alpha), Q(beta) themselves (:1073-1079)
    }
    return true;
}

}  // namespace lsr
