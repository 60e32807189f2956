// lsr_quotient.cu -- SURVEY row N1: the quotient-polynomial pipeline of the prover on the GPU.
//
// Replaces, on the NTT path, rust-api/lambda-snark/src/r1cs.rs:474-503 (compute_quotient_poly):
//   compute_constraint_evals      (:296-304, SparseMatrix::mul_vec sparse_matrix.rs:259-289)  -> spmv3_kernel
//   lagrange_interpolate_ntt      (:746-793, ntt_inverse of size m over H = {omega^j})        -> inverse cyclic NTT
//   poly_mul / poly_sub           (:846-895, O(m^2) schoolbook)                               -> size-2m NTTs + pointwise
//   poly_div_vanishing(use_ntt)   (:995-1065, long division by X^m - 1)                       -> split of the 2m coefficients
// The quotient of an exact division is unique, so the result is the reference's coefficient vector bit
// for bit (canonical residues, trailing zeros trimmed by the caller-facing wrapper).
//
// Data flow for a batch of W witnesses of one R1CS instance (m constraints, m a power of two).  The
// reference multiplies A_z and B_z as coefficient vectors (degree 2m - 2) and long-divides by X^m - 1.
// Here the division is done where it is free: on the coset psi * H of the m-th roots of unity H (psi any
// primitive 2m-th root), X^m - 1 is the constant psi^m - 1 = -2, and deg Q <= m - 2, so m coset values
// determine Q.  Evaluating at psi^(2j+1) is exactly the NEGACYCLIC transform of size m with root psi:
//   E[3][W][m]   evaluations of A_z, B_z, C_z on H, written by the mat-vec directly in bit-reversed row
//                order (the order the inverse transform consumes); the same kernel checks
//                a_i * b_i = c_i (is_satisfied, r1cs.rs:477-481  <=>  zero remainder, :1054-1060)
//   -> inverse cyclic transform (size m, batch 3W)      coefficients of A_z, B_z, C_z
//   -> forward negacyclic transform (size m, batch 2W)  values of A_z, B_z on psi * H (bit-reversed order)
//   N^[W][m]     a * b pointwise
//   -> inverse negacyclic transform (size m, batch W)   N = A_z * B_z mod (X^m + 1), natural order
//   Q[W][m]      (C_z - N) / 2 pointwise on coefficients: the transforms are linear, so the coset values of C_z
//                ((a * b - c) * (-2)^-1 on the coset) are never needed; equivalently, with
//                A_z * B_z = P_lo + X^m * P_hi:  C_z = P_lo + P_hi (it agrees with the product on H),
//                N = P_lo - P_hi, and Q = P_hi
// 6 W transforms of size m and no zero padding, against 3 W of size m + 4 W of size 2m.
#include <algorithm>
#include <cstdlib>
#include <new>
#include <utility>

#include "lsr_arith.cuh"
#include "lsr_engine.h"
#include "lsr_r1cs.h"

namespace lsr {

bool fs_challenge_launch(const u64* d_pub, size_t n_pub, const u64* d_containers, size_t words, size_t count,
                         u64 modulus, bool chain, u64* d_ab, u64* d_hashes, cudaStream_t s);
bool poly_eval_launch(u64 modulus, const u64* d_coeffs, size_t len, size_t polys, const u64* d_points, size_t npts,
                      size_t point_rows, u64* d_out, cudaStream_t s);

struct DeviceCsr {
    uint32_t* row_ptr = nullptr;   // [3][rows + 1], offsets into col / val of the concatenated A|B|C entries
    uint32_t* col = nullptr;
    u64* val = nullptr;            // reduced mod q (sparse_matrix.rs:279 `val % modulus`)
};

struct QuotientState {
    NttContext* small = nullptr;   // cyclic, size m (interpolation on H)
    NttContext* coset = nullptr;   // negacyclic, size m (evaluation / interpolation on psi * H)
    DeviceCsr csr;
    DeviceScratch z, e, qbuf, flags, seeds, containers, coef, fs;
    PinnedScratch h_flags;
    u64 omega = 0;
    int device = 0;
    // host-io pipeline of prover_commit_quotient: copy-in / compute / copy-out streams, two buffers each
    cudaStream_t pipe[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr},
                ev_zfree[2] = {nullptr, nullptr};
    bool pipe_ready = false;
};

void quotient_state_free(QuotientState* s) {
    if (!s) return;
    cudaSetDevice(s->device);
    if (s->small) ntt_destroy(s->small);
    if (s->coset) ntt_destroy(s->coset);
    if (s->csr.row_ptr) cudaFree(s->csr.row_ptr);
    if (s->csr.col) cudaFree(s->csr.col);
    if (s->csr.val) cudaFree(s->csr.val);
    s->z.release(); s->e.release(); s->qbuf.release(); s->flags.release(); s->h_flags.release();
    s->seeds.release(); s->containers.release(); s->coef.release(); s->fs.release();
    for (auto& st : s->pipe) if (st) cudaStreamDestroy(st);
    for (int b = 0; b < 2; b++)
        for (cudaEvent_t e : {s->ev_in[b], s->ev_comp[b], s->ev_out[b], s->ev_zfree[b]}) if (e) cudaEventDestroy(e);
    delete s;
}

// ------------------------------------------------------------------ kernels
// one thread per (witness, row); the three matrices share the witness loads.  flags[w] |= (a * b != c).
__global__ void __launch_bounds__(256)
spmv3_kernel(const ModParams mp, const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
             const u64* __restrict__ val, const u64* __restrict__ z, uint32_t rows, uint32_t cols, int logm,
             size_t witnesses, u64* __restrict__ E, unsigned* __restrict__ flags) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= witnesses * rows) return;
    const size_t w = idx / rows;
    const uint32_t r = (uint32_t)(idx % rows);
    const u64* __restrict__ zw = z + w * cols;
    const uint32_t dst = logm ? (__brev(r) >> (32 - logm)) : 0u;
    u64 abc[3];
#pragma unroll
    for (int mat = 0; mat < 3; mat++) {
        const uint32_t* rp = row_ptr + (size_t)mat * (rows + 1);
        u64 acc = 0;
        for (uint32_t k = rp[r]; k < rp[r + 1]; k++)
            acc = field_add(acc, field_mul(val[k], reduce64(zw[col[k]], mp), mp), mp);     // v[col] % modulus
        E[((size_t)mat * witnesses + w) * rows + dst] = acc;
        abc[mat] = acc;
    }
    if (field_mul(abc[0], abc[1], mp) != abc[2]) atomicOr(flags + w, 1u);
}

// Same mat-vec for m >= 1024 with both sides coalesced: a CTA owns the 1024 rows r = (a, mid, b) with a fixed
// middle field (a = top 5 bits, b = low 5 bits); warp a reads 32 consecutive rows, the values cross a padded
// shared-memory tile, and warp b writes the 32 consecutive destinations brv(r) = (brv5(b), brv(mid), brv5(a)),
// a = 0..31 -- 256-byte segments instead of 32 scattered 8-byte words.  grid = (m / 1024, witnesses)
__global__ void __launch_bounds__(256, 4)
spmv3_tiled_kernel(const ModParams mp, const uint32_t* __restrict__ row_ptr, const uint32_t* __restrict__ col,
                   const u64* __restrict__ val, const u64* __restrict__ z, uint32_t rows, uint32_t cols, int logm,
                   size_t witnesses, u64* __restrict__ E, unsigned* __restrict__ flags) {
    __shared__ u64 tile[3][32][33];
    const uint32_t b = threadIdx.x, mid = blockIdx.x;
    const size_t w = blockIdx.y;
    const int midbits = logm - 10;
    const u64* __restrict__ zw = z + w * cols;
    bool bad = false;
    // 256 threads, four rows each: a = threadIdx.y + 8 i.  The kernel is bound by load latency, not by bytes (ncu: 3 TB/s at
    // 24 % of the warp slots, long-scoreboard stalls 6.6 per issue with the entry loop of every (row, matrix) pair as a
    // chain of three dependent loads at 96 registers).  So the FIRST entry of all twelve pairs goes as waves of independent
    // loads -- row pointers, then column indices and values, then the witness words -- and only the first products stay
    // live (64 registers, four CTAs per SM); rows with further entries (gate-style circuits have one or two per row) finish
    // in the loop below, which fetches its row pointers again (L1).
    u64 acc[4][3];
    bool more = false;
    {
        uint32_t c0[4][3];
        u64 v0[4][3];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const uint32_t r = ((threadIdx.y + 8u * i) << (logm - 5)) | (mid << 5) | b;
#pragma unroll
            for (int mat = 0; mat < 3; mat++) {
                const uint32_t* rp = row_ptr + (size_t)mat * (rows + 1);
                const uint32_t lo = __ldg(rp + r), hi = __ldg(rp + r + 1);
                const bool has = lo < hi;
                more |= hi - lo > 1u;
                c0[i][mat] = has ? __ldg(col + lo) : 0u;
                v0[i][mat] = has ? __ldg(val + lo) : 0ull;          // weight 0: column 0 (it exists) contributes nothing
            }
        }
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int mat = 0; mat < 3; mat++) acc[i][mat] = __ldg(zw + c0[i][mat]);
#pragma unroll
        for (int i = 0; i < 4; i++)
#pragma unroll
            for (int mat = 0; mat < 3; mat++) acc[i][mat] = field_mul(v0[i][mat], reduce64(acc[i][mat], mp), mp);
    }
    if (more) {
#pragma unroll 1
        for (int i = 0; i < 4; i++) {
            const uint32_t r = ((threadIdx.y + 8u * i) << (logm - 5)) | (mid << 5) | b;
#pragma unroll 1
            for (int mat = 0; mat < 3; mat++) {
                const uint32_t* rp = row_ptr + (size_t)mat * (rows + 1);
                const uint32_t hi = __ldg(rp + r + 1);
                u64 s = 0;
                for (uint32_t k = __ldg(rp + r) + 1u; k < hi; k++)
                    s = field_add(s, field_mul(__ldg(val + k), reduce64(__ldg(zw + __ldg(col + k)), mp), mp), mp);
                // acc[i][mat] with run-time indices would send the array to local memory: select instead
#pragma unroll
                for (int ii = 0; ii < 4; ii++)
#pragma unroll
                    for (int mm = 0; mm < 3; mm++)
                        if (ii == i && mm == mat) acc[ii][mm] = field_add(acc[ii][mm], s, mp);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t a = threadIdx.y + 8u * i;
#pragma unroll
        for (int mat = 0; mat < 3; mat++) tile[mat][a][b] = acc[i][mat];
        bad |= field_mul(acc[i][0], acc[i][1], mp) != acc[i][2];
    }
    if (bad) atomicOr(flags + w, 1u);
    __syncthreads();
    // now threadIdx.y (+ 8 i) plays b, threadIdx.x plays a
    const uint32_t ra = __brev(threadIdx.x) >> 27;
    const uint32_t rmid = midbits ? (__brev(mid) >> (32 - midbits)) : 0u;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const uint32_t bb = threadIdx.y + 8u * i;
        const uint32_t dst = ((__brev(bb) >> 27) << (logm - 5)) | (rmid << 5) | ra;
#pragma unroll
        for (int mat = 0; mat < 3; mat++)
            E[((size_t)mat * witnesses + w) * rows + dst] = tile[mat][threadIdx.x][bb];
    }
}

// N^ = A * B on the coset values (any order: pointwise).  C takes no part on the coset: the transforms are
// linear, so its share of the quotient is subtracted in the coefficient domain (quotient_finish_kernel) and
// its forward transform is never computed.
__global__ void __launch_bounds__(256)
coset_product_kernel(const ModParams mp, const u64* __restrict__ E, u64* __restrict__ Nh, size_t per_matrix) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per_matrix) return;
    Nh[idx] = field_mul(E[idx], E[per_matrix + idx], mp);
}

// Q = (C - N) / 2, in place over N.  N = A_z * B_z mod (X^m + 1) (the inverse negacyclic transform of N^),
// C = A_z * B_z mod (X^m - 1) for a satisfied witness; with A_z * B_z = P_lo + X^m * P_hi that is
// C - N = 2 * P_hi = 2 * Q.  For any witness it equals the inverse transform of (A * B - C) * (-2)^-1 on the coset.
__global__ void __launch_bounds__(256)
quotient_finish_kernel(const ModParams mp, const u64* __restrict__ C, u64* __restrict__ Q, size_t per_matrix, u64 inv2) {
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= per_matrix) return;
    Q[idx] = field_mul(field_sub(C[idx], Q[idx], mp), inv2, mp);
}

// Test hook: the Goldilocks primitives on caller-supplied operands (any 64-bit words), so that the carry / borrow
// corner cases (operands and intermediate sums at 0, q - 1, q, 2^64 - 1, 2^32 boundaries) can be compared with
// arbitrary-precision arithmetic.  out[i] = {a*b, canonical(a +lazy (b mod q)), canonical(a - (b mod q)), (a mod q) + (b mod q)}
__global__ void gold_probe_kernel(const u64* __restrict__ a, const u64* __restrict__ b, size_t count, u64* __restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const u64 x = a[i], y = b[i];
    const u64 yc = gold_canonical(y), xc = gold_canonical(x);     // y < 2^64 < 2q: one conditional subtraction
    out[4 * i + 0] = gold_mul(x, y);
    out[4 * i + 1] = gold_canonical(gold_add_lazy(x, yc));
    out[4 * i + 2] = gold_canonical(gold_sub(x, yc));
    out[4 * i + 3] = gold_add(xc, yc);
}

bool gold_probe_host(const u64* a, const u64* b, size_t count, u64* out) {
    if (count == 0) return true;
    if (!cuda_ok(cudaSetDevice(current_device_choice()), "cudaSetDevice")) return false;
    u64 *da = nullptr, *db = nullptr, *dout = nullptr;
    bool ok = cuda_ok(cudaMalloc(&da, count * 8), "cudaMalloc") && cuda_ok(cudaMalloc(&db, count * 8), "cudaMalloc") &&
              cuda_ok(cudaMalloc(&dout, count * 32), "cudaMalloc") &&
              cuda_ok(cudaMemcpy(da, a, count * 8, cudaMemcpyHostToDevice), "H2D") &&
              cuda_ok(cudaMemcpy(db, b, count * 8, cudaMemcpyHostToDevice), "H2D");
    if (ok) {
        gold_probe_kernel<<<(unsigned)((count + 255) / 256), 256>>>(da, db, count, dout);
        ok = cuda_ok(cudaGetLastError(), "gold_probe_kernel") && cuda_ok(cudaMemcpy(out, dout, count * 32, cudaMemcpyDeviceToHost), "D2H");
    }
    cudaFree(da); cudaFree(db); cudaFree(dout);
    return ok;
}

// ------------------------------------------------------------------ host side
static bool build_csr(const R1csHandle* h, QuotientState* st) {
    const uint32_t rows = h->rows;
    const std::vector<SparseEntry>* mats[3] = {&h->A, &h->B, &h->C};
    size_t total = 0;
    for (auto* m : mats) total += m->size();
    if (total > 0xfffffff0ull) { set_error("R1CS too large"); return false; }
    std::vector<uint32_t> row_ptr((size_t)3 * (rows + 1), 0), col(std::max<size_t>(total, 1));
    std::vector<u64> val(std::max<size_t>(total, 1));
    size_t base = 0;
    for (int k = 0; k < 3; k++) {
        uint32_t* rp = row_ptr.data() + (size_t)k * (rows + 1);
        std::vector<uint32_t> count(rows + 1, 0);
        for (const SparseEntry& e : *mats[k]) {
            if (e.row >= rows || e.col >= h->cols) { set_error("R1CS entry out of range"); return false; }
            count[e.row + 1]++;
        }
        for (uint32_t r = 0; r < rows; r++) count[r + 1] += count[r];
        for (uint32_t r = 0; r <= rows; r++) rp[r] = (uint32_t)(base + count[r]);
        std::vector<uint32_t> fill(count.begin(), count.end() - 1);
        for (const SparseEntry& e : *mats[k]) {                 // stable: entries of a row keep their order
            const size_t at = base + fill[e.row]++;
            col[at] = e.col;
            val[at] = e.value % h->q;
        }
        base += mats[k]->size();
    }
    bool ok = cuda_ok(cudaMalloc(&st->csr.row_ptr, row_ptr.size() * 4), "cudaMalloc(csr)") &&
              cuda_ok(cudaMalloc(&st->csr.col, col.size() * 4), "cudaMalloc(csr)") &&
              cuda_ok(cudaMalloc(&st->csr.val, val.size() * 8), "cudaMalloc(csr)") &&
              cuda_ok(cudaMemcpy(st->csr.row_ptr, row_ptr.data(), row_ptr.size() * 4, cudaMemcpyHostToDevice), "H2D csr") &&
              cuda_ok(cudaMemcpy(st->csr.col, col.data(), col.size() * 4, cudaMemcpyHostToDevice), "H2D csr") &&
              cuda_ok(cudaMemcpy(st->csr.val, val.data(), val.size() * 8, cudaMemcpyHostToDevice), "H2D csr");
    return ok;
}

// The reference's roots: NTT_PRIMITIVE_ROOT^(2^32 / n) for Goldilocks (lambda-snark-core/src/lib.rs:78,
// ntt.rs:214-221 compute_root_of_unity); 3^((q-1)/n) for 17592169062401 (r1cs.rs:534-547 ROOTS_OF_UNITY);
// the minimal primitive n-th root otherwise.
u64 reference_root_of_unity(u64 q, uint32_t n) {
    if (n < 2 || (n & (n - 1)) || (q - 1) % n) return 0;
    if (q == kGoldilocks) return host::powmod(1753635133440165772ULL, (1ull << 32) / n, q);
    if (q == kDefaultModulusSmall) return host::powmod(3, (q - 1) / n, q);
    return host::min_primitive_root(q, n);
}

static QuotientState* get_state(R1csHandle* h, u64 omega) {
    if (h->quotient && h->quotient->omega == omega) return h->quotient;
    if (h->quotient) { quotient_state_free(h->quotient); h->quotient = nullptr; }
    const uint32_t m = h->rows;
    QuotientState* st = new (std::nothrow) QuotientState;
    if (!st) return nullptr;
    st->device = current_device_choice();
    st->omega = omega;
    bool ok = cuda_ok(cudaSetDevice(st->device), "cudaSetDevice");
    if (ok && m >= 2) {
        st->small = ntt_create_cyclic(h->q, m, omega);
        // any primitive 2m-th root serves as the coset shift: psi * H does not depend on the generator of H
        u64 psi = reference_root_of_unity(h->q, 2 * m);
        st->coset = psi ? ntt_create_negacyclic(h->q, m, psi) : nullptr;
        ok = st->small && st->coset;
        if (!ok) set_error("quotient: the modulus has no NTT of size 2m (need 2m | q-1, q prime < 2^61 or Goldilocks)");
    }
    ok = ok && build_csr(h, st);
    if (!ok) { quotient_state_free(st); return nullptr; }
    h->quotient = st;
    return st;
}

// Witnesses dz [W][cols] on the device -> evaluations / coefficients in dE [3][W][m], Q in dQ [W][m] (zero-padded),
// dF[w] != 0 when witness w does not satisfy the constraints.  Asynchronous on s.
static bool quotient_compute(R1csHandle* h, QuotientState* st, const u64* dz, size_t W, u64* dE, u64* dQ, unsigned* dF,
                             bool keep_coeffs, cudaStream_t s) {
    const uint32_t m = h->rows, cols = h->cols;
    const ModParams mp = host::make_mod_params(h->q, 1);
    const size_t em = (size_t)3 * W * m;
    int logm = 0;
    while ((1u << logm) < m) ++logm;
    auto grid = [](size_t n) { return (unsigned)((n + 255) / 256); };
    bool ok = cuda_ok(cudaMemsetAsync(dF, 0, W * 4, s), "memset");
    if (ok) {
        if (logm >= 10 && W <= 65535)
            spmv3_tiled_kernel<<<dim3(m >> 10, (unsigned)W), dim3(32, 8), 0, s>>>(mp, st->csr.row_ptr, st->csr.col, st->csr.val,
                                                                                  dz, m, cols, logm, W, dE, dF);
        else
            spmv3_kernel<<<grid(W * m), 256, 0, s>>>(mp, st->csr.row_ptr, st->csr.col, st->csr.val, dz, m, cols, logm, W, dE, dF);
        ok = cuda_ok(cudaGetLastError(), "spmv3_kernel");
    }
    if (ok && m >= 2) {
        const u64 inv2 = h->q / 2 + 1;                         // 2 * (q + 1) / 2 = q + 1 = 1 (mod q), q odd
        ok = ntt_inverse_launch(st->small, dE, 3 * W, s);
        if (ok && keep_coeffs)      // A_z, B_z, C_z as coefficient vectors [3][W][m] for the evaluations of prove_r1cs
            ok = st->coef.reserve(em * 8) &&
                 cuda_ok(cudaMemcpyAsync(st->coef.ptr, dE, em * 8, cudaMemcpyDeviceToDevice, s), "D2D coefficients");
        ok = ok && ntt_forward_launch(st->coset, dE, 2 * W, s);            // A_z, B_z only (C_z stays as coefficients)
#ifdef LSR_PROFILING    // tools/ build only: LSR_QUOT_FUSE bit 0 = product fused, bit 1 = finishing step fused (default 3)
        static const int fuse = [] { const char* e = std::getenv("LSR_QUOT_FUSE"); return e ? std::atoi(e) : 3; }();
#else
        constexpr int fuse = 3;
#endif
        if (ok && fuse && ntt_inverse_fused_supported(st->coset)) {
            // one inverse transform with both neighbours fused in: its first kernel reads a * b (the coset product is
            // never written) and writes to Q, its last kernel stores (C - N) / 2 -- two 24 B/coefficient sweeps less
            if (!(fuse & 1)) {
                coset_product_kernel<<<grid(W * m), 256, 0, s>>>(mp, dE, dQ, W * (size_t)m);
                ok = cuda_ok(cudaGetLastError(), "coset_product_kernel");
            }
            const InvFusion fz{(fuse & 1) ? dE + W * (size_t)m : nullptr, (fuse & 1) ? dQ : nullptr,
                               (fuse & 2) ? dE + 2 * W * (size_t)m : nullptr, inv2};
            ok = ok && ntt_inverse_fused_launch(st->coset, (fuse & 1) ? dE : dQ, W, s, fz);
            if (ok && !(fuse & 2)) {
                quotient_finish_kernel<<<grid(W * m), 256, 0, s>>>(mp, dE + 2 * W * (size_t)m, dQ, W * (size_t)m, inv2);
                ok = cuda_ok(cudaGetLastError(), "quotient_finish_kernel");
            }
        } else if (ok) {
            coset_product_kernel<<<grid(W * m), 256, 0, s>>>(mp, dE, dQ, W * (size_t)m);
            ok = cuda_ok(cudaGetLastError(), "coset_product_kernel");
            ok = ok && ntt_inverse_launch(st->coset, dQ, W, s);
            if (ok) {
                quotient_finish_kernel<<<grid(W * m), 256, 0, s>>>(mp, dE + 2 * W * (size_t)m, dQ, W * (size_t)m, inv2);
                ok = cuda_ok(cudaGetLastError(), "quotient_finish_kernel");
            }
        }
    } else if (ok) {
        // m = 1: A_z, B_z, C_z are constants; the numerator a*b - c has degree 0 < deg(X - 1), so the quotient is 0
        // and the division is exact iff the numerator vanishes (r1cs.rs:1010-1020): the flag of the mat-vec
        ok = cuda_ok(cudaMemsetAsync(dQ, 0, W * 8, s), "memset");
        if (ok && keep_coeffs)      // degree-0 interpolants: the evaluations themselves
            ok = st->coef.reserve(em * 8) &&
                 cuda_ok(cudaMemcpyAsync(st->coef.ptr, dE, em * 8, cudaMemcpyDeviceToDevice, s), "D2D coefficients");
    }
    return ok;
}

// Device part of the pipeline: witnesses [count][cols] (HOST words, or DEVICE words when witnesses_on_device)
// -> Q [count][m] zero-padded, left in st->qbuf on the device, status[count] on the host (0 ok, 1 the witness
// does not satisfy the constraints).  The caller holds h->mu.  Synchronises the stream before returning.
static int quotient_prepare(R1csHandle* h, u64 omega, QuotientState** out_state) {
    const uint32_t m = h->rows, cols = h->cols;
    if (m == 0 || (m & (m - 1)) || m > (1u << (kMaxEngineLogN - 1)) || cols == 0) { set_error("quotient: m must be a power of two <= 2^23"); return 2; }
    if (h->q != kGoldilocks && (h->q >> 61)) { set_error("quotient: unsupported modulus"); return 2; }
    if (m >= 2 && omega == 0) omega = reference_root_of_unity(h->q, m);
    if (m >= 2 && !host::cyclic_params_ok(h->q, m, omega)) { set_error("quotient: omega is not a primitive m-th root of unity"); return 2; }
    QuotientState* st = get_state(h, m >= 2 ? omega : 0);
    if (!st) return 4;
    *out_state = st;
    return cuda_ok(cudaSetDevice(st->device), "cudaSetDevice") ? 0 : 4;
}

static int quotient_device(R1csHandle* h, const u64* witnesses, bool witnesses_on_device, size_t count, u64 omega,
                           int* status, QuotientState** out_state, bool keep_coeffs = false) {
    const int rc = quotient_prepare(h, omega, out_state);
    if (rc != 0 || count == 0) return rc;
    QuotientState* st = *out_state;
    const uint32_t m = h->rows, cols = h->cols;
    cudaStream_t s = nullptr;      // legacy default stream: the cyclic contexts are private to this handle
    const size_t W = count;
    bool ok = st->z.reserve(W * cols * 8) && st->e.reserve((size_t)3 * W * m * 8) &&
              st->qbuf.reserve(W * m * 8) && st->flags.reserve(W * 4) && st->h_flags.reserve(W * 4);
    if (!ok) return 3;
    u64* dz = static_cast<u64*>(st->z.ptr);
    u64* dE = static_cast<u64*>(st->e.ptr);
    u64* dQ = static_cast<u64*>(st->qbuf.ptr);
    unsigned* dF = static_cast<unsigned*>(st->flags.ptr);
    ok = cuda_ok(cudaMemcpyAsync(dz, witnesses, W * cols * 8,
                                 witnesses_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s), "H2D witness") &&
         quotient_compute(h, st, dz, W, dE, dQ, dF, keep_coeffs, s);
    ok = ok && cuda_ok(cudaMemcpyAsync(st->h_flags.ptr, dF, W * 4, cudaMemcpyDeviceToHost, s), "D2H flags") &&
         cuda_ok(cudaStreamSynchronize(s), "sync");
    if (!ok) return 4;
    const unsigned* hf = static_cast<const unsigned*>(st->h_flags.ptr);
    for (size_t w = 0; w < W; w++) status[w] = hf[w] ? 1 : 0;
    return 0;
}

// witnesses: [count][cols] host words; out: [count][m] host words (Q zero-padded to m); status[count]: 0 ok,
// 1 the witness does not satisfy the constraints (non-zero remainder)
int r1cs_quotient_batch(R1csHandle* h, const u64* witnesses, size_t count, u64 omega, u64* out, int* status) {
    std::lock_guard<std::mutex> lock(h->mu);
    QuotientState* st = nullptr;
    const int rc = quotient_device(h, witnesses, false, count, omega, status, &st);
    if (rc != 0 || count == 0) return rc;
    return cuda_ok(cudaMemcpy(out, st->qbuf.ptr, count * (size_t)h->rows * 8, cudaMemcpyDeviceToHost), "D2H quotient") ? 0 : 4;
}

// Host-io pipeline of the commitment phase: witnesses are taken in groups; while group g is computed, the witnesses of
// group g + 1 come in and the containers of group g - 1 go out (three streams, two buffers of each kind).  With
// page-locked caller buffers the call is bound by the slower PCIe direction instead of the sum of both copies and the
// compute.  Groups hold at least ~8 MB of witness words so that the copies run at full link rate.
static std::pair<size_t, size_t> pipeline_groups(const R1csHandle* h, size_t count) {
    const size_t wbytes = (size_t)h->cols * 8;
    const size_t G = std::max<size_t>(1, std::min<size_t>(count, ((size_t)8 << 20) / std::max<size_t>(wbytes, 1)));
    return {G, G ? (count + G - 1) / G : 0};
}

static bool pipeline_init(QuotientState* st) {
    if (st->pipe_ready) return true;
    bool ok = true;
    for (auto& s : st->pipe) ok = ok && cuda_ok(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking), "cudaStreamCreate");
    for (int b = 0; b < 2 && ok; b++)
        for (cudaEvent_t* e : {&st->ev_in[b], &st->ev_comp[b], &st->ev_out[b], &st->ev_zfree[b]})
            ok = ok && cuda_ok(cudaEventCreateWithFlags(e, cudaEventDisableTiming), "cudaEventCreate");
    st->pipe_ready = ok;
    return ok;
}

static int prover_commit_pipelined(R1csHandle* h, const LweContext* lwe, const u64* witnesses, size_t count, u64 omega,
                                   const u64* seeds, size_t chunk_lo, size_t chunk_hi, uint32_t planes, u64* out, int* status) {
    QuotientState* st = nullptr;
    const int rc = quotient_prepare(h, omega, &st);
    if (rc != 0) return rc;
    if (st->device != lwe->device) { set_error("prover_commit_quotient: R1CS and LWE context live on different devices"); return 2; }
    if (!pipeline_init(st)) return 4;
    const uint32_t m = h->rows, cols = h->cols, n = lwe->n;
    // units = (ring-element chunk, digit plane) pairs; [chunk_lo, chunk_hi) ranges over units
    const size_t chunks = (size_t)m / n * planes, mine = chunk_hi - chunk_lo, words = lwe_words(lwe);
    const auto [G, NG] = pipeline_groups(h, count);
    bool ok = st->z.reserve(2 * G * cols * 8) && st->e.reserve((size_t)3 * G * m * 8) && st->qbuf.reserve(G * (size_t)m * 8) &&
              st->flags.reserve(count * 4) && st->h_flags.reserve(count * 4) && st->seeds.reserve(2 * G * mine * 8) &&
              st->containers.reserve(2 * G * mine * words * 8);
    if (!ok) return 3;
    u64* dE = static_cast<u64*>(st->e.ptr);
    u64* dQ = static_cast<u64*>(st->qbuf.ptr);
    unsigned* dF = static_cast<unsigned*>(st->flags.ptr);
    cudaStream_t sin = st->pipe[0], scomp = st->pipe[1], sout = st->pipe[2];
    for (size_t g = 0; ok && g < NG; g++) {
        const int b = (int)(g & 1);
        const size_t w0 = g * G, gw = std::min(G, count - w0);
        u64* dz = static_cast<u64*>(st->z.ptr) + (size_t)b * G * cols;
        u64* dseed = static_cast<u64*>(st->seeds.ptr) + (size_t)b * G * mine;
        u64* dcont = static_cast<u64*>(st->containers.ptr) + (size_t)b * G * mine * words;
        // copy-in: the buffers of group g - 2 must have been consumed
        if (g >= 2) ok = cuda_ok(cudaStreamWaitEvent(sin, st->ev_zfree[b], 0), "wait");
        ok = ok && cuda_ok(cudaMemcpyAsync(dz, witnesses + w0 * cols, gw * cols * 8, cudaMemcpyHostToDevice, sin), "H2D witness");
        // seeds of the group: rows of `chunks` words, of which [chunk_lo, chunk_hi) are this call's -- one strided copy
        ok = ok && cuda_ok(cudaMemcpy2DAsync(dseed, mine * 8, seeds + w0 * chunks + chunk_lo, chunks * 8, mine * 8, gw,
                                             cudaMemcpyHostToDevice, sin), "H2D seeds");
        ok = ok && cuda_ok(cudaEventRecord(st->ev_in[b], sin), "record");
        // compute: quotient of the group, then its commitments into container buffer b (free once group g - 2 is out)
        ok = ok && cuda_ok(cudaStreamWaitEvent(scomp, st->ev_in[b], 0), "wait");
        if (ok && g >= 2) ok = cuda_ok(cudaStreamWaitEvent(scomp, st->ev_out[b], 0), "wait");
        ok = ok && quotient_compute(h, st, dz, gw, dE, dQ, dF + w0, false, scomp);
        if (ok && mine == chunks) {
            ok = lwe_commit_launch(lwe, dQ, n, dseed, gw * chunks, dcont, scomp, planes, 0);
        } else for (size_t w = 0; ok && w < gw; w++) {
            ok = lwe_commit_launch(lwe, dQ + w * (size_t)m, n, dseed + w * mine, mine, dcont + w * mine * words, scomp,
                                   planes, chunk_lo);
        }
        // the seeds live in the same double buffer as the witnesses: both are free after the commitments
        ok = ok && cuda_ok(cudaEventRecord(st->ev_zfree[b], scomp), "record") &&
             cuda_ok(cudaEventRecord(st->ev_comp[b], scomp), "record");
        // copy-out
        ok = ok && cuda_ok(cudaStreamWaitEvent(sout, st->ev_comp[b], 0), "wait") &&
             cuda_ok(cudaMemcpyAsync(out + w0 * mine * words, dcont, gw * mine * words * 8, cudaMemcpyDeviceToHost, sout), "D2H containers") &&
             cuda_ok(cudaEventRecord(st->ev_out[b], sout), "record");
    }
    ok = ok && cuda_ok(cudaMemcpyAsync(st->h_flags.ptr, dF, count * 4, cudaMemcpyDeviceToHost, scomp), "D2H flags");
    // drain all three streams even after a failure: the buffers must be quiescent when the call returns
    for (cudaStream_t s : {sin, scomp, sout}) ok = cuda_ok(cudaStreamSynchronize(s), "sync") && ok;
    if (!ok) return 4;
    const unsigned* hf = static_cast<const unsigned*>(st->h_flags.ptr);
    for (size_t w = 0; w < count; w++) status[w] = hf[w] ? 1 : 0;
    return 0;
}

// Commitment phase of the prover (BASELINE configs[4]; replaces compute_quotient_poly + Commitment::new of
// prove_r1cs, rust-api/lambda-snark/src/lib.rs:747-757, for quotients longer than one ring element, which the
// reference silently truncates -- SURVEY F6).  Q of every witness is cut into chunks = ceil(m / n) messages of
// n = ring_degree coefficients (the last one shorter when m < n); message (w, j) is committed with seed
// seeds[w * chunks + j]; only chunk indices [chunk_lo, chunk_hi) are committed (a rank's slice of a sharded
// job).  Nothing but the witness and the containers crosses PCIe.
// out: [count][chunk_hi - chunk_lo][1 + k n] host words (or device words when out_on_device).
int prover_commit_quotient(R1csHandle* h, const LweContext* lwe, const u64* witnesses, size_t count, u64 omega,
                           const u64* seeds, size_t chunk_lo, size_t chunk_hi, u64* out, bool io_on_device,
                           int* status) {
    std::lock_guard<std::mutex> lock(h->mu);
    const uint32_t m = h->rows, n = lwe->n;
    // Digit planes: a commitment binds its message words modulo p, so every ring-element chunk of Q is
    // committed as L = ceil(log_p(field modulus)) commitments to its base-p digits; together they bind the whole quotient.
    const uint32_t planes = message_planes(lwe->p, h->q);
    if (planes == 0) { set_error("prover_commit_quotient: more than 4 digit planes needed (plaintext modulus too small)"); return 2; }
    const size_t chunks = (m <= n ? 1 : (size_t)m / n) * planes;       // units: (chunk, plane) pairs
    if (chunk_lo > chunk_hi || chunk_hi > chunks) { set_error("prover_commit_quotient: bad chunk range"); return 2; }
    QuotientState* st = nullptr;
    const size_t mine = chunk_hi - chunk_lo;
    if (!io_on_device && mine > 0 && lwe->n <= m && pipeline_groups(h, count).second >= 2)
        return prover_commit_pipelined(h, lwe, witnesses, count, omega, seeds, chunk_lo, chunk_hi, planes, out, status);
    const int rc = quotient_device(h, witnesses, io_on_device, count, omega, status, &st);
    if (rc != 0) return rc;
    if (count == 0 || mine == 0) return 0;
    if (st->device != lwe->device) { set_error("prover_commit_quotient: R1CS and LWE context live on different devices"); return 2; }
    const size_t words = lwe_words(lwe);
    const size_t msg_len = std::min<uint32_t>(m, n);
    cudaStream_t s = nullptr;
    if (!st->seeds.reserve(count * mine * 8)) return 3;
    u64* d_out = out;
    if (!io_on_device) {
        if (!st->containers.reserve(count * mine * words * 8)) return 3;
        d_out = static_cast<u64*>(st->containers.ptr);
    }
    u64* d_seeds = static_cast<u64*>(st->seeds.ptr);
    const u64* dQ = static_cast<const u64*>(st->qbuf.ptr);
    bool ok = true;
    if (mine == chunks) {
        // the whole quotient of every witness: Q [count][m] is [count * chunks][msg_len] -- one launch
        ok = cuda_ok(cudaMemcpyAsync(d_seeds, seeds, count * chunks * 8,
                                     io_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s), "H2D seeds") &&
             lwe_commit_launch(lwe, dQ, msg_len, d_seeds, count * chunks, d_out, s, planes, 0);
    } else {
        ok = cuda_ok(cudaMemcpy2DAsync(d_seeds, mine * 8, seeds + chunk_lo, chunks * 8, mine * 8, count,
                                       io_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, s), "H2D seeds");
        for (size_t w = 0; ok && w < count; w++)
            ok = lwe_commit_launch(lwe, dQ + w * (size_t)m, msg_len, d_seeds + w * mine, mine, d_out + w * mine * words, s,
                                   planes, chunk_lo);
    }
    if (ok && !io_on_device)
        ok = cuda_ok(cudaMemcpyAsync(out, d_out, count * mine * words * 8, cudaMemcpyDeviceToHost, s), "D2H containers");
    ok = ok && cuda_ok(cudaStreamSynchronize(s), "sync");
    return ok ? 0 : 4;
}

// prove_r1cs for a batch of witnesses of one circuit (rust-api/lambda-snark/src/lib.rs:747-809), on the path
// where the reference is self-consistent: interpolation over the roots of unity (NTT modulus), quotient no
// longer than one ring element (m <= ring_degree; longer quotients are truncated by the reference, F6).
// Per witness w:  Q -> commitment (seed seeds[w]) -> alpha = FS(z[0..n_public), C), beta = FS([alpha], C)
// -> evals[w] = {Q(alpha), Q(beta), A_z(alpha), B_z(alpha), C_z(alpha), A_z(beta), B_z(beta), C_z(beta)}
// (the field order of ProofR1CS::new, lib.rs:795-809).  The commitment never visits the host before its
// challenges exist; containers, challenges [count][2], hashes [count][2][4], evals [count][8] are copied out at
// the end (HOST pointers).
int prove_r1cs_batch(R1csHandle* h, const LweContext* lwe, const u64* witnesses, size_t count, size_t n_public,
                     u64 omega, const u64* seeds, u64* containers, u64* challenges, u64* hashes, u64* evals, int* status) {
    std::lock_guard<std::mutex> lock(h->mu);
    const uint32_t m = h->rows, cols = h->cols;
    if (m > lwe->n) { set_error("prove_r1cs_batch: quotient longer than one ring element (use lsr_prover_commit_quotient)"); return 2; }
    if (n_public > cols) { set_error("prove_r1cs_batch: more public inputs than variables"); return 2; }
    QuotientState* st = nullptr;
    const int rc = quotient_device(h, witnesses, false, count, omega, status, &st, true);
    if (rc != 0 || count == 0) return rc;
    if (st->device != lwe->device) { set_error("prove_r1cs_batch: R1CS and LWE context live on different devices"); return 2; }
    const size_t words = lwe_words(lwe), W = count;
    cudaStream_t s = nullptr;
    // fs scratch: pub [W][n_public] | ab [W][2] | hashes [W][8] | evals: Q [W][2], ABC [3][W][2]
    const size_t pub_w = W * std::max<size_t>(n_public, 1);
    if (!st->seeds.reserve(W * 8) || !st->containers.reserve(W * words * 8) ||
        !st->fs.reserve((pub_w + W * 2 + W * 8 + W * 2 + 3 * W * 2) * 8)) return 3;
    u64* d_seeds = static_cast<u64*>(st->seeds.ptr);
    u64* d_cont = static_cast<u64*>(st->containers.ptr);
    u64* d_pub = static_cast<u64*>(st->fs.ptr);
    u64* d_ab = d_pub + pub_w;
    u64* d_hash = d_ab + W * 2;
    u64* d_evq = d_hash + W * 8;
    u64* d_evabc = d_evq + W * 2;
    const u64* dz = static_cast<const u64*>(st->z.ptr);
    const u64* dQ = static_cast<const u64*>(st->qbuf.ptr);
    bool ok = cuda_ok(cudaMemcpyAsync(d_seeds, seeds, W * 8, cudaMemcpyHostToDevice, s), "H2D seeds") &&
              lwe_commit_launch(lwe, dQ, m, d_seeds, W, d_cont, s);
    if (ok && n_public)     // public inputs = z[0 .. l) of every witness (r1cs.rs:178-181)
        ok = cuda_ok(cudaMemcpy2DAsync(d_pub, n_public * 8, dz, (size_t)cols * 8, n_public * 8, W, cudaMemcpyDeviceToDevice, s), "D2D public inputs");
    ok = ok && fs_challenge_launch(d_pub, n_public, d_cont, words, W, h->q, true, d_ab, d_hash, s) &&
         poly_eval_launch(h->q, dQ, m, W, d_ab, 2, W, d_evq, s) &&
         poly_eval_launch(h->q, static_cast<const u64*>(st->coef.ptr), m, 3 * W, d_ab, 2, W, d_evabc, s);
    std::vector<u64> evq(W * 2), evabc(3 * W * 2);
    ok = ok && cuda_ok(cudaMemcpyAsync(containers, d_cont, W * words * 8, cudaMemcpyDeviceToHost, s), "D2H containers") &&
         cuda_ok(cudaMemcpyAsync(challenges, d_ab, W * 16, cudaMemcpyDeviceToHost, s), "D2H challenges") &&
         cuda_ok(cudaMemcpyAsync(hashes, d_hash, W * 64, cudaMemcpyDeviceToHost, s), "D2H hashes") &&
         cuda_ok(cudaMemcpyAsync(evq.data(), d_evq, W * 16, cudaMemcpyDeviceToHost, s), "D2H evaluations") &&
         cuda_ok(cudaMemcpyAsync(evabc.data(), d_evabc, 3 * W * 16, cudaMemcpyDeviceToHost, s), "D2H evaluations") &&
         cuda_ok(cudaStreamSynchronize(s), "sync");
    if (!ok) return 4;
    for (size_t w = 0; w < W; w++) {
        u64* e = evals + 8 * w;
        e[0] = evq[2 * w]; e[1] = evq[2 * w + 1];
        for (size_t mat = 0; mat < 3; mat++) {
            e[2 + mat] = evabc[2 * (mat * W + w)];          // at alpha
            e[5 + mat] = evabc[2 * (mat * W + w) + 1];      // at beta
        }
    }
    return 0;
}

}  // namespace lsr
