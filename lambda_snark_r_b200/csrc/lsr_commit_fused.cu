// lsr_commit_fused.cu -- K4: one kernel per batch of Module-LWE commitments.
//
// One CTA per commitment.  Everything between the message / seed coming in and
// t = A*s + e + Delta*m going out stays on chip:
//   phase 1  sample s_0..s_{k-1} (u64 residues) and e_0..e_{k-1} (int8) into
//            shared memory: ChaCha keystream + constant-time CDT (K7 inlined;
//            the CDT is searched with warp shuffles, see cdt_magnitude_shfl)
//   phase 2  forward NTT of all k secret polynomials, one multi-polynomial tile
//   phase 3  NTT-domain mat-vec IN PLACE: s-hat[.][x] <- A-hat[.][.][x] * s-hat[.][x]
//            (Shoup MACs; A-hat read once per commitment from L2)
//   phase 4  inverse NTT of the k rows, same tile
//   phase 5  + e_i (+ Delta*m on the last row) fused into the coalesced store of
//            the LweCommitment container
// Shared memory: k*n*8 (residues) + k*n (errors) = 72 KiB at n=4096, k=2 in the general form; the n = 4096
// form (fused_fast) samples s and e in the registers that consume them and needs 68 KiB (padded layout).
// Three CTAs share an SM either way.  HBM traffic per commitment: message in (8n) +
// container out (8kn + 8).
//
// Replaces Encryptor::encrypt_symmetric + BatchEncoder::encode +
// ciphertext_to_commitment of the reference (cpp-core/src/commitment.cpp:44-60,
// 152-156).
#include <algorithm>
#include <cstdlib>

#include "lsr_engine.h"
#include "lsr_ntt.cuh"
#include "lsr_sampler.cuh"

#ifndef LSR_FUSED_FAST
#define LSR_FUSED_FAST 1
#endif

namespace lsr {

struct FusedParams {
    ModParams mp;
    NttTables tbl;
    ChaChaKey key;
    const ulonglong2* A2;     // [K][K][n]  NTT domain; (a, floor(a*2^64/q)), or for POL_F64 the doubles (a, a/q)
    u64 delta;
    u64 p;
    u64 pinv;                 // floor((2^64 - 1) / p)
    const u64* msgs;          // [count][msg_len]
    const u64* seeds;         // [count]
    u64* out;                 // [count][1 + K*n]
    u32 msg_len;              // true stride of msgs
    u32 msg_used;             // min(msg_len, n)
    // digit planes: commitment b of the launch is unit g = unit0 + b; it commits digit g % planes
    // (base p) of message row g / planes.  planes = 1: the message itself (mod p).
    u32 planes;
    u32 unit0;
    u64 pdiv[4];              // p^l
    u64 pdinv[4];             // floor((2^64 - 1) / p^l)
#ifdef LSR_PROFILING
    u32 skip;                 // tools/ build only (-DLSR_PROFILING, env LSR_FUSED_SKIP): bit i set = phase i+1 not executed
#endif
};

// Phase switches exist only in the profiling build of tools/fused_phases.py; in the shipped library every phase runs.
#ifdef LSR_PROFILING
#define LSR_SKIP(fp, bits) (((fp).skip & (bits)) != 0u)
#else
#define LSR_SKIP(fp, bits) false
#endif

// FAST shape: n = 4096.  n/16 chunks = kNttThreads and plan<12> = 4+4+4, so the radix-16 first forward pass and
// last inverse pass give thread tau exactly the coefficients tau + 256 j of every polynomial -- the chunk tau of the
// randomness layout.  s is then sampled straight into the registers of the first forward pass and e
// straight into the registers of the epilogue: neither crosses shared memory, the e buffer disappears, and the 4 KiB
// per polynomial that frees pay for the padded layout (lsr_ntt.cuh padx) and leave the L1 some room for twiddles.
template <int LOGN, int K>
__host__ __device__ constexpr bool fused_fast() { return LOGN == 12 && kNttThreads == 256 && LSR_FUSED_FAST; }

template <int LOGN, int K>
__host__ __device__ constexpr size_t fused_smem() {
    return fused_fast<LOGN, K>() ? (((size_t)K << LOGN) + ((size_t)K << (LOGN - 4))) * sizeof(u64)
                                 : ((size_t)K << LOGN) * sizeof(u64) + ((size_t)K << LOGN);
}

template <int LOGN, int K>
__host__ __device__ constexpr int fused_min_blocks() {
    return fused_smem<LOGN, K>() <= 75 * 1024 ? 3 : (fused_smem<LOGN, K>() <= 113 * 1024 ? 2 : 1);
}

// canonical row coefficient v at tile index idx -> + e (+ Delta * (m[x] mod p) on row K-1), reduced
// pd, pdi: divisor p^plane of the commitment's digit plane and its reciprocal (pd = 0: plane 0, no division)
template <int LOGN, int K>
__device__ __forceinline__ u64 commit_finish(u32 idx, u64 v, int ev, const u64* __restrict__ msg, u64 q, u64 delta,
                                             u64 p, u64 pinv, u32 msg_used, u64 pd, u64 pdi) {
    v += ev < 0 ? q - (u64)(-ev) : (u64)ev;                                    // < 2q
    const u32 x = idx - ((u32)(K - 1) << LOGN);                                // wraps for earlier rows
    if ((K == 1 || idx >= ((u32)(K > 1 ? K - 1 : 1) << LOGN)) && x < msg_used) {
        u64 word = __ldcs(msg + x);
        if (pd) word = div_small(word, pd, pdi);                               // uniform over the CTA
        // messages are field elements in practice (>= p more often than not): Barrett, not a 64-bit division
        const u64 m = p < (1ull << 21) ? (u64)mod_small(word, (u32)p, pinv) : word % p;
        v = csub(v + delta * m, q);                                            // delta*m <= q-1
    }
    return csub(v, q);
}

// same without the error term (already added): v canonical -> + Delta * (m[x] mod p) on row K-1
template <int LOGN, int K>
__device__ __forceinline__ u64 commit_finish_msg(u32 idx, u64 v, const u64* __restrict__ msg, u64 q, u64 delta,
                                                 u64 p, u64 pinv, u32 msg_used, u64 pd, u64 pdi) {
    const u32 x = idx - ((u32)(K - 1) << LOGN);                                // wraps for earlier rows
    if ((K == 1 || idx >= ((u32)(K > 1 ? K - 1 : 1) << LOGN)) && x < msg_used) {
        u64 word = __ldcs(msg + x);
        if (pd) word = div_small(word, pd, pdi);                               // uniform over the CTA
        const u64 m = p < (1ull << 21) ? (u64)mod_small(word, (u32)p, pinv) : word % p;
        v = csub(v + delta * m, q);                                            // delta*m <= q-1
    }
    return v;
}

// legacy epilogue: e comes from the int8 buffer in shared memory
template <int LOGN, int K>
struct CommitEpilogue {
    static constexpr bool kWholeItem = false;
    static constexpr bool kRawF64 = false;
    struct Pre {};
    const signed char* E;
    const u64* msg;
    u64 q, delta, p, pinv;
    u32 msg_used;
    u64 pd, pdi;
    __device__ __forceinline__ u64 operator()(u32 idx, u64 v) const {
        return commit_finish<LOGN, K>(idx, v, (int)E[idx], msg, q, delta, p, pinv, msg_used, pd, pdi);
    }
};

// one CDT sample from a 64-bit uniform word; all 32 lanes of the warp call together
// NCH8 = 0: the compact search (tables with at most 31 distinct values; the host picks the instantiation, so only one
// search is compiled into a kernel -- the fused kernel is 150 KB of code and its three resident CTAs are rarely in the
// same phase); NCH8 = 5, 8: binary search over 31 entries + tail scan.
template <int NCH8>
struct CdtLanes {
    const CdtParam& cdt;
    u64 lane_entry;
    u32 lane_cum;
    u32 lane_fast;                // ~fast[lane] (compact tables only)
    __device__ __forceinline__ u32 operator()(u64 u) const {
        if constexpr (NCH8 == 0) return cdt_magnitude_compact(cdt.dval[15], lane_entry, lane_cum, u);
        else return cdt_magnitude_shfl<NCH8>(cdt, lane_entry, u);
    }
};

__device__ __forceinline__ u32 pack_sample(u32 mag, u32 w, u32 j) {
    const int sv = (w & 1u) ? -(int)mag : (int)mag;                               // -0 == 0: no test of mag needed
    return ((u32)sv & 0xffu) << (8u * (j & 3u));
}

// The 16 samples of chunk tau of polynomial P as signed bytes packed into four words (lane j in byte j & 3 of word
// j >> 2): one ChaCha block, word j = sign draw + top 31 bits of the magnitude draw of lane j.
// Compact tables: cdt_fast_probe decides every sample from the top 25 bits; only when one of the 512 draws of the warp
// ties with a table prefix (probability ~ 4e-4 per call at sigma = 3.19) the refinement blocks are generated and the
// whole chunk is redone with the full 64-bit search -- the branch is warp-uniform (__any_sync) and what it reveals
// is that SOME draw of the warp shared 25 leading bits with a table entry, not which nor its value.
// Sampling into 4 registers instead of 16 doubles keeps the register pressure low enough for the compiler to interleave
// the searches of a block.
template <int NCH8>
__device__ __forceinline__ void sample_chunk_packed(const ChaChaKey& key, u32 s_lo, u32 s_hi, u32 tau, u32 P,
                                                    const CdtLanes<NCH8>& cdtl, u32 (&pk)[4]) {
    u32 x[16];
    chacha_block(key, s_lo, s_hi, tau, kDomCommit | P, x);
    bool redo = NCH8 != 0;
    if constexpr (NCH8 == 0) {
        const u32 mid_e = ~cdtl.cdt.fast[15];
        u32 tie = 0u;
        pk[0] = pk[1] = pk[2] = pk[3] = 0u;
#pragma unroll
        for (u32 j = 0; j < 16; j++) {
            const u32 e = cdt_fast_probe(mid_e, cdtl.lane_fast, x[j]);
            tie = max(tie, (e ^ x[j]) | 0x7fu);
            pk[j >> 2] |= pack_sample((e & 0x7fu) ^ 0x7fu, x[j], j);
        }
        redo = __any_sync(0xffffffffu, tie == 0xffffffffu) != 0;
    }
    if (redo) {
        pk[0] = pk[1] = pk[2] = pk[3] = 0u;
#pragma unroll 1
        for (u32 h = 0; h < 2; h++) {                 // one copy of the block + searches in the code, run twice
            u32 f[16];
            chacha_block(key, s_lo, s_hi, tau, kDomCommit | kDomCommitFine | (2u * P + h), f);
            u32 a = 0u, b = 0u;
#pragma unroll
            for (u32 w = 0; w < 8; w++) {
                const u32 xw = h ? x[8 + w] : x[w];
                const u32 byte = pack_sample(cdtl(commit_draw(xw, f[2 * w], f[2 * w + 1])), xw, w);
                if (w < 4) a |= byte; else b |= byte;
            }
            if (h == 0) { pk[0] = a; pk[1] = b; } else { pk[2] = a; pk[3] = b; }
        }
    }
}
__device__ __forceinline__ int unpack_s8(const u32 (&pk)[4], u32 j) {
    return (int)(signed char)(pk[j >> 2] >> (8u * (j & 3u)));
}

// FAST epilogue: the work item (row, tau) samples its own 16 error terms -- before its coefficients are loaded --
// and finishes + stores its coefficients
// POL_F64: the last inverse pass hands over balanced doubles r (|r| <= 0.75 q, lsr_ntt.cuh RAW) and the error term is
// added BEFORE the one canonicalisation: e + 128 sits in the low word of the double 2^52 + (e + 128) (no conversion
// instruction), x = r + that is exact, and the compare / select / add of f_to_canonical works on the biased value --
// 6 FP64-pipe and select instructions per coefficient where the integer form (canonicalise, 64-bit e mod q, add,
// conditional subtract) took 19.
template <int LOGN, int K, int NCH8, int POL>
struct CommitEpilogueFast {
    static constexpr bool kWholeItem = true;
    static constexpr bool kRawF64 = POL == POL_F64;
    // PARKED (k >= 2): the last inverse pass takes the rows in order, and once a thread has loaded its 16 coefficients of
    // row 0 into registers those 16 shared-memory slots are its own and dead.  It parks the 16 message words it will need
    // on the message row (same coefficient indices) there with cp.async: no register holds them, and by the time the
    // message row's work item reaches its epilogue they have long arrived -- the epilogue reads shared memory instead of
    // waiting on L2 / HBM (the first use of a message word was 7.9 % of the kernel's stall samples).
    static constexpr bool kParkMsg = kRawF64 && K >= 2;
    struct Pre { u32 pk[4]; };
    const FusedParams& fp;
    const CdtLanes<NCH8>& cdtl;
    const u64* msg;
    u32 s_lo, s_hi;
    u64 pd, pdi;
    u64* sm;                      // the tile (padded layout)
    static constexpr u32 kStep = (1u << (LOGN - 4)) + (1u << (LOGN - 8));      // padded stride of n/16 coefficients
    __device__ __forceinline__ Pre pre(u32 W) const {
        constexpr u32 LG = LOGN - 4;
        const u32 row = W >> LG, tau = W & ((1u << LG) - 1u);
        Pre r;
        if (LSR_SKIP(fp, 33u)) { r.pk[0] = r.pk[1] = r.pk[2] = r.pk[3] = 0u; }
        else sample_chunk_packed<NCH8>(fp.key, s_lo, s_hi, tau, (u32)K + row, cdtl, r.pk);
        return r;
    }
    // called right after the work item's coefficients left shared memory
    __device__ __forceinline__ void loaded(u32 W, u32 base) const {
        if constexpr (kParkMsg) {
            constexpr u32 LG = LOGN - 4;
            if ((W >> LG) == 0u) {
                const u32 tau = W;
#pragma unroll
                for (u32 j = 0; j < 16; j++) {
                    const u32 x = tau + (j << LG);
                    if (x < fp.msg_used) {
                        const u32 dst = (u32)__cvta_generic_to_shared(sm + padx(base) + j * kStep);
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"(dst), "l"(msg + x) : "memory");
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
        }
    }
    __device__ __forceinline__ void item(u64* __restrict__ g, u32, u32 base, const u64 (&v)[16], const Pre& pre) const {
        constexpr u32 LG = LOGN - 4;
        const ModParams& mp = fp.mp;
        if constexpr (kRawF64) {
            const double bias = kTwo52 + 128.0;                         // x = r + e + bias
            const double off_neg = mp.qd - 128.0, off_pos = -128.0;     // -> (r + e [+ q]) + 2^52
            auto canonical = [&](u32 j) -> u64 {                        // (r + e) mod q of coefficient j of the work item
                const u32 byte = ((pre.pk[j >> 2] >> (8u * (j & 3u))) & 0xffu) ^ 0x80u;        // e + 128
                const double x = __dadd_rn(as_d(v[j]), as_d(kTwo52Bits | (u64)byte));
                const double off = x < bias ? off_neg : off_pos;
                return as_u(__dadd_rn(x, off)) & 0x000fffffffffffffull;
            };
            // the row is uniform over the work item (and over the warp): rows without a message take a loop with no
            // message code in it
            const bool msg_row = K == 1 || base >= ((u32)(K - 1) << LOGN);
            if (!msg_row) {
#pragma unroll
                for (u32 j = 0; j < 16; j++) __stcs(g + base + (j << LG), canonical(j));
            } else if constexpr (kParkMsg) {
                asm volatile("cp.async.wait_all;" ::: "memory");       // this thread's own copies (loaded())
                const u32 x0 = base - ((u32)(K - 1) << LOGN);
#pragma unroll
                for (u32 j = 0; j < 16; j++) {
                    u64 c = canonical(j);
                    if (x0 + (j << LG) < fp.msg_used) {
                        u64 word = sm[padx(x0) + j * kStep];           // parked in row 0's slots of this thread
                        if (pd) word = div_small(word, pd, pdi);                               // uniform over the CTA
                        const u64 m = fp.p < (1ull << 21) ? (u64)mod_small(word, (u32)fp.p, fp.pinv) : word % fp.p;
                        c = csub(c + fp.delta * m, mp.q);                                      // delta * m <= q - 1
                    }
                    __stcs(g + base + (j << LG), c);
                }
            } else {
#pragma unroll
                for (u32 j = 0; j < 16; j++) {
                    const u32 idx = base + (j << LG);
                    __stcs(g + idx, commit_finish_msg<LOGN, K>(idx, canonical(j), msg, mp.q, fp.delta, fp.p, fp.pinv, fp.msg_used, pd, pdi));
                }
            }
        } else {
#pragma unroll
            for (u32 j = 0; j < 16; j++) {
                const u32 idx = base + (j << LG);
                __stcs(g + idx, commit_finish<LOGN, K>(idx, v[j], unpack_s8(pre.pk, j), msg, mp.q, fp.delta, fp.p, fp.pinv, fp.msg_used, pd, pdi));
            }
        }
    }
};

template <int LOGN, int K, int NCH8, int POL>
__global__ void __launch_bounds__(kNttThreads, fused_min_blocks<LOGN, K>())
fused_commit_kernel(const __grid_constant__ FusedParams fp, const __grid_constant__ CdtParam cdt) {
    constexpr u32 n = 1u << LOGN;
    constexpr bool FAST = fused_fast<LOGN, K>();
    static_assert((n >> 4) % 32 == 0, "whole warps must take part in the shuffle search");
    extern __shared__ __align__(16) u64 sm[];
    u64* S = sm;                                                     // [K][n], swizzled (FAST: padded)
    const ModParams& mp = fp.mp;
    const size_t b = blockIdx.x;
    const u64 seed = fp.seeds[b];
    const u32 s_lo = (u32)seed, s_hi = (u32)(seed >> 32);
    const u64 lane_entry = NCH8 == 0 ? cdt.dval[threadIdx.x & 31u]
                                     : cdt.cdf[(threadIdx.x & 31u) < 31u ? (threadIdx.x & 31u) : (u32)(kCdtInline - 1)];
    const CdtLanes<NCH8> cdtl{cdt, lane_entry, cdt.dcum[threadIdx.x & 31u], ~cdt.fast[threadIdx.x & 31u]};

    // digit planes: unit g = unit0 + b commits digit (g % planes) of message row g / planes
    const size_t unit = (size_t)fp.unit0 + b;
    const size_t mrow = fp.planes > 1 ? unit / fp.planes : unit;
    const u32 plane = fp.planes > 1 ? (u32)(unit % fp.planes) : 0u;
    const u64 pd = plane ? fp.pdiv[plane & 3u] : 0ull, pdi = plane ? fp.pdinv[plane & 3u] : 0ull;
    const u64* const msg_row = fp.msgs + mrow * (size_t)fp.msg_len;
    // the message is consumed by the last pass only: pull it towards L2 now (one 128-byte line per thread)
    // so that the epilogue's loads do not wait on HBM
    {
        const u64* m = msg_row;
        for (u32 x = threadIdx.x * 16u; x < fp.msg_used; x += kNttThreads * 16u)
            asm volatile("prefetch.global.L2 [%0];" :: "l"(m + x));
    }
    u64* o = fp.out + b * (1 + (size_t)K * n);
    if (threadIdx.x == 0) o[0] = (u64)K * n * 8;

    if constexpr (FAST) {
        // ---- phase 1+2a: chunk tau = threadIdx.x.  s_P[tau + 256 j] is sampled into v[j], the radix-16 first
        // pass (stages 0-3, grid-uniform twiddles from the kernel parameters) runs on it, and only its output
        // goes to shared memory.
        const u32 tau = threadIdx.x;
#pragma unroll 1
        for (u32 P = 0; P < (u32)K; P++) {
            u32 pk[4] = {0u, 0u, 0u, 0u};
            if (!LSR_SKIP(fp, 1u)) sample_chunk_packed<NCH8>(fp.key, s_lo, s_hi, tau, P, cdtl, pk);
            u64 v[16];
#pragma unroll
            for (u32 j = 0; j < 16; j++) {
                const int sv = unpack_s8(pk, j);
                v[j] = POL == POL_F64 ? as_u((double)sv) : (sv < 0 ? mp.q - (u64)(-sv) : (u64)sv);
            }
            if (!LSR_SKIP(fp, 2u)) fwd_network<4, POL, false, true, (LSR_SMEM_FWD_XR != 0)>(v, fp.tbl.fwd, 1u, mp, 0u, fp.tbl.head_fwd);
            const u32 pb = padx((P << LOGN) + tau);
#pragma unroll
            for (u32 j = 0; j < 16; j++) S[pb + j * 272u] = v[j];              // padded stride of 256 coefficients
        }
        __syncthreads();
        // ---- phase 2b: remaining forward passes (POL_F64: evaluations stay unreduced, the mat-vec product reduces)
        if (!LSR_SKIP(fp, 2u)) tile_forward<LOGN, LOGN, POL, true, 1>(S, fp.tbl, mp, (u32)K * n, 0u);

        // ---- phase 3: mat-vec in place (each coefficient index x is owned by one thread)
        // (requesting the A-hat pairs of the next x ahead of the products of the current one was measured: no gain in the
        // phase alone, 2 % slower overall -- 16 more live registers)
        for (u32 x = threadIdx.x; x < (LSR_SKIP(fp, 4u) ? 0u : n); x += kNttThreads) {
            u64 sv[K];
#pragma unroll
            for (u32 j = 0; j < (u32)K; j++) sv[j] = S[padx((j << LOGN) + x)];
#pragma unroll
            for (u32 i = 0; i < (u32)K; i++) {
                if (POL == POL_F64) {
                    double acc = 0.0;
#pragma unroll
                    for (u32 j = 0; j < (u32)K; j++) {
                        const ulonglong2 a = __ldg(fp.A2 + ((size_t)(i * K + j) << LOGN) + x);
                        acc = __dadd_rn(acc, mulmod_f(as_d(sv[j]), as_d(a.x), as_d(a.y), mp.qd));   // each term <= 0.75 q
                    }
                    S[padx((i << LOGN) + x)] = as_u(acc);                          // |acc| <= 0.75 K q <= 3 q
                } else {
                    u64 acc = 0;
#pragma unroll
                    for (u32 j = 0; j < (u32)K; j++) {
                        const ulonglong2 a = __ldg(fp.A2 + ((size_t)(i * K + j) << LOGN) + x);
                        acc += mulred4(sv[j], a.x, a.y, mp.nq);                   // each term < 4q
                    }
                    S[padx((i << LOGN) + x)] = reduce_small(acc, mp);              // 4Kq <= 16q < 2^7 q
                }
            }
        }
        __syncthreads();

        // ---- phase 4+5: inverse transform of the K rows; its last pass (thread tau again owns tau + 256 j) samples
        // e in registers, adds it (and Delta*m on the last row) and stores the container straight to HBM
        if (!LSR_SKIP(fp, 8u)) {
            const CommitEpilogueFast<LOGN, K, NCH8, POL> epi{fp, cdtl, msg_row, s_lo, s_hi, pd, pdi, S};
            tile_inverse_to_global<LOGN, LOGN, POL, CommitEpilogueFast<LOGN, K, NCH8, POL>, true>(S, o + 1, fp.tbl, mp, (u32)K * n, epi);
        }
        return;
    }

    // =============================== other shapes: sampler -> shared memory -> transforms ===============================
    signed char* E = reinterpret_cast<signed char*>(sm + (size_t)K * n);   // [K][n]
    constexpr u32 CH = n >> 4;                                              // chunks; chunk tau = coefficients tau + CH j
    // ---- phase 1: randomness (DESIGN.md 3.3 layout, same as sample_se_kernel)
    for (u32 tau = threadIdx.x; tau < (LSR_SKIP(fp, 1u) ? 0u : CH); tau += kNttThreads) {
#pragma unroll 1
        for (u32 P = 0; P < 2 * K; P++) {
            u32 pk[4];
            sample_chunk_packed<NCH8>(fp.key, s_lo, s_hi, tau, P, cdtl, pk);
#pragma unroll
            for (u32 j = 0; j < 16; j++) {
                const int sv = unpack_s8(pk, j);
                if (P < K) {
                    S[swz((P << LOGN) + tau + CH * j)] =
                        POL == POL_F64 ? as_u((double)sv) : (sv < 0 ? mp.q - (u64)(-sv) : (u64)sv);
                } else {
                    E[((P - K) << LOGN) + tau + CH * j] = (signed char)sv;
                }
            }
        }
    }
    __syncthreads();

    // ---- phase 2: s-hat = NTT(s), all K polynomials as one multi-polynomial tile
    // (POL_F64: evaluations stay unreduced, |s-hat| < (1 + 0.75 logn) q; the mat-vec product reduces)
    if (!LSR_SKIP(fp, 2u)) tile_forward<LOGN, LOGN, POL>(S, fp.tbl, mp, (u32)K * n, 0u);

    // ---- phase 3: mat-vec in place (each coefficient index x is owned by one thread)
    for (u32 x = threadIdx.x; x < (LSR_SKIP(fp, 4u) ? 0u : n); x += kNttThreads) {
        u64 sv[K];
#pragma unroll
        for (u32 j = 0; j < (u32)K; j++) sv[j] = S[swz((j << LOGN) + x)];
#pragma unroll
        for (u32 i = 0; i < (u32)K; i++) {
            if (POL == POL_F64) {
                double acc = 0.0;
#pragma unroll
                for (u32 j = 0; j < (u32)K; j++) {
                    const ulonglong2 a = __ldg(fp.A2 + ((size_t)(i * K + j) << LOGN) + x);
                    acc = __dadd_rn(acc, mulmod_f(as_d(sv[j]), as_d(a.x), as_d(a.y), mp.qd));   // each term <= 0.75 q
                }
                S[swz((i << LOGN) + x)] = as_u(acc);                           // |acc| <= 0.75 K q <= 3 q
            } else {
                u64 acc = 0;
#pragma unroll
                for (u32 j = 0; j < (u32)K; j++) {
                    const ulonglong2 a = __ldg(fp.A2 + ((size_t)(i * K + j) << LOGN) + x);
                    acc += mulred4(sv[j], a.x, a.y, mp.nq);                   // each term < 4q
                }
                S[swz((i << LOGN) + x)] = reduce_small(acc, mp);               // 4Kq <= 16q < 2^7 q
            }
        }
    }
    __syncthreads();

    // ---- phase 4+5: rows of A*s back to coefficients; the last inverse pass (coalesced
    // thread -> coefficient map) adds e (and Delta*m on the last row) in registers and
    // stores the LweCommitment container straight to HBM
    if (!LSR_SKIP(fp, 8u)) {
        const CommitEpilogue<LOGN, K> epi{E, msg_row, mp.q, fp.delta, fp.p, fp.pinv, fp.msg_used, pd, pdi};
        tile_inverse_to_global<LOGN, LOGN, POL>(S, o + 1, fp.tbl, mp, (u32)K * n, epi);
    }
}

// ---------------------------------------------------------------------------
// K6 fused: lwe_verify_opening for a batch, one CTA per commitment (n = 4096, 2 <= k <= 4).  Replaces decrypt + decode +
// compare of the reference (cpp-core/src/commitment.cpp:200-232) by the trapdoor opening of DESIGN.md 3.4:
//   rows t_0 .. t_{k-2} of the container -> shared memory (range check) -> forward transforms (one multi-polynomial tile)
//   -> u^ = sum_i z'^_i * t^_i in place -> inverse transform, whose last pass (thread tau owns coefficients tau + 256 j)
//   adds t_{k-1}, decodes round(u / Delta) mod p and ORs the difference with the caller's words mod p into a register
//   -> one warp reduction and one atomicOr per warp.  Nothing but the container, the message and two flag words touches HBM
//   (the generic path is five kernels with the k - 1 transforms making two round trips each).
// ---------------------------------------------------------------------------
struct VerifyParams {
    ModParams mp;
    NttTables tbl;
    const u64* zh;            // [K-1][n] canonical residues, NTT domain
    const u64* comm;          // [count][stride] containers
    const u64* msgs;          // [count][cmp_len] (already cut to the compared length)
    unsigned long long* diff; // [count], zeroed
    int* invalid;             // [count], zeroed
    size_t stride;
    u64 delta, dinv;          // dinv = floor((2^64 - 1) / delta)
    u64 p, pinv;
    u32 cmp_len;
};

template <int LOGN, int K>
struct VerifyEpilogue {
    static constexpr bool kWholeItem = true;
    static constexpr bool kRawF64 = false;
    struct Pre {};
    const VerifyParams& vp;
    const u64* last;          // row K-1 of this commitment's container
    const u64* msg;           // this commitment's message words
    unsigned long long* acc;  // per-thread OR of the differences
    int* bad;                 // per-thread: a container word out of range
    __device__ __forceinline__ Pre pre(u32) const { return Pre{}; }
    __device__ __forceinline__ void loaded(u32, u32) const {}
    __device__ __forceinline__ void item(u64*, u32, u32 base, const u64 (&v)[16], const Pre&) const {
        constexpr u32 LG = LOGN - 4;
        const u64 q = vp.mp.q;
        // four quarters of four coefficients: the last-row words and the message words of a quarter are requested together
        // (both were pulled towards L2 at kernel start), so a work item waits four times, not once per coefficient (eight at a
        // time spill at the 80-register budget)
#pragma unroll
        for (u32 h = 0; h < 4; h++) {
            u64 t[4], mw[4];
#pragma unroll
            for (u32 jj = 0; jj < 4; jj++) {
                const u32 x = base + ((4u * h + jj) << LG);
                t[jj] = __ldcs(last + x);
                mw[jj] = x < vp.cmp_len ? __ldcs(msg + x) : 0ull;
            }
#pragma unroll
            for (u32 jj = 0; jj < 4; jj++) {
                const u32 x = base + ((4u * h + jj) << LG);
                u64 w = t[jj];
                if (w >= q) { *bad = 1; w = 0; }                               // flagged invalid; keep the arithmetic in range
                const u64 u = addmod(v[4u * h + jj], w, q);
                u64 d = div_small(u + vp.delta / 2, vp.delta, vp.dinv);        // round(u / Delta) <= p
                d = d >= vp.p ? d - vp.p : d;
                const u64 m = vp.p < (1ull << 21) ? (u64)mod_small(mw[jj], (u32)vp.p, vp.pinv) : mw[jj] % vp.p;
                *acc |= x < vp.cmp_len ? (d ^ m) : 0ull;                        // no branch on the outcome
            }
        }
    }
};

template <int LOGN, int K, int POL>
__global__ void __launch_bounds__(kNttThreads, 3)
fused_verify_kernel(const __grid_constant__ VerifyParams vp) {
    constexpr u32 n = 1u << LOGN;
    static_assert(K >= 2, "k = 1 has no trapdoor transform: the generic path decodes row 0 directly");
    extern __shared__ __align__(16) u64 sm[];                                  // [K-1][n], swizzled
    const ModParams& mp = vp.mp;
    const size_t b = blockIdx.x;
    const u64* c = vp.comm + b * vp.stride;
    int bad = (threadIdx.x == 0 && c[0] != (u64)K * n * 8) ? 1 : 0;
    {   // the last row and the message are consumed by the last pass only: pull them towards L2 now (one 128-byte line each)
        const u64* lastrow = c + 1 + (size_t)(K - 1) * n;
        const u64* m = vp.msgs + b * (size_t)vp.cmp_len;
        for (u32 x = threadIdx.x * 16u; x < n; x += kNttThreads * 16u) {
            asm volatile("prefetch.global.L2 [%0];" :: "l"(lastrow + x));
            if (x < vp.cmp_len) asm volatile("prefetch.global.L2 [%0];" :: "l"(m + x));
        }
    }
    // rows t_0 .. t_{k-2} into shared memory: all the words of a row that a thread owns are requested before the first is
    // checked (one word per iteration left the kernel waiting on HBM at 20 % of its bandwidth: 26 % of the stall samples sat
    // on the range check behind the load, profiles/r02_ncu_verify_before.txt)
    constexpr u32 PER_THREAD = n / kNttThreads;
#pragma unroll 1
    for (u32 row = 0; row < (u32)(K - 1); row++) {
        const u64* __restrict__ src = c + 1 + ((size_t)row << LOGN) + threadIdx.x;
        u64 x[PER_THREAD];
#pragma unroll
        for (u32 k = 0; k < PER_THREAD; k++) x[k] = __ldcs(src + k * kNttThreads);
#pragma unroll
        for (u32 k = 0; k < PER_THREAD; k++) {
            const bool out_of_range = x[k] >= mp.q;                            // flagged invalid; keep the arithmetic in range
            bad |= out_of_range ? 1 : 0;
            sm[swz((row << LOGN) + threadIdx.x + k * kNttThreads)] = to_working<POL>(out_of_range ? 0ull : x[k]);
        }
    }
    __syncthreads();
    tile_forward<LOGN, LOGN, POL>(sm, vp.tbl, mp, (u32)(K - 1) * n, 0u);
    // u^ = sum_i z'^_i * t^_i in place in row 0 (each coefficient index is owned by one thread).  The trapdoor words a thread
    // needs from a row are requested together (one dependent L2 load per coefficient was a quarter of the kernel's stall samples)
#pragma unroll 1
    for (u32 i = 0; i < (u32)(K - 1); i++) {
        const u64* __restrict__ zr = vp.zh + ((size_t)i << LOGN) + threadIdx.x;
        u64 zw[PER_THREAD];
#pragma unroll
        for (u32 k = 0; k < PER_THREAD; k++) zw[k] = __ldg(zr + k * kNttThreads);
#pragma unroll
        for (u32 k = 0; k < PER_THREAD; k++) {
            const u32 x = threadIdx.x + k * kNttThreads;
            if (POL == POL_F64) {
                const double z = u64_to_f(zw[k]);
                // z / q formed on the fly: two roundings instead of one, |x| 2^-52 <= 2^-4 for the unreduced evaluations (< 2^48)
                const double term = mulmod_f(as_d(sm[swz((i << LOGN) + x)]), z, __dmul_rn(z, mp.invq), mp.qd);
                sm[swz(x)] = as_u(i ? __dadd_rn(as_d(sm[swz(x)]), term) : term);        // |acc| <= 0.75 (K-1) q
            } else {
                const u64 term = mulmod_exact(zw[k], sm[swz((i << LOGN) + x)], mp);
                sm[swz(x)] = i ? addmod(sm[swz(x)], term, mp.q) : term;
            }
        }
    }
    __syncthreads();
    unsigned long long acc = 0ull;
    const VerifyEpilogue<LOGN, K> epi{vp, c + 1 + (size_t)(K - 1) * n, vp.msgs + b * (size_t)vp.cmp_len, &acc, &bad};
    tile_inverse_to_global<LOGN, LOGN, POL, VerifyEpilogue<LOGN, K>>(sm, nullptr, vp.tbl, mp, n, epi);
    // warp OR, one atomic per warp (every warp issues it: no branch on the outcome)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        acc |= __shfl_xor_sync(0xffffffffu, acc, o);
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    if ((threadIdx.x & 31u) == 0u) {
        atomicOr(vp.diff + b, acc);
        atomicOr(vp.invalid + b, bad);
    }
}

bool fused_verify_supported(const LweContext* c) {
    const bool arith_ok = (c->ntt->mp.lazy_fwd && c->ntt->mp.lazy_inv) || (c->ntt->mp.f64_ok && c->ntt->arith != 1);
    return c->logn == 12 && c->k >= 2 && c->k <= 4 && arith_ok && c->commit_path != 1;
}

template <int K, int POL>
static bool launch_fused_verify_pol(const VerifyParams& vp, size_t count, cudaStream_t s) {
    constexpr size_t smem = ((size_t)(K - 1) << 12) * sizeof(u64);
    auto kernel = fused_verify_kernel<12, K, POL>;
    if (smem > 48 * 1024 &&
        !cuda_ok(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute"))
        return false;
    kernel<<<(unsigned)count, kNttThreads, smem, s>>>(vp);
    return cuda_ok(cudaGetLastError(), "fused_verify_kernel launch");
}

// d_comm [count][stride], d_msgs [count][cmp_len], d_diff / d_invalid zeroed by the caller; asynchronous on s
bool fused_verify_launch(const LweContext* c, const u64* d_comm, size_t stride, const u64* d_msgs, size_t cmp_len,
                         size_t count, unsigned long long* d_diff, int* d_invalid, cudaStream_t s) {
    if (count == 0) return true;
    if (count > 0x7fffffffull) { set_error("batch too large for one launch"); return false; }
    const bool f64 = c->ntt->mp.f64_ok && c->ntt->arith != 1;
    VerifyParams vp;
    vp.mp = c->ntt->mp;
    vp.tbl = f64 ? c->ntt->tables_f : c->ntt->tables;
    vp.zh = c->d_zh; vp.comm = d_comm; vp.msgs = d_msgs; vp.diff = d_diff; vp.invalid = d_invalid;
    vp.stride = stride; vp.delta = c->delta; vp.dinv = ~0ull / c->delta; vp.p = c->p; vp.pinv = ~0ull / c->p;
    vp.cmp_len = (u32)cmp_len;
    switch (c->k) {
        case 2: return f64 ? launch_fused_verify_pol<2, POL_F64>(vp, count, s) : launch_fused_verify_pol<2, POL_LAZY>(vp, count, s);
        case 3: return f64 ? launch_fused_verify_pol<3, POL_F64>(vp, count, s) : launch_fused_verify_pol<3, POL_LAZY>(vp, count, s);
        case 4: return f64 ? launch_fused_verify_pol<4, POL_F64>(vp, count, s) : launch_fused_verify_pol<4, POL_LAZY>(vp, count, s);
        default: set_error("fused verify: unsupported module rank"); return false;
    }
}

// ------------------------------------------------------------------- host side
static bool build_cdt_param(const LweContext* c, CdtParam& out) {
    size_t used = c->cdf.size();
    while (used > 0 && c->cdf[used - 1] == ~0ull) --used;     // entries == 2^64-1 are never below u
    if (used > (size_t)kCdtInline) return false;
    if (c->cdf.size() - 1 > 127) return false;                // e held as int8
    out.count = (u32)used;
    for (int i = 0; i < kCdtInline; i++) out.cdf[i] = (size_t)i < used ? c->cdf[i] : ~0ull;
    out.pad = 1;
    for (int i = 31; i < kCdtInline; i++) if ((out.cdf[i] >> 32) != 0xffffffffull) out.pad = 0;
    // compact form: distinct values with cumulative multiplicities (the table is non-decreasing)
    out.pad2 = 0;
    for (int i = 0; i < 32; i++) { out.dval[i] = ~0ull; out.dcum[i] = 0; }
    size_t distinct = 0;
    out.compact = 1;
    for (size_t i = 0; i < used; i++) {
        if (i > 0 && c->cdf[i] == c->cdf[i - 1]) continue;
        if (distinct == 31) { out.compact = 0; break; }
        out.dval[distinct] = c->cdf[i];
        out.dcum[distinct] = (u32)i;                          // entries strictly below this value
        ++distinct;
    }
    for (size_t i = distinct; i < 32; i++) out.dcum[i] = (u32)used;
    for (int i = 0; i < 32; i++) out.fast[i] = ((u32)(out.dval[i] >> 39) << 7) | out.dcum[i];   // dcum <= 64 < 128
    return true;
}

// Test hook: evaluates the three CDT searches on caller-supplied u values so that the
// boundary cases (u = cdf[k] - 1, cdf[k], cdf[k] + 1, including the 2^-60-probability tail
// the keystream never reaches in a test) can be compared with the reference's linear scan.
// cycles (optional): clock64() ticks each warp spent in the search -- the device analogue of the reference's dudect
// harness (cpp-core/tools/dudect_sampler.cpp:105-141): the caller compares the distributions of two input classes
template <int NCH8>
__global__ void cdt_probe_kernel(const __grid_constant__ CdtParam cdt, const u64* __restrict__ cdf_full, u32 cdf_n,
                                 const u64* __restrict__ u, size_t count, u32* __restrict__ out, int variant,
                                 unsigned long long* __restrict__ cycles) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;      // count is padded to a multiple of 32
    const bool compact = variant >= 3 && cdt.compact != 0;
    const u64 lane_entry = compact ? cdt.dval[threadIdx.x & 31u]
                                   : cdt.cdf[(threadIdx.x & 31u) < 31u ? (threadIdx.x & 31u) : (u32)(kCdtInline - 1)];
    const u32 lane_cum = cdt.dcum[threadIdx.x & 31u];
    const u64 x = i < count ? u[i] : 0;
    u32 r;
    __syncwarp();
    const long long t0 = clock64();
    if (variant == 0) r = cdt_magnitude_global(cdf_full, cdf_n, x);
    else if (variant == 1) r = cdt_magnitude<NCH8>(cdt, x);
    else if (compact && variant == 4) {
        // the commitment sampler's decision: 25-bit prefix first, full draw only when some lane of the warp ties
        const u32 w = (u32)(x >> 32);                                 // bits 31..1 = top 31 bits of the draw
        const u32 e = cdt_fast_probe(~cdt.fast[15], ~cdt.fast[threadIdx.x & 31u], w);
        r = (e & 0x7fu) ^ 0x7fu;
        if (__any_sync(0xffffffffu, ((e ^ w) | 0x7fu) == 0xffffffffu))
            r = cdt_magnitude_compact(cdt.dval[15], lane_entry, lane_cum, x);
    }
    else if (compact) r = cdt_magnitude_compact(cdt.dval[15], lane_entry, lane_cum, x);
    else r = cdt_magnitude_shfl<NCH8>(cdt, lane_entry, x);
    // the result feeds the time stamp's dependency chain so that the search cannot be scheduled past it
    const long long t1 = clock64() + (long long)(__shfl_sync(0xffffffffu, r, 0) >> 31);
    if (i < count) out[i] = r;
    if (cycles && (threadIdx.x & 31u) == 0u) cycles[i >> 5] = (unsigned long long)(t1 - t0);
}

bool cdt_probe_host(double sigma, const u64* u, size_t count, uint32_t* out, int variant, u64* cycles_per_warp) {
    LweContext fake;
    fake.cdf = host::build_cdt(sigma);
    if (fake.cdf.empty()) return false;
    CdtParam cdt;
    const bool inline_ok = build_cdt_param(&fake, cdt);
    if (variant != 0 && !inline_ok) { set_error("CDT does not fit the inline table"); return false; }
    if (!cuda_ok(cudaSetDevice(current_device_choice()), "cudaSetDevice")) return false;
    u64 *d_u = nullptr, *d_cdf = nullptr;
    u32* d_out = nullptr;
    unsigned long long* d_cyc = nullptr;
    const size_t warps = (count + 31) / 32;
    if (cycles_per_warp && !cuda_ok(cudaMalloc(&d_cyc, warps * 8), "cudaMalloc")) return false;
    bool ok = cuda_ok(cudaMalloc(&d_u, count * 8), "cudaMalloc") && cuda_ok(cudaMalloc(&d_out, count * 4), "cudaMalloc") &&
              cuda_ok(cudaMalloc(&d_cdf, fake.cdf.size() * 8), "cudaMalloc") &&
              cuda_ok(cudaMemcpy(d_u, u, count * 8, cudaMemcpyHostToDevice), "H2D") &&
              cuda_ok(cudaMemcpy(d_cdf, fake.cdf.data(), fake.cdf.size() * 8, cudaMemcpyHostToDevice), "H2D");
    if (ok) {
        const unsigned blocks = (unsigned)((count + 127) / 128);
        if (!inline_ok || (cdt.count + 7) / 8 <= 5) cdt_probe_kernel<5><<<blocks, 128>>>(cdt, d_cdf, (u32)fake.cdf.size(), d_u, count, d_out, variant, d_cyc);
        else cdt_probe_kernel<8><<<blocks, 128>>>(cdt, d_cdf, (u32)fake.cdf.size(), d_u, count, d_out, variant, d_cyc);
        ok = cuda_ok(cudaGetLastError(), "cdt_probe_kernel") && cuda_ok(cudaMemcpy(out, d_out, count * 4, cudaMemcpyDeviceToHost), "D2H");
        if (ok && d_cyc) ok = cuda_ok(cudaMemcpy(cycles_per_warp, d_cyc, warps * 8, cudaMemcpyDeviceToHost), "D2H");
    }
    cudaFree(d_u); cudaFree(d_out); cudaFree(d_cdf); cudaFree(d_cyc);
    return ok;
}

static bool fused_shape_ok(uint32_t logn, uint32_t k) {
    if (logn == 12) return k >= 1 && k <= 4;
    if (logn == 10 || logn == 11 || logn == 13) return k == 2;
    return false;
}

bool fused_commit_supported(const LweContext* c) {
    CdtParam tmp;
    const bool arith_ok = (c->ntt->mp.lazy_fwd && c->ntt->mp.lazy_inv) || (c->ntt->mp.f64_ok && c->ntt->arith != 1);
    return fused_shape_ok(c->logn, c->k) && arith_ok && build_cdt_param(c, tmp);
}

// (a, floor(a * 2^64 / q)) pairs of A-hat for the fused kernel's Shoup MACs
__global__ void shoup_table_kernel(u64 q, const u64* __restrict__ A, size_t total, ulonglong2* __restrict__ A2) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const u64 a = A[i];
    // long division of (a : 0) by q, 64 shift-subtract steps (q < 2^61, no overflow)
    u64 rem = a % q, quo = 0;
    for (int bit = 0; bit < 64; bit++) {
        rem <<= 1;
        quo <<= 1;
        if (rem >= q) { rem -= q; quo |= 1; }
    }
    A2[i] = make_ulonglong2(a, quo);
}

// (a, a / q) as doubles for the POL_F64 mat-vec (a < 2^45 converts exactly; IEEE division)
__global__ void f64_table_kernel(double qd, const u64* __restrict__ A, size_t total, ulonglong2* __restrict__ A2f) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const double a = (double)A[i];
    A2f[i] = make_ulonglong2(as_u(a), as_u(__ddiv_rn(a, qd)));
}

// called once from lwe_create (context set-up), so launches never race on it
bool fused_prepare(LweContext* mc, cudaStream_t s) {
    const LweContext* c = mc;
    if (c->d_A2 || !fused_shape_ok(c->logn, c->k)) return true;
    const size_t total = (size_t)c->k * c->k * c->n;
    if (!cuda_ok(cudaMalloc(&mc->d_A2, total * sizeof(ulonglong2)), "cudaMalloc(A2)")) { mc->d_A2 = nullptr; return false; }
    shoup_table_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(c->q, c->d_A, total, mc->d_A2);
    if (!cuda_ok(cudaGetLastError(), "shoup_table_kernel")) return false;
    if (c->ntt->mp.f64_ok) {
        if (!cuda_ok(cudaMalloc(&mc->d_A2f, total * sizeof(ulonglong2)), "cudaMalloc(A2f)")) { mc->d_A2f = nullptr; return false; }
        f64_table_kernel<<<(unsigned)((total + 255) / 256), 256, 0, s>>>(c->ntt->mp.qd, c->d_A, total, mc->d_A2f);
        if (!cuda_ok(cudaGetLastError(), "f64_table_kernel")) return false;
    }
    return true;
}

template <int LOGN, int K, int POL>
static bool launch_fused_pol(const FusedParams& fp, const CdtParam& cdt, size_t count, cudaStream_t s) {
    constexpr size_t smem = fused_smem<LOGN, K>();
    const int nch8 = (int)((cdt.count + 7) / 8);
    auto run = [&](auto kernel) -> bool {
        if (smem > 48 * 1024 &&
            !cuda_ok(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cudaFuncSetAttribute"))
            return false;
        kernel<<<(unsigned)count, kNttThreads, smem, s>>>(fp, cdt);
        return cuda_ok(cudaGetLastError(), "fused_commit_kernel launch");
    };
    if (cdt.compact) return run(fused_commit_kernel<LOGN, K, 0, POL>);
    if (nch8 <= 5) return run(fused_commit_kernel<LOGN, K, 5, POL>);
    return run(fused_commit_kernel<LOGN, K, 8, POL>);
}

template <int LOGN, int K>
static bool launch_fused(const FusedParams& fp, const CdtParam& cdt, size_t count, cudaStream_t s, bool f64) {
    if (f64) return launch_fused_pol<LOGN, K, POL_F64>(fp, cdt, count, s);
    return launch_fused_pol<LOGN, K, POL_LAZY>(fp, cdt, count, s);
}

bool fused_commit_launch(const LweContext* c, const u64* d_msgs, size_t msg_len, const u64* d_seeds,
                         size_t count, u64* d_out, cudaStream_t s, uint32_t planes, size_t unit0) {
    if (count == 0) return true;
    if (planes < 1 || planes > 4 || unit0 > 0xffffffffull) { set_error("fused path: bad digit planes"); return false; }
    if (count > 0x7fffffffull) { set_error("batch too large for one launch"); return false; }
    if (msg_len > 0xffffffffull) { set_error("msg_len too large"); return false; }
    CdtParam cdt;
    if (!build_cdt_param(c, cdt)) { set_error("fused path: CDT too large"); return false; }
    if (!c->d_A2) { set_error("fused path: context not prepared"); return false; }
    const bool f64 = c->ntt->mp.f64_ok && c->ntt->arith != 1;
    if (f64 && !c->d_A2f) { set_error("fused path: context not prepared"); return false; }
    FusedParams fp;
    fp.mp = c->ntt->mp;
    fp.tbl = f64 ? c->ntt->tables_f : c->ntt->tables;
    for (int i = 0; i < 8; i++) fp.key.k[i] = c->key[i];
    fp.A2 = f64 ? c->d_A2f : c->d_A2;
    fp.delta = c->delta;
    fp.p = c->p;
    fp.pinv = ~0ull / c->p;
    fp.msgs = d_msgs;
    fp.seeds = d_seeds;
    fp.out = d_out;
    fp.msg_len = (u32)msg_len;
    fp.msg_used = (u32)std::min<size_t>(msg_len, c->n);
    fp.planes = planes;
    fp.unit0 = (u32)unit0;
    plane_divisors(c->p, fp.pdiv, fp.pdinv);
#ifdef LSR_PROFILING
    const char* skip = std::getenv("LSR_FUSED_SKIP");
    fp.skip = skip ? (u32)std::strtoul(skip, nullptr, 0) : 0u;
#endif
    switch (c->logn * 16 + c->k) {
        case 12 * 16 + 1: return launch_fused<12, 1>(fp, cdt, count, s, f64);
        case 12 * 16 + 2: return launch_fused<12, 2>(fp, cdt, count, s, f64);
        case 12 * 16 + 3: return launch_fused<12, 3>(fp, cdt, count, s, f64);
        case 12 * 16 + 4: return launch_fused<12, 4>(fp, cdt, count, s, f64);
        case 10 * 16 + 2: return launch_fused<10, 2>(fp, cdt, count, s, f64);
        case 11 * 16 + 2: return launch_fused<11, 2>(fp, cdt, count, s, f64);
        case 13 * 16 + 2: return launch_fused<13, 2>(fp, cdt, count, s, f64);
        default: set_error("fused path: unsupported (n, k)"); return false;
    }
}

}  // namespace lsr
