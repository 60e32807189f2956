// lsr_keccak.h -- SHA3-256 of the Fiat-Shamir transcript, written once for host and device.
//
// Replaces Challenge::derive (rust-api/lambda-snark/src/challenge.rs:102-134; the `sha3` crate's
// Sha3_256, FIPS 202): the transcript
//     "LAMBDA-SNARK-R-FS-v1" || le64(#public) || public inputs (le64 each) || le64(#words) || words (le64 each)
// is 20 bytes of domain tag followed by 64-bit little-endian words only, so byte offset 20 + 8k of the
// transcript always sits 4 bytes into a Keccak lane: lane L of rate block b is
//     hi32(word(17 b + L - 3)) | lo32(word(17 b + L - 2)) << 32          (17 lanes = 136-byte rate)
// where word() runs over [#public, public..., #words, words...], followed by the virtual word 0x06 (SHA3
// domain byte) and zeros; 0x80 is XORed into the last byte of the last block.  The three lanes that hold
// the tag are constants.  The state lives in 25 scalar variables (registers on the device).
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define LSR_HD __host__ __device__ __forceinline__
#else
#define LSR_HD inline
#endif

namespace lsr {

typedef unsigned long long kw64;     // the library's u64 (lsr_common.h); same width as uint64_t

LSR_HD kw64 keccak_rotl(kw64 x, int r) { return (x << r) | (x >> (64 - r)); }

// Keccak-f[1600], 24 rounds, fully unrolled state a[x + 5 y]
LSR_HD void keccak_f1600(kw64 (&a)[25]) {
    constexpr kw64 RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int round = 0; round < 24; round++) {
        // theta
        const kw64 c0 = a[0] ^ a[5] ^ a[10] ^ a[15] ^ a[20];
        const kw64 c1 = a[1] ^ a[6] ^ a[11] ^ a[16] ^ a[21];
        const kw64 c2 = a[2] ^ a[7] ^ a[12] ^ a[17] ^ a[22];
        const kw64 c3 = a[3] ^ a[8] ^ a[13] ^ a[18] ^ a[23];
        const kw64 c4 = a[4] ^ a[9] ^ a[14] ^ a[19] ^ a[24];
        const kw64 d0 = c4 ^ keccak_rotl(c1, 1);
        const kw64 d1 = c0 ^ keccak_rotl(c2, 1);
        const kw64 d2 = c1 ^ keccak_rotl(c3, 1);
        const kw64 d3 = c2 ^ keccak_rotl(c4, 1);
        const kw64 d4 = c3 ^ keccak_rotl(c0, 1);
        // rho + pi: b[y + 5 ((2x + 3y) % 5)] = rotl(a[x + 5y] ^ d[x], r[x][y])
        const kw64 b0 = a[0] ^ d0;
        const kw64 b10 = keccak_rotl(a[1] ^ d1, 1);
        const kw64 b20 = keccak_rotl(a[2] ^ d2, 62);
        const kw64 b5 = keccak_rotl(a[3] ^ d3, 28);
        const kw64 b15 = keccak_rotl(a[4] ^ d4, 27);
        const kw64 b16 = keccak_rotl(a[5] ^ d0, 36);
        const kw64 b1 = keccak_rotl(a[6] ^ d1, 44);
        const kw64 b11 = keccak_rotl(a[7] ^ d2, 6);
        const kw64 b21 = keccak_rotl(a[8] ^ d3, 55);
        const kw64 b6 = keccak_rotl(a[9] ^ d4, 20);
        const kw64 b7 = keccak_rotl(a[10] ^ d0, 3);
        const kw64 b17 = keccak_rotl(a[11] ^ d1, 10);
        const kw64 b2 = keccak_rotl(a[12] ^ d2, 43);
        const kw64 b12 = keccak_rotl(a[13] ^ d3, 25);
        const kw64 b22 = keccak_rotl(a[14] ^ d4, 39);
        const kw64 b23 = keccak_rotl(a[15] ^ d0, 41);
        const kw64 b8 = keccak_rotl(a[16] ^ d1, 45);
        const kw64 b18 = keccak_rotl(a[17] ^ d2, 15);
        const kw64 b3 = keccak_rotl(a[18] ^ d3, 21);
        const kw64 b13 = keccak_rotl(a[19] ^ d4, 8);
        const kw64 b14 = keccak_rotl(a[20] ^ d0, 18);
        const kw64 b24 = keccak_rotl(a[21] ^ d1, 2);
        const kw64 b9 = keccak_rotl(a[22] ^ d2, 61);
        const kw64 b19 = keccak_rotl(a[23] ^ d3, 56);
        const kw64 b4 = keccak_rotl(a[24] ^ d4, 14);
        // chi
        a[0] = b0 ^ (~b1 & b2);    a[1] = b1 ^ (~b2 & b3);    a[2] = b2 ^ (~b3 & b4);    a[3] = b3 ^ (~b4 & b0);    a[4] = b4 ^ (~b0 & b1);
        a[5] = b5 ^ (~b6 & b7);    a[6] = b6 ^ (~b7 & b8);    a[7] = b7 ^ (~b8 & b9);    a[8] = b8 ^ (~b9 & b5);    a[9] = b9 ^ (~b5 & b6);
        a[10] = b10 ^ (~b11 & b12); a[11] = b11 ^ (~b12 & b13); a[12] = b12 ^ (~b13 & b14); a[13] = b13 ^ (~b14 & b10); a[14] = b14 ^ (~b10 & b11);
        a[15] = b15 ^ (~b16 & b17); a[16] = b16 ^ (~b17 & b18); a[17] = b17 ^ (~b18 & b19); a[18] = b18 ^ (~b19 & b15); a[19] = b19 ^ (~b15 & b16);
        a[20] = b20 ^ (~b21 & b22); a[21] = b21 ^ (~b22 & b23); a[22] = b22 ^ (~b23 & b24); a[23] = b23 ^ (~b24 & b20); a[24] = b24 ^ (~b20 & b21);
        // iota
        a[0] ^= RC[round];
    }
}

// word k of the transcript's word sequence; k may run past the end (SHA3 pad word, then zeros)
struct FsTranscript {
    const kw64* pub;       // [n_pub]
    kw64 n_pub;
    const kw64* words;     // [n_words]  (LweCommitment::data, commitment.rs:87-93)
    kw64 n_words;
    LSR_HD kw64 total() const { return n_pub + 2 + n_words; }
    LSR_HD kw64 word(long long k) const {
        if (k < 0) return 0;
        const kw64 u = (kw64)k;
        if (u == 0) return n_pub;
        if (u <= n_pub) return pub[u - 1];
        if (u == n_pub + 1) return n_words;
        if (u < total()) return words[u - n_pub - 2];
        return u == total() ? 0x06ull : 0ull;
    }
};

// SHA3-256 of the transcript; out[4] = the digest as four little-endian lanes (digest byte i = byte i%8 of out[i/8])
LSR_HD void fs_sha3_256(const FsTranscript& t, kw64 (&out)[4]) {
    kw64 a[25];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 25; i++) a[i] = 0;
    const kw64 K = t.total();
    const kw64 nblocks = (20 + 8 * K) / 136 + 1;
    for (kw64 b = 0; b < nblocks; b++) {
        const long long k0 = (long long)(17 * b) - 3;
        // bulk blocks: all 18 words word(k0) .. word(k0 + 17) come from words[]
        if (k0 >= (long long)(t.n_pub + 2) && (kw64)k0 + 18 <= K) {
            const kw64* w = t.words + ((kw64)k0 - t.n_pub - 2);
            kw64 prev = w[0];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int L = 0; L < 17; L++) {
                const kw64 next = w[L + 1];
                a[L] ^= (prev >> 32) | (next << 32);
                prev = next;
            }
        } else {
            kw64 prev = t.word(k0);
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
            for (int L = 0; L < 17; L++) {
                const kw64 next = t.word(k0 + L + 1);
                kw64 lane = (prev >> 32) | (next << 32);
                if (b == 0) {
                    // "LAMBDA-SNARK-R-FS-v1": lanes 0, 1 and the low half of lane 2
                    if (L == 0) lane = 0x532d4144424d414cULL;                      // "LAMBDA-S"
                    if (L == 1) lane = 0x462d522d4b52414eULL;                      // "NARK-R-F"
                    if (L == 2) lane = 0x31762d53ULL | (next << 32);               // "S-v1" | lo32(word 0)
                }
                a[L] ^= lane;
                prev = next;
            }
        }
        if (b == nblocks - 1) a[16] ^= 0x8000000000000000ULL;
        keccak_f1600(a);
    }
    out[0] = a[0]; out[1] = a[1]; out[2] = a[2]; out[3] = a[3];
}

// ---------------------------------------------------------------------------------------------------------
// Lane-parallel form: state lane a[x + 5y] lives in execution lane l = x + 5y of a warp (lanes 25 .. 31 idle), so one
// statement's hash advances with the latency of ~10 shuffles per round instead of ~150 dependent instructions of one
// thread: 3x lower latency for small batches, where one thread per statement leaves the GPU empty and a proof waits
// on its own transcript (fs_challenge_warp_kernel).  The index maps below are the whole difference to keccak_f1600
// above; the host emulation (same maps, arrays instead of shuffles) is pinned to hashlib by the CPU suite.
// ---------------------------------------------------------------------------------------------------------
LSR_HD unsigned keccak_rho_rot(unsigned l) {        // rotation of a[l] in rho (the constants of keccak_f1600 above)
    const unsigned char t[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    return t[l];
}
LSR_HD unsigned keccak_pi_src(unsigned l) {         // b[l] = rotated a[keccak_pi_src(l)]  (inverse of pi)
    const unsigned char t[25] = {0, 6, 12, 18, 24, 3, 9, 10, 16, 22, 1, 7, 13, 19, 20, 4, 5, 11, 17, 23, 2, 8, 14, 15, 21};
    return t[l];
}
LSR_HD unsigned keccak_row_lane(unsigned l, unsigned dx) { return l - l % 5 + (l % 5 + dx) % 5; }   // (x + dx, y)
LSR_HD unsigned keccak_col_lane(unsigned l, unsigned dy) { return (l + 5 * dy) % 25; }               // (x, y + dy)
#if defined(__CUDACC__)
// the device copy lives in the constant bank (a function-local table is rebuilt on the stack at every call)
static __constant__ kw64 kKeccakRoundConstants[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
#endif
LSR_HD kw64 keccak_round_constant(int round) {
#if defined(__CUDA_ARCH__)
    return kKeccakRoundConstants[round];
#endif
    const kw64 RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    return RC[round];
}
// most significant word of (hi:lo) << (s mod 32)
LSR_HD unsigned keccak_funnel_l(unsigned lo, unsigned hi, unsigned s) {
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, s);
#else
    s &= 31u;
    return s ? (hi << s) | (lo >> (32u - s)) : hi;
#endif
}
// rotation by a per-lane amount r in [0, 63]: swap the halves for r >= 32, then two funnel shifts by r mod 32
LSR_HD kw64 keccak_rotl_var(kw64 x, unsigned r) {
    unsigned lo = (unsigned)x, hi = (unsigned)(x >> 32);
    if (r & 32u) { const unsigned t = lo; lo = hi; hi = t; }
    const unsigned nh = keccak_funnel_l(lo, hi, r), nl = keccak_funnel_l(hi, lo, r);
    return ((kw64)nh << 32) | nl;
}

// number of 136-byte rate blocks of the padded transcript, and lane L (0 .. 16) of block b
LSR_HD kw64 fs_nblocks(const FsTranscript& t) { return (20 + 8 * t.total()) / 136 + 1; }
LSR_HD kw64 fs_block_lane(const FsTranscript& t, kw64 b, int L) {
    const long long k = (long long)(17 * b) - 3 + L;
    const kw64 prev = t.word(k), next = t.word(k + 1);
    kw64 lane = (prev >> 32) | (next << 32);
    if (b == 0) {
        if (L == 0) lane = 0x532d4144424d414cULL;                      // "LAMBDA-S"
        if (L == 1) lane = 0x462d522d4b52414eULL;                      // "NARK-R-F"
        if (L == 2) lane = 0x31762d53ULL | (next << 32);               // "S-v1" | lo32(word 0)
    }
    return lane;
}

#if !defined(__CUDA_ARCH__)
// host emulation of the lane-parallel hash (test infrastructure for the index maps): arrays stand in for shuffles
inline void fs_sha3_256_lanes_emulated(const FsTranscript& t, kw64 (&out)[4]) {
    kw64 a[25] = {0};
    const kw64 nblocks = fs_nblocks(t);
    for (kw64 b = 0; b < nblocks; b++) {
        for (int L = 0; L < 17; L++) a[L] ^= fs_block_lane(t, b, L);
        if (b == nblocks - 1) a[16] ^= 0x8000000000000000ULL;
        for (int round = 0; round < 24; round++) {
            kw64 c[25], bb[25], tt[25];
            for (unsigned l = 0; l < 25; l++)
                c[l] = a[l] ^ a[keccak_col_lane(l, 1)] ^ a[keccak_col_lane(l, 2)] ^ a[keccak_col_lane(l, 3)] ^ a[keccak_col_lane(l, 4)];
            for (unsigned l = 0; l < 25; l++)
                tt[l] = keccak_rotl_var(a[l] ^ c[keccak_row_lane(l, 4)] ^ keccak_rotl_var(c[keccak_row_lane(l, 1)], 1), keccak_rho_rot(l));
            for (unsigned l = 0; l < 25; l++) bb[l] = tt[keccak_pi_src(l)];
            for (unsigned l = 0; l < 25; l++) a[l] = bb[l] ^ (~bb[keccak_row_lane(l, 1)] & bb[keccak_row_lane(l, 2)]);
            a[0] ^= keccak_round_constant(round);
        }
    }
    out[0] = a[0]; out[1] = a[1]; out[2] = a[2]; out[3] = a[3];
}
#endif

#if defined(__CUDACC__)
// One warp, one transcript: every lane passes the same FsTranscript; the digest is returned in all lanes.
__device__ __forceinline__ void fs_sha3_256_warp(const FsTranscript& t, kw64 (&out)[4]) {
    constexpr unsigned kFull = 0xffffffffu;
    const unsigned lane = threadIdx.x & 31u;
    const unsigned l = lane < 25u ? lane : 0u;           // idle lanes mirror lane 0's sources; their values are never read
    const unsigned c1 = keccak_col_lane(l, 1), c2 = keccak_col_lane(l, 2), c3 = keccak_col_lane(l, 3), c4 = keccak_col_lane(l, 4);
    const unsigned xm1 = keccak_row_lane(l, 4), xp1 = keccak_row_lane(l, 1), xp2 = keccak_row_lane(l, 2);
    const unsigned rot = keccak_rho_rot(l), src = keccak_pi_src(l);
    kw64 a = 0;
    const kw64 nblocks = fs_nblocks(t);
    kw64 in = lane < 17u ? fs_block_lane(t, 0, (int)lane) : 0ull;
    for (kw64 b = 0; b < nblocks; b++) {
        if (b == nblocks - 1 && lane == 16u) in ^= 0x8000000000000000ULL;
        a ^= in;
        // the next block's words are requested before the 24 rounds of this one: with one warp per scheduler nothing
        // else would hide the load latency
        if (b + 1 < nblocks) in = lane < 17u ? fs_block_lane(t, b + 1, (int)lane) : 0ull;
#pragma unroll 1
        for (int round = 0; round < 24; round++) {
            const kw64 c = a ^ __shfl_sync(kFull, a, c1) ^ __shfl_sync(kFull, a, c2) ^ __shfl_sync(kFull, a, c3) ^ __shfl_sync(kFull, a, c4);
            const kw64 d = __shfl_sync(kFull, c, xm1) ^ keccak_rotl_var(__shfl_sync(kFull, c, xp1), 1u);
            const kw64 bb = __shfl_sync(kFull, keccak_rotl_var(a ^ d, rot), src);
            const kw64 b1 = __shfl_sync(kFull, bb, xp1), b2 = __shfl_sync(kFull, bb, xp2);
            a = bb ^ (~b1 & b2) ^ (lane == 0u ? keccak_round_constant(round) : 0ull);     // iota, no divergence
        }
    }
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = __shfl_sync(kFull, a, i);
}
#endif

}  // namespace lsr
