// lsr_sampler.cuh -- device side of K7: ChaCha keystream + CDT discrete Gaussian.
//
// Replaces the reference's sample_gaussian / sample_single
// (cpp-core/src/utils.cpp:95-146).  The CDT itself is built on the host with
// the reference's long-double code (lsr_host.cpp build_cdt) and only evaluated
// here.  Entropy: std::random_device (utils.cpp:138) becomes a keyed ChaCha
// stream so CPU oracle and GPU agree bit for bit and results do not depend on
// how a batch is sharded.
//
// Constant time: like the reference's linear scan, the sample is computed as
// the COUNT of table entries below u (the table is non-decreasing and ends at
// 2^64-1, so the count equals the reference's "first k with cdf[k] >= u"), with
// no data-dependent branch or address.
#pragma once
#include "lsr_common.h"

namespace lsr {

__device__ __forceinline__ u32 rotl32(u32 v, int c) { return __funnelshift_l(v, v, c); }

#define LSR_QR(a, b, c, d)                       \
    a += b; d ^= a; d = rotl32(d, 16);           \
    c += d; b ^= c; b = rotl32(b, 12);           \
    a += b; d ^= a; d = rotl32(d, 8);            \
    c += d; b ^= c; b = rotl32(b, 7);

struct ChaChaKey { u32 k[8]; };

// one ChaCha block, kChaChaRounds rounds, output x[16]
__device__ __forceinline__ void chacha_block(const ChaChaKey& key, u32 w12, u32 w13, u32 w14,
                                             u32 w15, u32 (&x)[16]) {
    x[0] = 0x61707865u; x[1] = 0x3320646eu; x[2] = 0x79622d32u; x[3] = 0x6b206574u;
#pragma unroll
    for (int i = 0; i < 8; i++) x[4 + i] = key.k[i];
    x[12] = w12; x[13] = w13; x[14] = w14; x[15] = w15;
#pragma unroll
    for (int r = 0; r < kChaChaRounds; r += 2) {
        LSR_QR(x[0], x[4], x[8], x[12])  LSR_QR(x[1], x[5], x[9], x[13])
        LSR_QR(x[2], x[6], x[10], x[14]) LSR_QR(x[3], x[7], x[11], x[15])
        LSR_QR(x[0], x[5], x[10], x[15]) LSR_QR(x[1], x[6], x[11], x[12])
        LSR_QR(x[2], x[7], x[8], x[13])  LSR_QR(x[3], x[4], x[9], x[14])
    }
    x[0] += 0x61707865u; x[1] += 0x3320646eu; x[2] += 0x79622d32u; x[3] += 0x6b206574u;
#pragma unroll
    for (int i = 0; i < 8; i++) x[4 + i] += key.k[i];
    x[12] += w12; x[13] += w13; x[14] += w14; x[15] += w15;
}

// CDT table passed by value: lives in the kernel-parameter constant bank, so
// with the loops unrolled every entry is an immediate constant operand.
constexpr int kCdtInline = 64;
struct CdtParam {
    u32 count;                 // entries in use (<= kCdtInline); trailing 2^64-1 entries trimmed
    u32 pad;                   // 1 if every entry with index >= 31 has high word 0xffffffff (32-bit tail scan)
    u64 cdf[kCdtInline];       // unused entries = 2^64-1 (never "below u")
    // Compact form (build_cdt_param): the DISTINCT table values below 2^64-1 in ascending order with the
    // number of table entries they stand for.  The reference's long-double table saturates (sigma = 3.19:
    // cdf[29..38] are all 2^64-2), so 39 entries are 30 distinct values and the whole table fits one value
    // per lane: no tail scan at all.  #{k : cdf[k] < u} = dcum[#{i : dval[i] < u}].
    u32 compact;               // 1 if there are at most 31 distinct values
    u32 pad2;
    u64 dval[32];              // unused = 2^64-1; dval[31] is always 2^64-1
    u32 dcum[32];              // dcum[i] = number of table entries < dval[i] (= entries counted by dval[0..i-1])
    u32 fast[32];              // (dval[i] >> 39) << 7 | dcum[i]: 25-bit prefix + count in one word (cdt_fast_probe)
};

// Magnitude draw of a commitment sample: top 31 bits from the sample's own keystream word w
// (whose bit 0 is the sign draw), low 33 bits from the refinement block (bit 0 of f0, all of f1).
__host__ __device__ __forceinline__ u64 commit_draw(u32 w, u32 f0, u32 f1) {
    return ((u64)(w >> 1) << 33) | ((u64)(f0 & 1u) << 32) | (u64)f1;
}

// magnitude = #{k : cdf[k] < u}; NCH8 = ceil(count / 8)
template <int NCH8>
__device__ __forceinline__ u32 cdt_magnitude(const CdtParam& t, u64 u) {
    u32 below = 0;
#pragma unroll
    for (int k = 0; k < NCH8 * 8; k++) below += (u32)(t.cdf[k] < u);
    return below;
}

// Same count, ~3x fewer instructions: entries 0..30 live one per lane
// (lane_entry = cdf[lane], lane 31 unused) and are searched with a 5-step
// branch-free binary search whose probes are warp shuffles -- no memory access
// and no branch depends on u, so the constant-time property of the linear scan
// is kept.  Entries 31.. (mass < 2^-60 for sigma = 3.19) are scanned linearly
// from the constant bank.  Exactly equal to cdt_magnitude because the table is
// non-decreasing: #{k : cdf[k] < u} = #{k < 31 : ...} + #{k >= 31 : ...}.
// Must be called by all 32 lanes of a warp.
template <int NCH8>
__device__ __forceinline__ u32 cdt_magnitude_shfl(const CdtParam& t, u64 lane_entry, u64 u) {
    const u32 e_lo = (u32)lane_entry, e_hi = (u32)(lane_entry >> 32);
    // the search tracks the next probe index directly: probe' = probe + (below ? step/2 : -step/2)
    // is one select between two immediates and one add, against select + add + add for pos / probe
    u32 probe = (t.cdf[15] < u) ? 23u : 7u;        // first probe (entry 15) is the same for every lane
#pragma unroll
    for (int step = 8; step >= 2; step >>= 1) {
        const u32 v_lo = __shfl_sync(0xffffffffu, e_lo, probe);
        const u32 v_hi = __shfl_sync(0xffffffffu, e_hi, probe);
        const u64 v = ((u64)v_hi << 32) | v_lo;
        probe += (v < u) ? (u32)(step / 2) : (u32)(-(step / 2));
    }
    u32 pos;
    {
        const u32 v_lo = __shfl_sync(0xffffffffu, e_lo, probe);
        const u32 v_hi = __shfl_sync(0xffffffffu, e_hi, probe);
        const u64 v = ((u64)v_hi << 32) | v_lo;
        pos = probe + ((v < u) ? 1u : 0u);         // #{k < 31 : cdf[k] < u}
    }
    if (t.pad) {
        // every entry from 31 on has an all-ones high word (true for sigma = 3.19: 1 - cdf[31] < 2^-32), so
        // cdf[k] < u  <=>  u_hi == 0xffffffff and cdf_lo[k] < u_lo.  Each 32-bit comparison is the borrow of a
        // subtraction, accumulated with subc: two instructions per entry, one gate at the end.
        // t.pad is a property of the table (uniform over the grid), not of u.
        const u32 u_lo = (u32)u, u_hi = (u32)(u >> 32);
        u32 neg = 0;                                // minus the number of tail entries below u_lo
#pragma unroll
        for (int k = 31; k < NCH8 * 8; k++) {
            u32 scratch;
            asm("{sub.cc.u32 %1, %2, %3; subc.u32 %0, %0, 0;}" : "+r"(neg), "=r"(scratch) : "r"((u32)t.cdf[k]), "r"(u_lo));
        }
        pos += (u_hi == 0xffffffffu) ? (0u - neg) : 0u;
    } else {
#pragma unroll
        for (int k = 31; k < NCH8 * 8; k++) pos += (u32)(t.cdf[k] < u);
    }
    return pos;
}

// Compact search (CdtParam::compact): 31 distinct values, one per lane (lane_val = dval[lane]), and the
// cumulative entry counts (lane_cum = dcum[lane]).  Five 64-bit comparisons, each the carry of
// u + ~v (= [v < u]; an add.cc / addc.cc pair on the complemented table value, so only add-type carries
// are chained) shifted into the position with addc (pos = 2 pos + carry): a step is two shuffles, two
// ALU-pipe additions, one FMA-pipe add-with-carry and the probe index -- the sampler phase is bound by
// the ALU pipe (2 cycles per warp instruction), which the compare / select / add formulation of
// cdt_magnitude_shfl loads twice as much.  The first probe (dval[15]) is the same for every lane and
// comes from the constant bank; the last shuffle turns the position into the entry count.  No branch
// and no address depends on u.  All 32 lanes must call.
__device__ __forceinline__ u32 cdt_magnitude_compact(u64 mid, u64 lane_val, u32 lane_cum, u64 u) {
    const u32 e_lo = ~(u32)lane_val, e_hi = ~(u32)(lane_val >> 32);
    const u32 u_lo = (u32)u, u_hi = (u32)(u >> 32);
    u32 p;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %3;\n\taddc.cc.u32 t, %2, %4;\n\taddc.u32 %0, 0, 0;\n\t}"
        : "=r"(p) : "r"(~(u32)mid), "r"(~(u32)(mid >> 32)), "r"(u_lo), "r"(u_hi));
#pragma unroll
    for (int i = 1; i < 5; i++) {
        const u32 probe = p * (1u << (5 - i)) + ((1u << (4 - i)) - 1u);       // pos + half - 1, pos = p << (5 - i)
        const u32 v_lo = __shfl_sync(0xffffffffu, e_lo, probe);
        const u32 v_hi = __shfl_sync(0xffffffffu, e_hi, probe);
        asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %3;\n\taddc.cc.u32 t, %2, %4;\n\taddc.u32 %0, %0, %0;\n\t}"
            : "+r"(p) : "r"(v_lo), "r"(v_hi), "r"(u_lo), "r"(u_hi));
    }
    return __shfl_sync(0xffffffffu, lane_cum, p);                             // p = #{i < 31 : dval[i] < u}
}

// Fast form of the compact search for the commitment sampler: the top 25 bits of the draw (= w >> 7) against
// the 25-bit prefixes of the distinct table values, prefix and entry count packed in ONE word per lane
// (lane_e = ~fast[lane], complemented for the carry trick), so a step is one shuffle, one add.cc and one addc.
// [fast[i] < (w & ~0x7f)]  <=>  prefix[i] < w >> 7  (the count occupies the low 7 bits, which are cleared in the
// key).  Returns e = ~fast[p] with p = #{i : prefix[i] < w >> 7}:
//   magnitude = (e & 0x7f) ^ 0x7f      -- valid iff no table prefix equals w >> 7;
//   tie       = ((e ^ w) | 0x7f) == 0xffffffff   (prefix[p] == w >> 7: the low bits of the draw decide; the
//               caller then forms the full 64-bit draw and runs cdt_magnitude_compact).
// The sentinel lanes (value 2^64-1, prefix 0x1ffffff) make the all-ones prefix a tie, which is correct if
// conservative.  Probability of a tie: (#distinct prefixes) * 2^-25 per sample.  All 32 lanes must call.
__device__ __forceinline__ u32 cdt_fast_probe(u32 mid_e, u32 lane_e, u32 w) {
    const u32 key = w & ~0x7fu;
    u32 p;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, 0, 0;\n\t}" : "=r"(p) : "r"(mid_e), "r"(key));
#pragma unroll
    for (int i = 1; i < 5; i++) {
        const u32 probe = p * (1u << (5 - i)) + ((1u << (4 - i)) - 1u);
        const u32 v = __shfl_sync(0xffffffffu, lane_e, probe);
        asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %1, %2;\n\taddc.u32 %0, %0, %0;\n\t}" : "+r"(p) : "r"(v), "r"(key));
    }
    return __shfl_sync(0xffffffffu, lane_e, p);
}

// table in global memory, any size
__device__ __forceinline__ u32 cdt_magnitude_global(const u64* __restrict__ cdf, u32 count, u64 u) {
    u32 below = 0;
    for (u32 k = 0; k < count; k++) below += (u32)(__ldg(cdf + k) < u);
    return below;
}

// signed sample as a residue mod q: sign applies only when magnitude != 0
__device__ __forceinline__ u64 signed_residue(u32 mag, u32 sign_bit, u64 q) {
    const u64 neg = q - (u64)mag;
    return (sign_bit & (u32)(mag != 0)) ? neg : (u64)mag;
}
__device__ __forceinline__ long long signed_value(u32 mag, u32 sign_bit) {
    const long long m = (long long)mag;
    return (sign_bit & (u32)(mag != 0)) ? -m : m;
}

}  // namespace lsr
