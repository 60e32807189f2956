// lsr_arith.cuh -- modular arithmetic on u64 residues for sm_100a.
//
// B200 has no 64-bit integer multiplier: everything here is spelled in 32-bit
// IMAD (mad.lo / mad.hi / mad.wide), the pipe that bounds the NTT.  Counts in
// comments are IMAD-pipe instructions per call.
#pragma once
#include "lsr_common.h"

namespace lsr {

__device__ __forceinline__ u32 lo32(u64 x) { return (u32)x; }
__device__ __forceinline__ u32 hi32(u64 x) { return (u32)(x >> 32); }
__device__ __forceinline__ u64 pack64(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }

__device__ __forceinline__ u64 mad_wide(u32 a, u32 b, u64 c) {
    u64 d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}
__device__ __forceinline__ u64 mul_wide(u32 a, u32 b) {
    u64 d;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ u32 mad_lo(u32 a, u32 b, u32 c) {
    u32 d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// low 64 bits of  x*w + y*z   (6 IMAD, no carry chains: the two wide products
// accumulate in one 64-bit register pair, the four cross terms only touch the
// high word)
__device__ __forceinline__ u64 mullo2_acc(u64 x, u64 w, u64 y, u64 z) {
    u64 acc = mul_wide(lo32(x), lo32(w));
    acc = mad_wide(lo32(y), lo32(z), acc);
    u32 h = hi32(acc);
    h = mad_lo(lo32(x), hi32(w), h);
    h = mad_lo(hi32(x), lo32(w), h);
    h = mad_lo(lo32(y), hi32(z), h);
    h = mad_lo(hi32(y), lo32(z), h);
    return pack64(lo32(acc), h);
}

// Truncated Shoup multiplication: x * w mod q, lazily.
//   ws = floor(w * 2^64 / q), nq = 2^64 - q, any x < 2^64, w < q < 2^62.
// The quotient estimate drops the low partial product and the carries of the
// two cross products, so it undershoots floor(x*ws/2^64) by at most 2; the
// exact Shoup remainder is < 2q, hence the result is in [0, 4q).
// 9 IMAD + 2 ALU.
//
// Measured on B200 (tools/imad_microbench.cu): IMAD / IMAD.WIDE issue at 64
// lanes/clk/SM, IMAD.HI at 32.  The two cross products are therefore taken as
// full-rate mul.wide results whose high registers are added on the ALU pipe
// (one 3-input IADD3 + one IADD3.X) instead of two half-rate mul.hi.
__device__ __forceinline__ u64 mulred4(u64 x, u64 w, u64 ws, u64 nq) {
    const u32 x0 = lo32(x), x1 = hi32(x);
    const u32 s0 = lo32(ws), s1 = hi32(ws);
    const u64 ta = mul_wide(x1, s0);
    const u64 tb = mul_wide(x0, s1);
    // the 33-bit sum of the two high halves is formed first so that it is born
    // as a register pair and can ride in IMAD.WIDE's 64-bit addend without the
    // zero-extension moves ptxas otherwise inserts (they land on the IMAD pipe)
    const u64 ab = (u64)hi32(ta) + (u64)hi32(tb);
    const u64 qh = mad_wide(x1, s1, ab);
    return mullo2_acc(x, w, qh, nq);
}

// Exact Shoup multiplication (SEAL multiply_uint_mod_lazy): result in [0, 2q).
__device__ __forceinline__ u64 mulred2(u64 x, u64 w, u64 ws, u64 nq) {
    const u64 qh = __umul64hi(x, ws);
    return mullo2_acc(x, w, qh, nq);
}

// x >= m ? x - m : x, written so the subtraction's borrow is the predicate
__device__ __forceinline__ u64 csub(u64 x, u64 m) {
    const u64 t = x - m;
    return x < m ? x : t;
}

// v < 2^7 * q  ->  [0, q).   3 IMAD + shifts + one conditional subtract.
__device__ __forceinline__ u64 reduce_small(u64 v, const ModParams& mp) {
    const u32 vs = (u32)(v >> mp.red_sh);
    const u32 est = hi32(mul_wide(vs, mp.red_c));     // floor(v/q) - 1 <= est <= floor(v/q)
    u64 acc = mad_wide(est, lo32(mp.nq), v);          // v - est*q  (mod 2^64)
    const u32 h = mad_lo(est, hi32(mp.nq), hi32(acc));
    return csub(pack64(lo32(acc), h), mp.q);
}

// Exact reduction of a 128-bit value (SEAL barrett_reduce_128, used by
// multiply_uint_mod): (z1:z0) mod q for any z, q < 2^61.
__device__ __forceinline__ u64 barrett128(u64 z0, u64 z1, const ModParams& mp) {
    // tmp3 = floor(z * ratio / 2^128) up to an error of 1
    const u64 c1 = __umul64hi(z0, mp.bar_lo);
    const u64 t2lo = z0 * mp.bar_hi;
    const u64 t2hi = __umul64hi(z0, mp.bar_hi);
    const u64 s1 = t2lo + c1;
    const u64 carry1 = s1 < t2lo;
    const u64 t3lo = z1 * mp.bar_lo;
    const u64 t3hi = __umul64hi(z1, mp.bar_lo);
    const u64 s2 = s1 + t3lo;
    const u64 carry2 = s2 < s1;
    const u64 qhat = z1 * mp.bar_hi + t2hi + carry1 + t3hi + carry2;
    const u64 r = z0 - qhat * mp.q;
    return csub(r, mp.q);
}

// ---------------------------------------------------------------------------
// POL_GOLD: q = 2^64 - 2^32 + 1.  Residues are canonical u64 in [0, q) at every interface and throughout the inverse
// transforms; inside the forward transforms sums are any representative in [0, 2^64) (gold_add_lazy).  Sums that
// overflow 2^64 are fixed up with 2^64 = 2^32 - 1 (mod q), i.e. a wrapped subtraction of q.
// ---------------------------------------------------------------------------
// Canonical residues in, canonical residue out.  Every correction is a carry / borrow folded back with
// 2^64 = 2^32 - 1 (mod q): the borrow of a subtraction chain turned into the mask 0xffffffff (= eps) by
// `subc m, 0, 0`, the carry of an addition chain by `addc` -- no 64-bit compare / select pairs, which is what
// the ALU-pipe-bound Goldilocks butterflies were mostly made of (12 ISETP + 10 SEL of 48 instructions).
// Chains are homogeneous (sub.cc ... subc, add.cc ... addc): the two flag conventions are never mixed.
__device__ __forceinline__ u64 gold_sub(u64 a, u64 b) {
    // d = a - b; on borrow the wrapped value is 2^64 = eps too large
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t.reg .u32 m;\n\t"
        "sub.cc.u32 %0, %0, %2;\n\tsubc.cc.u32 %1, %1, %3;\n\tsubc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\tsubc.u32 %1, %1, 0;\n\t}"
        : "+r"(a0), "+r"(a1) : "r"(b0), "r"(b1));
    return ((u64)a1 << 32) | a0;
}
__device__ __forceinline__ u64 gold_neg_raw(u64 b) { return kGoldilocks - b; }      // in [1, q] for canonical b
// a + b = a - (q - b); q - b = q for b = 0 is handled by the same borrow correction
__device__ __forceinline__ u64 gold_add(u64 a, u64 b) { return gold_sub(a, gold_neg_raw(b)); }

// a + b for a ANYWHERE in [0, 2^64) and b <= q: on carry the wrapped sum is 2^64 = eps too small; adding eps cannot carry
// again (a + b - 2^64 <= q - 2).  The result is some representative in [0, 2^64), not necessarily canonical: the forward
// (Cooley-Tukey) butterflies keep their sums this way -- X +- T with T canonical never needs X canonical, and products accept
// any 64-bit operand -- and only the last pass canonicalises (gold_canonical).  gold_sub(a, b) has the same property: the
// borrow correction a - b + 2^64 - eps cannot borrow again when b <= q, whatever a is.
__device__ __forceinline__ u64 gold_add_lazy(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t.reg .u32 m;\n\t"
        "add.cc.u32 %0, %0, %2;\n\taddc.cc.u32 %1, %1, %3;\n\taddc.u32 m, 0, 0;\n\tneg.s32 m, m;\n\t"
        "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t}"
        : "+r"(a0), "+r"(a1) : "r"(b0), "r"(b1));
    return ((u64)a1 << 32) | a0;
}
#ifndef LSR_GOLD_CANON_PRED
#define LSR_GOLD_CANON_PRED 1
#endif
// r >= q  <=>  high word all ones and low word non-zero (q = 0xffffffff00000001); then r - q = low word - 1
// (as carries, not compares: the butterfly networks keep every predicate register busy with carry chains, and
// compare results that live across them get spilled into general registers)
__device__ __forceinline__ u64 gold_canonical(u64 r) {
    u32 r0 = (u32)r, r1 = (u32)(r >> 32);
#if LSR_GOLD_CANON_PRED
    if (r1 == 0xffffffffu && r0 != 0u) { r0 -= 1u; r1 = 0u; }
#else
    // r >= q  <=>  r + eps carries; then r - q = r + eps (mod 2^64)
    asm("{\n\t.reg .u32 m, s0, s1;\n\t"
        "add.cc.u32 s0, %0, 0xffffffff;\n\taddc.cc.u32 s1, %1, 0;\n\taddc.u32 m, 0, 0;\n\tneg.s32 m, m;\n\t"
        "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t}"
        : "+r"(r0), "+r"(r1));
#endif
    return ((u64)r1 << 32) | r0;
}

// (a * b) mod q for ANY 64-bit a, b, canonical result.  Product: four 32 x 32 -> 64 multiplies whose halves are summed
// column-wise with two carry chains (4 IMAD.WIDE + 5 IADD3; the earlier form with 64-bit accumulators cost 12 instructions,
// half of them moves that set up the accumulator register pairs).  Reduction of (r3 r2 r1 r0), 2^64 = eps, 2^96 = -1:
// (r1 r0) - r3 + r2 * eps; the borrow of the subtraction is folded back as - eps (cannot borrow again:
// lo - r3 + 2^64 >= 2^64 - eps), the carry of the addition as + eps (cannot carry again: the sum is then < q).
__device__ __forceinline__ u64 gold_mul(u64 a, u64 b) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u32 r0, r1, r2, r3;
    asm("{\n\t.reg .u64 p0, p1, p2, p3;\n\t.reg .u32 p0l, p0h, p1l, p1h, p2l, p2h, p3l, p3h;\n\t"
        "mul.wide.u32 p0, %4, %6;\n\tmul.wide.u32 p1, %4, %7;\n\tmul.wide.u32 p2, %5, %6;\n\tmul.wide.u32 p3, %5, %7;\n\t"
        "mov.b64 {p0l, p0h}, p0;\n\tmov.b64 {p1l, p1h}, p1;\n\tmov.b64 {p2l, p2h}, p2;\n\tmov.b64 {p3l, p3h}, p3;\n\t"
        "mov.u32 %0, p0l;\n\t"
        "add.cc.u32 %1, p0h, p1l;\n\taddc.cc.u32 %2, p1h, p3l;\n\taddc.u32 %3, p3h, 0;\n\t"
        "add.cc.u32 %1, %1, p2l;\n\taddc.cc.u32 %2, %2, p2h;\n\taddc.u32 %3, %3, 0;\n\t}"
        : "=r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    asm("{\n\t.reg .u32 m;\n\t"
        "sub.cc.u32 %0, %0, %2;\n\tsubc.cc.u32 %1, %1, 0;\n\tsubc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\tsubc.u32 %1, %1, 0;\n\t}"
        : "+r"(r0), "+r"(r1) : "r"(r3));
    asm("{\n\t.reg .u32 m, t0, t1;\n\t.reg .u64 t;\n\t"
        "mul.wide.u32 t, %2, 0xffffffff;\n\tmov.b64 {t0, t1}, t;\n\t"
        "add.cc.u32 %0, %0, t0;\n\taddc.cc.u32 %1, %1, t1;\n\taddc.u32 m, 0, 0;\n\tneg.s32 m, m;\n\t"
        "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t}"
        : "+r"(r0), "+r"(r1) : "r"(r2));
    return gold_canonical(((u64)r1 << 32) | r0);
}

// exact (a * b) mod q for any u64 a, b  (ntt.cpp:116-118 -> multiply_uint_mod)
__device__ __forceinline__ u64 mulmod_exact(u64 a, u64 b, const ModParams& mp) {
    if (mp.gold) return gold_mul(a, b);       // any 64-bit operands
    return barrett128(a * b, __umul64hi(a, b), mp);
}

// exact x mod q for any u64 x
__device__ __forceinline__ u64 reduce64(u64 x, const ModParams& mp) {
    if (mp.gold) return x >= kGoldilocks ? x - kGoldilocks : x;
    return barrett128(x, 0, mp);
}

// Input sanitiser: SEAL documents [0, 4q) inputs for the lazy forward
// transform; anything larger is reduced exactly (rare, warp-divergent branch)
// so that every u64 input has a defined result: NTT(x mod q).
__device__ __forceinline__ u64 sanitize(u64 x, u64 limit, const ModParams& mp) {
    if (__builtin_expect(x >= limit, 0)) x = reduce64(x, mp);
    return x;
}

// ---------------------------------------------------------------------------
// POL_F64: residues as doubles.  Every value is an integer-valued double with
// |x| < 2^51; products are formed exactly with the fma error-free
// transformation, so the results are exact integers mod q whatever the
// rounding of the quotient estimate does -- the estimate only decides which
// representative comes out.  Nothing here is approximate.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double as_d(u64 x) { return __longlong_as_double((long long)x); }
__device__ __forceinline__ u64 as_u(double x) { return (u64)__double_as_longlong(x); }

constexpr double kRintMagic = 6755399441055744.0;   // 1.5 * 2^52: (v + M) - M = rint(v) for |v| < 2^51
constexpr double kTwo52 = 4503599627370496.0;
constexpr u64 kTwo52Bits = 0x4330000000000000ull;   // bit pattern of 2^52

// x * w - c * q with c = rint(x * wq), wq = RN(w / q), 0 <= w < q < 2^45, |x| < 2^51.
//   c  : (x*wq + M) is computed exactly and rounded to the unit grid, so |c - x*wq| <= 1/2 and
//        |c - x*w/q| <= 1/2 + |x| 2^-53
//   h,l: h + l = x*w exactly (l = fma(x, w, -h) is the rounding error of the product)
//   r  : h is an integer (a product of integers, rounded to a coarser grid when >= 2^53) and
//        |h - c*q| < 2^53, so the fma is exact; r + l = x*w - c*q is an integer below q in
//        magnitude, exact again.
// Result: balanced representative, |result| <= q (1/2 + |x| 2^-53) <= 0.75 q.   6 FP64 instructions.
// c = an integer within 3/4 of x * wq.  Two formulations: the magic-number pair DFMA + DADD (two FP64-pipe instructions, c =
// nearest integer to the exact x * wq), or -- XR -- DMUL + FRND.F64 (one FP64-pipe instruction + one conversion-pipe
// instruction; the product is rounded to 53 bits first: for |x| < 2^51 that adds at most 2^-3 to the 1/2, and with the
// rounding of wq itself, |x| 2^-54 <= 2^-3, c stays within 3/4 of x * w / q -- the bound everything below assumes).
// Results are exact either way; the two can pick different representatives only on a tie, and outputs are canonical.
// Measured (tools/variant_bench.py): FRND issues in one cycle where DADD holds the port for two, +3.4 % on the stand-alone
// forward transform (61.8 -> 63.9 M NTT/s at n = 4096), but its latency sits in the c -> r chain: +0.6 % on the inverse,
// -1.4 % on the fused commitment.  So XR is a template flag and only the stand-alone forward kernels set it.
template <bool XR = false>
__device__ __forceinline__ double rint_prod(double x, double wq) {
    if constexpr (XR) {
        double c;
        asm("cvt.rni.f64.f64 %0, %1;" : "=d"(c) : "d"(__dmul_rn(x, wq)));
        return c;
    } else {
        return __dadd_rn(__fma_rn(x, wq, kRintMagic), -kRintMagic);
    }
}

template <bool XR = false>
__device__ __forceinline__ double mulmod_f(double x, double w, double wq, double q) {
    const double c = rint_prod<XR>(x, wq);
    const double h = __dmul_rn(x, w);
    const double l = __fma_rn(x, w, -h);
    const double r = __fma_rn(-c, q, h);
    return __dadd_rn(r, l);
}

// |x| < 2^51  ->  balanced representative, |result| <= q (1/2 + |x| 2^-53).   3 FP64 instructions.
template <bool XR = false>
__device__ __forceinline__ double reduce_f(double x, double invq, double q) {
    const double c = rint_prod<XR>(x, invq);
    return __fma_rn(-c, q, x);
}

// integer x < 2^52 -> double, without the (quarter-rate) conversion instruction
__device__ __forceinline__ double u64_to_f(u64 x) { return __dadd_rn(as_d(x | kTwo52Bits), -kTwo52); }

// integer-valued r in (-q, q) -> canonical residue in [0, q) as u64.  The comparison (not the
// sign bit) decides, so a negative zero cannot come out as q.
__device__ __forceinline__ u64 f_to_canonical(double r, const ModParams& mp) {
    const double off = r < 0.0 ? mp.q52 : kTwo52;
    return as_u(__dadd_rn(r, off)) & 0x000fffffffffffffull;
}

// m mod p for any u64 m and a small modulus p < 2^21 (the plaintext modulus), pinv = floor((2^64-1)/p):
// the quotient estimate is at most 2 short, so the 32-bit remainder is below 3p; min(r, r - p) is the
// branch-free conditional subtract (r - p wraps to a huge value when r < p).
__device__ __forceinline__ u32 mod_small(u64 m, u32 p, u64 pinv) {
    const u64 qh = __umul64hi(m, pinv);
    u32 r = (u32)m - (u32)qh * p;
    r = min(r, r - p);
    r = min(r, r - p);
    return r;
}

// floor(v / d) for any u64 v and a divisor 2 <= d < 2^63, dinv = floor((2^64-1)/d): the estimate
// umulhi(v, dinv) is at most 1 short (v/d - v*dinv/2^64 <= v/2^64 < 1), two fix-ups for good measure.
// Digit planes of the commitment messages: d = p^l.
__device__ __forceinline__ u64 div_small(u64 v, u64 d, u64 dinv) {
    u64 qh = __umul64hi(v, dinv);
    u64 r = v - qh * d;
    if (r >= d) { r -= d; ++qh; }
    if (r >= d) { ++qh; }
    return qh;
}

// modular add/sub on canonical residues
__device__ __forceinline__ u64 addmod(u64 a, u64 b, u64 q) { return csub(a + b, q); }
__device__ __forceinline__ u64 submod(u64 a, u64 b, u64 q) { return a >= b ? a - b : a + q - b; }

// field operations on canonical residues for any supported modulus (q < 2^61, or Goldilocks)
__device__ __forceinline__ u64 field_add(u64 a, u64 b, const ModParams& mp) { return mp.gold ? gold_add(a, b) : addmod(a, b, mp.q); }
__device__ __forceinline__ u64 field_sub(u64 a, u64 b, const ModParams& mp) { return mp.gold ? gold_sub(a, b) : submod(a, b, mp.q); }
__device__ __forceinline__ u64 field_mul(u64 a, u64 b, const ModParams& mp) { return mulmod_exact(a, b, mp); }

}  // namespace lsr
