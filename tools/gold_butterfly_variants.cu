// tools/gold_butterfly_variants.cu -- one Goldilocks butterfly (T = Y * w; X + T; X - T) under the formulations that were
// weighed for csrc/lsr_arith.cuh; compile one at a time and count the SASS:
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -DVAR=<n> -cubin -o v.cubin tools/gold_butterfly_variants.cu
//   cuobjdump -sass v.cubin | grep -cP "^\s+/\*[0-9a-f]{4}\*/"
// Instructions of the whole kernel (16 of them are address arithmetic, loads and stores), nvcc 12.9:
//   the form before this study (64-bit accumulators in the product, carry-chain canonicalisation)        66
//   VAR=2 predicated corrections: ptxas turns the flag into a mask and back (ISETP + SEL)          63
//   VAR=3 masks by `subc m,0,0` after an ADD chain: 52, but ptxas emits m = carry - 1 (it keeps the SASS carry
//         convention of the subtraction), i.e. the correction fires on NO carry -- mixing the two flag
//         conventions is mistranslated, so this form is WRONG on the device and only listed as a warning
//   VAR=4 carry folded in by mad.wide(c, 0xffffffff, s): ptxas splits the constant multiply (IMAD.HI + moves)    57
//   VAR=5 inverted mask (addc m, 0xffffffff, 0; not): the `not` is not folded into the IADD3 operand            56
//   VAR=6 column-sum product + homogeneous chains + compare-based canonicalisation (adopted)                     56
// Inside the radix-16 networks the adopted form gives 44.5 (forward, lazy sums) / 48.7 (inverse) instructions per
// butterfly against 54.7 / 50.0 before; a carry-chain canonicalisation instead of the compare costs 2 - 5 % more there
// even though the compares spill predicates into a register (measured: 532 vs 487 G butterflies/s forward at 2^20).
#include <cstdint>
typedef unsigned long long u64; typedef unsigned int u32;
constexpr u64 Q = 0xFFFFFFFF00000001ull;
#ifndef VAR
#define VAR 6
#endif
#if VAR == 3
// wrapped difference / sum fixes with masks taken straight from the flag (subc m = -CF)
__device__ __forceinline__ u64 gold_sub(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t.reg .u32 m;\n\t"
        "sub.cc.u32 %0, %0, %2;\n\tsubc.cc.u32 %1, %1, %3;\n\tsubc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\tsubc.u32 %1, %1, 0;\n\t}"
        : "+r"(a0), "+r"(a1) : "r"(b0), "r"(b1));
    return ((u64)a1 << 32) | a0;
}
// a arbitrary in [0, 2^64), b <= q: a + b, carry -> += eps (cannot carry again)
__device__ __forceinline__ u64 gold_add_lazy(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t.reg .u32 m;\n\t"
        "add.cc.u32 %0, %0, %2;\n\taddc.cc.u32 %1, %1, %3;\n\tsubc.u32 m, 0, 0;\n\t"
        "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t}"
        : "+r"(a0), "+r"(a1) : "r"(b0), "r"(b1));
    return ((u64)a1 << 32) | a0;
}
__device__ __forceinline__ u64 gold_add(u64 a, u64 b) { return gold_add_lazy(a, b); }
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
        "add.cc.u32 %0, %0, t0;\n\taddc.cc.u32 %1, %1, t1;\n\tsubc.u32 m, 0, 0;\n\t"
        "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t}"
        : "+r"(r0), "+r"(r1) : "r"(r2));
    // canonical: r >= q  <=>  r1 == 0xffffffff and r0 != 0; then r - q = r0 - 1
    if (r1 == 0xffffffffu && r0 != 0u) { r0 -= 1u; r1 = 0u; }
    return ((u64)r1 << 32) | r0;
}
#elif VAR == 6
// wrapped difference / sum fixes with masks taken straight from the flag (subc m = -CF)
__device__ __forceinline__ u64 gold_sub(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t.reg .u32 m;\n\t"
        "sub.cc.u32 %0, %0, %2;\n\tsubc.cc.u32 %1, %1, %3;\n\tsubc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\tsubc.u32 %1, %1, 0;\n\t}"
        : "+r"(a0), "+r"(a1) : "r"(b0), "r"(b1));
    return ((u64)a1 << 32) | a0;
}
// a arbitrary in [0, 2^64), b <= q: a + b, carry -> += eps (cannot carry again)
__device__ __forceinline__ u64 gold_add_lazy(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t.reg .u32 m;\n\t"
        "add.cc.u32 %0, %0, %2;\n\taddc.cc.u32 %1, %1, %3;\n\taddc.u32 m, 0, 0;\n\tneg.s32 m, m;\n\t"
        "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t}"
        : "+r"(a0), "+r"(a1) : "r"(b0), "r"(b1));
    return ((u64)a1 << 32) | a0;
}
__device__ __forceinline__ u64 gold_add(u64 a, u64 b) { return gold_add_lazy(a, b); }
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
    // canonical: r >= q  <=>  r1 == 0xffffffff and r0 != 0; then r - q = r0 - 1
    if (r1 == 0xffffffffu && r0 != 0u) { r0 -= 1u; r1 = 0u; }
    return ((u64)r1 << 32) | r0;
}
#elif VAR == 5
// wrapped difference / sum fixes with masks taken straight from the flag (subc m = -CF)
__device__ __forceinline__ u64 gold_sub(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t.reg .u32 m;\n\t"
        "sub.cc.u32 %0, %0, %2;\n\tsubc.cc.u32 %1, %1, %3;\n\tsubc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\tsubc.u32 %1, %1, 0;\n\t}"
        : "+r"(a0), "+r"(a1) : "r"(b0), "r"(b1));
    return ((u64)a1 << 32) | a0;
}
// a arbitrary in [0, 2^64), b <= q: a + b, carry -> += eps (cannot carry again)
__device__ __forceinline__ u64 gold_add_lazy(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t.reg .u32 m;\n\t"
        "add.cc.u32 %0, %0, %2;\n\taddc.cc.u32 %1, %1, %3;\n\taddc.u32 m, 0xffffffff, 0;\n\tnot.b32 m, m;\n\t"
        "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t}"
        : "+r"(a0), "+r"(a1) : "r"(b0), "r"(b1));
    return ((u64)a1 << 32) | a0;
}
__device__ __forceinline__ u64 gold_add(u64 a, u64 b) { return gold_add_lazy(a, b); }
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
        "add.cc.u32 %0, %0, t0;\n\taddc.cc.u32 %1, %1, t1;\n\taddc.u32 m, 0xffffffff, 0;\n\tnot.b32 m, m;\n\t"
        "add.cc.u32 %0, %0, m;\n\taddc.u32 %1, %1, 0;\n\t}"
        : "+r"(r0), "+r"(r1) : "r"(r2));
    // canonical: r >= q  <=>  r1 == 0xffffffff and r0 != 0; then r - q = r0 - 1
    if (r1 == 0xffffffffu && r0 != 0u) { r0 -= 1u; r1 = 0u; }
    return ((u64)r1 << 32) | r0;
}
#elif VAR == 4
__device__ __forceinline__ u64 gold_sub(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t.reg .u32 m;\n\t"
        "sub.cc.u32 %0, %0, %2;\n\tsubc.cc.u32 %1, %1, %3;\n\tsubc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\tsubc.u32 %1, %1, 0;\n\t}"
        : "+r"(a0), "+r"(a1) : "r"(b0), "r"(b1));
    return ((u64)a1 << 32) | a0;
}
// a + b + (carry ? eps : 0): the carry (0/1) times eps goes in with one wide multiply-add
__device__ __forceinline__ u64 add_fold(u64 a, u64 b) {
    u64 r;
    asm("{\n\t.reg .u32 a0, a1, b0, b1, c;\n\t.reg .u64 s;\n\t"
        "mov.b64 {a0, a1}, %1;\n\tmov.b64 {b0, b1}, %2;\n\t"
        "add.cc.u32 a0, a0, b0;\n\taddc.cc.u32 a1, a1, b1;\n\taddc.u32 c, 0, 0;\n\t"
        "mov.b64 s, {a0, a1};\n\tmad.wide.u32 %0, c, 0xffffffff, s;\n\t}"
        : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 gold_add(u64 a, u64 b) { return add_fold(a, b); }
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
    const u64 r = add_fold(((u64)r1 << 32) | r0, (u64)r2 * 0xFFFFFFFFull);
    r0 = (u32)r; r1 = (u32)(r >> 32);
    if (r1 == 0xffffffffu && r0 != 0u) { r0 -= 1u; r1 = 0u; }
    return ((u64)r1 << 32) | r0;
}
#else
// ---- sub: a - b (canonical in, canonical out); borrow -> += q  (i.e. -= eps on the wrapped value)
__device__ __forceinline__ u64 gold_sub(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t.reg .pred p;\n\t.reg .u32 c;\n\t"
        "sub.cc.u32 %0, %0, %2;\n\tsubc.cc.u32 %1, %1, %3;\n\tsubc.u32 c, 0, 0;\n\t"
        "setp.ne.u32 p, c, 0;\n\t"
        "@p sub.cc.u32 %0, %0, 0xffffffff;\n\t@p subc.u32 %1, %1, 0;\n\t}"
        : "+r"(a0), "+r"(a1) : "r"(b0), "r"(b1));
    return ((u64)a1 << 32) | a0;
}
__device__ __forceinline__ u64 gold_add(u64 a, u64 b) { return gold_sub(a, Q - b); }
__device__ __forceinline__ u64 gold_mul(u64 a, u64 b) {
    const u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u32 r0, r1, r2, r3;
    // 4 wide products, then column sums with carry chains
    asm("{\n\t.reg .u64 p0, p1, p2, p3;\n\t.reg .u32 p0l, p0h, p1l, p1h, p2l, p2h, p3l, p3h;\n\t"
        "mul.wide.u32 p0, %4, %6;\n\tmul.wide.u32 p1, %4, %7;\n\tmul.wide.u32 p2, %5, %6;\n\tmul.wide.u32 p3, %5, %7;\n\t"
        "mov.b64 {p0l, p0h}, p0;\n\tmov.b64 {p1l, p1h}, p1;\n\tmov.b64 {p2l, p2h}, p2;\n\tmov.b64 {p3l, p3h}, p3;\n\t"
        "mov.u32 %0, p0l;\n\t"
        "add.cc.u32 %1, p0h, p1l;\n\taddc.cc.u32 %2, p1h, p3l;\n\taddc.u32 %3, p3h, 0;\n\t"
        "add.cc.u32 %1, %1, p2l;\n\taddc.cc.u32 %2, %2, p2h;\n\taddc.u32 %3, %3, 0;\n\t}"
        : "=r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3) : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    // reduce: lo = r1:r0, hl = r2, hh = r3:  lo - hh + hl * eps
    // t = lo - hh: borrow -> -= eps
    asm("{\n\t.reg .pred p;\n\t.reg .u32 c;\n\t"
        "sub.cc.u32 %0, %0, %2;\n\tsubc.cc.u32 %1, %1, 0;\n\tsubc.u32 c, 0, 0;\n\tsetp.ne.u32 p, c, 0;\n\t"
        "@p sub.cc.u32 %0, %0, 0xffffffff;\n\t@p subc.u32 %1, %1, 0;\n\t}"
        : "+r"(r0), "+r"(r1) : "r"(r3));
    // + hl * eps = (hl << 32) - hl:  lo -= hl (borrow into hi), hi += hl  -> net carry/borrow
    // do it as: t1 = hl * 0xffffffff (wide), r += t1, carry -> += eps
    asm("{\n\t.reg .pred p;\n\t.reg .u32 c, t0, t1;\n\t.reg .u64 t;\n\t"
        "mul.wide.u32 t, %2, 0xffffffff;\n\tmov.b64 {t0, t1}, t;\n\t"
        "add.cc.u32 %0, %0, t0;\n\taddc.cc.u32 %1, %1, t1;\n\taddc.u32 c, 0, 0;\n\tsetp.ne.u32 p, c, 0;\n\t"
        "@p add.cc.u32 %0, %0, 0xffffffff;\n\t@p addc.u32 %1, %1, 0;\n\t}"
        : "+r"(r0), "+r"(r1) : "r"(r2));
    // canonical: r >= q <=> r + eps carries
    asm("{\n\t.reg .pred p;\n\t.reg .u32 c, s0, s1;\n\t"
        "add.cc.u32 s0, %0, 0xffffffff;\n\taddc.cc.u32 s1, %1, 0;\n\taddc.u32 c, 0, 0;\n\tsetp.ne.u32 p, c, 0;\n\t"
        "@p mov.u32 %0, s0;\n\t@p mov.u32 %1, s1;\n\t}"
        : "+r"(r0), "+r"(r1));
    return ((u64)r1 << 32) | r0;
}
#endif
__global__ void bf(u64* x, const u64* w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    u64 X = x[2 * i], Y = x[2 * i + 1];
    const u64 T = gold_mul(Y, w[i]);
    x[2 * i] = gold_add(X, T);
    x[2 * i + 1] = gold_sub(X, T);
}
