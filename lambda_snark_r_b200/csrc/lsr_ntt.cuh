// lsr_ntt.cuh -- negacyclic NTT building blocks for sm_100a.
//
// Replaces seal::util::ntt_negacyclic_harvey / inverse_ntt_negacyclic_harvey as
// called from the reference's cpp-core/src/ntt.cpp:84,99.  Same contract:
// forward natural -> bit-reversed, inverse bit-reversed -> natural with n^-1
// folded in, outputs fully reduced to [0, q).
//
// Structure: a polynomial (or a 2^TLOG-coefficient block of a larger one) sits
// in shared memory; the log2 stages are grouped into "passes" of R <= 4 stages.
// In a pass each thread owns 2^R coefficients in registers, runs the R stages
// on them (radix-2^R butterfly network, twiddles read as 16-byte (w, w') pairs
// through the read-only path), and writes them back.  Only the pass boundaries
// touch shared memory.  The XOR swizzle below makes every pass of the plans
// chosen in plan_for() bank-conflict free.
#pragma once
#include <type_traits>

#include "lsr_arith.cuh"

// Unit-stride pass talking to HBM directly with 16-byte accesses instead of restaging
// through shared memory.  Measured at n = 4096 (tools/ntt_variant_bench.cu) with the u64
// butterflies, which are bound by the integer pipes: the forward store gains 3 % (36.6 vs 35.4
// M NTT/s), the inverse load loses 3 % (33.4 vs 34.5).  With the FP64 butterflies the L1 data
// pipe is the scarcer resource and the direct store (each lane writes its own 128-byte line,
// 32 lines per instruction) costs 17 %: 48.4 vs 58.7 M NTT/s, so POL_F64 always restages.
#ifndef LSR_NTT_DIRECT_OUT
#define LSR_NTT_DIRECT_OUT 1
#endif
#ifndef LSR_NTT_DIRECT_IN
#define LSR_NTT_DIRECT_IN 0
#endif
#ifndef LSR_SMEM_FWD_XR   // experiment: FRND rounding in the shared-memory-only forward passes too (fused commitment)
#define LSR_SMEM_FWD_XR 0
#endif
#ifndef LSR_NTT_MINB      // tools/ntt_variant_bench.cu overrides this to explore occupancy
#define LSR_NTT_MINB 3
#endif

namespace lsr {

#ifndef LSR_NTT_THREADS   // tools/ntt_variant_bench.cu overrides this
#define LSR_NTT_THREADS 256
#endif
constexpr int kNttThreads = LSR_NTT_THREADS;
constexpr int kTileLogMin = 12;   // a CTA always works on >= 4096 coefficients
#ifndef LSR_NTT_PAD
#define LSR_NTT_PAD 1
#endif
// shared-memory layout of the stand-alone transform kernels (see padx): measured +3.7 % on the forward kernel
// (61.6 vs 59.4 M NTT/s at n = 4096) and -2.8 % on the inverse one, so only the forward direction pads
// (the inverse kernel with the padded layout: -2.8 % at n = 4096, -2.9 % at 1024, -2.8 % at 2048, +2.0 % at 8192 -- so it
// pads at 8192 only, where its 32-value middle pass gains most from the constant offsets)
template <bool INVERSE, int LT = 12> __host__ __device__ constexpr bool ntt_pad() {
    return LSR_NTT_PAD != 0 && (!INVERSE || LT == 13);
}

// shared-memory swizzle: bits 0-3 ^= bits 4-7
__device__ __forceinline__ u32 swz(u32 i) { return i ^ ((i >> 4) & 15u); }
// Alternative layout of the stand-alone transforms (PAD): one unused slot after every 16 coefficients,
// index i + (i >> 4).  Equally conflict-free (strides 16 and 256 become 17 and 272), and the 2^R coefficients
// of a work item sit at CONSTANT offsets from one base register (every pass has stride 1 or >= 16), so the
// per-element XOR / shift address arithmetic of the swizzle disappears.  Costs 1/16 more shared memory,
// which the fused commitment kernel (3 CTAs of 72 KiB per SM) does not have: it keeps the swizzle.
__device__ __forceinline__ u32 padx(u32 i) { return i + (i >> 4); }
template <bool PAD> __device__ __forceinline__ u32 sidx(u32 i) { return PAD ? padx(i) : swz(i); }

__device__ __forceinline__ ulonglong2 ld_tw(const ulonglong2* p) {
#ifdef LSR_DIAG_NOTW    // tools/ntt_variant_bench.cu only: what would the kernel do if twiddle loads were free? (wrong results)
    return make_ulonglong2(0x40c81c8000000000ull /* 12345.0 */, 0x3d081c8000000000ull /* ~7e-13 */);
#else
    return __ldg(p);
#endif
}

// ---------------------------------------------------------------------------
// Register butterfly networks on v[0 .. 2^R), element j <-> index base + j*g.
//   forward stage r (0..R-1): half = 2^(R-1-r), twiddle index (T0 << r) + t
//   inverse stage r (0..R-1): half = 2^r,       twiddle index (T0 << (R-1-r)) + t
// T0 = 2^S + block, S = first forward stage of the pass.
// ---------------------------------------------------------------------------
// One forward stage r of a radix-2^R pass; recursion over r keeps every loop
// bound a template constant (plain nested `#pragma unroll` loops were being
// re-rolled by nvcc, which sent v[] to local memory).
//
// Twiddle addressing.  LL = false: heap layout, entry (T0 << r) + t.  LL = true
// ("last-pass layout", unit-stride pass where every work item has its own 2^R-1
// twiddles): entry ((2^r - 1 + t) * stride + T0) of a table transposed on the
// host so that the 32 lanes of a warp read 32 consecutive 16-byte pairs -- 4
// L1 wavefronts per load instead of up to 32.  T0 is then the work-item index.
template <bool LL>
__device__ __forceinline__ u32 tw_index(u32 T0, int r, int t, u32 stride) {
    return LL ? ((u32)((1 << r) - 1 + t) * stride + T0) : ((T0 << r) + (u32)t);
}

// HEAD: the pass starts at forward stage 0 of a whole polynomial, so T0 == 1 and
// the entries (1 << r) + t are the same for every thread of the grid: they are
// read from `head`, a by-value kernel parameter (constant bank), and reach the
// IMADs as uniform operands -- no load, no register, one RF read less each.
template <int R, int r, int POL, bool LL, bool HEAD, bool XR = false>
__device__ __forceinline__ void fwd_stage(u64 (&v)[1 << R], const ulonglong2* __restrict__ tw,
                                          const ulonglong2* __restrict__ head,
                                          u32 T0, u32 stride, const ModParams& mp) {
    if constexpr (r < R) {
        constexpr int half = 1 << (R - 1 - r);
#pragma unroll
        for (int t = 0; t < (1 << r); t++) {
            const ulonglong2 w = HEAD ? head[(1 << r) + t] : ld_tw(tw + tw_index<LL>(T0, r, t, stride));
#pragma unroll
            for (int jl = 0; jl < half; jl++) {
                const int j = (t << (R - r)) + jl;
                const int jj = j + half;
                if (POL == POL_F64) {
                    // balanced doubles: |T| <= 0.75 q, growth +0.75 q per stage, no offset needed
                    const double T = mulmod_f<XR>(as_d(v[jj]), as_d(w.x), as_d(w.y), mp.qd);
                    const double X = as_d(v[j]);
                    v[j] = as_u(__dadd_rn(X, T));
                    v[jj] = as_u(__dadd_rn(X, -T));
                } else if (POL == POL_GOLD) {
                    // T canonical, X any representative in [0, 2^64): sums stay lazy until fwd_final
                    const u64 T = gold_mul(v[jj], w.x);
                    const u64 X = v[j];
                    v[j] = gold_add_lazy(X, T);
                    v[jj] = gold_sub(X, T);
                } else if (POL == POL_LAZY) {
                    const u64 T = mulred4(v[jj], w.x, w.y, mp.nq);
                    const u64 X = v[j];
                    v[j] = X + T;
                    v[jj] = X + mp.q4 - T;
                } else {
                    const u64 X = csub(v[j], mp.q2);
                    const u64 T = mulred2(v[jj], w.x, w.y, mp.nq);
                    v[j] = X + T;
                    v[jj] = X + mp.q2 - T;
                }
            }
        }
        fwd_stage<R, r + 1, POL, LL, HEAD, XR>(v, tw, head, T0, stride, mp);
    }
}

// XR: POL_F64 products round their quotient estimate with FRND.F64 (lsr_arith.cuh rint_prod); the stand-alone forward kernels
template <int R, int POL, bool LL = false, bool HEAD = false, bool XR = false>
__device__ __forceinline__ void fwd_network(u64 (&v)[1 << R], const ulonglong2* __restrict__ tw,
                                            u32 T0, const ModParams& mp, u32 stride = 0,
                                            const ulonglong2* __restrict__ head = nullptr) {
    fwd_stage<R, 0, POL, LL, HEAD, XR>(v, tw, head, T0, stride, mp);
}

// last inverse stage (m = 1): scalar n^-1 folded in (SEAL transform_from_rev
// with scalar; w_scaled = inv[1] is already multiplied by n^-1), then the
// final correction to [0, q)
// RAW (POL_F64 only): leave the two outputs as balanced doubles, |r| <= 0.75 q, for an epilogue that keeps working
// in FP64 and canonicalises itself (the fused commitment adds its error term first)
template <int POL, bool RAW = false>
__device__ __forceinline__ void inv_last_butterfly(u64& x, u64& y, const ulonglong2 w_scaled,
                                                   const ulonglong2 n_inv, int sigma,
                                                   const ModParams& mp) {
    const u64 X = x, Y = y;
    if (POL == POL_F64) {
        const double Xd = as_d(X), Yd = as_d(Y);
        const double rx = mulmod_f(__dadd_rn(Xd, Yd), as_d(n_inv.x), as_d(n_inv.y), mp.qd);
        const double ry = mulmod_f(__dadd_rn(Xd, -Yd), as_d(w_scaled.x), as_d(w_scaled.y), mp.qd);
        x = RAW ? as_u(rx) : f_to_canonical(rx, mp);
        y = RAW ? as_u(ry) : f_to_canonical(ry, mp);
    } else if (POL == POL_GOLD) {
        x = gold_mul(gold_add(X, Y), n_inv.x);
        y = gold_mul(gold_sub(X, Y), w_scaled.x);
    } else if (POL == POL_LAZY) {
        const u64 C = mp.q4 << sigma;
        x = csub(csub(mulred4(X + Y, n_inv.x, n_inv.y, mp.nq), mp.q2), mp.q);
        y = csub(csub(mulred4(X + C - Y, w_scaled.x, w_scaled.y, mp.nq), mp.q2), mp.q);
    } else {
        x = csub(mulred2(csub(X + Y, mp.q2), n_inv.x, n_inv.y, mp.nq), mp.q);
        y = csub(mulred2(X + mp.q2 - Y, w_scaled.x, w_scaled.y, mp.nq), mp.q);
    }
}

// sigma0 = number of inverse stages already done before this pass (growth
// bound: a value entering inverse stage sigma is < 4q * 2^sigma on the lazy
// path).  FINAL: the pass ends with the m = 1 stage (forward stage 0).
template <int R, int r, int POL, bool FINAL, bool LL, bool HEAD, bool RAW = false>
__device__ __forceinline__ void inv_stage(u64 (&v)[1 << R], const ulonglong2* __restrict__ tw,
                                          const ulonglong2* __restrict__ head,
                                          u32 T0, u32 stride, int sigma0, const ulonglong2 n_inv,
                                          const ModParams& mp) {
    if constexpr (r < R) {
        constexpr int half = 1 << r;
        constexpr int fr = R - 1 - r;          // matching forward stage within the pass
        const u64 C = mp.q4 << (sigma0 + r);
#pragma unroll
        for (int t = 0; t < (1 << fr); t++) {
            const ulonglong2 w = HEAD ? head[(1 << fr) + t] : ld_tw(tw + tw_index<LL>(T0, fr, t, stride));
#pragma unroll
            for (int jl = 0; jl < half; jl++) {
                const int j = (t << (r + 1)) + jl;
                const int jj = j + half;
                if constexpr (FINAL && r == R - 1) {
                    inv_last_butterfly<POL, RAW>(v[j], v[jj], w, n_inv, sigma0 + r, mp);
                } else {
                    const u64 X = v[j], Y = v[jj];
                    if (POL == POL_F64) {
                        const double Xd = as_d(X), Yd = as_d(Y);
                        v[j] = as_u(__dadd_rn(Xd, Yd));
                        v[jj] = as_u(mulmod_f(__dadd_rn(Xd, -Yd), as_d(w.x), as_d(w.y), mp.qd));
                    } else if (POL == POL_GOLD) {
                        v[j] = gold_add(X, Y);
                        v[jj] = gold_mul(gold_sub(X, Y), w.x);
                    } else if (POL == POL_LAZY) {
                        v[j] = X + Y;
                        v[jj] = mulred4(X + C - Y, w.x, w.y, mp.nq);
                    } else {
                        v[j] = csub(X + Y, mp.q2);
                        v[jj] = mulred2(X + mp.q2 - Y, w.x, w.y, mp.nq);
                    }
                }
            }
        }
        inv_stage<R, r + 1, POL, FINAL, LL, HEAD, RAW>(v, tw, head, T0, stride, sigma0, n_inv, mp);
    }
}

// POL_F64 growth control.  In a GS pass element j is a product at stage r iff bit r of j is
// set, and a sum (doubling) otherwise: leaving the pass, element 0 is bounded by 2^R b (b = bound
// of the inputs) and element j != 0 by 2^(R-1-h) P, h = highest set bit of j, P = 0.75 q the
// bound of a product.  Reducing the 2^(R-2) elements with the largest bounds (j < 2^(R-2)) leaves
// every element <= 2P = 1.5 q, so the next pass (R <= 5) sees |X - Y| <= 2^R * 1.5 q <= 48 q
// < 2^51 for q < 2^45.  Cost: 3 FP64 instructions on a quarter of the elements per pass.
template <int R, int POL, bool FINAL, bool LL = false, bool HEAD = false, bool RAW = false>
__device__ __forceinline__ void inv_network(u64 (&v)[1 << R], const ulonglong2* __restrict__ tw,
                                            u32 T0, int sigma0, const ulonglong2 n_inv,
                                            const ModParams& mp, u32 stride = 0,
                                            const ulonglong2* __restrict__ head = nullptr) {
    inv_stage<R, 0, POL, FINAL, LL, HEAD, RAW>(v, tw, head, T0, stride, sigma0, n_inv, mp);
    if constexpr (POL == POL_F64 && !FINAL) {
        constexpr int NRED = R >= 2 ? (1 << (R - 2)) : 1;
#pragma unroll
        for (int j = 0; j < NRED; j++) v[j] = as_u(reduce_f(as_d(v[j]), mp.invq, mp.qd));
    }
}

template <int POL, bool XR = false>
__device__ __forceinline__ u64 fwd_final(u64 v, const ModParams& mp) {
    if (POL == POL_F64) return f_to_canonical(reduce_f<XR>(as_d(v), mp.invq, mp.qd), mp);
    if (POL == POL_GOLD) return gold_canonical(v);
    if (POL == POL_LAZY) return reduce_small(v, mp);
    return csub(csub(v, mp.q2), mp.q);
}

// raw u64 from the caller -> the policy's working representation
template <int POL>
__device__ __forceinline__ u64 to_working(u64 v) {
    if (POL == POL_F64) return as_u(u64_to_f(v));
    return v;
}

// ---------------------------------------------------------------------------
// Pass plans.  LT = number of stages done inside the tile (= log2 tile size
// per polynomial block).  The last forward pass has R = min(4, LT) and unit
// stride; every earlier pass has stride g >= 16, which is what the swizzle
// needs.  plan<LT>::R[i] lists forward passes in execution order.
// ---------------------------------------------------------------------------
template <int LT> struct plan;
template <> struct plan<1>  { static constexpr int N = 1; static constexpr int R[4] = {1, 0, 0, 0}; };
template <> struct plan<2>  { static constexpr int N = 1; static constexpr int R[4] = {2, 0, 0, 0}; };
template <> struct plan<3>  { static constexpr int N = 1; static constexpr int R[4] = {3, 0, 0, 0}; };
template <> struct plan<4>  { static constexpr int N = 1; static constexpr int R[4] = {4, 0, 0, 0}; };
template <> struct plan<5>  { static constexpr int N = 2; static constexpr int R[4] = {1, 4, 0, 0}; };
template <> struct plan<6>  { static constexpr int N = 2; static constexpr int R[4] = {2, 4, 0, 0}; };
template <> struct plan<7>  { static constexpr int N = 2; static constexpr int R[4] = {3, 4, 0, 0}; };
template <> struct plan<8>  { static constexpr int N = 2; static constexpr int R[4] = {4, 4, 0, 0}; };
template <> struct plan<9>  { static constexpr int N = 3; static constexpr int R[4] = {4, 1, 4, 0}; };   // forward 492 -> 550 M NTT/s against 3+2+4 (the inverse keeps it)
// n = 1024, 2048: the two directions want different splits (tools/plan_bench.py, M NTT/s forward / inverse at n = 1024:
// 3+3+4 236 / 242, 4+2+4 268 / 229, 2+4+4 209 / 239; n = 2048: 4+3+4 127.5 / 108.7, 3+4+4 116.6 / 114.4) -- the forward
// transform reads its first pass straight from HBM and gains from 16 loads in flight per thread, the inverse ends on that
// pass and prefers the even split.  plan_inv<LT> is the plan the inverse direction runs (in reverse order).
template <> struct plan<10> { static constexpr int N = 3; static constexpr int R[4] = {4, 2, 4, 0}; };
template <> struct plan<11> { static constexpr int N = 3; static constexpr int R[4] = {4, 3, 4, 0}; };
template <> struct plan<12> { static constexpr int N = 3; static constexpr int R[4] = {4, 4, 4, 0}; };
// n = 8192: three passes, one of them on 32 register values per thread, instead of the four of 3+3+3+4 (21.7 M NTT/s
// forward).  The 32-value pass goes in the middle: 4+5+4 runs at 26.2 / 22.3 M NTT/s forward / inverse against 24.3 / 21.4
// for 5+4+4 -- the pass that talks to HBM (first forward, last inverse) is better off with 16 values and the registers
// that leaves for addressing (tools/plan_bench.py)
template <> struct plan<13> { static constexpr int N = 3; static constexpr int R[4] = {4, 5, 4, 0}; };
template <> struct plan<14> { static constexpr int N = 4; static constexpr int R[4] = {4, 3, 3, 4}; };
template <int LT> struct plan_inv : plan<LT> {};
template <> struct plan_inv<9>  { static constexpr int N = 3; static constexpr int R[4] = {3, 2, 4, 0}; };
template <> struct plan_inv<10> { static constexpr int N = 3; static constexpr int R[4] = {3, 3, 4, 0}; };
template <> struct plan_inv<11> { static constexpr int N = 3; static constexpr int R[4] = {3, 4, 4, 0}; };
// every plan covers its LT stages, ends on the unit-stride radix-16 pass the transposed twiddle table is laid out for
// (LT > 4), and keeps every earlier pass at stride >= 16 (what the swizzle and the padded layout need)
template <typename P, int LT> constexpr bool plan_ok() {
    int sum = 0;
    for (int i = 0; i < P::N; i++) sum += P::R[i];
    if (sum != LT) return false;
    if (LT > 4 && P::R[P::N - 1] != 4) return false;
    int done = 0;
    for (int i = 0; i + 1 < P::N; i++) { done += P::R[i]; if (LT - done < 4) return false; }
    return true;
}
template <int LT> constexpr bool plans_ok() {
    if constexpr (LT == 0) return true;
    else return plan_ok<plan<LT>, LT>() && plan_ok<plan_inv<LT>, LT>() && plan<LT>::N == plan_inv<LT>::N && plans_ok<LT - 1>();
}
static_assert(plans_ok<14>(), "pass plan table");

// ---------------------------------------------------------------------------
// One pass over a tile.
//   LOGN  : log2 of the full transform size
//   S     : first forward stage covered by this pass (global numbering)
//   R     : stages in the pass
//   LT    : log2 of the per-polynomial block held in the tile (LT <= LOGN);
//           the block covers forward stages LOGN-LT .. LOGN-1
//   items : work items in the tile = tile_elems >> R
//   tb    : index of the tile's block within its polynomial (0 when LT==LOGN)
//           (for multi-polynomial tiles, LT == LOGN and tb == 0)
//   IN/OUT: where the pass reads / writes its coefficients.  IO_GLOBAL fuses
//           the HBM load (store) of the tile into the first (last) pass whose
//           thread->coefficient map is coalesced, so the tile crosses shared
//           memory one time less.  All global loads of a work item are issued
//           before any is consumed (16 independent requests per thread).
// ---------------------------------------------------------------------------
enum : int { IO_SMEM = 0, IO_GLOBAL = 1 };

template <int POL>
__host__ __device__ constexpr bool ntt_direct_out() { return LSR_NTT_DIRECT_OUT && POL != POL_F64; }

struct TileIo {
    u64* g;        // tile base in global memory (IO_GLOBAL only)
    u32 valid;     // coefficients of the tile that exist (multiple of 2^LT)
    u32 sanitize;  // the tile is raw caller data: reduce out-of-range inputs on load and convert
                   // to the policy's working representation (POL_F64: doubles)
    u64 limit;     // sanitiser threshold
};

// Epilogue hook for the pass that stores to global memory: value = epi(tile index, value).
// kWholeItem = true: the hook takes the whole work item instead -- item(g, W, base, v) stores the 2^R coefficients
// base + (j << LG) itself (the fused commitment samples its error terms right there, in registers).
// pre(W) runs before the coefficients of the work item are loaded (few live registers) and hands its result to item().
struct NoEpilogue {
    static constexpr bool kWholeItem = false;
    static constexpr bool kRawF64 = false;    // true: POL_F64 final inverse pass hands balanced doubles to item()
    struct Pre {};
    __device__ __forceinline__ u64 operator()(u32, u64 v) const { return v; }
};

// Final pass of a fused inverse transform (InvFusion::fin_c): value = (c[idx] - value) * scale, or the value itself when c is null.
struct FinEpilogue {
    static constexpr bool kWholeItem = false;
    static constexpr bool kRawF64 = false;
    struct Pre {};
    const u64* c;          // tile-relative
    u64 scale;
    const ModParams* mp;
    __device__ __forceinline__ u64 operator()(u32 idx, u64 v) const {
        return c ? field_mul(field_sub(__ldcs(c + idx), v, *mp), scale, *mp) : v;
    }
};

// WHOLE: the tile holds whole polynomials (LT == log n).  Otherwise it is one 2^LT block of a larger
// polynomial and d = log n - LT (a run-time value: one kernel serves every big ring degree) only enters
// the twiddle indices.  SL = first stage of the pass, counted inside the block.
template <int LT, bool WHOLE, int SL, int R, int POL, bool INVERSE, bool FINAL, int IN, int OUT, typename Epi = NoEpilogue, bool PAD = false,
          bool XR = false>
__device__ __forceinline__ void tile_pass(u64* __restrict__ sm, const TileIo& io, const NttTables& tb_,
                                          const ModParams& mp, u32 items, u32 tb, u32 d_, const Epi& epi = Epi()) {
    constexpr int LG = LT - SL - R;           // log2 of the element stride g
    constexpr u32 g = 1u << LG;
    static_assert(LG >= 0 && SL >= 0, "bad pass");
    static_assert(!PAD || IN != IO_SMEM || LG == 0 || LG >= 4, "padded layout needs stride 1 or >= 16");
    static_assert(!PAD || OUT != IO_SMEM || LG == 0 || LG >= 4, "padded layout needs stride 1 or >= 16");
    constexpr u32 pstep = LG >= 4 ? (1u << LG) + (1u << (LG >= 4 ? LG - 4 : 0)) : 1u;   // padded element stride
    const u32 d = WHOLE ? 0u : d_;
    // unit-stride radix-16 pass of a multi-pass plan: per-work-item twiddles, transposed table
    constexpr bool LL = (LG == 0) && (R == 4) && (LT > 4);
    constexpr bool HEAD = (SL == 0) && WHOLE && !LL;           // twiddles 1 .. 2^R - 1, grid-uniform
    const u32 ll_stride = 1u << (d + LT - R);
    const ulonglong2* __restrict__ tw = LL ? (INVERSE ? tb_.inv_last : tb_.fwd_last) : (INVERSE ? tb_.inv : tb_.fwd);
    for (u32 W = threadIdx.x; W < items; W += blockDim.x) {
        const u32 poly = W >> (LT - R);                   // polynomial slot within the tile
        const u32 w = W & ((1u << (LT - R)) - 1u);
        const u32 blk = w >> LG;
        const u32 c = w & (g - 1u);
        const u32 base = (poly << LT) + (blk << (LG + R)) + c;
        const u32 T0 = LL ? ((tb << (LT - R)) + w) : ((1u << (d + SL)) + (tb << SL) + blk);
        const bool live = (IN == IO_GLOBAL || OUT == IO_GLOBAL) ? ((poly << LT) < io.valid) : true;
        typename Epi::Pre pre;
        if constexpr (OUT == IO_GLOBAL && Epi::kWholeItem) pre = epi.pre(W);
        u64 v[1 << R];
        if constexpr (IN == IO_GLOBAL && LG == 0 && R >= 1) {
            // unit-stride pass: each thread owns 2^R consecutive coefficients -> 16-byte accesses
            const ulonglong2* __restrict__ gp = reinterpret_cast<const ulonglong2*>(io.g + base);
#pragma unroll
            for (int j = 0; j < (1 << R); j += 2) {
                const ulonglong2 t2 = live ? __ldcs(gp + (j >> 1)) : make_ulonglong2(0ull, 0ull);
                v[j] = t2.x; v[j + 1] = t2.y;
            }
            if (io.sanitize) {
                u64 any = 0;
#pragma unroll
                for (int j = 0; j < (1 << R); j++) any |= v[j];
                if (__builtin_expect(any >= io.limit, 0)) {
#pragma unroll
                    for (int j = 0; j < (1 << R); j++) v[j] = sanitize(v[j], io.limit, mp);
                }
#pragma unroll
                for (int j = 0; j < (1 << R); j++) v[j] = to_working<POL>(v[j]);
            }
        } else if constexpr (IN == IO_GLOBAL) {
#pragma unroll
            for (int j = 0; j < (1 << R); j++) v[j] = live ? __ldcs(io.g + base + ((u32)j << LG)) : 0ull;
            if (io.sanitize) {
                // every element <= the OR of all: one test per work item, slow path only if it can matter
                u64 any = 0;
#pragma unroll
                for (int j = 0; j < (1 << R); j++) any |= v[j];
                if (__builtin_expect(any >= io.limit, 0)) {
#pragma unroll
                    for (int j = 0; j < (1 << R); j++) v[j] = sanitize(v[j], io.limit, mp);
                }
#pragma unroll
                for (int j = 0; j < (1 << R); j++) v[j] = to_working<POL>(v[j]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < (1 << R); j++) v[j] = PAD ? sm[padx(base) + (u32)j * pstep] : sm[swz(base + ((u32)j << LG))];
            // whole-item epilogues may reuse the slots just read (they belong to this thread alone in this pass)
            if constexpr (OUT == IO_GLOBAL && Epi::kWholeItem) epi.loaded(W, base);
        }
        if constexpr (!INVERSE) {
            fwd_network<R, POL, LL, HEAD, XR>(v, tw, T0, mp, ll_stride, tb_.head_fwd);
            if (FINAL) {
#pragma unroll
                for (int j = 0; j < (1 << R); j++) v[j] = fwd_final<POL, XR>(v[j], mp);
            }
        } else {
            constexpr int sigma0 = LT - SL - R;           // inverse stages already done
            static_assert(!INVERSE || !FINAL || (SL == 0 && WHOLE), "final inverse pass must contain stage 0");
            constexpr bool RAW = FINAL && OUT == IO_GLOBAL && POL == POL_F64 && Epi::kWholeItem && Epi::kRawF64;
            inv_network<R, POL, FINAL, LL, HEAD, RAW>(v, tw, T0, sigma0, tb_.n_inv, mp, ll_stride, tb_.head_inv);
        }
        if constexpr (OUT == IO_GLOBAL && LG == 0 && R >= 1 && std::is_same<Epi, NoEpilogue>::value) {
            if (live) {
                ulonglong2* __restrict__ gp = reinterpret_cast<ulonglong2*>(io.g + base);
#pragma unroll
                for (int j = 0; j < (1 << R); j += 2) __stcs(gp + (j >> 1), make_ulonglong2(v[j], v[j + 1]));
            }
        } else if constexpr (OUT == IO_GLOBAL) {
            if constexpr (Epi::kWholeItem) {
                if (live) epi.item(io.g, W, base, v, pre);
            } else if (live) {
#pragma unroll
                for (int j = 0; j < (1 << R); j++) {
                    const u32 idx = base + ((u32)j << LG);
                    __stcs(io.g + idx, epi(idx, v[j]));
                }
            }
        } else {
#pragma unroll
            for (int j = 0; j < (1 << R); j++) {
                if (PAD) sm[padx(base) + (u32)j * pstep] = v[j];
                else sm[swz(base + ((u32)j << LG))] = v[j];
            }
        }
    }
}

// ---------------------------------------------------------------------------
// All passes of a tile with a barrier after each pass that wrote shared memory.
// tile_elems = coefficients of the tile (multiple of 2^LT).
// Forward: stages LOGN-LT .. LOGN-1 (always ends with the final reduction).
// Inverse: the same stages in reverse; FINAL (n^-1 + correction) iff LT==LOGN.
// GIO = false: shared memory in, shared memory out (caller synchronised after
//              filling `sm`; used by the fused commitment kernel).
// GIO = true : forward reads its first pass from global memory and leaves the
//              result in shared memory (global when the plan has one pass);
//              inverse expects the tile in shared memory (global when one
//              pass) and writes its last pass to global memory.
// ---------------------------------------------------------------------------
template <int LT, bool WHOLE, int POL, bool GIO, int I, bool FIN = true, bool PAD = false>
__device__ __forceinline__ void tile_forward_from(u64* sm, const TileIo& io, const NttTables& t,
                                                  const ModParams& mp, u32 tile_elems, u32 tb, u32 d) {
    using P = plan<LT>;
    if constexpr (I < P::N) {
        constexpr int SL = (I > 0 ? P::R[0] : 0) + (I > 1 ? P::R[1] : 0) + (I > 2 ? P::R[2] : 0);
        constexpr int R = P::R[I];
        constexpr int IN = (GIO && I == 0) ? IO_GLOBAL : IO_SMEM;
        constexpr int OUT = (GIO && (P::N == 1 || (ntt_direct_out<POL>() && I == P::N - 1))) ? IO_GLOBAL : IO_SMEM;
        tile_pass<LT, WHOLE, SL, R, POL, false, (FIN && I == P::N - 1), IN, OUT, NoEpilogue, PAD, ((GIO && WHOLE && LT <= 12) || LSR_SMEM_FWD_XR)>(sm, io, t, mp, tile_elems >> R, tb, d);      // LT = 13: FRND in the 32-value pass spills, in the 16-value passes alone it gains 0.6 %; block tiles of the big ring degrees: -1 %
        if constexpr (OUT == IO_SMEM) __syncthreads();
        tile_forward_from<LT, WHOLE, POL, GIO, I + 1, FIN, PAD>(sm, io, t, mp, tile_elems, tb, d);
    }
}

template <int LT, bool WHOLE, int POL, bool GIO, int I, typename Epi = NoEpilogue, bool PAD = false>
__device__ __forceinline__ void tile_inverse_from(u64* sm, const TileIo& io, const NttTables& t,
                                                  const ModParams& mp, u32 tile_elems, u32 tb, u32 d,
                                                  const Epi& epi = Epi()) {
    using P = plan_inv<LT>;
    static_assert(P::N == plan<LT>::N, "both directions make the same number of passes");
    if constexpr (I >= 0) {
        constexpr int SL = (I > 0 ? P::R[0] : 0) + (I > 1 ? P::R[1] : 0) + (I > 2 ? P::R[2] : 0);
        constexpr int R = P::R[I];
        constexpr int IN = (GIO && (P::N == 1 || (LSR_NTT_DIRECT_IN && I == P::N - 1))) ? IO_GLOBAL : IO_SMEM;
        constexpr int OUT = (GIO && I == 0) ? IO_GLOBAL : IO_SMEM;
        tile_pass<LT, WHOLE, SL, R, POL, true, (I == 0 && WHOLE), IN, OUT, Epi, PAD>(sm, io, t, mp, tile_elems >> R, tb, d, epi);
        if constexpr (OUT == IO_SMEM) __syncthreads();
        tile_inverse_from<LT, WHOLE, POL, GIO, I - 1, Epi, PAD>(sm, io, t, mp, tile_elems, tb, d, epi);
    }
}

// shared-memory-only forms (fused commitment kernel).  Values in shared memory are in the
// policy's working representation.  POL_F64 forward leaves the evaluations unreduced (bounded by
// (b + 0.75 logn) q for inputs bounded by b q): the consumer multiplies them, which reduces.
// I0 = first pass to run (the fused commitment kernel does pass 0 itself, on freshly sampled registers)
template <int LOGN, int LT, int POL, bool PAD = false, int I0 = 0>
__device__ __forceinline__ void tile_forward(u64* sm, const NttTables& t, const ModParams& mp,
                                             u32 tile_elems, u32 tb) {
    static_assert(LOGN == LT, "whole polynomials only");
    const TileIo io{nullptr, tile_elems, 0u, 0ull};
    tile_forward_from<LT, true, POL, false, I0, POL != POL_F64, PAD>(sm, io, t, mp, tile_elems, tb, 0u);
}
template <int LOGN, int LT, int POL>
__device__ __forceinline__ void tile_inverse(u64* sm, const NttTables& t, const ModParams& mp,
                                             u32 tile_elems, u32 tb) {
    static_assert(LOGN == LT, "whole polynomials only");
    const TileIo io{nullptr, tile_elems, 0u, 0ull};
    tile_inverse_from<LT, true, POL, false, plan<LT>::N - 1>(sm, io, t, mp, tile_elems, tb, 0u);
}

// shared memory in, last pass straight to global memory through an epilogue
// (fused commitment kernel: + e, + Delta*m, container store); needs a multi-pass plan
template <int LOGN, int LT, int POL, typename Epi, bool PAD = false>
__device__ __forceinline__ void tile_inverse_to_global(u64* sm, u64* g, const NttTables& t, const ModParams& mp,
                                                       u32 tile_elems, const Epi& epi) {
    static_assert(plan<LT>::N > 1, "single-pass plans read from global memory");
    static_assert(LOGN == LT, "whole polynomials only");
    const TileIo io{g, tile_elems, 0u, 0ull};
    tile_inverse_from<LT, true, POL, true, plan<LT>::N - 1, Epi, PAD>(sm, io, t, mp, tile_elems, 0u, 0u, epi);
}

// ---------------------------------------------------------------------------
// K1 / K2: batched transform, one tile (>= 4096 coefficients) per CTA.
//   data        : [batch][n] contiguous u64, transformed in place
//   total_elems : batch * n
// LT == LOGN: a tile is max(1, 4096/n) whole polynomials.
// LT <  LOGN: a tile is one 2^LT block of a polynomial (big-n second kernel
//             forward / first kernel inverse); values are already lazy, so the
//             forward direction must not sanitise.
// Forward: HBM -> registers (first pass, coalesced) ... last pass -> shared
//          memory -> coalesced store.  Inverse: coalesced load -> shared memory
//          -> first pass ... last pass (coalesced map) -> HBM.
// ---------------------------------------------------------------------------
#ifndef LSR_NTT_MINB_GOLD      // Goldilocks tiles: resident CTAs per SM the register allocation is held to
#define LSR_NTT_MINB_GOLD LSR_NTT_MINB
#endif
template <int LT, int POL = POL_F64>
constexpr int ntt_min_blocks() { return LT <= 12 ? (POL == POL_GOLD ? LSR_NTT_MINB_GOLD : LSR_NTT_MINB) : (LT == 13 ? 2 : 1); }

// FUSED (inverse only): the InvFusion hooks are compiled in; the plain instantiation is exactly the stand-alone transform.
template <int LT, bool WHOLE, int POL, bool INVERSE, bool FUSED = false>
__global__ void __launch_bounds__(kNttThreads, ntt_min_blocks<LT, POL>())
ntt_tile_kernel(const ModParams mp, const NttTables tbl, u64* __restrict__ data, size_t total_elems, u32 d, const InvFusion fz) {
    extern __shared__ __align__(16) u64 sm[];
    constexpr int TL = LT > kTileLogMin ? LT : kTileLogMin;
    constexpr u32 TILE = 1u << TL;
    constexpr u32 PER_THREAD = TILE / kNttThreads;
    constexpr bool ONE_PASS = plan<LT>::N == 1;
    const size_t tile0 = (size_t)blockIdx.x << TL;
    const u32 valid = (u32)((total_elems - tile0) < TILE ? (total_elems - tile0) : TILE);
    u64* __restrict__ g = data + tile0;
    const u32 tb = WHOLE ? 0u : (blockIdx.x & ((1u << d) - 1u));
    const bool clean = (!INVERSE && !WHOLE);      // already-lazy values: no sanitiser
    const TileIo io{g, valid, clean ? 0u : 1u, POL == POL_GOLD ? mp.q : (INVERSE ? mp.q2 : mp.q4)};

    if constexpr (!INVERSE) {
        tile_forward_from<LT, WHOLE, POL, true, 0, true, ntt_pad<false>()>(sm, io, tbl, mp, TILE, tb, d);
        if constexpr (!ONE_PASS && !ntt_direct_out<POL>()) {
#pragma unroll 4
            for (u32 k = 0; k < PER_THREAD; k++) {
                const u32 i = threadIdx.x + k * kNttThreads;
                if (i < valid) __stcs(g + i, sm[sidx<ntt_pad<false>()>(i)]);
            }
        }
    } else {
        if constexpr (!ONE_PASS && !LSR_NTT_DIRECT_IN) {
            u64 x[PER_THREAD];
            if constexpr (FUSED) {
                // fused pointwise product: input = data * mul.  Both operands of all the thread's coefficients are requested
                // before the first product (with the second operand behind the first one's use the tile kernel of the
                // fused transform ran at 19 us per 2^20 transform against 14 for the plain one)
                const bool has_mul = fz.mul != nullptr;
                const u64* __restrict__ g2 = has_mul ? fz.mul + tile0 : g;
                u64 y[PER_THREAD];
#pragma unroll
                for (u32 k = 0; k < PER_THREAD; k++) {
                    const u32 i = threadIdx.x + k * kNttThreads;
                    x[k] = i < valid ? __ldcs(g + i) : 0ull;
                    y[k] = (has_mul && i < valid) ? __ldcs(g2 + i) : 0ull;
                }
                if (has_mul) {
#pragma unroll
                    for (u32 k = 0; k < PER_THREAD; k++) x[k] = mulmod_exact(x[k], y[k], mp);
                }
            } else {
#pragma unroll
                for (u32 k = 0; k < PER_THREAD; k++) {
                    const u32 i = threadIdx.x + k * kNttThreads;
                    x[k] = i < valid ? __ldcs(g + i) : 0ull;
                }
            }
#pragma unroll
            for (u32 k = 0; k < PER_THREAD; k++) {
                const u32 i = threadIdx.x + k * kNttThreads;
                sm[sidx<ntt_pad<true, LT>()>(i)] = to_working<POL>(sanitize(x[k], io.limit, mp));
            }
            __syncthreads();
            if constexpr (FUSED) {
                // out of place (first kernel of a fused transform) and / or the finishing epilogue on the last pass
                const TileIo io_out{fz.dst ? fz.dst + tile0 : g, valid, io.sanitize, io.limit};
                const FinEpilogue fin{(WHOLE && fz.fin_c) ? fz.fin_c + tile0 : nullptr, fz.fin_scale, &mp};
                tile_inverse_from<LT, WHOLE, POL, true, plan<LT>::N - 1, FinEpilogue, ntt_pad<true, LT>()>(sm, io_out, tbl, mp, TILE, tb, d, fin);
            } else {
                tile_inverse_from<LT, WHOLE, POL, true, plan<LT>::N - 1, NoEpilogue, ntt_pad<true, LT>()>(sm, io, tbl, mp, TILE, tb, d);
            }
        } else {
            tile_inverse_from<LT, WHOLE, POL, true, plan<LT>::N - 1, NoEpilogue, ntt_pad<true, LT>()>(sm, io, tbl, mp, TILE, tb, d);
        }
    }
}

// ---------------------------------------------------------------------------
// Big-n helper (n >= 2^14): S forward stages s0 .. s0+S-1 ahead of the 4096-blocks of the tile kernel
// (or the matching inverse stages behind it).  At stage s0 a polynomial is 2^s0 blocks of
// n >> s0 coefficients; the S stages couple coefficients g = n >> (s0 + S) apart inside a block.
// A thread owns the 2^S coefficients of one column, straight from / to global memory (consecutive
// threads take consecutive columns: coalesced, g >= 4096), no shared memory.  log n, s0 are
// run-time values: n <= 2^17 needs one such pass (S <= 5), the cyclic transforms of the quotient
// pipeline (up to 2^24) chain two or three.  FIRST: the pass reads raw caller data (forward s0 == 0)
// or ends the inverse transform (n^-1 folded in, canonical output).
// ---------------------------------------------------------------------------
// FIN (the kernel that ends an inverse transform only): the InvFusion finishing step is compiled in; the plain instantiation
// carries no register for it
template <int S, int POL, bool INVERSE, bool FIRST, bool FIN = false>
__global__ void __launch_bounds__(kNttThreads)
ntt_column_kernel(const ModParams mp, const NttTables tbl, u64* __restrict__ data, size_t batch, u32 logn, u32 s0,
                  const u64* __restrict__ fin_c, u64 fin_scale) {
    const u32 LG = logn - s0 - (u32)S;                   // log2 g
    const size_t cols = batch << (logn - (u32)S);        // (polynomial, block, column) triples
    const size_t idx = (size_t)blockIdx.x * kNttThreads + threadIdx.x;
    if (idx >= cols) return;
    const size_t pb = idx >> LG;                         // polynomial * 2^s0 + block
    const u32 c = (u32)(idx & (((size_t)1 << LG) - 1u));
    const u32 blk = (u32)(pb & (((size_t)1 << s0) - 1u));
    u64* __restrict__ g = data + (pb << (LG + (u32)S)) + c;
    const u32 T0 = (1u << s0) + blk;
    u64 v[1 << S];
    if (!INVERSE) {
        // all loads are in flight before the first value is looked at
#pragma unroll
        for (int j = 0; j < (1 << S); j++) v[j] = g[(size_t)j << LG];
        if (FIRST) {
            // raw caller data: every element <= the OR of all, so one test per thread decides whether the
            // exact reduction of out-of-range inputs (rare, divergent) can matter at all
            const u64 limit = POL == POL_GOLD ? mp.q : mp.q4;
            u64 any = 0;
#pragma unroll
            for (int j = 0; j < (1 << S); j++) any |= v[j];
            if (__builtin_expect(any >= limit, 0)) {
#pragma unroll
                for (int j = 0; j < (1 << S); j++) v[j] = sanitize(v[j], limit, mp);
            }
#pragma unroll
            for (int j = 0; j < (1 << S); j++) v[j] = to_working<POL>(v[j]);
        }
        fwd_network<S, POL>(v, tbl.fwd, T0, mp);
    } else {
        // fused finishing step (InvFusion): (c - x) * scale; the c words are requested together with the coefficients
        const u64* __restrict__ cf = FIN ? fin_c + (pb << (LG + (u32)S)) + c : g;
        u64 cv[FIN ? (1 << S) : 1];
#pragma unroll
        for (int j = 0; j < (1 << S); j++) {
            v[j] = g[(size_t)j << LG];
            if (FIN) cv[j] = __ldcs(cf + ((size_t)j << LG));
        }
        inv_network<S, POL, FIRST>(v, tbl.inv, T0, (int)LG, tbl.n_inv, mp);
        if constexpr (FIN) {
#pragma unroll
            for (int j = 0; j < (1 << S); j++) v[j] = field_mul(field_sub(cv[FIN ? j : 0], v[j], mp), fin_scale, mp);
        }
    }
#pragma unroll
    for (int j = 0; j < (1 << S); j++) g[(size_t)j << LG] = v[j];
}

// ---------------------------------------------------------------------------
// Big-n helper for n >= 2^18: S1 + S2 column stages (6 .. 8) in ONE HBM round trip.  A CTA takes the
// 2^(S1+S2) rows x COLS adjacent columns (rows are g = n >> (s0 + S1 + S2) apart, COLS * 8 >= 128 bytes of
// every row are contiguous) = 4096 coefficients through shared memory: phase A is the register pass of
// ntt_column_kernel<S1> on rows r0 + (j << S2), phase B the one of ntt_column_kernel<S2> on rows
// (b << S2) + j of sub-block b (twiddle group blk * 2^S1 + b), i.e. exactly the two chained register passes,
// with the exchange between them on chip.  Inverse: phase B first, then phase A.
// ---------------------------------------------------------------------------
// FIN (the kernel that ends an inverse transform only): the InvFusion finishing step is compiled in
template <int S1, int S2, int POL, bool INVERSE, bool FIRST, bool FIN = false>
__global__ void __launch_bounds__(kNttThreads, FIN ? 2 : (INVERSE ? 4 : 3))
ntt_column2_kernel(const ModParams mp, const NttTables tbl, u64* __restrict__ data, size_t batch, u32 logn, u32 s0,
                   const u64* __restrict__ fin_c, u64 fin_scale) {
    constexpr u32 LR = S1 + S2, ROWS = 1u << LR, ELEMS = 4096u, COLS = ELEMS / ROWS, LC = 12 - LR;
    static_assert(S1 >= S2 && S1 <= 4 && COLS >= 16, "row segments of at least 128 bytes");
    __shared__ __align__(16) u64 sm[ELEMS];
    const u32 LG = logn - s0 - LR;                       // log2 g
    const size_t grp = blockIdx.x;                       // (polynomial, block, column group) triples
    const size_t pb = grp >> (LG - LC);                  // polynomial * 2^s0 + block
    const u32 c0 = (u32)(grp & (((size_t)1 << (LG - LC)) - 1u)) << LC;
    const u32 blk = (u32)(pb & (((size_t)1 << s0) - 1u));
    u64* __restrict__ g = data + (pb << (LG + LR)) + c0;
    (void)batch;

    auto phaseA = [&](bool from_global, bool to_global) {
        const u32 T0 = (1u << s0) + blk;
#pragma unroll 1
        for (u32 it = threadIdx.x; it < (ELEMS >> S1); it += kNttThreads) {
            const u32 cc = it & (COLS - 1u), r0 = it >> LC;          // r0 < 2^S2
            u64 v[1 << S1];
            // fused finishing step (c - x) * scale: the c words are requested before the butterflies, used after them
            u64 cv[FIN ? (1 << S1) : 1];
            if constexpr (FIN) {
                if (to_global) {
                    const u64* __restrict__ cf = fin_c + (pb << (LG + LR)) + c0;
#pragma unroll
                    for (int j = 0; j < (1 << S1); j++) cv[j] = __ldcs(cf + ((size_t)(r0 + ((u32)j << S2)) << LG) + cc);
                }
            }
            if (from_global) {
#pragma unroll
                for (int j = 0; j < (1 << S1); j++) v[j] = g[((size_t)(r0 + ((u32)j << S2)) << LG) + cc];
                if (FIRST && !INVERSE) {                             // raw caller data (see ntt_column_kernel)
                    const u64 limit = POL == POL_GOLD ? mp.q : mp.q4;
                    u64 any = 0;
#pragma unroll
                    for (int j = 0; j < (1 << S1); j++) any |= v[j];
                    if (__builtin_expect(any >= limit, 0)) {
#pragma unroll
                        for (int j = 0; j < (1 << S1); j++) v[j] = sanitize(v[j], limit, mp);
                    }
#pragma unroll
                    for (int j = 0; j < (1 << S1); j++) v[j] = to_working<POL>(v[j]);
                }
            } else {
#pragma unroll
                for (int j = 0; j < (1 << S1); j++) v[j] = sm[((r0 + ((u32)j << S2)) << LC) + cc];
            }
            if (!INVERSE) fwd_network<S1, POL>(v, tbl.fwd, T0, mp);
            else inv_network<S1, POL, FIRST>(v, tbl.inv, T0, (int)(LG + S2), tbl.n_inv, mp);
            if (to_global) {
                if constexpr (FIN) {
#pragma unroll
                    for (int j = 0; j < (1 << S1); j++) v[j] = field_mul(field_sub(cv[FIN ? j : 0], v[j], mp), fin_scale, mp);
                }
#pragma unroll
                for (int j = 0; j < (1 << S1); j++) g[((size_t)(r0 + ((u32)j << S2)) << LG) + cc] = v[j];
            } else {
#pragma unroll
                for (int j = 0; j < (1 << S1); j++) sm[((r0 + ((u32)j << S2)) << LC) + cc] = v[j];
            }
        }
    };
    auto phaseB = [&](bool from_global, bool to_global) {
#pragma unroll 1
        for (u32 it = threadIdx.x; it < (ELEMS >> S2); it += kNttThreads) {
            const u32 cc = it & (COLS - 1u), b = it >> LC;           // b < 2^S1
            const u32 T0 = (1u << (s0 + S1)) + (blk << S1) + b;
            u64 v[1 << S2];
            if (from_global) {
#pragma unroll
                for (int j = 0; j < (1 << S2); j++) v[j] = g[((size_t)((b << S2) + (u32)j) << LG) + cc];
            } else {
#pragma unroll
                for (int j = 0; j < (1 << S2); j++) v[j] = sm[(((b << S2) + (u32)j) << LC) + cc];
            }
            if (!INVERSE) fwd_network<S2, POL>(v, tbl.fwd, T0, mp);
            else inv_network<S2, POL, false>(v, tbl.inv, T0, (int)LG, tbl.n_inv, mp);
            if (to_global) {
#pragma unroll
                for (int j = 0; j < (1 << S2); j++) g[((size_t)((b << S2) + (u32)j) << LG) + cc] = v[j];
            } else {
#pragma unroll
                for (int j = 0; j < (1 << S2); j++) sm[(((b << S2) + (u32)j) << LC) + cc] = v[j];
            }
        }
    };
    if (!INVERSE) {
        phaseA(true, false);
        __syncthreads();
        phaseB(false, true);
    } else {
        phaseB(true, false);
        __syncthreads();
        phaseA(false, true);
    }
}

// ---------------------------------------------------------------------------
// K3: result[i] = a[i] * b[i] mod q, exact for any u64 inputs
// (ntt.cpp:106-119).  16-byte accesses, result may alias a or b (same index).
// A CTA takes tiles of 256 * U pairs; a thread issues its 2U loads before the first product (streaming: every word is
// touched once) and stores after the last load, so exact aliasing is safe.  GOLD is a template parameter so that the
// modulus test is not in the loop.  tools/pointwise_variants.cu: U = 4 with streaming hints 6.4 TB/s against 6.16 for
// the one-pair-per-iteration form at the bench's 3 x 512 MiB working set.
// ---------------------------------------------------------------------------
constexpr int kPointwiseU = 4;
template <bool GOLD>
static __global__ void __launch_bounds__(256)
pointwise_mul_kernel(const ModParams mp, u64* result, const u64* a, const u64* b, size_t total) {
    const size_t pairs = total >> 1;
    const bool aligned = ((((size_t)result) | ((size_t)a) | ((size_t)b)) & 15u) == 0;
    auto mul = [&](u64 x, u64 y) { return GOLD ? gold_mul(x, y) : barrett128(x * y, __umul64hi(x, y), mp); };
    if (aligned) {
        const ulonglong2* a2 = reinterpret_cast<const ulonglong2*>(a);
        const ulonglong2* b2 = reinterpret_cast<const ulonglong2*>(b);
        ulonglong2* r2 = reinterpret_cast<ulonglong2*>(result);
        const size_t tile = (size_t)blockDim.x * kPointwiseU;
        for (size_t base = (size_t)blockIdx.x * tile; base < pairs; base += (size_t)gridDim.x * tile) {
            ulonglong2 x[kPointwiseU], y[kPointwiseU];
#pragma unroll
            for (int u = 0; u < kPointwiseU; u++) {
                const size_t p = base + (size_t)u * blockDim.x + threadIdx.x;
                if (p < pairs) { x[u] = __ldcs(a2 + p); y[u] = __ldcs(b2 + p); }
            }
#pragma unroll
            for (int u = 0; u < kPointwiseU; u++) {
                const size_t p = base + (size_t)u * blockDim.x + threadIdx.x;
                if (p < pairs) __stcs(r2 + p, make_ulonglong2(mul(x[u].x, y[u].x), mul(x[u].y, y[u].y)));
            }
        }
        if (blockIdx.x == 0 && threadIdx.x == 0 && (total & 1)) result[total - 1] = mul(a[total - 1], b[total - 1]);
    } else {
        const size_t stride = (size_t)gridDim.x * blockDim.x;
        for (size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += stride) result[p] = mul(a[p], b[p]);
    }
}

}  // namespace lsr
