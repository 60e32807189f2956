// lsr_common.h -- types shared by the host runtime and the sm_100a kernels.
//
// Everything the kernels need about a modulus or an NTT table travels BY VALUE
// in these structs (kernel parameters live in the constant bank, so q, 2^64-q,
// the Barrett ratio etc. are uniform constant operands, not loads).
#pragma once
#include <cstddef>
#include <cstdint>

#include <vector_types.h>   // ulonglong2 (CUDA toolkit header, host-safe)

namespace lsr {

typedef unsigned long long u64;
typedef unsigned int u32;

constexpr u64 kDefaultModulusSmall = 17592169062401ULL;  // reference r1cs.rs:527, 2-adicity 13
constexpr u64 kDefaultModulusLarge = 17592180539393ULL;  // 44-bit, 2-adicity 18
constexpr int kMaxLogN = 17;                              // SEAL_POLY_MOD_DEGREE_MAX = 131072
constexpr int kMaxEngineLogN = 24;                        // cyclic transforms of the quotient pipeline (3 column passes + tile)
constexpr int kChaChaRounds = 8;

constexpr u32 kDomMatrix = 0x01000000u;
constexpr u32 kDomTrap   = 0x02000000u;
constexpr u32 kDomCommit = 0x03000000u;
constexpr u32 kDomCommitFine = 0x100u;    // refinement blocks of the commitment randomness
constexpr u32 kDomSample = 0x05000000u;

// Modulus constants.  "lazy" policies: butterflies carry unreduced values and
// only the last stage reduces; valid while the worst-case growth fits in 63
// bits.  Otherwise the SEAL-style guarded Harvey
// butterfly ([0,4q) forward, [0,2q) inverse) is used.
struct ModParams {
    u64 q;
    u64 nq;        // 2^64 - q
    u64 q2;        // 2q
    u64 q4;        // 4q
    u64 bar_lo;    // floor(2^128 / q), low word
    u64 bar_hi;    //                   high word
    u32 red_sh;    // reduce_small(v): est = umulhi32(v >> red_sh, red_c), v < 2^7 q
    u32 red_c;
    u32 lazy_fwd;  // (4 + 4 logn) q < 2^63
    u32 lazy_inv;  // 2^(logn+2) q < 2^63
    u32 f64_ok;    // q < 2^45: the FP64-pipe butterfly is exact
    u32 gold;      // q = 2^64 - 2^32 + 1 (the reference's NTT_MODULUS): POL_GOLD arithmetic
    double qd;     // (double)q
    double invq;   // RN(1 / q)
    double q52;    // q + 2^52
};

// Arithmetic policy of the NTT kernels (template parameter POL):
//   POL_GUARD  SEAL's Harvey butterfly on u64, any q < 2^61
//   POL_LAZY   u64 residues, truncated Shoup product, no reduction inside the butterfly
//   POL_F64    residues held as doubles (balanced representatives), exact modular product by
//              fma error-free multiplication; q < 2^45.  The FP64 pipe of B200 issues 64
//              lanes/clk/SM and a butterfly costs 8 of its instructions, against 5 half-rate
//              IMAD.WIDE + 4 IMAD + 8 ALU for POL_LAZY: measured 2.0x (tools/fp64_microbench.cu)
//   POL_GOLD   q = 2^64 - 2^32 + 1 (Goldilocks, lambda-snark-core/src/lib.rs:58): canonical u64 residues,
//              128-bit product folded with 2^64 = 2^32 - 1, 2^96 = -1 (mod q); cyclic transforms of the
//              quotient pipeline (rust-api/lambda-snark/src/ntt.rs)
enum : int { POL_GUARD = 0, POL_LAZY = 1, POL_F64 = 2, POL_GOLD = 3 };
constexpr u64 kGoldilocks = 0xFFFFFFFF00000001ULL;

// Device twiddle tables, both indexed [m + group] for the stage with m groups
// (m = 1, 2, 4, ..., n/2): .x = w, .y = floor(w * 2^64 / q).
//   fwd[m + g] = SEAL root_powers[m + g]          = psi^brv(m + g)
//   inv[m + g] = SEAL inv_root_powers[n - 2m + 1 + g]  (SEAL consumes them
//                sequentially from index 1; re-indexed so both directions
//                address twiddles the same way); inv[1] is pre-multiplied by
//                n^-1 (SEAL folds the scalar into the last stage the same way)
//   fwd_last / inv_last: the twiddles of the unit-stride radix-16 pass (forward
//                stages logn-4 .. logn-1), transposed to [15][n/16] so that work
//                item w reads entry [(2^r - 1 + t)][w]: coalesced across a warp
// For POL_F64 a second instance of this struct holds the same tables as bit patterns of
// doubles: .x = (double)w, .y = RN(w / q).
struct NttTables {
    const ulonglong2* fwd;
    const ulonglong2* inv;
    const ulonglong2* fwd_last;
    const ulonglong2* inv_last;
    ulonglong2 n_inv;   // (n^-1, shoup)
    ulonglong2 head_fwd[32];   // fwd[0..31] / inv[0..31] by value: the twiddles of the
    ulonglong2 head_inv[32];   // first forward pass (up to 5 stages) are uniform over the whole grid
};

// Optional fusions around an INVERSE transform (the quotient pipeline, lsr_quotient.cu):
//   mul       the transform's input is data[i] * mul[i] (the pointwise product of two evaluation vectors) instead of data[i]
//   dst       where the FIRST kernel of the transform writes (nullptr: in place); later kernels work in place on dst
//   fin_c     the LAST kernel stores (fin_c[i] - x) * fin_scale instead of x
// All-zero: the plain transform.  Forward kernels ignore it.
struct InvFusion {
    const u64* mul;
    u64* dst;
    const u64* fin_c;
    u64 fin_scale;
};

}  // namespace lsr
