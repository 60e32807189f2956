// lsr_host.h -- host-side number theory and table construction (K0).
//
// Product code: builds what seal::util::NTTTables builds for the reference
// (cpp-core/src/ntt.cpp:55-59), the CDT of cpp-core/src/utils.cpp:26-75, and
// the Module-LWE context material of DESIGN.md section 3.  Independent of the
// oracle under oracle/ (which exists to check this code, not to feed it).
#pragma once
#include <cstdint>
#include <vector>

#include "lsr_common.h"

namespace lsr {
namespace host {

u64 mulmod(u64 a, u64 b, u64 q);
u64 powmod(u64 a, u64 e, u64 q);
bool is_prime(u64 q);
// smallest primitive (2n)-th root of unity mod prime q, 0 if none
u64 min_primitive_root(u64 q, u64 two_n);
// (q, n) acceptable to ntt_context_create (SEAL Modulus + NTTTables rules)
bool ntt_params_ok(u64 q, uint32_t n);
bool ntt_friendly_prime(u64 q, uint32_t n);

struct NttHostTables {
    u64 q = 0;
    uint32_t n = 0, logn = 0;
    u64 psi = 0;
    std::vector<ulonglong2> fwd;   // [n], index m+g
    std::vector<ulonglong2> inv;   // [n], index m+g, inv[1] scaled by n^-1
    std::vector<ulonglong2> fwd_last, inv_last;   // [15][n/16] transposed last-pass twiddles (logn > 4)
    ulonglong2 n_inv{0, 0};
};
bool build_ntt_tables(u64 q, uint32_t n, NttHostTables& out);
// same layout for a caller-chosen primitive 2n-th root psi; q < 2^61 or Goldilocks, n <= 2^kMaxEngineLogN
bool build_negacyclic_tables(u64 q, uint32_t n, u64 psi, NttHostTables& out);
// cyclic (X^n - 1) tables for a given primitive n-th root omega; q < 2^61 prime or Goldilocks
bool cyclic_params_ok(u64 q, uint32_t n, u64 omega);
bool build_cyclic_tables(u64 q, uint32_t n, u64 omega, NttHostTables& out);
ModParams make_mod_params(u64 q, uint32_t logn);

// utils.cpp:26-75 (long double CDT).  Empty vector on invalid sigma.
std::vector<u64> build_cdt(double sigma, size_t cap = 32768);

// ChaCha block, kChaChaRounds rounds
void chacha_block(const uint32_t key[8], uint32_t w12, uint32_t w13, uint32_t w14, uint32_t w15,
                  uint32_t out[16]);
void load_key(const uint8_t seed32[32], uint32_t key[8]);
bool os_entropy(uint8_t* out, size_t len);

// largest divisor of q-1 not exceeding 2^20 (plaintext modulus p; delta=(q-1)/p)
u64 plain_modulus(u64 q);

// branch-free CDT select + sign (utils.cpp:95-121), host copy for context setup
int64_t cdt_sample(const std::vector<u64>& cdf, u64 u1, u64 u2);

// uniform residues in [0,q) by rejection from stream (key, w14=stream_id, dom=kDomMatrix)
void uniform_poly(const uint32_t key[8], uint32_t stream_id, u64 q, uint32_t n, u64* out);
// CDT-sampled small polynomial as residues, stream (key, w14=stream_id, dom=kDomTrap),
// coefficient i uses 64-bit draws (2i, 2i+1)
void gaussian_poly(const uint32_t key[8], uint32_t stream_id, const std::vector<u64>& cdf, u64 q,
                   uint32_t n, u64* out);

}  // namespace host
}  // namespace lsr
