// lsr_host.cpp -- host-side number theory and table construction (K0).
#include "lsr_host.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <random>

namespace lsr {
namespace host {

typedef unsigned __int128 u128;

u64 mulmod(u64 a, u64 b, u64 q) { return (u64)(((u128)a * b) % q); }

u64 powmod(u64 a, u64 e, u64 q) {
    u64 result = 1 % q;
    u64 base = a % q;
    for (; e; e >>= 1) {
        if (e & 1) result = mulmod(result, base, q);
        base = mulmod(base, base, q);
    }
    return result;
}

bool is_prime(u64 n) {
    if (n < 2) return false;
    static const u64 witnesses[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (u64 p : witnesses) {
        if (n == p) return true;
        if (n % p == 0) return false;
    }
    u64 odd = n - 1;
    int twos = 0;
    while (!(odd & 1)) { odd >>= 1; ++twos; }
    for (u64 a : witnesses) {
        u64 x = powmod(a, odd, n);
        if (x == 1 || x == n - 1) continue;
        bool witness = true;
        for (int i = 1; i < twos && witness; ++i) {
            x = mulmod(x, x, n);
            if (x == n - 1) witness = false;
        }
        if (witness) return false;
    }
    return true;
}

// SEAL util/numth.cpp try_minimal_primitive_root: the minimum over all
// primitive 2n-th roots, i.e. over root * (root^2)^j.  Any starting root gives
// the same minimum, so the search for one is deterministic here.
u64 min_primitive_root(u64 q, u64 two_n) {
    if (q < 3 || two_n < 2 || (two_n & (two_n - 1))) return 0;
    if ((q - 1) % two_n) return 0;
    if (!is_prime(q)) return 0;   // composite moduli: SEAL's table would depend on its RNG
    const u64 cofactor = (q - 1) / two_n;
    u64 root = 0;
    for (u64 cand = 2; cand < q && !root; ++cand) {
        const u64 r = powmod(cand, cofactor, q);
        if (powmod(r, two_n >> 1, q) == q - 1) root = r;
    }
    if (!root) return 0;
    const u64 step = mulmod(root, root, q);
    u64 smallest = root, walk = root;
    for (u64 i = 1; i < (two_n >> 1); ++i) {
        walk = mulmod(walk, step, q);
        smallest = std::min(smallest, walk);
    }
    return smallest;
}

bool ntt_params_ok(u64 q, uint32_t n) {
    // ntt.cpp:31,41 (n == 0, non power of two); SEAL NTTTables: 2 <= n <= 2^17;
    // SEAL Modulus: 2 <= q < 2^61
    if (n < 2 || (n & (n - 1)) || n > (1u << kMaxLogN)) return false;
    if (q < 2 || (q >> 61)) return false;
    return true;
}

bool ntt_friendly_prime(u64 q, uint32_t n) {
    if (!ntt_params_ok(q, n) || q < 3) return false;
    if ((q - 1) % (2ull * n)) return false;
    return is_prime(q);
}

static inline u64 shoup_quotient(u64 w, u64 q) { return (u64)(((u128)w << 64) / q); }
static void transpose_last_pass_impl(NttHostTables& out);
static void transpose_last_pass(NttHostTables& out) { transpose_last_pass_impl(out); }

static inline uint32_t bitrev(uint32_t x, uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t b = 0; b < bits; ++b) r |= ((x >> b) & 1u) << (bits - 1 - b);
    return r;
}

ModParams make_mod_params(u64 q, uint32_t logn) {
    ModParams mp{};
    mp.q = q;
    mp.nq = 0 - q;
    mp.q2 = 2 * q;
    mp.q4 = 4 * q;
    const u128 ratio = (~(u128)0) / q;     // floor((2^128 - 1)/q) == floor(2^128/q) for q not a power of two
    // q is odd here (NTT prime) or at least > 1 and not a power of two matters only for exactness of
    // the identity above; for a power of two q the value is off by one, which barrett128 tolerates
    mp.bar_lo = (u64)ratio;
    mp.bar_hi = (u64)(ratio >> 64);
    const int qbits = 64 - __builtin_clzll(q);
    mp.red_sh = qbits > 24 ? (uint32_t)(qbits - 24) : 0u;
    mp.red_c = (uint32_t)(((u128)1 << (32 + mp.red_sh)) / q);
    const u128 lim = (u128)1 << 63;
    mp.lazy_fwd = ((u128)(4 + 4 * logn) * q < lim) ? 1u : 0u;
    mp.lazy_inv = (((u128)q << (logn + 2)) < lim) ? 1u : 0u;
    mp.f64_ok = (q < ((u64)1 << 45)) ? 1u : 0u;
    mp.gold = (q == kGoldilocks) ? 1u : 0u;
    mp.qd = (double)q;
    mp.invq = 1.0 / (double)q;
    mp.q52 = (double)q + 4503599627370496.0;
    return mp;
}

// Negacyclic tables for a given primitive 2n-th root psi: out[i] = f(psi^(2 brv(i) + 1)).
// SEAL's NTTTables (q < 2^61, n <= 2^17, minimal psi) is the special case build_ntt_tables below; the
// quotient pipeline evaluates on the coset psi * <psi^2> of the m-th roots of unity with any psi, over
// Goldilocks too, and up to 2^kMaxEngineLogN points.
bool build_negacyclic_tables(u64 q, uint32_t n, u64 psi, NttHostTables& out) {
    if (n < 2 || (n & (n - 1)) || n > (1u << kMaxEngineLogN)) return false;
    if (q != kGoldilocks && (q < 3 || (q >> 61))) return false;
    if (psi == 0 || psi >= q || powmod(psi, n, q) != q - 1) return false;
    uint32_t logn = 0;
    while ((1u << logn) < n) ++logn;
    const u64 psi_inv = powmod(psi, q - 2, q);
    const u64 n_inv = powmod(n % q, q - 2, q);
    const bool small = !(q >> 61);
    auto pair = [&](u64 w) { return ulonglong2{w, small ? shoup_quotient(w, q) : 0ull}; };

    out.q = q; out.n = n; out.logn = logn; out.psi = psi;
    out.fwd.assign(n, ulonglong2{0, 0});
    out.inv.assign(n, ulonglong2{0, 0});

    // SEAL NTTTables::initialize: root_powers[brv(i)] = psi^i
    std::vector<u64> seal_inv(n, 0);   // inv_root_powers[brv(i-1)+1] = psi^-i
    u64 pw = 1, ipw = 1;
    for (uint32_t i = 0; i < n; ++i) {
        if (i > 0) {
            pw = mulmod(pw, psi, q);
            ipw = mulmod(ipw, psi_inv, q);
            seal_inv[bitrev(i - 1, logn) + 1] = ipw;
        }
        out.fwd[bitrev(i, logn)] = pair(pw);
    }
    seal_inv[0] = 1;
    // re-index the inverse table: stage with m groups reads SEAL slots
    // n-2m+1 .. n-m (transform_from_rev consumes them sequentially from 1)
    out.inv[0] = pair(1);
    for (uint32_t m = 1; m < n; m <<= 1) {
        for (uint32_t g = 0; g < m; ++g) {
            u64 w = seal_inv[n - 2 * m + 1 + g];
            if (m == 1) w = mulmod(w, n_inv, q);      // scalar folded into the last stage
            out.inv[m + g] = pair(w);
        }
    }
    out.n_inv = pair(n_inv);
    transpose_last_pass(out);
    return true;
}

bool build_ntt_tables(u64 q, uint32_t n, NttHostTables& out) {
    if (!ntt_params_ok(q, n)) return false;
    const u64 psi = min_primitive_root(q, 2ull * n);
    if (!psi) return false;
    return build_negacyclic_tables(q, n, psi, out);
}

// transposed copy of the unit-stride radix-16 pass (forward stages logn-4 .. logn-1):
// work item w, stage r, group t  <-  heap entry ((2^(logn-4) + w) << r) + t
static void transpose_last_pass_impl(NttHostTables& out) {
    const uint32_t n = out.n, logn = out.logn;
    out.fwd_last.clear(); out.inv_last.clear();
    if (logn > 4) {
        const uint32_t items = n >> 4, first = 1u << (logn - 4);
        out.fwd_last.assign((size_t)15 * items, ulonglong2{0, 0});
        out.inv_last.assign((size_t)15 * items, ulonglong2{0, 0});
        for (uint32_t w = 0; w < items; ++w)
            for (uint32_t r = 0; r < 4; ++r)
                for (uint32_t t = 0; t < (1u << r); ++t) {
                    const size_t dst = (size_t)((1u << r) - 1 + t) * items + w;
                    const size_t src = ((size_t)(first + w) << r) + t;
                    out.fwd_last[dst] = out.fwd[src];
                    out.inv_last[dst] = out.inv[src];
                }
    }
}

bool cyclic_params_ok(u64 q, uint32_t n, u64 omega) {
    if (n < 2 || (n & (n - 1)) || n > (1u << kMaxEngineLogN)) return false;
    if (q != kGoldilocks && (q < 3 || (q >> 61))) return false;
    if ((q - 1) % n) return false;
    if (omega == 0 || omega >= q) return false;
    return powmod(omega, n >> 1, q) == q - 1;      // primitive n-th root of unity (q prime is the caller's claim
                                                   // for Goldilocks; checked below for everything else)
}

// Cyclic transform of rust-api/lambda-snark/src/ntt.rs:117-201 on the same butterfly networks: the
// forward network with twiddle  w[m + g] = omega^(brv_{log m}(g) * n / 2m)  maps natural-order
// coefficients to f(omega^brv(i)) at index i (the negacyclic table is the same with 2*brv + 1 and
// psi = sqrt(omega)); the inverse network takes w^-1 and n^-1 in its last stage.
bool build_cyclic_tables(u64 q, uint32_t n, u64 omega, NttHostTables& out) {
    if (!cyclic_params_ok(q, n, omega)) return false;
    if (q != kGoldilocks && !is_prime(q)) return false;
    uint32_t logn = 0;
    while ((1u << logn) < n) ++logn;
    const u64 omega_inv = powmod(omega, q - 2, q);
    const u64 n_inv = powmod(n % q, q - 2, q);
    out.q = q; out.n = n; out.logn = logn; out.psi = omega;
    out.fwd.assign(n, ulonglong2{0, 0});
    out.inv.assign(n, ulonglong2{0, 0});
    std::vector<u64> pw(n / 2 + 1), ipw(n / 2 + 1);          // omega^e, omega^-e for e < n/2
    pw[0] = ipw[0] = 1;
    for (uint32_t e = 1; e <= n / 2; ++e) { pw[e] = mulmod(pw[e - 1], omega, q); ipw[e] = mulmod(ipw[e - 1], omega_inv, q); }
    const bool small = !(q >> 61);
    auto pair = [&](u64 w) { return ulonglong2{w, small ? shoup_quotient(w, q) : 0ull}; };
    out.fwd[0] = out.inv[0] = pair(1);
    uint32_t logm = 0;
    for (uint32_t m = 1; m < n; m <<= 1, ++logm) {
        for (uint32_t g = 0; g < m; ++g) {
            const uint32_t e = bitrev(g, logm) * (n / (2 * m));   // < n/2
            out.fwd[m + g] = pair(pw[e]);
            u64 wi = ipw[e];
            if (m == 1) wi = mulmod(wi, n_inv, q);               // scalar folded into the last inverse stage
            out.inv[m + g] = pair(wi);
        }
    }
    out.n_inv = pair(n_inv);
    transpose_last_pass(out);
    return true;
}

// ---------------------------------------------------------------------------
// utils.cpp:26-75 restated with identical floating-point operations
// (x87 80-bit long double, std::exp, std::ceil) so the table is bit-identical.
// ---------------------------------------------------------------------------
std::vector<u64> build_cdt(double sigma, size_t cap) {
    std::vector<u64> cdf;
    if (!(sigma > 0.0) || !std::isfinite(sigma)) return cdf;
    const long double s = static_cast<long double>(sigma);
    const long double s2 = s * s;
    long double bound = std::ceil(12.0L * s);
    if (bound < 8.0L) bound = 8.0L;
    const size_t top = static_cast<size_t>(bound);
    if (top + 1 > cap) return cdf;

    std::vector<long double> wt(top + 1, 0.0L);
    long double total = 0.0L;
    for (size_t k = 0; k <= top; ++k) {
        const long double kk = static_cast<long double>(k) * static_cast<long double>(k);
        long double w = std::exp(-kk / (2.0L * s2));
        if (k > 0) w *= 2.0L;
        wt[k] = w;
        total += w;
    }
    const u64 kMax = std::numeric_limits<u64>::max();
    cdf.assign(top + 1, 0);
    if (total == 0.0L) {
        cdf[top] = kMax;
        return cdf;
    }
    const long double scale = static_cast<long double>(kMax) / total;
    long double run = 0.0L;
    for (size_t k = 0; k <= top; ++k) {
        run += wt[k];
        const long double val = run * scale;
        if (val >= static_cast<long double>(kMax)) cdf[k] = kMax;
        else if (val <= 0.0L) cdf[k] = 0;
        else cdf[k] = static_cast<u64>(val);
    }
    cdf.back() = kMax;
    return cdf;
}

int64_t cdt_sample(const std::vector<u64>& cdf, u64 u1, u64 u2) {
    // smallest k with cdf[k] >= u1 == number of entries strictly below u1
    // (the table is non-decreasing and ends at 2^64-1)
    uint32_t below = 0;
    for (u64 entry : cdf) below += (uint32_t)(entry < u1);
    const int64_t mag = (int64_t)below;
    const int64_t neg = -(int64_t)((u2 & 1ull) & (u64)(below != 0));
    return (mag ^ neg) - neg;
}

// ---------------------------------------------------------------------------
static inline uint32_t rotl(uint32_t v, int c) { return (v << c) | (v >> (32 - c)); }
static inline void quarter(uint32_t& a, uint32_t& b, uint32_t& c, uint32_t& d) {
    a += b; d ^= a; d = rotl(d, 16);
    c += d; b ^= c; b = rotl(b, 12);
    a += b; d ^= a; d = rotl(d, 8);
    c += d; b ^= c; b = rotl(b, 7);
}

void chacha_block(const uint32_t key[8], uint32_t w12, uint32_t w13, uint32_t w14, uint32_t w15,
                  uint32_t out[16]) {
    uint32_t init[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    std::memcpy(init + 4, key, 32);
    init[12] = w12; init[13] = w13; init[14] = w14; init[15] = w15;
    uint32_t s[16];
    std::memcpy(s, init, sizeof(s));
    for (int round = 0; round < kChaChaRounds; round += 2) {
        quarter(s[0], s[4], s[8], s[12]);
        quarter(s[1], s[5], s[9], s[13]);
        quarter(s[2], s[6], s[10], s[14]);
        quarter(s[3], s[7], s[11], s[15]);
        quarter(s[0], s[5], s[10], s[15]);
        quarter(s[1], s[6], s[11], s[12]);
        quarter(s[2], s[7], s[8], s[13]);
        quarter(s[3], s[4], s[9], s[14]);
    }
    for (int i = 0; i < 16; ++i) out[i] = s[i] + init[i];
}

void load_key(const uint8_t seed32[32], uint32_t key[8]) {
    for (int i = 0; i < 8; ++i) {
        key[i] = (uint32_t)seed32[4 * i] | ((uint32_t)seed32[4 * i + 1] << 8) |
                 ((uint32_t)seed32[4 * i + 2] << 16) | ((uint32_t)seed32[4 * i + 3] << 24);
    }
}

bool os_entropy(uint8_t* out, size_t len) {
    try {
        std::random_device rd;     // same source the reference sampler uses (utils.cpp:138)
        for (size_t i = 0; i < len; i += 4) {
            const uint32_t v = rd();
            for (size_t b = 0; b < 4 && i + b < len; ++b) out[i + b] = (uint8_t)(v >> (8 * b));
        }
        return true;
    } catch (...) {
        return false;
    }
}

u64 plain_modulus(u64 q) {
    const u64 cap = 1ull << 20;
    u64 rest = q - 1;
    std::vector<u64> divs{1};
    for (u64 p = 2; p <= cap && p * p <= rest; ++p) {
        if (rest % p) continue;
        const size_t base = divs.size();
        u64 pk = 1;
        while (rest % p == 0) {
            rest /= p;
            if (pk <= cap) pk *= p;
            if (pk <= cap) {
                for (size_t i = 0; i < base; ++i) {
                    if (divs[i] * pk <= cap) divs.push_back(divs[i] * pk);
                }
            }
        }
    }
    if (rest > 1 && rest <= cap) {
        const size_t base = divs.size();
        for (size_t i = 0; i < base; ++i) {
            if (divs[i] * rest <= cap) divs.push_back(divs[i] * rest);
        }
    }
    return *std::max_element(divs.begin(), divs.end());
}

namespace {
struct Stream {
    const uint32_t* key;
    uint32_t w14, dom;
    u64 have_block = ~0ull;
    uint32_t buf[16];
    u64 draw(u64 idx) {
        const u64 blk = idx >> 3;
        if (blk != have_block) {
            chacha_block(key, (uint32_t)blk, (uint32_t)(blk >> 32), w14, dom, buf);
            have_block = blk;
        }
        const uint32_t w = (uint32_t)(idx & 7);
        return (u64)buf[2 * w] | ((u64)buf[2 * w + 1] << 32);
    }
};
}  // namespace

void uniform_poly(const uint32_t key[8], uint32_t stream_id, u64 q, uint32_t n, u64* out) {
    Stream st{key, stream_id, kDomMatrix};
    const int qbits = 64 - __builtin_clzll(q);
    const u64 mask = qbits >= 64 ? ~0ull : ((1ull << qbits) - 1);
    u64 idx = 0;
    for (uint32_t got = 0; got < n;) {
        const u64 cand = st.draw(idx++) & mask;
        if (cand < q) out[got++] = cand;
    }
}

void gaussian_poly(const uint32_t key[8], uint32_t stream_id, const std::vector<u64>& cdf, u64 q,
                   uint32_t n, u64* out) {
    Stream st{key, stream_id, kDomTrap};
    for (uint32_t i = 0; i < n; ++i) {
        const u64 u1 = st.draw(2ull * i);
        const u64 u2 = st.draw(2ull * i + 1);
        const int64_t v = cdt_sample(cdf, u1, u2);
        out[i] = v < 0 ? q - (u64)(-v) : (u64)v;
    }
}

}  // namespace host
}  // namespace lsr
