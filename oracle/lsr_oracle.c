/*
 * lsr_oracle.c -- CPU ORACLE (test infrastructure, see lsr_oracle.h header).
 *
 * Plain C restatement of the reference's algorithm for the prover hot path.
 * Every function cites the reference file:line (relative to /root/reference)
 * or the SEAL 4.1 source file whose published algorithm it restates.
 *
 * Parity: sampler PINNED (oracle/_ref, reference utils.cpp); NTT pinned to
 * SURVEY 8c KATs / closed form only ("parity unpinned" versus a SEAL binary);
 * commitment bytes "parity unpinned" (reference output is randomised, F2).
 */
#include "lsr_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------ */
/* modular arithmetic: SEAL util/uintarithsmallmod.h (multiply_uint_mod is   */
/* exact (a*b) mod q for all u64 a,b; used by ntt.cpp:116-118)               */
/* ------------------------------------------------------------------------ */
uint64_t lsro_mulmod(uint64_t a, uint64_t b, uint64_t q) {
    return (uint64_t)(((u128)a * b) % q);
}

uint64_t lsro_powmod(uint64_t a, uint64_t e, uint64_t q) {
    uint64_t r = 1 % q;
    a %= q;
    while (e) {
        if (e & 1) r = lsro_mulmod(r, a, q);
        a = lsro_mulmod(a, a, q);
        e >>= 1;
    }
    return r;
}

/* deterministic Miller-Rabin for 64-bit integers */
int lsro_is_prime(uint64_t n) {
    static const uint64_t bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    if (n < 2) return 0;
    for (size_t i = 0; i < sizeof(bases) / sizeof(bases[0]); i++) {
        if (n % bases[i] == 0) return n == bases[i];
    }
    uint64_t d = n - 1;
    int r = 0;
    while ((d & 1) == 0) { d >>= 1; r++; }
    for (size_t i = 0; i < sizeof(bases) / sizeof(bases[0]); i++) {
        uint64_t x = lsro_powmod(bases[i], d, n);
        if (x == 1 || x == n - 1) continue;
        int composite = 1;
        for (int j = 1; j < r; j++) {
            x = lsro_mulmod(x, x, n);
            if (x == n - 1) { composite = 0; break; }
        }
        if (composite) return 0;
    }
    return 1;
}

/*
 * SEAL util/numth.cpp try_minimal_primitive_root(degree = 2n): take any
 * primitive degree-th root r, walk r*(r^2)^j over all degree/2 odd powers and
 * keep the numerically smallest.  SEAL finds r by random trials; the minimum
 * over the full coset does not depend on which r was found, so a deterministic
 * search gives the identical table.  Composite q is rejected here (SEAL's
 * answer would depend on its RNG): see DESIGN.md "deviations".
 */
uint64_t lsro_min_primitive_root(uint64_t q, uint64_t two_n) {
    if (q < 3 || two_n < 2 || (two_n & (two_n - 1)) != 0) return 0;
    if ((q - 1) % two_n != 0) return 0;
    if (!lsro_is_prime(q)) return 0;
    uint64_t quot = (q - 1) / two_n;
    uint64_t root = 0;
    for (uint64_t g = 2; g < q; g++) {
        uint64_t r = lsro_powmod(g, quot, q);
        /* is_primitive_root: r^(degree/2) == -1 */
        if (lsro_powmod(r, two_n / 2, q) == q - 1) { root = r; break; }
    }
    if (!root) return 0;
    uint64_t gsq = lsro_mulmod(root, root, q);
    uint64_t cur = root, best = root;
    for (uint64_t i = 0; i < two_n / 2; i++) {
        if (cur < best) best = cur;
        cur = lsro_mulmod(cur, gsq, q);
    }
    return best;
}

/* ------------------------------------------------------------------------ */
/* NTT tables: SEAL util/ntt.cpp NTTTables::initialize                       */
/* ------------------------------------------------------------------------ */
struct lsro_ntt {
    uint64_t q;
    uint32_t n, logn;
    uint64_t psi, psi_inv;
    uint64_t n_inv, n_inv_shoup;
    uint64_t *rp, *rps;   /* root_powers_[brv(i)] = psi^i, and floor(w*2^64/q) */
    uint64_t *irp, *irps; /* inv_root_powers_[brv(i-1)+1] = psi^-i              */
};

static uint32_t brv(uint32_t x, uint32_t bits) {
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* MultiplyUIntModOperand::set_quotient: floor(operand * 2^64 / q) */
static uint64_t shoup(uint64_t w, uint64_t q) { return (uint64_t)((((u128)w) << 64) / q); }

lsro_ntt *lsro_ntt_create(uint64_t q, uint32_t n) {
    /* ntt.cpp:31 n == 0; :41 non power of two; SEAL Modulus: q < 2^61, q != 1;
       NTTTables: 1 <= log2 n <= 17; throws if no primitive 2n-th root.        */
    if (n == 0 || (n & (n - 1)) != 0) return NULL;
    uint32_t logn = 0;
    while ((1u << logn) < n) logn++;
    if (logn < 1 || logn > 17) return NULL;
    if (q < 2 || (q >> 61) != 0) return NULL;
    uint64_t psi = lsro_min_primitive_root(q, 2ull * n);
    if (!psi) return NULL;

    lsro_ntt *c = (lsro_ntt *)calloc(1, sizeof(*c));
    if (!c) return NULL;
    c->q = q; c->n = n; c->logn = logn; c->psi = psi;
    c->psi_inv = lsro_powmod(psi, q - 2, q);
    c->rp = (uint64_t *)malloc(sizeof(uint64_t) * n);
    c->rps = (uint64_t *)malloc(sizeof(uint64_t) * n);
    c->irp = (uint64_t *)malloc(sizeof(uint64_t) * n);
    c->irps = (uint64_t *)malloc(sizeof(uint64_t) * n);
    if (!c->rp || !c->rps || !c->irp || !c->irps) { lsro_ntt_free(c); return NULL; }

    uint64_t power = psi;
    for (uint32_t i = 1; i < n; i++) {
        c->rp[brv(i, logn)] = power;
        power = lsro_mulmod(power, psi, q);
    }
    c->rp[0] = 1;
    power = c->psi_inv;
    for (uint32_t i = 1; i < n; i++) {
        c->irp[brv(i - 1, logn) + 1] = power;
        power = lsro_mulmod(power, c->psi_inv, q);
    }
    c->irp[0] = 1;
    for (uint32_t i = 0; i < n; i++) {
        c->rps[i] = shoup(c->rp[i], q);
        c->irps[i] = shoup(c->irp[i], q);
    }
    c->n_inv = lsro_powmod(n % q, q - 2, q);
    c->n_inv_shoup = shoup(c->n_inv, q);
    return c;
}

void lsro_ntt_free(lsro_ntt *c) {
    if (!c) return;
    free(c->rp); free(c->rps); free(c->irp); free(c->irps);
    free(c);
}

uint64_t lsro_ntt_psi(const lsro_ntt *c) { return c ? c->psi : 0; }
uint64_t lsro_ntt_modulus(const lsro_ntt *c) { return c ? c->q : 0; }
uint32_t lsro_ntt_degree(const lsro_ntt *c) { return c ? c->n : 0; }
const uint64_t *lsro_ntt_table(const lsro_ntt *c, int which) {
    if (!c) return NULL;
    switch (which) {
        case 0: return c->rp;
        case 1: return c->rps;
        case 2: return c->irp;
        case 3: return c->irps;
        default: return NULL;
    }
}

/* SEAL multiply_uint_mod_lazy: result in [0, 2q) for any 64-bit y */
static inline uint64_t mul_root(uint64_t y, uint64_t w, uint64_t ws, uint64_t q) {
    uint64_t hi = (uint64_t)(((u128)y * ws) >> 64);
    return y * w - hi * q;
}

/*
 * SEAL util/dwthandler.h transform_to_rev with the lazy Arithmetic of
 * util/ntt.h (guard: x >= 2q ? x-2q : x; add: a+b; sub: a+2q-b), followed by
 * the final correction of ntt_negacyclic_harvey (ntt.cpp:84 calls it).
 */
static void fwd_one(const lsro_ntt *c, uint64_t *x) {
    const uint64_t q = c->q, two_q = 2 * q;
    const uint32_t n = c->n;
    uint32_t gap = n >> 1;
    uint32_t root_idx = 0;
    for (uint32_t m = 1; m < n; m <<= 1) {
        uint32_t offset = 0;
        for (uint32_t i = 0; i < m; i++) {
            ++root_idx;
            const uint64_t w = c->rp[root_idx], ws = c->rps[root_idx];
            uint64_t *px = x + offset, *py = px + gap;
            for (uint32_t j = 0; j < gap; j++) {
                uint64_t u = px[j] >= two_q ? px[j] - two_q : px[j];
                uint64_t v = mul_root(py[j], w, ws, q);
                px[j] = u + v;
                py[j] = u + two_q - v;
            }
            offset += gap << 1;
        }
        gap >>= 1;
    }
    for (uint32_t i = 0; i < n; i++) {
        uint64_t v = x[i];
        if (v >= two_q) v -= two_q;
        if (v >= q) v -= q;
        x[i] = v;
    }
}

/*
 * SEAL transform_from_rev (Gentleman-Sande), scalar n^-1 folded into the last
 * stage, then inverse_ntt_negacyclic_harvey's single conditional subtract
 * (ntt.cpp:99 calls it).
 */
static void inv_one(const lsro_ntt *c, uint64_t *x) {
    const uint64_t q = c->q, two_q = 2 * q;
    const uint32_t n = c->n;
    uint32_t gap = 1;
    uint32_t root_idx = 0;
    uint32_t m = n >> 1;
    for (; m > 1; m >>= 1) {
        uint32_t offset = 0;
        for (uint32_t i = 0; i < m; i++) {
            ++root_idx;
            const uint64_t w = c->irp[root_idx], ws = c->irps[root_idx];
            uint64_t *px = x + offset, *py = px + gap;
            for (uint32_t j = 0; j < gap; j++) {
                uint64_t u = px[j], v = py[j];
                uint64_t s = u + v;
                px[j] = s >= two_q ? s - two_q : s;
                py[j] = mul_root(u + two_q - v, w, ws, q);
            }
            offset += gap << 1;
        }
        gap <<= 1;
    }
    {
        ++root_idx;
        const uint64_t w = c->irp[root_idx];
        const uint64_t sw = lsro_mulmod(w, c->n_inv, q);   /* mul_root_scalar */
        const uint64_t sws = shoup(sw, q);
        uint64_t *px = x, *py = x + gap;
        for (uint32_t j = 0; j < gap; j++) {
            uint64_t u = px[j] >= two_q ? px[j] - two_q : px[j];
            uint64_t v = py[j];
            uint64_t s = u + v;
            if (s >= two_q) s -= two_q;
            px[j] = mul_root(s, c->n_inv, c->n_inv_shoup, q);
            py[j] = mul_root(u + two_q - v, sw, sws, q);
        }
    }
    for (uint32_t i = 0; i < n; i++) {
        if (x[i] >= q) x[i] -= q;
    }
}

int lsro_ntt_forward(const lsro_ntt *c, uint64_t *x, uint32_t n) {
    if (!c || !x || n != c->n) return -1;      /* ntt.cpp:81 */
    fwd_one(c, x);
    return 0;
}

int lsro_ntt_inverse(const lsro_ntt *c, uint64_t *x, uint32_t n) {
    if (!c || !x || n != c->n) return -1;      /* ntt.cpp:96 */
    inv_one(c, x);
    return 0;
}

void lsro_ntt_mul_pointwise(const lsro_ntt *c, uint64_t *r, const uint64_t *a,
                            const uint64_t *b, uint32_t n) {
    if (!c || !r || !a || !b) return;          /* ntt.cpp:113: silent */
    for (uint32_t i = 0; i < n; i++) r[i] = lsro_mulmod(a[i], b[i], c->q);
}

int lsro_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

int lsro_ntt_forward_batch(const lsro_ntt *c, uint64_t *x, size_t batch, int threads) {
    if (!c || !x) return -1;
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long b = 0; b < (long)batch; b++) fwd_one(c, x + (size_t)b * c->n);
    return 0;
}

int lsro_ntt_inverse_batch(const lsro_ntt *c, uint64_t *x, size_t batch, int threads) {
    if (!c || !x) return -1;
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long b = 0; b < (long)batch; b++) inv_one(c, x + (size_t)b * c->n);
    return 0;
}

void lsro_ntt_mul_pointwise_batch(const lsro_ntt *c, uint64_t *r, const uint64_t *a,
                                  const uint64_t *b, size_t total, int threads) {
    if (!c || !r || !a || !b) return;
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (long i = 0; i < (long)total; i++) r[i] = lsro_mulmod(a[i], b[i], c->q);
}

/* ------------------------------------------------------------------------ */
/* discrete Gaussian: utils.cpp:24-75 (build_cdf), :95-121 (sample_single)   */
/* ------------------------------------------------------------------------ */
size_t lsro_cdt_build(double sigma, uint64_t *cdf, size_t cap) {
    if (!(sigma > 0.0) || !isfinite(sigma) || !cdf) return 0;   /* utils.cpp:133 */
    const long double sigma_ld = (long double)sigma;
    const long double sigma_sq = sigma_ld * sigma_ld;
    long double bound = ceill(12.0L * sigma_ld);                /* kTailCutoff */
    if (bound < 8.0L) bound = 8.0L;
    const size_t max_index = (size_t)bound;
    if (max_index + 1 > cap) return 0;

    long double *weights = (long double *)malloc(sizeof(long double) * (max_index + 1));
    if (!weights) return 0;
    long double sum = 0.0L;
    for (size_t k = 0; k <= max_index; k++) {
        const long double kk = (long double)k * (long double)k;
        const long double exponent = -kk / (2.0L * sigma_sq);
        long double weight = expl(exponent);
        if (k > 0) weight *= 2.0L;
        weights[k] = weight;
        sum += weight;
    }
    memset(cdf, 0, sizeof(uint64_t) * (max_index + 1));
    if (sum == 0.0L) {
        cdf[max_index] = UINT64_MAX;
        free(weights);
        return max_index + 1;
    }
    const long double scale = (long double)UINT64_MAX / sum;
    long double cumulative = 0.0L;
    for (size_t k = 0; k <= max_index; k++) {
        cumulative += weights[k];
        long double value = cumulative * scale;
        if (value >= (long double)UINT64_MAX) cdf[k] = UINT64_MAX;
        else if (value <= 0.0L) cdf[k] = 0;
        else cdf[k] = (uint64_t)value;
    }
    cdf[max_index] = UINT64_MAX;
    free(weights);
    return max_index + 1;
}

int64_t lsro_cdt_sample(const uint64_t *cdf, size_t count, uint64_t u1, uint64_t u2) {
    uint32_t chosen = (uint32_t)(count - 1);
    uint64_t found = 0;
    for (size_t k = 0; k < count; k++) {
        const uint64_t ge = (uint64_t)(cdf[k] >= u1);
        const uint64_t sel = ge & (1ULL ^ found);
        const uint32_t mask32 = (uint32_t)(-(int32_t)sel);
        chosen = (chosen & ~mask32) | ((uint32_t)k & mask32);
        found |= sel;
    }
    const uint64_t sign_bit = u2 & 1ULL;
    const uint64_t nonzero = (uint64_t)(chosen != 0);
    const uint64_t sign_mask = sign_bit & nonzero;
    const int64_t magnitude = (int64_t)chosen;
    const int64_t mask = -(int64_t)sign_mask;
    return (magnitude & ~mask) | ((-magnitude) & mask);
}

/* ------------------------------------------------------------------------ */
/* ChaCha (RFC 7539 quarter round), LSRO_CHACHA_ROUNDS rounds                */
/* replaces std::random_device of utils.cpp:77-93 with a reproducible stream */
/* ------------------------------------------------------------------------ */
#define ROTL32(v, c) (((v) << (c)) | ((v) >> (32 - (c))))
#define QR(a, b, c, d)                                   \
    a += b; d ^= a; d = ROTL32(d, 16);                   \
    c += d; b ^= c; b = ROTL32(b, 12);                   \
    a += b; d ^= a; d = ROTL32(d, 8);                    \
    c += d; b ^= c; b = ROTL32(b, 7);

void lsro_chacha_block(const uint32_t key[8], uint32_t w12, uint32_t w13,
                       uint32_t w14, uint32_t w15, uint32_t out[16]) {
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u,
                       key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                       w12, w13, w14, w15};
    uint32_t x[16];
    memcpy(x, in, sizeof(x));
    for (int r = 0; r < LSRO_CHACHA_ROUNDS; r += 2) {
        QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13])
        QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
        QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12])
        QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + in[i];
}

/* ------------------------------------------------------------------------ */
/* Module-LWE commitment                               */
/* ------------------------------------------------------------------------ */
#define LSRO_Q0 17592169062401ULL   /* r1cs.rs:527 NTT_FRIENDLY_MODULUS, 2-adicity 13 */
#define LSRO_Q1 17592180539393ULL   /* 44-bit, 2-adicity 18: n in (4096, 131072]     */
#define DOM_MATRIX 0x01000000u
#define DOM_TRAP   0x02000000u
#define DOM_COMMIT 0x03000000u
#define DOM_SAMPLE 0x05000000u

struct lsro_lwe {
    uint64_t q, p, delta;
    uint32_t n, k;
    double sigma;
    uint32_t key[8];
    lsro_ntt *ntt;
    uint64_t *cdf; size_t cdf_n;
    uint64_t *A;   /* [k][k][n], NTT domain */
    uint64_t *zh;  /* [k-1][n], NTT domain  */
};

static int ntt_friendly(uint64_t q, uint32_t n) {
    if (q < 3 || (q >> 61) != 0) return 0;
    if ((q - 1) % (2ull * n) != 0) return 0;
    return lsro_is_prime(q);
}

/* largest divisor of q-1 that is <= 2^20 */
static uint64_t plain_modulus(uint64_t q) {
    const uint64_t cap = 1ull << 20;
    uint64_t m = q - 1;
    uint64_t primes[64]; int exps[64]; int np = 0;
    for (uint64_t d = 2; d <= cap && d * d <= m; d++) {
        if (m % d == 0) {
            primes[np] = d; exps[np] = 0;
            while (m % d == 0) { m /= d; exps[np]++; }
            np++;
        }
    }
    if (m > 1 && m <= cap) { primes[np] = m; exps[np] = 1; np++; }
    /* enumerate divisors of the cap-smooth part */
    uint64_t best = 1;
    uint64_t *divs = (uint64_t *)malloc(sizeof(uint64_t) * (1u << 20));
    size_t nd = 1; divs[0] = 1;
    for (int i = 0; i < np; i++) {
        size_t cur = nd;
        uint64_t pw = 1;
        for (int e = 1; e <= exps[i]; e++) {
            pw *= primes[i];
            if (pw > cap) break;
            for (size_t j = 0; j < cur; j++) {
                uint64_t v = divs[j] * pw;
                if (v <= cap) { divs[nd++] = v; if (v > best) best = v; }
            }
        }
    }
    free(divs);
    return best;
}

static inline uint64_t to_residue(int64_t v, uint64_t q) {
    return v < 0 ? q - (uint64_t)(-v) : (uint64_t)v;
}

/* u64 draw #idx of a (w13,w14,dom) stream: 8 draws per block, counter in w12 */
typedef struct { const uint32_t *key; uint32_t w13, w14, w15; uint32_t blk; uint32_t buf[16]; int have; } stream_t;
static uint64_t stream_u64(stream_t *s, uint64_t idx) {
    uint32_t b = (uint32_t)(idx >> 3);
    if (!s->have || s->blk != b) {
        lsro_chacha_block(s->key, b, s->w13, s->w14, s->w15, s->buf);
        s->blk = b; s->have = 1;
    }
    uint32_t w = (uint32_t)(idx & 7);
    return (uint64_t)s->buf[2 * w] | ((uint64_t)s->buf[2 * w + 1] << 32);
}

static void load_key(const uint8_t seed32[32], uint32_t key[8]) {
    for (int i = 0; i < 8; i++) {
        key[i] = (uint32_t)seed32[4 * i] | ((uint32_t)seed32[4 * i + 1] << 8) |
                 ((uint32_t)seed32[4 * i + 2] << 16) | ((uint32_t)seed32[4 * i + 3] << 24);
    }
}

/* utils.cpp:132-146 with std::random_device replaced by a keyed stream */
int lsro_sample_gaussian_seeded(uint64_t *out, size_t len, double sigma, const uint8_t seed32[32]) {
    if (!out || len == 0 || !(sigma > 0.0) || !isfinite(sigma) || !seed32) return -1;
    uint64_t *cdf = (uint64_t *)malloc(sizeof(uint64_t) * 32768);
    if (!cdf) return -1;
    size_t cn = lsro_cdt_build(sigma, cdf, 32768);
    if (!cn) { free(cdf); return -1; }
    uint32_t key[8];
    load_key(seed32, key);
    uint32_t blk[16];
    for (size_t i = 0; i < len; i++) {
        if ((i & 3) == 0) {
            uint64_t B = i >> 2;
            lsro_chacha_block(key, (uint32_t)B, (uint32_t)(B >> 32), 0, DOM_SAMPLE, blk);
        }
        uint32_t w = (uint32_t)(i & 3) * 4;
        uint64_t u1 = (uint64_t)blk[w] | ((uint64_t)blk[w + 1] << 32);
        uint64_t u2 = (uint64_t)blk[w + 2] | ((uint64_t)blk[w + 3] << 32);
        out[i] = (uint64_t)lsro_cdt_sample(cdf, cn, u1, u2);
    }
    free(cdf);
    return 0;
}

lsro_lwe *lsro_lwe_create(uint64_t modulus_req, uint32_t n, uint32_t k, double sigma,
                          const uint8_t seed32[32]) {
    if (!seed32) return NULL;
    if (n < 16 || (n & (n - 1)) != 0 || n > 131072) return NULL;
    if (k < 1 || k > 16) return NULL;
    if (!(sigma > 0.0) || !isfinite(sigma)) return NULL;
    lsro_lwe *c = (lsro_lwe *)calloc(1, sizeof(*c));
    if (!c) return NULL;
    c->n = n; c->k = k; c->sigma = sigma;
    c->q = ntt_friendly(modulus_req, n) ? modulus_req : (n <= 4096 ? LSRO_Q0 : LSRO_Q1);
    c->p = plain_modulus(c->q);
    c->delta = (c->q - 1) / c->p;
    load_key(seed32, c->key);
    c->ntt = lsro_ntt_create(c->q, n);
    c->cdf = (uint64_t *)malloc(sizeof(uint64_t) * 32768);
    c->cdf_n = c->cdf ? lsro_cdt_build(sigma, c->cdf, 32768) : 0;
    c->A = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)k * k * n);
    c->zh = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)(k > 1 ? k - 1 : 1) * n);
    if (!c->ntt || !c->cdf_n || !c->A || !c->zh) { lsro_lwe_free(c); return NULL; }

    const uint64_t q = c->q;
    int qbits = 64 - __builtin_clzll(q);
    const uint64_t mask = qbits == 64 ? ~0ull : ((1ull << qbits) - 1);
    /* top k-1 rows: uniform, sampled directly in the NTT domain by rejection */
    for (uint32_t i = 0; i + 1 < k; i++) {
        for (uint32_t j = 0; j < k; j++) {
            stream_t s = {c->key, 0, i * k + j, DOM_MATRIX, 0, {0}, 0};
            uint64_t *dst = c->A + ((size_t)i * k + j) * n;
            uint64_t idx = 0;
            for (uint32_t got = 0; got < n;) {
                uint64_t v = stream_u64(&s, idx++) & mask;
                if (v < q) dst[got++] = v;
            }
        }
    }
    /* trapdoor polys z'_0..z'_{k-2}, f_0..f_{k-1}: CDT samples, draws (2i, 2i+1) */
    uint64_t *f = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)k * n);
    if (!f) { lsro_lwe_free(c); return NULL; }
    for (uint32_t P = 0; P < 2 * k - 1; P++) {
        stream_t s = {c->key, 0, P, DOM_TRAP, 0, {0}, 0};
        uint64_t *dst = P < k - 1 ? c->zh + (size_t)P * n : f + (size_t)(P - (k - 1)) * n;
        for (uint32_t i = 0; i < n; i++) {
            uint64_t u1 = stream_u64(&s, 2ull * i), u2 = stream_u64(&s, 2ull * i + 1);
            dst[i] = to_residue(lsro_cdt_sample(c->cdf, c->cdf_n, u1, u2), q);
        }
        fwd_one(c->ntt, dst);
    }
    /* last row: A[k-1][j] = f_j - sum_i z'_i * A[i][j] */
    for (uint32_t j = 0; j < k; j++) {
        uint64_t *dst = c->A + ((size_t)(k - 1) * k + j) * n;
        for (uint32_t x = 0; x < n; x++) {
            uint64_t acc = f[(size_t)j * n + x];
            for (uint32_t i = 0; i + 1 < k; i++) {
                uint64_t prod = lsro_mulmod(c->zh[(size_t)i * n + x],
                                            c->A[((size_t)i * k + j) * n + x], q);
                acc = acc >= prod ? acc - prod : acc + q - prod;
            }
            dst[x] = acc;
        }
    }
    free(f);
    return c;
}

void lsro_lwe_free(lsro_lwe *c) {
    if (!c) return;
    lsro_ntt_free(c->ntt);
    free(c->cdf); free(c->A); free(c->zh);
    free(c);
}

uint64_t lsro_lwe_modulus(const lsro_lwe *c) { return c->q; }
uint64_t lsro_lwe_plain_modulus(const lsro_lwe *c) { return c->p; }
uint64_t lsro_lwe_delta(const lsro_lwe *c) { return c->delta; }
uint32_t lsro_lwe_words(const lsro_lwe *c) { return 1 + c->k * c->n; }
const uint64_t *lsro_lwe_matrix(const lsro_lwe *c) { return c->A; }
const uint64_t *lsro_lwe_trapdoor(const lsro_lwe *c) { return c->zh; }

/*
 * Commitment randomness layout.  Coefficient c of
 * polynomial P (P<k: s_P, P>=k: e_{P-k}) lives in chunk tau = c mod (n/16), lane j = c div (n/16)
 * (a chunk is the 16 coefficients tau + (n/16) j that one thread of the radix-16 first / last NTT pass owns).
 * One ChaCha block per (P, tau) carries all 16 samples of the chunk: with w = word j of
 *   block(P)            = ChaCha(key, w12=seed_lo, w13=seed_hi, w14=tau, w15=DOM_COMMIT|P)
 * the two draws of utils.cpp:95-121 (sample_single) are
 *   u2 (sign draw)      = w            (only bit 0 is looked at, utils.cpp:113)
 *   u1 (magnitude draw) = (w >> 1) * 2^33 + lo33,
 *   lo33                = 33 bits of the REFINEMENT block 0x100 | (2P + (j>>3)): bit 0 of word 2(j&7) as
 *                         bit 32, word 2(j&7)+1 as bits 31..0.
 * u1 is a uniform 64-bit word independent of u2.  Its top 31 bits decide the sample unless they equal the
 * top 31 bits of a table entry (probability < 2^-26 per sample), so an implementation may fetch the refinement
 * block only then; this restatement always forms the full u1.
 */
static void sample_chunk(const lsro_lwe *c, uint64_t seed, uint32_t tau, int64_t *out /*[2k][16]*/) {
    const uint32_t k = c->k;
    for (uint32_t P = 0; P < 2 * k; P++) {
        uint32_t blk[16], fine[2][16];
        lsro_chacha_block(c->key, (uint32_t)seed, (uint32_t)(seed >> 32), tau, DOM_COMMIT | P, blk);
        for (uint32_t h = 0; h < 2; h++)
            lsro_chacha_block(c->key, (uint32_t)seed, (uint32_t)(seed >> 32), tau,
                              DOM_COMMIT | 0x100u | (2 * P + h), fine[h]);
        for (uint32_t j = 0; j < 16; j++) {
            const uint32_t w = blk[j];
            const uint32_t *f = fine[j >> 3];
            const uint64_t lo33 = ((uint64_t)(f[2 * (j & 7)] & 1u) << 32) | (uint64_t)f[2 * (j & 7) + 1];
            const uint64_t u1 = ((uint64_t)(w >> 1) << 33) | lo33;
            const uint64_t u2 = (uint64_t)w;
            out[P * 16 + j] = lsro_cdt_sample(c->cdf, c->cdf_n, u1, u2);
        }
    }
}

void lsro_lwe_sample_se(const lsro_lwe *c, uint64_t seed, int64_t *s, int64_t *e) {
    const uint32_t n = c->n, k = c->k;
    int64_t buf[32 * 16];
    for (uint32_t tau = 0; tau < n / 16; tau++) {
        sample_chunk(c, seed, tau, buf);
        for (uint32_t P = 0; P < k; P++) {
            for (uint32_t j = 0; j < 16; j++) {
                s[(size_t)P * n + tau + (size_t)(n / 16) * j] = buf[P * 16 + j];
                e[(size_t)P * n + tau + (size_t)(n / 16) * j] = buf[(k + P) * 16 + j];
            }
        }
    }
}

/* any int64 -> [0,q) (explicit mode takes caller-supplied values, not only CDT samples) */
static inline uint64_t to_residue_any(int64_t v, uint64_t q) {
    const uint64_t mag = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
    const uint64_t r = mag % q;
    return (v < 0 && r) ? q - r : r;
}

/* t = A*s + e + Delta*m for given s, e (the embedding of the message follows
 * commitment.cpp:146-149: first min(len, n) slots, the rest zero) */
static int commit_core(const lsro_lwe *c, const uint64_t *msg, size_t msg_len, uint64_t *out,
                       const int64_t *s, const int64_t *e, uint64_t *sh, uint64_t *acc) {
    const uint32_t n = c->n, k = c->k;
    const uint64_t q = c->q;
    for (uint32_t j = 0; j < k; j++) {
        for (uint32_t x = 0; x < n; x++) sh[(size_t)j * n + x] = to_residue_any(s[(size_t)j * n + x], q);
        fwd_one(c->ntt, sh + (size_t)j * n);
    }
    out[0] = (uint64_t)k * n * 8;
    const size_t L = msg_len < n ? msg_len : n;     /* commitment.cpp:146-149 */
    for (uint32_t i = 0; i < k; i++) {
        for (uint32_t x = 0; x < n; x++) {
            uint64_t a = 0;
            for (uint32_t j = 0; j < k; j++) {
                a += lsro_mulmod(c->A[((size_t)i * k + j) * n + x], sh[(size_t)j * n + x], q);
                if (a >= q) a -= q;
            }
            acc[x] = a;
        }
        inv_one(c->ntt, acc);
        uint64_t *t = out + 1 + (size_t)i * n;
        for (uint32_t x = 0; x < n; x++) {
            uint64_t v = acc[x] + to_residue_any(e[(size_t)i * n + x], q);
            if (v >= q) v -= q;
            if (i == k - 1 && x < L) {
                v += lsro_mulmod(c->delta, msg[x] % c->p, q);
                if (v >= q) v -= q;
            }
            t[x] = v;
        }
    }
    return 0;
}

static int commit_one(const lsro_lwe *c, const uint64_t *msg, size_t msg_len,
                      uint64_t seed, uint64_t *out, int64_t *s, int64_t *e, uint64_t *sh,
                      uint64_t *acc) {
    lsro_lwe_sample_se(c, seed, s, e);
    return commit_core(c, msg, msg_len, out, s, e, sh, acc);
}

int lsro_lwe_commit_explicit(const lsro_lwe *c, const uint64_t *msg, size_t msg_len,
                             const int64_t *s, const int64_t *e, uint64_t *out_words) {
    if (!c || (!msg && msg_len) || !s || !e || !out_words) return -1;
    const size_t kn = (size_t)c->k * c->n;
    uint64_t *sh = (uint64_t *)malloc(sizeof(uint64_t) * (kn + c->n));
    if (!sh) return -1;
    int rc = commit_core(c, msg, msg_len, out_words, s, e, sh, sh + kn);
    free(sh);
    return rc;
}

int lsro_lwe_commit(const lsro_lwe *c, const uint64_t *msg, size_t msg_len,
                    uint64_t seed, uint64_t *out_words) {
    if (!c || !msg || !out_words) return -1;       /* commitment.cpp:144 */
    const size_t kn = (size_t)c->k * c->n;
    int64_t *s = (int64_t *)malloc(sizeof(int64_t) * kn * 2);
    uint64_t *sh = (uint64_t *)malloc(sizeof(uint64_t) * (kn + c->n));
    if (!s || !sh) { free(s); free(sh); return -1; }
    int rc = commit_one(c, msg, msg_len, seed, out_words, s, s + kn, sh, sh + kn);
    free(s); free(sh);
    return rc;
}

int lsro_lwe_commit_batch(const lsro_lwe *c, const uint64_t *msgs, size_t msg_len,
                          const uint64_t *seeds, size_t count, uint64_t *out_words,
                          int threads) {
    if (!c || !msgs || !seeds || !out_words) return -1;
    if (threads < 1) threads = 1;
    const size_t kn = (size_t)c->k * c->n;
    const size_t words = 1 + kn;
    int bad = 0;
#pragma omp parallel num_threads(threads)
    {
        int64_t *s = (int64_t *)malloc(sizeof(int64_t) * kn * 2);
        uint64_t *sh = (uint64_t *)malloc(sizeof(uint64_t) * (kn + c->n));
        if (!s || !sh) {
#pragma omp atomic write
            bad = 1;
        } else {
#pragma omp for schedule(static)
            for (long b = 0; b < (long)count; b++) {
                commit_one(c, msgs + (size_t)b * msg_len, msg_len, seeds[b],
                           out_words + (size_t)b * words, s, s + kn, sh, sh + kn);
            }
        }
        free(s); free(sh);
    }
    return bad ? -1 : 0;
}

static int container_ok(const lsro_lwe *c, const uint64_t *w, size_t len) {
    /* commitment.cpp:66-75: len >= 1, 0 < byte_len <= available; plus our format */
    if (!w || len < 1) return 0;
    const uint64_t byte_len = w[0];
    if (byte_len == 0 || byte_len > (len - 1) * 8) return 0;
    if (byte_len != (uint64_t)c->k * c->n * 8) return 0;
    for (size_t i = 0; i < (size_t)c->k * c->n; i++) if (w[1 + i] >= c->q) return 0;
    return 1;
}

int lsro_lwe_verify(const lsro_lwe *c, const uint64_t *w, size_t len,
                    const uint64_t *msg, size_t msg_len) {
    if (!c || !w || !msg) return -1;               /* commitment.cpp:207 */
    if (!container_ok(c, w, len)) return -1;       /* :210-212 */
    const uint32_t n = c->n, k = c->k;
    const uint64_t q = c->q;
    if (n < msg_len) return 0;                     /* :219-221 */
    uint64_t *u = (uint64_t *)calloc(n, sizeof(uint64_t));
    uint64_t *tmp = (uint64_t *)malloc(sizeof(uint64_t) * n);
    if (!u || !tmp) { free(u); free(tmp); return -1; }
    for (uint32_t i = 0; i + 1 < k; i++) {
        memcpy(tmp, w + 1 + (size_t)i * n, sizeof(uint64_t) * n);
        fwd_one(c->ntt, tmp);
        for (uint32_t x = 0; x < n; x++) {
            u[x] += lsro_mulmod(c->zh[(size_t)i * n + x], tmp[x], q);
            if (u[x] >= q) u[x] -= q;
        }
    }
    inv_one(c->ntt, u);
    uint64_t diff = 0;
    const uint64_t *last = w + 1 + (size_t)(k - 1) * n;
    for (uint32_t x = 0; x < (uint32_t)msg_len; x++) {
        uint64_t v = u[x] + last[x];
        if (v >= q) v -= q;
        uint64_t d = ((v + c->delta / 2) / c->delta) % c->p;
        diff |= d ^ (msg[x] % c->p);               /* :223-226; words are bound mod p */
    }
    free(u); free(tmp);
    return diff == 0 ? 1 : 0;
}

int lsro_lwe_linear_combine(const lsro_lwe *c, const uint64_t *const *comms,
                            const size_t *comm_lens, const uint64_t *coeffs,
                            size_t count, uint64_t *out) {
    if (!c || !comms || !comm_lens || !coeffs || count == 0 || !out) return -1;  /* :240 */
    const size_t kn = (size_t)c->k * c->n;
    int has = 0;
    memset(out, 0, sizeof(uint64_t) * (1 + kn));
    for (size_t i = 0; i < count; i++) {
        if (!comms[i]) continue;                   /* :248-250 */
        if (!container_ok(c, comms[i], comm_lens[i])) return -1;   /* :253-255 */
        uint64_t cf = coeffs[i] % c->p;            /* :90 coeff %= plain_modulus */
        if (cf > c->p / 2) cf = c->q - (c->p - cf); /* centred representative */
        for (size_t x = 0; x < kn; x++) {
            uint64_t v = out[1 + x] + lsro_mulmod(cf, comms[i][1 + x], c->q);
            if (v >= c->q) v -= c->q;
            out[1 + x] = v;
        }
        has = 1;
    }
    if (!has) return -1;                           /* :268-270 */
    out[0] = (uint64_t)kn * 8;
    return 0;
}
