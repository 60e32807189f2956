// test_host_mirror.cpp -- the reference's Rust unit tests that sit directly on the FFI, ported assertion for assertion to the
// C++ mirror of its safe wrappers (include/lambda_snark_b200.hpp):
//   context.rs:83-101      test_context_create_and_drop
//   commitment.rs:140-219  test_commitment_create, test_linear_combination_roundtrip
//   opening.rs:289-421     test_opening_generation / _evaluation_tv1 / _correctness / _soundness_* / _empty_witness / _out_of_field
//   challenge.rs:160-322   deterministic, collision, public-input sensitivity, in field, domain separation, 100 distinct
// Exit code 0 = every assertion holds.  Needs a CUDA device (the library has no CPU path).
#include <algorithm>
#include <cstdio>
#include <set>

#include "lambda_snark_b200.hpp"

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } } while (0)

static const uint64_t TEST_MODULUS = 17592186044417ull;      // 2^44 + 1, the modulus of every reference test

static lsr::Params test_params() { lsr::Params p; p.n = 4096; p.k = 2; p.q = TEST_MODULUS; p.sigma = 3.19; return p; }

static void context_tests() {
    lsr::LweContext ctx(test_params());
    CHECK(ctx.modulus() == TEST_MODULUS);
    bool threw = false;
    try { lsr::Params bad = test_params(); bad.n = 1000; lsr::LweContext c(bad); } catch (const lsr::InvalidInput&) { threw = true; }
    CHECK(threw);                                            // Params::validate: n must be a power of two
}

static void commitment_tests() {
    lsr::LweContext ctx(test_params());
    lsr::Commitment c = lsr::Commitment::create(ctx, {1, 2, 3}, 0x1234);
    CHECK(c.len() > 0 && c.data()[0] == (c.len() - 1) * 8);
    lsr::Commitment copy(c);                                 // Clone: deep copy
    CHECK(copy.len() == c.len() && copy.data() != c.data() && std::equal(c.data(), c.data() + c.len(), copy.data()));

    // test_linear_combination_roundtrip: 2 * m1 + 3 * m2 opens to the combined message
    std::vector<uint64_t> m1{1, 2, 3, 4}, m2{2, 4, 6, 8};
    lsr::Commitment c1 = lsr::Commitment::create(ctx, m1, 0), c2 = lsr::Commitment::create(ctx, m2, 1);
    std::vector<uint64_t> coeffs{2, 3};
    lsr::Commitment comb = lsr::Commitment::linear_combine(ctx, {&c1, &c2}, coeffs);
    std::vector<uint64_t> expected(4);
    for (size_t i = 0; i < 4; i++)
        expected[i] = lsr::add_mod(lsr::mul_mod(2, m1[i], TEST_MODULUS), lsr::mul_mod(3, m2[i], TEST_MODULUS), TEST_MODULUS);
    LweOpening none{nullptr, 0};
    CHECK(lwe_verify_opening(ctx.as_ptr(), comb.as_ffi_ptr(), expected.data(), expected.size(), &none) == 1);
    bool threw = false;
    try { lsr::Commitment::linear_combine(ctx, {}, {}); } catch (const lsr::InvalidInput&) { threw = true; }
    CHECK(threw);
    threw = false;
    try { lsr::Commitment::linear_combine(ctx, {&c1}, coeffs); } catch (const lsr::InvalidInput&) { threw = true; }
    CHECK(threw);
}

static void opening_tests() {
    lsr::LweContext ctx(test_params());
    const std::vector<uint64_t> poly{1, 7, 13, 91};          // TV-1 witness as coefficients
    const uint64_t randomness = 0x1234, alpha = 12345;
    lsr::Opening o = lsr::generate_opening(poly, alpha, randomness, TEST_MODULUS);
    const uint64_t want = (1 + 7 * alpha % TEST_MODULUS + lsr::mul_mod(13, lsr::mul_mod(alpha, alpha, TEST_MODULUS), TEST_MODULUS) +
                           lsr::mul_mod(91, lsr::mul_mod(alpha, lsr::mul_mod(alpha, alpha, TEST_MODULUS), TEST_MODULUS), TEST_MODULUS)) % TEST_MODULUS;
    CHECK(o.evaluation == want);
    CHECK(!o.witness.empty() && o.witness[0] == randomness && o.witness.size() == 5);

    lsr::Commitment c = lsr::Commitment::create(ctx, poly, randomness);
    CHECK(lsr::verify_opening_with_context(c, alpha, o, TEST_MODULUS, ctx));                       // correctness
    lsr::Opening forged{o.evaluation + 1, o.witness};
    CHECK(!lsr::verify_opening(c, alpha, forged, TEST_MODULUS));                                    // wrong evaluation
    lsr::Opening other = lsr::generate_opening({1, 7, 13, 92}, alpha, randomness, TEST_MODULUS);
    CHECK(lsr::verify_opening(c, alpha, other, TEST_MODULUS));                                      // consistent by itself ...
    CHECK(!lsr::verify_opening_with_context(c, alpha, other, TEST_MODULUS, ctx));                   // ... but not bound to c
    lsr::Commitment c12 = lsr::Commitment::create(ctx, {1, 2}, 0x1234);
    CHECK(!lsr::verify_opening(c12, 100, lsr::Opening{42, {}}, TEST_MODULUS));                      // empty witness
    CHECK(!lsr::verify_opening(c12, 100, lsr::Opening{TEST_MODULUS, {0x1234, 1, 2}}, TEST_MODULUS)); // out of field
}

static void challenge_tests() {
    lsr::LweContext ctx(test_params());
    lsr::Commitment c1 = lsr::Commitment::create(ctx, {1, 2, 3}, 0x1234), c2 = lsr::Commitment::create(ctx, {4, 5, 6}, 0x1234);
    const std::vector<uint64_t> pub{1, 91};
    lsr::Challenge a = lsr::Challenge::derive(pub, c1, TEST_MODULUS), b = lsr::Challenge::derive(pub, c1, TEST_MODULUS);
    CHECK(a.alpha == b.alpha && a.hash == b.hash);                                                  // deterministic
    CHECK(lsr::Challenge::derive(pub, c2, TEST_MODULUS).alpha != a.alpha);                          // commitment sensitivity
    CHECK(lsr::Challenge::derive({1, 92}, c1, TEST_MODULUS).alpha != a.alpha);                      // public-input sensitivity
    CHECK(lsr::Challenge::derive({1}, c1, TEST_MODULUS).alpha != a.alpha);
    CHECK(a.alpha < TEST_MODULUS);                                                                  // in field
    // length prefixes separate ([1, 91], ...) from ([1], [91, ...]): a different public-input count changes the hash
    CHECK(lsr::Challenge::derive({}, c1, TEST_MODULUS).hash != a.hash);
    std::set<uint64_t> seen;
    for (uint64_t i = 0; i < 100; i++) seen.insert(lsr::Challenge::derive({i}, c1, TEST_MODULUS).alpha);
    CHECK(seen.size() == 100);                                                                      // all distinct
}

static void ntt_tests() {                                                                           // lambda-snark-sys lib.rs:36-43
    lsr::NttContext ntt(12289, 256);
    std::vector<uint64_t> v(256, 0);
    for (int i = 0; i < 8; i++) v[(size_t)i] = (uint64_t)i + 1;
    const std::vector<uint64_t> orig = v;
    ntt.forward(v);
    CHECK(v[0] == 26 && v[1] == 11046 && v[255] == 11454);
    ntt.inverse(v);
    CHECK(v == orig);
    CHECK(ntt.mul_pointwise(std::vector<uint64_t>(256, 2), std::vector<uint64_t>(256, 3)) == std::vector<uint64_t>(256, 6));
    bool threw = false;
    try { lsr::NttContext bad(12289, 100); } catch (const lsr::FfiError&) { threw = true; }
    CHECK(threw);
}

int main() {
    context_tests();
    commitment_tests();
    opening_tests();
    challenge_tests();
    ntt_tests();
    if (failures) { std::fprintf(stderr, "%d failures\n", failures); return 1; }
    std::printf("host mirror: all ported Rust assertions hold\n");
    return 0;
}
