/*
 * test_cabi.c -- drop-in check in the reference's own language family: a plain C
 * program that includes the reference's header paths (<lambda_snark/...>) and links
 * liblambda_snark_core exactly as the cpp-core tests and lambda-snark-sys do.
 * Ports the assertions of cpp-core/tests/test_ntt.cpp, test_commitment.cpp,
 * test_utils.cpp and rust-api/lambda-snark-sys/src/lib.rs:28-43.
 * Exit code 0 = all assertions hold.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lambda_snark/commitment.h"
#include "lambda_snark/ntt.h"
#include "lambda_snark/r1cs.h"
#include "lambda_snark/types.h"
#include "lambda_snark/utils.h"

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); failures++; } } while (0)

static void test_ntt(void) {                       /* test_ntt.cpp:13-90 */
    const uint64_t q = 12289; const uint32_t n = 256;
    NttContext* ctx = ntt_context_create(q, n);
    CHECK(ctx != NULL);
    uint64_t a[256], b[256], r[256], orig[256];
    for (uint32_t i = 0; i < n; i++) a[i] = 1;
    CHECK(ntt_forward(ctx, a, n) == 0);
    for (uint32_t i = 0; i < n; i++) a[i] = 1;
    CHECK(ntt_inverse(ctx, a, n) == 0);
    memset(orig, 0, sizeof(orig));
    for (uint32_t i = 0; i < 8; i++) orig[i] = i + 1;
    memcpy(a, orig, sizeof(a));
    CHECK(ntt_forward(ctx, a, n) == 0);
    CHECK(a[0] == 26 && a[1] == 11046 && a[2] == 1743 && a[3] == 1098 && a[255] == 11454);   /* SURVEY 8c KAT */
    CHECK(ntt_inverse(ctx, a, n) == 0);
    CHECK(memcmp(a, orig, sizeof(a)) == 0);
    for (uint32_t i = 0; i < n; i++) { a[i] = 2; b[i] = 3; }
    ntt_mul_pointwise(ctx, r, a, b, n);
    for (uint32_t i = 0; i < n; i++) CHECK(r[i] == 6);
    CHECK(ntt_forward(NULL, a, n) == -1);
    CHECK(ntt_forward(ctx, NULL, n) == -1);
    ntt_context_free(NULL);
    ntt_context_free(ctx);
    CHECK(ntt_context_create(12289, 0) == NULL);
    CHECK(ntt_context_create(12289, 100) == NULL);
}

static void test_commitment(void) {                /* test_commitment.cpp:12-166 */
    PublicParams params;
    params.profile = PROFILE_RING_B; params.security_level = 128; params.modulus = 12289;
    params.ring_degree = 4096; params.module_rank = 2; params.sigma = 3.19;
    CHECK(lwe_context_create(NULL) == NULL);       /* lambda-snark-sys lib.rs:28-34 */
    LweContext* ctx = lwe_context_create(&params);
    CHECK(ctx != NULL);
    if (!ctx) return;

    uint64_t message[] = {1, 2, 3, 4};
    LweCommitment* comm = lwe_commit(ctx, message, 4, 0x1234);
    CHECK(comm != NULL && comm->len > 0 && comm->data != NULL);
    CHECK(comm->data[0] == (comm->len - 1) * 8);   /* types.h:32-34 */
    LweCommitment* clone = lwe_commitment_clone(comm);
    CHECK(clone && clone->len == comm->len && memcmp(clone->data, comm->data, comm->len * 8) == 0);
    lwe_commitment_free(clone);
    lwe_commitment_free(comm);

    uint64_t msg1[] = {1, 2, 3}, msg2[] = {4, 5, 6};
    LweCommitment* c1 = lwe_commit(ctx, msg1, 3, 0x1234);
    LweCommitment* c2 = lwe_commit(ctx, msg2, 3, 0x1234);
    CHECK(c1 && c2 && memcmp(c1->data, c2->data, c1->len * 8) != 0);
    lwe_commitment_free(c1); lwe_commitment_free(c2);

    CHECK(lwe_commit(NULL, NULL, 0, 0) == NULL);
    CHECK(lwe_commit(ctx, NULL, 10, 0) == NULL);
    lwe_commitment_free(NULL);

    uint64_t m[] = {7, 11, 13, 17};
    comm = lwe_commit(ctx, m, 4, 0);
    uint64_t randomness = 0;
    LweOpening opening = {&randomness, 1};
    CHECK(lwe_verify_opening(ctx, comm, m, 4, &opening) == 1);
    uint64_t wrong[] = {7, 11 ^ 1, 13, 17};
    CHECK(lwe_verify_opening(ctx, comm, wrong, 4, &opening) == 0);
    lwe_commitment_free(comm);

    uint64_t a[] = {1, 2, 3, 4}, b[] = {5, 6, 7, 8};
    c1 = lwe_commit(ctx, a, 4, 0); c2 = lwe_commit(ctx, b, 4, 0);
    const LweCommitment* inputs[] = {c1, c2};
    uint64_t coeffs[] = {2, 3};
    LweCommitment* combined = lwe_linear_combine(ctx, inputs, coeffs, 2);
    CHECK(combined != NULL);
    uint64_t expected[4];
    for (int i = 0; i < 4; i++) expected[i] = coeffs[0] * a[i] + coeffs[1] * b[i];
    CHECK(lwe_verify_opening(ctx, combined, expected, 4, &opening) == 1);
    expected[0] += 1;
    CHECK(lwe_verify_opening(ctx, combined, expected, 4, &opening) == 0);
    lwe_commitment_free(c1); lwe_commitment_free(c2); lwe_commitment_free(combined);
    lwe_context_free(ctx);
    lwe_context_free(NULL);
}

static void test_sampler(void) {                   /* test_utils.cpp:26-70 */
    enum { N = 4096 };
    static uint64_t buf[N];
    CHECK(sample_gaussian(NULL, 16, 3.2) == -1);
    CHECK(sample_gaussian(buf, 0, 3.2) == -1);
    CHECK(sample_gaussian(buf, 16, 0.0) == -1);
    CHECK(sample_gaussian(buf, 16, INFINITY) == -1);
    CHECK(sample_gaussian(buf, N, 3.2) == 0);
    double mean = 0.0, m2 = 0.0; size_t pos = 0, neg = 0;
    for (size_t i = 0; i < N; i++) {
        const int64_t v = (int64_t)buf[i];
        const double x = (double)v, d = x - mean;
        mean += d / (double)(i + 1);
        m2 += d * (x - mean);
        if (v > 0) pos++; else if (v < 0) neg++;
    }
    const double sd = sqrt(m2 / (double)(N - 1));
    CHECK(fabs(mean) < 0.5);
    CHECK(fabs(sd - 3.2) < 0.8);
    CHECK(pos > N / 4 && neg > N / 4);
    CHECK(llabs((long long)pos - (long long)neg) < N / 5);
}

static void test_r1cs(void) {                      /* test_vectors.rs:70-93 through the C ABI */
    SparseEntry ea[] = {{0, 1, 1}}, eb[] = {{0, 2, 1}}, ec[] = {{0, 3, 1}};
    SparseMatrix A = {ea, 1, 1, 4}, B = {eb, 1, 1, 4}, Cm = {ec, 1, 1, 4};
    void* h = NULL;
    CHECK(lambda_snark_r1cs_create(&A, &B, &Cm, 17592186044417ULL, &h) == LAMBDA_SNARK_OK);
    CHECK(lambda_snark_r1cs_num_constraints(h) == 1 && lambda_snark_r1cs_num_variables(h) == 4);
    uint64_t z[] = {1, 7, 13, 91};
    R1CSWitness w = {z, 4};
    bool ok = false;
    CHECK(lambda_snark_r1cs_validate_witness(h, &w, &ok) == LAMBDA_SNARK_OK && ok);
    z[1] = 8;
    CHECK(lambda_snark_r1cs_validate_witness(h, &w, &ok) == LAMBDA_SNARK_OK && !ok);
    lambda_snark_r1cs_free(h);
}

int main(void) {
    test_ntt();
    test_commitment();
    test_sampler();
    test_r1cs();
    if (failures) { fprintf(stderr, "%d assertion(s) failed\n", failures); return 1; }
    printf("test_cabi: all reference assertions hold\n");
    return 0;
}
