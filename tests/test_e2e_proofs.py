"""End-to-end proofs (BASELINE configs[0], [1]): the reference's Rust prover / verifier,
restated in tests/prover_restatement.py, run on top of the commitment produced by
  * the CPU oracle           (CPU test: pins the restatement itself), and
  * the B200 library's C ABI (GPU test: lwe_commit -> words the transcript hashes).
Pass criterion: the reference verifier's equations (lib.rs:1047-1078) hold with
alpha, beta recomputed from the commitment words; a tampered proof or commitment is rejected.
"""
import numpy as np
import pytest

import prover_restatement as P
from conftest import Q0
from oracle import oracle as O

TEST_MODULUS = 17592186044417      # 2^44 + 1 (composite; every reference test vector uses it)
CLI_MODULUS = 17592186044423       # prime used by lambda-snark-cli
SEED32 = bytes(range(32))

CIRCUITS = [("tv1", lambda: P.tv1(TEST_MODULUS)), ("tv2", lambda: P.tv2(TEST_MODULUS)),
            ("healthcare", lambda: P.healthcare(CLI_MODULUS)),
            ("mul_gates_m4", lambda: P.multiplication_gates(4, TEST_MODULUS)),
            ("mul_gates_m20", lambda: P.multiplication_gates(20, CLI_MODULUS))]


def check_circuit(name, build, commit):
    r1cs, witness = build()
    assert r1cs.is_satisfied(witness)
    proof, q_coeffs = P.prove_r1cs(r1cs, witness, commit, 0x1234)
    publics = r1cs.public_inputs(witness)
    assert P.verify_r1cs(proof, publics, r1cs), name
    # the quotient really is (A_z B_z - C_z) / Z_H at a random point
    for x in (12345, proof.alpha):
        a, b, c = r1cs.evals(witness)
        ap, bp, cp = (P.lagrange_interpolate(e, r1cs.modulus) for e in (a, b, c))
        lhs = P.mul_mod(r1cs.eval_poly(q_coeffs, x), r1cs.eval_vanishing(x), r1cs.modulus)
        rhs = P.sub_mod(P.mul_mod(r1cs.eval_poly(ap, x), r1cs.eval_poly(bp, x), r1cs.modulus),
                        r1cs.eval_poly(cp, x), r1cs.modulus)
        assert lhs == rhs
    # soundness smoke: tampering with an evaluation, the commitment, or the public input is rejected
    bad = P.ProofR1CS(**{**proof.__dict__, "q_alpha": (proof.q_alpha + 1) % r1cs.modulus})
    assert not P.verify_r1cs(bad, publics, r1cs)
    words = np.array(proof.commitment_words, copy=True)
    words[5] ^= np.uint64(1)
    assert not P.verify_r1cs(P.ProofR1CS(**{**proof.__dict__, "commitment_words": words}), publics, r1cs)
    if len(publics) > 1:
        assert not P.verify_r1cs(proof, publics[:-1] + [publics[-1] + 1], r1cs)
    # an unsatisfying witness has no quotient
    wrong = list(witness)
    wrong[-1] = (wrong[-1] + 1) % r1cs.modulus
    if not r1cs.is_satisfied(wrong):
        with pytest.raises(ValueError):
            r1cs.compute_quotient_poly(wrong)
    return proof


def test_restatement_helpers():
    q = TEST_MODULUS
    assert P.mod_inverse(3, q) * 3 % q == 1
    assert P.mod_inverse(2, 17) == 9
    assert P.lagrange_interpolate([7], q) == [7]
    poly = P.lagrange_interpolate([5, 11, 19], q)          # through (0,5), (1,11), (2,19)
    r = P.R1CS(3, 1, 1, [{}] * 3, [{}] * 3, [{}] * 3, q)
    assert [r.eval_poly(poly, x) for x in (0, 1, 2)] == [5, 11, 19]
    assert P.vanishing_poly(3, q) == [0, 2, q - 3, 1]        # X(X-1)(X-2)
    assert r.eval_vanishing(5) == 5 * 4 * 3
    # known transcript: empty inputs / one word
    a1, d1 = P.challenge_derive([1, 91], np.array([8, 1, 2], dtype=np.uint64), q)
    a2, d2 = P.challenge_derive([1, 91], np.array([8, 1, 3], dtype=np.uint64), q)
    assert a1 != a2 and len(d1) == 32 and a1 < q


@pytest.mark.parametrize("name,build", CIRCUITS)
def test_proofs_verify_on_oracle_commitment(name, build):
    ctx = O.OracleLwe(TEST_MODULUS, 4096, 2, 3.19, SEED32)      # falls back to q = 17592169062401

    def commit(fields, seed):
        r1cs, _ = build()
        return ctx.commit([f % r1cs.modulus for f in fields], seed)    # commitment.rs:31-45

    check_circuit(name, build, commit)


@pytest.mark.gpu
@pytest.mark.parametrize("name,build", CIRCUITS)
def test_proofs_verify_on_b200_commitment(gpu, name, build):
    from lambda_snark_r_b200 import api
    r1cs, _ = build()
    ctx = api.LweContext(api.Params(n=4096, k=2, q=r1cs.modulus, sigma=3.19), seed32=SEED32)
    orc = O.OracleLwe(r1cs.modulus, 4096, 2, 3.19, SEED32)
    keep = []

    def commit(fields, seed):
        c = api.Commitment.new(ctx, fields, seed)
        keep.append(c)
        return c.as_bytes().copy()

    proof = check_circuit(name, build, commit)
    # same words as the oracle's commitment -> same alpha, beta, same proof as the CPU run
    q_coeffs = r1cs.compute_quotient_poly(build()[1])
    want = orc.commit([f % r1cs.modulus for f in q_coeffs], 0x1234)
    assert np.array_equal(np.asarray(proof.commitment_words), want)
    ctx.close()
