"""GPU parity tests for SURVEY row N2 (-m gpu): Fiat-Shamir challenges, polynomial evaluations and the batched
prove_r1cs through the C ABI, against hashlib / the restated Rust (tests/prover_restatement.py, oracle/quotient.py)."""
import random

import numpy as np
import pytest

import prover_restatement as PR
from conftest import Q0
from lambda_snark_r_b200 import api
from oracle import oracle as O
from oracle import quotient as QO
from test_oracle_quotient import mult_gates

pytestmark = pytest.mark.gpu
P = QO.NTT_MODULUS
SEED32 = bytes(range(32))


@pytest.mark.parametrize("n_pub,words", [(0, 0), (1, 1), (2, 8193), (3, 17), (17, 100), (1, 4097)])
def test_challenges_match_challenge_rs(gpu, n_pub, words):
    rng = np.random.default_rng(n_pub * 100 + words)
    count = 37
    pub = rng.integers(0, 2**64, size=(count, n_pub), dtype=np.uint64)
    cont = rng.integers(0, 2**64, size=(count, words), dtype=np.uint64)
    for modulus in (P, Q0, 17592186044417):
        ch, hs = api.fs_challenge_batch(pub, cont, modulus, chain=True)
        for i in (0, 1, count - 1):
            alpha, d0 = PR.challenge_derive([int(v) for v in pub[i]], cont[i], modulus)
            beta, d1 = PR.challenge_derive([alpha], cont[i], modulus)
            assert (int(ch[i, 0]), int(ch[i, 1])) == (alpha, beta)
            assert hs[i, 0].tobytes() == d0 and hs[i, 1].tobytes() == d1
    ch, hs = api.fs_challenge_batch(pub, cont, P, chain=False)
    assert not ch[:, 1].any() and not hs[:, 1].any()


@pytest.mark.parametrize("q", [P, Q0, 17592186044423])
def test_poly_eval_matches_eval_poly(gpu, q):
    rng = random.Random(q % 1000)
    for length in (1, 2, 255, 256, 257, 4096, 5000):
        coeffs = [[rng.randrange(2**64) if j % 7 == 0 else rng.randrange(q) for j in range(length)] for _ in range(3)]
        pts = [[0, 1, q - 1, rng.randrange(q)] for _ in range(3)]
        got = api.poly_eval_batch(np.array(coeffs, dtype=np.uint64), np.array(pts, dtype=np.uint64), q)
        for p in range(3):
            for j in range(4):
                assert int(got[p, j]) == QO.horner([c % q for c in coeffs[p]], pts[p][j], q), (q, length, p, j)


@pytest.mark.parametrize("m", [1, 2, 16, 1024, 4096])
def test_prove_r1cs_batch_matches_the_restated_prover(gpu, m):
    """lib.rs:747-809 on the NTT path: commitment, alpha, beta and the eight evaluations bit for bit; the
    reference verifier's equation (lib.rs:1047-1078) holds at both challenges."""
    q, n, k = P, 4096, 2
    rng = random.Random(900 + m)
    cols, A, B, C, z = mult_gates(m, q, rng)
    l = min(2, cols)
    r = api.R1CS(m, cols, A, B, C, q)
    ctx = api.LweContext(api.Params(n=n, k=k, q=Q0, sigma=3.19), seed32=SEED32)
    bad = list(z); bad[3] = (bad[3] + 1) % q
    z2 = list(z); z2[1], z2[2] = z2[2], z2[1]                           # a second valid witness (a*b = b*a)
    W = np.array([z, bad, z2], dtype=np.uint64)
    seeds = np.array([0x1234, 7, 0xDEADBEEF], dtype=np.uint64)
    out = r.prove_batch(ctx, W, l, seeds)
    assert out["status"].tolist() == [0, 1, 0]
    orc = O.OracleLwe(Q0, n, k, 3.19, SEED32)
    omega = QO.reference_root(q, m) if m > 1 else 1
    for w, wit in ((0, z), (2, z2)):
        quo = QO.compute_quotient_poly(m, A, B, C, wit, q)
        quo_pad = quo + [0] * (m - len(quo))
        words = orc.commit_batch(np.array([quo_pad], dtype=np.uint64), seeds[w:w + 1])[0]
        assert np.array_equal(out["containers"][w], words)
        alpha, d0 = PR.challenge_derive(wit[:l], words, q)
        beta, d1 = PR.challenge_derive([alpha], words, q)
        assert (int(out["challenges"][w, 0]), int(out["challenges"][w, 1])) == (alpha, beta)
        assert out["hashes"][w, 0].tobytes() == d0 and out["hashes"][w, 1].tobytes() == d1
        ev = [QO.mul_vec(m, M, wit, q) for M in (A, B, C)]
        ap, bp, cp = (QO.ntt_inverse(e, q, omega) if m > 1 else e for e in ev)
        want = [QO.horner(quo, alpha, q), QO.horner(quo, beta, q)] + [QO.horner(pl, alpha, q) for pl in (ap, bp, cp)] + \
               [QO.horner(pl, beta, q) for pl in (ap, bp, cp)]
        got = [int(v) for v in out["evals"][w]]
        assert got == want
        for x, qx, ax, bx, cx in ((alpha, got[0], got[2], got[3], got[4]), (beta, got[1], got[5], got[6], got[7])):
            assert (qx * (pow(x, m, q) - 1)) % q == (ax * bx - cx) % q   # Q(x) Z_H(x) = A_z(x) B_z(x) - C_z(x)
    # the batched verifier accepts the two honest proofs and rejects tampered ones
    pubs = W[:, :l]
    ok = api.verify_r1cs_batch(m, q, pubs, out["containers"], out["challenges"], out["evals"])
    assert ok.tolist() == [1, 0, 1]                                    # row 1 is the unsatisfying witness
    bad_ev = out["evals"].copy(); bad_ev[0, 0] = (int(bad_ev[0, 0]) + 1) % q
    bad_ct = out["containers"].copy(); bad_ct[2, 5] ^= np.uint64(1)
    bad_pub = pubs.copy(); bad_pub[0, -1] = (int(bad_pub[0, -1]) + 1) % q
    assert api.verify_r1cs_batch(m, q, pubs, out["containers"], out["challenges"], bad_ev).tolist()[0] == 0
    assert api.verify_r1cs_batch(m, q, pubs, bad_ct, out["challenges"], out["evals"]).tolist()[2] == 0
    assert api.verify_r1cs_batch(m, q, bad_pub, out["containers"], out["challenges"], out["evals"]).tolist()[0] == 0
    ctx.close(); r.close()
