"""CPU tests (-m "not gpu"): the quotient oracle against the reference's own unit tests
(rust-api/lambda-snark/src/ntt.rs:284-376) and against first principles."""
import random

from oracle import quotient as Q

P = Q.NTT_MODULUS


def test_reference_ntt_rs_unit_tests():
    # test_compute_root_of_unity (ntt.rs:262-282)
    w2, w4, w8 = (Q.compute_root_of_unity(n) for n in (2, 4, 8))
    assert w2 == P - 1 and pow(w4, 4, P) == 1 and pow(w4, 2, P) == P - 1
    assert pow(w8, 8, P) == 1 and pow(w8, 4, P) == P - 1
    # test_ntt_2_point / 4 / 8 (ntt.rs:284-331)
    ev = Q.ntt_forward([1, 2], P, w2)
    assert ev == [3, P - 1] and Q.ntt_inverse(ev, P, w2) == [1, 2]
    ev = Q.ntt_forward([1, 2, 3, 4], P, w4)
    assert ev[0] == 10 and Q.ntt_inverse(ev, P, w4) == [1, 2, 3, 4]
    ev = Q.ntt_forward(list(range(1, 9)), P, w8)
    assert ev[0] == 36 and Q.ntt_inverse(ev, P, w8) == list(range(1, 9))
    # test_ntt_inverse_correctness (ntt.rs:333-347)
    for log_n in range(1, 11):
        n = 1 << log_n
        w = Q.compute_root_of_unity(n)
        coeffs = [(i * 123456789) % P for i in range(n)]
        assert Q.ntt_inverse(Q.ntt_forward(coeffs, P, w), P, w) == coeffs
    # bit reversal (ntt.rs:233-260)
    assert [Q.reverse_bits(i, 3) for i in range(8)] == [0, 4, 2, 6, 1, 5, 3, 7]
    d = list(range(8)); Q.bit_reverse_permutation(d)
    assert d == [0, 4, 2, 6, 1, 5, 3, 7]


def test_forward_is_evaluation_at_powers_of_omega():
    rng = random.Random(5)
    for q, n in ((P, 16), (Q.NTT_FRIENDLY_MODULUS, 32)):
        w = Q.reference_root(q, n)
        f = [rng.randrange(q) for _ in range(n)]
        assert Q.ntt_forward(f, q, w) == [Q.horner(f, pow(w, j, q), q) for j in range(n)]
    # r1cs.rs:534-547 ROOTS_OF_UNITY table
    assert Q.reference_root(Q.NTT_FRIENDLY_MODULUS, 4) == 981206394875
    assert Q.reference_root(Q.NTT_FRIENDLY_MODULUS, 8192) == 9037003627149


def mult_gates(m, q, rng):
    """m multiplication gates z[3i+1] * z[3i+2] = z[3i+3] (tests/integration_matrix.rs:60-75 shape)."""
    cols = 3 * m + 1
    z = [1] + [0] * (3 * m)
    A, B, C = [], [], []
    for i in range(m):
        a, b = rng.randrange(q), rng.randrange(q)
        z[3 * i + 1], z[3 * i + 2], z[3 * i + 3] = a, b, (a * b) % q
        A.append((i, 3 * i + 1, 1)); B.append((i, 3 * i + 2, 1)); C.append((i, 3 * i + 3, 1))
    return cols, A, B, C, z


def test_quotient_identity_and_rejection():
    rng = random.Random(11)
    for q in (P, Q.NTT_FRIENDLY_MODULUS):
        for m in (1, 2, 8, 64):
            cols, A, B, C, z = mult_gates(m, q, rng)
            quo = Q.compute_quotient_poly(m, A, B, C, z, q)
            assert len(quo) <= max(m - 1, 1)
            if m > 1:
                w = Q.reference_root(q, m)
                ap, bp, cp = (Q.ntt_inverse(Q.mul_vec(m, M, z, q), q, w) for M in (A, B, C))
                x = rng.randrange(q)
                lhs = (Q.horner(quo, x, q) * (pow(x, m, q) - 1)) % q
                assert lhs == (Q.horner(ap, x, q) * Q.horner(bp, x, q) - Q.horner(cp, x, q)) % q
            bad = list(z); bad[3] = (bad[3] + 1) % q
            try:
                Q.compute_quotient_poly(m, A, B, C, bad, q)
                assert False, "invalid witness accepted"
            except ValueError:
                pass


# ---------------------------------------------------------------- C restatement (oracle/lsr_oracle_quotient.c)
def _triples(M):
    import numpy as np
    return (np.array([e[0] for e in M], dtype=np.uint32), np.array([e[1] for e in M], dtype=np.uint32),
            np.array([e[2] % 2**64 for e in M], dtype=np.uint64))


def test_c_cyclic_transform_passes_the_reference_unit_tests_and_matches_the_python_restatement():
    from oracle import oracle as O
    # ntt.rs:284-331
    assert [int(v) for v in O.cyclic_ntt_forward([1, 2], P, Q.compute_root_of_unity(2))] == [3, P - 1]
    assert int(O.cyclic_ntt_forward([1, 2, 3, 4], P, Q.compute_root_of_unity(4))[0]) == 10
    assert int(O.cyclic_ntt_forward(list(range(1, 9)), P, Q.compute_root_of_unity(8))[0]) == 36
    rng = random.Random(3)
    for q, n in ((P, 1), (P, 2), (P, 64), (P, 2048), (Q.NTT_FRIENDLY_MODULUS, 512)):
        w = Q.reference_root(q, n) if n > 1 else 1
        f = [rng.randrange(q) for _ in range(n)]
        ev = O.cyclic_ntt_forward(f, q, w)
        assert [int(v) for v in ev] == Q.ntt_forward(f, q, w)
        assert [int(v) for v in O.cyclic_ntt_inverse(ev, q, w)] == f
    # ntt.rs:333-347 round trips
    for log_n in range(1, 11):
        n = 1 << log_n
        w = Q.compute_root_of_unity(n)
        coeffs = [(i * 123456789) % P for i in range(n)]
        assert [int(v) for v in O.cyclic_ntt_inverse(O.cyclic_ntt_forward(coeffs, P, w), P, w)] == coeffs


def test_c_quotient_matches_the_schoolbook_python_restatement():
    from oracle import oracle as O
    for q in (P, Q.NTT_FRIENDLY_MODULUS):
        for m in (1, 2, 4, 32, 128):
            rng = random.Random(1000 + m)
            cols, A, B, C, z = mult_gates(m, q, rng)
            if m >= 4:
                A.append((1, 0, 5)); C.append((1, 0, (5 * z[3 * 1 + 2]) % q))
                B.append((2, 0, 2**64 - 1)); C.append((2, 0, ((2**64 - 1) % q) * z[3 * 2 + 1] % q))   # wrapped -1
            want = Q.compute_quotient_poly(m, A, B, C, z, q)
            om = Q.reference_root(q, m) if m > 1 else 1
            got, st = O.r1cs_quotient(m, cols, _triples(A), _triples(B), _triples(C), z, q, om, Q.reference_root(q, 2 * m))
            assert st == 0
            assert [int(v) for v in got] == want + [0] * (m - len(want)), (q, m)
            bad = list(z); bad[3] = (bad[3] + 1) % q
            assert O.r1cs_quotient(m, cols, _triples(A), _triples(B), _triples(C), bad, q, om, Q.reference_root(q, 2 * m))[1] == 1


def test_six_transform_identity_of_the_device_pipeline():
    """The device pipeline never transforms C_z to the coset: Q = (C_z - N) / 2 with N = A_z * B_z mod (X^m + 1)
    (DESIGN.md 4.7).  Checked here against the reference's own route (schoolbook product, long division by X^m - 1),
    for satisfied witnesses; for an unsatisfied one (C_z - N) / 2 still equals what the seven-transform flow
    ((a*b - c) * (-2)^-1 on the coset, inverse transform) produced, which is what the status flag accompanies."""
    rng = random.Random(23)
    for q in (P, Q.NTT_FRIENDLY_MODULUS):
        inv2 = q // 2 + 1
        assert (2 * inv2) % q == 1
        for m in (2, 8, 64):
            cols, A, B, C, z = mult_gates(m, q, rng)
            w = Q.reference_root(q, m)
            for witness in (z, [v if i != 3 else (v + 1) % q for i, v in enumerate(z)]):
                ap, bp, cp = (Q.ntt_inverse(Q.mul_vec(m, M, witness, q), q, w) for M in (A, B, C))
                prod = Q.poly_mul(ap, bp, q) + [0] * (2 * m)
                nega = [(prod[i] - prod[i + m]) % q for i in range(m)]               # A_z * B_z mod (X^m + 1)
                got = [((cp[i] - nega[i]) * inv2) % q for i in range(m)]
                if witness is z:
                    want = Q.compute_quotient_poly(m, A, B, C, z, q)
                    assert got[:len(want)] == want and not any(got[len(want):])
                else:
                    # seven-transform statement: values of (A*B - C) * (-2)^-1 at psi^(2j+1), interpolated back
                    psi = Q.reference_root(q, 2 * m)
                    half = (q - 1) // 2
                    assert (half * (q - 2)) % q == 1
                    pts = [pow(psi, 2 * j + 1, q) for j in range(m)]
                    vals = [((Q.horner(ap, x, q) * Q.horner(bp, x, q) - Q.horner(cp, x, q)) * half) % q for x in pts]
                    # the unique polynomial of degree < m with these values on the coset is `got`
                    assert [Q.horner(got, x, q) for x in pts] == vals
