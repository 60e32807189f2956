"""GPU parity tests for SURVEY row N1 (-m gpu): cyclic transforms and the quotient pipeline through the C ABI,
bit for bit against oracle/quotient.py (the restated Rust of ntt.rs / r1cs.rs)."""
import random

import numpy as np
import pytest

from conftest import Q0, Q1
from lambda_snark_r_b200 import api, capi
from oracle import quotient as QO
from test_oracle_quotient import mult_gates

pytestmark = pytest.mark.gpu
P = QO.NTT_MODULUS


def test_reference_ntt_rs_unit_tests_on_device(gpu):
    """rust-api/lambda-snark/src/ntt.rs:284-347, assertion for assertion, through lsr_cyclic_ntt_*."""
    for n, coeffs in ((2, [1, 2]), (4, [1, 2, 3, 4]), (8, list(range(1, 9)))):
        c = api.CyclicNtt(P, n)
        ev = c.forward_natural(np.array(coeffs, dtype=np.uint64))
        assert int(ev[0]) == sum(coeffs)
        if n == 2:
            assert [int(v) for v in ev] == [3, P - 1]
        assert [int(v) for v in c.inverse_natural(ev)] == coeffs
        c.close()
    for log_n in range(1, 11):
        n = 1 << log_n
        c = api.CyclicNtt(P, n)
        coeffs = np.array([(i * 123456789) % P for i in range(n)], dtype=np.uint64)
        ev = c.forward_natural(coeffs)
        assert [int(v) for v in ev] == QO.ntt_forward([int(v) for v in coeffs], P, QO.compute_root_of_unity(n))
        assert np.array_equal(c.inverse_natural(ev), coeffs)
        c.close()
    assert api.reference_root_of_unity(P, 8) == QO.compute_root_of_unity(8)
    assert api.reference_root_of_unity(Q0, 8192) == 9037003627149            # r1cs.rs:546


@pytest.mark.parametrize("q,n", [(P, 2), (P, 16), (P, 256), (P, 4096), (P, 8192), (P, 65536), (P, 131072),
                                 (Q0, 64), (Q0, 4096), (Q0, 8192), (Q1, 32768)])
def test_cyclic_transform_matches_oracle(gpu, q, n):
    rng = np.random.Generator(np.random.PCG64(n))
    w = api.reference_root_of_unity(q, n)
    assert pow(w, n, q) == 1 and pow(w, n // 2, q) == q - 1
    c = api.CyclicNtt(q, n)
    batch = 3 if n <= 8192 else 2
    x = rng.integers(0, q, size=(batch, n), dtype=np.uint64)
    x[0, :] = q - 1
    if q == P:
        x[1, : min(n, 8)] = np.uint64(2**64 - 1)       # >= q: reduced on load
    ev = c.forward_natural(x)
    rows = 1 if n > 8192 else batch                     # the Python oracle is O(n log n) big-integer work
    for b in range(rows):
        want = QO.ntt_forward([int(v) % q for v in x[b]], q, w)
        assert [int(v) for v in ev[b]] == want, (q, n, b)
    assert np.array_equal(c.inverse_natural(ev), x % np.uint64(q) if q == P else x)
    # both arithmetic policies for the 44-bit modulus
    if q == Q0:
        c.set_arith(1)
        assert np.array_equal(c.forward_natural(x), ev)
    c.close()


def test_goldilocks_pointwise_is_exact(gpu):
    rng = np.random.Generator(np.random.PCG64(3))
    c = api.CyclicNtt(P, 16)
    a = rng.integers(0, 2**64, size=4099, dtype=np.uint64)
    b = rng.integers(0, 2**64, size=4099, dtype=np.uint64)
    a[:4] = [2**64 - 1, P - 1, P, 0]
    b[:4] = [2**64 - 1, P - 1, P, 5]
    got = c.mul_pointwise_batch(a, b)
    assert [int(v) for v in got] == [(int(x) * int(y)) % P for x, y in zip(a, b)]
    c.close()


@pytest.mark.parametrize("q", [P, Q0])
@pytest.mark.parametrize("m", [1, 2, 4, 16, 128, 1024])
def test_quotient_matches_restated_rust(gpu, q, m):
    rng = random.Random(100 + m)
    cols, A, B, C, z = mult_gates(m, q, rng)
    # mix in linear terms and wrapped "negative" values the way tests/test_vectors.rs:63 builds them
    if m >= 4:
        A.append((1, 0, 5)); C.append((1, 0, (5 * z[3 * 1 + 2]) % q))          # (z4 + 5) * z5 = z6 + 5 z5
    r = api.R1CS(m, cols, A, B, C, q)
    want = QO.compute_quotient_poly(m, A, B, C, z, q)
    got = r.quotient(np.array(z, dtype=np.uint64))
    assert [int(v) for v in got] == want
    bad = list(z); bad[3] = (bad[3] + 1) % q
    with pytest.raises(api.LambdaSnarkError):
        r.quotient(np.array(bad, dtype=np.uint64))
    # batch entry point: valid, invalid, valid
    W = np.array([z, bad, z], dtype=np.uint64)
    out, status = r.quotient_batch(W)
    assert status.tolist() == [0, 1, 0]
    pad = want + [0] * (m - len(want))
    assert [int(v) for v in out[0]] == pad and [int(v) for v in out[2]] == pad
    r.close()


def test_quotient_test_vectors(gpu):
    """TV-1 (7 * 13 = 91) and a plaquette-style instance with wrapped negative coefficients."""
    q = P
    r = api.R1CS(1, 4, [(0, 1, 1)], [(0, 2, 1)], [(0, 3, 1)], q)        # test-vectors/tv-1-multiplication
    assert [int(v) for v in r.quotient(np.array([1, 7, 13, 91], dtype=np.uint64))] == [0]
    with pytest.raises(api.LambdaSnarkError):
        r.quotient(np.array([1, 7, 13, 92], dtype=np.uint64))
    r.close()
    # two constraints, B has a wrapped -1 (2^64 - 1 as a word, reduced as an unsigned word: sparse_matrix.rs:279)
    neg1 = 2**64 - 1
    A = [(0, 1, 1), (1, 3, 1)]
    B = [(0, 2, 1), (1, 0, neg1)]
    z = [1, 314, 628, (314 * 628) % q, 0]
    z[4] = (z[3] * (neg1 % q)) % q
    C = [(0, 3, 1), (1, 4, 1)]
    r = api.R1CS(2, 5, A, B, C, q)
    want = QO.compute_quotient_poly(2, A, B, C, z, q)
    assert [int(v) for v in r.quotient(np.array(z, dtype=np.uint64))] == want
    r.close()


def test_quotient_full_size_properties(gpu):
    """m = 2^16 (the largest size: 2m = 2^17 transforms), checked by the polynomial identity at a random point
    and against the oracle's inverse transforms on the same evaluations."""
    q, m = P, 1 << 16
    rng = random.Random(7)
    cols, A, B, C, z = mult_gates(m, q, rng)
    r = api.R1CS(m, cols, A, B, C, q)
    quo = [int(v) for v in r.quotient(np.array(z, dtype=np.uint64))]
    assert 1 <= len(quo) <= m - 1
    w = QO.reference_root(q, m)
    c = api.CyclicNtt(q, m)
    evals = np.array([QO.mul_vec(m, M, z, q) for M in (A, B, C)], dtype=np.uint64)
    ap, bp, cp = ([int(v) for v in row] for row in c.inverse_natural(evals))
    assert QO.horner(ap, w, q) == int(evals[0][1])                  # interpolation really passes through the evaluations
    for _ in range(3):
        x = rng.randrange(q)
        lhs = (QO.horner(quo, x, q) * (pow(x, m, q) - 1)) % q
        assert lhs == (QO.horner(ap, x, q) * QO.horner(bp, x, q) - QO.horner(cp, x, q)) % q
    c.close(); r.close()


def test_quotient_rejects_bad_shapes(gpu):
    lib = capi.load()
    r = api.R1CS(3, 4, [(0, 1, 1)], [(0, 2, 1)], [(0, 3, 1)], P)          # m = 3 is not a power of two
    with pytest.raises(api.LambdaSnarkError):
        r.quotient(np.array([1, 7, 13, 91], dtype=np.uint64))
    r.close()
    r = api.R1CS(2, 4, [(0, 1, 1)], [(0, 2, 1)], [(0, 3, 1)], 2**44 + 1)  # composite modulus of the reference's tests
    with pytest.raises(api.LambdaSnarkError):
        r.quotient(np.array([1, 7, 13, 91], dtype=np.uint64))
    r.close()
    assert not lib.lsr_cyclic_ntt_context_create(P, 6, 0)
    assert not lib.lsr_cyclic_ntt_context_create(Q0, 1 << 14, 0)          # 2-adicity of q0 is 13
    assert not lib.lsr_cyclic_ntt_context_create(P, 8, 3)                 # not a primitive 8th root
