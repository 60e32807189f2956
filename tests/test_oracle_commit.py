"""CPU tests (-m "not gpu"): the oracle's Module-LWE commitment behaves as the
reference's commitment API is tested to behave (cpp-core/tests/test_commitment.cpp,
rust-api/lambda-snark/src/commitment.rs:136-219, tests/lwe_verification.rs), and
is pinned against the committed digests (tests/golden/commit_kat.json)."""
import hashlib
import json
from pathlib import Path

import numpy as np
import pytest

from conftest import Q0, Q1
from oracle import oracle as O

GOLD = Path(__file__).resolve().parent / "golden"
SEED32 = bytes(range(32))


@pytest.fixture(scope="module")
def ctx():
    # test_commitment.cpp:12-18 uses modulus 12289 with n = 4096: not NTT-friendly for
    # that degree, so the ring modulus falls back to 17592169062401 (the reference
    # ignores the field altogether, commitment.cpp:108-111)
    return O.OracleLwe(12289, 4096, 2, 3.19, SEED32)


def test_parameters(ctx):
    assert ctx.q == Q0
    assert ctx.p == 204800 and ctx.delta == 85899263 and ctx.p * ctx.delta == Q0 - 1
    assert ctx.words == 1 + 2 * 4096
    assert O.OracleLwe(Q1, 8192, 2, 3.19, SEED32).q == Q1
    assert O.OracleLwe(12345, 8192, 2, 3.19, SEED32).q == Q1


def test_golden_digests():
    g = json.loads((GOLD / "commit_kat.json").read_text())
    for c in g["cases"]:
        o = O.OracleLwe(Q0, c["n"], c["k"], c["sigma"], SEED32)
        assert (o.q, o.p, o.delta) == (c["q"], c["p"], c["delta"])
        assert hashlib.sha256(o.matrix().tobytes()).hexdigest() == c["matrix_sha256"]
        cm = o.commit(c["msg"], c["seed"])
        assert cm.size == c["words"]
        assert [str(int(v)) for v in cm[:5]] == c["first"]
        assert hashlib.sha256(cm.tobytes()).hexdigest() == c["sha256"]


def test_golden_digests_explicit_mode():
    sys_path_golden = str(GOLD)
    import sys
    if sys_path_golden not in sys.path:
        sys.path.insert(0, sys_path_golden)
    from make_golden import explicit_inputs
    g = json.loads((GOLD / "commit_kat.json").read_text())
    assert len(g["explicit"]) >= 3
    for c in g["explicit"]:
        o = O.OracleLwe(Q0, c["n"], c["k"], c["sigma"], SEED32)
        s, e, msg = explicit_inputs(o.q, c["n"], c["k"], c["pcg64_seed"])
        cm = o.commit_explicit(msg, s, e)
        assert [str(int(v)) for v in cm[:5]] == c["first"]
        assert hashlib.sha256(cm.tobytes()).hexdigest() == c["sha256"]


def test_container_layout(ctx):
    cm = ctx.commit([1, 2, 3, 4], 0x1234)                 # test_commitment.cpp:37-47
    assert cm.size == ctx.words and int(cm[0]) == 2 * 4096 * 8     # types.h:32-34
    assert (cm[1:] < np.uint64(ctx.q)).all()


def test_binding_and_determinism(ctx):
    a = ctx.commit([1, 2, 3], 0x1234)
    b = ctx.commit([4, 5, 6], 0x1234)                     # test_commitment.cpp:77-100
    assert not np.array_equal(a, b)
    assert np.array_equal(a, ctx.commit([1, 2, 3], 0x1234))
    assert not np.array_equal(a, ctx.commit([1, 2, 3], 0x1235))


def test_verify_opening_roundtrip(ctx):
    msg = [7, 11, 13, 17]                                 # test_commitment.cpp:115-132
    cm = ctx.commit(msg, 5)
    assert ctx.verify(cm, msg) == 1
    wrong = list(msg); wrong[1] ^= 1
    assert ctx.verify(cm, wrong) == 0
    assert ctx.verify(cm, msg[:2]) == 1                   # only the first msg_len slots are compared
    assert ctx.verify(cm, msg + [0, 0]) == 1              # padding slots decode to zero
    assert ctx.verify(cm, []) == 1
    assert ctx.verify(cm, [0] * 4097) == 0                # commitment.cpp:219-221


def test_verify_rejects_malformed_container(ctx):
    cm = ctx.commit([1], 9)
    bad = cm.copy(); bad[0] = 0
    assert ctx.verify(bad, [1]) == -1                     # commitment.cpp:71-73
    bad = cm.copy(); bad[0] = cm[0] + np.uint64(8)
    assert ctx.verify(bad, [1]) == -1
    bad = cm.copy(); bad[5] = np.uint64(ctx.q)
    assert ctx.verify(bad, [1]) == -1
    assert ctx.verify(cm[:100], [1]) == -1


def test_linear_combination_homomorphism(ctx):
    m1, m2 = [1, 2, 3, 4], [5, 6, 7, 8]                   # test_commitment.cpp:134-166
    c1, c2 = ctx.commit(m1, 11), ctx.commit(m2, 12)
    lc = ctx.linear_combine([c1, c2], [2, 3])
    expected = [2 * a + 3 * b for a, b in zip(m1, m2)]
    assert ctx.verify(lc, expected) == 1
    expected[0] += 1
    assert ctx.verify(lc, expected) == 0
    # commitment.rs:163-219: seeds 0/1, coefficients 2 and 3, messages i+1 and 2(i+1)
    lc2 = ctx.linear_combine([c1, None, c2], [2, 99, 3])   # NULL entries are skipped (commitment.cpp:248-250)
    assert np.array_equal(lc, lc2)
    assert ctx.linear_combine([None, None], [1, 2]) is None
    # messages wrap modulo the plaintext modulus, coefficients are reduced mod p (commitment.cpp:90)
    big = ctx.commit([ctx.p - 1], 13)
    assert ctx.verify(ctx.linear_combine([big], [2]), [ctx.p - 2]) == 1
    assert np.array_equal(ctx.linear_combine([c1], [ctx.p + 2]), ctx.linear_combine([c1], [2]))


def test_truncation_and_padding(ctx):
    n = 4096                                              # commitment.cpp:146-149
    long_msg = list(range(1, n + 50))
    assert np.array_equal(ctx.commit(long_msg, 3), ctx.commit(long_msg[:n], 3))
    assert np.array_equal(ctx.commit([1, 2], 3), ctx.commit([1, 2] + [0] * 10, 3))
    assert ctx.verify(ctx.commit(long_msg, 3), [v % ctx.p for v in long_msg[:n]]) == 1


def test_message_words_are_bound_modulo_the_plain_modulus(ctx):
    # a message word >= p is encoded mod p (lambda_snark_b200.h, lwe_commit MESSAGE RANGE) and verification compares
    # mod p, so the library's own commitment to such a word opens; whole words are bound through digit planes
    cm = ctx.commit([ctx.p + 5], 21)
    assert ctx.verify(cm, [ctx.p + 5]) == 1
    assert ctx.verify(cm, [5]) == 1
    assert ctx.verify(cm, [6]) == 0 and ctx.verify(cm, [ctx.p + 6]) == 0


def test_noise_is_small_and_samples_match_definition(ctx):
    s, e = ctx.sample_se(77)
    assert s.shape == e.shape == (2, 4096)
    assert np.abs(s).max() <= 39 and np.abs(e).max() <= 39          # ceil(12 sigma) = 39
    assert 2.8 < s.std() < 3.6 and 2.8 < e.std() < 3.6
    # t - A*s - e - Delta*m == 0, recomputed with the oracle's own NTT
    ntt = O.OracleNtt(ctx.q, 4096)
    msg = np.arange(10, dtype=np.uint64)
    cm = ctx.commit(msg, 77)
    A = ctx.matrix()
    sh = ntt.forward(np.where(s < 0, s + ctx.q, s).astype(np.uint64))
    for i in range(2):
        acc = np.zeros(4096, dtype=object)
        for j in range(2):
            acc = (acc + A[i, j].astype(object) * sh[j].astype(object)) % ctx.q
        t = ntt.inverse(np.array(acc, dtype=np.uint64)).astype(object)
        t = (t + e[i].astype(object)) % ctx.q
        if i == 1:
            t[:10] = (t[:10] + ctx.delta * msg.astype(object)) % ctx.q
        assert [int(v) for v in cm[1 + i * 4096: 1 + (i + 1) * 4096]] == [int(v) for v in t]


@pytest.mark.parametrize("n,k", [(16, 1), (64, 2), (1024, 3), (8192, 2)])
def test_other_shapes(n, k):
    o = O.OracleLwe(Q0, n, k, 3.19, SEED32)
    cm = o.commit([3, 1, 4, 1, 5], 42)
    assert cm.size == 1 + n * k and o.verify(cm, [3, 1, 4, 1, 5]) == 1 and o.verify(cm, [3, 1, 4, 1, 6]) == 0


def test_batch_equals_single_and_thread_count_is_irrelevant(ctx, rng):
    msgs = rng.integers(0, 2**64, size=(6, 33), dtype=np.uint64)
    seeds = np.arange(100, 106, dtype=np.uint64)
    one = ctx.commit_batch(msgs, seeds, threads=1)
    many = ctx.commit_batch(msgs, seeds, threads=4)
    assert np.array_equal(one, many)
    for i in range(6):
        assert np.array_equal(one[i], ctx.commit(msgs[i], int(seeds[i])))


# ------------------------------------------------------------ explicit mode (s, e supplied)
def _schoolbook_negacyclic(a, b, q):
    n = len(a)
    out = [0] * n
    for i, x in enumerate(a):
        if x == 0:
            continue
        for j, y in enumerate(b):
            t = i + j
            if t < n:
                out[t] = (out[t] + x * y) % q
            else:
                out[t - n] = (out[t - n] - x * y) % q
    return out


def test_explicit_mode_reproduces_the_seeded_commitment(ctx):
    msg = np.arange(50, dtype=np.uint64) * 977
    for seed in (1, 0xC0FFEE):
        s, e = ctx.sample_se(seed)
        assert np.array_equal(ctx.commit_explicit(msg, s, e), ctx.commit(msg, seed))


def test_explicit_mode_is_A_s_plus_e_plus_delta_m_by_schoolbook_products():
    # small ring so that the O(n^2) negacyclic product finishes; A is taken back to coefficients with the
    # oracle's own inverse transform (pinned in test_oracle_pinning.py)
    n, k = 64, 2
    o = O.OracleLwe(Q0, n, k, 3.19, SEED32)
    ntt = O.OracleNtt(o.q, n)
    q = o.q
    rng = np.random.Generator(np.random.PCG64(7))
    s = rng.integers(-(q - 1), q, size=(k, n), dtype=np.int64)          # the whole residue range, both signs
    e = rng.integers(-(q - 1), q, size=(k, n), dtype=np.int64)
    s[0, 0], s[0, 1], e[1, 0], e[1, 1] = np.iinfo(np.int64).min, np.iinfo(np.int64).max, np.iinfo(np.int64).min, -q
    msg = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    cm = o.commit_explicit(msg, s, e)
    assert int(cm[0]) == 8 * k * n
    A = [[[int(v) for v in ntt.inverse(o.matrix()[i, j])] for j in range(k)] for i in range(k)]
    for i in range(k):
        t = [int(v) % q for v in e[i]]
        for j in range(k):
            prod = _schoolbook_negacyclic(A[i][j], [int(v) % q for v in s[j]], q)
            t = [(a + b) % q for a, b in zip(t, prod)]
        if i == k - 1:
            t = [(a + o.delta * (int(m) % o.p)) % q for a, m in zip(t, msg)]
        assert [int(v) for v in cm[1 + i * n: 1 + (i + 1) * n]] == t


def test_explicit_mode_is_linear_in_s_and_e(ctx, rng):
    q = ctx.q
    s1, e1 = ctx.sample_se(5)
    s2 = rng.integers(-1000, 1000, size=s1.shape, dtype=np.int64)
    e2 = rng.integers(-1000, 1000, size=s1.shape, dtype=np.int64)
    m1 = rng.integers(0, 1000, size=4096, dtype=np.uint64)                # no wrap of the plaintext modulus
    m2 = rng.integers(0, 1000, size=4096, dtype=np.uint64)
    a = ctx.commit_explicit(m1, s1, e1)[1:].astype(object)
    b = ctx.commit_explicit(m2, s2, e2)[1:].astype(object)
    both = ctx.commit_explicit(m1 + m2, s1 + s2, e1 + e2)[1:].astype(object)
    assert np.array_equal((a + b) % q, both)
    zero = np.zeros_like(s1)
    only_m = ctx.commit_explicit(m1, zero, zero)
    assert not only_m[1:1 + 4096].any()
    assert np.array_equal(only_m[1 + 4096:], (m1 % np.uint64(ctx.p)) * np.uint64(ctx.delta))
