"""GPU parity tests for SURVEY row N1 (-m gpu): cyclic transforms and the quotient pipeline through the C ABI,
bit for bit against oracle/quotient.py (the restated Rust of ntt.rs / r1cs.rs)."""
import random

import numpy as np
import pytest

from conftest import Q0, Q1, Q45, Q50, Q60
from lambda_snark_r_b200 import api, capi, sharding
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


@pytest.mark.parametrize("q,m", [(P, 1024), (Q0, 2048), (Q50, 1024), (Q60, 1024), (Q0, 32), (P, 64)])
def test_quotient_with_ragged_rows(gpu, q, m):
    """Rows with 0 .. 4 entries per matrix (the mat-vec kernel's batched first entry + the loop for the rest), empty
    rows, wrapped 'negative' values: quotient word for word against the restated Rust.  The moduli cover every arithmetic
    policy of the fused inverse transform (Goldilocks, FP64 butterflies, lazy and guarded u64), the sizes its one-kernel
    and multi-kernel shapes."""
    rng = random.Random(7 * m + 1)
    v = 40                                                     # free variables z[1 .. v]; z[v + 1 + i] closes constraint i
    z = [1] + [rng.randrange(q) for _ in range(v)] + [0] * m
    A, B, C = [], [], []

    def row(mat, i, lo):
        acc = 0
        for _ in range(rng.choice([lo, 1, 1, 2, 3, 4])):
            col, val = rng.randrange(v + 1), rng.choice([1, 2, q - 1, rng.randrange(q), (1 << 64) - 1])
            mat.append((i, col, val))
            acc = (acc + (val % q) * z[col]) % q
        return acc

    for i in range(m):
        a, b = row(A, i, 0), row(B, i, 0)
        extra = row(C, i, 0) if i % 3 == 0 else 0              # C row = closing variable (+ a few more entries)
        z[v + 1 + i] = (a * b - extra) % q
        C.append((i, v + 1 + i, 1))
    r = api.R1CS(m, v + 1 + m, A, B, C, q)
    omega = api.reference_root_of_unity(q, m)                  # the interpolation domain's generator, same on both sides
    want = QO.compute_quotient_poly(m, A, B, C, z, q, omega)
    got = r.quotient(np.array(z, dtype=np.uint64), omega)
    assert [int(x) for x in got] == want
    bad = list(z); bad[v + 1 + m // 2] = (bad[v + 1 + m // 2] + 1) % q
    _, status = r.quotient_batch(np.array([z, bad], dtype=np.uint64), omega)
    assert status.tolist() == [0, 1]
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


# ------------------------------------------------------------------ sizes beyond 2^17 (BASELINE configs[4])
def np_mult_gates(m, q, seed):
    """mult_gates as numpy triples (tests/integration_matrix.rs:60-75 shape), for millions of constraints."""
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, q, size=m, dtype=np.uint64)
    b = rng.integers(0, q, size=m, dtype=np.uint64)
    c = np.array([(int(x) * int(y)) % q for x, y in zip(a, b)], dtype=np.uint64)
    z = np.zeros(3 * m + 1, dtype=np.uint64)
    z[0] = 1
    z[1::3], z[2::3], z[3::3] = a, b, c
    rows = np.arange(m, dtype=np.uint32)
    one = np.ones(m, dtype=np.uint64)
    return 3 * m + 1, (rows, 3 * rows + 1, one), (rows, 3 * rows + 2, one), (rows, 3 * rows + 3, one), z


@pytest.mark.parametrize("q,logn", [(P, 18), (P, 19), (P, 20), (P, 21), (P, 23), (P, 24), (Q1, 18), (Q45, 19), (Q50, 19),
                                    (Q60, 18)])
def test_big_cyclic_transform_matches_c_oracle(gpu, q, logn):
    """Column stages ahead of the 4096-blocks: log n - 12 = 6, 7, 8 (one shared-memory pass of 3+3, 4+3, 4+4
    stages), 9, 11, 12 (a register pass + the 4+4 pass), under all four arithmetic policies (Goldilocks, FP64,
    lazy u64, guarded u64), against the C restatement of ntt.rs."""
    from oracle import oracle as O
    n = 1 << logn
    rng = np.random.Generator(np.random.PCG64(logn))
    w = api.reference_root_of_unity(q, n)
    assert pow(w, n // 2, q) == q - 1
    c = api.CyclicNtt(q, n)
    x = rng.integers(0, q, size=(2, n), dtype=np.uint64)
    x[1, :] = q - 1
    x[1, 5] = 0
    ev = c.forward_natural(x)
    for b in range(2):
        assert np.array_equal(ev[b], O.cyclic_ntt_forward(x[b], q, w)), (q, logn, b)
    assert np.array_equal(c.inverse_natural(ev), x)
    c.close()


@pytest.mark.parametrize("q,logm", [(P, 17), (P, 20), (Q1, 17)])
def test_big_quotient_matches_c_oracle(gpu, q, logm):
    from oracle import oracle as O
    m = 1 << logm
    cols, A, B, C, z = np_mult_gates(m, q, 40 + logm)
    r = api.R1CS.from_arrays(m, cols, A, B, C, q)
    w, w2 = api.reference_root_of_unity(q, m), api.reference_root_of_unity(q, 2 * m)
    want, st = O.r1cs_quotient(m, cols, A, B, C, z, q, w, w2)
    assert st == 0
    bad = z.copy(); bad[3] = (int(bad[3]) + 1) % q
    out, status = r.quotient_batch(np.stack([z, bad]))
    assert status.tolist() == [0, 1]
    assert np.array_equal(out[0], want)
    r.close()


@pytest.mark.parametrize("logm", [10, 14, 20])
def test_prover_commit_phase_matches_oracle_and_is_shard_invariant(gpu, logm):
    """BASELINE configs[4]: quotient -> ring-element chunks -> Module-LWE commitments without leaving the
    device.  Equal to (C oracle quotient) -> (oracle commitment) bit for bit, and to the concatenation of
    disjoint chunk slices (what ranks of a sharded job produce)."""
    from oracle import oracle as O
    q, m, n, k = P, 1 << logm, 4096, 2
    seed32 = bytes(range(32))
    cols, A, B, C, z = np_mult_gates(m, q, 70 + logm)
    r = api.R1CS.from_arrays(m, cols, A, B, C, q)
    ctx = api.LweContext(api.Params(n=n, k=k, q=Q0, sigma=3.19), seed32=seed32)
    chunks, planes = r.quotient_chunks(ctx), r.quotient_planes(ctx)
    assert planes == sharding.message_planes(ctx.p, q) == 4 and chunks == max(1, m // n) * planes
    seeds = np.arange(1, 2 * chunks + 1, dtype=np.uint64).reshape(2, chunks) * np.uint64(0x9E3779B97F4A7C15)
    bad = z.copy(); bad[3] = (int(bad[3]) + 1) % q
    got, status = r.commit_quotient(ctx, np.stack([z, bad]), seeds)
    assert status.tolist() == [0, 1] and got.shape == (2, chunks, ctx.words)
    want_q, st = O.r1cs_quotient(m, cols, A, B, C, z, q, api.reference_root_of_unity(q, m) if m > 1 else 1,
                                 api.reference_root_of_unity(q, 2 * m))
    assert st == 0
    orc = O.OracleLwe(Q0, n, k, 3.19, seed32)
    # units = (ring element, digit plane): the base-p digits of the quotient coefficients, which together bind them
    msgs = sharding.message_digits(want_q.reshape(chunks // planes, -1), ctx.p, planes)     # [chunks][min(m, n)]
    assert int(msgs.max()) < ctx.p
    recomposed = sum(msgs[l::planes].astype(object) * ctx.p ** l for l in range(planes))
    assert np.array_equal(recomposed, want_q.reshape(chunks // planes, -1).astype(object))
    assert np.array_equal(got[0], orc.commit_batch(msgs, seeds[0]))
    # the context opens what the prover committed
    assert ctx.verify_batch(got[0], msgs).tolist() == [1] * chunks
    if chunks >= 4:
        lo, _ = r.commit_quotient(ctx, z[None, :], seeds[:1], 0, chunks // 4)
        hi, _ = r.commit_quotient(ctx, z[None, :], seeds[:1], chunks // 4, chunks)
        assert np.array_equal(np.concatenate([lo[0], hi[0]]), got[0])
    with pytest.raises(api.LambdaSnarkError):
        r.commit_quotient(ctx, z[None, :], seeds[:1], 0, chunks + 1)
    ctx.close(); r.close()


@pytest.mark.parametrize("logm,count", [(18, 5), (16, 12)])
def test_prover_commit_pipeline_equals_the_serial_path(gpu, logm, count):
    """Host-io pipeline of lsr_prover_commit_quotient (groups of witnesses over copy-in / compute / copy-out streams, two
    buffers each): five groups of one witness (2^18) and groups of 5, 5, 2 witnesses (2^16), an unsatisfied witness in
    the middle, whole quotients and a chunk slice -- bit-identical to the one-shot path (one call per witness: a single
    group is never pipelined) and repeatable on the same handle."""
    q, m, n, k = P, 1 << logm, 4096, 2
    cols, A, B, C, z = np_mult_gates(m, q, 99)
    r = api.R1CS.from_arrays(m, cols, A, B, C, q)
    ctx = api.LweContext(api.Params(n=n, k=k, q=Q0, sigma=3.19), seed32=bytes(range(32)))
    chunks = r.quotient_chunks(ctx)
    zs = np.tile(z, (count, 1))
    rng = np.random.Generator(np.random.PCG64(5))
    for w in range(1, count):                           # different satisfied witnesses: a_i, b_i re-drawn per witness
        a = rng.integers(0, q, size=m, dtype=np.uint64)
        zs[w, 1::3] = a
        zs[w, 3::3] = np.array([(int(x) * int(y)) % q for x, y in zip(a.tolist(), zs[w, 2::3].tolist())], dtype=np.uint64)
    zs[2, 3] = (int(zs[2, 3]) + 1) % q                  # witness 2 violates constraint 0
    seeds = (np.arange(1, count * chunks + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)).reshape(count, chunks)
    one = [r.commit_quotient(ctx, zs[w:w + 1], seeds[w:w + 1]) for w in range(count)]
    want, st_want = np.concatenate([c for c, _ in one]), np.concatenate([s for _, s in one])
    want_slice = np.concatenate([r.commit_quotient(ctx, zs[w:w + 1], seeds[w:w + 1], 3, chunks - 5)[0] for w in range(count)])
    assert st_want.tolist() == [0, 0, 1] + [0] * (count - 3)
    for _ in range(2):
        got, st = r.commit_quotient(ctx, zs, seeds)
        assert st.tolist() == st_want.tolist() and np.array_equal(got, want)
        got_slice, st = r.commit_quotient(ctx, zs, seeds, 3, chunks - 5)
        assert st.tolist() == st_want.tolist() and np.array_equal(got_slice, want_slice)
    assert np.array_equal(want[:, 3:chunks - 5], want_slice)
    assert not np.array_equal(want[0], want[1])
    ctx.close(); r.close()


def test_goldilocks_primitives_on_boundary_operands(gpu):
    """gold_mul / gold_add_lazy / gold_sub / gold_add on every pair of boundary words (0, 1, 2^32 +- 1, q - 1, q, q + 1,
    2^64 - 1, ...) and on random pairs, against Python integers."""
    q = P
    eps = 2**32 - 1
    edge = [0, 1, 2, eps - 1, eps, eps + 1, 2**32 + 1, 2**33, 2**63 - 1, 2**63, 2**63 + 1, q - eps - 1, q - eps, q - 2, q - 1, q,
            q + 1, q + 2, 2**64 - eps - 1, 2**64 - eps, 2**64 - 2**32, 2**64 - 2, 2**64 - 1, (q - 1) // 2, (q + 1) // 2,
            0xFFFFFFFF_00000000, 0x00000000_FFFFFFFF, 0xFFFFFFFE_FFFFFFFF, 0x80000000_80000000, 0xFFFF0000_0000FFFF]
    rng = np.random.Generator(np.random.PCG64(64))
    a = [x for x in edge for _ in edge] + rng.integers(0, 2**64, size=200000, dtype=np.uint64).tolist()
    b = [y for _ in edge for y in edge] + rng.integers(0, 2**64, size=200000, dtype=np.uint64).tolist()
    # random operands concentrated where the corrections fire: high words all ones / zero
    hi = rng.integers(0, 2**32, size=50000, dtype=np.uint64).tolist()
    a += [(0xFFFFFFFF << 32) | x for x in hi]
    b += [(0xFFFFFFFF << 32) | y for y in reversed(hi)]
    a += [x for x in hi]
    b += [(0xFFFFFFFF << 32) | y for y in hi]
    A, B = np.array(a, dtype=np.uint64), np.array(b, dtype=np.uint64)
    out = np.zeros((A.size, 4), dtype=np.uint64)
    lib = capi.load()
    assert lib.lsr_goldilocks_probe_device(A.ctypes.data_as(capi.u64p), B.ctypes.data_as(capi.u64p), A.size,
                                           out.ctypes.data_as(capi.u64p)) == 0
    want = np.array([[x * y % q, (x + y % q) % q, (x - y % q) % q, (x % q + y % q) % q] for x, y in zip(a, b)], dtype=np.uint64)
    bad = np.nonzero((out != want).any(axis=1))[0]
    assert bad.size == 0, [(hex(a[i]), hex(b[i]), out[i].tolist(), want[i].tolist()) for i in bad[:5]]
