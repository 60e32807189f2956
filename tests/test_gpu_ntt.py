"""GPU parity tests for K1/K2/K3 (-m gpu): every call goes through the C ABI
and is compared BIT FOR BIT with the CPU oracle (integer work: no tolerance)."""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

from conftest import Q0, Q1, Q31, Q45, Q50, Q60, uniform
from lambda_snark_r_b200 import api, capi
from oracle import oracle as O

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).resolve().parent / "golden"


# ------------------------------------------------- the reference's own tests
def test_reference_test_ntt_cpp(gpu):
    """cpp-core/tests/test_ntt.cpp, assertion for assertion."""
    lib = capi.load()
    q, n = 12289, 256
    ctx = lib.ntt_context_create(q, n)
    assert ctx                                                            # CreateAndFree
    ones = np.ones(n, dtype=np.uint64)
    assert lib.ntt_forward(ctx, ones.ctypes.data_as(capi.u64p), n) == 0   # ForwardNttBasic
    ones = np.ones(n, dtype=np.uint64)
    assert lib.ntt_inverse(ctx, ones.ctypes.data_as(capi.u64p), n) == 0   # InverseNttBasic
    orig = np.zeros(n, dtype=np.uint64); orig[:8] = np.arange(1, 9)
    t = orig.copy()
    assert lib.ntt_forward(ctx, t.ctypes.data_as(capi.u64p), n) == 0      # ForwardInverseIdentity
    assert lib.ntt_inverse(ctx, t.ctypes.data_as(capi.u64p), n) == 0
    assert np.array_equal(t, orig)
    a = np.full(n, 2, dtype=np.uint64); b = np.full(n, 3, dtype=np.uint64); r = np.zeros(n, dtype=np.uint64)
    lib.ntt_mul_pointwise(ctx, r.ctypes.data_as(capi.u64p), a.ctypes.data_as(capi.u64p), b.ctypes.data_as(capi.u64p), n)
    assert (r == 6).all()                                                 # PointwiseMultiplication
    dummy = np.zeros(n, dtype=np.uint64)
    assert lib.ntt_forward(None, dummy.ctypes.data_as(capi.u64p), n) == -1   # NullPointerHandling
    assert lib.ntt_forward(ctx, None, n) == -1
    assert lib.ntt_forward(ctx, dummy.ctypes.data_as(capi.u64p), n // 2) == -1   # ntt.cpp:81 n != degree
    assert lib.ntt_inverse(ctx, dummy.ctypes.data_as(capi.u64p), n * 2) == -1
    lib.ntt_context_free(None)
    lib.ntt_context_free(ctx)


def test_lambda_snark_sys_smoke(gpu):
    """rust-api/lambda-snark-sys/src/lib.rs:28-43."""
    lib = capi.load()
    assert not lib.lwe_context_create(None)
    ctx = lib.ntt_context_create(12289, 256)
    if ctx:
        lib.ntt_context_free(ctx)


# ------------------------------------------------------------------- KATs
def test_survey_kats_on_device(gpu):
    g = json.loads((GOLD / "ntt_kat.json").read_text())
    for c in g["cases"]:
        ctx = api.NttContext(c["q"], c["n"])
        assert ctx.root == c["psi"]
        if "fwd_1to8" in c:
            x = np.zeros(c["n"], dtype=np.uint64); x[:8] = np.arange(1, 9)
            y = ctx.forward(x)
            assert [int(v) for v in y[:4]] == c["fwd_1to8"] and int(y[-1]) == c["fwd_1to8_last"]
            assert [int(v) for v in ctx.forward(np.ones(c["n"], dtype=np.uint64))[:3]] == c["fwd_ones"]
        ctx.close()
    for f in g["full"]:
        ctx = api.NttContext(f["q"], f["n"])
        x = np.array([int(v) for v in f["input"]], dtype=np.uint64)
        assert [str(int(v)) for v in ctx.forward(x)] == f["forward"]
        ctx.close()


def test_seal_unit_test_constants_on_device(gpu):
    """SEAL's own unit-test vectors (native/tests/seal/util/ntt.cpp NTTPrimitiveRootsTest / NegacyclicNTTTest, restated in
    tests/golden/ntt_kat.json) through ntt_context_create / ntt_forward / ntt_inverse -- what cpp-core/src/ntt.cpp:46,84,99
    would return with SEAL behind it.  q = 0xffffffffffc0001 is a 60-bit modulus: the guarded u64 butterflies."""
    g = json.loads((GOLD / "ntt_kat.json").read_text())["seal_unit_tests"]
    q = g["q"]
    for n, powers in g["root_powers"].items():
        ctx = api.NttContext(q, int(n))
        assert ctx.root == (powers[2] if int(n) == 4 else powers[1])
        # forward(X^j)[i] = psi^((2 brv(i) + 1) j): with j = 1 the outputs are the odd powers, and output 0 is psi itself
        x = np.zeros(int(n), dtype=np.uint64); x[1] = 1
        y = ctx.forward(x)
        assert int(y[0]) == ctx.root and int(y[1]) == q - ctx.root
        ctx.close()
    ctx = api.NttContext(q, 2)
    for v in g["forward_n2"]:
        y = ctx.forward(np.array(v["in"], dtype=np.uint64))
        assert [int(t) for t in y] == v["out"]
        assert [int(t) for t in ctx.inverse(y)] == v["in"]
    ctx.close()


# ------------------------------------------------------------ parity sweeps
SWEEP = [(12289, 256), (257, 2), (Q0, 2), (Q0, 4), (Q0, 8), (Q0, 16), (Q0, 32), (Q0, 64), (Q0, 128), (Q0, 512),
         (Q0, 1024), (Q0, 2048), (Q0, 4096), (Q1, 8192), (Q1, 16384), (Q1, 32768), (Q1, 65536), (Q1, 131072),
         (Q60, 16), (Q60, 4096), (Q60, 16384), (Q60, 65536), (Q50, 4096), (Q50, 32768), (Q31, 1024), (Q31, 65536),
         (Q45, 4096), (Q45, 8192), (Q45, 65536)]


@pytest.mark.parametrize("q,n", SWEEP)
def test_forward_inverse_match_oracle(gpu, rng, q, n):
    g = api.NttContext(q, n)
    o = O.OracleNtt(q, n)
    assert g.root == o.psi
    batch = max(1, min(37, (1 << 17) // n))          # odd counts: ragged last tile for n < 4096
    x = uniform(rng, q, (batch, n))
    x[0, :] = 0                                       # adversarial rows
    if batch > 1:
        x[1, :] = q - 1
    if batch > 2:
        x[2, :] = 0; x[2, 0] = 1
    if batch > 3:
        x[3, :] = 0; x[3, n - 1] = 1
    y = uniform(rng, q, (batch, n))
    want_f, want_i = o.forward(x), o.inverse(y)
    # both arithmetic policies where both exist: FP64 butterflies (auto for q < 2^45) and u64 Shoup
    for arith in ((0, 1) if q < 2**45 else (0,)):
        g.set_arith(arith)
        assert g.arith == (2 if (arith == 0 and q < 2**45) else 1)
        fx = g.forward_batch(x)
        assert np.array_equal(fx, want_f), f"arith {arith}"
        assert np.array_equal(g.inverse_batch(fx), x), f"arith {arith}"
        assert np.array_equal(g.inverse_batch(y), want_i), f"arith {arith}"
    g.close()


@pytest.mark.parametrize("q,n", [(Q0, 4096), (Q0, 1024), (Q0, 16), (Q1, 8192), (Q1, 131072), (Q31, 1024), (12289, 256),
                                 (Q45, 16), (Q45, 4096), (Q45, 16384), (Q45, 131072)])
def test_fp64_butterflies_on_extreme_inputs(gpu, rng, q, n):
    """The FP64 policy keeps balanced representatives whose bounds grow with the stage count;
    these rows push every bound: all q-1, alternating 0 / q-1 in every period, values around q/2,
    lazy inputs up to 4q-1 (forward) and 2q-1 (inverse), and rows that make one butterfly output
    sum coherently through all stages (the all-ones evaluation vector)."""
    g = api.NttContext(q, n)
    o = O.OracleNtt(q, n)
    assert g.arith == 2
    rows = [np.full(n, q - 1, dtype=np.uint64), np.full(n, q // 2, dtype=np.uint64), np.full(n, q // 2 + 1, dtype=np.uint64),
            np.ones(n, dtype=np.uint64)]
    for period in (1, 2, 4, 16, 256):
        if period < n:
            r = np.zeros(n, dtype=np.uint64); r[(np.arange(n) // period) % 2 == 1] = q - 1
            rows.append(r)
    for pos in (0, 1, n // 2, n - 1):
        r = np.zeros(n, dtype=np.uint64); r[pos] = q - 1
        rows.append(r)
    rows += [uniform(rng, q, n) for _ in range(4)]
    x = np.stack(rows)
    want_f, want_i = o.forward(x), o.inverse(x)
    assert np.array_equal(g.forward_batch(x), want_f)
    assert np.array_equal(g.inverse_batch(x), want_i)
    assert np.array_equal(g.forward_batch(x + np.uint64(3 * q)), want_f)        # lazy inputs, still below 4q
    assert np.array_equal(g.inverse_batch(x + np.uint64(q)), want_i)            # below 2q
    assert np.array_equal(g.inverse_batch(want_f), x)
    g.set_arith(1)
    assert np.array_equal(g.forward_batch(x), want_f) and np.array_equal(g.inverse_batch(x), want_i)
    g.close()


def test_arith_policy_selection(gpu):
    lib = capi.load()
    for q, n, pol in ((Q0, 4096, 2), (Q31, 1024, 2), (Q50, 4096, 1), (Q60, 4096, 1)):
        h = lib.ntt_context_create(q, n)
        assert lib.lsr_ntt_arith(h) == pol
        assert lib.lsr_ntt_set_arith(h, 2) == (0 if pol == 2 else -1)
        assert lib.lsr_ntt_set_arith(h, 1) == 0 and lib.lsr_ntt_arith(h) == 1
        assert lib.lsr_ntt_set_arith(h, 0) == 0 and lib.lsr_ntt_arith(h) == pol
        assert lib.lsr_ntt_set_arith(h, 3) == -1 and lib.lsr_ntt_set_arith(None, 0) == -1
        lib.ntt_context_free(h)


@pytest.mark.parametrize("q,n", [(Q0, 4096), (Q60, 4096), (Q1, 32768), (12289, 256)])
def test_single_call_entry_points(gpu, rng, q, n):
    g = api.NttContext(q, n)
    o = O.OracleNtt(q, n)
    x = uniform(rng, q, n)
    assert np.array_equal(g.forward(x), o.forward(x))
    assert np.array_equal(g.inverse(x), o.inverse(x))
    g.close()


@pytest.mark.parametrize("q", [Q0, Q60, 12289])
def test_lazy_and_wild_inputs(gpu, rng, q):
    """SEAL accepts [0,4q) inputs; anything else is defined here as NTT(x mod q)."""
    n = 1024 if q != 12289 else 256
    g = api.NttContext(q, n)
    o = O.OracleNtt(q, n)
    x = uniform(rng, q, (4, n))
    lazy = x + np.uint64(q) * rng.integers(0, 4, (4, n), dtype=np.uint64)
    assert np.array_equal(g.forward_batch(lazy), o.forward(x))
    wild = rng.integers(0, 2**64, (4, n), dtype=np.uint64)
    reduced = (wild % np.uint64(q)).astype(np.uint64)
    assert np.array_equal(g.forward_batch(wild), o.forward(reduced))
    assert np.array_equal(g.inverse_batch(wild), o.inverse(reduced))
    g.close()


def test_negacyclic_convolution_on_device(gpu, rng):
    for q, n in ((Q0, 4096), (Q0, 1024), (12289, 256)):
        g = api.NttContext(q, n)
        a = np.zeros(n, dtype=np.uint64); a[n - 1] = 1
        b = np.zeros(n, dtype=np.uint64); b[1] = 1
        r = g.inverse(g.mul_pointwise(g.forward(a), g.forward(b)))
        assert int(r[0]) == q - 1 and not r[1:].any()          # X^(n-1) * X = -1
        g.close()
    q, n = Q0, 64
    g = api.NttContext(q, n)
    a, b = uniform(rng, q, n), uniform(rng, q, n)
    got = g.inverse(g.mul_pointwise(g.forward(a), g.forward(b)))
    assert [int(v) for v in got] == O.py_negacyclic_mul(a, b, q)
    g.close()


@pytest.mark.parametrize("q", [Q0, Q60, 12289, Q31])
def test_pointwise_exact_for_any_u64(gpu, rng, q):
    g = api.NttContext(q, 16)
    o = O.OracleNtt(q, 16)
    for total in (1, 2, 3, 255, 4097, 100003):
        a = rng.integers(0, 2**64, total, dtype=np.uint64)
        b = rng.integers(0, 2**64, total, dtype=np.uint64)
        a[0], b[0] = 2**64 - 1, 2**64 - 1
        assert np.array_equal(g.mul_pointwise_batch(a, b), o.mul_pointwise(a, b))
    # no n check and aliasing allowed (ntt.cpp:106-119)
    lib = capi.load()
    a = rng.integers(0, q, 40, dtype=np.uint64); b = rng.integers(0, q, 40, dtype=np.uint64)
    want = o.mul_pointwise(a, b)
    lib.ntt_mul_pointwise(g.handle, a.ctypes.data_as(capi.u64p), a.ctypes.data_as(capi.u64p), b.ctypes.data_as(capi.u64p), 40)
    assert np.array_equal(a, want)
    keep = b.copy()
    lib.ntt_mul_pointwise(g.handle, None, a.ctypes.data_as(capi.u64p), b.ctypes.data_as(capi.u64p), 40)   # silent
    assert np.array_equal(b, keep)
    g.close()


def test_context_create_rejections_on_device(gpu):
    lib = capi.load()
    for q, n in ((12289, 0), (12289, 3), (12289, 1), (12289, 1 << 18), (12289, 4096), (Q0, 8192), (1 << 61, 16),
                 (17592186044417, 4096), (1, 16), (0, 16)):
        assert not lib.ntt_context_create(q, n), (q, n)
    h = lib.ntt_context_create(Q0, 4096)
    assert h and lib.lsr_ntt_modulus(h) == Q0 and lib.lsr_ntt_degree(h) == 4096 and lib.lsr_ntt_device(h) >= 0
    lib.ntt_context_free(h)


def test_full_size_properties(gpu, rng):
    """BASELINE batch sizes, checked through size-independent properties: round trip,
    linearity, and agreement with the oracle on a sampled subset."""
    q, n, batch = Q0, 4096, 8192
    g = api.NttContext(q, n)
    o = O.OracleNtt(q, n)
    x = uniform(rng, q, (batch, n))
    y = uniform(rng, q, (batch, n))
    fx = g.forward_batch(x)
    fy = g.forward_batch(y)
    s = ((x.astype(object) + y.astype(object)) % q).astype(np.uint64)
    fs = g.forward_batch(s)
    assert np.array_equal(fs, ((fx.astype(object) + fy.astype(object)) % q).astype(np.uint64))   # linearity
    assert np.array_equal(g.inverse_batch(fx), x)                                                 # round trip
    pick = rng.choice(batch, 64, replace=False)
    assert np.array_equal(fx[pick], o.forward(x[pick]))
    g.close()


def test_device_pointer_entry_points(gpu, rng):
    import torch
    q, n, batch = Q0, 4096, 33
    g = api.NttContext(q, n)
    o = O.OracleNtt(q, n)
    x = uniform(rng, q, (batch, n))
    d = torch.from_numpy(x.view(np.int64)).cuda()
    s = torch.cuda.current_stream().cuda_stream
    g.forward_device(d.data_ptr(), batch, s)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint64), o.forward(x))
    g.inverse_device(d.data_ptr(), batch, s)
    torch.cuda.synchronize()
    assert np.array_equal(d.cpu().numpy().view(np.uint64), x)
    a = torch.from_numpy(rng.integers(0, 2**63, batch * n, dtype=np.int64)).cuda()
    b = torch.from_numpy(rng.integers(0, 2**63, batch * n, dtype=np.int64)).cuda()
    r = torch.empty_like(a)
    g.mul_pointwise_device(r.data_ptr(), a.data_ptr(), b.data_ptr(), batch * n, s)
    torch.cuda.synchronize()
    assert np.array_equal(r.cpu().numpy().view(np.uint64),
                          o.mul_pointwise(a.cpu().numpy().view(np.uint64), b.cpu().numpy().view(np.uint64)))
    g.close()


def test_concurrent_contexts_on_threads(gpu, rng):
    """Rust marks the handles Send (context.rs:76); cargo test drives distinct contexts from parallel threads."""
    import threading
    errors = []
    data = uniform(rng, Q0, (8, 16, 1024))

    def work(i):
        try:
            g = api.NttContext(Q0, 1024)
            o = O.OracleNtt(Q0, 1024)
            for _ in range(5):
                assert np.array_equal(g.forward_batch(data[i]), o.forward(data[i]))
            g.close()
        except Exception as e:       # noqa: BLE001
            errors.append(e)

    ts = [threading.Thread(target=work, args=(i,)) for i in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errors, errors
