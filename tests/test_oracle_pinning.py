"""CPU tests (-m "not gpu"): pin the ORACLE before trusting it.

 * sampler: against vectors produced by the reference's own utils.cpp
   (tests/golden/sampler_ref.json) and, when oracle/_ref is present, live.
 * NTT: against the SURVEY 8c KATs, the closed form out[i] = f(psi^(2 brv(i)+1)),
   schoolbook negacyclic products, and the reference's own test assertions
   (cpp-core/tests/test_ntt.cpp:47-81).
 * root table: rust-api/lambda-snark/src/r1cs.rs:534-547.
"""
import json
from pathlib import Path

import numpy as np
import pytest

from conftest import Q0, Q1, Q31, Q50, Q60, uniform
from oracle import oracle as O

GOLD = Path(__file__).resolve().parent / "golden"


# ------------------------------------------------------------------ sampler
def test_cdt_tables_match_reference_golden():
    g = json.loads((GOLD / "sampler_ref.json").read_text())
    for sigma_s, table in g["tables"].items():
        mine = O.cdt_build(float(sigma_s))
        assert [str(int(v)) for v in mine] == table, sigma_s


def test_cdt_kats_from_survey():
    t = O.cdt_build(3.19)
    assert len(t) == 40
    assert [int(t[i]) for i in (0, 1, 2, 10, 20, 38)] == [
        2306954904936458487, 6699640861611186643, 10490287658698886018,
        18429224519687107125, 18446744071673367447, 18446744073709551614]
    assert int(t[39]) == 2**64 - 1
    t = O.cdt_build(3.2)
    assert [int(t[i]) for i in (0, 1, 2, 10, 20)] == [
        2299745670858531888, 6680047332305665407, 10463485708449051034,
        18428562412348008226, 18446744071381690596]
    assert int(t[38]) == int(t[39]) == 2**64 - 1


def test_sample_mapping_matches_reference_golden():
    g = json.loads((GOLD / "sampler_ref.json").read_text())
    for sigma_s, rec in g["samples"].items():
        cdf = O.cdt_build(float(sigma_s))
        draws = [int(d) for d in rec["draws"]]
        got = [O.cdt_sample(cdf, draws[2 * i], draws[2 * i + 1]) for i in range(len(draws) // 2)]
        assert got == rec["samples"], sigma_s


def test_sampler_live_against_reference_build(rng):
    try:
        ref = O.RefSampler()
    except (FileNotFoundError, OSError, RuntimeError):
        pytest.skip("oracle/_ref not built and /root/reference absent")
    for sigma in (3.19, 3.2, 0.7, 6.5, 12.0):
        assert np.array_equal(ref.build_cdf(sigma), O.cdt_build(sigma))
    cdf = O.cdt_build(3.19)
    draws = rng.integers(0, 2**64, 4000, dtype=np.uint64)
    want = ref.sample(3.19, draws)
    got = np.array([O.cdt_sample(cdf, int(draws[2 * i]), int(draws[2 * i + 1])) for i in range(2000)])
    assert np.array_equal(got, want)


def test_sampler_rejects_bad_sigma():          # utils.cpp:133-135
    for bad in (0.0, -1.0, float("inf"), float("nan")):
        with pytest.raises(ValueError):
            O.cdt_build(bad)
        with pytest.raises(ValueError):
            O.sample_gaussian_seeded(4, bad, bytes(32))


def test_seeded_sampler_moments():             # cpp-core/tests/test_utils.cpp:35-70
    x = O.sample_gaussian_seeded(4096, 3.2, bytes(range(32))).astype(np.float64)
    assert abs(x.mean()) < 0.5
    assert abs(x.std(ddof=1) - 3.2) < 0.8
    pos, neg = int((x > 0).sum()), int((x < 0).sum())
    assert pos > 1024 and neg > 1024 and abs(pos - neg) < 4096 // 5


def test_chacha8_known_answer():
    # ChaCha8, zero key / counter / nonce: first keystream bytes 3e 00 ef 2f 89 5f 40 d6 ...
    out = O.chacha_block([0] * 8, 0, 0, 0, 0)
    assert out[:4].tobytes().hex() == "3e00ef2f895f40d67f5bb8e81f09a5a1"


# --------------------------------------------------------------------- NTT
def test_ntt_kats():
    g = json.loads((GOLD / "ntt_kat.json").read_text())
    for c in g["cases"]:
        ctx = O.OracleNtt(c["q"], c["n"])
        assert ctx.psi == c["psi"]
        if "fwd_1to8" in c:
            x = np.zeros(c["n"], dtype=np.uint64)
            x[:8] = np.arange(1, 9)
            y = ctx.forward(x)
            assert [int(v) for v in y[:4]] == c["fwd_1to8"]
            assert int(y[-1]) == c["fwd_1to8_last"]
            assert [int(v) for v in ctx.forward(np.ones(c["n"], dtype=np.uint64))[:3]] == c["fwd_ones"]
    for f in g["full"]:
        ctx = O.OracleNtt(f["q"], f["n"])
        x = np.array([int(v) for v in f["input"]], dtype=np.uint64)
        assert [str(int(v)) for v in ctx.forward(x)] == f["forward"]


def test_seal_unit_test_constants():
    """The constants SEAL's own unit tests assert (native/tests/seal/util/ntt.cpp, restated in ntt_kat.json): root-power
    tables for n = 2, 4 and the n = 2 forward transforms at q = 0xffffffffffc0001.  cpp-core/src/ntt.cpp:46,84 is SEAL."""
    g = json.loads((GOLD / "ntt_kat.json").read_text())["seal_unit_tests"]
    q = g["q"]
    for n, powers in g["root_powers"].items():
        ctx = O.OracleNtt(q, int(n))
        assert ctx.psi == powers[2 if int(n) == 4 else 1]            # root_powers[brv(1)] = psi for n = 2; [2] = psi^1 for n = 4
        assert [int(v) for v in ctx.table(0)] == powers
        if int(n) == 2:                                               # the same SEAL test: inv_root_powers(1) inverts root_powers(1)
            assert (int(ctx.table(2)[1]) * powers[1]) % q == 1
    ctx = O.OracleNtt(q, 2)
    for v in g["forward_n2"]:
        assert [int(x) for x in ctx.forward(np.array(v["in"], dtype=np.uint64))] == v["out"]


@pytest.mark.parametrize("q,n", [(12289, 256), (Q0, 64), (Q0, 512), (Q60, 128), (Q31, 32), (257, 2), (Q0, 2)])
def test_forward_is_closed_form(q, n, rng):
    ctx = O.OracleNtt(q, n)
    x = uniform(rng, q, n)
    assert [int(v) for v in ctx.forward(x)] == O.py_forward_closed_form(x, q, ctx.psi)


@pytest.mark.parametrize("q,n", [(12289, 256), (Q0, 1024), (Q0, 4096), (Q1, 8192), (Q60, 4096), (Q50, 2048)])
def test_roundtrip_and_negacyclic_wrap(q, n, rng):
    ctx = O.OracleNtt(q, n)
    x = uniform(rng, q, (3, n))
    assert np.array_equal(ctx.inverse(ctx.forward(x)), x)
    a = np.zeros(n, dtype=np.uint64); a[n - 1] = 1
    b = np.zeros(n, dtype=np.uint64); b[1] = 1
    r = ctx.inverse(ctx.mul_pointwise(ctx.forward(a), ctx.forward(b)))
    assert int(r[0]) == q - 1 and not r[1:].any()        # X^(n-1) * X = -1


def test_convolution_matches_schoolbook(rng):
    q, n = Q0, 64
    ctx = O.OracleNtt(q, n)
    a, b = uniform(rng, q, n), uniform(rng, q, n)
    got = ctx.inverse(ctx.mul_pointwise(ctx.forward(a), ctx.forward(b)))
    assert [int(v) for v in got] == O.py_negacyclic_mul(a, b, q)


def test_reference_test_ntt_assertions():
    # cpp-core/tests/test_ntt.cpp:13-14,47-81
    q, n = 12289, 256
    ctx = O.OracleNtt(q, n)
    x = np.zeros(n, dtype=np.uint64); x[:8] = np.arange(1, 9)
    assert np.array_equal(ctx.inverse(ctx.forward(x)), x)
    r = ctx.mul_pointwise(np.full(n, 2, dtype=np.uint64), np.full(n, 3, dtype=np.uint64))
    assert (r == 6).all()


def test_lazy_inputs_up_to_4q(rng):
    # SEAL's forward transform accepts inputs in [0, 4q): result is the NTT of x mod q
    q, n = Q0, 256
    ctx = O.OracleNtt(q, n)
    x = uniform(rng, q, n)
    lazy = x + np.uint64(q) * rng.integers(0, 4, n, dtype=np.uint64)
    assert np.array_equal(ctx.forward(lazy), ctx.forward(x))


def test_context_create_rejections():
    # ntt.cpp:31,41 and SEAL Modulus / NTTTables rules
    assert O.OracleNtt.try_create(12289, 0) is None
    assert O.OracleNtt.try_create(12289, 3) is None
    assert O.OracleNtt.try_create(12289, 1) is None
    assert O.OracleNtt.try_create(12289, 1 << 18) is None
    assert O.OracleNtt.try_create(12289, 4096) is None            # 8192 does not divide 12288
    assert O.OracleNtt.try_create(Q0, 8192) is None               # 2-adicity 13
    assert O.OracleNtt.try_create(1 << 61, 16) is None
    assert O.OracleNtt.try_create(17592186044417, 4096) is None   # 2^44+1 is composite
    assert O.OracleNtt.try_create(1, 16) is None
    assert O.OracleNtt.try_create(Q0, 4096) is not None


def test_pointwise_exact_for_any_u64(rng):
    ctx = O.OracleNtt(Q60, 16)
    a = rng.integers(0, 2**64, 1000, dtype=np.uint64)
    b = rng.integers(0, 2**64, 1000, dtype=np.uint64)
    got = ctx.mul_pointwise(a, b)
    assert [int(v) for v in got] == [(int(x) * int(y)) % Q60 for x, y in zip(a, b)]


def test_roots_of_unity_table_r1cs_rs():
    # rust-api/lambda-snark/src/r1cs.rs:534-547: omega_m = 3^((q-1)/m), exact order m
    table = {4: 981206394875, 8: 4268641988953, 16: 9400386778549, 32: 15690227524213, 64: 8332322609789,
             128: 9249819209096, 256: 5221410271124, 512: 9594533594163, 1024: 11016271016603,
             2048: 14373677444369, 4096: 11176258803537, 8192: 9037003627149}
    for m, w in table.items():
        assert pow(3, (Q0 - 1) // m, Q0) == w
        assert pow(w, m, Q0) == 1 and pow(w, m // 2, Q0) == Q0 - 1
    # the negacyclic psi for n is a primitive 2n-th root: psi^2 generates the same group as omega_n
    for n in (1024, 2048, 4096):
        psi = O.OracleNtt(Q0, n).psi
        assert pow(psi, n, Q0) == Q0 - 1
