"""Regenerates tests/golden/*.json.

Sources of truth, in order of strength:
  1. sampler_ref.json   -- produced by the REFERENCE's own code: cpp-core/src/utils.cpp
                           compiled in place into oracle/_ref (oracle/ref_sampler_shim.cpp).
                           Needs /root/reference; the resulting vectors travel with the repo.
  2. ntt_kat.json       -- the KATs SURVEY.md 8c derived from the SEAL 4.1 specification
                           (psi_min and forward-NTT samples), copied verbatim, plus
                           reference-test inputs ([1..8,0...] round trip, 2*3=6).
  3. commit_kat.json    -- digests of oracle commitments for fixed (context seed, seed,
                           message): pins the oracle against silent drift; the reference's
                           own commitments are randomised and cannot be pinned.

Run from the repo root:  python tests/golden/make_golden.py
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import oracle as O  # noqa: E402

HERE = Path(__file__).resolve().parent
Q0 = 17592169062401


def sampler_ref():
    ref = O.RefSampler()
    out = {"source": "reference cpp-core/src/utils.cpp build_cdf + sample_single via oracle/_ref", "tables": {}, "samples": {}}
    for sigma in (3.19, 3.2, 1.0, 0.5, 10.0):
        out["tables"][repr(sigma)] = [str(int(v)) for v in ref.build_cdf(sigma)]
    rng = np.random.Generator(np.random.PCG64(2024))
    for sigma in (3.19, 3.2):
        cdf = ref.build_cdf(sigma)
        draws = rng.integers(0, 2**64, 512, dtype=np.uint64).tolist()
        for v in cdf[:24]:                      # boundary cases: u = cdf[k] - 1, cdf[k], cdf[k] + 1
            for d in (-1, 0, 1):
                draws += [(int(v) + d) % 2**64, 1, (int(v) + d) % 2**64, 0]
        draws = np.array(draws, dtype=np.uint64)
        samples = ref.sample(sigma, draws)
        out["samples"][repr(sigma)] = {"draws": [str(int(d)) for d in draws], "samples": [int(s) for s in samples]}
    (HERE / "sampler_ref.json").write_text(json.dumps(out, indent=0))


def ntt_kat():
    kat = {
        "source": "SURVEY.md section 8c (derived from the SEAL 4.1 spec; not from a SEAL binary)",
        "cases": [
            {"q": 12289, "n": 256, "psi": 3, "fwd_1to8": [26, 11046, 1743, 1098], "fwd_1to8_last": 11454,
             "fwd_ones": [12288, 6145, 8261]},
            {"q": Q0, "n": 1024, "psi": 60934826393,
             "fwd_1to8": [8085186849839, 16063790770996, 11396965722667, 14084617137215],
             "fwd_1to8_last": 1423427964781, "fwd_ones": [4619811266306, 10650564614633, 12662679278915]},
            {"q": Q0, "n": 2048, "psi": 11696237686},
            {"q": Q0, "n": 4096, "psi": 1299579534,
             "fwd_1to8": [4906668228709, 13352284639367, 1243151528753, 7252051638976],
             "fwd_1to8_last": 17473155690403, "fwd_ones": [10796465977081, 15311896912402, 8589703109061]},
        ],
        # Constants asserted by SEAL's OWN unit tests (Microsoft SEAL 4.1, native/tests/seal/util/ntt.cpp:
        # NTTTablesTest.NTTPrimitiveRootsTest and NTTTablesTest.NegacyclicNTTTest), restated here from that file -- SEAL
        # is the library cpp-core/src/ntt.cpp:46,84 calls and is not vendored under /root/reference.  They were NOT
        # produced by this repo's code: the oracle and the device are checked against them.  They pin the choice of
        # the minimal primitive 2n-th root, the bit-reversed table order and the output order of the forward transform.
        "seal_unit_tests": {
            "source": "SEAL 4.1 native/tests/seal/util/ntt.cpp (NTTPrimitiveRootsTest, NegacyclicNTTTest)",
            "q": 0xffffffffffc0001,
            "root_powers": {"2": [1, 288794978602139552],
                            "4": [1, 288794978602139552, 178930308976060547, 748001537669050592]},
            "forward_n2": [{"in": [0, 0], "out": [0, 0]}, {"in": [1, 0], "out": [1, 1]},
                           {"in": [1, 1], "out": [288794978602139553, 864126526004445282]}]},
        "roots_of_unity_rs": {      # rust-api/lambda-snark/src/r1cs.rs:534-547, omega_m = 3^((q-1)/m)
            "q": Q0, "generator": 3, "m": [4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192]},
    }
    # full forward vectors (oracle) for a seeded input, so the GPU path can be checked offline too
    rng = np.random.Generator(np.random.PCG64(0x5EED))
    for q, n in ((12289, 256), (Q0, 1024)):
        x = rng.integers(0, q, n, dtype=np.uint64)
        y = O.OracleNtt(q, n).forward(x)
        kat.setdefault("full", []).append({"q": q, "n": n, "input_sha256": hashlib.sha256(x.tobytes()).hexdigest(),
                                           "seed": "PCG64(0x5EED) sequential", "input": [str(int(v)) for v in x],
                                           "forward": [str(int(v)) for v in y]})
    (HERE / "ntt_kat.json").write_text(json.dumps(kat, indent=0))


def explicit_inputs(q, n, k, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    s = rng.integers(-(q - 1), q, size=(k, n), dtype=np.int64)
    e = rng.integers(-(q - 1), q, size=(k, n), dtype=np.int64)
    msg = rng.integers(0, 2**64, size=n, dtype=np.uint64)
    return s, e, msg


def commit_kat():
    cases = []
    for (n, k, sigma, seed, msg) in [(4096, 2, 3.19, 0x1234, [1, 2, 3, 4]), (4096, 2, 3.19, 0xC0FFEE, list(range(100))),
                                     (1024, 2, 3.2, 7, [7, 11, 13, 17]), (256, 3, 3.19, 99, [5])]:
        ctx = O.OracleLwe(Q0, n, k, sigma, bytes(range(32)))
        c = ctx.commit(msg, seed)
        cases.append({"n": n, "k": k, "sigma": sigma, "ctx_seed": "bytes(range(32))", "seed": seed, "msg": msg,
                      "q": ctx.q, "p": ctx.p, "delta": ctx.delta, "words": int(c.size), "first": [str(int(v)) for v in c[:5]],
                      "sha256": hashlib.sha256(c.tobytes()).hexdigest(),
                      "matrix_sha256": hashlib.sha256(ctx.matrix().tobytes()).hexdigest()})
    # explicit mode (s, e supplied): s, e uniform over (-q, q) and the message uniform over u64, all from PCG64(seed)
    explicit = []
    for (n, k, seed) in [(4096, 2, 1), (1024, 2, 2), (64, 3, 3)]:
        ctx = O.OracleLwe(Q0, n, k, 3.19, bytes(range(32)))
        s, e, msg = explicit_inputs(ctx.q, n, k, seed)
        c = ctx.commit_explicit(msg, s, e)
        explicit.append({"n": n, "k": k, "sigma": 3.19, "ctx_seed": "bytes(range(32))", "pcg64_seed": seed,
                         "inputs": "tests/golden/make_golden.py::explicit_inputs", "first": [str(int(v)) for v in c[:5]],
                         "sha256": hashlib.sha256(c.tobytes()).hexdigest()})
    (HERE / "commit_kat.json").write_text(json.dumps({"source": "oracle/lsr_oracle.c (self-pin)", "cases": cases,
                                                      "explicit": explicit}, indent=0))


if __name__ == "__main__":
    sampler_ref()
    ntt_kat()
    commit_kat()
    print("golden vectors written to", HERE)
