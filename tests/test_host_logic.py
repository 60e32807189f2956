"""CPU tests (-m "not gpu"): host-side logic -- parameter validation mirrored from
lambda-snark-core, the NTL-free R1CS handle (host code inside the C-ABI library),
and the multi-GPU sharding rules exercised with a world_size-2 gloo group."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import Q0
from lambda_snark_r_b200 import api, capi, sharding

ROOT = Path(__file__).resolve().parents[1]
TEST_MODULUS = 17592186044417          # 2^44 + 1, used by every reference test vector


# ----------------------------------------------------------------- params
def test_params_validation_mirrors_lambda_snark_core():
    api.Params().validate()
    for bad in (dict(n=1000), dict(k=0), dict(q=1 << 20), dict(sigma=2.5)):
        with pytest.raises(api.LambdaSnarkError):
            api.Params(**bad).validate()
    pp = api.Params(n=4096, k=2, q=TEST_MODULUS, sigma=3.19).to_ffi()
    assert (pp.profile, pp.security_level, pp.modulus, pp.ring_degree, pp.module_rank) == (1, 128, TEST_MODULUS, 4096, 2)


def test_opening_helpers():
    # opening.rs:104-115 and polynomial.rs:97-113 on TV-1's witness
    op = api.generate_opening([1, 7, 13, 91], 12345, 0x1234, TEST_MODULUS)
    assert op.witness == [0x1234, 1, 7, 13, 91]
    assert op.evaluation == (1 + 7 * 12345 + 13 * 12345**2 + 91 * 12345**3) % TEST_MODULUS


# ------------------------------------------------------------------- R1CS
def _matrix(entries, rows, cols):
    arr = (capi.SparseEntry * max(len(entries), 1))(*[capi.SparseEntry(r, c, v % (1 << 64)) for r, c, v in entries])
    return capi.SparseMatrix(arr, len(entries), rows, cols), arr


def _r1cs(A, B, Cm, rows, cols, q=TEST_MODULUS):
    lib = capi.load()
    keep = []
    mats = []
    for e in (A, B, Cm):
        m, arr = _matrix(e, rows, cols)
        keep.append(arr); mats.append(m)
    h = C.c_void_p()
    rc = lib.lambda_snark_r1cs_create(C.byref(mats[0]), C.byref(mats[1]), C.byref(mats[2]), q, C.byref(h))
    return rc, h, keep


def _validate(h, witness):
    lib = capi.load()
    w = np.array(witness, dtype=np.uint64)
    wit = capi.R1CSWitness(w.ctypes.data_as(capi.u64p), w.size)
    ok = C.c_bool(False)
    rc = lib.lambda_snark_r1cs_validate_witness(h, C.byref(wit), C.byref(ok))
    return rc, bool(ok.value)


def test_r1cs_tv1_multiplication():
    # test-vectors/tv-1-multiplication/constraints.json; rust tests/test_vectors.rs:70-93,118-133
    lib = capi.load()
    rc, h, _ = _r1cs([(0, 1, 1)], [(0, 2, 1)], [(0, 3, 1)], 1, 4)
    assert rc == 0
    assert lib.lambda_snark_r1cs_num_constraints(h) == 1 and lib.lambda_snark_r1cs_num_variables(h) == 4
    assert _validate(h, [1, 7, 13, 91]) == (0, True)
    assert _validate(h, [1, 8, 13, 91]) == (0, False)
    lib.lambda_snark_r1cs_free(h)


def test_r1cs_tv2_plaquette_negative_entries():
    # tv-2: B row holds -1 entries, stored as u64 wrap (test_vectors.rs:63); r1cs.cpp:165-167
    lib = capi.load()
    rc, h, _ = _r1cs([(0, 0, 1)], [(0, 1, 1), (0, 2, 1), (0, 3, -1), (0, 4, -1)], [], 1, 5)
    assert rc == 0
    assert _validate(h, [1, 314, 628, 471, 471]) == (0, True)
    assert _validate(h, [1, 314, 628, 471, 472]) == (0, False)
    lib.lambda_snark_r1cs_free(h)


def test_r1cs_modular_wrap_and_linear_combination():
    # cpp-core/tests/test_r1cs.cpp:203-241 ((q-1)^2 = 1) and :248-268 ((a+2b)*c = d)
    lib = capi.load()
    rc, h, _ = _r1cs([(0, 1, 1)], [(0, 2, 1)], [(0, 3, 1)], 1, 4)
    assert _validate(h, [1, TEST_MODULUS - 1, TEST_MODULUS - 1, 1]) == (0, True)
    lib.lambda_snark_r1cs_free(h)
    rc, h, _ = _r1cs([(0, 1, 1), (0, 2, 2)], [(0, 3, 1)], [(0, 4, 1)], 1, 5)
    assert _validate(h, [1, 3, 5, 7, 91]) == (0, True)
    assert _validate(h, [1, 3, 5, 7, 90]) == (0, False)
    lib.lambda_snark_r1cs_free(h)


def test_r1cs_error_codes():
    lib = capi.load()
    h = C.c_void_p()
    assert lib.lambda_snark_r1cs_create(None, None, None, TEST_MODULUS, C.byref(h)) == 1        # NULL_PTR, ffi.cpp:34
    a, ka = _matrix([(0, 1, 1)], 1, 4)
    b, kb = _matrix([(0, 2, 1)], 2, 4)                                                         # row mismatch
    assert lib.lambda_snark_r1cs_create(C.byref(a), C.byref(b), C.byref(a), TEST_MODULUS, C.byref(h)) == 2
    rc, h, _ = _r1cs([(0, 1, 1)], [(0, 2, 1)], [(0, 3, 1)], 1, 4)
    assert _validate(h, [1, 7, 13])[0] == 2                # length mismatch -> INVALID_PARAMS (r1cs.cpp:100-105)
    assert _validate(h, [2, 7, 13, 91])[0] == 2            # witness[0] != 1 (r1cs.cpp:108-110)
    assert lib.lambda_snark_r1cs_validate_witness(None, None, None) == 1
    lib.lambda_snark_r1cs_free(h)
    lib.lambda_snark_r1cs_free(None)
    assert lib.lambda_snark_r1cs_num_constraints(None) == 0 and lib.lambda_snark_r1cs_num_variables(None) == 0
    rc, h, _ = _r1cs([(0, 9, 1)], [(0, 2, 1)], [(0, 3, 1)], 1, 4)     # column out of range -> CRYPTO_FAILED (ffi.cpp:73-75)
    assert _validate(h, [1, 7, 13, 91])[0] == 4
    lib.lambda_snark_r1cs_free(h)


# --------------------------------------------------------------- sharding
def test_shard_ranges_partition_the_batch():
    for count in (0, 1, 7, 256, 1000003):
        for world in (1, 2, 3, 8):
            rs = [sharding.shard_range(count, r, world) for r in range(world)]
            assert rs[0][0] == 0 and rs[-1][1] == count
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in rs]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_global_seeds_do_not_depend_on_world_size():
    full = sharding.global_seeds(0xC0FFEE, 0, 100)
    for world in (2, 3, 8):
        parts = [sharding.global_seeds(0xC0FFEE, *sharding.shard_range(100, r, world)) for r in range(world)]
        assert np.array_equal(np.concatenate(parts), full)
    assert (sharding.global_seeds(-5 % (1 << 64), 0, 10) != 0).all()


def _gloo_worker(rank, world, port, count, out_dir):
    # the N>1 path on CPU: gloo group, each rank commits ITS slice (the oracle stands in
    # for the device here -- the GPU twin of this test is tests/test_gpu_commit.py),
    # then the final all-gather; rank 0 saves the gathered batch
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = O.OracleLwe(Q0, 256, 2, 3.19, bytes(range(32)))
    start, stop = sharding.shard_range(count, rank, world)
    rng = np.random.Generator(np.random.PCG64(99))
    msgs = rng.integers(0, Q0, size=(count, 256), dtype=np.uint64)           # same synthetic batch on all ranks
    mine = ctx.commit_batch(msgs[start:stop], sharding.global_seeds(0xC0FFEE, start, stop))
    # all ranks own equal slices here (count % world == 0), as in the weak-scaling bench
    local = torch.from_numpy(mine.view(np.int64))
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    if rank == 0:
        full = sharding.gather_slices([g.numpy().view(np.uint64) for g in gathered])
        np.save(Path(out_dir) / "gathered.npy", full)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_gather_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    count, world = 8, 2
    port = 29500 + os.getpid() % 2000
    mp.spawn(_gloo_worker, args=(world, port, count, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "gathered.npy")
    ctx = O.OracleLwe(Q0, 256, 2, 3.19, bytes(range(32)))
    rng = np.random.Generator(np.random.PCG64(99))
    msgs = rng.integers(0, Q0, size=(count, 256), dtype=np.uint64)
    want = ctx.commit_batch(msgs, sharding.global_seeds(0xC0FFEE, 0, count))
    assert np.array_equal(got, want)


# ------------------------------------------------------------- Lean export
def test_lean_export_formats_match_the_reference():
    """cpp-core/src/lean_ffi.cpp:46-78,152-232 and rust-api/lambda-snark/src/lean_export.rs:119-199 (same terms)."""
    lib = capi.load()
    buf = C.create_string_buffer(4096)
    pp = capi.PublicParams(1, 128, 12289, 4096, 2, 3.2)
    n = lib.export_params_to_lean(C.byref(pp), buf, len(buf))
    text = "{ n := 4096, k := 2, q := 12289, σ := 3.2, λ := 128 }"            # lean_ffi.cpp:66 doc comment
    assert n == len(text.encode()) and buf.value.decode() == text
    pp2 = capi.PublicParams(1, 128, 17592169062401, 4096, 2, 3.19)
    lib.export_params_to_lean(C.byref(pp2), buf, len(buf))
    assert buf.value.decode() == "{ n := 4096, k := 2, q := 17592169062401, σ := 3.2, λ := 128 }"   # setprecision(1)
    assert lib.export_params_to_lean(C.byref(pp), buf, 10) == -1                 # buffer too small
    assert lib.export_params_to_lean(None, buf, len(buf)) == -1

    # TV-1 (7 * 13 = 91): one constraint, four variables, two public inputs (lean_export.rs:263-280 header check)
    def mat(entries, rows, cols):
        arr = (capi.SparseEntry * len(entries))(*[capi.SparseEntry(r, c, v) for r, c, v in entries])
        return capi.SparseMatrix(arr, len(entries), rows, cols), arr
    (A, ka), (B, kb), (Cm, kc) = mat([(0, 1, 1)], 1, 4), mat([(0, 2, 1)], 1, 4), mat([(0, 3, 1), (0, 0, 5)], 1, 4)
    sys_ = capi.R1CSConstraintSystem(A, B, Cm, 4, 2, 1)
    pp3 = capi.PublicParams(1, 128, TEST_MODULUS, 4096, 2, 3.19)
    n = lib.export_vk_to_lean(C.byref(sys_), C.byref(pp3), buf, len(buf))
    want = ("⟨1, 4, 2, 17592186044417, SparseMatrix.mk 1 4 [(0, 1, 1)], SparseMatrix.mk 1 4 [(0, 2, 1)], "
            "SparseMatrix.mk 1 4 [(0, 3, 1), (0, 0, 5)]⟩")
    assert buf.value.decode() == want and n == len(want.encode())
    assert lib.export_vk_to_lean(C.byref(sys_), C.byref(pp3), buf, 16) == -1
    sys_.n_public_inputs = 5                                                     # lean_ffi.cpp:161-167
    assert lib.export_vk_to_lean(C.byref(sys_), C.byref(pp3), buf, len(buf)) == -1
    # SEAL-specific exporters: no SEAL object in this library, -1 like the reference without a SEAL context
    assert lib.export_seal_context_to_lean(None, buf, len(buf)) == -1
    assert lib.export_seal_pubkey_to_lean(None, buf, len(buf)) == -1


def _gloo_prover_worker(rank, world, port, out_dir):
    # BASELINE configs[4] on CPU: every rank forms the quotient of the same witness and commits ITS chunk range
    # (seeds indexed by the global chunk number); the gathered containers must be the single-process result.
    # The oracle stands in for the device; the GPU twin is test_prover_commit_phase_matches_oracle_and_is_shard_invariant.
    import random
    import torch
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    from oracle import oracle as O
    from oracle import quotient as QO
    from test_oracle_quotient import mult_gates
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    q, m, n, k = QO.NTT_MODULUS, 128, 16, 2
    cols, A, B, Cm, z = mult_gates(m, q, random.Random(5))
    quo = QO.compute_quotient_poly(m, A, B, Cm, z, q)
    quo = quo + [0] * (m - len(quo))
    ctx = O.OracleLwe(Q0, n, k, 3.19, bytes(range(32)))
    # units = (ring element, base-p digit plane) pairs, the order lsr_prover_commit_quotient commits them in
    planes = sharding.message_planes(ctx.p, q)
    chunks = m // n * planes
    lo, hi = sharding.shard_range(chunks, rank, world)
    msgs = sharding.message_digits(np.array(quo, dtype=np.uint64).reshape(m // n, n), ctx.p, planes)
    mine = ctx.commit_batch(msgs[lo:hi], sharding.global_seeds(0xC0FFEE, lo, hi))
    local = torch.from_numpy(mine.view(np.int64))
    gathered = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    if rank == 0:
        np.save(Path(out_dir) / "prover_gathered.npy", sharding.gather_slices([g.numpy().view(np.uint64) for g in gathered]))
        np.save(Path(out_dir) / "prover_msgs.npy", msgs)
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_prover_chunk_ranges(tmp_path):
    import torch.multiprocessing as mp
    from oracle import oracle as O
    world = 2
    port = 31500 + os.getpid() % 2000
    mp.spawn(_gloo_prover_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    got = np.load(tmp_path / "prover_gathered.npy")
    msgs = np.load(tmp_path / "prover_msgs.npy")
    ctx = O.OracleLwe(Q0, 16, 2, 3.19, bytes(range(32)))
    want = ctx.commit_batch(msgs, sharding.global_seeds(0xC0FFEE, 0, msgs.shape[0]))
    assert got.shape == want.shape and np.array_equal(got, want)


# ------------------------------------------------------------- Fiat-Shamir transcript hash
def test_transcript_hash_source_matches_sha3_256(tmp_path):
    """csrc/lsr_keccak.h (the code fs_challenge_kernel runs, compiled here for the host) against hashlib on
    every alignment case of the 136-byte rate: challenge.rs:102-134's transcript."""
    import hashlib
    import struct
    import subprocess
    so = tmp_path / "libkeccak_host.so"
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-I", str(ROOT / "lambda_snark_r_b200" / "csrc"), "-o", str(so),
                    str(ROOT / "tests" / "cabi" / "keccak_host.cpp")], check=True)
    lib = C.CDLL(str(so))
    rng = np.random.default_rng(1)
    for n_pub in (0, 1, 2, 5, 16, 17, 40):
        for n_words in (0, 1, 11, 12, 13, 14, 15, 16, 17, 18, 29, 30, 31, 33, 34, 35, 100, 8193):
            pub = rng.integers(0, 2**64, n_pub, dtype=np.uint64)
            words = rng.integers(0, 2**64, n_words, dtype=np.uint64)
            out = np.zeros(4, dtype=np.uint64)
            lib.lsr_test_fs_hash(pub.ctypes.data_as(capi.u64p), C.c_uint64(n_pub), words.ctypes.data_as(capi.u64p),
                                 C.c_uint64(n_words), out.ctypes.data_as(capi.u64p))
            h = hashlib.sha3_256(b"LAMBDA-SNARK-R-FS-v1" + struct.pack("<Q", n_pub) + pub.tobytes() +
                                 struct.pack("<Q", n_words) + words.tobytes())
            assert out.tobytes() == h.digest(), (n_pub, n_words)
            # the lane-parallel formulation of the warp-per-statement kernel (same index maps, arrays for shuffles)
            out2 = np.zeros(4, dtype=np.uint64)
            lib.lsr_test_fs_hash_lanes(pub.ctypes.data_as(capi.u64p), C.c_uint64(n_pub), words.ctypes.data_as(capi.u64p),
                                       C.c_uint64(n_words), out2.ctypes.data_as(capi.u64p))
            assert out2.tobytes() == h.digest(), ("lanes", n_pub, n_words)


def test_copy_pool_stress(tmp_path):
    """csrc/lsr_copy_pool.h: parallel memcpy used by the staged pageable path of lwe_commit_batch."""
    import subprocess
    exe = tmp_path / "copy_pool_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-I", str(ROOT / "lambda_snark_r_b200" / "csrc"), "-o", str(exe),
                    str(ROOT / "tests" / "cabi" / "copy_pool_test.cpp")], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "copy pool ok" in r.stderr, r.stderr
