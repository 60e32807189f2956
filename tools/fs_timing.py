"""Timing of the device Fiat-Shamir challenge kernel and of lsr_prove_r1cs_batch."""
import sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from lambda_snark_r_b200 import api, capi
import ctypes as C

P = 2**64 - 2**32 + 1
Q0 = 17592169062401
api.set_device(0)
lib = capi.load()
words = 8193
for count in (32, 1024, 16384):
    cont = torch.randint(0, 2**62, (count, words), dtype=torch.int64, device="cuda")
    pub = torch.randint(0, 2**62, (count, 2), dtype=torch.int64, device="cuda")
    ab = torch.empty((count, 2), dtype=torch.int64, device="cuda")
    hs = torch.empty((count, 8), dtype=torch.int64, device="cuda")
    def run():
        rc = lib.lsr_fs_challenge_batch_device(pub.data_ptr(), 2, cont.data_ptr(), words, count, P, 1, ab.data_ptr(), hs.data_ptr(), None)
        assert rc == 0
    run(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): run()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"fs_challenge (alpha+beta, {words} words): count={count}: {ms:.3f} ms, {count / ms * 1e3:.0f} statements/s, "
          f"{2 * count * words * 8 / ms / 1e6:.1f} GB/s hashed")
# whole proofs, m = 4096
import random
m = 4096
rng = random.Random(1)
cols, z, A, B, Cm = 3 * m + 1, [1] + [0] * (3 * m), [], [], []
for i in range(m):                      # gates z[3i+1] * z[3i+2] = z[3i+3]
    a, b = rng.randrange(P), rng.randrange(P)
    z[3 * i + 1], z[3 * i + 2], z[3 * i + 3] = a, b, (a * b) % P
    A.append((i, 3 * i + 1, 1)); B.append((i, 3 * i + 2, 1)); Cm.append((i, 3 * i + 3, 1))
r = api.R1CS(m, cols, A, B, Cm, P)
ctx = api.LweContext(api.Params(n=4096, k=2, q=Q0, sigma=3.19), seed32=bytes(range(32)))
for count in (64, 1024):
    W = np.tile(np.array(z, dtype=np.uint64), (count, 1))
    seeds = np.arange(1, count + 1, dtype=np.uint64)
    r.prove_batch(ctx, W, 2, seeds)
    t0 = time.perf_counter()
    out = r.prove_batch(ctx, W, 2, seeds)
    dt = time.perf_counter() - t0
    assert not out["status"].any()
    print(f"lsr_prove_r1cs_batch m={m}: count={count}: {dt * 1e3:.2f} ms wall (host pointers), {count / dt:.0f} proofs/s")
