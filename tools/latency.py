"""Single-call latency of the drop-in entry points (what a caller that commits one polynomial at a time sees)."""
import sys, time
from pathlib import Path
import numpy as np
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from lambda_snark_r_b200 import api

Q, N, K = 17592169062401, 4096, 2
api.set_device(0)
ctx = api.LweContext(api.Params(n=N, k=K, q=Q, sigma=3.19), seed32=bytes(range(32)))
ntt = api.NttContext(Q, N)
rng = np.random.default_rng(1)
msg = rng.integers(0, Q, size=N, dtype=np.uint64)


def bench(fn, reps=200):
    for _ in range(20): fn()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    return (time.perf_counter() - t0) / reps * 1e6

from lambda_snark_r_b200 import capi
lib = capi.load()
mp = msg.ctypes.data_as(capi.u64p)


def one():
    c = lib.lwe_commit(ctx.as_ptr(), mp, N, 7)
    lib.lwe_commitment_free(c)

print(f"lwe_commit + lwe_commitment_free (raw C ABI, 1 commitment, pageable host memory): {bench(one):8.1f} us per call")
for path, name in ((1, 'generic'), (2, 'fused')):
    ctx.set_commit_path(path)
    print(f"  forced {name:7s} path:                      {bench(one):8.1f} us per call")
ctx.set_commit_path(0)
for b in (1, 4, 16, 64):
    m = np.tile(msg, (b, 1)); s = np.arange(1, b + 1, dtype=np.uint64)
    us = bench(lambda: ctx.commit_batch(m, s), 50)
    print(f"lwe_commit_batch count={b:4d}: {us:8.1f} us per call, {us / b:7.1f} us per commitment")
for b in (128, 256, 1024, 4096):
    m = rng.integers(0, Q, size=(b, N), dtype=np.uint64); s = np.arange(1, b + 1, dtype=np.uint64)
    o = np.zeros((b, ctx.words), dtype=np.uint64)
    us = bench(lambda: ctx.commit_batch(m, s, out=o), 5)
    print(f"lwe_commit_batch count={b:4d} (pageable numpy memory, staged path): {us / 1e3:8.2f} ms per call, {b / us:6.3f} M commitments/s")
x = msg.copy()
print(f"ntt_forward (1 polynomial, host pointers): {bench(lambda: ntt.forward(x)):8.1f} us per call")
c = api.Commitment.new(ctx, msg % np.uint64(ctx.p), 9)
print(f"lwe_verify_opening: {bench(lambda: api.verify_commitment(ctx, c, msg % np.uint64(ctx.p)), 100):8.1f} us per call")
