"""Timing of the prover's commitment phase (BASELINE configs[4]): m constraints -> quotient -> m/n commitments.
Device-resident witnesses, CUDA events on the default stream (the pipeline's stream)."""
import sys, time
from pathlib import Path
import numpy as np
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from lambda_snark_r_b200 import capi
if "--lib" in sys.argv:          # kernel-variant experiments: a library built by `_build --variant` in place of the shipped one
    capi._lib = capi.load(Path(sys.argv.pop(sys.argv.index("--lib") + 1))); sys.argv.remove("--lib")
from lambda_snark_r_b200 import api

P = 2**64 - 2**32 + 1
Q0, N, K = 17592169062401, 4096, 2


def gates(m, q, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.integers(0, q, size=m, dtype=np.uint64)
    b = rng.integers(0, q, size=m, dtype=np.uint64)
    # a*b mod q without Python loops: object arrays are slow at 2^20, so use exact integer arithmetic in chunks
    c = np.fromiter(((int(x) * int(y)) % q for x, y in zip(a.tolist(), b.tolist())), dtype=np.uint64, count=m)
    z = np.zeros(3 * m + 1, dtype=np.uint64); z[0] = 1
    z[1::3], z[2::3], z[3::3] = a, b, c
    rows = np.arange(m, dtype=np.uint32); one = np.ones(m, dtype=np.uint64)
    return 3 * m + 1, (rows, 3 * rows + 1, one), (rows, 3 * rows + 2, one), (rows, 3 * rows + 3, one), z


def main():
    logm = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    W = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    m = 1 << logm
    api.set_device(0)
    cols, A, B, C, z = gates(m, P, 1)
    r = api.R1CS.from_arrays(m, cols, A, B, C, P)
    ctx = api.LweContext(api.Params(n=N, k=K, q=Q0, sigma=3.19), seed32=bytes(range(32)))
    chunks = r.quotient_chunks(ctx)
    zs = torch.from_numpy(np.tile(z.view(np.int64), (W, 1))).cuda()
    seeds = torch.arange(1, W * chunks + 1, dtype=torch.int64, device="cuda")
    out = torch.empty((W, chunks, ctx.words), dtype=torch.int64, device="cuda")

    def run(lo, hi):
        st = r.commit_quotient_device(ctx, zs.data_ptr(), W, seeds.data_ptr(), out.data_ptr(), lo, hi)
        assert st.tolist() == [0] * W

    def timed(lo, hi, reps=5):
        for _ in range(2): run(lo, hi)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): run(lo, hi)
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    tq = timed(0, 0)
    tf = timed(0, chunks)
    alg = 240 * m * W
    print(f"m=2^{logm} W={W} chunks={chunks}: quotient {tq:.3f} ms ({alg / tq / 1e6:.0f} GB/s on 240 B/constraint), "
          f"quotient+commit {tf:.3f} ms -> {W / tf * 1e3:.1f} witnesses/s, {W * chunks / tf * 1e3:.0f} commitments/s")


if __name__ == "__main__":
    main()
