"""Small driver for ncu: a few launches of each hot kernel at bench sizes.
usage: python tools/prof_driver.py [ntt|commit|all] [batch]"""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from lambda_snark_r_b200 import api  # noqa: E402

Q, N = 17592169062401, 4096
what = sys.argv[1] if len(sys.argv) > 1 else "all"
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
api.set_device(0)
s = torch.cuda.current_stream().cuda_stream
if what in ("ntt", "all"):
    ntt = api.NttContext(Q, N)
    d = torch.randint(0, Q, (batch, N), device="cuda", dtype=torch.int64)
    for _ in range(3):
        ntt.forward_device(d.data_ptr(), batch, s)
    for _ in range(3):
        ntt.inverse_device(d.data_ptr(), batch, s)
    a = torch.randint(0, Q, (batch, N), device="cuda", dtype=torch.int64)
    for _ in range(2):
        ntt.mul_pointwise_device(a.data_ptr(), a.data_ptr(), d.data_ptr(), batch * N, s)
    torch.cuda.synchronize()
if what in ("commit", "all"):
    ctx = api.LweContext(api.Params(n=N, k=2, q=Q, sigma=3.19), seed32=bytes(range(32)))
    cb = min(batch, 4096)
    msgs = torch.randint(0, Q, (cb, N), device="cuda", dtype=torch.int64)
    seeds = torch.arange(1, cb + 1, device="cuda", dtype=torch.int64)
    out = torch.empty((cb, ctx.words), device="cuda", dtype=torch.int64)
    for _ in range(3):
        ctx.commit_batch_device(msgs.data_ptr(), N, seeds.data_ptr(), cb, out.data_ptr(), s)
    torch.cuda.synchronize()
print("done")
