"""Tiny driver for ncu captures: python tools/prof_driver.py {commit|verify|ntt|prover} [batch]
commit: 3 launches of the fused commitment kernel; verify: one commitment launch + 3 of the fused verification kernel;
ntt: forward, inverse, pointwise x3; prover: the 2^20-constraint commitment phase x3.  Deterministic synthetic inputs, no host-side work between launches."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from lambda_snark_r_b200 import api  # noqa: E402

Q, N, K = 17592169062401, 4096, 2
what = sys.argv[1] if len(sys.argv) > 1 else "commit"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
api.set_device(0)
s = torch.cuda.current_stream().cuda_stream
if what == "commit":
    ctx = api.LweContext(api.Params(n=N, k=K, q=Q, sigma=3.19), seed32=bytes(range(32)))
    msgs = torch.randint(0, Q, (B, N), device="cuda", dtype=torch.int64)
    seeds = torch.arange(1, B + 1, device="cuda", dtype=torch.int64)
    out = torch.empty((B, ctx.words), device="cuda", dtype=torch.int64)
    for _ in range(3):
        ctx.commit_batch_device(msgs.data_ptr(), N, seeds.data_ptr(), B, out.data_ptr(), s)
elif what == "verify":
    ctx = api.LweContext(api.Params(n=N, k=K, q=Q, sigma=3.19), seed32=bytes(range(32)))
    msgs = torch.randint(0, Q, (B, N), device="cuda", dtype=torch.int64)
    seeds = torch.arange(1, B + 1, device="cuda", dtype=torch.int64)
    out = torch.empty((B, ctx.words), device="cuda", dtype=torch.int64)
    diff = torch.zeros(B, device="cuda", dtype=torch.int64)
    inv = torch.zeros(B, device="cuda", dtype=torch.int32)
    ctx.commit_batch_device(msgs.data_ptr(), N, seeds.data_ptr(), B, out.data_ptr(), s)
    for _ in range(3):
        ctx.verify_batch_device(out.data_ptr(), msgs.data_ptr(), N, B, diff.data_ptr(), inv.data_ptr(), s)
    torch.cuda.synchronize()
    assert int(diff.abs().sum()) == 0 and int(inv.sum()) == 0, "an honest opening failed to verify"
elif what == "ntt":
    ntt = api.NttContext(Q, N)
    a = torch.randint(0, Q, (B, N), device="cuda", dtype=torch.int64)
    b = torch.randint(0, Q, (B, N), device="cuda", dtype=torch.int64)
    for _ in range(3):
        ntt.forward_device(a.data_ptr(), B, s)
        ntt.inverse_device(a.data_ptr(), B, s)
        ntt.mul_pointwise_device(b.data_ptr(), a.data_ptr(), b.data_ptr(), B * N, s)
else:
    import subprocess
    subprocess.run([sys.executable, str(ROOT / "tools" / "prover_phase.py"), "20", "4"], check=True)
torch.cuda.synchronize()
