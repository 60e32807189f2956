"""Device-resident rate of the fused verification kernel (lwe_verify_opening for a batch): python tools/verify_bench.py [batch]
Commits a batch, verifies it (every opening must pass), then flips one message word per commitment (every opening must fail)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from lambda_snark_r_b200 import api  # noqa: E402

Q, N, K = 17592169062401, 4096, 2
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
api.set_device(0)
s = torch.cuda.current_stream().cuda_stream
ctx = api.LweContext(api.Params(n=N, k=K, q=Q, sigma=3.19), seed32=bytes(range(32)))
g = torch.Generator(device="cuda"); g.manual_seed(1)
msgs = torch.randint(0, Q, (B, N), device="cuda", dtype=torch.int64, generator=g)
seeds = torch.arange(1, B + 1, device="cuda", dtype=torch.int64)
out = torch.empty((B, ctx.words), device="cuda", dtype=torch.int64)
diff = torch.zeros(B, device="cuda", dtype=torch.int64)
inv = torch.zeros(B, device="cuda", dtype=torch.int32)
ctx.commit_batch_device(msgs.data_ptr(), N, seeds.data_ptr(), B, out.data_ptr(), s)


def verify():
    ctx.verify_batch_device(out.data_ptr(), msgs.data_ptr(), N, B, diff.data_ptr(), inv.data_ptr(), s)


for _ in range(3):
    verify()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20):
    verify()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
assert int(diff.count_nonzero()) == 0 and int(inv.count_nonzero()) == 0, "an honest opening failed to verify"
msgs[:, 17] = (msgs[:, 17] + 1) % ctx.p
verify(); torch.cuda.synchronize()
assert int(diff.count_nonzero()) == B, "a changed message word went unnoticed"
print(f"verify_opening, batch {B}: {ms:.4f} ms, {B / ms / 1e3:.2f} M openings/s")
