"""End-to-end commit rate of one process on one GPU (lwe_commit_batch, page-locked buffers); run several at once
(one per GPU) to see how the host side shares: python tools/e2e_probe.py <device> [batch] [seconds]"""
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from lambda_snark_r_b200 import capi  # noqa: E402
if "--lib" in sys.argv:          # kernel / pipeline variant built by `_build --variant`
    capi._lib = capi.load(Path(sys.argv.pop(sys.argv.index("--lib") + 1))); sys.argv.remove("--lib")
from lambda_snark_r_b200 import api  # noqa: E402

dev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
secs = float(sys.argv[3]) if len(sys.argv) > 3 else 3.0
Q, N, K = 17592169062401, 4096, 2
torch.cuda.set_device(dev)
api.set_device(dev)
ctx = api.LweContext(api.Params(n=N, k=K, q=Q, sigma=3.19), seed32=bytes(range(32)))
h_msgs = torch.randint(0, Q, (B, N), dtype=torch.int64).pin_memory()
h_seeds = torch.arange(1, B + 1, dtype=torch.int64).pin_memory()
h_out = torch.empty((B, ctx.words), dtype=torch.int64).pin_memory()
for _ in range(2):
    ctx.commit_batch_ptr(h_msgs.data_ptr(), N, h_seeds.data_ptr(), B, h_out.data_ptr())
t0 = time.perf_counter(); reps = 0
while time.perf_counter() - t0 < secs:
    ctx.commit_batch_ptr(h_msgs.data_ptr(), N, h_seeds.data_ptr(), B, h_out.data_ptr()); reps += 1
dt = (time.perf_counter() - t0) / reps
print(f"gpu{dev}: {B / dt / 1e6:.3f} M commitments/s ({dt * 1e3:.2f} ms per {B}; D2H {B * ctx.words * 8 / dt / 1e9:.1f} GB/s + H2D {B * N * 8 / dt / 1e9:.1f} GB/s)", flush=True)
