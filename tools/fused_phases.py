"""Phase cost breakdown of the fused commitment kernel: time it with individual phases
switched off (LSR_FUSED_SKIP bit mask; results are garbage, timing only).  The switch exists only in the
profiling build of the library (python -m lambda_snark_r_b200._build --profiling), loaded here in place
of the shipped one."""
import os, sys
from pathlib import Path
import torch
ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from lambda_snark_r_b200 import _build, capi
capi._lib = capi.load(_build.build_profiling())     # before api binds: every wrapper goes through capi.load()
from lambda_snark_r_b200 import api
Q, N, B = 17592169062401, 4096, 16384
api.set_device(0)
ctx = api.LweContext(api.Params(n=N, k=2, q=Q, sigma=3.19), seed32=bytes(range(32)))
s = torch.cuda.current_stream().cuda_stream
msgs = torch.randint(0, Q, (B, N), device="cuda", dtype=torch.int64)
seeds = torch.arange(1, B + 1, device="cuda", dtype=torch.int64)
out = torch.empty((B, ctx.words), device="cuda", dtype=torch.int64)
def t(mask):
    os.environ["LSR_FUSED_SKIP"] = str(mask)
    for _ in range(2): ctx.commit_batch_device(msgs.data_ptr(), N, seeds.data_ptr(), B, out.data_ptr(), s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): ctx.commit_batch_device(msgs.data_ptr(), N, seeds.data_ptr(), B, out.data_ptr(), s)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5
full = t(0)
print(f"full            {full:7.3f} ms  {B/full/1e3:6.2f} M/s")
for name, m in (("no sampler", 1), ("no fwd NTT", 2), ("no matvec", 4), ("no inv NTT", 8), ("no store", 16),
                ("only sampler", 30), ("only fwd", 29), ("only matvec", 27), ("only inv", 23), ("only store", 15), ("nothing", 31)):
    x = t(m)
    print(f"{name:15s} {x:7.3f} ms  (delta vs full {full - x:+7.3f})")
