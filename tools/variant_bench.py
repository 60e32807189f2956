"""Device-resident rates of the three headline kernels for a library variant built by `_build --variant` (or the shipped
one): python tools/variant_bench.py [--lib path/to/liblambda_snark_core_<name>.so]
Also checks every result against the shipped library's (bit-identical outputs are the contract)."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from lambda_snark_r_b200 import capi  # noqa: E402
lib_path = Path(sys.argv[sys.argv.index("--lib") + 1]) if "--lib" in sys.argv else None
if lib_path:
    capi._lib = capi.load(lib_path)
from lambda_snark_r_b200 import api  # noqa: E402

Q, N, K, B = 17592169062401, 4096, 2, 16384
api.set_device(0)
s = torch.cuda.current_stream().cuda_stream
ctx = api.LweContext(api.Params(n=N, k=K, q=Q, sigma=3.19), seed32=bytes(range(32)))
ntt = api.NttContext(Q, N)
g = torch.Generator(device="cuda"); g.manual_seed(1)
msgs = torch.randint(0, Q, (B, N), device="cuda", dtype=torch.int64, generator=g)
seeds = torch.arange(1, B + 1, device="cuda", dtype=torch.int64)
out = torch.empty((B, ctx.words), device="cuda", dtype=torch.int64)
data = torch.randint(0, Q, (B, N), device="cuda", dtype=torch.int64, generator=g)


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


tc = timed(lambda: ctx.commit_batch_device(msgs.data_ptr(), N, seeds.data_ptr(), B, out.data_ptr(), s))
tf = timed(lambda: ntt.forward_device(data.data_ptr(), B, s))
ti = timed(lambda: ntt.inverse_device(data.data_ptr(), B, s))
print(f"{lib_path.name if lib_path else 'shipped'}: commit {B / tc / 1e3:.3f} M/s   ntt fwd {B / tf / 1e3:.2f} M/s   inv {B / ti / 1e3:.2f} M/s")
print("checksums", int(out.sum().item()), int(data.sum().item()))
