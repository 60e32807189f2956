"""Forward / inverse transform rates at small ring degrees for a library variant (pass-plan experiments):
python tools/plan_bench.py [--lib path] [log2 n ...]   (default 10 11; 2^26 coefficients per launch, checksums for comparison)"""
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

Q_SMALL, Q_LARGE = 17592169062401, 17592180539393          # 2-adicity 13 / 18
logs = [int(a) for a in sys.argv[1:] if a.isdigit()] or [10, 11]
api.set_device(0)
s = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for logn in logs:
    n, B = 1 << logn, (1 << 26) >> logn
    Q = Q_SMALL if logn <= 12 else Q_LARGE
    ntt = api.NttContext(Q, n)
    g = torch.Generator(device="cuda"); g.manual_seed(logn)
    data = torch.randint(0, Q, (B, n), device="cuda", dtype=torch.int64, generator=g)
    ntt.forward_device(data.data_ptr(), B, s); torch.cuda.synchronize()
    cf = int(data.sum().item())
    ntt.inverse_device(data.data_ptr(), B, s); torch.cuda.synchronize()
    ci = int(data.sum().item())
    tf = timed(lambda: ntt.forward_device(data.data_ptr(), B, s))
    ti = timed(lambda: ntt.inverse_device(data.data_ptr(), B, s))
    print(f"{lib_path.name if lib_path else 'shipped'}: n=2^{logn} batch {B}: fwd {B / tf / 1e3:.2f} M/s  inv {B / ti / 1e3:.2f} M/s  checksums {cf} {ci}")
    ntt.close()
