"""BASELINE configs[2]: batched forward / inverse NTT sweep, n = 2^10 .. 2^16, batch 1 .. 65536 (capped at 4 GiB),
device-resident data, CUDA events, median of 10 after 3 warm-ups.  q = 17592169062401 for n <= 4096 (2-adicity 13),
17592180539393 above.  Prints one JSON object; commit it under profiles/."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from lambda_snark_r_b200 import api  # noqa: E402

Q0, Q1 = 17592169062401, 17592180539393
api.set_device(0)
s = torch.cuda.current_stream().cuda_stream
hbm = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
rows = []
import os
LOGNS = [int(x) for x in os.environ.get("LSR_SWEEP_LOGN", "10,11,12,13,14,15,16").split(",")]
BATCHES = [int(x) for x in os.environ.get("LSR_SWEEP_BATCH", "1,16,256,4096,65536").split(",")]
for logn in LOGNS:
    n = 1 << logn
    q = Q0 if n <= 4096 else Q1
    ctx = api.NttContext(q, n)
    for batch in BATCHES:
        if batch * n * 8 > (4 << 30):
            batch = (4 << 30) // (n * 8)
        d = torch.randint(0, q, (batch, n), device="cuda", dtype=torch.int64)
        rec = {"n": n, "batch": batch}
        for name, fn in (("forward", ctx.forward_device), ("inverse", ctx.inverse_device)):
            for _ in range(3):
                fn(d.data_ptr(), batch, s)
            ts = []
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(d.data_ptr(), batch, s); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            passes = 1 if logn <= 14 else 2          # n > 2^14: column kernel + tile kernel, two HBM round trips
            rec[name] = {"ms": ms, "ntt_per_s": batch / (ms * 1e-3), "algorithmic_GBps": batch * n * 16 / (ms * 1e-3) / 1e9,
                         "hbm_frac": batch * n * 16 / (ms * 1e-3) / 1e9 / hbm, "hbm_round_trips": passes,
                         "gbutterfly_per_s": batch * (n // 2) * logn / (ms * 1e-3) / 1e9}
        rows.append(rec)
        del d
    ctx.close()
print(json.dumps({"q_small": Q0, "q_large": Q1, "hbm_peak_gbs": hbm, "rows": rows}))
