"""BASELINE configs[2]: batched forward / inverse NTT sweep, ring degree 2^10 .. 2^16 (plus the cyclic 2^20 of the
quotient pipeline), batch 1 .. 65536 (capped at 4 GiB of coefficients), data resident in HBM, CUDA events on the
launching stream, median of 9 after 3 warm-ups.  q = 17592169062401 for n <= 4096 (2-adicity 13),
17592180539393 (2-adicity 18) above.  Writes one JSON document to stdout."""
import json
import statistics
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from lambda_snark_r_b200 import api  # noqa: E402

Q_SMALL, Q_LARGE, GOLD = 17592169062401, 17592180539393, 2**64 - 2**32 + 1


def hbm_peak():
    try:
        return json.load(open(ROOT / "MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception:
        return 6549.4


def timed(fn, reps=9, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    out = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    return statistics.median(out)


def main():
    api.set_device(0)
    s = torch.cuda.current_stream().cuda_stream
    peak = hbm_peak()
    rows = []
    shapes = [(logn, Q_SMALL if logn <= 12 else Q_LARGE, False) for logn in range(10, 17)] + [(20, GOLD, True)]
    only = [int(a) for a in sys.argv[1:] if a.isdigit()]          # optional: ring-degree exponents to run
    arith = 1 if "--u64" in sys.argv else 0                       # --u64: integer butterflies instead of the FP64 default
    if only:
        shapes = [sh for sh in shapes if sh[0] in only]
    for logn, q, cyclic in shapes:
        n = 1 << logn
        ctx = api.CyclicNtt(q, n) if cyclic else api.NttContext(q, n)
        if arith and not cyclic:
            ctx.set_arith(arith)
        for batch in (1, 16, 256, 4096, 65536):
            if batch * n * 8 > (4 << 30):
                continue
            data = torch.randint(0, min(q, 2**62), (batch, n), device="cuda", dtype=torch.int64)
            trips = 1 if logn <= 13 else 1 + (logn - 12 + 4) // 5
            row = {"n": n, "batch": batch, "q": q, "cyclic": cyclic}
            for name, fn in (("forward", lambda: ctx.forward_device(data.data_ptr(), batch, s)),
                             ("inverse", lambda: ctx.inverse_device(data.data_ptr(), batch, s))):
                ms = timed(fn)
                gbs = batch * n * 16 / (ms * 1e-3) / 1e9
                row[name] = {"ms": ms, "ntt_per_s": batch / (ms * 1e-3), "algorithmic_GBps": gbs, "hbm_frac": gbs / peak,
                             "hbm_round_trips": trips, "gbutterfly_per_s": batch * (n // 2) * logn / (ms * 1e-3) / 1e9}
            rows.append(row)
            del data
        ctx.close()
    print(json.dumps({"q_small": Q_SMALL, "q_large": Q_LARGE, "goldilocks": GOLD, "hbm_peak_gbs": peak, "rows": rows}))


if __name__ == "__main__":
    main()
