"""Small batches (BASELINE configs[2], batch 1 .. 256): where the ~16 us per call of tools/sweep.py goes.
Times ntt_forward at n = 4096 three ways: one call per CUDA-event pair (what sweep.py reports), 200 calls back to back
(launch overhead overlapped with the previous kernel), and the same 200 launches replayed from a CUDA graph (no host
in the loop).  Under `ncu --metrics gpu__time_duration.sum` the same script gives the kernel-only durations."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
from lambda_snark_r_b200 import api  # noqa: E402

Q, N = 17592169062401, 4096
api.set_device(0)
ntt = api.NttContext(Q, N)
for batch in (1, 16, 148, 256, 444, 1024, 4096):
    data = torch.randint(0, Q, (batch, N), device="cuda", dtype=torch.int64)
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(5):
        ntt.forward_device(data.data_ptr(), batch, s)
    torch.cuda.synchronize()
    single = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); ntt.forward_device(data.data_ptr(), batch, s); e1.record(); torch.cuda.synchronize()
        single.append(e0.elapsed_time(e1) * 1e3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        ntt.forward_device(data.data_ptr(), batch, s)
    e1.record(); torch.cuda.synchronize()
    stream_us = e0.elapsed_time(e1) * 1e3 / 200
    g = torch.cuda.CUDAGraph()
    cs = torch.cuda.Stream()
    with torch.cuda.stream(cs):
        with torch.cuda.graph(g, stream=cs):
            for _ in range(200):
                ntt.forward_device(data.data_ptr(), batch, cs.cuda_stream)
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    graph_us = e0.elapsed_time(e1) * 1e3 / 200
    single.sort()
    print(f"n=4096 batch {batch:5d}: one call per event pair {single[len(single) // 2]:6.2f} us | 200 calls back to back "
          f"{stream_us:6.2f} us/call | CUDA graph of 200 launches {graph_us:6.2f} us/launch = {batch / graph_us:7.2f} M NTT/s", flush=True)
ntt.close()
