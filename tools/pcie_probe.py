"""PCIe probe: pinned host <-> device copy rates of one GPU (argv[1] = device index, argv[2] = seconds per leg).
Run one process per GPU at the same time to see what the host side sustains in aggregate."""
import sys
import time

import torch

dev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
secs = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
torch.cuda.set_device(dev)
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); reps = 0
    while time.perf_counter() - t0 < secs:
        fn(); torch.cuda.synchronize(); reps += 1
    return (time.perf_counter() - t0) / reps


def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)


h2d = t(lambda: d_in.copy_(h_in, non_blocking=True))
d2h = t(lambda: h_out.copy_(d_out, non_blocking=True))
bi = t(both)
print(f"gpu{dev}: H2D {n/h2d/1e9:.1f} GB/s  D2H {n/d2h/1e9:.1f} GB/s  bidirectional {2*n/bi/1e9:.1f} GB/s total", flush=True)
