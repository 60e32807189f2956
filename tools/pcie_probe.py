"""PCIe probe: pinned host <-> device copy rates of one GPU.

    python tools/pcie_probe.py <device> [seconds per leg] [--cpu-threads T]

Run one process per GPU at the same time (tools/multi_gpu_probe.sh) to see what the HOST sustains in aggregate.
Legs: H2D alone, D2H alone, 1:1 both ways, and the mix of lwe_commit_batch (64 KiB out per 32 KiB in per
commitment: 2 bytes device->host for every byte host->device).  With --cpu-threads T a fifth leg repeats the
commitment mix while T host threads stream-copy 112 bytes per 80 bytes of DMA in the background: what a packed wire
format (48-bit planes, unpacked by host threads) would ask of the host's memory system on top of the DMA.
"""
import sys
import threading
import time

import torch

dev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
secs = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 1.0
cpu_threads = int(sys.argv[sys.argv.index("--cpu-threads") + 1]) if "--cpu-threads" in sys.argv else 0
torch.cuda.set_device(dev)
n = 256 << 20
h_in = torch.empty(n, dtype=torch.uint8).pin_memory(); h_out = torch.empty(2 * n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); reps = 0
    while time.perf_counter() - t0 < secs:
        fn(); torch.cuda.synchronize(); reps += 1
    return (time.perf_counter() - t0) / reps


def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out[:n].copy_(d_out[:n], non_blocking=True)


def commit_mix():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)


h2d = t(lambda: d_in.copy_(h_in, non_blocking=True))
d2h = t(lambda: h_out[:n].copy_(d_out[:n], non_blocking=True))
bi = t(both)
mix = t(commit_mix)
line = (f"gpu{dev}: H2D {n/h2d/1e9:.1f} GB/s  D2H {n/d2h/1e9:.1f} GB/s  1:1 both ways {2*n/bi/1e9:.1f} GB/s total  "
        f"commit mix (2 out : 1 in) {3*n/mix/1e9:.1f} GB/s total")
if cpu_threads:
    torch.set_num_threads(cpu_threads)
    src = torch.empty(96 << 20, dtype=torch.uint8); dst = torch.empty(96 << 20, dtype=torch.uint8)
    stop = False
    copied = [0]

    def cpu_loop():
        while not stop:
            dst.copy_(src); copied[0] += src.numel()

    th = threading.Thread(target=cpu_loop); th.start()
    c0 = copied[0]; t0 = time.perf_counter()
    mix2 = t(commit_mix)
    cpu_rate = (copied[0] - c0) / (time.perf_counter() - t0)
    stop = True; th.join()
    line += (f"  | commit mix with {cpu_threads} host copy threads running: DMA {3*n/mix2/1e9:.1f} GB/s, "
             f"host copy {cpu_rate/1e9:.1f} GB/s (read + write each)")
print(line, flush=True)
