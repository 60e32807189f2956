"""Does the host's DMA ceiling move with 2 MiB pages?  Same legs as tools/pcie_probe.py (D2H alone, the commitment's
2 out : 1 in mix), but the page-locked buffers are anonymous mappings with madvise(MADV_HUGEPAGE), touched, then registered
(cudaHostRegister) -- transparent huge pages, no reservation or privilege needed -- instead of cudaHostAlloc's 4 KiB pages.
    python tools/thp_probe.py <device> [seconds per leg] [--small]      (--small: plain 4 KiB pages through the same code path)
"""
import ctypes
import mmap
import sys
import time

import torch

dev = int(sys.argv[1]) if len(sys.argv) > 1 else 0
secs = float(sys.argv[2]) if len(sys.argv) > 2 and not sys.argv[2].startswith("--") else 1.0
small = "--small" in sys.argv
torch.cuda.set_device(dev)
n = 256 << 20
libc = ctypes.CDLL(None, use_errno=True)
MADV_HUGEPAGE, MADV_NOHUGEPAGE = 14, 15


def pinned(nbytes):
    mm = mmap.mmap(-1, nbytes + (2 << 20), flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
    base = ctypes.addressof(ctypes.c_char.from_buffer(mm))
    start = (base + (2 << 20) - 1) & ~((2 << 20) - 1)                      # 2 MiB aligned
    rc = libc.madvise(ctypes.c_void_p(start), ctypes.c_size_t(nbytes), MADV_NOHUGEPAGE if small else MADV_HUGEPAGE)
    t = torch.frombuffer(mm, dtype=torch.uint8, count=nbytes, offset=start - base)
    t.zero_()                                                              # touch: the pages materialise now
    err = torch.cuda.cudart().cudaHostRegister(start, nbytes, 0)
    assert int(err) == 0, f"cudaHostRegister -> {err}"
    return t, mm, rc


h_in, _k1, rc1 = pinned(n)
h_out, _k2, rc2 = pinned(2 * n)
huge = 0
for line in open("/proc/self/smaps_rollup"):
    if line.startswith("AnonHugePages"):
        huge = int(line.split()[1]) // 1024
d_in = torch.empty(n, dtype=torch.uint8, device="cuda"); d_out = torch.empty(2 * n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
assert h_in.is_pinned() and h_out.is_pinned()


def t(fn):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter(); reps = 0
    while time.perf_counter() - t0 < secs:
        fn(); torch.cuda.synchronize(); reps += 1
    return (time.perf_counter() - t0) / reps


def commit_mix():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)


d2h = t(lambda: h_out[:n].copy_(d_out[:n], non_blocking=True))
h2d = t(lambda: d_in.copy_(h_in, non_blocking=True))
mix = t(commit_mix)
print(f"gpu{dev} [{'4 KiB pages' if small else 'THP'}; madvise rc {rc1},{rc2}; AnonHugePages {huge} MiB of {3 * n >> 20}]: "
      f"H2D {n/h2d/1e9:.1f} GB/s  D2H {n/d2h/1e9:.1f} GB/s  commit mix (2 out : 1 in) {3*n/mix/1e9:.1f} GB/s total", flush=True)
