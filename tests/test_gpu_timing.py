"""GPU test (-m gpu): timing independence of the CDT searches -- the device analogue of the reference's dudect harness
(cpp-core/tools/dudect_sampler.cpp:105-141: 20 000 traces of 64 samples, classes by the parity of the first sample of
the trace, Welch's t, |t| < 4.5).  Here a trace is a warp (32 samples), its time the clock64() ticks the warp spent
inside the search (lsr_cdt_timing_device), and two classifications are tested: the reference's (parity of the first
sample of the trace, random inputs throughout) and dudect's own fixed-vs-random one."""
import ctypes as C

import numpy as np
import pytest

from lambda_snark_r_b200 import capi

pytestmark = pytest.mark.gpu
WARPS = 16384


def timing(sigma, u, variant):
    lib = capi.load()
    out = np.zeros(u.size, dtype=np.uint32)
    cyc = np.zeros((u.size + 31) // 32, dtype=np.uint64)
    rc = lib.lsr_cdt_timing_device(sigma, u.ctypes.data_as(capi.u64p), u.size, variant,
                                   out.ctypes.data_as(C.POINTER(C.c_uint32)), cyc.ctypes.data_as(capi.u64p))
    assert rc == 0
    return out, cyc.astype(np.float64)


def welch_t(a, b):
    # dudect crops the slow tail (interrupts / preemption) before the test
    cut = np.percentile(np.concatenate([a, b]), 99.0)
    a, b = a[a <= cut], b[b <= cut]
    return float((a.mean() - b.mean()) / np.sqrt(a.var(ddof=1) / a.size + b.var(ddof=1) / b.size))


def t_statistics(sigma, variant, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    u = rng.integers(0, 2**64, size=32 * WARPS, dtype=np.uint64)
    timing(sigma, u, variant)                                          # warm-up launch (instruction cache, clocks)
    out, cyc = timing(sigma, u, variant)
    first = out[::32] & 1                                              # the reference's classes
    t_ref = welch_t(cyc[first == 0], cyc[first == 1])
    fixed = rng.integers(0, 2, size=WARPS).astype(bool)                # dudect's classes: fixed input vs random input
    u2 = u.copy().reshape(WARPS, 32)
    u2[fixed, :] = np.uint64(0x8000000000000000)
    _, cyc2 = timing(sigma, u2.reshape(-1), variant)
    return t_ref, welch_t(cyc2[fixed], cyc2[~fixed])


@pytest.mark.parametrize("variant", [2, 3, 4])
def test_cdt_search_time_does_not_depend_on_the_samples(gpu, variant):
    runs = [t_statistics(3.2, variant, 1000 + i) for i in range(5)]      # median of five: one noisy launch cannot fail the suite
    t_ref = float(np.median([abs(r[0]) for r in runs]))
    t_fix = float(np.median([abs(r[1]) for r in runs]))
    print(f"variant {variant}: |t| by first-sample parity {t_ref:.2f}, fixed vs random {t_fix:.2f}")
    assert t_ref < 4.5 and t_fix < 4.5, (variant, runs)
