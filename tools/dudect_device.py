"""Report of the device-side dudect analogue (tests/test_gpu_timing.py): Welch t of the warp timing of the CDT searches
for the reference's classification (parity of the first sample) and for fixed-vs-random inputs.  |t| < 4.5 passes."""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import test_gpu_timing as T  # noqa: E402

print("device dudect analogue: 16384 traces (warps) of 32 samples per class pair, sigma = 3.2, clock64() per warp, top 1 % cropped")
for variant, name in ((2, "shuffle binary search + tail scan"), (3, "compact carry-chain search"), (4, "25-bit prefix search (fused kernel)")):
    for i in range(3):
        t_ref, t_fix = T.t_statistics(3.2, variant, 2000 + i)
        print(f"  variant {variant} ({name}), run {i}: t(first-sample parity) = {t_ref:+.2f}   t(fixed vs random) = {t_fix:+.2f}")
