"""pytest configuration.

`-m "not gpu"`: oracle vs golden vectors / reference sampler, host logic,
C-ABI library loads and exports every declared symbol (no compute calls).
`-m gpu`: the parity tests proper -- every call goes through the C ABI of
lambda_snark_r_b200/lib/liblambda_snark_core.so and is compared with the CPU
oracle under oracle/ (test infrastructure).
"""
import ctypes
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

Q0 = 17592169062401            # reference r1cs.rs:527 NTT_FRIENDLY_MODULUS
Q1 = 17592180539393            # 44-bit, 2-adicity 18
Q60 = 1152921504606584833      # 60-bit, 2-adicity 18 (guarded Harvey path)
Q50 = 1125899902124033         # 50-bit
Q31 = 2146959361               # 31-bit
Q45 = 35184365273089           # largest prime < 2^45 with 2^18 | q-1: worst case for the FP64 butterflies


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _device_count() -> int:
    from lambda_snark_r_b200 import capi
    try:
        return int(capi.load().lsr_device_count())
    except FileNotFoundError:
        return -1


@pytest.fixture(scope="session")
def gpu():
    n = _device_count()
    if n < 0:
        pytest.fail("liblambda_snark_core.so is not built: the CUDA path is mandatory (python -m lambda_snark_r_b200._build)")
    if n == 0:
        pytest.skip("no CUDA device visible")
    return n


@pytest.fixture(scope="session")
def rng():
    return np.random.Generator(np.random.PCG64(0x5EED))


def uniform(rng, q, shape):
    return rng.integers(0, q, size=shape, dtype=np.uint64)
