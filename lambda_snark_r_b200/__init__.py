"""lambda_snark_r_b200 -- B200-native prover hot path of LambdaSNARK-R.

(The task names the package directory `lambda-snark-r_b200/`; Python cannot
import a hyphenated name, so the directory is `lambda_snark_r_b200/`.)

Contents: csrc/ (hand-written sm_100a CUDA kernels + the extern "C" ABI),
_build.py (in-tree nvcc build), capi.py (raw ctypes declarations, the analogue
of lambda-snark-sys), api.py (the analogue of the Rust safe wrappers).
There is no CPU fallback: importing the API without the built CUDA library
raises.
"""
from .api import (Commitment, LambdaSnarkError, LweContext, NttContext, Opening, Params,  # noqa: F401
                  device_count, generate_opening, last_error, sample_gaussian, set_device,
                  verify_commitment, verify_opening_with_context)
