"""GPU test (-m gpu): compile tests/cabi/test_cabi.c against include/ with the
reference's header paths and link it to liblambda_snark_core.so -- the same way
cpp-core's own tests and lambda-snark-sys consume the library -- then run it."""
import subprocess
from pathlib import Path

import pytest

from lambda_snark_r_b200 import capi

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def test_c_harness_links_and_passes(gpu, tmp_path):
    exe = tmp_path / "test_cabi"
    lib_dir = capi.LIB_PATH.parent
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cabi" / "test_cabi.c"),
           "-L", str(lib_dir), "-llambda_snark_core", f"-Wl,-rpath,{lib_dir}", "-lm", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all reference assertions hold" in r.stdout
