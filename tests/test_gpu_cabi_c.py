"""GPU test (-m gpu): compile tests/cabi/test_cabi.c against include/ with the
reference's header paths and link it to liblambda_snark_core.so -- the same way
cpp-core's own tests consume the library -- and to the STATIC archive with exactly
the link line lambda-snark-sys would emit (INTEGRATION.md section 2;
rust-api/lambda-snark-sys/build.rs:106-124,180), then run both."""
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


STATIC_LINK_LINE = ["-l:liblambda_snark_core.a", "-L/usr/local/cuda/lib64", "-lcudart_static", "-ldl", "-lrt", "-lpthread",
                    "-lstdc++"]


def test_static_archive_links_with_the_cargo_link_line_and_passes(gpu, tmp_path):
    """A12: Cargo links `static=lambda_snark_core` + `static=cudart_static` + dl, rt, pthread, stdc++ -- no shared
    object of this repo at run time."""
    exe = tmp_path / "test_cabi_static"
    lib_dir = capi.LIB_PATH.parent
    assert (lib_dir / "liblambda_snark_core.a").exists()
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cabi" / "test_cabi.c"),
           "-L", str(lib_dir), *STATIC_LINK_LINE, "-lm", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ldd = subprocess.run(["ldd", str(exe)], capture_output=True, text=True).stdout
    assert "lambda_snark_core" not in ldd and "libcudart" not in ldd, ldd
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all reference assertions hold" in r.stdout


def test_cxx_host_mirror_of_the_rust_wrappers(gpu, tmp_path):
    """include/lambda_snark_b200.hpp mirrors context.rs / commitment.rs / opening.rs / challenge.rs in C++ (no Rust
    toolchain here); tests/cabi/test_host_mirror.cpp ports the Rust unit tests that sit on the FFI."""
    exe = tmp_path / "test_host_mirror"
    lib_dir = capi.LIB_PATH.parent
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cabi" / "test_host_mirror.cpp"),
           "-L", str(lib_dir), "-llambda_snark_core", f"-Wl,-rpath,{lib_dir}", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "all ported Rust assertions hold" in r.stdout
