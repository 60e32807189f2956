"""CPU tests (-m "not gpu"): the C-ABI library loads without a GPU, exports every
symbol include/lambda_snark_b200.h declares, and keeps the reference's struct
layouts (cpp-core/include/lambda_snark/types.h, r1cs.h).  No GPU compute here."""
import ctypes as C
import re
import subprocess
from pathlib import Path

import pytest

from lambda_snark_r_b200 import capi

ROOT = Path(__file__).resolve().parents[1]
HEADER = ROOT / "include" / "lambda_snark_b200.h"

# the symbols the reference exports for this path, Lean exporters included
REFERENCE_SYMBOLS = [
    "export_vk_to_lean", "export_params_to_lean", "export_seal_context_to_lean", "export_seal_pubkey_to_lean",
    "lwe_context_create", "lwe_context_free", "lwe_commit", "lwe_commitment_free", "lwe_commitment_clone",
    "lwe_verify_opening", "lwe_linear_combine", "ntt_context_create", "ntt_context_free", "ntt_forward",
    "ntt_inverse", "ntt_mul_pointwise", "sample_gaussian", "lambda_snark_r1cs_create",
    "lambda_snark_r1cs_validate_witness", "lambda_snark_r1cs_free", "lambda_snark_r1cs_num_constraints",
    "lambda_snark_r1cs_num_variables",
]


def declared_functions() -> list[str]:
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    text = re.sub(r"typedef\s+(struct|enum)\s*\{.*?\}\s*\w+\s*;", "", text, flags=re.S)
    names = re.findall(r"\b([a-z_][a-z0-9_]*)\s*\([^;{}]*\)\s*(?:LSR_NOEXCEPT)?\s*;", text)
    return sorted(set(names))


def test_library_is_built_and_loads():
    assert capi.LIB_PATH.exists(), "build the CUDA library first: python -m lambda_snark_r_b200._build"
    lib = capi.load()
    assert b"sm_100a" in lib.lsr_version()
    assert lib.lsr_device_count() >= 0


def test_every_declared_symbol_is_exported():
    lib = C.CDLL(str(capi.LIB_PATH))
    declared = declared_functions()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    for name in REFERENCE_SYMBOLS:
        assert name in declared and hasattr(lib, name)


def test_ctypes_table_covers_the_header():
    assert sorted(capi.SIGNATURES) == declared_functions()


def test_struct_layouts_match_the_reference():
    assert C.sizeof(capi.PublicParams) == 32                        # types.h:60-67
    assert capi.PublicParams.modulus.offset == 8
    assert capi.PublicParams.ring_degree.offset == 16
    assert capi.PublicParams.module_rank.offset == 20
    assert capi.PublicParams.sigma.offset == 24
    assert C.sizeof(capi.LweCommitment) == 16 and capi.LweCommitment.len.offset == 8
    assert C.sizeof(capi.LweOpening) == 16
    assert C.sizeof(capi.SparseEntry) == 16 and capi.SparseEntry.value.offset == 8
    assert C.sizeof(capi.SparseMatrix) == 24 and capi.SparseMatrix.n_cols.offset == 20
    assert C.sizeof(capi.R1CSWitness) == 16


def test_header_compiles_as_c_and_cxx_with_static_layout_asserts(tmp_path):
    src = r'''
#include "lambda_snark/types.h"
#include "lambda_snark/ntt.h"
#include "lambda_snark/commitment.h"
#include "lambda_snark/utils.h"
#include "lambda_snark/r1cs.h"
#include <stddef.h>
#ifdef __cplusplus
#define SA static_assert
#else
#define SA _Static_assert
#endif
SA(sizeof(PublicParams) == 32, "PublicParams");
SA(offsetof(PublicParams, sigma) == 24, "sigma");
SA(sizeof(LweCommitment) == 16, "LweCommitment");
SA(sizeof(SparseEntry) == 16, "SparseEntry");
SA(sizeof(SparseMatrix) == 24, "SparseMatrix");
SA(LAMBDA_SNARK_ERR_CRYPTO_FAILED == 4, "error codes");
int main(void) { NttContext* (*f)(uint64_t, uint32_t) = ntt_context_create; return sizeof(f) == 0; }
'''
    for name, cc, std in (("t.c", "gcc", "-std=c11"), ("t.cpp", "g++", "-std=c++17")):
        f = tmp_path / name
        f.write_text(src)
        r = subprocess.run([cc, std, "-Wall", "-Werror", "-I", str(ROOT / "include"), "-c", str(f), "-o",
                            str(tmp_path / (name + ".o"))], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_null_and_invalid_arguments_need_no_gpu():
    lib = capi.load()
    # lambda-snark-sys/src/lib.rs:28-34
    assert not lib.lwe_context_create(None)
    # ntt.cpp:31,41 / SEAL rules: rejected before any device work
    for q, n in ((12289, 0), (12289, 3), (12289, 1), (12289, 4096), (1 << 61, 16), (17592186044417, 4096)):
        assert not lib.ntt_context_create(q, n)
    lib.ntt_context_free(None)                                       # test_ntt.cpp:89
    lib.lwe_context_free(None)
    lib.lwe_commitment_free(None)                                    # test_commitment.cpp:112
    assert lib.ntt_forward(None, None, 4) == -1                      # test_ntt.cpp:86
    assert not lib.lwe_commit(None, None, 0, 0)                      # test_commitment.cpp:104
    assert not lib.lwe_commitment_clone(None)
    assert lib.lwe_verify_opening(None, None, None, 0, None) == -1
    assert not lib.lwe_linear_combine(None, None, None, 0)
    buf = (C.c_uint64 * 16)()
    assert lib.sample_gaussian(None, 16, 3.2) == -1                  # test_utils.cpp:29-33
    assert lib.sample_gaussian(buf, 0, 3.2) == -1
    assert lib.sample_gaussian(buf, 16, 0.0) == -1
    assert lib.sample_gaussian(buf, 16, float("inf")) == -1
    # extensions added for explicit-mode commitments and the Goldilocks probe: NULL arguments are refused up front
    assert lib.lsr_lwe_commit_explicit(None, None, 4, None, None, 1, None) == -1
    assert lib.lsr_lwe_commit_explicit_device(None, None, 4, None, None, 1, None, None) == -1
    assert lib.lsr_goldilocks_probe_device(None, None, 4, None) == -1
    # peer-memory gather: NULL handles / pointers are refused, NULL frees are no-ops
    assert lib.lsr_lwe_verify_opening_batch_device(None, None, None, 4, 1, None, None, None) == -1
    assert lib.lsr_peer_export(None, None) == -1
    assert not lib.lsr_peer_open(None)
    assert lib.lsr_peer_close(None) == 0
    assert not lib.lsr_device_alloc(0)
    lib.lsr_device_free(None)


def test_missing_library_fails_loudly(tmp_path):
    with pytest.raises(FileNotFoundError):
        capi.load(tmp_path / "liblambda_snark_core.so")


def test_static_archive_resolves_with_the_cargo_link_line(tmp_path):
    """A12 (rust-api/lambda-snark-sys/build.rs:106-124,180 as patched in INTEGRATION.md section 2): the static archive,
    cudart_static, dl, rt, pthread, stdc++ resolve every symbol of a C consumer.  Linking needs no GPU; the binary is
    run by tests/test_gpu_cabi_c.py."""
    import subprocess
    exe = tmp_path / "test_cabi_static"
    lib_dir = capi.LIB_PATH.parent
    cmd = ["gcc", "-std=c11", "-O1", "-I", str(ROOT / "include"), str(ROOT / "tests" / "cabi" / "test_cabi.c"),
           "-L", str(lib_dir), "-l:liblambda_snark_core.a", "-L/usr/local/cuda/lib64", "-lcudart_static", "-ldl", "-lrt",
           "-lpthread", "-lstdc++", "-lm", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ldd = subprocess.run(["ldd", str(exe)], capture_output=True, text=True).stdout
    assert "lambda_snark_core" not in ldd and "libcudart" not in ldd, ldd


def test_no_profiling_switch_ships_in_the_library():
    """The phase-skip / kernel-selection switches of tools/ exist only in the -DLSR_PROFILING build: the shipped
    library reads no LSR_* environment variable."""
    blob = capi.LIB_PATH.read_bytes() + (capi.LIB_PATH.parent / "liblambda_snark_core.a").read_bytes()
    for name in (b"LSR_FUSED_SKIP", b"LSR_COMMIT_CHUNK", b"LSR_PROVER_PIPELINE", b"LSR_NTT_COLUMN2", b"LSR_FS_KERNEL",
                 b"LSR_QUOT_FUSE"):
        assert name not in blob, name


def test_cxx_host_mirror_compiles_and_links_without_a_gpu(tmp_path):
    """The C++ mirror of the Rust safe wrappers (include/lambda_snark_b200.hpp) and its test program build against the
    library here; tests/test_gpu_cabi_c.py runs the program on the GPU."""
    exe = tmp_path / "test_host_mirror"
    lib_dir = capi.LIB_PATH.parent
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", str(ROOT / "include"),
           str(ROOT / "tests" / "cabi" / "test_host_mirror.cpp"), "-L", str(lib_dir), "-llambda_snark_core",
           f"-Wl,-rpath,{lib_dir}", "-o", str(exe)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
