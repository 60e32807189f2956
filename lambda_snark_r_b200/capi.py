"""ctypes binding of the C ABI (include/lambda_snark_b200.h).

This is the Python equivalent of what rust-api/lambda-snark-sys generates with
bindgen (build.rs:184-207): raw declarations only, no logic.  The library is
built in-tree by lambda_snark_r_b200._build; loading fails loudly when it is
missing -- there is no CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "liblambda_snark_core.so"

u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)


class LweCommitment(C.Structure):          # types.h:36-39
    _fields_ = [("data", u64p), ("len", C.c_size_t)]


class LweOpening(C.Structure):             # types.h:44-47
    _fields_ = [("randomness", u64p), ("rand_len", C.c_size_t)]


class PublicParams(C.Structure):           # types.h:60-67
    _fields_ = [("profile", C.c_int), ("security_level", C.c_uint32), ("modulus", C.c_uint64),
                ("ring_degree", C.c_uint32), ("module_rank", C.c_uint32), ("sigma", C.c_double)]


class SparseEntry(C.Structure):            # r1cs.h:38-42
    _fields_ = [("row", C.c_uint32), ("col", C.c_uint32), ("value", C.c_uint64)]


class SparseMatrix(C.Structure):           # r1cs.h:49-54
    _fields_ = [("entries", C.POINTER(SparseEntry)), ("n_entries", C.c_size_t),
                ("n_rows", C.c_uint32), ("n_cols", C.c_uint32)]


class R1CSConstraintSystem(C.Structure):   # r1cs.h:62-69
    _fields_ = [("A", SparseMatrix), ("B", SparseMatrix), ("C", SparseMatrix), ("n_vars", C.c_uint32),
                ("n_public_inputs", C.c_uint32), ("n_constraints", C.c_uint32)]


class R1CSWitness(C.Structure):            # r1cs.h:76-79
    _fields_ = [("values", u64p), ("len", C.c_size_t)]


PROFILE_SCALAR_A, PROFILE_RING_B = 0, 1
LweCommitmentP = C.POINTER(LweCommitment)

# name -> (restype, argtypes); every symbol include/lambda_snark_b200.h declares
SIGNATURES = {
    # part 1: drop-in surface
    "ntt_context_create": (C.c_void_p, [C.c_uint64, C.c_uint32]),
    "ntt_context_free": (None, [C.c_void_p]),
    "ntt_forward": (C.c_int, [C.c_void_p, u64p, C.c_uint32]),
    "ntt_inverse": (C.c_int, [C.c_void_p, u64p, C.c_uint32]),
    "ntt_mul_pointwise": (None, [C.c_void_p, u64p, u64p, u64p, C.c_uint32]),
    "lwe_context_create": (C.c_void_p, [C.POINTER(PublicParams)]),
    "lwe_context_free": (None, [C.c_void_p]),
    "lwe_commit": (LweCommitmentP, [C.c_void_p, u64p, C.c_size_t, C.c_uint64]),
    "lwe_commitment_free": (None, [LweCommitmentP]),
    "lwe_commitment_clone": (LweCommitmentP, [LweCommitmentP]),
    "lwe_verify_opening": (C.c_int, [C.c_void_p, LweCommitmentP, u64p, C.c_size_t, C.POINTER(LweOpening)]),
    "lwe_linear_combine": (LweCommitmentP, [C.c_void_p, C.POINTER(LweCommitmentP), u64p, C.c_size_t]),
    "sample_gaussian": (C.c_int, [u64p, C.c_size_t, C.c_double]),
    "lambda_snark_r1cs_create": (C.c_uint32, [C.POINTER(SparseMatrix)] * 3 + [C.c_uint64, C.POINTER(C.c_void_p)]),
    "lambda_snark_r1cs_validate_witness": (C.c_uint32, [C.c_void_p, C.POINTER(R1CSWitness), C.POINTER(C.c_bool)]),
    "lambda_snark_r1cs_free": (None, [C.c_void_p]),
    "lambda_snark_r1cs_num_constraints": (C.c_uint32, [C.c_void_p]),
    "lambda_snark_r1cs_num_variables": (C.c_uint32, [C.c_void_p]),
    "export_vk_to_lean": (C.c_int, [C.POINTER(R1CSConstraintSystem), C.POINTER(PublicParams), C.c_char_p, C.c_size_t]),
    "export_params_to_lean": (C.c_int, [C.POINTER(PublicParams), C.c_char_p, C.c_size_t]),
    "export_seal_context_to_lean": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    "export_seal_pubkey_to_lean": (C.c_int, [C.c_void_p, C.c_char_p, C.c_size_t]),
    # part 2: batched / device extensions
    "lsr_device_count": (C.c_int, []),
    "lsr_set_device": (C.c_int, [C.c_int]),
    "lsr_version": (C.c_char_p, []),
    "lsr_last_error": (C.c_char_p, []),
    "lsr_measure_imad_peak": (C.c_int, [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "lsr_measure_fp64_peak": (C.c_int, [C.POINTER(C.c_double)]),
    "lsr_ntt_modulus": (C.c_uint64, [C.c_void_p]),
    "lsr_ntt_degree": (C.c_uint32, [C.c_void_p]),
    "lsr_ntt_root": (C.c_uint64, [C.c_void_p]),
    "lsr_ntt_device": (C.c_int, [C.c_void_p]),
    "ntt_forward_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ntt_inverse_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ntt_mul_pointwise_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "lsr_ntt_forward_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lsr_ntt_inverse_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lsr_ntt_mul_pointwise_device": (C.c_int, [C.c_void_p] * 4 + [C.c_size_t, C.c_void_p]),
    "lwe_context_create_seeded": (C.c_void_p, [C.POINTER(PublicParams), C.c_char_p]),
    "lsr_lwe_modulus": (C.c_uint64, [C.c_void_p]),
    "lsr_lwe_plain_modulus": (C.c_uint64, [C.c_void_p]),
    "lsr_lwe_delta": (C.c_uint64, [C.c_void_p]),
    "lsr_lwe_commitment_words": (C.c_size_t, [C.c_void_p]),
    "lsr_lwe_copy_matrix": (C.c_int, [C.c_void_p, u64p]),
    "lsr_lwe_set_commit_path": (C.c_int, [C.c_void_p, C.c_int]),
    "lsr_reference_root_of_unity": (C.c_uint64, [C.c_uint64, C.c_uint32]),
    "lsr_cyclic_ntt_context_create": (C.c_void_p, [C.c_uint64, C.c_uint32, C.c_uint64]),
    "lsr_cyclic_ntt_forward": (C.c_int, [C.c_void_p, u64p, C.c_size_t]),
    "lsr_cyclic_ntt_inverse": (C.c_int, [C.c_void_p, u64p, C.c_size_t]),
    "lsr_r1cs_quotient": (C.c_int, [C.c_void_p, u64p, C.c_size_t, C.c_uint64, u64p, C.c_size_t, C.POINTER(C.c_size_t)]),
    "lsr_r1cs_quotient_batch": (C.c_int, [C.c_void_p, u64p, C.c_size_t, C.c_size_t, C.c_uint64, u64p, C.POINTER(C.c_int)]),
    "lsr_prover_quotient_chunks": (C.c_size_t, [C.c_void_p, C.c_void_p]),
    "lsr_prover_quotient_planes": (C.c_uint32, [C.c_void_p, C.c_void_p]),
    "lsr_lwe_set_strict_messages": (C.c_int, [C.c_void_p, C.c_int]),
    "lsr_lwe_lincomb_budget": (C.c_uint64, [C.c_void_p]),
    "lsr_lwe_message_planes": (C.c_uint32, [C.c_void_p, C.c_uint64]),
    "lsr_lwe_commit_digits_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32,
                                                     C.c_void_p, C.c_void_p]),
    "lsr_prover_commit_quotient": (C.c_int, [C.c_void_p, C.c_void_p, u64p, C.c_size_t, C.c_size_t, C.c_uint64, u64p,
                                            C.c_size_t, C.c_size_t, u64p, C.POINTER(C.c_int)]),
    "lsr_prover_commit_quotient_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t,
                                                   C.c_uint64, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                                   C.POINTER(C.c_int)]),
    "lsr_fs_challenge_batch": (C.c_int, [u64p, C.c_size_t, u64p, C.c_size_t, C.c_size_t, C.c_uint64, C.c_int, u64p, u64p]),
    "lsr_fs_challenge_batch_device": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint64,
                                               C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]),
    "lsr_poly_eval_batch": (C.c_int, [C.c_uint64, u64p, C.c_size_t, C.c_size_t, u64p, C.c_size_t, u64p]),
    "lsr_prove_r1cs_batch": (C.c_int, [C.c_void_p, C.c_void_p, u64p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_uint64, u64p,
                                      u64p, u64p, u64p, u64p, C.POINTER(C.c_int)]),
    "lsr_verify_r1cs_batch": (C.c_int, [C.c_uint64, C.c_uint64, u64p, C.c_size_t, u64p, C.c_size_t, u64p, u64p, C.c_size_t,
                                       C.POINTER(C.c_int)]),
    "lsr_host_alloc": (C.c_void_p, [C.c_size_t]),
    "lsr_host_free": (None, [C.c_void_p]),
    "lsr_device_alloc": (C.c_void_p, [C.c_size_t]),
    "lsr_device_free": (None, [C.c_void_p]),
    "lsr_peer_export": (C.c_int, [C.c_void_p, C.c_char_p]),
    "lsr_peer_open": (C.c_void_p, [C.c_char_p]),
    "lsr_peer_close": (C.c_int, [C.c_void_p]),
    "lsr_ntt_set_arith": (C.c_int, [C.c_void_p, C.c_int]),
    "lsr_ntt_arith": (C.c_int, [C.c_void_p]),
    "lsr_lwe_set_arith": (C.c_int, [C.c_void_p, C.c_int]),
    "lwe_commit_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "lsr_lwe_commit_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t,
                                               C.c_void_p, C.c_void_p]),
    "lsr_lwe_verify_opening_batch_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                                      C.c_void_p, C.c_void_p]),
    "lwe_verify_opening_batch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t,
                                            C.POINTER(C.c_int)]),
    "lsr_sample_gaussian_seeded": (C.c_int, [u64p, C.c_size_t, C.c_double, C.c_char_p]),
    "lsr_cdt_magnitude_device": (C.c_int, [C.c_double, u64p, C.c_size_t, C.POINTER(C.c_uint32), C.c_int]),
    "lsr_cdt_timing_device": (C.c_int, [C.c_double, u64p, C.c_size_t, C.c_int, C.POINTER(C.c_uint32), u64p]),
    "lsr_goldilocks_probe_device": (C.c_int, [u64p, u64p, C.c_size_t, u64p]),
    "lsr_lwe_sample_se": (C.c_int, [C.c_void_p, C.c_uint64, i64p, i64p]),
    "lsr_lwe_commit_explicit": (C.c_int, [C.c_void_p, u64p, C.c_size_t, i64p, i64p, C.c_size_t, u64p]),
    "lsr_lwe_commit_explicit_device": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                                                 C.c_void_p, C.c_void_p]),
}

_lib = None


def load(path: Path | None = None) -> C.CDLL:
    """dlopen the C-ABI library and attach the prototypes.  Raises if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = Path(path) if path else LIB_PATH
    if not p.exists():
        raise FileNotFoundError(
            f"{p} is missing: build it with `python -m lambda_snark_r_b200._build` "
            "(the CUDA extension is mandatory; there is no CPU fallback)")
    lib = C.CDLL(str(p))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if a declared symbol is not exported
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib
