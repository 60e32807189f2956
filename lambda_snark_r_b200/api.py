"""Host-side mirror of the reference's safe wrappers over the C ABI.

The reference's host side is Rust (rust-api/lambda-snark/src/{context,commitment,
opening}.rs); there is no Rust toolchain in this image, so the same thin RAII
layer is written in Python with the same names, argument meaning and error
behaviour:

    LweContext            context.rs:14-76     (validate, pack PublicParams, create/free)
    Commitment            commitment.rs:13-110 (new, linear_combine, as_bytes, clone, drop)
    verify_opening_with_context, generate_opening, Opening
                          opening.rs:104-222
    NttContext            the ntt_* FFI (bound by lambda-snark-sys, no safe Rust wrapper)

All arithmetic happens in the CUDA library; numpy / torch only carry buffers.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import capi
from .capi import LweCommitment, LweOpening, PublicParams, u64p


class LambdaSnarkError(RuntimeError):
    pass


def _lib():
    return capi.load()


def _u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _p(a: np.ndarray):
    return a.ctypes.data_as(u64p)


def _vp(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def last_error() -> str:
    return (_lib().lsr_last_error() or b"").decode()


def device_count() -> int:
    return int(_lib().lsr_device_count())


def set_device(dev: int) -> None:
    if _lib().lsr_set_device(dev) != 0:
        raise LambdaSnarkError(f"lsr_set_device({dev}) failed: {last_error()}")


# --------------------------------------------------------------------------- NTT
class NttContext:
    """ntt_context_create / ntt_context_free (ntt.h:34-41)."""

    def __init__(self, q: int, n: int):
        self._h = _lib().ntt_context_create(q, n)
        if not self._h:
            raise LambdaSnarkError(f"ntt_context_create({q}, {n}) returned NULL")
        self.q, self.n = q, n

    def close(self) -> None:
        if getattr(self, "_h", None):
            try:
                _lib().ntt_context_free(self._h)
            except TypeError:        # interpreter shutdown: module globals already cleared
                pass
            self._h = None

    __del__ = close

    @property
    def handle(self):
        return self._h

    @property
    def root(self) -> int:
        return int(_lib().lsr_ntt_root(self._h))

    def set_arith(self, arith: int) -> None:
        """0 auto, 1 u64 butterflies only, 2 require the FP64-pipe butterflies (q < 2^45)."""
        if _lib().lsr_ntt_set_arith(self._h, arith) != 0:
            raise LambdaSnarkError("lsr_ntt_set_arith failed")

    @property
    def arith(self) -> int:
        """Policy in effect: 1 = u64 Shoup butterflies, 2 = FP64 butterflies."""
        return int(_lib().lsr_ntt_arith(self._h))

    # single-polynomial drop-in calls (in place on a copy)
    def forward(self, coeffs) -> np.ndarray:
        a = _u64(coeffs).copy()
        if _lib().ntt_forward(self._h, _p(a), a.size) != 0:
            raise LambdaSnarkError("ntt_forward returned -1")
        return a

    def inverse(self, evals) -> np.ndarray:
        a = _u64(evals).copy()
        if _lib().ntt_inverse(self._h, _p(a), a.size) != 0:
            raise LambdaSnarkError("ntt_inverse returned -1")
        return a

    def mul_pointwise(self, a, b) -> np.ndarray:
        a, b = _u64(a), _u64(b)
        r = np.zeros_like(a)
        _lib().ntt_mul_pointwise(self._h, _p(r), _p(a), _p(b), a.size)
        return r

    # batched host-memory calls, [batch][n]
    def forward_batch(self, polys, inplace: bool = False) -> np.ndarray:
        a = _u64(polys) if inplace else _u64(polys).copy()
        assert a.size % self.n == 0
        if _lib().ntt_forward_batch(self._h, _vp(a), a.size // self.n) != 0:
            raise LambdaSnarkError(f"ntt_forward_batch failed: {last_error()}")
        return a

    def inverse_batch(self, polys, inplace: bool = False) -> np.ndarray:
        a = _u64(polys) if inplace else _u64(polys).copy()
        assert a.size % self.n == 0
        if _lib().ntt_inverse_batch(self._h, _vp(a), a.size // self.n) != 0:
            raise LambdaSnarkError(f"ntt_inverse_batch failed: {last_error()}")
        return a

    def mul_pointwise_batch(self, a, b) -> np.ndarray:
        a, b = _u64(a), _u64(b)
        r = np.zeros_like(a)
        if _lib().ntt_mul_pointwise_batch(self._h, _vp(r), _vp(a), _vp(b), a.size) != 0:
            raise LambdaSnarkError(f"ntt_mul_pointwise_batch failed: {last_error()}")
        return r

    # device-memory calls: raw pointers (e.g. torch_tensor.data_ptr()) and a cudaStream_t
    def forward_device(self, dptr: int, batch: int, stream: int = 0) -> None:
        if _lib().lsr_ntt_forward_device(self._h, C.c_void_p(dptr), batch, C.c_void_p(stream)) != 0:
            raise LambdaSnarkError(f"lsr_ntt_forward_device failed: {last_error()}")

    def inverse_device(self, dptr: int, batch: int, stream: int = 0) -> None:
        if _lib().lsr_ntt_inverse_device(self._h, C.c_void_p(dptr), batch, C.c_void_p(stream)) != 0:
            raise LambdaSnarkError(f"lsr_ntt_inverse_device failed: {last_error()}")

    def mul_pointwise_device(self, r: int, a: int, b: int, total: int, stream: int = 0) -> None:
        if _lib().lsr_ntt_mul_pointwise_device(self._h, C.c_void_p(r), C.c_void_p(a), C.c_void_p(b), total,
                                               C.c_void_p(stream)) != 0:
            raise LambdaSnarkError(f"lsr_ntt_mul_pointwise_device failed: {last_error()}")


# ------------------------------------------------------------------- commitment
class CyclicNtt(NttContext):
    """Cyclic transform over X^n - 1, natural order in and out: ntt_forward / ntt_inverse of
    rust-api/lambda-snark/src/ntt.rs:117-201.  omega = 0 picks the reference's root of unity."""

    def __init__(self, q: int, n: int, omega: int = 0):
        self._h = _lib().lsr_cyclic_ntt_context_create(q, n, omega)
        if not self._h:
            raise LambdaSnarkError(f"lsr_cyclic_ntt_context_create({q}, {n}, {omega}) returned NULL")
        self.q, self.n = q, n

    def forward_natural(self, polys) -> np.ndarray:
        a = np.ascontiguousarray(_u64(polys)).copy()
        if _lib().lsr_cyclic_ntt_forward(self._h, _p(a), a.size // self.n) != 0:
            raise LambdaSnarkError("lsr_cyclic_ntt_forward returned -1")
        return a

    def inverse_natural(self, evals) -> np.ndarray:
        a = np.ascontiguousarray(_u64(evals)).copy()
        if _lib().lsr_cyclic_ntt_inverse(self._h, _p(a), a.size // self.n) != 0:
            raise LambdaSnarkError("lsr_cyclic_ntt_inverse returned -1")
        return a


def reference_root_of_unity(q: int, n: int) -> int:
    return int(_lib().lsr_reference_root_of_unity(q, n))


class R1CS:
    """lambda_snark_r1cs_* handle (lambda-snark-core/src/r1cs.rs:121-141) plus the GPU quotient pipeline
    that replaces R1CS::compute_quotient_poly (lambda-snark/src/r1cs.rs:474-503) on its NTT path.
    Matrices are lists of (row, col, value)."""

    def __init__(self, rows: int, cols: int, A, B, C, modulus: int):
        from .capi import SparseEntry, SparseMatrix
        import ctypes as C_
        self._keep = []
        mats = []
        for entries in (A, B, C):
            arr = (SparseEntry * max(len(entries), 1))()
            for i, (r, c, v) in enumerate(entries):
                arr[i].row, arr[i].col, arr[i].value = r, c, v % (1 << 64)
            self._keep.append(arr)
            mats.append(SparseMatrix(C_.cast(arr, C_.POINTER(SparseEntry)), len(entries), rows, cols))
        h = C_.c_void_p()
        rc = _lib().lambda_snark_r1cs_create(C_.byref(mats[0]), C_.byref(mats[1]), C_.byref(mats[2]), modulus, C_.byref(h))
        if rc != 0:
            raise LambdaSnarkError(f"lambda_snark_r1cs_create failed with code {rc}")
        self._h = h
        self.rows, self.cols, self.modulus = rows, cols, modulus

    @classmethod
    def from_arrays(cls, rows: int, cols: int, A, B, C, modulus: int) -> "R1CS":
        """Same handle from numpy triples (row[], col[], value[]) per matrix -- no per-entry Python work, for
        circuits with millions of constraints.  The buffer has the SparseEntry layout of r1cs.h:38-43."""
        from .capi import SparseEntry, SparseMatrix
        import ctypes as C_
        self = cls.__new__(cls)
        self._keep = []
        mats = []
        dt = np.dtype([("row", np.uint32), ("col", np.uint32), ("value", np.uint64)])
        assert dt.itemsize == C_.sizeof(SparseEntry)
        for r, c, v in (A, B, C):
            arr = np.zeros(max(len(r), 1), dtype=dt)
            arr["row"][: len(r)], arr["col"][: len(r)], arr["value"][: len(r)] = r, c, v
            self._keep.append(arr)
            mats.append(SparseMatrix(arr.ctypes.data_as(C_.POINTER(SparseEntry)), len(r), rows, cols))
        h = C_.c_void_p()
        rc = _lib().lambda_snark_r1cs_create(C_.byref(mats[0]), C_.byref(mats[1]), C_.byref(mats[2]), modulus, C_.byref(h))
        if rc != 0:
            raise LambdaSnarkError(f"lambda_snark_r1cs_create failed with code {rc}")
        self._h = h
        self.rows, self.cols, self.modulus = rows, cols, modulus
        return self

    def quotient_chunks(self, ctx: "LweContext") -> int:
        """Commitment units per witness: ring elements of the quotient x digit planes."""
        return int(_lib().lsr_prover_quotient_chunks(self._h, ctx.as_ptr()))

    def quotient_planes(self, ctx: "LweContext") -> int:
        return int(_lib().lsr_prover_quotient_planes(self._h, ctx.as_ptr()))

    def commit_quotient(self, ctx: "LweContext", witnesses, seeds, chunk_lo: int = 0, chunk_hi: int | None = None,
                        omega: int = 0):
        """Commitment phase of prove_r1cs (lib.rs:747-757) for quotients longer than one ring element:
        witnesses [count][cols], seeds [count][chunks] (global unit index; a unit is one digit plane of one ring
        element of Q, sharding.message_digits order) -> (containers [count][chunk_hi - chunk_lo][1 + k n],
        status [count]).  Q stays on the device."""
        import ctypes as C_
        w = np.ascontiguousarray(_u64(witnesses)).reshape(-1, self.cols)
        chunks = self.quotient_chunks(ctx)
        chunk_hi = chunks if chunk_hi is None else chunk_hi
        sd = np.ascontiguousarray(_u64(seeds)).reshape(w.shape[0], chunks)
        out = np.zeros((w.shape[0], chunk_hi - chunk_lo, ctx.words), dtype=np.uint64)
        status = np.zeros(w.shape[0], dtype=np.int32)
        rc = _lib().lsr_prover_commit_quotient(self._h, ctx.as_ptr(), _p(w), self.cols, w.shape[0], omega, _p(sd),
                                               chunk_lo, chunk_hi, _p(out), status.ctypes.data_as(C_.POINTER(C_.c_int)))
        if rc != 0:
            raise LambdaSnarkError(f"lsr_prover_commit_quotient failed with code {rc}: {last_error()}")
        return out, status

    def commit_quotient_device(self, ctx: "LweContext", witnesses_ptr: int, count: int, seeds_ptr: int, out_ptr: int,
                               chunk_lo: int, chunk_hi: int, omega: int = 0) -> np.ndarray:
        import ctypes as C_
        status = np.zeros(count, dtype=np.int32)
        rc = _lib().lsr_prover_commit_quotient_device(self._h, ctx.as_ptr(), witnesses_ptr, self.cols, count, omega,
                                                      seeds_ptr, chunk_lo, chunk_hi, out_ptr,
                                                      status.ctypes.data_as(C_.POINTER(C_.c_int)))
        if rc != 0:
            raise LambdaSnarkError(f"lsr_prover_commit_quotient_device failed with code {rc}: {last_error()}")
        return status

    def prove_batch(self, ctx: "LweContext", witnesses, n_public: int, seeds, omega: int = 0):
        """prove_r1cs (lib.rs:747-809) for a batch of witnesses, m <= ring_degree, NTT path.  Returns a dict of
        arrays: containers [count][words], challenges [count][2] (alpha, beta), hashes [count][2][32] bytes,
        evals [count][8] (ProofR1CS field order), status [count]."""
        import ctypes as C_
        w = np.ascontiguousarray(_u64(witnesses)).reshape(-1, self.cols)
        count = w.shape[0]
        sd = np.ascontiguousarray(_u64(seeds)).reshape(count)
        out = {"containers": np.zeros((count, ctx.words), dtype=np.uint64), "challenges": np.zeros((count, 2), dtype=np.uint64),
               "hashes": np.zeros((count, 2, 4), dtype=np.uint64), "evals": np.zeros((count, 8), dtype=np.uint64),
               "status": np.zeros(count, dtype=np.int32)}
        rc = _lib().lsr_prove_r1cs_batch(self._h, ctx.as_ptr(), _p(w), self.cols, count, n_public, omega, _p(sd),
                                         _p(out["containers"]), _p(out["challenges"]), _p(out["hashes"]), _p(out["evals"]),
                                         out["status"].ctypes.data_as(C_.POINTER(C_.c_int)))
        if rc != 0:
            raise LambdaSnarkError(f"lsr_prove_r1cs_batch failed with code {rc}: {last_error()}")
        out["hashes"] = out["hashes"].view(np.uint8).reshape(count, 2, 32)
        return out

    def close(self) -> None:
        if getattr(self, "_h", None):
            try:
                _lib().lambda_snark_r1cs_free(self._h)
            except TypeError:
                pass
            self._h = None

    __del__ = close

    def quotient(self, witness, omega: int = 0) -> np.ndarray:
        """compute_quotient_poly: coefficients with trailing zeros removed; raises if the witness is invalid."""
        import ctypes as C_
        w = np.ascontiguousarray(_u64(witness))
        out = np.zeros(max(self.rows, 1), dtype=np.uint64)
        n = C_.c_size_t(0)
        rc = _lib().lsr_r1cs_quotient(self._h, _p(w), w.size, omega, _p(out), out.size, C_.byref(n))
        if rc != 0:
            raise LambdaSnarkError(f"lsr_r1cs_quotient failed with code {rc}: {last_error()}")
        return out[: n.value].copy()

    def quotient_batch(self, witnesses, omega: int = 0):
        """witnesses [count][cols] -> (Q [count][m] zero-padded, status [count])."""
        import ctypes as C_
        w = np.ascontiguousarray(_u64(witnesses)).reshape(-1, self.cols)
        out = np.zeros((w.shape[0], self.rows), dtype=np.uint64)
        status = np.zeros(w.shape[0], dtype=np.int32)
        rc = _lib().lsr_r1cs_quotient_batch(self._h, _p(w), self.cols, w.shape[0], omega, _p(out),
                                            status.ctypes.data_as(C_.POINTER(C_.c_int)))
        if rc != 0:
            raise LambdaSnarkError(f"lsr_r1cs_quotient_batch failed with code {rc}: {last_error()}")
        return out, status


@dataclass
class Params:
    """lambda-snark-core Params/Profile::RingB (lambda-snark-core/src/lib.rs:129-196)."""
    n: int = 4096
    k: int = 2
    q: int = 17592169062401
    sigma: float = 3.19
    security_level: int = 128

    def validate(self) -> None:          # lib.rs validation: n pow2, k>0, q>=2^24, sigma>=3.0
        if self.n <= 0 or self.n & (self.n - 1):
            raise LambdaSnarkError("invalid params: n must be a power of two")
        if self.k <= 0:
            raise LambdaSnarkError("invalid params: k must be positive")
        if self.q < (1 << 24):
            raise LambdaSnarkError("invalid params: q must be at least 2^24")
        if not self.sigma >= 3.0:
            raise LambdaSnarkError("invalid params: sigma must be at least 3.0")

    def to_ffi(self) -> PublicParams:
        return PublicParams(capi.PROFILE_RING_B, self.security_level, self.q, self.n, self.k, self.sigma)


class LweContext:
    """context.rs:14-76.  `seed32` (extension) makes the context reproducible."""

    def __init__(self, params: Params, seed32: bytes | None = None, validate: bool = True):
        if validate:
            params.validate()
        self.params = params
        pp = params.to_ffi()
        if seed32 is None:
            self._h = _lib().lwe_context_create(C.byref(pp))
        else:
            assert len(seed32) == 32
            self._h = _lib().lwe_context_create_seeded(C.byref(pp), seed32)
        if not self._h:
            raise LambdaSnarkError(f"lwe_context_create returned NULL: {last_error()}")
        lib = _lib()
        self.q = int(lib.lsr_lwe_modulus(self._h))            # ring modulus actually in use
        self.p = int(lib.lsr_lwe_plain_modulus(self._h))
        self.delta = int(lib.lsr_lwe_delta(self._h))
        self.words = int(lib.lsr_lwe_commitment_words(self._h))
        self.n, self.k = params.n, params.k

    def modulus(self) -> int:            # context.rs: ctx.modulus() is the caller's field modulus
        return self.params.q

    def as_ptr(self):
        return self._h

    def close(self) -> None:
        if getattr(self, "_h", None):
            try:
                _lib().lwe_context_free(self._h)
            except TypeError:
                pass
            self._h = None

    __del__ = close

    def set_arith(self, arith: int) -> None:
        """Arithmetic of the NTT butterflies inside the commitment kernels (see NttContext.set_arith)."""
        if _lib().lsr_lwe_set_arith(self._h, arith) != 0:
            raise LambdaSnarkError("lsr_lwe_set_arith failed")

    def set_commit_path(self, path: int) -> None:
        if _lib().lsr_lwe_set_commit_path(self._h, path) != 0:
            raise LambdaSnarkError("lsr_lwe_set_commit_path failed")

    def set_strict_messages(self, strict: bool) -> None:
        """Host-pointer calls reject message words >= p instead of reducing them (lsr_lwe_set_strict_messages)."""
        if _lib().lsr_lwe_set_strict_messages(self._h, 1 if strict else 0) != 0:
            raise LambdaSnarkError("lsr_lwe_set_strict_messages failed")

    def lincomb_budget(self) -> int:
        """Largest sum of |centred coefficients| lwe_linear_combine accepts."""
        return int(_lib().lsr_lwe_lincomb_budget(self._h))

    def message_planes(self, modulus: int) -> int:
        """Base-p digit planes that bind a whole element of Z_modulus (0: more than 4)."""
        return int(_lib().lsr_lwe_message_planes(self._h, modulus))

    def commit_digits_batch_device(self, msgs_ptr: int, msg_len: int, seeds_ptr: int, count: int, planes: int,
                                   out_ptr: int, stream: int = 0) -> None:
        if _lib().lsr_lwe_commit_digits_batch_device(self._h, C.c_void_p(msgs_ptr), msg_len, C.c_void_p(seeds_ptr), count,
                                                     planes, C.c_void_p(out_ptr), C.c_void_p(stream)) != 0:
            raise LambdaSnarkError(f"lsr_lwe_commit_digits_batch_device failed: {last_error()}")

    def matrix(self) -> np.ndarray:
        out = np.zeros((self.k, self.k, self.n), dtype=np.uint64)
        if _lib().lsr_lwe_copy_matrix(self._h, _p(out)) != 0:
            raise LambdaSnarkError("lsr_lwe_copy_matrix failed")
        return out

    def sample_se(self, seed: int):
        s = np.zeros((self.k, self.n), dtype=np.int64)
        e = np.zeros((self.k, self.n), dtype=np.int64)
        if _lib().lsr_lwe_sample_se(self._h, seed, s.ctypes.data_as(capi.i64p), e.ctypes.data_as(capi.i64p)) != 0:
            raise LambdaSnarkError("lsr_lwe_sample_se failed")
        return s, e

    def commit_explicit(self, messages, s, e) -> np.ndarray:
        """Explicit mode: containers of t = A*s + e + Delta*m for caller-supplied s, e.
        messages [count][msg_len], s / e [count][k][n] int64 -> [count][words]."""
        m = _u64(messages)
        count, msg_len = m.shape
        s = np.ascontiguousarray(s, dtype=np.int64).reshape(count, self.k, self.n)
        e = np.ascontiguousarray(e, dtype=np.int64).reshape(count, self.k, self.n)
        out = np.zeros((count, self.words), dtype=np.uint64)
        rc = _lib().lsr_lwe_commit_explicit(self._h, _p(m) if m.size else None, msg_len, s.ctypes.data_as(capi.i64p),
                                            e.ctypes.data_as(capi.i64p), count, _p(out))
        if rc != 0:
            raise LambdaSnarkError(f"lsr_lwe_commit_explicit failed: {last_error()}")
        return out

    # batched extension: messages [count][msg_len], seeds [count] -> containers [count][words]
    def commit_batch(self, messages, seeds, out: np.ndarray | None = None) -> np.ndarray:
        m = _u64(messages)
        count, msg_len = m.shape
        s = _u64(seeds)
        assert s.size == count
        if out is None:
            out = np.zeros((count, self.words), dtype=np.uint64)
        if _lib().lwe_commit_batch(self._h, _vp(m), msg_len, _vp(s), count, _vp(out)) != 0:
            raise LambdaSnarkError(f"lwe_commit_batch failed: {last_error()}")
        return out

    def commit_batch_ptr(self, msgs_ptr: int, msg_len: int, seeds_ptr: int, count: int, out_ptr: int) -> None:
        """Host pointers (e.g. pinned torch tensors)."""
        if _lib().lwe_commit_batch(self._h, C.c_void_p(msgs_ptr), msg_len, C.c_void_p(seeds_ptr), count,
                                   C.c_void_p(out_ptr)) != 0:
            raise LambdaSnarkError(f"lwe_commit_batch failed: {last_error()}")

    def commit_batch_device(self, msgs_ptr: int, msg_len: int, seeds_ptr: int, count: int, out_ptr: int,
                            stream: int = 0) -> None:
        if _lib().lsr_lwe_commit_batch_device(self._h, C.c_void_p(msgs_ptr), msg_len, C.c_void_p(seeds_ptr), count,
                                              C.c_void_p(out_ptr), C.c_void_p(stream)) != 0:
            raise LambdaSnarkError(f"lsr_lwe_commit_batch_device failed: {last_error()}")

    def verify_batch_device(self, comm_ptr: int, msgs_ptr: int, msg_len: int, count: int, diff_ptr: int, invalid_ptr: int,
                            stream: int = 0) -> None:
        """lsr_lwe_verify_opening_batch_device: opening i verifies iff diff[i] == 0 and invalid[i] == 0 (asynchronous)."""
        if _lib().lsr_lwe_verify_opening_batch_device(self._h, comm_ptr, msgs_ptr, msg_len, count, diff_ptr, invalid_ptr, stream) != 0:
            raise LambdaSnarkError(f"lsr_lwe_verify_opening_batch_device failed: {last_error()}")

    def verify_batch(self, containers, messages) -> np.ndarray:
        cw = _u64(containers)
        m = _u64(messages)
        count, msg_len = m.shape
        res = np.zeros(count, dtype=np.int32)
        if _lib().lwe_verify_opening_batch(self._h, _vp(cw), _vp(m), msg_len, count,
                                           res.ctypes.data_as(C.POINTER(C.c_int))) != 0:
            raise LambdaSnarkError(f"lwe_verify_opening_batch failed: {last_error()}")
        return res


class Commitment:
    """commitment.rs:13-110 -- owns an `LweCommitment*` returned by the library."""

    def __init__(self, inner):
        self._inner = inner

    @classmethod
    def new(cls, ctx: LweContext, message, seed: int) -> "Commitment":
        # commitment.rs:31-45: reduce each field element mod ctx.modulus(), then lwe_commit
        modulus = ctx.modulus()
        msg = _u64([int(v) % modulus for v in message])
        ptr = _p(msg) if msg.size else C.cast(C.c_void_p(8), u64p)     # non-null dangling, as Vec::as_ptr
        inner = _lib().lwe_commit(ctx.as_ptr(), ptr, msg.size, seed)
        if not inner:
            raise LambdaSnarkError("CommitmentFailed")
        return cls(inner)

    @classmethod
    def linear_combine(cls, ctx: LweContext, commitments, coeffs) -> "Commitment":
        # commitment.rs:48-84
        if len(commitments) == 0:
            raise LambdaSnarkError("no commitments provided")
        if len(commitments) != len(coeffs):
            raise LambdaSnarkError("commitments/coeffs length mismatch")
        modulus = ctx.modulus()
        ptrs = (capi.LweCommitmentP * len(commitments))(*[c._inner for c in commitments])
        cf = _u64([int(v) % modulus for v in coeffs])
        inner = _lib().lwe_linear_combine(ctx.as_ptr(), ptrs, _p(cf), len(commitments))
        if not inner:
            raise LambdaSnarkError("CommitmentFailed")
        return cls(inner)

    def as_bytes(self) -> np.ndarray:
        """commitment.rs:87-93: the u64 words the Fiat-Shamir transcript hashes."""
        raw = self._inner.contents
        return np.ctypeslib.as_array(raw.data, shape=(raw.len,))

    def as_ffi_ptr(self):
        return self._inner

    def clone(self) -> "Commitment":
        inner = _lib().lwe_commitment_clone(self._inner)
        if not inner:
            raise LambdaSnarkError("lwe_commitment_clone returned null")
        return Commitment(inner)

    def close(self) -> None:
        if getattr(self, "_inner", None):
            try:
                _lib().lwe_commitment_free(self._inner)
            except TypeError:
                pass
            self._inner = None

    __del__ = close


def verify_commitment(ctx: LweContext, commitment: Commitment, message, randomness=None) -> int:
    """Raw lwe_verify_opening: 1 / 0 / -1."""
    msg = _u64(message)
    mp = _p(msg) if msg.size else C.cast(C.c_void_p(8), u64p)
    if randomness is None:
        opening = LweOpening(None, 0)
    else:
        r = _u64(randomness)
        opening = LweOpening(_p(r), r.size)
    return int(_lib().lwe_verify_opening(ctx.as_ptr(), commitment.as_ffi_ptr(), mp, msg.size, C.byref(opening)))


# ------------------------------------------------------------------ openings
@dataclass
class Opening:                              # opening.rs:20-60
    evaluation: int
    witness: list


def _horner(coeffs, alpha: int, q: int) -> int:   # polynomial.rs:97-113
    acc = 0
    for c in reversed(list(coeffs)):
        acc = (acc * alpha + int(c)) % q
    return acc


def generate_opening(coeffs, alpha: int, randomness: int, modulus: int) -> Opening:
    """opening.rs:104-115: y = f(alpha); witness = [randomness, coeffs...]."""
    cs = [int(c) % modulus for c in coeffs]
    return Opening(_horner(cs, alpha % modulus, modulus), [randomness] + cs)


def verify_opening_with_context(commitment: Commitment, alpha: int, opening: Opening, modulus: int,
                                ctx: LweContext) -> bool:
    """opening.rs:160-222."""
    if opening.evaluation >= modulus:
        return False
    if len(opening.witness) < 2:
        return False
    coeffs = [int(c) % modulus for c in opening.witness[1:]]
    if (_horner(coeffs, alpha % modulus, modulus) - opening.evaluation) % modulus != 0:
        return False
    return verify_commitment(ctx, commitment, coeffs, [opening.witness[0]]) == 1


# -------------------------------------------------------------------- sampler
def sample_gaussian(length: int, sigma: float, seed32: bytes | None = None) -> np.ndarray:
    """sample_gaussian (utils.h:27); returns int64 samples.  Raises on -1."""
    out = np.zeros(max(length, 1), dtype=np.uint64)
    if seed32 is None:
        rc = _lib().sample_gaussian(_p(out), length, float(sigma))
    else:
        rc = _lib().lsr_sample_gaussian_seeded(_p(out), length, float(sigma), seed32)
    if rc != 0:
        raise LambdaSnarkError("sample_gaussian returned -1")
    return out[:length].view(np.int64)


def fs_challenge_batch(public_inputs, containers, modulus: int, chain: bool = True):
    """Challenge::derive (challenge.rs:102-134) for a batch on the device: public_inputs [count][n_public],
    containers [count][words] -> (challenges [count][2], hashes [count][2][32] bytes)."""
    c = np.ascontiguousarray(_u64(containers))
    c = c.reshape(1, -1) if c.ndim == 1 else c
    count = c.shape[0]
    pub = np.ascontiguousarray(_u64(public_inputs)).reshape(count, -1)
    ch = np.zeros((count, 2), dtype=np.uint64)
    hs = np.zeros((count, 2, 4), dtype=np.uint64)
    rc = _lib().lsr_fs_challenge_batch(_p(pub), pub.shape[1], _p(c), c.shape[1], count, modulus, 1 if chain else 0, _p(ch), _p(hs))
    if rc != 0:
        raise LambdaSnarkError(f"lsr_fs_challenge_batch failed: {last_error()}")
    return ch, hs.view(np.uint8).reshape(count, 2, 32)


def poly_eval_batch(coeffs, points, modulus: int) -> np.ndarray:
    """eval_poly (r1cs.rs:362-373): coeffs [polys][len], points [polys][npts] -> [polys][npts]."""
    c = np.ascontiguousarray(_u64(coeffs))
    c = c.reshape(1, -1) if c.ndim == 1 else c
    pts = np.ascontiguousarray(_u64(points)).reshape(c.shape[0], -1)
    out = np.zeros(pts.shape, dtype=np.uint64)
    rc = _lib().lsr_poly_eval_batch(modulus, _p(c), c.shape[1], c.shape[0], _p(pts), pts.shape[1], _p(out))
    if rc != 0:
        raise LambdaSnarkError(f"lsr_poly_eval_batch failed: {last_error()}")
    return out


def verify_r1cs_batch(n_constraints: int, modulus: int, public_inputs, containers, challenges, evals) -> np.ndarray:
    """verify_r1cs (lib.rs:1016-1082) for a batch of proofs of one circuit on the NTT path; 1 accept / 0 reject."""
    import ctypes as C_
    c = np.ascontiguousarray(_u64(containers))
    c = c.reshape(1, -1) if c.ndim == 1 else c
    count = c.shape[0]
    pub = np.ascontiguousarray(_u64(public_inputs)).reshape(count, -1)
    ch = np.ascontiguousarray(_u64(challenges)).reshape(count, 2)
    ev = np.ascontiguousarray(_u64(evals)).reshape(count, 8)
    res = np.zeros(count, dtype=np.int32)
    rc = _lib().lsr_verify_r1cs_batch(n_constraints, modulus, _p(pub), pub.shape[1], _p(c), c.shape[1], _p(ch), _p(ev), count,
                                      res.ctypes.data_as(C_.POINTER(C_.c_int)))
    if rc != 0:
        raise LambdaSnarkError(f"lsr_verify_r1cs_batch failed: {last_error()}")
    return res
