"""ctypes front-end for the CPU ORACLE (oracle/lsr_oracle.c) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.  The product package
(lambda_snark_r_b200) never does.

Also holds a few pure-Python (big-int) restatements used to cross-check the C
oracle on small cases: the closed form of the SEAL forward NTT
(out[i] = f(psi^(2*brv(i)+1))) and schoolbook negacyclic
multiplication.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_BUILD = _HERE / "_build"
_REF = _HERE / "_ref"

u64p = C.POINTER(C.c_uint64)
i64p = C.POINTER(C.c_int64)


def build(native: bool = False, quiet: bool = True) -> Path:
    """Compile the C oracle (and oracle/_ref when /root/reference is present)."""
    target = "_build/liblsr_oracle_native.so" if native else "all"
    out = subprocess.run(["make", "-C", str(_HERE), target], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("oracle build failed:\n" + out.stdout + out.stderr)
    if not quiet:
        print(out.stdout)
    return _BUILD / ("liblsr_oracle_native.so" if native else "liblsr_oracle.so")


def _np_u64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(u64p)


class _Lib:
    def __init__(self, path: Path):
        lib = C.CDLL(str(path))
        self.lib = lib
        lib.lsro_mulmod.restype = C.c_uint64
        lib.lsro_mulmod.argtypes = [C.c_uint64] * 3
        lib.lsro_powmod.restype = C.c_uint64
        lib.lsro_powmod.argtypes = [C.c_uint64] * 3
        lib.lsro_is_prime.restype = C.c_int
        lib.lsro_is_prime.argtypes = [C.c_uint64]
        lib.lsro_min_primitive_root.restype = C.c_uint64
        lib.lsro_min_primitive_root.argtypes = [C.c_uint64, C.c_uint64]
        lib.lsro_ntt_create.restype = C.c_void_p
        lib.lsro_ntt_create.argtypes = [C.c_uint64, C.c_uint32]
        lib.lsro_ntt_free.argtypes = [C.c_void_p]
        lib.lsro_ntt_psi.restype = C.c_uint64
        lib.lsro_ntt_psi.argtypes = [C.c_void_p]
        lib.lsro_ntt_table.restype = u64p
        lib.lsro_ntt_table.argtypes = [C.c_void_p, C.c_int]
        lib.lsro_ntt_forward.restype = C.c_int
        lib.lsro_ntt_forward.argtypes = [C.c_void_p, u64p, C.c_uint32]
        lib.lsro_ntt_inverse.restype = C.c_int
        lib.lsro_ntt_inverse.argtypes = [C.c_void_p, u64p, C.c_uint32]
        lib.lsro_ntt_mul_pointwise.argtypes = [C.c_void_p, u64p, u64p, u64p, C.c_uint32]
        lib.lsro_ntt_forward_batch.restype = C.c_int
        lib.lsro_ntt_forward_batch.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_int]
        lib.lsro_ntt_inverse_batch.restype = C.c_int
        lib.lsro_ntt_inverse_batch.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_int]
        lib.lsro_ntt_mul_pointwise_batch.argtypes = [C.c_void_p, u64p, u64p, u64p, C.c_size_t, C.c_int]
        lib.lsro_cdt_build.restype = C.c_size_t
        lib.lsro_cdt_build.argtypes = [C.c_double, u64p, C.c_size_t]
        lib.lsro_cdt_sample.restype = C.c_int64
        lib.lsro_cdt_sample.argtypes = [u64p, C.c_size_t, C.c_uint64, C.c_uint64]
        lib.lsro_sample_gaussian_seeded.restype = C.c_int
        lib.lsro_sample_gaussian_seeded.argtypes = [u64p, C.c_size_t, C.c_double, C.c_char_p]
        lib.lsro_chacha_block.argtypes = [C.POINTER(C.c_uint32)] + [C.c_uint32] * 4 + [C.POINTER(C.c_uint32)]
        lib.lsro_lwe_create.restype = C.c_void_p
        lib.lsro_lwe_create.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_double, C.c_char_p]
        lib.lsro_lwe_free.argtypes = [C.c_void_p]
        for name in ("lsro_lwe_modulus", "lsro_lwe_plain_modulus", "lsro_lwe_delta"):
            getattr(lib, name).restype = C.c_uint64
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.lsro_lwe_words.restype = C.c_uint32
        lib.lsro_lwe_words.argtypes = [C.c_void_p]
        lib.lsro_lwe_matrix.restype = u64p
        lib.lsro_lwe_matrix.argtypes = [C.c_void_p]
        lib.lsro_lwe_trapdoor.restype = u64p
        lib.lsro_lwe_trapdoor.argtypes = [C.c_void_p]
        lib.lsro_lwe_sample_se.argtypes = [C.c_void_p, C.c_uint64, i64p, i64p]
        lib.lsro_lwe_commit.restype = C.c_int
        lib.lsro_lwe_commit.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_uint64, u64p]
        lib.lsro_lwe_commit_explicit.restype = C.c_int
        lib.lsro_lwe_commit_explicit.argtypes = [C.c_void_p, u64p, C.c_size_t, i64p, i64p, u64p]
        lib.lsro_lwe_commit_batch.restype = C.c_int
        lib.lsro_lwe_commit_batch.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p, C.c_size_t, u64p, C.c_int]
        lib.lsro_lwe_verify.restype = C.c_int
        lib.lsro_lwe_verify.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p, C.c_size_t]
        lib.lsro_lwe_linear_combine.restype = C.c_int
        lib.lsro_lwe_linear_combine.argtypes = [C.c_void_p, C.POINTER(u64p), C.POINTER(C.c_size_t), u64p,
                                                C.c_size_t, u64p]
        lib.lsro_max_threads.restype = C.c_int
        for name in ("lsro_cyclic_ntt_forward", "lsro_cyclic_ntt_inverse"):
            getattr(lib, name).restype = C.c_int
            getattr(lib, name).argtypes = [u64p, C.c_size_t, C.c_uint64, C.c_uint64]
        u32pp, u64pp = C.POINTER(C.c_uint32) * 3, u64p * 3
        lib.lsro_r1cs_quotient.restype = C.c_int
        lib.lsro_r1cs_quotient.argtypes = [C.c_size_t, C.c_size_t, u32pp, u32pp, u64pp, C.c_size_t * 3, u64p,
                                           C.c_uint64, C.c_uint64, C.c_uint64, u64p]


_lib_cache: dict[bool, _Lib] = {}


def lib(native: bool = False) -> _Lib:
    if native not in _lib_cache:
        path = _BUILD / ("liblsr_oracle_native.so" if native else "liblsr_oracle.so")
        src_mtime = max((_HERE / f).stat().st_mtime for f in ("lsr_oracle.c", "lsr_oracle_quotient.c", "lsr_oracle.h"))
        if not path.exists() or path.stat().st_mtime < src_mtime:
            build(native=native)
        _lib_cache[native] = _Lib(path)
    return _lib_cache[native]


def max_threads() -> int:
    return int(lib().lib.lsro_max_threads())


# ------------------------------------------------------------------ arithmetic
def mulmod(a: int, b: int, q: int) -> int:
    return int(lib().lib.lsro_mulmod(a, b, q))


def is_prime(q: int) -> bool:
    return bool(lib().lib.lsro_is_prime(q))


def min_primitive_root(q: int, two_n: int) -> int:
    return int(lib().lib.lsro_min_primitive_root(q, two_n))


# ------------------------------------------------------------------------ NTT
class OracleNtt:
    """Mirror of the reference's NttContext (cpp-core/src/ntt.cpp:21-74)."""

    def __init__(self, q: int, n: int, native: bool = False):
        self._l = lib(native).lib
        self._h = self._l.lsro_ntt_create(q, n)
        if not self._h:
            raise ValueError(f"ntt_context_create({q}, {n}) -> NULL")
        self.q, self.n = q, n

    @staticmethod
    def try_create(q: int, n: int):
        try:
            return OracleNtt(q, n)
        except ValueError:
            return None

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.lsro_ntt_free(self._h)
            self._h = None

    @property
    def psi(self) -> int:
        return int(self._l.lsro_ntt_psi(self._h))

    def table(self, which: int) -> np.ndarray:
        p = self._l.lsro_ntt_table(self._h, which)
        return np.ctypeslib.as_array(p, shape=(self.n,)).copy()

    def forward(self, x) -> np.ndarray:
        a = _np_u64(x).copy()
        batch = a.size // self.n
        assert a.size == batch * self.n
        rc = self._l.lsro_ntt_forward_batch(self._h, _ptr(a), batch, 1)
        assert rc == 0
        return a.reshape(np.shape(x))

    def inverse(self, x) -> np.ndarray:
        a = _np_u64(x).copy()
        batch = a.size // self.n
        assert a.size == batch * self.n
        rc = self._l.lsro_ntt_inverse_batch(self._h, _ptr(a), batch, 1)
        assert rc == 0
        return a.reshape(np.shape(x))

    def forward_inplace(self, a: np.ndarray, threads: int = 1) -> None:
        self._l.lsro_ntt_forward_batch(self._h, _ptr(a), a.size // self.n, threads)

    def inverse_inplace(self, a: np.ndarray, threads: int = 1) -> None:
        self._l.lsro_ntt_inverse_batch(self._h, _ptr(a), a.size // self.n, threads)

    def mul_pointwise(self, a, b) -> np.ndarray:
        a = _np_u64(a)
        b = _np_u64(b)
        r = np.empty_like(a)
        self._l.lsro_ntt_mul_pointwise_batch(self._h, _ptr(r), _ptr(a), _ptr(b), a.size, 1)
        return r

    def mul_pointwise_into(self, r, a, b, threads: int = 1) -> None:
        self._l.lsro_ntt_mul_pointwise_batch(self._h, _ptr(r), _ptr(a), _ptr(b), a.size, threads)


# -------------------------------------------------------------------- sampler
def cdt_build(sigma: float) -> np.ndarray:
    buf = np.zeros(32768, dtype=np.uint64)
    n = lib().lib.lsro_cdt_build(float(sigma), _ptr(buf), buf.size)
    if n == 0:
        raise ValueError("bad sigma")
    return buf[:n].copy()


def cdt_sample(cdf: np.ndarray, u1: int, u2: int) -> int:
    cdf = _np_u64(cdf)
    return int(lib().lib.lsro_cdt_sample(_ptr(cdf), cdf.size, u1, u2))


def sample_gaussian_seeded(length: int, sigma: float, seed32: bytes) -> np.ndarray:
    out = np.zeros(length, dtype=np.uint64)
    rc = lib().lib.lsro_sample_gaussian_seeded(_ptr(out), length, float(sigma), seed32)
    if rc != 0:
        raise ValueError("sample_gaussian -> -1")
    return out.view(np.int64)


def chacha_block(key8, w12: int, w13: int, w14: int, w15: int) -> np.ndarray:
    key = np.ascontiguousarray(key8, dtype=np.uint32)
    out = np.zeros(16, dtype=np.uint32)
    u32p = C.POINTER(C.c_uint32)
    lib().lib.lsro_chacha_block(key.ctypes.data_as(u32p), w12, w13, w14, w15, out.ctypes.data_as(u32p))
    return out


# ----------------------------------------------------------------- commitment
class OracleLwe:
    """Mirror of the reference's LweContext (cpp-core/src/commitment.cpp:31-40,102-136)
    for the Module-LWE definition of DESIGN.md section 3."""

    def __init__(self, modulus: int, n: int, k: int, sigma: float, seed32: bytes, native: bool = False):
        assert len(seed32) == 32
        self._l = lib(native).lib
        self._h = self._l.lsro_lwe_create(modulus, n, k, float(sigma), seed32)
        if not self._h:
            raise ValueError("lwe_context_create -> NULL")
        self.n, self.k = n, k
        self.q = int(self._l.lsro_lwe_modulus(self._h))
        self.p = int(self._l.lsro_lwe_plain_modulus(self._h))
        self.delta = int(self._l.lsro_lwe_delta(self._h))
        self.words = int(self._l.lsro_lwe_words(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            self._l.lsro_lwe_free(self._h)
            self._h = None

    def matrix(self) -> np.ndarray:
        p = self._l.lsro_lwe_matrix(self._h)
        return np.ctypeslib.as_array(p, shape=(self.k, self.k, self.n)).copy()

    def trapdoor(self) -> np.ndarray:
        p = self._l.lsro_lwe_trapdoor(self._h)
        return np.ctypeslib.as_array(p, shape=(max(self.k - 1, 1), self.n))[: self.k - 1].copy()

    def sample_se(self, seed: int):
        s = np.zeros((self.k, self.n), dtype=np.int64)
        e = np.zeros((self.k, self.n), dtype=np.int64)
        self._l.lsro_lwe_sample_se(self._h, seed, s.ctypes.data_as(i64p), e.ctypes.data_as(i64p))
        return s, e

    def commit(self, msg, seed: int) -> np.ndarray:
        m = _np_u64(msg)
        out = np.zeros(self.words, dtype=np.uint64)
        mp = _ptr(m) if m.size else C.cast(C.c_void_p(1), u64p)   # non-NULL for empty msg
        rc = self._l.lsro_lwe_commit(self._h, mp, m.size, seed, _ptr(out))
        assert rc == 0
        return out

    def commit_explicit(self, msg, s, e) -> np.ndarray:
        """explicit mode: t = A*s + e + Delta*m for caller-supplied s, e ([k][n] int64)"""
        m = _np_u64(msg)
        s = np.ascontiguousarray(s, dtype=np.int64).reshape(self.k, self.n)
        e = np.ascontiguousarray(e, dtype=np.int64).reshape(self.k, self.n)
        out = np.zeros(self.words, dtype=np.uint64)
        mp = _ptr(m) if m.size else C.cast(C.c_void_p(1), u64p)
        rc = self._l.lsro_lwe_commit_explicit(self._h, mp, m.size, s.ctypes.data_as(i64p), e.ctypes.data_as(i64p), _ptr(out))
        assert rc == 0
        return out

    def commit_batch(self, msgs: np.ndarray, seeds, threads: int = 1, out: np.ndarray | None = None) -> np.ndarray:
        msgs = _np_u64(msgs)
        count, msg_len = msgs.shape
        seeds = _np_u64(seeds)
        assert seeds.size == count
        if out is None:
            out = np.zeros((count, self.words), dtype=np.uint64)
        rc = self._l.lsro_lwe_commit_batch(self._h, _ptr(msgs), msg_len, _ptr(seeds), count, _ptr(out), threads)
        assert rc == 0
        return out

    def verify(self, comm, msg) -> int:
        c = _np_u64(comm)
        m = _np_u64(msg)
        mp = _ptr(m) if m.size else C.cast(C.c_void_p(1), u64p)
        return int(self._l.lsro_lwe_verify(self._h, _ptr(c), c.size, mp, m.size))

    def linear_combine(self, comms, coeffs):
        arrs = [None if c is None else _np_u64(c) for c in comms]
        n = len(arrs)
        ptrs = (u64p * n)(*[C.cast(None, u64p) if a is None else _ptr(a) for a in arrs])
        lens = (C.c_size_t * n)(*[0 if a is None else a.size for a in arrs])
        cf = _np_u64(coeffs)
        out = np.zeros(self.words, dtype=np.uint64)
        rc = self._l.lsro_lwe_linear_combine(self._h, ptrs, lens, _ptr(cf), n, _ptr(out))
        return out if rc == 0 else None


# ---------------------------------------------------------- reference sampler
class RefSampler:
    """The reference's own utils.cpp compiled into oracle/_ref (see ref_sampler_shim.cpp)."""

    def __init__(self):
        path = _REF / "libref_sampler.so"
        if not path.exists():
            if Path("/root/reference/cpp-core/src/utils.cpp").exists():
                build()
            if not path.exists():
                raise FileNotFoundError(str(path))
        self._l = C.CDLL(str(path))
        self._l.ref_build_cdf.restype = C.c_size_t
        self._l.ref_build_cdf.argtypes = [C.c_double, u64p, C.c_size_t]
        self._l.ref_sample.argtypes = [C.c_double, u64p, C.c_size_t, u64p]

    def build_cdf(self, sigma: float) -> np.ndarray:
        buf = np.zeros(32768, dtype=np.uint64)
        n = self._l.ref_build_cdf(float(sigma), _ptr(buf), buf.size)
        return buf[:n].copy()

    def sample(self, sigma: float, draws: np.ndarray) -> np.ndarray:
        draws = _np_u64(draws)
        count = draws.size // 2
        out = np.zeros(count, dtype=np.uint64)
        self._l.ref_sample(float(sigma), _ptr(draws), count, _ptr(out))
        return out.view(np.int64)


# ------------------------------------------------ pure-Python cross-check code
def brv(x: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def py_forward_closed_form(coeffs, q: int, psi: int) -> list[int]:
    """out[i] = f(psi^(2*brv(i)+1)) mod q  (SURVEY.md 8c normative spec)."""
    n = len(coeffs)
    bits = n.bit_length() - 1
    out = []
    for i in range(n):
        x = pow(psi, 2 * brv(i, bits) + 1, q)
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * x + int(c)) % q
        out.append(acc)
    return out


def py_negacyclic_mul(a, b, q: int) -> list[int]:
    n = len(a)
    r = [0] * n
    for i in range(n):
        ai = int(a[i])
        if ai == 0:
            continue
        for j in range(n):
            k = i + j
            v = ai * int(b[j])
            if k >= n:
                r[k - n] = (r[k - n] - v) % q
            else:
                r[k] = (r[k] + v) % q
    return r


# ---------------------------------------------------------------- quotient pipeline (C restatement)
def cyclic_ntt_forward(coeffs, q: int, omega: int) -> np.ndarray:
    """ntt.rs:117-160 (natural order in and out); coeffs reduced mod q first."""
    x = _np_u64(coeffs).copy()
    if lib().lib.lsro_cyclic_ntt_forward(_ptr(x), x.size, q, omega) != 0:
        raise ValueError("bad size")
    return x


def cyclic_ntt_inverse(evals, q: int, omega: int) -> np.ndarray:
    x = _np_u64(evals).copy()
    if lib().lib.lsro_cyclic_ntt_inverse(_ptr(x), x.size, q, omega) != 0:
        raise ValueError("bad size")
    return x


def r1cs_quotient(rows: int, cols: int, A, B, C_, witness, q: int, omega: int, omega2: int):
    """r1cs.rs:474-503 on the NTT path for big m.  A, B, C_: (row[], col[], val[]) numpy triples.
    Returns (Q zero-padded to m, status) with status 1 when the witness does not satisfy the constraints."""
    u32p = C.POINTER(C.c_uint32)
    mats = [(np.ascontiguousarray(r, dtype=np.uint32), np.ascontiguousarray(c, dtype=np.uint32), _np_u64(v))
            for r, c, v in (A, B, C_)]
    rp = (u32p * 3)(*[m[0].ctypes.data_as(u32p) for m in mats])
    cp = (u32p * 3)(*[m[1].ctypes.data_as(u32p) for m in mats])
    vp = (u64p * 3)(*[_ptr(m[2]) for m in mats])
    nn = (C.c_size_t * 3)(*[m[2].size for m in mats])
    w = _np_u64(witness)
    out = np.zeros(rows, dtype=np.uint64)
    st = lib().lib.lsro_r1cs_quotient(rows, cols, rp, cp, vp, nn, _ptr(w), q, omega, omega2, _ptr(out))
    if st < 0:
        raise ValueError("bad arguments")
    return out, st
