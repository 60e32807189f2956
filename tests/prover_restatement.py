"""TEST INFRASTRUCTURE: Python restatement of the reference's Rust host-side prover
and verifier (the part that "stays Rust" and calls the C ABI), so the end-to-end
claim -- proofs built on the B200 commitment still satisfy the reference
verifier's equations -- can be checked without a Rust toolchain.

Restated, function for function (rust-api/lambda-snark/src/):
    arith.rs:8-106          mul_mod / add_mod / sub_mod / mod_pow / mod_inverse (Fermat, then Euclid)
    sparse_matrix.rs:259-289  SparseMatrix::mul_vec
    r1cs.rs:148-173,296-304   is_satisfied, compute_constraint_evals
    r1cs.rs:362-373           eval_poly (power-accumulating, not Horner)
    r1cs.rs:424-440           eval_vanishing (baseline domain {0..m-1})
    r1cs.rs:474-503           compute_quotient_poly
    r1cs.rs:641-696,808-828   lagrange_basis_sequential, lagrange_interpolate_baseline
    r1cs.rs:846-863,959-1065  poly_mul, poly_sub, vanishing_poly, poly_div_vanishing
    challenge.rs:102-134      Challenge::derive  (SHA3-256, "LAMBDA-SNARK-R-FS-v1")
    lib.rs:747-809            prove_r1cs
    lib.rs:1016-1082          verify_r1cs
Only the baseline (sequential-domain) path is restated: it is the one taken for every
modulus the reference's tests and CLI use (2^44+1, 17592186044423), r1cs.rs:386-389.
"""
from __future__ import annotations

import hashlib
import struct
from dataclasses import dataclass


# ------------------------------------------------------------------ arith.rs
def mul_mod(a, b, q):
    return (a * b) % q


def add_mod(a, b, q):
    return (a + b) % q


def sub_mod(a, b, q):
    return (a - b) % q


def mod_inverse(v, q):
    if v == 0 or q <= 1:
        raise ZeroDivisionError
    r = v % q
    if r == 0:
        raise ZeroDivisionError
    if q & 1:
        cand = pow(r, q - 2, q)
        if (cand * r) % q == 1 % q:
            return cand
    t, nt, a, b = 0, 1, q, r          # arith.rs:86-106, needed for the composite 2^44+1
    while b:
        k = a // b
        t, nt = nt, t - k * nt
        a, b = b, a - k * b
    if a != 1:
        raise ZeroDivisionError
    return t % q


# ------------------------------------------------------------------ r1cs.rs
@dataclass
class R1CS:
    m: int
    n: int
    l: int
    A: list      # rows of {col: value}
    B: list
    C: list
    modulus: int

    @staticmethod
    def _mul_vec(rows, z, q):
        return [sum((v % q) * (z[c] % q) for c, v in row.items()) % q for row in rows]

    def evals(self, z):
        q = self.modulus
        return self._mul_vec(self.A, z, q), self._mul_vec(self.B, z, q), self._mul_vec(self.C, z, q)

    def is_satisfied(self, z):
        a, b, c = self.evals(z)
        return all((x * y) % self.modulus == w for x, y, w in zip(a, b, c))

    def public_inputs(self, z):
        return list(z[: self.l])

    def eval_poly(self, poly, x):
        q = self.modulus
        res, power = 0, 1
        for c in poly:
            res = add_mod(res, mul_mod(c, power, q), q)
            power = mul_mod(power, x, q)
        return res

    def eval_vanishing(self, x):
        q = self.modulus
        res = 1
        for i in range(self.m):
            res = mul_mod(res, sub_mod(x, i % q, q), q)
        return res

    def compute_quotient_poly(self, z):
        if not self.is_satisfied(z):
            raise ValueError("Witness does not satisfy R1CS constraints")
        q = self.modulus
        a, b, c = self.evals(z)
        ap, bp, cp = (lagrange_interpolate(e, q) for e in (a, b, c))
        num = poly_sub(poly_mul(ap, bp, q), cp, q)
        return poly_div_vanishing(num, self.m, q)


def poly_mul_linear(poly, root, q):      # poly * (X - root)
    out = [0] * (len(poly) + 1)
    for i, c in enumerate(poly):
        out[i + 1] = add_mod(out[i + 1], c, q)
        out[i] = sub_mod(out[i], mul_mod(c, root % q, q), q)
    return out


def lagrange_basis_sequential(i, m, q):
    poly = [1]
    for j in range(m):
        if j != i:
            poly = poly_mul_linear(poly, j, q)
    den = 1
    for j in range(m):
        if j != i:
            den = mul_mod(den, sub_mod(i % q, j % q, q), q)
    inv = mod_inverse(den, q) if m > 1 else 1
    poly = [mul_mod(c, inv, q) for c in poly]
    return (poly + [0] * m)[:m]


def lagrange_interpolate(evals, q):
    m = len(evals)
    res = [0] * m
    for i in range(m):
        basis = lagrange_basis_sequential(i, m, q)
        for j in range(m):
            res[j] = add_mod(res[j], mul_mod(evals[i], basis[j], q), q)
    return res


def poly_mul(a, b, q):
    if not a or not b:
        return [0]
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        for j, y in enumerate(b):
            out[i + j] = add_mod(out[i + j], mul_mod(x % q, y % q, q), q)
    return out


def poly_sub(a, b, q):
    n = max(len(a), len(b))
    a = a + [0] * (n - len(a))
    b = b + [0] * (n - len(b))
    return [sub_mod(x, y, q) for x, y in zip(a, b)]


def vanishing_poly(m, q):
    poly = [1]
    for i in range(m):
        poly = poly_mul_linear(poly, i, q)
    return poly


def poly_div_vanishing(num, m, q):
    if not num:
        return [0]
    div = vanishing_poly(m, q)
    rem = list(num)
    deg_num, deg_div = len(rem) - 1, len(div) - 1
    if deg_num < deg_div:
        if all(x == 0 for x in rem):
            return [0]
        raise ValueError("Polynomial division by Z_H: remainder non-zero (witness invalid)")
    deg_q = deg_num - deg_div
    quo = [0] * (deg_q + 1)
    lead_inv = mod_inverse(div[deg_div], q)
    for i in range(deg_q, -1, -1):
        idx = i + deg_div
        if idx < len(rem) and idx > 0:
            qc = mul_mod(rem[idx] % q, lead_inv, q)
            quo[i] = qc
            for j, d in enumerate(div):
                if i + j < len(rem):
                    rem[i + j] = sub_mod(rem[i + j], mul_mod(qc, d % q, q), q)
    if any(rem):
        raise ValueError("Polynomial division by Z_H: remainder non-zero (witness invalid)")
    while len(quo) > 1 and quo[-1] == 0:
        quo.pop()
    return quo


# --------------------------------------------------------------- challenge.rs
def challenge_derive(public_inputs, commitment_words, modulus):
    h = hashlib.sha3_256()
    h.update(b"LAMBDA-SNARK-R-FS-v1")
    h.update(struct.pack("<Q", len(public_inputs)))
    for v in public_inputs:
        h.update(struct.pack("<Q", int(v)))
    h.update(struct.pack("<Q", len(commitment_words)))
    import numpy as np
    h.update(np.asarray(commitment_words, dtype="<u8").tobytes())      # each word little-endian, challenge.rs:118-122
    digest = h.digest()
    return struct.unpack("<Q", digest[:8])[0] % modulus, digest


# --------------------------------------------------------------------- lib.rs
@dataclass
class ProofR1CS:
    commitment_words: object
    alpha: int
    beta: int
    q_alpha: int
    q_beta: int
    a_z_alpha: int
    b_z_alpha: int
    c_z_alpha: int
    a_z_beta: int
    b_z_beta: int
    c_z_beta: int
    opening_alpha: int
    opening_beta: int


def prove_r1cs(r1cs: R1CS, witness, commit, seed):
    """lib.rs:747-809.  `commit(field_elements, seed)` is Commitment::new -> as_bytes() words."""
    q = r1cs.modulus
    q_coeffs = r1cs.compute_quotient_poly(witness)
    words = commit(q_coeffs, seed)
    alpha, _ = challenge_derive(r1cs.public_inputs(witness), words, q)
    beta, _ = challenge_derive([alpha], words, q)
    a, b, c = r1cs.evals(witness)
    ap, bp, cp = (lagrange_interpolate(e, q) for e in (a, b, c))
    ev = r1cs.eval_poly
    return ProofR1CS(words, alpha, beta, ev(q_coeffs, alpha), ev(q_coeffs, beta),
                     ev(ap, alpha), ev(bp, alpha), ev(cp, alpha), ev(ap, beta), ev(bp, beta), ev(cp, beta),
                     ev(q_coeffs, alpha), ev(q_coeffs, beta)), q_coeffs


def verify_r1cs(proof: ProofR1CS, public_inputs, r1cs: R1CS) -> bool:
    """lib.rs:1016-1082, equation for equation."""
    q = r1cs.modulus
    alpha, _ = challenge_derive(public_inputs, proof.commitment_words, q)
    if alpha != proof.alpha:
        return False
    beta, _ = challenge_derive([alpha], proof.commitment_words, q)
    if beta != proof.beta:
        return False
    for x, qx, ax, bx, cx in ((alpha, proof.q_alpha, proof.a_z_alpha, proof.b_z_alpha, proof.c_z_alpha),
                              (beta, proof.q_beta, proof.a_z_beta, proof.b_z_beta, proof.c_z_beta)):
        if mul_mod(qx, r1cs.eval_vanishing(x), q) != sub_mod(mul_mod(ax, bx, q), cx, q):
            return False
    return proof.opening_alpha == proof.q_alpha and proof.opening_beta == proof.q_beta


# ------------------------------------------------------------------- circuits
def tv1(q):      # test-vectors/tv-1-multiplication/constraints.json
    return R1CS(1, 4, 2, [{1: 1}], [{2: 1}], [{3: 1}], q), [1, 7, 13, 91]


def tv2(q):      # test-vectors/tv-2-plaquette/constraints.json (-1 entries stored as q-1 by the Rust side)
    return R1CS(1, 5, 1, [{0: 1}], [{1: 1, 2: 1, 3: q - 1, 4: q - 1}], [{}], q), [1, 314, 628, 471, 471]


def multiplication_gates(m, q):      # tests/integration_matrix.rs:25-75
    n = 1 + 3 * m
    A, B, C, z = [], [], [], [1]
    for i in range(m):
        A.append({1 + 3 * i: 1}); B.append({2 + 3 * i: 1}); C.append({3 + 3 * i: 1})
        x, y = (2 * i + 3) % q, (3 * i + 5) % q
        z += [x, y, (x * y) % q]
    return R1CS(m, n, 1, A, B, C, q), z


def healthcare(q=17592186044423):    # lambda-snark-cli/src/main.rs:774-886
    one, risk, gh, ah, bh, temp, allh = 0, 1, 5, 6, 7, 8, 9
    A = [{gh: 1}, {ah: 1}, {bh: 1}, {gh: 1}, {temp: 1}, {one: 1, allh: 2}]
    B = [{gh: 1, one: q - 1}, {ah: 1, one: q - 1}, {bh: 1, one: q - 1}, {ah: 1}, {bh: 1}, {one: 1}]
    C = [{}, {}, {}, {temp: 1}, {allh: 1}, {risk: 1}]
    return R1CS(6, 10, 2, A, B, C, q), [1, 3, 142, 45, 31, 1, 1, 1, 1, 1]
