"""TEST INFRASTRUCTURE (never imported by the product): CPU restatement of the Rust side of the
quotient pipeline, function for function, in plain Python integers.

    rust-api/lambda-snark/src/ntt.rs:68-96      reverse_bits, bit_reverse_permutation
    rust-api/lambda-snark/src/ntt.rs:117-165    ntt_forward (radix-2 DIT, natural in / natural out)
    rust-api/lambda-snark/src/ntt.rs:185-201    ntt_inverse
    rust-api/lambda-snark/src/ntt.rs:214-221    compute_root_of_unity
    rust-api/lambda-snark/src/sparse_matrix.rs:259-289   SparseMatrix::mul_vec
    rust-api/lambda-snark/src/r1cs.rs:474-503   compute_quotient_poly (NTT path: fft-ntt feature)
    rust-api/lambda-snark/src/r1cs.rs:746-771   lagrange_interpolate_ntt
    rust-api/lambda-snark/src/r1cs.rs:846-895   poly_mul, poly_sub
    rust-api/lambda-snark/src/r1cs.rs:953-965   vanishing_poly(use_ntt = true) = X^m - 1
    rust-api/lambda-snark/src/r1cs.rs:995-1065  poly_div_vanishing

Pinned against the reference's own unit tests (ntt.rs:284-347, restated in tests/test_oracle_quotient.py):
f = 1 + 2X -> [3, q-1]; [1,2,3,4] -> out[0] = 10; [1..8] -> out[0] = 36; round trips for n = 2 .. 1024.
The Rust crate itself cannot be built here (no rustc), so beyond those KATs this is "parity unpinned"
against a cargo build; the quotient of an exact division is unique, which is what the GPU is held to.
"""
from __future__ import annotations

NTT_MODULUS = 18_446_744_069_414_584_321          # lambda-snark-core/src/lib.rs:58, 2^64 - 2^32 + 1
NTT_PRIMITIVE_ROOT = 1_753_635_133_440_165_772    # lib.rs:78, primitive 2^32-th root of unity
NTT_FRIENDLY_MODULUS = 17592169062401             # r1cs.rs:527


def reverse_bits(x: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def bit_reverse_permutation(data: list) -> None:
    n = len(data)
    log_n = n.bit_length() - 1
    for i in range(n):
        j = reverse_bits(i, log_n)
        if i < j:
            data[i], data[j] = data[j], data[i]


def ntt_forward(coeffs, modulus: int, omega: int) -> list:
    n = len(coeffs)
    assert n & (n - 1) == 0 and n > 0
    if n == 1:
        return list(coeffs)
    data = list(coeffs)
    bit_reverse_permutation(data)
    log_n = n.bit_length() - 1
    for s in range(1, log_n + 1):
        m = 1 << s
        m_half = m >> 1
        omega_m = pow(omega, n // m, modulus)
        for k in range(0, n, m):
            omega_power = 1
            for j in range(m_half):
                t = (data[k + j + m_half] * omega_power) % modulus
                u = data[k + j]
                data[k + j] = (u + t) % modulus
                data[k + j + m_half] = (u - t) % modulus
                omega_power = (omega_power * omega_m) % modulus
    return data


def ntt_inverse(evals, modulus: int, omega: int) -> list:
    n = len(evals)
    if n == 1:
        return list(evals)
    omega_inv = pow(omega, modulus - 2, modulus)
    coeffs = ntt_forward(evals, modulus, omega_inv)
    n_inv = pow(n, modulus - 2, modulus)
    return [(c * n_inv) % modulus for c in coeffs]


def compute_root_of_unity(n: int, modulus: int = NTT_MODULUS, primitive_root: int = NTT_PRIMITIVE_ROOT) -> int:
    assert n & (n - 1) == 0 and n <= (1 << 32)
    return pow(primitive_root, (1 << 32) // n, modulus)


def reference_root(q: int, n: int) -> int:
    """The root the reference would use for a size-n domain over q."""
    if q == NTT_MODULUS:
        return compute_root_of_unity(n)
    if q == NTT_FRIENDLY_MODULUS:
        return pow(3, (q - 1) // n, q)          # r1cs.rs:534-547 ROOTS_OF_UNITY (generator 3)
    raise ValueError("no reference root for this modulus")


def mul_vec(rows: int, entries, v, modulus: int) -> list:
    """entries: iterable of (row, col, value) in CSR order (row-major, insertion order within a row)."""
    out = [0] * rows
    for r, c, val in entries:
        out[r] = (out[r] + (val % modulus) * (v[c] % modulus)) % modulus
    return out


def poly_mul(a, b, q):
    if not a or not b:
        return [0]
    res = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x == 0:
            continue
        for j, y in enumerate(b):
            res[i + j] = (res[i + j] + (x % q) * (y % q)) % q
    return res


def poly_sub(a, b, q):
    n = max(len(a), len(b))
    res = [((a[i] if i < len(a) else 0) % q - (b[i] if i < len(b) else 0) % q) % q for i in range(n)]
    while len(res) > 1 and res[-1] == 0:
        res.pop()
    return res


def poly_div_vanishing_ntt(numerator, m: int, q: int):
    """Long division by X^m - 1 (r1cs.rs:995-1065 with use_ntt = true).  Raises ValueError on a remainder."""
    if not numerator:
        return [0]
    divisor = [0] * (m + 1)
    divisor[0] = q - 1
    divisor[m] = 1
    rem = list(numerator)
    deg_num, deg_div = len(rem) - 1, m
    if deg_num < deg_div:
        if all(x == 0 for x in rem):
            return [0]
        raise ValueError("remainder non-zero (witness invalid)")
    deg_quot = deg_num - deg_div
    quot = [0] * (deg_quot + 1)
    for i in range(deg_quot, -1, -1):
        idx = i + deg_div
        if idx < len(rem) and idx > 0:
            qc = rem[idx] % q                       # lead coefficient of the divisor is 1
            quot[i] = qc
            for j in (0, m):                        # the only non-zero divisor coefficients
                pos = i + j
                if pos < len(rem):
                    rem[pos] = (rem[pos] - qc * divisor[j]) % q
    if any(x != 0 for x in rem):
        raise ValueError("remainder non-zero (witness invalid)")
    while len(quot) > 1 and quot[-1] == 0:
        quot.pop()
    return quot


def compute_quotient_poly(rows: int, A, B, C, witness, q: int, omega: int | None = None):
    """r1cs.rs:474-503 on the NTT path.  A, B, C: lists of (row, col, value)."""
    m = rows
    a = mul_vec(m, A, witness, q)
    b = mul_vec(m, B, witness, q)
    c = mul_vec(m, C, witness, q)
    if any((x * y) % q != z for x, y, z in zip(a, b, c)):                   # is_satisfied, r1cs.rs:477-481
        raise ValueError("Witness does not satisfy R1CS constraints")
    if omega is None:
        omega = reference_root(q, m) if m > 1 else 1
    ap, bp, cp = (ntt_inverse(e, q, omega) for e in (a, b, c))
    num = poly_sub(poly_mul(ap, bp, q), cp, q)
    return poly_div_vanishing_ntt(num, m, q)


def horner(poly, x, q):
    acc = 0
    for c in reversed(poly):
        acc = (acc * x + c) % q
    return acc
