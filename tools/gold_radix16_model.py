"""Model of a radix-16 regrouping of the Goldilocks transforms (studied for the quotient pipeline, not built).

The register networks of csrc/lsr_ntt.cuh use one twiddle per (stage, group): tw[(T0 << r) + t].  For every pass the
model checks, against the plain networks, that these factor as
        tw(r, t) = B^(2^(R-1-r)) * zeta_r^brv_r(t),      B = tw[T0 << (R-1)],  zeta_r a primitive 2^(r+1)-th root of unity,
so a pass equals: scale input j by B^j (one general product per element, 2^R - 1 powers of B from a table), then run the
network with the zeta powers only -- and in Goldilocks every root of unity of order <= 64 is a power of two (2 has order
192), so the inner twiddles are shifts.  For the reference's root the shift exponents are the same for every pass, block,
size and for both the cyclic and the negacyclic tables: ((0), (0, 48), (0, 48, 120, 168), (0, 48, 120, 168, 156, 12, 84, 132)).
Inverse passes: the network with the negated shifts, then scale output j by B^-j (and n^-1 on the pass that ends the transform).

Cost model per radix-16 work item (instructions, from the SASS counts of DESIGN.md 4.7): now 32 butterflies x (general
product 32 + add / sub 11) = 1376; regrouped 15 general products (480) + 32 butterflies on canonical values (12 each, 384) +
17 non-trivial shift products (~18 each, 306) + negations = ~1190: -13 %.  The kernels are ALU-pipe bound, so that is
the gain to expect; it needs new power tables per pass shape and a second set of networks, and was left for later.

    python tools/gold_radix16_model.py        (checks n = 32, 256, 4096, cyclic and negacyclic, R = 1 .. 5)
"""
import random
P = 2**64 - 2**32 + 1
G = 1753635133440165772          # 2^32-th root used by the reference

def brv(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1); x >>= 1
    return r

def tables(n, cyclic, root):
    logn = n.bit_length() - 1
    fwd = [0] * n; inv = [0] * n
    if cyclic:
        om = root
        fwd[0] = 1
        for s in range(logn):
            M = 1 << s
            for i in range(M):
                fwd[M + i] = pow(om, brv(i, s) * (n // (2 * M)), P)
    else:
        psi = root
        for i in range(n):
            fwd[brv(i, logn)] = pow(psi, i, P)
    for k in range(1, n):
        inv[k] = pow(fwd[k], P - 2, P)
    inv[0] = 1
    ninv = pow(n, P - 2, P)
    inv[1] = inv[1] * ninv % P           # scalar folded into the last stage
    return fwd, inv, ninv

def fwd_network(v, R, tw, T0):
    for r in range(R):
        half = 1 << (R - 1 - r)
        for t in range(1 << r):
            w = tw[(T0 << r) + t]
            for jl in range(half):
                j = (t << (R - r)) + jl; jj = j + half
                T = v[jj] * w % P; X = v[j]
                v[j] = (X + T) % P; v[jj] = (X - T) % P

def inv_network(v, R, tw, T0, final, ninv):
    for r in range(R):
        half = 1 << r; fr = R - 1 - r
        for t in range(1 << fr):
            w = tw[(T0 << fr) + t]
            for jl in range(half):
                j = (t << (r + 1)) + jl; jj = j + half
                X, Y = v[j], v[jj]
                if final and r == R - 1:
                    v[j] = (X + Y) * ninv % P; v[jj] = (X - Y) * w % P     # w = inv[1] already carries n^-1
                else:
                    v[j] = (X + Y) % P; v[jj] = (X - Y) * w % P

# ---- regrouped forms
def zeta_tables(R, fwd, T0):
    """shift twiddles z[r][t] = tw(r,t) / tw(r,0) -- must be powers of two (returned as exponents mod 192)"""
    logs = {pow(2, e, P): e for e in range(192)}
    z = []
    for r in range(R):
        row = []
        b = fwd[T0 << r]
        for t in range(1 << r):
            q = fwd[(T0 << r) + t] * pow(b, P - 2, P) % P
            row.append(logs[q])            # KeyError if not a power of two
        z.append(row)
    return z

def fwd_network_regrouped(v, R, fwd, T0):
    B = fwd[T0 << (R - 1)]
    z = zeta_tables(R, fwd, T0)
    for j in range(1, 1 << R):
        v[j] = v[j] * pow(B, j, P) % P
    for r in range(R):
        half = 1 << (R - 1 - r)
        for t in range(1 << r):
            e = z[r][t]
            for jl in range(half):
                j = (t << (R - r)) + jl; jj = j + half
                T = v[jj] * pow(2, e, P) % P; X = v[j]
                v[j] = (X + T) % P; v[jj] = (X - T) % P
    return z

def inv_network_regrouped(v, R, fwd, inv, T0, final, ninv):
    Binv = pow(fwd[T0 << (R - 1)], P - 2, P)
    z = zeta_tables(R, fwd, T0)
    for r in range(R):
        half = 1 << r; fr = R - 1 - r
        for t in range(1 << fr):
            e = (192 - z[fr][t]) % 192
            for jl in range(half):
                j = (t << (r + 1)) + jl; jj = j + half
                X, Y = v[j], v[jj]
                v[j] = (X + Y) % P; v[jj] = (X - Y) * pow(2, e, P) % P
    for j in range(1 << R):
        s = pow(Binv, j, P) * (ninv if final else 1) % P
        v[j] = v[j] * s % P

def check(n, cyclic):
    logn = n.bit_length() - 1
    root = pow(G, 2**32 // n, P) if cyclic else pow(G, 2**32 // (2 * n), P)
    fwd, inv, ninv = tables(n, cyclic, root)
    rng = random.Random(n + cyclic)
    zsets = set()
    for R in (1, 2, 3, 4, 5):
        for s0 in range(0, logn - R + 1):
            for i0 in {0, (1 << s0) - 1, rng.randrange(1 << s0)}:
                T0 = (1 << s0) + i0
                x = [rng.randrange(P) for _ in range(1 << R)]
                a = list(x); fwd_network(a, R, fwd, T0)
                b = list(x); z = fwd_network_regrouped(b, R, fwd, T0)
                assert a == b, ("fwd", n, cyclic, R, s0, i0)
                zsets.add((R, tuple(tuple(r) for r in z)))
                final = (s0 == 0)
                a = list(x); inv_network(a, R, inv, T0, final, ninv)
                b = list(x); inv_network_regrouped(b, R, fwd, inv, T0, final, ninv)
                assert a == b, ("inv", n, cyclic, R, s0, i0)
    return zsets

for n in (32, 256, 4096):
    for cyc in (True, False):
        zs = check(n, cyc)
        print(n, "cyclic" if cyc else "negacyclic", "ok; distinct shift tables per R:", {R: len([1 for (r, _) in zs if r == R]) for R in (1, 2, 3, 4, 5)})
        for (R, z) in sorted(zs):
            if R == 4: print("   R=4 shifts:", z)
