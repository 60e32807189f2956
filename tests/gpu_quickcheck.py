"""Fast end-to-end sanity run on a GPU box (not a pytest file): every product
call goes through the C ABI and is compared with the CPU oracle.  Used while
developing; the real parity suite is tests/test_gpu_*.py."""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from lambda_snark_r_b200 import api  # noqa: E402
from oracle import oracle as O  # noqa: E402

Q0 = 17592169062401
Q1 = 17592180539393
Q60 = 1152921504606584833


def check_ntt(q, n, batch, rng):
    g = api.NttContext(q, n)
    o = O.OracleNtt(q, n)
    assert g.root == o.psi, (g.root, o.psi)
    x = rng.integers(0, q, size=(batch, n), dtype=np.uint64)
    x[0, :] = 0
    if batch > 1:
        x[1, :] = q - 1
    f_g = g.forward_batch(x)
    f_o = o.forward(x)
    ok_f = np.array_equal(f_g, f_o)
    i_g = g.inverse_batch(f_g)
    ok_i = np.array_equal(i_g, x)
    y = rng.integers(0, q, size=(batch, n), dtype=np.uint64)
    i2_g = g.inverse_batch(y)
    ok_i2 = np.array_equal(i2_g, o.inverse(y))
    a = rng.integers(0, 2**64, size=batch * n, dtype=np.uint64)
    b = rng.integers(0, 2**64, size=batch * n, dtype=np.uint64)
    ok_m = np.array_equal(g.mul_pointwise_batch(a, b), o.mul_pointwise(a, b))
    print(f"ntt q={q} n={n} batch={batch}: fwd={ok_f} roundtrip={ok_i} inv={ok_i2} mul={ok_m}", flush=True)
    g.close()
    return ok_f and ok_i and ok_i2 and ok_m


def check_commit(q, n, k, sigma, rng, paths=(1, 2)):
    seed32 = bytes(range(32))
    ctx = api.LweContext(api.Params(n=n, k=k, q=q, sigma=sigma), seed32=seed32)
    orc = O.OracleLwe(q, n, k, sigma, seed32)
    ok = (ctx.q, ctx.p, ctx.delta) == (orc.q, orc.p, orc.delta)
    ok = ok and np.array_equal(ctx.matrix(), orc.matrix())
    print(f"commit n={n} k={k}: params/matrix {ok}", flush=True)
    s_g, e_g = ctx.sample_se(0xC0FFEE)
    s_o, e_o = orc.sample_se(0xC0FFEE)
    ok_s = np.array_equal(s_g, s_o) and np.array_equal(e_g, e_o)
    print(f"   sampler {ok_s}", flush=True)
    ok = ok and ok_s
    count = 5
    msgs = rng.integers(0, 2**64, size=(count, n), dtype=np.uint64)
    msgs[0, :] = 0
    seeds = np.array([(i * 0x9E3779B97F4A7C15) % 2**64 for i in range(1, count + 1)], dtype=np.uint64)
    want = orc.commit_batch(msgs, seeds)
    for path in paths:
        try:
            ctx.set_commit_path(path)
            got = ctx.commit_batch(msgs, seeds)
            same = np.array_equal(got, want)
        except api.LambdaSnarkError as e:
            print(f"   path {path}: {e}")
            same = path == 2   # fused may be unsupported for this shape
        print(f"   path {path}: {same}", flush=True)
        ok = ok and same
    ctx.set_commit_path(0)
    small = rng.integers(0, ctx.p, size=(count, 7), dtype=np.uint64)
    cm = ctx.commit_batch(small, seeds)
    res = ctx.verify_batch(cm, small)
    bad = small.copy(); bad[:, 3] ^= 1
    res_bad = ctx.verify_batch(cm, bad)
    ok_v = res.tolist() == [1] * count and res_bad.tolist() == [0] * count
    ok_v = ok_v and [orc.verify(cm[i], small[i]) for i in range(count)] == [1] * count
    print(f"   verify {ok_v} {res.tolist()} {res_bad.tolist()}", flush=True)
    ctx.close()
    return ok and ok_v


def main():
    rng = np.random.Generator(np.random.PCG64(0x5EED))
    print("devices:", api.device_count())
    ok = True
    t0 = time.time()
    for q, n, batch in [(12289, 256, 3), (Q0, 4096, 5), (Q0, 1024, 9), (Q0, 2, 7), (Q0, 16, 300), (Q0, 512, 3),
                        (Q0, 2048, 3), (Q1, 8192, 3), (Q1, 16384, 2), (Q1, 32768, 2), (Q1, 65536, 2), (Q1, 131072, 1),
                        (Q60, 4096, 3), (Q60, 65536, 1), (Q60, 64, 5)]:
        ok = check_ntt(q, n, batch, rng) and ok
    for q, n, k in [(Q0, 4096, 2), (Q0, 1024, 2), (Q0, 4096, 3), (Q0, 256, 2), (Q1, 8192, 2), (Q0, 4096, 1)]:
        ok = check_commit(q, n, k, 3.19, rng) and ok
    s = api.sample_gaussian(4099, 3.2, seed32=bytes(32))
    ok_s = np.array_equal(s, O.sample_gaussian_seeded(4099, 3.2, bytes(32)))
    print("sample_gaussian seeded", ok_s)
    ok = ok and ok_s
    print("ALL OK" if ok else "FAILURES", f"({time.time() - t0:.1f}s)")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
