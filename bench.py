#!/usr/bin/env python
"""bench.py -- throughput of the prover hot path on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            (this repo's CUDA path)
    python bench.py --impl reference --gpus N --steps K ...  (reference arm: CPU, host cores)

A "step" is one pass of the hot path over one batch of synthetic input: a
batch of Module-LWE commitments t = A*s + e + Delta*m at n=4096, k=2,
q=17592169062401, sigma=3.19 (BASELINE.json configs[3]); each commitment runs
the sampler, k forward NTTs, the k x k NTT-domain mat-vec and k inverse NTTs
inside ONE fused kernel.  The batched NTT sweep at the same n (configs[2]) is
reported beside it under "ntt".

value      : commitments/s, whole job, inputs resident in HBM (CUDA events)
e2e        : same metric through the C ABI's host-pointer call lwe_commit_batch
             with pinned HOST buffers, copies inside the timed region
roofline   : dominant kernel vs the measured HBM peak (MEASURED_PEAKS.json), plus the
             compute rooflines: the FP64 pipe the butterflies run on (q < 2^45: exact
             modular products by error-free fma multiplication, 8 FP64 instructions per
             butterfly) and, for continuity with SURVEY 8d, the IMAD normalisation
cpu_baseline / --impl reference : the oracle's C port of the same algorithm on
             the host cores (the reference's SEAL path cannot be built here)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_RING, K_RANK, Q_MOD, SIGMA = 4096, 2, 17592169062401, 3.19
CTX_SEED = bytes(range(32))
SEED_BASE = 0xC0FFEE
ALG_BYTES_COMMIT = 8 * N_RING + 8 * K_RANK * N_RING + 8          # message in + container out
ALG_BYTES_NTT = 16 * N_RING                                       # read + write, in place
# quotient pipeline, every stage touching its operands once (multiplication-gate R1CS: 3 non-zeros per constraint):
# mat-vec 24 (witness) + 48 (CSR: 3 x (col 4 + val 8) + 3 row pointers) + 24 (evaluations out) = 96; 3 inverse + 2 forward
# + 1 inverse size-m transforms 16 each = 96; coset product 16 in + 8 out = 24; (C - N) / 2 on coefficients 16 in + 8 out = 24
QUOTIENT_BYTES_PER_CONSTRAINT = 96 + 96 + 24 + 24
BUTTERFLIES_NTT = (N_RING // 2) * 12
IMAD_PER_MODMUL = 10                                              # SURVEY 8d normalisation
MODMUL_COMMIT = 2 * K_RANK * BUTTERFLIES_NTT + K_RANK * (N_RING // 2) + K_RANK * K_RANK * N_RING   # 118784
FP64_PER_BUTTERFLY = 8                                            # mulmod_f (6) + add + sub, DESIGN.md 4.2
FP64_PER_MODMUL = 6
FP64_NTT_FWD = BUTTERFLIES_NTT * FP64_PER_BUTTERFLY
FP64_NTT_INV = FP64_NTT_FWD + (N_RING // 2) * FP64_PER_MODMUL     # n^-1 folded into the last stage
FP64_COMMIT = K_RANK * (FP64_NTT_FWD + FP64_NTT_INV) + K_RANK * K_RANK * N_RING * (FP64_PER_MODMUL + 1)
PARITY_N, PARITY_NTT_ROWS = 32, 4


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def traffic_note():
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        return json.loads(p.read_text())
    return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------ CPU arm
def host_threads() -> int:
    """All host cores this process may use (torchrun exports OMP_NUM_THREADS=1; the CPU legs ignore it)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_commit_rate(target_s: float, threads: int | None = None):
    """Oracle C port (kind "port"), all host threads, bounded sample of the same workload."""
    from oracle import oracle as O
    try:
        O.build(native=True)
        native = True
    except Exception:
        native = False
    ctx = O.OracleLwe(Q_MOD, N_RING, K_RANK, SIGMA, CTX_SEED, native=native)
    threads = threads or host_threads()
    rng = np.random.Generator(np.random.PCG64(0x5EED))

    def run(count):
        msgs = rng.integers(0, Q_MOD, size=(count, N_RING), dtype=np.uint64)
        seeds = (np.arange(count, dtype=np.uint64) + np.uint64(SEED_BASE))
        out = np.zeros((count, ctx.words), dtype=np.uint64)
        t0 = time.perf_counter()
        ctx.commit_batch(msgs, seeds, threads=threads, out=out)
        return time.perf_counter() - t0

    calib = max(threads * 4, 16)
    dt = run(calib)
    count = int(max(calib, min(200000, calib * target_s / max(dt, 1e-6))))
    count -= count % threads or 0
    dt = run(count)
    return count / dt, threads, count, dt, native


def cpu_ntt_rate(target_s: float, threads: int | None = None):
    from oracle import oracle as O
    try:
        O.build(native=True)
        native = True
    except Exception:
        native = False
    ctx = O.OracleNtt(Q_MOD, N_RING, native=native)
    threads = threads or host_threads()
    rng = np.random.Generator(np.random.PCG64(0x5EED))
    calib = threads * 64
    x = rng.integers(0, Q_MOD, size=(calib, N_RING), dtype=np.uint64)
    t0 = time.perf_counter(); ctx.forward_inplace(x, threads); dt = time.perf_counter() - t0
    count = int(max(calib, min(2_000_000, calib * target_s / max(dt, 1e-6))))
    x = rng.integers(0, Q_MOD, size=(count, N_RING), dtype=np.uint64)
    t0 = time.perf_counter(); ctx.forward_inplace(x, threads); dt = time.perf_counter() - t0
    return count / dt, threads, count, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    per_step = max(2.0, min(20.0, 60.0 / max(args.steps + args.warmup, 1)))
    rates = []
    meta = None
    for i in range(args.warmup + args.steps):
        rate, threads, count, dt, native = cpu_commit_rate(per_step)
        meta = (threads, count, native)
        if i >= args.warmup:
            rates.append((count, dt))
    total = sum(c for c, _ in rates)
    secs = sum(d for _, d in rates)
    value = total / secs
    threads, count, native = meta
    sample = (f"{count} commitments per step (n={N_RING}, k={K_RANK}, full-length messages), "
              f"{threads} OpenMP threads, oracle C port of the same algorithm"
              f"{' (-march=native)' if native else ''}; SEAL/BFV reference cannot be built here")
    line = {
        "impl": "reference", "metric": "lwe_commitments_per_sec", "value": value, "unit": "commitments/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * secs / max(len(rates), 1), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(count, "cpu"),
        "cpu_baseline": {"value": value, "unit": "commitments/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "commitments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(batch, where):
    return {"workload": f"Module-LWE commitment batch, fused NTT-domain ring mat-vec (BASELINE configs[3]); "
                        f"n={N_RING}, k={K_RANK}, q={Q_MOD}, sigma={SIGMA}, seeded s/e, full-length uniform messages",
            "n": N_RING, "k": K_RANK, "q": Q_MOD, "sigma": SIGMA, "batch_per_gpu": int(batch), "where": where,
            "l2": "inputs+outputs per step exceed the 126 MB L2 (no flush needed)"}


def bind_to_gpu_numa_node(props):
    """Multi-rank runs: keep a rank's threads (and therefore its pinned host buffers, first-touch) on the NUMA node its
    GPU hangs off, so the end-to-end leg does not cross the socket interconnect.  Best effort; returns the node or None."""
    try:
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return None
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return node
    except Exception:
        pass
    return None


# ------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from lambda_snark_r_b200 import api, capi
    import ctypes as C

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    api.set_device(local)
    all_cpus = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(torch.cuda.get_device_properties(local)) if world > 1 else None
    if world > 1:
        # stdout carries the one JSON line: NCCL's own log (it prints its version banner at NCCL_DEBUG >= VERSION) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("WARN", "VERSION"):
            os.environ.pop("NCCL_DEBUG")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms: float) -> float:
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ctx = api.LweContext(api.Params(n=N_RING, k=K_RANK, q=Q_MOD, sigma=SIGMA), seed32=CTX_SEED)
    ntt = api.NttContext(Q_MOD, N_RING)
    if args.arith == "u64":          # comparison runs only: pin both contexts to the u64 Shoup butterflies
        ctx.set_arith(1)
        ntt.set_arith(1)
    words = ctx.words
    B = args.batch
    # weak scaling: every rank owns B commitments; global index = rank*B + i decides seed and message
    g = torch.Generator(device=dev); g.manual_seed(0x5EED + rank)
    msgs = torch.randint(0, Q_MOD, (B, N_RING), device=dev, dtype=torch.int64, generator=g)
    seeds = (torch.arange(B, device=dev, dtype=torch.int64) + (SEED_BASE + rank * B))
    out = torch.empty((B, words), device=dev, dtype=torch.int64)

    def step():
        ctx.commit_batch_device(msgs.data_ptr(), N_RING, seeds.data_ptr(), B, out.data_ptr(), stream)

    def timed(fn, warmup, steps):
        for _ in range(warmup):
            fn()
        barrier()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1))

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_total = timed(step, args.warmup, args.steps)
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)

    # ---- parity of what the timed region produced (outside it): `out` is the last timed step's output; rank 0 puts
    # PARITY_N randomly chosen containers of it next to the CPU oracle's commitments to the same (message, seed)
    parity = None
    if rank == 0 and not args.no_parity:
        from oracle import oracle as O
        prng = np.random.Generator(np.random.PCG64(int(time.time_ns()) & 0xffffffff))
        pick = np.sort(prng.choice(B, size=min(PARITY_N, B), replace=False))
        pt = torch.from_numpy(pick).to(dev)
        got = out.index_select(0, pt).cpu().numpy().view(np.uint64)
        pm = msgs.index_select(0, pt).cpu().numpy().view(np.uint64)
        ps = seeds.index_select(0, pt).cpu().numpy().view(np.uint64)
        want = O.OracleLwe(Q_MOD, N_RING, K_RANK, SIGMA, CTX_SEED).commit_batch(pm, ps, threads=host_threads())
        same = int((got == want).all(axis=1).sum())
        parity = {"containers_compared": int(pick.size), "containers_identical": same,
                  "what": "random containers of the last timed step vs oracle/lsr_oracle.c (bit for bit)"}
        if same != pick.size:
            raise SystemExit(f"bench.py: PARITY FAILURE: {pick.size - same} of {pick.size} containers of the timed step "
                             f"differ from the oracle")

    # ---- K6 on the containers of the timed step: lwe_verify_opening for the whole batch, device-resident (one fused kernel);
    # every opening must verify (outside the timed region: another whole-batch check of the step's output)
    v_diff = torch.empty(B, device=dev, dtype=torch.int64)
    v_inv = torch.empty(B, device=dev, dtype=torch.int32)

    def verify_step():
        ctx.verify_batch_device(out.data_ptr(), msgs.data_ptr(), N_RING, B, v_diff.data_ptr(), v_inv.data_ptr(), stream)

    ms_verify = timed(verify_step, args.warmup, args.steps) / args.steps
    if int(v_diff.count_nonzero().item()) or int(v_inv.count_nonzero().item()):
        raise SystemExit("bench.py: PARITY FAILURE: a commitment of the timed step does not open to its message")

    # ---- NTT sweep point at the same n (BASELINE configs[2])
    NB = args.ntt_batch
    data = torch.randint(0, Q_MOD, (NB, N_RING), device=dev, dtype=torch.int64, generator=g)
    data2 = torch.randint(0, Q_MOD, (NB, N_RING), device=dev, dtype=torch.int64, generator=g)
    ms_fwd = timed(lambda: ntt.forward_device(data.data_ptr(), NB, stream), args.warmup, args.steps) / args.steps
    ms_inv = timed(lambda: ntt.inverse_device(data.data_ptr(), NB, stream), args.warmup, args.steps) / args.steps
    ms_mul = timed(lambda: ntt.mul_pointwise_device(data2.data_ptr(), data.data_ptr(), data2.data_ptr(), NB * N_RING, stream),
                   args.warmup, args.steps) / args.steps
    if parity is not None:
        # one more full-batch launch of each timed NTT call; PARITY_NTT_ROWS rows of it against the oracle
        from oracle import oracle as O
        rows = torch.from_numpy(np.sort(prng.choice(NB, size=min(PARITY_NTT_ROWS, NB), replace=False))).to(dev)
        orc_ntt = O.OracleNtt(Q_MOD, N_RING)
        before = data.index_select(0, rows).cpu().numpy().view(np.uint64)
        ntt.forward_device(data.data_ptr(), NB, stream); torch.cuda.synchronize()
        fwd = data.index_select(0, rows).cpu().numpy().view(np.uint64)
        ntt.inverse_device(data.data_ptr(), NB, stream); torch.cuda.synchronize()
        back = data.index_select(0, rows).cpu().numpy().view(np.uint64)
        b2 = data2.index_select(0, rows).cpu().numpy().view(np.uint64)
        ntt.mul_pointwise_device(data2.data_ptr(), data.data_ptr(), data2.data_ptr(), NB * N_RING, stream); torch.cuda.synchronize()
        prod = data2.index_select(0, rows).cpu().numpy().view(np.uint64)
        ok_rows = int(sum(bool(np.array_equal(f, orc_ntt.forward(x)) and np.array_equal(y, x) and
                               np.array_equal(pr, orc_ntt.mul_pointwise(x, m2)))
                          for x, f, y, m2, pr in zip(before, fwd, back, b2, prod)))
        parity.update({"ntt_rows_compared": int(rows.numel()), "ntt_rows_identical": ok_rows,
                       "ntt_what": "rows of one extra full-batch forward / inverse / pointwise launch (same calls as the timed ones) "
                                   "vs the oracle: forward values, inverse(forward) = input, pointwise products"})
        if ok_rows != rows.numel():
            raise SystemExit("bench.py: PARITY FAILURE in the NTT block")

    # ---- BASELINE configs[2] in brief: ring degrees 2^10 .. 2^16 at 2^26 coefficients per launch (512 MiB > L2), forward and
    # inverse; q = 17592169062401 has 2-adicity 13 (n <= 4096), above that 17592180539393.  tools/sweep.py is the full grid.
    sweep = []
    if args.sweep:
        for logn in range(10, 17):
            sn, sq = 1 << logn, (Q_MOD if logn <= 12 else 17592180539393)
            sctx = ntt if logn == 12 else api.NttContext(sq, sn)
            sb = (1 << 26) >> logn
            sd = data[: (sb * sn) // N_RING].view(sb, sn) if NB * N_RING >= sb * sn else torch.randint(
                0, sq, (sb, sn), device=dev, dtype=torch.int64, generator=g)
            tf = timed(lambda: sctx.forward_device(sd.data_ptr(), sb, stream), 3, 5) / 5
            ti = timed(lambda: sctx.inverse_device(sd.data_ptr(), sb, stream), 3, 5) / 5
            sweep.append((logn, sb, tf, ti))
            if sctx is not ntt:
                sctx.close()

    # ---- end to end through the host-pointer C ABI, pinned host buffers
    EB = args.e2e_batch
    h_msgs = torch.randint(0, Q_MOD, (EB, N_RING), dtype=torch.int64).pin_memory()
    h_seeds = (torch.arange(EB, dtype=torch.int64) + (SEED_BASE + rank * EB)).pin_memory()
    h_out = torch.empty((EB, words), dtype=torch.int64).pin_memory()

    def e2e_step():
        ctx.commit_batch_ptr(h_msgs.data_ptr(), N_RING, h_seeds.data_ptr(), EB, h_out.data_ptr())

    e2e_steps = max(1, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    e2e_value = world * EB / (e2e_ms * 1e-3)
    # N > 1, same job with dynamic sharding: the ranks pull chunks of the N * EB * steps commitments from one shared counter
    # instead of each taking a fixed share.  On a host whose GPUs do not get equal DMA bandwidth (the 8-GPU node: four GPUs
    # at half the share of the other four, profiles/r02_multi_gpu_probe_8gpu.txt) the fixed split waits for the slowest
    # rank; a dispatcher that feeds GPUs as they free up does not.  `e2e.value` stays the fixed (weak-scaling) split.
    e2e_dyn = None
    if world > 1:
        import fcntl
        DCH = 1024
        total_chunks = world * (EB // DCH) * e2e_steps
        path = f"/tmp/lsr_bench_counter_{os.environ.get('MASTER_PORT', '0')}"
        if rank == 0:
            with open(path, "wb") as f:
                f.write((0).to_bytes(8, "little"))
        barrier()
        fd = os.open(path, os.O_RDWR)

        def pull():
            fcntl.flock(fd, fcntl.LOCK_EX)
            os.lseek(fd, 0, os.SEEK_SET)
            c = int.from_bytes(os.read(fd, 8), "little")
            if c < total_chunks:
                os.lseek(fd, 0, os.SEEK_SET)
                os.write(fd, (c + 1).to_bytes(8, "little"))
            fcntl.flock(fd, fcntl.LOCK_UN)
            return c

        barrier()
        t0 = time.perf_counter()
        mine = 0
        while True:
            c = pull()
            if c >= total_chunks:
                break
            slot = (c % (EB // DCH)) * DCH                       # this rank's buffers, slot by slot
            ctx.commit_batch_ptr(h_msgs.data_ptr() + slot * N_RING * 8, N_RING, h_seeds.data_ptr() + slot * 8, DCH,
                                 h_out.data_ptr() + slot * words * 8)
            mine += 1
        torch.cuda.synchronize()
        dyn_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        os.close(fd)
        counts = torch.tensor([mine], device=dev, dtype=torch.int64)
        allc = torch.empty(world, device=dev, dtype=torch.int64)
        dist.all_gather_into_tensor(allc, counts)
        e2e_dyn = {"what": "the same commitments pulled in chunks of 1024 from one shared counter (dynamic sharding over the ranks)",
                   "value": total_chunks * DCH / (dyn_ms * 1e-3), "unit": "commitments/s", "ms_total": dyn_ms,
                   "chunks_per_rank": allc.cpu().tolist()}
        barrier()
        if rank == 0:
            os.unlink(path)
    # the same bytes as raw page-locked copies (H2D of the messages and D2H of the containers at the same time, all ranks
    # at once): what the host's DMA path sustains for this mix, the ceiling of any end-to-end path with this container format
    d_raw_in = torch.empty((EB, N_RING), dtype=torch.int64, device=dev)
    d_raw_out = out[:EB] if B >= EB else torch.empty((EB, words), dtype=torch.int64, device=dev)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()

    def raw_copies():
        with torch.cuda.stream(s_in):
            d_raw_in.copy_(h_msgs, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_out.copy_(d_raw_out, non_blocking=True)

    raw_copies(); barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        raw_copies()
    torch.cuda.synchronize()
    raw_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    del d_raw_in
    # ---- forward NTT end to end: ntt_forward_batch on page-locked host polynomials (in place: 32 KiB in, 32 KiB out each)
    h_polys = torch.randint(0, Q_MOD, (EB, N_RING), dtype=torch.int64).pin_memory()

    def ntt_e2e_step():
        if capi.load().ntt_forward_batch(ntt.handle, C.c_void_p(h_polys.data_ptr()), EB) != 0:
            raise RuntimeError("ntt_forward_batch failed")

    for _ in range(2):
        ntt_e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ntt_e2e_step()
    torch.cuda.synchronize()
    ntt_e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / e2e_steps
    del h_polys
    # ---- prover commitment phase (BASELINE configs[4]): 2^20-constraint R1CS over Goldilocks, quotient polynomial on
    # the device, cut into m / n ring elements, each committed; weak scaling over witnesses (every rank proves its own)
    prover = None
    if args.prover_logm > 0:
        PM = 1 << args.prover_logm
        PW = args.prover_witnesses
        GOLD = 2**64 - 2**32 + 1
        rng = np.random.Generator(np.random.PCG64(0xC5 + rank))
        ga = rng.integers(0, GOLD, size=PM, dtype=np.uint64)
        gb = rng.integers(0, GOLD, size=PM, dtype=np.uint64)
        gc = np.fromiter(((int(x) * int(y)) % GOLD for x, y in zip(ga.tolist(), gb.tolist())), dtype=np.uint64, count=PM)
        z = np.zeros(3 * PM + 1, dtype=np.uint64); z[0] = 1
        z[1::3], z[2::3], z[3::3] = ga, gb, gc              # gates z[3i+1] * z[3i+2] = z[3i+3] (tests/integration_matrix.rs:60-75)
        rows = np.arange(PM, dtype=np.uint32); one = np.ones(PM, dtype=np.uint64)
        r1cs = api.R1CS.from_arrays(PM, 3 * PM + 1, (rows, 3 * rows + 1, one), (rows, 3 * rows + 2, one),
                                    (rows, 3 * rows + 3, one), GOLD)
        chunks = r1cs.quotient_chunks(ctx)
        zs = torch.from_numpy(np.tile(z.view(np.int64), (PW, 1))).to(dev)
        pseeds = torch.arange(1, PW * chunks + 1, dtype=torch.int64, device=dev) + rank * PW * chunks
        pout = torch.empty((PW, chunks, words), dtype=torch.int64, device=dev)

        def prover_step(hi=chunks):
            st = r1cs.commit_quotient_device(ctx, zs.data_ptr(), PW, pseeds.data_ptr(), pout.data_ptr(), 0, hi)
            assert not st.any()

        psteps = max(3, min(args.steps, 10))
        ms_p = timed(prover_step, args.warmup, psteps) / psteps
        ms_q = timed(lambda: prover_step(0), args.warmup, psteps) / psteps
        h_z = torch.from_numpy(np.tile(z.view(np.int64), (PW, 1))).pin_memory()
        h_ps = pseeds.cpu().pin_memory()
        h_pout = torch.empty((PW, chunks, words), dtype=torch.int64).pin_memory()
        lib = capi.load()
        st_buf = (C.c_int * PW)()

        def prover_e2e():
            rc = lib.lsr_prover_commit_quotient(r1cs._h, ctx.as_ptr(), C.cast(h_z.data_ptr(), capi.u64p), 3 * PM + 1, PW, 0,
                                                C.cast(h_ps.data_ptr(), capi.u64p), 0, chunks,
                                                C.cast(h_pout.data_ptr(), capi.u64p), st_buf)
            assert rc == 0
        prover_e2e()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            prover_e2e()
        pe2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3) / 3
        prover = {"workload": f"prover commitment phase (BASELINE configs[4]): R1CS with 2^{args.prover_logm} multiplication gates over "
                              f"Goldilocks, quotient polynomial -> {chunks} ring elements of n={N_RING} -> {chunks} commitments per witness",
                  "witnesses_per_gpu": PW, "constraints": PM, "commitments_per_witness": chunks,
                  "value": world * PW / (ms_p * 1e-3), "unit": "witnesses/s",
                  "constraints_per_s": world * PW * PM / (ms_p * 1e-3),
                  "commitments_per_s": world * PW * chunks / (ms_p * 1e-3),
                  "ms_per_step": ms_p, "ms_quotient": ms_q, "ms_commit": ms_p - ms_q,
                  "quotient_bytes_per_constraint": QUOTIENT_BYTES_PER_CONSTRAINT,
                  "quotient_GBps": PW * PM * QUOTIENT_BYTES_PER_CONSTRAINT / (ms_q * 1e-3) / 1e9,
                  "e2e": {"value": world * PW / (pe2e_ms * 1e-3), "unit": "witnesses/s", "ms_per_step": pe2e_ms,
                          "h2d_bytes_per_step": PW * ((3 * PM + 1) + chunks) * 8, "d2h_bytes_per_step": PW * chunks * words * 8,
                          "api": "lsr_prover_commit_quotient (C ABI, pinned host buffers)"}}
        # CPU baseline for this block: the C restatement of ntt.rs / r1cs.rs (oracle/lsr_oracle_quotient.c), one witness,
        # one host thread -- the reference's own quotient is an O(m^2) schoolbook product (r1cs.rs:846-863) and does not
        # finish at this size, so the oracle's O(m log m) form is the generous baseline
        if rank == 0 and args.cpu_seconds > 0:
            from oracle import oracle as O
            w1 = api.reference_root_of_unity(GOLD, PM)
            w2 = api.reference_root_of_unity(GOLD, 2 * PM)
            tri = [(rows, 3 * rows + 1 + j, one) for j in range(3)]
            t0 = time.perf_counter()
            _, st_cpu = O.r1cs_quotient(PM, 3 * PM + 1, tri[0], tri[1], tri[2], z, GOLD, w1, w2)
            dt_q = time.perf_counter() - t0
            assert st_cpu == 0
            prover["cpu_baseline"] = {"value": 1.0 / dt_q, "unit": "quotients/s", "cores": 1, "kind": "port",
                                      "sample": f"1 quotient of 2^{args.prover_logm} constraints in {dt_q:.2f}s, oracle C port (O(m log m)); "
                                                f"commitments of the phase are covered by the top-level cpu_baseline",
                                      "gpu_quotients_per_s": PW / (ms_q * 1e-3)}
        # one proof, strong scaling (BASELINE configs[4] as written: ONE 2^20-constraint witness over N GPUs): every rank
        # computes the quotient (it does not shard: six dependent size-m transforms) and commits ITS slice of the
        # `chunks` units; at N > 1 the containers land in rank 0's HBM through peer stores (no collective)
        lo, hi = (rank * chunks) // world, ((rank + 1) * chunks) // world
        sw_pg = None
        sw_out = pout.data_ptr()
        if world > 1:
            from lambda_snark_r_b200 import gather as G
            sw_pg = G.PeerGather(rank, world, ((chunks + world - 1) // world) * words * 8, G.torch_bcast())
            sw_out = sw_pg.slice_ptr

        def single_witness_step():
            st = r1cs.commit_quotient_device(ctx, zs.data_ptr(), 1, pseeds.data_ptr(), sw_out, lo, hi)
            assert not st.any()

        ms_sw = timed(single_witness_step, args.warmup, psteps) / psteps
        ms_sw_q = timed(lambda: r1cs.commit_quotient_device(ctx, zs.data_ptr(), 1, pseeds.data_ptr(), sw_out, lo, lo),
                        args.warmup, psteps) / psteps
        prover["single_witness"] = {"what": "latency of the commitment phase of ONE 2^20-constraint proof: quotient on every rank "
                                            "(replicated), units [r*C/N, (r+1)*C/N) committed by rank r"
                                            + (", containers stored into rank 0's HBM over NVLink" if world > 1 else ""),
                                    "ms": ms_sw, "ms_quotient_replicated": ms_sw_q, "ms_commit_slice": ms_sw - ms_sw_q,
                                    "units_per_rank": hi - lo, "proofs_per_s": 1e3 / ms_sw}
        if sw_pg:
            sw_pg.close()
        del zs, pout
        r1cs.close()

    # the timed regions are a few tens of ms each: nvidia-smi samples every 100 ms, so the sampler stays on across all of
    # them (commit, NTT, e2e) and the loop below keeps the commit kernel running until it has seen >= 10 samples
    if rank == 0:
        t_end = time.perf_counter() + 1.2
        while time.perf_counter() < t_end:
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None

    # ---- final gather over NVLink / NVSwitch, for the WHOLE step batch, inside the timed step
    #  (a) containers -> rank 0's HBM, fused: the commitment kernel of every rank stores its container rows straight into
    #      rank 0's buffer through peer memory (lambda_snark_r_b200/gather.py): no collective follows the kernel.  Bounded by
    #      rank 0's NVLink ingress: (N-1) * B * 64 KiB per step.
    #  (b) digests: SHA3-256 of every container on its owner (the Fiat-Shamir transcript hash of N2 -- work a single GPU does
    #      too), then an NCCL all-gather of 32 bytes per commitment; the containers stay with their owners.
    gather = None
    if world > 1:
        from lambda_snark_r_b200 import gather as G
        lib = capi.load()
        pg = G.PeerGather(rank, world, B * words * 8, G.torch_bcast())

        def step_peer():
            ctx.commit_batch_device(msgs.data_ptr(), N_RING, seeds.data_ptr(), B, pg.slice_ptr, stream)

        ms_peer = timed(step_peer, args.warmup, args.steps) / args.steps
        # checksum of checksums: every rank's slice in rank 0's buffer equals what the rank computed locally
        local_sum = out.sum(dim=1).sum().reshape(1)
        sums = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(sums, local_sum)
        gathered_ok = None
        if rank == 0:
            view = G.device_view(pg.base, world * B * words, dev).view(world, B * words)
            gathered_ok = bool(torch.equal(view.sum(dim=1), sums))
            if not gathered_ok:
                raise SystemExit("bench.py: the peer-memory gather does not hold what the ranks computed")
        d_ab = torch.empty((B, 2), dtype=torch.int64, device=dev)
        d_hash = torch.empty((B, 2, 4), dtype=torch.int64, device=dev)
        d_all = torch.empty((world * B, 2, 4), dtype=torch.int64, device=dev)
        d_pub = torch.zeros(8, dtype=torch.int64, device=dev)

        def transcript():
            if lib.lsr_fs_challenge_batch_device(d_pub.data_ptr(), 0, out.data_ptr(), words, B, Q_MOD, 0, d_ab.data_ptr(),
                                                 d_hash.data_ptr(), stream) != 0:
                raise RuntimeError("lsr_fs_challenge_batch_device failed")

        def step_digest_gather():
            step(); dist.all_gather_into_tensor(d_all, d_hash)

        def step_transcript_gather():
            step(); transcript(); dist.all_gather_into_tensor(d_all, d_hash)

        transcript()
        ms_dg = timed(step_digest_gather, args.warmup, args.steps) / args.steps
        ms_tg = timed(step_transcript_gather, args.warmup, max(3, args.steps // 2)) / max(3, args.steps // 2)
        # the NCCL all-gather of whole containers the first round used, for comparison (one eighth of the step batch)
        GB_ = min(B, 2048)
        part = out[:GB_].contiguous()
        full = torch.empty((world * GB_, words), device=dev, dtype=torch.int64)
        ms_ag = timed(lambda: dist.all_gather_into_tensor(full, part), 2, 3) / 3
        del full
        gather = {
            "containers_to_rank0_peer_stores": {
                "what": "fused: every rank's commitment kernel writes its containers into rank 0's HBM over NVLink (no collective)",
                "commitments_per_rank": B, "ms_per_step": ms_peer, "value_with_gather": world * B / (ms_peer * 1e-3),
                "frac_of_value": (ms_step / ms_peer), "rank0_ingress_GBps": (world - 1) * B * words * 8 / (ms_peer * 1e-3) / 1e9,
                "nvlink_peak_GBps_per_direction": 900.0, "verified": gathered_ok},
            "digests_all_gather": {
                "what": "owner hashes its containers (SHA3-256 transcript, N2), NCCL all-gather of 32 B per commitment",
                "commitments_per_rank": B, "ms_per_step_commit_plus_gather": ms_dg,
                "value_with_gather": world * B / (ms_dg * 1e-3), "frac_of_value": ms_step / ms_dg,
                "ms_per_step_commit_plus_transcript_plus_gather": ms_tg,
                "value_with_transcript_and_gather": world * B / (ms_tg * 1e-3)},
            "containers_nccl_all_gather": {"commitments_per_rank": GB_, "ms": ms_ag,
                                           "GBps_in_per_rank": (world - 1) * GB_ * words * 8 / (ms_ag * 1e-3) / 1e9},
        }
        pg.close()

    # ---- integer roofline denominator (measured on this GPU)
    imad_wide = C.c_double(0); imad_lo = C.c_double(0); mhz = C.c_double(0)
    capi.load().lsr_measure_imad_peak(0, C.byref(imad_lo), C.byref(mhz))     # IMAD (32-bit result): 64 lanes/clk/SM
    capi.load().lsr_measure_imad_peak(1, C.byref(imad_wide), None)           # IMAD.WIDE (64-bit result): 32 lanes/clk/SM
    imad_peak = imad_lo.value
    # a Shoup modmul on 2-limb residues needs 5 full 32x32->64 products + 4 low products; with IMAD.WIDE at half
    # rate that is 14 issue slots of the IMAD pipe, not the 10 the SURVEY normalisation assumes
    wide_cost = imad_lo.value / imad_wide.value if imad_wide.value else 2.0
    slots_per_modmul = 5 * wide_cost + 4
    fp64_peak = C.c_double(0)
    capi.load().lsr_measure_fp64_peak(C.byref(fp64_peak))                     # DFMA/DADD/DMUL: 64 lanes/clk/SM
    fp64_peak = fp64_peak.value
    arith = "fp64" if ntt.arith == 2 else "u64"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    hbm_peak, hbm_src = peaks()
    tr = traffic_note()
    commit_gbs = B * ALG_BYTES_COMMIT / (ms_step * 1e-3) / 1e9
    commit_imad = B * MODMUL_COMMIT * IMAD_PER_MODMUL / (ms_step * 1e-3) / 1e9
    commit_fp64 = B * FP64_COMMIT / (ms_step * 1e-3) / 1e9

    def ntt_block(ms, fp64_inst, extra_modmul=0):
        rate = world * NB / (ms * 1e-3)
        per_gpu = NB / (ms * 1e-3)
        gbs = NB * ALG_BYTES_NTT / (ms * 1e-3) / 1e9
        gimad = NB * (BUTTERFLIES_NTT + extra_modmul) * IMAD_PER_MODMUL / (ms * 1e-3) / 1e9
        gfp64 = per_gpu * fp64_inst / 1e9
        bounds = {"hbm": hbm_peak * 1e9 / ALG_BYTES_NTT, "fp64": fp64_peak * 1e9 / fp64_inst if arith == "fp64" else None,
                  "imad_survey": imad_peak * 1e9 / ((BUTTERFLIES_NTT + extra_modmul) * IMAD_PER_MODMUL)}
        live = {k: v for k, v in bounds.items() if v and (k != "imad_survey" or arith != "fp64")}
        slow = min(live, key=live.get)
        return {"value": rate, "unit": "NTT/s", "ms_per_step": ms,
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                             "fp64_achieved_ginst_s": gfp64 if arith == "fp64" else None, "fp64_peak_ginst_s": fp64_peak,
                             "fp64_frac": gfp64 / fp64_peak if (arith == "fp64" and fp64_peak) else None,
                             "bounds_ntt_per_s": bounds, "slower_bound": slow, "frac_of_slower_bound": per_gpu / live[slow],
                             "imad_survey_frac": gimad / imad_peak if imad_peak else None}}

    # CPU baseline: bounded sample of the same workload on this box's host cores (all of them again)
    os.sched_setaffinity(0, all_cpus)
    cpu_rate, cpu_threads, cpu_count, cpu_dt, native = cpu_commit_rate(args.cpu_seconds)
    ntt_cpu_rate, _, ntt_cpu_count, ntt_cpu_dt = cpu_ntt_rate(min(args.cpu_seconds, 5.0))

    mul_gbs = NB * N_RING * 24 / (ms_mul * 1e-3) / 1e9
    fwd_block, inv_block = ntt_block(ms_fwd, FP64_NTT_FWD), ntt_block(ms_inv, FP64_NTT_INV, N_RING // 2)
    # The roof that binds the fused kernel is the one its multiplications run on: the FP64 pipe (arith fp64: 8 FP64
    # instructions per butterfly at 64 lanes/clk/SM, measured on this GPU in this run) or, with --arith u64, the IMAD pipe
    # under SURVEY 8d's normalisation.  HBM is reported beside it (`hbm`), not as the headline: the kernel moves 98 312 B
    # per commitment and sits at ~0.14 of the copy peak by construction.
    if arith == "fp64":
        bound = {"bound": "fp64", "achieved": commit_fp64, "peak": fp64_peak, "unit": "Ginst/s",
                 "frac": commit_fp64 / fp64_peak if fp64_peak else None,
                 "peak_source": "lsr_measure_fp64_peak on this GPU in this run (dependent-free DFMA chains, 64 lanes/clk/SM)",
                 "model": f"{FP64_COMMIT} FP64 instructions/commitment = {K_RANK} fwd + {K_RANK} inv NTT x {FP64_PER_BUTTERFLY}/butterfly"
                          f" + {K_RANK * K_RANK * N_RING} mat-vec MACs x {FP64_PER_MODMUL + 1}; sampler (ALU pipe) not counted"}
    else:
        bound = {"bound": "imad", "achieved": commit_imad, "peak": imad_peak, "unit": "GIMAD/s",
                 "frac": commit_imad / imad_peak if imad_peak else None,
                 "peak_source": "lsr_measure_imad_peak on this GPU in this run (64 lanes/clk/SM)",
                 "model": f"{MODMUL_COMMIT} modmul/commitment x {IMAD_PER_MODMUL} IMAD (SURVEY 8d normalisation)"}
    tr_per = tr.get("fused_commit_bytes_per_commitment")
    roofline = {
        "kernel": "fused_commit_kernel<12,2,0,%s>" % ("POL_F64" if arith == "fp64" else "POL_LAZY"), **bound,
        "arith": arith,
        "traffic": tr_per * B if tr_per else None,
        "traffic_source": "static: dram bytes per commitment of the ncu --set full capture recorded in profiles/traffic.json x batch "
                          "(not measured in this run)" if tr_per else None,
        "algorithmic_bytes_per_commitment": ALG_BYTES_COMMIT,
        "hbm": {"achieved": commit_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": commit_gbs / hbm_peak, "peak_source": hbm_src},
        "imad_survey": {"frac": commit_imad / imad_peak if imad_peak else None,
                        "model": f"{MODMUL_COMMIT} modmul/commitment x {IMAD_PER_MODMUL} IMAD (SURVEY 8d normalisation; with "
                                 f"arith=fp64 the multiplications run on the FP64 pipe instead); IMAD {imad_lo.value:.0f} GIMAD/s"
                                 f" (= 64 lanes/clk/SM at {mhz.value:.0f} MHz), IMAD.WIDE {imad_wide.value:.0f} GIMAD/s measured"},
        "bounds_commitments_per_s": {"hbm": hbm_peak * 1e9 / ALG_BYTES_COMMIT,
                                     "fp64": fp64_peak * 1e9 / FP64_COMMIT if fp64_peak else None,
                                     "imad_survey": imad_peak * 1e9 / (MODMUL_COMMIT * IMAD_PER_MODMUL) if imad_peak else None},
        # the other half of BASELINE.json's metric (batched NTTs/s at n = 4096) and the other kernels of the path, one
        # line each: value per GPU, fraction of the slower of its compute and HBM roofs
        "kernels": {
            "ntt_forward": {"value": NB / (ms_fwd * 1e-3), "unit": "NTT/s", "batch": NB,
                            "bound": fwd_block["roofline"]["slower_bound"], "frac": fwd_block["roofline"]["frac_of_slower_bound"],
                            "hbm_frac": fwd_block["roofline"]["frac"]},
            "ntt_inverse": {"value": NB / (ms_inv * 1e-3), "unit": "NTT/s", "batch": NB,
                            "bound": inv_block["roofline"]["slower_bound"], "frac": inv_block["roofline"]["frac_of_slower_bound"],
                            "hbm_frac": inv_block["roofline"]["frac"]},
            "pointwise": {"value": NB * N_RING / (ms_mul * 1e-3), "unit": "coefficients/s", "bound": "hbm",
                          "frac": mul_gbs / hbm_peak},
            # K6: k - 1 forward + 1 inverse transform + (k - 1) n products per opening; 8 k n + 8 n bytes in
            "verify_opening": {"value": B / (ms_verify * 1e-3), "unit": "openings/s", "bound": "fp64" if arith == "fp64" else "imad",
                               "frac": (B / (ms_verify * 1e-3)) * ((K_RANK - 1) * FP64_NTT_FWD + FP64_NTT_INV
                                                                   + (K_RANK - 1) * N_RING * (FP64_PER_MODMUL + 1)) / 1e9 / fp64_peak
                               if (arith == "fp64" and fp64_peak) else None,
                               "hbm_frac": B * (8 * K_RANK * N_RING + 8 * N_RING) / (ms_verify * 1e-3) / 1e9 / hbm_peak,
                               "all_openings_of_the_timed_step_verify": True},
        },
    }
    if sweep:
        # per ring degree: NTT/s per GPU and the fraction of the slower of the two roofs (FP64 pipe: 8 instructions per
        # butterfly, + 6 per coefficient pair of the inverse's n^-1 stage; HBM: 16 B per coefficient, one round trip)
        rows = []
        for logn, sb, tf, ti in sweep:
            sn = 1 << logn
            bf = (sn // 2) * logn
            roofs_f = {"hbm": hbm_peak * 1e9 / (16 * sn), "fp64": fp64_peak * 1e9 / (bf * FP64_PER_BUTTERFLY)}
            roofs_i = {"hbm": roofs_f["hbm"], "fp64": fp64_peak * 1e9 / (bf * FP64_PER_BUTTERFLY + (sn // 2) * FP64_PER_MODMUL)}
            if arith != "fp64":
                roofs_f = {"hbm": roofs_f["hbm"], "imad_survey": imad_peak * 1e9 / (bf * IMAD_PER_MODMUL)}
                roofs_i = {"hbm": roofs_f["hbm"], "imad_survey": imad_peak * 1e9 / ((bf + sn // 2) * IMAD_PER_MODMUL)}
            bfk, bik = min(roofs_f, key=roofs_f.get), min(roofs_i, key=roofs_i.get)
            rows.append({"n": sn, "batch": sb, "fwd_per_s": sb / (tf * 1e-3), "inv_per_s": sb / (ti * 1e-3),
                         "fwd_bound": bfk, "fwd_frac": sb / (tf * 1e-3) / roofs_f[bfk],
                         "inv_bound": bik, "inv_frac": sb / (ti * 1e-3) / roofs_i[bik],
                         "fwd_hbm_frac": sb / (tf * 1e-3) / roofs_f["hbm"]})
        roofline["kernels"]["ntt_sweep"] = rows
    if prover:
        prover["quotient_hbm_frac"] = prover["quotient_GBps"] / hbm_peak
        roofline["kernels"]["quotient_pipeline"] = {
            "value": PW * PM / (prover["ms_quotient"] * 1e-3), "unit": "constraints/s", "bound": "hbm",
            "frac": prover["quotient_hbm_frac"], "ms": prover["ms_quotient"],
            "model": f"{QUOTIENT_BYTES_PER_CONSTRAINT} algorithmic bytes per constraint, 2^{args.prover_logm} constraints x {PW} witnesses"}
    e2e = {"value": e2e_value, "unit": "commitments/s", "h2d_bytes_per_step": EB * (N_RING + 1) * 8,
           "d2h_bytes_per_step": EB * words * 8, "batch_per_gpu": EB, "ms_per_step": e2e_ms,
           "api": "lwe_commit_batch (C ABI, pinned host buffers)", "numa_node": numa,
           "raw_copy_ceiling": {"what": "the step's H2D + D2H bytes as bare page-locked copies on two streams, all ranks at once, "
                                        "slowest rank (no kernels): the host's DMA ceiling for this byte mix",
                                "ms_per_step": raw_ms, "commitments_per_s_equivalent": world * EB / (raw_ms * 1e-3),
                                "GBps_total": world * EB * (N_RING + words) * 8 / (raw_ms * 1e-3) / 1e9},
           "frac_of_raw_copy_ceiling": raw_ms / e2e_ms,
           "dynamic_sharding": e2e_dyn,
           "ntt_forward": {"value": world * EB / (ntt_e2e_ms * 1e-3), "unit": "NTT/s",
                           "api": "ntt_forward_batch (C ABI, pinned host buffers, in place)"}}
    if prover:
        e2e["prover"] = {"value": prover["e2e"]["value"], "unit": "witnesses/s", "api": "lsr_prover_commit_quotient"}
    line = {
        "metric": "lwe_commitments_per_sec", "value": value, "unit": "commitments/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(B, "hbm-resident"),
        "roofline": roofline,
        "cpu_baseline": {"value": cpu_rate, "unit": "commitments/s", "cores": cpu_threads, "kind": "port",
                         "sample": f"{cpu_count} commitments in {cpu_dt:.1f}s, oracle C port, {cpu_threads} OpenMP threads"
                                   f"{' -march=native' if native else ''}",
                         "ntt_forward_per_s": ntt_cpu_rate},
        "e2e": e2e,
        "gpu_launches": args.steps,
        "clocks": clocks,
        "parity_checked": (parity or {}).get("containers_identical", 0) + (parity or {}).get("ntt_rows_identical", 0),
        "parity": parity,
        "ntt": {"batch_per_gpu": NB, "forward": fwd_block, "inverse": inv_block,
                "e2e": {"value": world * EB / (ntt_e2e_ms * 1e-3), "unit": "NTT/s", "ms_per_step": ntt_e2e_ms, "batch_per_gpu": EB,
                        "h2d_bytes_per_step": EB * N_RING * 8, "d2h_bytes_per_step": EB * N_RING * 8,
                        "api": "ntt_forward_batch (C ABI, pinned host buffers, in place)"},
                "pointwise": {"value": world * NB * N_RING / (ms_mul * 1e-3), "unit": "coefficients/s", "ms_per_step": ms_mul,
                              "roofline": {"bound": "hbm", "achieved": mul_gbs, "peak": hbm_peak, "unit": "GB/s",
                                           "frac": mul_gbs / hbm_peak}}},
    }
    if prover:
        line["prover"] = prover
    if gather:
        line["gather"] = gather
        best = max(gather["containers_to_rank0_peer_stores"]["value_with_gather"], gather["digests_all_gather"]["value_with_gather"])
        line["value_with_gather"] = best
        line["config"]["gather"] = ("value excludes the final gather; value_with_gather = the better of the two in-step forms under "
                                    "`gather` (containers into rank 0's HBM by peer stores / digest all-gather)")
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16384, help="commitments per GPU per step")
    ap.add_argument("--ntt-batch", type=int, default=16384, help="polynomials per GPU per NTT step")
    ap.add_argument("--e2e-batch", type=int, default=8192)
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--prover-logm", type=int, default=20, help="log2 constraints of the prover-phase block (0 = skip)")
    ap.add_argument("--prover-witnesses", type=int, default=4, help="witnesses per GPU per prover step")
    ap.add_argument("--no-sweep", dest="sweep", action="store_false", help="skip the ring-degree sweep (configs[2] in brief)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle comparison of the timed step's output")
    ap.add_argument("--arith", default="auto", choices=["auto", "u64"],
                    help="auto: FP64-pipe butterflies (exact for q < 2^45); u64: integer Shoup butterflies (comparison)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_gpu(args)


if __name__ == "__main__":
    sys.exit(main())
