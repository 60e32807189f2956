"""Summarise an .ncu-rep: key raw metrics + top stall instructions (source page).
usage: python tools/ncu_summary.py report.ncu-rep [--top N]"""
import csv
import io
import subprocess
import sys
from collections import Counter

rep = sys.argv[1]
top_n = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 18

KEYS = ["gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__cycles_elapsed.max",
        "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"]


def run(page):
    out = subprocess.run(["ncu", "-i", rep, "--page", page, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


raw = run("raw")
hdr, units = raw[0], raw[1]
for row in raw[2:]:
    name = row[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
    print("==", name[:110])
    d = dict(zip(hdr, zip(units, row)))
    for k in KEYS:
        if k in d:
            print(f"  {k:86s} {d[k][1]:>16s} {d[k][0]}")
    stalls = sorted(((float(v[1]), k) for k, v in d.items() if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("per_issue_active.ratio")), reverse=True)
    print("  stalls/issue:", ", ".join(f"{k.split('stalled_')[1].split('_per_')[0]}={v:.2f}" for v, k in stalls[:7]))

src = run("source")
sections, cur = [], None
for r in src:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1] if len(r) > 1 else "?", "hdr": None, "rows": []}
        sections.append(cur)
    elif cur is not None and cur["hdr"] is None:
        cur["hdr"] = r
    elif cur is not None and len(r) == len(cur["hdr"]):
        cur["rows"].append(r)
for sec in sections:
    h, rows = sec["hdr"], sec["rows"]
    ix = {n: i for i, n in enumerate(h)}
    tot = sum(int(r[ix["# Samples"]]) for r in rows)
    print(f"== source of {sec['name'][:90]}: {len(rows)} SASS instructions, {tot} samples")
    for r in sorted(rows, key=lambda r: -int(r[ix["# Samples"]]))[:top_n]:
        print(f"  {int(r[ix['# Samples']]):7d} {100.0 * int(r[ix['# Samples']]) / max(tot, 1):5.1f}%  {r[ix['Source']].strip()[:100]}")
    ops = Counter()
    execd = Counter()
    for r in rows:
        toks = [t for t in r[ix["Source"]].split() if not t.startswith("@")]
        op = toks[0].split(".")[0] if toks else "?"
        ops[op] += int(r[ix["# Samples"]])
        execd[op] += int(r[ix["Instructions Executed"]])
    te = sum(execd.values())
    print("  samples by opcode:", ops.most_common(10))
    print("  executed by opcode:", [(k, f"{100.0 * v / max(te, 1):.1f}%") for k, v in execd.most_common(16)])
    print(f"  total executed warp-instructions {te}")
