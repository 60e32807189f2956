"""Executed warp-instructions and stall samples per source line of an .ncu-rep captured with
--import-source on (kernels are built with -lineinfo).
usage: python tools/ncu_by_line.py report.ncu-rep [--top N]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top_n = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = "?"
hdr = None
per_line = defaultdict(lambda: [0, 0, ""])
per_file = defaultdict(lambda: [0, 0])
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = {n: i for i, n in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr) or not r[0]:
        continue
    def num(x):
        try:
            return int(x)
        except ValueError:
            return 0
    ex = num(r[hdr["Instructions Executed"]])
    sm = num(r[hdr["# Samples"]])
    k = (cur_file, int(r[0]))
    per_line[k][0] += ex
    per_line[k][1] += sm
    per_line[k][2] = r[1].strip()[:90]
    per_file[cur_file][0] += ex
    per_file[cur_file][1] += sm
tot_ex = sum(v[0] for v in per_file.values()) or 1
tot_sm = sum(v[1] for v in per_file.values()) or 1
print(f"total executed warp-instructions {tot_ex}, samples {tot_sm}")
for f, (ex, sm) in sorted(per_file.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:28s} exec {100.0 * ex / tot_ex:5.1f}%  samples {100.0 * sm / tot_sm:5.1f}%")
print("top lines by executed instructions:")
for (f, ln), (ex, sm, src) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"  {f:22s}:{ln:4d} exec {100.0 * ex / tot_ex:5.1f}%  smp {100.0 * sm / tot_sm:5.1f}%  {src}")
