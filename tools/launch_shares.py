"""Per-kernel share of the GPU time of one `ncu --metrics gpu__time_duration.sum --csv` launch list.
usage: python tools/launch_shares.py launches.csv "<command that was profiled>" """
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot = defaultdict(lambda: [0, 0.0])
peak = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0]
    if "_peak_kernel" in name:          # roofline denominators (lsr_measure_*_peak), run after the timed regions
        peak[name][0] += 1
        peak[name][1] += float(r[14].replace(",", "")) / 1e3
        continue
    tot[name][0] += 1
    tot[name][1] += float(r[14].replace(",", "")) / 1e3
s = sum(v[1] for v in tot.values()) or 1.0
print(f"ncu --metrics gpu__time_duration.sum --clock-control none, {sys.argv[2] if len(sys.argv) > 2 else ''} "
      f"(cold-cache, serialised: compare shares)")
for k, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:4d} launches {us:10.1f} us {100 * us / s:6.1f}%  {k[:90]}")
print("not part of any step (roofline-denominator microbenchmarks, run after the timed regions):")
for k, (n, us) in sorted(peak.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:4d} launches {us:10.1f} us          {k[:90]}")
