#!/bin/bash
# tools/thp_probe_all.sh N OUT : tools/thp_probe.py from N processes at once, 4 KiB pages then transparent huge pages
N=${1:-2}; OUT=${2:-gpurun_out/thp_probe.txt}
cd "$(dirname "$0")/.."
{
  echo "THP: $(cat /sys/kernel/mm/transparent_hugepage/enabled)  defrag: $(cat /sys/kernel/mm/transparent_hugepage/defrag)"
  echo "--- one process alone (gpu0), 4 KiB pages / THP"
  python tools/thp_probe.py 0 1.0 --small; python tools/thp_probe.py 0 1.0
  echo "--- $N processes at once, registered 4 KiB pages"
  for i in $(seq 0 $((N - 1))); do python tools/thp_probe.py $i 1.5 --small & done; wait
  echo "--- $N processes at once, transparent huge pages"
  for i in $(seq 0 $((N - 1))); do python tools/thp_probe.py $i 1.5 & done; wait
} > "$OUT" 2>&1
