#!/bin/bash
# tools/multi_gpu_probe.sh N OUT : one process per GPU at the same time, raw page-locked copies (tools/pcie_probe.py) and
# lwe_commit_batch end to end (tools/e2e_probe.py); writes OUT.  The host's DMA ceiling for N GPUs, next to what the
# library's pipeline reaches on the same box.
N=${1:-2}; OUT=${2:-gpurun_out/multi_gpu_probe.txt}
cd "$(dirname "$0")/.."
{
  echo "host: $(nproc) cpus, $(grep -m1 'model name' /proc/cpuinfo | cut -d: -f2), $(ls -d /sys/devices/system/node/node* | wc -l) NUMA node(s), $(free -g | awk '/Mem:/{print $2}') GB"
  nvidia-smi topo -m 2>/dev/null | head -n $((N + 3))
  echo "--- raw copies, one process alone (gpu0)"
  python tools/pcie_probe.py 0 1.0
  echo "--- raw copies, $N processes at once"
  for i in $(seq 0 $((N - 1))); do python tools/pcie_probe.py $i 1.5 --cpu-threads $(( $(nproc) / N > 4 ? 4 : $(nproc) / N )) & done; wait
  echo "--- lwe_commit_batch end to end, one process alone (gpu0)"
  python tools/e2e_probe.py 0 8192 2
  echo "--- lwe_commit_batch end to end, $N processes at once"
  for i in $(seq 0 $((N - 1))); do python tools/e2e_probe.py $i 8192 3 & done; wait
} > "$OUT" 2>&1
