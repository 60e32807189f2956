#!/bin/bash
# tools/sanitize.sh OUT_PREFIX : compute-sanitizer memcheck / racecheck / synccheck over smoke() (NTT forward + inverse,
# fused commitment, trapdoor verification), one fused-commitment test, one quotient test and one big-transform test
# ("new build: compute-sanitizer on K1-K4").  Writes <prefix>_{memcheck,racecheck,synccheck}.txt with
# each run's tail (the ERROR SUMMARY lines are what counts).
P=${1:-gpurun_out/sanitizer}
cd "$(dirname "$0")/.."
SMOKE='import __graft_entry__ as g; g.smoke()'
TESTS='tests/test_gpu_commit.py::test_device_pointer_commit tests/test_gpu_quotient.py::test_quotient_matches_restated_rust tests/test_gpu_ntt.py::test_seal_unit_test_constants_on_device tests/test_gpu_quotient.py::test_big_quotient_matches_c_oracle'
for tool in memcheck racecheck synccheck; do
  {
    echo "### compute-sanitizer --tool $tool : smoke()"
    timeout 900 compute-sanitizer --tool $tool --target-processes all python -c "$SMOKE" 2>&1 | grep -v "^$" | tail -n 12
    echo "### compute-sanitizer --tool $tool : pytest $TESTS"
    timeout 1500 compute-sanitizer --tool $tool --target-processes all python -m pytest -x -q $TESTS 2>&1 | grep -v "^$" | tail -n 14
  } > ${P}_$tool.txt 2>&1
done
