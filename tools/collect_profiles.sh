set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01e.json 2> gpurun_out/bench_r01e.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r01e.csv python bench.py --steps 10 --warmup 3 --cpu-seconds 1 > gpurun_out/ncu_bench_e.log 2>&1
python tools/sweep.py > gpurun_out/sweep_e.json 2> gpurun_out/sweep_e.err
ncu --set full --clock-control none --import-source on -k regex:fused_commit -s 1 -c 1 -o gpurun_out/r01e_commit -f python tools/prof_driver.py commit 8192 > gpurun_out/ncu_e3.log 2>&1
ncu --set full --clock-control none -k regex:pointwise -s 1 -c 1 -o gpurun_out/r01e_pointwise -f python tools/prof_driver.py ntt 8192 > gpurun_out/ncu_e4.log 2>&1
ls -la gpurun_out
