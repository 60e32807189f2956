# tools/collect_final.sh : end-of-round evidence on one B200 (run under gpurun): GPU tests, the bench line, the launch list of the
# bench command, and ncu --set full summaries of the fused commitment and verification kernels (each after a plain run of the
# same command that exited 0).  The .ncu-rep files are summarised on the box and deleted.
set -x
R=${1:-r02}
python -m pytest tests -x -q -m gpu > gpurun_out/${R}_pytest_gpu_final.log 2>&1; tail -3 gpurun_out/${R}_pytest_gpu_final.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_final.json 2> gpurun_out/${R}_bench_final.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 10 --warmup 3 --cpu-seconds 1 --no-parity > gpurun_out/ncu_bench.log 2>&1
python tools/launch_shares.py gpurun_out/${R}_launches_bench.csv "python bench.py --steps 10 --warmup 3 --cpu-seconds 1 --no-parity (first 900 launches)" > gpurun_out/${R}_launch_shares.txt
python tools/prof_driver.py commit 8192 && ncu --set full --clock-control none --import-source on -k regex:fused_commit -s 1 -c 1 -o gpurun_out/${R}_commit -f python tools/prof_driver.py commit 8192 > gpurun_out/ncu_1.log 2>&1
(python tools/ncu_summary.py gpurun_out/${R}_commit.ncu-rep --top 24; python tools/ncu_by_line.py gpurun_out/${R}_commit.ncu-rep --top 45) 2>&1 | cut -c1-200 > gpurun_out/${R}_ncu_commit.txt
python tools/prof_driver.py verify 8192 && ncu --set full --clock-control none --import-source on -k regex:fused_verify -s 1 -c 1 -o gpurun_out/${R}_verify -f python tools/prof_driver.py verify 8192 > gpurun_out/ncu_2.log 2>&1
(python tools/ncu_summary.py gpurun_out/${R}_verify.ncu-rep --top 24; python tools/ncu_by_line.py gpurun_out/${R}_verify.ncu-rep --top 30) 2>&1 | cut -c1-200 > gpurun_out/${R}_ncu_verify.txt
rm -f gpurun_out/*.ncu-rep
python tools/sweep.py > gpurun_out/${R}_ntt_sweep.json 2> gpurun_out/sweep.err
python tools/latency.py > gpurun_out/${R}_latency.txt 2>&1
tail -c 600 gpurun_out/${R}_bench_final.err; head -c 400 gpurun_out/${R}_bench_final.json
