# tools/collect_profiles.sh : the round's profile set on one B200 (run under gpurun).  Every ncu pass follows a plain run of
# the same command that exited 0; the .ncu-rep files are summarised on the box (tools/ncu_summary.py, tools/ncu_by_line.py)
# and deleted -- with imported sources they exceed what gpurun brings back.
set -x
R=${1:-r02}
python bench.py --steps 10 --warmup 3 > gpurun_out/${R}_bench_final.json 2> gpurun_out/${R}_bench_final.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${R}_launches_bench.csv python bench.py --steps 10 --warmup 3 --cpu-seconds 1 --no-parity > gpurun_out/ncu_bench.log 2>&1
python tools/launch_shares.py gpurun_out/${R}_launches_bench.csv "python bench.py --steps 10 --warmup 3 --cpu-seconds 1 --no-parity (first 900 launches)" > gpurun_out/${R}_launch_shares.txt
python tools/prof_driver.py commit 8192 && ncu --set full --clock-control none --import-source on -k regex:fused_commit -s 1 -c 1 -o gpurun_out/${R}_commit -f python tools/prof_driver.py commit 8192 > gpurun_out/ncu_1.log 2>&1
(python tools/ncu_summary.py gpurun_out/${R}_commit.ncu-rep --top 24; python tools/ncu_by_line.py gpurun_out/${R}_commit.ncu-rep --top 45) 2>&1 | cut -c1-200 > gpurun_out/${R}_ncu_commit.txt
python tools/prof_driver.py ntt 8192 && ncu --set full --clock-control none --import-source on -k regex:"ntt_tile|pointwise" -s 3 -c 3 -o gpurun_out/${R}_ntt -f python tools/prof_driver.py ntt 8192 > gpurun_out/ncu_2.log 2>&1
python tools/ncu_summary.py gpurun_out/${R}_ntt.ncu-rep --top 12 2>&1 | cut -c1-200 > gpurun_out/${R}_ncu_ntt.txt
python tools/prover_phase.py 20 4 && ncu --set full --clock-control none --import-source on -k regex:"spmv3|ntt_tile|ntt_column2" -s 14 -c 7 -o gpurun_out/${R}_quotient -f python tools/prover_phase.py 20 4 > gpurun_out/ncu_3.log 2>&1
python tools/ncu_summary.py gpurun_out/${R}_quotient.ncu-rep --top 10 2>&1 | cut -c1-200 > gpurun_out/${R}_ncu_quotient.txt
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out; du -sh gpurun_out
