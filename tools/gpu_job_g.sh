#!/bin/bash
# paced fstats + dense E-step: parity, FULL cfg5 bench (N=50M), launch list of the same command, ncu captures
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/g_tc.log 2>&1; echo "rc=$?" >> gpurun_out/g_tc.log
tail -3 gpurun_out/g_tc.log
( time timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/g_bench_cfg5.json 2> gpurun_out/g_bench_cfg5.err ) 2> gpurun_out/g_bench_time.txt
cat gpurun_out/g_bench_cfg5.json; tail -5 gpurun_out/g_bench_cfg5.err; cat gpurun_out/g_bench_time.txt
timeout 300 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/g_bench_ref.json 2> gpurun_out/g_bench_ref.err; cat gpurun_out/g_bench_ref.json
KREG='regex:tc_|softmax_kernel|nw_phase|gating_kernel|mean_over_k|stats_|label_|quad_loglik|diag_loglik'
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 400 --csv --log-file gpurun_out/g_launches_cfg5.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/g_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_fstats_kernel -s 1 -c 1 -o gpurun_out/g_prof_fstats -f python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/g_ncu_fstats.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_estep_kernel -s 1 -c 1 -o gpurun_out/g_prof_estep -f python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/g_ncu_estep.log 2>&1
ls -la gpurun_out | tail -12
