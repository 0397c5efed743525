#!/bin/bash
# feature-form statistics v2 (static schedule, FP16 Dekker products): parity, bench, ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s -k "stats or sweep" > gpurun_out/e_tc.log 2>&1; echo "rc=$?" >> gpurun_out/e_tc.log
grep -E "stats K|passed|failed|rc=|Error|error" gpurun_out/e_tc.log | tail -30
timeout 600 python bench.py --workload cfg5 --n-override 4000000 --steps 2 --warmup 1 --no-cpu > gpurun_out/e_bench_cfg5_4M.json 2> gpurun_out/e_bench_cfg5_4M.err
cat gpurun_out/e_bench_cfg5_4M.json; tail -5 gpurun_out/e_bench_cfg5_4M.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_fstats_kernel -s 1 -c 1 -o gpurun_out/e_prof_fstats -f python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/e_ncu_fstats.log 2>&1
