#!/bin/bash
# CTA-pair statistics kernel: parity, bench, ncu
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s > gpurun_out/i_tc.log 2>&1; echo "rc=$?" >> gpurun_out/i_tc.log
grep -E "stats K|passed|failed|rc=|Error|error|assert" gpurun_out/i_tc.log | tail -30
timeout 600 python bench.py --workload cfg5 --n-override 4000000 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/i_bench_cfg5_4M.json 2> gpurun_out/i_bench_cfg5_4M.err
cat gpurun_out/i_bench_cfg5_4M.json; tail -5 gpurun_out/i_bench_cfg5_4M.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_fstats_kernel -s 1 -c 1 -o gpurun_out/i_prof_fstats -f python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/i_ncu_fstats.log 2>&1
