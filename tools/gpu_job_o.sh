#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s -k "screened or sweep or loglik" > gpurun_out/o_tc.log 2>&1; echo "rc=$?" >> gpurun_out/o_tc.log
grep -E "screened sweep|passed|failed|rc=|Error|error|assert" gpurun_out/o_tc.log | tail -14
timeout 600 python bench.py --workload cfg5 --n-override 4000000 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/o_bench_cfg5_4M.json 2> gpurun_out/o_bench_cfg5_4M.err
cat gpurun_out/o_bench_cfg5_4M.json; tail -5 gpurun_out/o_bench_cfg5_4M.err
KREG='regex:tc_|softmax_kernel|screen_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 200 --csv --log-file gpurun_out/o_launches_1M.csv python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/o_ncu_list.log 2>&1
