#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/d2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/d2_tests.log
tail -3 gpurun_out/d2_tests.log
timeout 600 python bench.py --workload cfg5 --n-override 4000000 --steps 3 --warmup 2 --no-cpu --no-dense > gpurun_out/d2_bench_cfg5_4M.json 2> gpurun_out/d2_bench_cfg5_4M.err
cat gpurun_out/d2_bench_cfg5_4M.json | cut -c1-1400; tail -3 gpurun_out/d2_bench_cfg5_4M.err
KREG='regex:tc_|softmax_kernel|screen_|pair_stats'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 80 --csv --log-file gpurun_out/d2_launches_1M.csv python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-dense > gpurun_out/d2_ncu_list.log 2>&1
