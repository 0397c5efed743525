#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/c2_tests.log
tail -4 gpurun_out/c2_tests.log
timeout 600 python bench.py --workload cfg2 --steps 5 --warmup 3 --no-cpu > gpurun_out/c2_bench_cfg2.json 2> gpurun_out/c2_bench_cfg2.err
cat gpurun_out/c2_bench_cfg2.json | cut -c1-1200; tail -3 gpurun_out/c2_bench_cfg2.err
timeout 600 python bench.py --workload cfg5 --n-override 4000000 --steps 3 --warmup 2 --no-cpu --no-dense > gpurun_out/c2_bench_cfg5_4M.json 2> gpurun_out/c2_bench_cfg5_4M.err
cat gpurun_out/c2_bench_cfg5_4M.json | cut -c1-1600; tail -3 gpurun_out/c2_bench_cfg5_4M.err
KREG='regex:tc_|softmax_kernel|screen_|pair_stats'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 80 --csv --log-file gpurun_out/c2_launches_1M.csv python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-dense > gpurun_out/c2_ncu_list.log 2>&1
