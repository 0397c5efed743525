#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/p_smi.txt
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/p_tests.log 2>&1; echo "rc=$?" >> gpurun_out/p_tests.log
grep -E "screened sweep|passed|failed|rc=|Error|error|assert" gpurun_out/p_tests.log | tail -20
timeout 600 python bench.py --workload cfg5 --n-override 4000000 --steps 2 --warmup 2 --no-cpu --no-e2e > gpurun_out/p_bench_cfg5_4M_screen.json 2> gpurun_out/p_bench_cfg5_4M_screen.err
cat gpurun_out/p_bench_cfg5_4M_screen.json; tail -5 gpurun_out/p_bench_cfg5_4M_screen.err
timeout 600 python bench.py --workload cfg5 --n-override 4000000 --steps 2 --warmup 2 --no-cpu --no-e2e --tc-mode 3 > gpurun_out/p_bench_cfg5_4M_dense.json 2> gpurun_out/p_bench_cfg5_4M_dense.err
cat gpurun_out/p_bench_cfg5_4M_dense.json; tail -5 gpurun_out/p_bench_cfg5_4M_dense.err
KREG='regex:tc_|softmax_kernel|screen_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 200 --csv --log-file gpurun_out/p_launches_1M.csv python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/p_ncu_list.log 2>&1
tail -30 gpurun_out/p_launches_1M.csv | cut -c1-300
