#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/s_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s_tests.log
grep -E "screened sweep|passed|failed|rc=|Error|error|assert" gpurun_out/s_tests.log | tail -20
timeout 600 python bench.py --workload cfg5 --n-override 4000000 --steps 2 --warmup 2 --no-cpu > gpurun_out/s_bench_cfg5_4M.json 2> gpurun_out/s_bench_cfg5_4M.err
cat gpurun_out/s_bench_cfg5_4M.json; tail -5 gpurun_out/s_bench_cfg5_4M.err
KREG='regex:tc_|softmax_kernel|screen_|pais_stats'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 80 --csv --log-file gpurun_out/s_launches_1M.csv python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/s_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_estep2_kernel -s 2 -c 1 -o gpurun_out/s_prof_estep_screen -f python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/s_ncu_estep.log 2>&1
