#!/bin/bash
set -x
mkdir -p gpurun_out
KREG='regex:tc_|softmax_kernel|screen_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 200 --csv --log-file gpurun_out/n_launches_1M.csv python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/n_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_estep2_kernel -s 2 -c 1 -o gpurun_out/n_prof_estep1 -f python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/n_ncu_estep1.log 2>&1
