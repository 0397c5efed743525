#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/u_bench_cfg5_full.json 2> gpurun_out/u_bench_cfg5_full.err
cat gpurun_out/u_bench_cfg5_full.json; tail -5 gpurun_out/u_bench_cfg5_full.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/u_bench_reference.json 2> gpurun_out/u_bench_reference.err
cat gpurun_out/u_bench_reference.json
KREG='regex:tc_|softmax_kernel|screen_|pair_stats|nw_phase|gating|stats_hard|label_'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 400 --csv --log-file gpurun_out/u_launches_4M.csv python bench.py --workload cfg5 --n-override 4000000 --steps 2 --warmup 1 --no-cpu --no-e2e --no-dense > gpurun_out/u_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:pair_stats_kernel -s 2 -c 1 -o gpurun_out/u_prof_pairstats -f python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-dense > gpurun_out/u_ncu_pair.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:screen_refine_kernel -s 2 -c 1 -o gpurun_out/u_prof_refine -f python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-dense > gpurun_out/u_ncu_refine.log 2>&1
