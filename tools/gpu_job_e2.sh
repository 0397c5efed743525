#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/e2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/e2_tests.log
tail -4 gpurun_out/e2_tests.log
timeout 1500 python bench.py > gpurun_out/e2_bench_cfg5_full.json 2> gpurun_out/e2_bench_cfg5_full.err
cat gpurun_out/e2_bench_cfg5_full.json; tail -5 gpurun_out/e2_bench_cfg5_full.err
KREG='regex:tc_|softmax_kernel|screen_|pair_stats|nw_phase|gating|label_'
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 2000 --csv --log-file gpurun_out/e2_launches_full.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-dense > gpurun_out/e2_ncu_list.log 2>&1
for kk in screen_refine_kernel:2 pair_stats_kernel:3 screen_emit_kernel:1; do
k=${kk%%:*}; sk=${kk##*:}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s $sk -c 1 -o gpurun_out/e2_prof_$k -f python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-dense > gpurun_out/e2_ncu_$k.log 2>&1
done
ls -la gpurun_out | tail -12
