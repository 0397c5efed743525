#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/z_tests.log 2>&1; echo "rc=$?" >> gpurun_out/z_tests.log
tail -4 gpurun_out/z_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/z_smoke.log 2>&1; tail -5 gpurun_out/z_smoke.log
timeout 600 python bench.py --workload cfg5 --n-override 4000000 --steps 3 --warmup 2 --no-cpu --no-dense > gpurun_out/z_bench_cfg5_4M.json 2> gpurun_out/z_bench_cfg5_4M.err
cat gpurun_out/z_bench_cfg5_4M.json; tail -5 gpurun_out/z_bench_cfg5_4M.err
for c in cfg4 cfg2; do
timeout 600 python bench.py --workload $c --steps 5 --warmup 3 > gpurun_out/z_bench_$c.json 2> gpurun_out/z_bench_$c.err
cat gpurun_out/z_bench_$c.json | cut -c1-1500; tail -3 gpurun_out/z_bench_$c.err
timeout 600 python bench.py --workload $c --steps 5 --warmup 3 --tc-mode 0 --no-cpu --no-e2e > gpurun_out/z_bench_${c}_dense.json 2> gpurun_out/z_bench_${c}_dense.err
cat gpurun_out/z_bench_${c}_dense.json | cut -c1-900
done
KREG='regex:tc_|softmax_kernel|screen_|pair_stats'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREG" -c 80 --csv --log-file gpurun_out/z_launches_1M.csv python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e --no-dense > gpurun_out/z_ncu_list.log 2>&1
