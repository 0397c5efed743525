#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/b2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/b2_tests.log
tail -4 gpurun_out/b2_tests.log
timeout 900 python bench.py --workload cfg3 --n-override 20000000 --steps 3 --warmup 2 --no-cpu > gpurun_out/b2_bench_cfg3_20M.json 2> gpurun_out/b2_bench_cfg3_20M.err
cat gpurun_out/b2_bench_cfg3_20M.json | cut -c1-1400; tail -3 gpurun_out/b2_bench_cfg3_20M.err
for c in cfg4 cfg2; do
timeout 600 python bench.py --workload $c --steps 5 --warmup 3 > gpurun_out/b2_bench_$c.json 2> gpurun_out/b2_bench_$c.err
cat gpurun_out/b2_bench_$c.json | cut -c1-1200; tail -3 gpurun_out/b2_bench_$c.err
done
timeout 1500 python bench.py > gpurun_out/b2_bench_cfg5_full.json 2> gpurun_out/b2_bench_cfg5_full.err
cat gpurun_out/b2_bench_cfg5_full.json; tail -5 gpurun_out/b2_bench_cfg5_full.err
