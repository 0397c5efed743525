#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --workload cfg3 --steps 3 --warmup 3 > gpurun_out/g2_bench_cfg3.json 2> gpurun_out/g2_bench_cfg3.err
cat gpurun_out/g2_bench_cfg3.json | cut -c1-1800; tail -3 gpurun_out/g2_bench_cfg3.err
timeout 300 python bench.py --workload cfg2 --steps 5 --warmup 3 > gpurun_out/g2_bench_cfg2.json 2> gpurun_out/g2_bench_cfg2.err
cat gpurun_out/g2_bench_cfg2.json | cut -c1-600
