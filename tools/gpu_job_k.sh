#!/bin/bash
# smoke + slab-balanced statistics schedule
set -x
mkdir -p gpurun_out
timeout 600 python __graft_entry__.py smoke > gpurun_out/k_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/k_smoke.log; tail -5 gpurun_out/k_smoke.log
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "stats or sweep" > gpurun_out/k_tc.log 2>&1; echo "rc=$?" >> gpurun_out/k_tc.log; tail -3 gpurun_out/k_tc.log
timeout 600 python bench.py --workload cfg5 --n-override 4000000 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/k_bench_cfg5_4M.json 2> gpurun_out/k_bench_cfg5_4M.err
cat gpurun_out/k_bench_cfg5_4M.json; tail -5 gpurun_out/k_bench_cfg5_4M.err
