#!/bin/bash
# feature-form statistics kernel: parity, then bench lines on the tensor-core path
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s > gpurun_out/c_tc.log 2>&1; echo "rc=$?" >> gpurun_out/c_tc.log
tail -30 gpurun_out/c_tc.log
timeout 600 python bench.py --workload cfg5 --n-override 4000000 --steps 2 --warmup 1 --no-cpu > gpurun_out/c_bench_cfg5_4M.json 2> gpurun_out/c_bench_cfg5_4M.err
cat gpurun_out/c_bench_cfg5_4M.json; tail -5 gpurun_out/c_bench_cfg5_4M.err
