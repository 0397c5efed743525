#!/bin/bash
# second GPU job: parity of the tcgen05 kernels, then bench lines on the tensor-core path
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s > gpurun_out/b_tc.log 2>&1; echo "rc=$?" >> gpurun_out/b_tc.log
tail -40 gpurun_out/b_tc.log
timeout 600 python bench.py --workload cfg5 --n-override 4000000 --steps 2 --warmup 1 --no-cpu > gpurun_out/b_bench_cfg5_4M.json 2> gpurun_out/b_bench_cfg5_4M.err
cat gpurun_out/b_bench_cfg5_4M.json; tail -5 gpurun_out/b_bench_cfg5_4M.err
