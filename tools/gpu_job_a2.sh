#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s > gpurun_out/a2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/a2_tests.log
grep -E "screened sweep|all-rows|passed|failed|rc=|Error|error|assert" gpurun_out/a2_tests.log | tail -24
MIMO_HOST_DEBUG=1 timeout 600 python bench.py --workload cfg5 --n-override 4000000 --steps 3 --warmup 2 --no-cpu --no-dense > gpurun_out/a2_bench_cfg5_4M.json 2> gpurun_out/a2_bench_cfg5_4M.err
cat gpurun_out/a2_bench_cfg5_4M.json | cut -c1-200; grep mimo_sweep_host gpurun_out/a2_bench_cfg5_4M.err | tail -12
