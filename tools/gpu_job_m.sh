#!/bin/bash
# screened E-step: parity, bench
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s -k "screened" > gpurun_out/m_tc.log 2>&1; echo "rc=$?" >> gpurun_out/m_tc.log
grep -E "screened sweep|passed|failed|rc=|Error|error|assert" gpurun_out/m_tc.log | tail -30
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_api.py -m gpu -x -q > gpurun_out/m_tc2.log 2>&1; echo "rc=$?" >> gpurun_out/m_tc2.log; tail -3 gpurun_out/m_tc2.log
timeout 600 python bench.py --workload cfg5 --n-override 4000000 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/m_bench_cfg5_4M.json 2> gpurun_out/m_bench_cfg5_4M.err
cat gpurun_out/m_bench_cfg5_4M.json; tail -5 gpurun_out/m_bench_cfg5_4M.err
