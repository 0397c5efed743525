#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python bench.py > gpurun_out/v_bench_cfg5_full.json 2> gpurun_out/v_bench_cfg5_full.err
cat gpurun_out/v_bench_cfg5_full.json; tail -5 gpurun_out/v_bench_cfg5_full.err
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/v_tests.log 2>&1; echo "rc=$?" >> gpurun_out/v_tests.log
grep -E "passed|failed|rc=|Error|error|assert" gpurun_out/v_tests.log | tail -8
