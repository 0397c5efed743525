#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s -k "tiers or large" > gpurun_out/f2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/f2_tests.log
grep -E "tiers across|passed|failed|rc=|Error|error|assert" gpurun_out/f2_tests.log | tail
