#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/i2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/i2_tests.log
tail -4 gpurun_out/i2_tests.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/i2_smoke.log 2>&1; tail -5 gpurun_out/i2_smoke.log
