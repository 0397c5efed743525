#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s > gpurun_out/b_tc.log 2>&1; echo "rc=$?" >> gpurun_out/b_tc.log
tail -40 gpurun_out/b_tc.log
