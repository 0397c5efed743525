#!/bin/bash
# ncu captures of the two tensor-core kernels at cfg5 shape (1M points = one chunk), plus the rest of the tc parity tests
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q > gpurun_out/d_tc.log 2>&1; echo "rc=$?" >> gpurun_out/d_tc.log
tail -5 gpurun_out/d_tc.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_fstats_kernel -s 1 -c 1 -o gpurun_out/d_prof_fstats -f python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/d_ncu_fstats.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_estep_kernel -s 1 -c 1 -o gpurun_out/d_prof_estep -f python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/d_ncu_estep.log 2>&1
ls -la gpurun_out/
