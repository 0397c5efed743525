#!/bin/bash
# CTA-pair (cta_group::2) E-step: parity, bench, ncu
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -s -k "loglik" > gpurun_out/h_tc.log 2>&1; echo "rc=$?" >> gpurun_out/h_tc.log
grep -E "loglik K|passed|failed|rc=|Error|error|assert" gpurun_out/h_tc.log | tail -30
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "sweep" > gpurun_out/h_tc2.log 2>&1; echo "rc=$?" >> gpurun_out/h_tc2.log; tail -3 gpurun_out/h_tc2.log
timeout 600 python bench.py --workload cfg5 --n-override 4000000 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/h_bench_cfg5_4M.json 2> gpurun_out/h_bench_cfg5_4M.err
cat gpurun_out/h_bench_cfg5_4M.json; tail -5 gpurun_out/h_bench_cfg5_4M.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_estep2_kernel -s 1 -c 1 -o gpurun_out/h_prof_estep2 -f python bench.py --workload cfg5 --n-override 1000000 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/h_ncu_estep2.log 2>&1
