#!/bin/bash
# whole GPU suite + smoke + full cfg5 bench with the CTA-pair kernels
set -x
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/j_pytest.log 2>&1 ) 2> gpurun_out/j_pytest_time.txt; echo "rc=$?" >> gpurun_out/j_pytest.log
tail -5 gpurun_out/j_pytest.log; cat gpurun_out/j_pytest_time.txt
timeout 600 python __graft_entry__.py smoke > gpurun_out/j_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/j_smoke.log; tail -5 gpurun_out/j_smoke.log
timeout 1500 python bench.py --steps 3 --warmup 3 > gpurun_out/j_bench_cfg5.json 2> gpurun_out/j_bench_cfg5.err
cat gpurun_out/j_bench_cfg5.json; tail -5 gpurun_out/j_bench_cfg5.err
