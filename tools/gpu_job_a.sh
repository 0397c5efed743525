#!/bin/bash
# first GPU job of the round: parity tests, a bench line per workload family, ncu launch list + one full capture
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/a_gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py --workload cfg4 --steps 5 --warmup 3 > gpurun_out/a_bench_cfg4.json 2> gpurun_out/a_bench_cfg4.err
timeout 600 python bench.py --workload cfg5 --n-override 2000000 --steps 2 --warmup 1 --no-cpu > gpurun_out/a_bench_cfg5_2M.json 2> gpurun_out/a_bench_cfg5_2M.err
timeout 600 python bench.py --workload cfg3 --n-override 20000000 --steps 2 --warmup 1 --no-cpu > gpurun_out/a_bench_cfg3_20M.json 2> gpurun_out/a_bench_cfg3_20M.err
timeout 600 python bench.py --workload cfg2 --steps 3 --warmup 3 --no-cpu > gpurun_out/a_bench_cfg2.json 2> gpurun_out/a_bench_cfg2.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/a_launches_cfg4.csv python bench.py --workload cfg4 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/a_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:quad_loglik -s 2 -c 2 -o gpurun_out/a_prof_quad -f python bench.py --workload cfg4 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/a_ncu_full.log 2>&1
tail -3 gpurun_out/a_pytest.log; cat gpurun_out/a_bench_cfg4.json gpurun_out/a_bench_cfg5_2M.json
