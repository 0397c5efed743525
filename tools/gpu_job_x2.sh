#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/x_bench_cfg5_2gpu.json 2> gpurun_out/x_bench_cfg5_2gpu.err
cat gpurun_out/x_bench_cfg5_2gpu.json; tail -8 gpurun_out/x_bench_cfg5_2gpu.err
