#!/bin/bash
set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
nvidia-smi topo -m > gpurun_out/y_topo.txt 2>&1
for n in 8 4 2; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/y_bench_cfg5_${n}gpu.json 2> gpurun_out/y_bench_cfg5_${n}gpu.err
tail -1 gpurun_out/y_bench_cfg5_${n}gpu.json | cut -c1-400; tail -3 gpurun_out/y_bench_cfg5_${n}gpu.err | cut -c1-300
done
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --no-cpu --no-dense > gpurun_out/y_bench_cfg5_1gpu.json 2> gpurun_out/y_bench_cfg5_1gpu.err
tail -1 gpurun_out/y_bench_cfg5_1gpu.json | cut -c1-400
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 1 --warmup 1 --n-override 8000000 --no-e2e 2>&1 | grep -E "NVLS|Channel|Connected|nvls|Using network" | head -20 > gpurun_out/y_nccl_info.txt
