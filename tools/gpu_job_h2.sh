#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 --n-override 8000000 --no-e2e > gpurun_out/h2_shard.json 2> gpurun_out/h2_shard.err
tail -1 gpurun_out/h2_shard.json | cut -c1-300; tail -3 gpurun_out/h2_shard.err | cut -c1-300
MIMO_REPLICATED_POSTERIOR=1 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 2 --steps 3 --warmup 3 --n-override 8000000 --no-e2e > gpurun_out/h2_repl.json 2> gpurun_out/h2_repl.err
tail -1 gpurun_out/h2_repl.json | cut -c1-300; tail -3 gpurun_out/h2_repl.err | cut -c1-300
timeout 600 python bench.py --gpus 1 --steps 3 --warmup 3 --n-override 8000000 --no-e2e --no-cpu --no-dense > gpurun_out/h2_one.json 2> gpurun_out/h2_one.err
tail -1 gpurun_out/h2_one.json | cut -c1-300
