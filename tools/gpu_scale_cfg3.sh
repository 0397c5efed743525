#!/bin/bash
# cfg3 (diagonal GMM, Gibbs, N = 100M) on N GPUs of one box: tools/gpu_scale_cfg3.sh <tag> <N>
TAG=$1; N=$2
mkdir -p gpurun_out
if [ "$N" = "1" ]; then TR="python"; else TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"; fi
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 --workload cfg3 --no-cpu > gpurun_out/${TAG}_cfg3_${N}gpu.json 2> gpurun_out/${TAG}_cfg3_${N}gpu.err
tail -2 gpurun_out/${TAG}_cfg3_${N}gpu.err | cut -c1-300
python tools/bench_brief.py gpurun_out/${TAG}_cfg3_${N}gpu.json
