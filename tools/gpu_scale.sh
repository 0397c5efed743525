#!/bin/bash
# multi-GPU bench legs on one box: tools/gpu_scale.sh <tag> <N> [extra bench args for the cfg5 leg]
TAG=$1; N=$2; shift 2
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"
timeout 900 $TR bench.py --gpus $N --steps 5 --warmup 3 --no-cpu --no-others --no-overlap "$@" > gpurun_out/${TAG}_cfg5_${N}gpu.json 2> gpurun_out/${TAG}_cfg5_${N}gpu.err
tail -2 gpurun_out/${TAG}_cfg5_${N}gpu.err | cut -c1-300
python tools/bench_brief.py gpurun_out/${TAG}_cfg5_${N}gpu.json
timeout 600 $TR bench.py --gpus $N --steps 5 --warmup 3 --workload cfg3 --no-cpu > gpurun_out/${TAG}_cfg3_${N}gpu.json 2> gpurun_out/${TAG}_cfg3_${N}gpu.err
tail -2 gpurun_out/${TAG}_cfg3_${N}gpu.err | cut -c1-300
python tools/bench_brief.py gpurun_out/${TAG}_cfg3_${N}gpu.json
