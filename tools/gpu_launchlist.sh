#!/bin/bash
# ncu launch list (gpu__time_duration.sum) of this library's kernels inside a bench command
# usage: tools/gpu_launchlist.sh <tag> <skip> <count> <bench args...>
TAG=$1; SKIP=$2; CNT=$3; shift 3
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none \
    -k 'regex:^(tc|nw_|ng_|mnw_|gating|softmax|pair_|label_|stats_|feature_|td_|screen_|quad_|diag_|resp_|predict|studentt)' \
    -s $SKIP -c $CNT --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py "$@" > gpurun_out/${TAG}_ncu_bench.log 2>&1
tail -1 gpurun_out/${TAG}_launches.csv | cut -c1-200
