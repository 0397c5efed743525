#!/bin/bash
# ncu --set full capture of kernels matching a regex inside a short bench run
# usage: tools/gpu_ncu.sh <tag> <kernel regex> <skip> <count> <bench args...>
TAG=$1; RE=$2; SKIP=$3; CNT=$4; shift 4
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RE -s $SKIP -c $CNT -f -o gpurun_out/${TAG} \
    python bench.py "$@" > gpurun_out/${TAG}_ncu.log 2>&1
tail -3 gpurun_out/${TAG}_ncu.log | cut -c1-300
ls -la gpurun_out/${TAG}.ncu-rep
