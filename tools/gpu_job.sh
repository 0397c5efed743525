#!/bin/bash
# generic GPU job: tag + commands; everything lands in gpurun_out/<tag>_*
# usage: tools/gpu_job.sh <tag> [tests] [bench "<args>"] ...
TAG=$1; shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/${TAG}_gpu.txt 2>&1
