#!/bin/bash
# generic GPU job: tools/gpu_job.sh <tag> <pytest args or ""> [bench-variant "ENV=.. args"]...
# everything lands in gpurun_out/<tag>_*
TAG=$1; shift
TESTS=$1; shift
mkdir -p gpurun_out
if [ -n "$TESTS" ]; then
  timeout 900 python -m pytest $TESTS -m gpu -q > gpurun_out/${TAG}_tests.log 2>&1
  tail -8 gpurun_out/${TAG}_tests.log
fi
i=0
for V in "$@"; do
  i=$((i+1))
  echo "== bench $i: $V"
  timeout 1500 env $V > gpurun_out/${TAG}_bench$i.json 2> gpurun_out/${TAG}_bench$i.err
  tail -3 gpurun_out/${TAG}_bench$i.err
  python tools/bench_brief.py gpurun_out/${TAG}_bench$i.json
done
