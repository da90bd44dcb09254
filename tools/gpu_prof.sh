#!/bin/bash
# ncu full capture of the step kernel for the given configs. usage: bash tools/gpu_prof.sh <tag> cfg[:worlds] ...
set -u
TAG=${1:-p}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
for spec in "$@"; do
  cfg=${spec%%:*}; W=""
  if [[ "$spec" == *:* ]]; then W="--worlds ${spec##*:}"; fi
  CMD="python bench.py --config $cfg $W --steps 3 --warmup 3 --no-cpu-baseline"
  $CMD > $OUT/plain_$cfg.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:brs_step -s 4 -c 1 -o $OUT/prof_$cfg $CMD > $OUT/ncu_$cfg.log 2>&1
  echo "$cfg rc=$?"
done
ls -la $OUT
