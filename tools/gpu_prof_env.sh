#!/bin/bash
# ncu full capture with env overrides. usage: bash tools/gpu_prof_env.sh <tag> <name> <cfg[:worlds]> "ENV=.. ENV=.."
set -u
TAG=$1; NAME=$2; SPEC=$3; ENVS=${4:-}
cfg=${SPEC%%:*}; W=""
if [[ "$SPEC" == *:* ]]; then W="--worlds ${SPEC##*:}"; fi
OUT=gpurun_out/$TAG; mkdir -p $OUT
CMD="env $ENVS python bench.py --config $cfg $W --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain_$NAME.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:brs_step -s 4 -c 1 -o $OUT/prof_$NAME $CMD > $OUT/ncu_$NAME.log 2>&1
echo "$NAME rc=$?"
