#!/bin/bash
# env-var sweep of launch geometry. usage: bash tools/gpu_sweep.sh <tag> <config[:worlds]> "ENV1=a ENV2=b" "ENV1=c" ...
set -u
TAG=$1; SPEC=$2; shift 2
cfg=${SPEC%%:*}; W=""
if [[ "$SPEC" == *:* ]]; then W="--worlds ${SPEC##*:}"; fi
OUT=gpurun_out/$TAG; mkdir -p $OUT
for envs in "$@"; do
  name=$(echo "$envs" | tr ' =' '__')
  env $envs timeout 300 python bench.py --config $cfg $W --no-cpu-baseline --steps 10 --warmup 3 > $OUT/sw_${cfg}_$name.json 2> $OUT/sw_${cfg}_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/sw_${cfg}_$name.json").read().strip().splitlines()[-1])
    r=d["roofline"]
    print("$cfg [$envs]", "ms=%.4f hbm=%.3f fp32=%.4f"%(d["ms_per_step"],r["hbm"]["frac"],r["fp32"]["frac"]), d["config"]["launch"])
except Exception as e:
    print("$cfg [$envs] ERR", e, open("$OUT/sw_${cfg}_$name.err").read()[-300:])
PY
done
