#!/bin/bash
# compare library variants (build/var/libbrs_<name>.so). usage: bash tools/gpu_variants.sh <tag> "<variants>" "<cfg[:worlds]|ENV..>" ...
set -u
TAG=$1; VARS=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
for spec in "$@"; do
  cfgw=${spec%%|*}; envs=""; [[ "$spec" == *"|"* ]] && envs=${spec#*|}
  cfg=${cfgw%%:*}; W=""; [[ "$cfgw" == *:* ]] && W="--worlds ${cfgw##*:}"
  for v in $VARS; do
    f=$OUT/v_${cfg}_${v}_$(echo "$envs" | tr ' =' '__').json
    env BRS_LIB=build/var/libbrs_$v.so $envs timeout 300 python bench.py --config $cfg $W --no-cpu-baseline --steps 10 --warmup 3 > $f 2> ${f%.json}.err
    python - <<PY
import json
try:
    d=json.loads(open("$f").read().strip().splitlines()[-1]); r=d["roofline"]
    print("%-10s %-8s [%s] ms=%.4f hbm=%.3f fp32=%.4f"%("$cfg","$v","$envs",d["ms_per_step"],r["hbm"]["frac"],r["fp32"]["frac"]), d["config"]["launch"])
except Exception as e:
    print("$cfg $v [$envs] ERR", e, open("${f%.json}.err").read()[-300:])
PY
  done
done
