#!/bin/bash
# One gpurun call: (optional) GPU tests, then a matrix of bench lines, then (optional) ncu captures.
# usage: bash tools/gpu_matrix.sh <tag> <specfile>
# specfile lines:   bench <name> | ENV=.. ENV=.. | bench.py args
#                   ncu   <name> | ENV=..         | bench.py args | kernel-regex
#                   pytest [pytest args]
set -u
TAG=$1; SPEC=$2
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $OUT/smi.txt 2>&1
while IFS= read -r line; do
  [[ -z "$line" || "$line" == \#* ]] && continue
  kind=${line%% *}; rest=${line#* }
  [ "$rest" = "$line" ] && rest=""
  if [ "$kind" = "pytest" ]; then
    timeout 1500 python -m pytest tests -m gpu -x -q --timeout 900 $rest > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 $OUT/pytest_gpu.log
    continue
  fi
  IFS='|' read -r name envs args kre <<< "$rest"
  name=$(echo $name); envs=$(echo $envs); args=$(echo $args); kre=$(echo ${kre:-})
  if [ "$kind" = "bench" ]; then
    env $envs timeout 400 python bench.py $args --no-cpu-baseline --no-large-batch --steps 10 --warmup 3 > $OUT/b_$name.json 2> $OUT/b_$name.err
    python - <<PY
import json
try:
    d=json.loads(open("$OUT/b_$name.json").read().strip().splitlines()[-1]); r=d["roofline"]; l=d["config"].get("launch") or d.get("launch")
    print("%-30s ms=%.4f hbm=%.3f fp32=%.4f | %s thr=%d ctas=%d smem=%d tw=%d"%("$name",d["ms_per_step"],r["hbm"]["frac"],r["fp32"]["frac"],l["path"][:5],l["threads"],l["ctas"],l["smem_bytes"],l["worlds_per_tile"]))
except Exception as e:
    print("$name ERR", e, open("$OUT/b_$name.err").read()[-500:])
PY
  elif [ "$kind" = "ncu" ]; then
    CMD="python bench.py $args --steps 3 --warmup 3 --no-cpu-baseline --no-large-batch"
    env $envs $CMD > $OUT/plain_$name.log 2>&1 && \
    env $envs ncu --set full --clock-control none --import-source on -k regex:$kre -s 4 -c 1 -o $OUT/prof_$name $CMD > $OUT/ncu_$name.log 2>&1
    echo "ncu $name rc=$?"
  elif [ "$kind" = "launches" ]; then
    CMD="python bench.py $args --steps 3 --warmup 3 --no-cpu-baseline --no-large-batch"
    env $envs ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/launches_$name.csv $CMD > $OUT/ncul_$name.log 2>&1
    echo "launches $name rc=$?"
  fi
done < $SPEC
# the full default bench line (all sub-objects), as the driver runs it
if [ -n "${FULL_BENCH:-}" ]; then
  timeout 900 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "default bench rc=$?"; tail -c 600 $OUT/bench_default.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/bench_default.json").read().strip().splitlines()[-1])
    print("headline ms=%.4f frac=%.3f e2e=%.4g ceiling=%.1f GB/s launches=%d"%(d["ms_per_step"],d["roofline"]["frac"],d["e2e"]["value"],d["e2e"]["host_ceiling_gbs"],d["gpu_launches"]))
    for k,v in d.get("configs",{}).items(): print(k, {kk:vv for kk,vv in v.items() if kk in ("ms_per_step","hbm_frac","fp32_frac","fp32_executed_frac","world_steps_per_sec","seconds_for_1000_steps","gpu_seconds_for_1000_steps")})
except Exception as e: print("default bench parse ERR", e)
PY
fi
