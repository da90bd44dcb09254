#!/bin/bash
# quick GPU check: parity tests + bench lines (no ncu). usage: bash tools/gpu_quick.sh <tag> [configs...]
set -u
TAG=${1:-q}; shift
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 $OUT/pytest_gpu.log
CFGS=${@:-config2 config3 config4 config5-100-16 config5-2000-256 config5-10-1 config5-300-64}
for cfg in $CFGS; do
  W=""
  case $cfg in config5-2000-256) W="--worlds 8192";; config5-*) W="--worlds 32768";; esac
  timeout 300 python bench.py --config $cfg $W --no-cpu-baseline --steps 10 --warmup 3 > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err; echo "bench $cfg rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f.split("bench_")[1][:-5], "ws/s=%.3g tests/s=%.3g ms=%.4f hbm=%.3f fp32=%.4f e2e=%.3g"%(d["value"],d["ray_segment_tests_per_sec"],d["ms_per_step"],r["hbm"]["frac"],r["fp32"]["frac"],d["e2e"]["value"]), d["config"]["launch"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace('.json','.err')).read()[-400:])
PY
