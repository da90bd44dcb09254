#!/bin/bash
# One gpurun call: GPU parity tests, bench lines of every config, ncu launch list + full capture.
# usage (here): gpurun --timeout 1500 -- 'bash tools/gpu_round.sh <tag>'
set -u
TAG=${1:-r1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
tail -3 $OUT/pytest_gpu.log
python bench.py > $OUT/bench_config2.json 2> $OUT/bench_config2.err; echo "bench config2 rc=$?"
for cfg in config1 config3 config4 config5-100-16 config5-2000-256 config5-10-1 config5-300-64; do
  W=""
  case $cfg in config5-2000-256) W="--worlds 8192";; config5-*) W="--worlds 32768";; esac
  python bench.py --config $cfg $W --no-cpu-baseline --steps 10 --warmup 3 > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err; echo "bench $cfg rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f.split("bench_")[1][:-5], "ws/s=%.3g tests/s=%.3g ms=%.4f hbm=%.3f fp32=%.4f e2e=%.3g"%(d["value"],d["ray_segment_tests_per_sec"],d["ms_per_step"],r["hbm"]["frac"],r["fp32"]["frac"],d["e2e"]["value"]))
    except Exception as e:
        print(f, "ERR", e)
PY
# ncu: launch list, then one full capture of the step kernel (same command exited 0 just above)
CMD="python bench.py --steps 3 --warmup 3 --no-cpu-baseline"
$CMD > $OUT/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_config2.csv $CMD > $OUT/ncu1.log 2>&1
$CMD > $OUT/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:brs_step -s 4 -c 1 -o $OUT/prof_config2 $CMD > $OUT/ncu2.log 2>&1
CMD4="python bench.py --config config4 --steps 3 --warmup 3 --no-cpu-baseline"
$CMD4 > $OUT/plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:brs_step -s 4 -c 1 -o $OUT/prof_config4 $CMD4 > $OUT/ncu4.log 2>&1
ls -la $OUT
