#!/bin/bash
# round 2, call r3g: evidence for the final kernels -- tests, bench lines, launch lists, full captures
set -u
OUT=gpurun_out/${TAG:-r3g}; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 700 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"
for cfg in config3 config4 config1 config5-30-4 config5-100-16 config5-2000-256; do
  W=""; case $cfg in config5-2000-256) W="--worlds 16384";; config5-*) W="--worlds 32768";; esac
  timeout 300 python bench.py --config $cfg $W --no-cpu-baseline --steps 10 --warmup 3 > $OUT/bench_$cfg.json 2> $OUT/bench_$cfg.err; echo "bench $cfg rc=$?"
done
for spec in config2:10: config5-30-4:20:32768; do
  cfg=${spec%%:*}; rest=${spec#*:}; K=${rest%%:*}; W=${rest#*:}
  CMD="python tools/run_steps.py $cfg $K $W"
  $CMD > $OUT/plain_$cfg.log 2>&1; echo "plain $cfg rc=$?"; cat $OUT/plain_$cfg.log
  ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_steps_$cfg.csv $CMD > $OUT/ncul_$cfg.log 2>&1; echo "launches $cfg rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:brs_ -s 4 -c 1 -o $OUT/prof_${cfg}_steps $CMD > $OUT/ncu_$cfg.log 2>&1; echo "ncu $cfg rc=$?"
done
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-large-batch --no-extra"
$CMD > $OUT/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_bench_config2.csv $CMD > $OUT/ncul_bench.log 2>&1; echo "launches bench rc=$?"
du -sh $OUT
