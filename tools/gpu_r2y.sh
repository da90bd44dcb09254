#!/bin/bash
# round 2, call y: brs_steps evidence -- tests, default bench, launch lists, full captures of the one-launch kernels
set -u
OUT=gpurun_out/r2y; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench_default.json 2> $OUT/bench_default.err; echo "bench rc=$?"
timeout 300 python bench.py --config config4 --no-cpu-baseline --steps 10 --warmup 3 > $OUT/bench_config4.json 2> $OUT/bench_config4.err; echo "bench c4 rc=$?"
timeout 300 python bench.py --config config3 --no-cpu-baseline --steps 10 --warmup 3 > $OUT/bench_config3.json 2> $OUT/bench_config3.err; echo "bench c3 rc=$?"
for cfg in config2 config3; do
  CMD="python tools/run_steps.py $cfg 10"
  $CMD > $OUT/plain_$cfg.log 2>&1; echo "plain $cfg rc=$?"; cat $OUT/plain_$cfg.log
  ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_steps_$cfg.csv $CMD > $OUT/ncul_$cfg.log 2>&1; echo "launches $cfg rc=$?"
  ncu --set full --clock-control none --import-source on -k regex:brs_ -s 4 -c 1 -o $OUT/prof_${cfg}_steps $CMD > $OUT/ncu_$cfg.log 2>&1; echo "ncu $cfg rc=$?"
done
CMD="python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-large-batch --no-extra"
$CMD > $OUT/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_bench_config2.csv $CMD > $OUT/ncul_bench.log 2>&1; echo "launches bench rc=$?"
ls -la $OUT
