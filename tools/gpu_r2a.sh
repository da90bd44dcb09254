#!/bin/bash
# round 2, call a: first contact of the wide path -- parity tests, fused vs wide bench lines, variants
set -u
OUT=gpurun_out/r2a; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
# quick smoke of the wide path first (a hang here must not eat the budget)
BRS_WIDE=1 timeout 120 python - > $OUT/smoke_wide.log 2>&1 <<'PY'
import sys; sys.path.insert(0, "tests")
import __graft_entry__ as ge; ge.build()
import aitk_robots_b200 as rb, parity
for name, sc, st, cam in (("demo", rb.scenes.demo_scene(32), 3, True), ("c4", rb.scenes.config4(n_worlds=64), 3, True),
                          ("c3", rb.scenes.config3(n_worlds=32), 5, False)):
    print(name, parity.compare(sc, n_steps=st, worlds=range(0, sc.n_worlds, 4), camera=cam), flush=True)
PY
echo "smoke_wide rc=$?"; tail -5 $OUT/smoke_wide.log
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -8 $OUT/pytest_gpu.log
run() {  # name, env..., -- bench args
  local name=$1; shift
  local envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 300 python bench.py "$@" --no-cpu-baseline --no-large-batch --steps 10 --warmup 3 > $OUT/b_$name.json 2> $OUT/b_$name.err
  python - <<PY
import json
try:
    d=json.loads(open("$OUT/b_$name.json").read().strip().splitlines()[-1]); r=d["roofline"]
    print("%-28s ms=%.4f hbm=%.3f fp32=%.4f"%("$name",d["ms_per_step"],r["hbm"]["frac"],r["fp32"]["frac"]), d["config"]["launch"])
except Exception as e:
    print("$name ERR", e, open("$OUT/b_$name.err").read()[-400:])
PY
}
for cfg in config4 config3 config2 config5-100-16:32768 config5-2000-256:8192 config5-10-1:32768 config5-300-64:32768 config5-1000-4:32768 config1; do
  c=${cfg%%:*}; W=""; [[ "$cfg" == *:* ]] && W="--worlds ${cfg##*:}"
  run ${c}_fused BRS_WIDE=0 -- --config $c $W
  run ${c}_wide BRS_WIDE=1 -- --config $c $W
done
# variants of the wide path on config 4 / 3
for v in w24 nozones; do
  run config4_wide_$v BRS_WIDE=1 BRS_LIB=build/var/libbrs_$v.so -- --config config4
  run config2_fused_$v BRS_WIDE=0 BRS_LIB=build/var/libbrs_$v.so -- --config config2
done
run config4_wide_tw2 BRS_WIDE=1 BRS_WTW=2 BRS_WNT=64 -- --config config4
run config4_wide_nopdl BRS_WIDE=1 BRS_PDL=0 -- --config config4
run config3_wide_nt64 BRS_WIDE=1 BRS_WNT=64 BRS_WTW=2 -- --config config3
run config3_wide_tw1 BRS_WIDE=1 BRS_WTW=1 -- --config config3
# ncu: config 4 wide
CMD4="python bench.py --config config4 --steps 3 --warmup 3 --no-cpu-baseline --no-large-batch"
BRS_WIDE=1 $CMD4 > $OUT/plain4.log 2>&1 && \
BRS_WIDE=1 ncu --set full --clock-control none --import-source on -k regex:brs_wide -s 4 -c 1 -o $OUT/prof_config4_wide $CMD4 > $OUT/ncu4.log 2>&1
BRS_WIDE=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 30 --csv --log-file $OUT/launches_config4_wide.csv $CMD4 > $OUT/ncu4l.log 2>&1
ls -la $OUT | head -60
