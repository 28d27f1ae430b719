#!/bin/bash
# Short GPU visit while iterating on one kernel.  Usage: bash tools/gpu_quick.sh <tag> <workload> [pytest -k expr]
TAG=$1; WL=${2:-dxt1_rgba8}; KEXPR=${3:-"dxt or golden or stripes or full_size or medium"}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu -k "$KEXPR" > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -8 $OUT/pytest_gpu.log
for wl in $WL; do
  timeout 300 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "bench $wl exit $?"
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$wl.json")); r=d["roofline"]
    print("$wl: %.0f Mpix/s  kernel %.1f us  frac %.3f  e2e %.0f" % (d["value"], r["kernel_ms_avg"]*1e3, r["frac"], d["e2e"]["value"]))
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/bench_$wl.err").read()[-2000:])
PY
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:"encode4x4_tma|pvrtc" -s 4 -c 1 -o $OUT/prof_$wl \
      python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu_$wl.log 2>&1; echo "ncu exit $?"
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file $OUT/launches_$wl.csv \
      python bench.py --workload $wl --steps 4 --warmup 3 --no-cpu-baseline > $OUT/ncu_launches_$wl.log 2>&1
  python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("$OUT/launches_$wl.csv")) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows: agg[r[4][:70]].append(float(r[-1]))
for k, v in agg.items(): print("   %-70s n=%2d avg %.1f us" % (k, len(v), sum(v)/len(v)/ (1000.0 if max(v) > 5000 else 1.0)))
PY
done
if [ -x tools/microbench/pipe_rates ] && [ ! -f gpurun_out/pipe_rates.txt ]; then tools/microbench/pipe_rates > gpurun_out/pipe_rates.txt 2>&1; cat gpurun_out/pipe_rates.txt; fi
