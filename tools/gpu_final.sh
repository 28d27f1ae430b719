#!/bin/bash
# End-of-round evidence run (one GPU): full GPU suite, the default bench line, the reference arm, the ncu launch list of
# the bench command and sanitizer passes.  Usage (under gpurun):  bash tools/gpu_final.sh <tag>
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_gpu.log
timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench exit $?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference_arm.json 2> $OUT/bench_reference_arm.err; echo "reference arm exit $?"
for wl in dxt5_rgba8 dxt1_rgb8 etc1_rgb8 pvrtc2_rgba8; do
  timeout 300 python bench.py --workload $wl --no-others > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; echo "bench $wl exit $?"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_dxt1_rgba8.csv \
    python bench.py --steps 5 --warmup 3 --no-others --no-cpu-baseline > $OUT/ncu_launches.log 2>&1; echo "ncu launch list exit $?"
python - $OUT <<'PY'
import json, sys, os
out = sys.argv[1]
for f in sorted(os.listdir(out)):
    if f.startswith("bench_") and f.endswith(".json"):
        try:
            d = json.load(open(os.path.join(out, f)))
            print("%-28s value %.0f %s  step %.2f us  frac %s  parity %s  e2e %.0f" % (f, d["value"], d["unit"], d["ms_per_step"] * 1e3,
                  ("%.3f" % d["roofline"]["frac"]) if "roofline" in d else "-", (d.get("parity") or {}).get("equal"), d["e2e"]["value"]))
        except Exception as e:
            print(f, "parse failed", e)
PY
bash tools/gpu_sanitize.sh $TAG
