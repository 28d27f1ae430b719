#!/bin/bash
# A/B a kernel variant on the GPU box: parity subset + bench lines under different env settings + one ncu capture.
TAG=$1; shift
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python -m pytest tests -x -q -m gpu -k "dxt or golden or stripes or full_size or medium" > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -5 $OUT/pytest_gpu.log
for cfg in "$@"; do
  name=$(echo "$cfg" | tr ' =' '__')
  env $cfg timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > $OUT/bench_$name.json 2> $OUT/bench_$name.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/bench_$name.json")); r=d["roofline"]
    print("$cfg: %.0f Mpix/s  kernel avg %.1f us min %.1f us frac %.3f" % (d["value"], r["kernel_ms_avg"]*1e3, r["kernel_ms_min"]*1e3, r["frac"]))
except Exception as e:
    print("bench parse failed", e); print(open("$OUT/bench_$name.err").read()[-2000:])
PY
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"encode4x4_tma" -s 4 -c 1 -o $OUT/prof \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu.log 2>&1; echo "ncu exit $?"
