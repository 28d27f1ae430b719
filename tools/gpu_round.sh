#!/bin/bash
# One GPU-box visit while iterating: full GPU test suite, then every workload's bench line (with and without
# programmatic dependent launch for the 4x4 codecs), then one ncu --set full capture of the DXT1 kernel.
# Usage (under gpurun, from the repo root):  bash tools/gpu_round.sh <tag>
TAG=${1:-round}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
show() { python - "$1" "$2" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d["roofline"]
    print("%-28s %9.0f Mpix/s  step %.1f us  isolated avg %.1f min %.1f  frac %.3f  e2e %.0f" % (
        sys.argv[2], d["value"], r["kernel_ms_avg"] * 1e3, r["kernel_ms_isolated_avg"] * 1e3, r["kernel_ms_min"] * 1e3, r["frac"], d["e2e"]["value"]))
except Exception as e:
    print(sys.argv[2], "bench parse failed", e)
PY
}
for wl in dxt1_rgba8 dxt1_rgb8 dxt5_rgba8 etc1_rgb8 pvrtc2_rgba8; do
  timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench_$wl.json 2> $OUT/bench_$wl.err; show $OUT/bench_$wl.json $wl
done
for wl in dxt1_rgba8 dxt5_rgba8; do
  ICB_NO_PDL=1 timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline > $OUT/bench_${wl}_nopdl.json 2> $OUT/bench_${wl}_nopdl.err; show $OUT/bench_${wl}_nopdl.json "$wl (no PDL)"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"encode4x4_(tma|ring)" -s 4 -c 1 -o $OUT/prof_dxt1_rgba8 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/ncu.log 2>&1; echo "ncu exit $?"
