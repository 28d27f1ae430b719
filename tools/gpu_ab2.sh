#!/bin/bash
TAG=$1; OUT=gpurun_out/$TAG; mkdir -p $OUT
run() {  # name, workload, env...
  name=$1; wl=$2; shift 2
  env "$@" timeout 200 python bench.py --workload $wl --steps 30 --warmup 5 --no-cpu-baseline > $OUT/${name}_$wl.json 2> $OUT/${name}_$wl.err
  python - <<PY
import json
try:
    d=json.load(open("$OUT/${name}_$wl.json")); r=d["roofline"]
    print("%-22s %-12s b2b %.2f us  iso-med %.2f  min %.2f  frac %.3f  e2e-ok %s clocks %s %s" % ("$name", "$wl", r["kernel_ms_avg"]*1e3, r["kernel_ms_isolated_median"]*1e3, r["kernel_ms_min"]*1e3, r["frac"], d["e2e"]["output_equals_device_path"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e:
    print("$name $wl failed", e); print(open("$OUT/${name}_$wl.err").read()[-1500:])
PY
}
V=$PWD/image_compression_b200/lib/variants
timeout 700 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
ICB200_LIB=$V/libicb200_mbar.so timeout 700 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu_mbar.log 2>&1; echo "pytest(mbar) exit $?"; tail -2 $OUT/pytest_gpu_mbar.log
for wl in dxt1_rgba8 dxt5_rgba8 dxt1_rgb8; do
  run early_acqrel $wl A=1
  run early_acqrel_2st $wl ICB_TMA_STAGES=2
  run early_mbar $wl ICB200_LIB=$V/libicb200_mbar.so
  run early_mbar_2st $wl ICB200_LIB=$V/libicb200_mbar.so ICB_TMA_STAGES=2
  run late_mbar $wl ICB200_LIB=$V/libicb200_mbar_late.so
  run producer_warp $wl ICB_PRODUCER_WARP=1
done
run early_acqrel etc1_rgb8 A=1
run early_mbar etc1_rgb8 ICB200_LIB=$V/libicb200_mbar.so
