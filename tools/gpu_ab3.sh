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
run producer dxt1_rgba8 ICB_PRODUCER_WARP=1
run producer_sleep1000 dxt1_rgba8 ICB_PRODUCER_WARP=1 ICB200_LIB=$V/libicb200_sleep1000.so
run producer_sleep5000 dxt1_rgba8 ICB_PRODUCER_WARP=1 ICB200_LIB=$V/libicb200_sleep5000.so
run ring56 dxt1_rgba8 ICB200_LIB=$V/libicb200_ring56.so
run ring48 dxt1_rgba8 ICB200_LIB=$V/libicb200_ring48.so
run ring48_2st dxt1_rgba8 ICB200_LIB=$V/libicb200_ring48.so ICB_TMA_STAGES=2
run ring64_2st dxt5_rgba8 ICB_TMA_STAGES=2
run ring56_2st dxt5_rgba8 ICB200_LIB=$V/libicb200_ring56.so ICB_TMA_STAGES=2
run ring48_2st dxt5_rgba8 ICB200_LIB=$V/libicb200_ring48.so ICB_TMA_STAGES=2
run producer_sleep1000 dxt5_rgba8 ICB_PRODUCER_WARP=1 ICB200_LIB=$V/libicb200_sleep1000.so
run producer_sleep1000 etc1_rgb8 ICB_PRODUCER_WARP=1 ICB200_LIB=$V/libicb200_sleep1000.so
run producer etc1_rgb8 ICB_PRODUCER_WARP=1
