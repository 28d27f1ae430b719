#!/bin/bash
# A/B timing of kernel / driver configurations on the GPU box, one bench line each.
#   bash tools/gpu_ab.sh <tag> <workload> NAME[:ENV=VAL,ENV=VAL...] ...
# e.g. after `tools/build_variants.sh probe "-DICB_PROBE_ENCODER"`:
#   bash tools/gpu_ab.sh ab1 dxt1_rgba8 default ring:ICB_DRIVER=ring ring2:ICB_DRIVER=ring,ICB_TMA_STAGES=2 \
#        probe:ICB200_LIB=$PWD/image_compression_b200/lib/variants/libicb200_probe.so
TAG=$1; WL=$2; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
for spec in "$@"; do
  name=${spec%%:*}; envs=""
  [ "$spec" != "$name" ] && envs=$(echo "${spec#*:}" | tr ',' ' ')
  env $envs timeout 200 python bench.py --workload $WL --steps 30 --warmup 5 --no-cpu-baseline --no-others > $OUT/${name}_$WL.json 2> $OUT/${name}_$WL.err
  python - "$OUT/${name}_$WL.json" "$name" "$WL" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1])); r = d["roofline"]
    print("%-22s %-12s b2b %.2f us  iso-med %.2f  min %.2f  frac %.3f  parity %s clocks %s %s" % (
        sys.argv[2], sys.argv[3], r["kernel_ms_avg"] * 1e3, r["kernel_ms_isolated_median"] * 1e3, r["kernel_ms_min"] * 1e3,
        r["frac"], d["parity"]["equal"], d["clocks"]["sm_mhz"], d["clocks"]["reasons"]))
except Exception as e:
    print(sys.argv[2], sys.argv[3], "failed", e)
PY
done
