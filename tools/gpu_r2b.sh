#!/bin/bash
# Round-2 (second half) GPU-box visit: GPU test suite on the product library, then same-box A/B bench lines of the
# product library against named variants (tools/build_variants.sh) for the DXT workloads.
# Usage (under gpurun, from the repo root):  bash tools/gpu_r2b.sh <tag> "<workloads>" <variant specs for tools/gpu_ab.sh ...>
TAG=${1:-r2b}; WLS=${2:-"dxt1_rgba8 dxt1_rgb8 dxt5_rgba8"}; shift 2
OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi -L > $OUT/gpus.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -4 $OUT/pytest_gpu.log
for wl in $WLS; do
  bash tools/gpu_ab.sh $TAG $wl "$@"
done
