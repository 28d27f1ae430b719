#!/bin/bash
# First GPU visit of the next round: the experiments this round prepared on the CPU but could not measure.
#   (in the container)  tools/build_variants.sh dxt5x5 "-DICB_DXT5_RING_MIN_CTAS=5" mod7 "-DICB_PVRTC_MOD_MIN_CTAS=7" \
#                                               mod8 "-DICB_PVRTC_MOD_MIN_CTAS=8"
#   (under gpurun)      bash tools/gpu_next_round.sh <tag>
# 1. DXT5 ring kernel compiled for five resident CTAs per SM (48 registers, 8 bytes of spill; block4x4_kernels.cuh):
#    parity of the variant library through the whole 4x4 GPU suite, then A/B against the product build.
TAG=${1:-nr1}; OUT=gpurun_out/$TAG; mkdir -p $OUT
V=$PWD/image_compression_b200/lib/variants
if [ -f $V/libicb200_dxt5x5.so ]; then
  ICB200_LIB=$V/libicb200_dxt5x5.so timeout 300 python -m pytest tests -x -q -m gpu -k "not pvrtc and not multigpu" > $OUT/pytest_dxt5x5.log 2>&1
  echo "variant parity exit $?"; tail -2 $OUT/pytest_dxt5x5.log
  bash tools/gpu_ab.sh $TAG dxt5_rgba8 base dxt5x5:ICB200_LIB=$V/libicb200_dxt5x5.so base2 dxt5x5b:ICB200_LIB=$V/libicb200_dxt5x5.so
fi
# 2. PVRTC Modulate compiled for 7 / 8 resident CTAs per SM (72 registers, no spill / 64 registers, 32 bytes of spill)
#    instead of the 80 registers ptxas takes by itself (pvrtc_kernels.cuh).
if [ -f $V/libicb200_mod7.so ]; then
  for v in mod7 mod8; do
    ICB200_LIB=$V/libicb200_$v.so timeout 120 python -m pytest tests -x -q -m gpu -k "pvrtc or Pvrtc" > $OUT/pytest_$v.log 2>&1; echo "$v parity exit $?"
  done
  bash tools/gpu_ab.sh $TAG pvrtc2_rgba8 base mod7:ICB200_LIB=$V/libicb200_mod7.so mod8:ICB200_LIB=$V/libicb200_mod8.so base2
fi
if [ ! -f $V/libicb200_dxt5x5.so ] && [ ! -f $V/libicb200_mod7.so ]; then
  echo "build the variants first: tools/build_variants.sh dxt5x5 \"-DICB_DXT5_RING_MIN_CTAS=5\""
fi
