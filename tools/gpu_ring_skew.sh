V=$PWD/image_compression_b200/lib/variants
for wl in dxt5_rgba8 dxt1_rgba8 dxt1_rgb8 etc1_rgb8; do
for drv in ring "ring,ICB_TMA_STAGES=3"; do
  bash tools/gpu_ab.sh r2g $wl skew_$drv:ICB200_LIB=$V/libicb200_skew.so,ICB_DRIVER=$drv
done; done
