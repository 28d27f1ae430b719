bash tools/gpu_r2.sh r2c quick
bash tools/gpu_ab.sh r2c dxt5_rgba8 default x5:ICB200_LIB=$PWD/image_compression_b200/lib/variants/libicb200_dxt5x5.so ring3:ICB_TMA_STAGES=3 producer:ICB_DRIVER=producer
