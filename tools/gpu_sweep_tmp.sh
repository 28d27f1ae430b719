ICB200_LIB=image_compression_b200/lib/variants/libicb200_x5np8.so timeout 300 python tools/debug_parity.py dxt5_rgba8 4
ICB200_LIB=image_compression_b200/lib/variants/libicb200_x5.so timeout 300 python tools/debug_parity.py dxt5_rgba8 3
timeout 300 python tools/debug_parity.py dxt5_rgba8 3
timeout 300 python tools/debug_parity.py dxt1_rgba8 3
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
bash tools/gpu_ab.sh v3 dxt5_rgba8 new np8:ICB200_LIB=image_compression_b200/lib/variants/libicb200_np8.so x5:ICB200_LIB=image_compression_b200/lib/variants/libicb200_x5.so base:ICB200_LIB=image_compression_b200/lib/variants/libicb200_base.so new2
bash tools/gpu_ab.sh v3 dxt1_rgba8 new base:ICB200_LIB=image_compression_b200/lib/variants/libicb200_base.so new2
bash tools/gpu_ab.sh v3 dxt1_rgb8 new new2
