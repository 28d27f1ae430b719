#!/bin/bash
# Builds experimental variants of libicb200.so (extra -D flags) next to the product library, for A/B timing on the GPU
# box in one visit:  tools/build_variants.sh name1 "-DFLAG1 -DFLAG2" name2 "-DFLAG3" ...
# Select one with ICB200_LIB=image_compression_b200/lib/variants/libicb200_<name>.so python bench.py ...
set -e
cd "$(dirname "$0")/../image_compression_b200/csrc"
mkdir -p ../lib/variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-Wall,-fvisibility=hidden \
      --expt-relaxed-constexpr $flags -shared -cudart static -o ../lib/variants/libicb200_$name.so icb_api.cu &
done
wait
ls -la ../lib/variants
