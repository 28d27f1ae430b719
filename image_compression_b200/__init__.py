"""image_compression_b200 -- B200 (sm_100a) texture-block encoder behind the google/image-compression API.

Python here is only the test/bench harness over the C ABI (include/icb200.h); the product is
lib/libicb200.so (CUDA kernels + extern "C" entry points) and lib/libimagecompression_b200.so (the C++
`image_codec_compression::*Compressor` classes that call it).  There is no CPU fallback: loading fails loudly if
the CUDA library has not been built, and every compute call fails if no CUDA device is present.
"""
from .binding import (  # noqa: F401
    BGR, BGRA, RGB, RGBA, CODEC_DXT1, CODEC_DXT5, CODEC_ETC1, CODEC_PVRTC2,
    ETC_HEURISTIC, ETC_SMALLER_ERROR, ETC_SPLIT_HORIZONTALLY, ETC_SPLIT_VERTICALLY,
    IcbError, compress_host, compressed_size, decode_device, decompress_host, encode_device, encode_stripe_device, fill_synthetic, launch_count,
    lib, lib_path, pvrtc_encode_device, pvrtc_encode_stripe_device, set_tma_mode, stripe_rows,
    OP_COPY_SUBIMAGE, OP_DOWNSAMPLE, OP_PAD, OP_SOLID, OP_TRANSCODE, block_bytes, blockop_host, copy_subimage_device,
    downsample_device, fill_solid_device, pad_device, transcode_dxt1_to_etc1_device,
    ShardContext, root_share_permille, stripe_partition,
)
