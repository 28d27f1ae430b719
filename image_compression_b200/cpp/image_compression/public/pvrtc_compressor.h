// PvrtcCompressor: PVRTC1 2 bits-per-pixel RGBA; square power-of-two images only, no row
// padding.  Only Compress() is provided (as in the reference); everything else returns false.
//
// Interface-compatible with the reference's image_compression/public/pvrtc_compressor.h.  Compress() runs on the GPU
// (sm_100a CUDA kernels behind include/icb200.h) and produces byte-identical blocks; the other methods return false,
// as the reference's do (internal/pvrtc_compressor.cc:669-704).
#ifndef IMAGE_COMPRESSION_PUBLIC_PVRTC_COMPRESSOR_H_
#define IMAGE_COMPRESSION_PUBLIC_PVRTC_COMPRESSOR_H_

#include <stddef.h>

#include <vector>

#include "base/integral_types.h"
#include "image_compression/public/compressed_image.h"
#include "image_compression/public/compressor.h"

namespace image_codec_compression {

class PvrtcCompressor : public Compressor {
 public:
  PvrtcCompressor();
  ~PvrtcCompressor() override;

  IMAGE_CODEC_COMPRESSION_OVERRIDE_ALL();
};

}  // namespace image_codec_compression

#endif  // IMAGE_COMPRESSION_PUBLIC_PVRTC_COMPRESSOR_H_
