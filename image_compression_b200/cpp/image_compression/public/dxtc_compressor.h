// DxtcCompressor: DXT1 (3-component formats) / DXT5 (4-component formats), any image size;
// windows that cross the image edge replicate the last row / column.
//
// Interface-compatible with the reference's image_compression/public/dxtc_compressor.h.  Compress() and
// CompressAndPad() run on the GPU (sm_100a CUDA kernels behind include/icb200.h) and produce byte-identical
// blocks; so do Decompress(), Downsample(), Pad(), CopySubimage() and CreateSolidImage() (DESIGN.md sections 4.6, 4.7).
#ifndef IMAGE_COMPRESSION_PUBLIC_DXTC_COMPRESSOR_H_
#define IMAGE_COMPRESSION_PUBLIC_DXTC_COMPRESSOR_H_

#include <stddef.h>

#include <vector>

#include "base/integral_types.h"
#include "image_compression/public/compressed_image.h"
#include "image_compression/public/compressor.h"

namespace image_codec_compression {

class DxtcCompressor : public Compressor {
 public:
  DxtcCompressor();
  ~DxtcCompressor() override;

  IMAGE_CODEC_COMPRESSION_OVERRIDE_ALL();
};

}  // namespace image_codec_compression

#endif  // IMAGE_COMPRESSION_PUBLIC_DXTC_COMPRESSOR_H_
