// EtcCompressor: ETC1 for kRGB images, any image size.
//
// Interface-compatible with the reference's image_compression/public/etc_compressor.h.  Compress() and
// CompressAndPad() run on the GPU (sm_100a CUDA kernels behind include/icb200.h) and produce byte-identical
// blocks; so do Decompress(), Downsample(), Pad(), CopySubimage() and CreateSolidImage() (DESIGN.md sections 4.6, 4.7).
#ifndef IMAGE_COMPRESSION_PUBLIC_ETC_COMPRESSOR_H_
#define IMAGE_COMPRESSION_PUBLIC_ETC_COMPRESSOR_H_

#include <stddef.h>

#include <vector>

#include "base/integral_types.h"
#include "image_compression/public/compressed_image.h"
#include "image_compression/public/compressor.h"

namespace image_codec_compression {

class EtcCompressor : public Compressor {
 public:
  // How each 4x4 block is split into two sub-blocks (default kSmallerError).
  enum CompressionStrategy {
    kSplitHorizontally,  // two 4x2 halves, top / bottom
    kSplitVertically,    // two 2x4 halves, left / right
    kSmallerError,       // try both, keep the smaller error
    kHeuristic,          // pick split and codewords by heuristic, no exhaustive search
  };

  EtcCompressor();
  ~EtcCompressor() override;

  void SetCompressionStrategy(CompressionStrategy strategy) { compression_strategy_ = strategy; }
  CompressionStrategy GetCompressionStrategy() const { return compression_strategy_; }

  IMAGE_CODEC_COMPRESSION_OVERRIDE_ALL();

 private:
  CompressionStrategy compression_strategy_;
};

}  // namespace image_codec_compression

#endif  // IMAGE_COMPRESSION_PUBLIC_ETC_COMPRESSOR_H_
