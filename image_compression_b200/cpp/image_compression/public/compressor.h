// Compressor: abstract interface of the block-based texture compressors.
//
// Interface-compatible with the reference's image_compression/public/compressor.h:46-138.  Source images are
// 8 bits per component, interleaved RGB (3 bytes) or RGBA (4 bytes), row-major, top row first, with
// padding_bytes_per_row extra bytes after each row.  Argument order is (height, width) everywhere.  Functions
// returning a CompressedImage through an out-parameter allocate into a default-constructed instance or fill
// caller storage of exactly ComputeCompressedDataSize() bytes.  Errors are reported as `false`.
//
// In this build every method runs on the GPU (CUDA, sm_100a) through the C ABI in include/icb200.h; see
// INTEGRATION.md.
#ifndef IMAGE_COMPRESSION_PUBLIC_COMPRESSOR_H_
#define IMAGE_COMPRESSION_PUBLIC_COMPRESSOR_H_

#include <stddef.h>

#include <vector>

#include "base/integral_types.h"
#include "image_compression/public/compressed_image.h"

namespace image_codec_compression {

class Compressor {
 public:
  virtual ~Compressor() {}

  virtual bool SupportsFormat(CompressedImage::Format format) const = 0;
  virtual bool IsValidCompressedImage(const CompressedImage &image) = 0;
  virtual size_t ComputeCompressedDataSize(CompressedImage::Format format, uint32 height, uint32 width) = 0;

  // Declaration order IS the vtable layout and follows the reference header line for line
  // (public/compressor.h:52,61,68,77,85,95,105,114,125,134), so that an object compiled against the reference's
  // headers dispatches into this library correctly (tests/test_cpp_api.py::test_vtable_layout_matches_reference).
  // Compress and CompressAndPad are the hot path (GPU).
  virtual bool Compress(CompressedImage::Format format, uint32 height, uint32 width, uint32 padding_bytes_per_row,
                        const uint8 *buffer, CompressedImage *image) = 0;
  virtual bool Decompress(const CompressedImage &image, std::vector<uint8> *decompressed_buffer) = 0;
  virtual bool Downsample(const CompressedImage &image, CompressedImage *downsampled_image) = 0;
  virtual bool Pad(const CompressedImage &image, uint32 padded_height, uint32 padded_width,
                   CompressedImage *padded_image) = 0;
  virtual bool CompressAndPad(CompressedImage::Format format, uint32 height, uint32 width, uint32 padded_height,
                              uint32 padded_width, uint32 padding_bytes_per_row, const uint8 *buffer,
                              CompressedImage *padded_image) = 0;
  virtual bool CreateSolidImage(CompressedImage::Format format, uint32 height, uint32 width, const uint8 *color,
                                CompressedImage *image) = 0;
  virtual bool CopySubimage(const CompressedImage &image, uint32 start_row, uint32 start_column, uint32 height,
                            uint32 width, CompressedImage *subimage) = 0;
};

}  // namespace image_codec_compression

// The concrete compressors override the whole interface with identical signatures; they declare it through this
// list instead of repeating it three times.
#define IMAGE_CODEC_COMPRESSION_OVERRIDE_ALL()                                                                          \
  bool SupportsFormat(CompressedImage::Format format) const override;                                                   \
  bool IsValidCompressedImage(const CompressedImage &image) override;                                                   \
  size_t ComputeCompressedDataSize(CompressedImage::Format format, uint32 height, uint32 width) override;               \
  bool Compress(CompressedImage::Format format, uint32 height, uint32 width, uint32 padding_bytes_per_row,              \
                const uint8 *buffer, CompressedImage *image) override;                                                  \
  bool Decompress(const CompressedImage &image, std::vector<uint8> *decompressed_buffer) override;                      \
  bool Downsample(const CompressedImage &image, CompressedImage *downsampled_image) override;                           \
  bool Pad(const CompressedImage &image, uint32 padded_height, uint32 padded_width, CompressedImage *padded_image)      \
      override;                                                                                                         \
  bool CompressAndPad(CompressedImage::Format format, uint32 height, uint32 width, uint32 padded_height,                \
                      uint32 padded_width, uint32 padding_bytes_per_row, const uint8 *buffer,                           \
                      CompressedImage *padded_image) override;                                                          \
  bool CreateSolidImage(CompressedImage::Format format, uint32 height, uint32 width, const uint8 *color,                \
                        CompressedImage *image) override;                                                               \
  bool CopySubimage(const CompressedImage &image, uint32 start_row, uint32 start_column, uint32 height, uint32 width,   \
                    CompressedImage *subimage) override

#endif  // IMAGE_COMPRESSION_PUBLIC_COMPRESSOR_H_
