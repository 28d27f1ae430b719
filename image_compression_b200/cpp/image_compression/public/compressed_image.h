// CompressedImage: result container of the Compressor API.
//
// Interface-compatible with the reference's image_compression/public/compressed_image.h:32-204 (same names,
// signatures, ownership rules), re-implemented for the B200 build:
//   * default-constructed  -> the compressor allocates with new uint8[] (CreateOwnedData) and the image frees it;
//   * (size, pointer) ctor -> caller-owned storage; the compressor checks GetDataSize() and only sets metadata.
#ifndef IMAGE_COMPRESSION_PUBLIC_COMPRESSED_IMAGE_H_
#define IMAGE_COMPRESSION_PUBLIC_COMPRESSED_IMAGE_H_

#include <stddef.h>

#include <cstring>
#include <string>

#include "base/integral_types.h"
#include "base/logging.h"

namespace image_codec_compression {

class CompressedImage {
 public:
  enum Format { kRGB, kBGR, kRGBA, kBGRA };

  struct Metadata {
    Metadata(Format format_in, const std::string &compressor_name_in, uint32 uncompressed_height_in,
             uint32 uncompressed_width_in, uint32 compressed_height_in, uint32 compressed_width_in,
             uint32 padding_bytes_per_row_in)
        : format(format_in),
          compressor_name(compressor_name_in),
          uncompressed_height(uncompressed_height_in),
          uncompressed_width(uncompressed_width_in),
          compressed_height(compressed_height_in),
          compressed_width(compressed_width_in),
          padding_bytes_per_row(padding_bytes_per_row_in) {}

    Format format;
    std::string compressor_name;   // "dxtc", "etc" or "pvrtc"
    uint32 uncompressed_height;    // source size in pixels
    uint32 uncompressed_width;
    uint32 compressed_height;      // size covered by the block grid
    uint32 compressed_width;
    uint32 padding_bytes_per_row;  // of the source rows
  };

  CompressedImage() : metadata_(kRGB, "", 0, 0, 0, 0, 0), data_size_(0), data_(NULL), owns_data_(true) {}

  CompressedImage(size_t data_size, uint8 *external_data)
      : metadata_(kRGB, "", 0, 0, 0, 0, 0), data_size_(data_size), data_(external_data), owns_data_(false) {
    DCHECK(external_data);
  }

  ~CompressedImage() {
    if (owns_data_) delete[] data_;
  }

  // Deep copy; this instance owns the copy whatever `from` did.
  void Duplicate(const CompressedImage &from) {
    if (&from == this && owns_data_) return;
    const uint8 *bytes = from.data_;
    const size_t size = from.data_size_;
    DCHECK(bytes);
    uint8 *fresh = new uint8[size];
    std::memcpy(fresh, bytes, size);
    const Metadata meta = from.metadata_;
    if (owns_data_) delete[] data_;
    metadata_ = meta;
    data_size_ = size;
    data_ = fresh;
    owns_data_ = true;
  }

  void CreateOwnedData(const Metadata &metadata, size_t data_size) {
    if (owns_data_) delete[] data_;
    metadata_ = metadata;
    data_size_ = data_size;
    data_ = new uint8[data_size];
    owns_data_ = true;
  }

  void SetMetadata(const Metadata &metadata) {
    DCHECK(!owns_data_);
    metadata_ = metadata;
  }

  const Metadata &GetMetadata() const { return metadata_; }
  bool OwnsData() const { return owns_data_; }
  size_t GetDataSize() const { return data_size_; }
  const uint8 *GetData() const { return data_; }
  uint8 *GetMutableData() { return data_; }

 private:
  Metadata metadata_;
  size_t data_size_;
  uint8 *data_;
  bool owns_data_;

  CompressedImage(const CompressedImage &);
  void operator=(const CompressedImage &);
};

inline int GetNumFormatComponents(CompressedImage::Format format) {
  switch (format) {
    case CompressedImage::kRGB:
    case CompressedImage::kBGR:
      return 3;
    case CompressedImage::kRGBA:
    case CompressedImage::kBGRA:
      return 4;
    default:
      return 0;
  }
}

inline bool NeedsRedAndBlueSwapped(CompressedImage::Format format) {
  return format == CompressedImage::kBGR || format == CompressedImage::kBGRA;
}

}  // namespace image_codec_compression

#endif  // IMAGE_COMPRESSION_PUBLIC_COMPRESSED_IMAGE_H_
