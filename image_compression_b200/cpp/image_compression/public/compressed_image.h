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

  // Empty image that will own whatever a compressor allocates for it.
  CompressedImage();
  // Image over caller-owned storage of exactly the size the compressor will need.
  CompressedImage(size_t data_size, uint8 *external_data);
  ~CompressedImage() { ReleaseOwned(); }

  // Deep copy; this instance owns the copy whatever `from` did.
  void Duplicate(const CompressedImage &from);
  // Replaces the contents by a fresh, owned, uninitialised buffer of data_size bytes.
  void CreateOwnedData(const Metadata &metadata, size_t data_size) { Adopt(metadata, data_size, new uint8[data_size]); }
  // Only for images over external storage: the bytes are the caller's, the description is ours.
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
  void ReleaseOwned() {
    if (owns_data_) delete[] data_;
  }
  void Adopt(const Metadata &metadata, size_t data_size, uint8 *owned_bytes) {
    ReleaseOwned();
    metadata_ = metadata;
    data_size_ = data_size;
    data_ = owned_bytes;
    owns_data_ = true;
  }

  Metadata metadata_;
  size_t data_size_;
  uint8 *data_;
  bool owns_data_;

  CompressedImage(const CompressedImage &);  // not copyable: use Duplicate
  void operator=(const CompressedImage &);
};

inline CompressedImage::CompressedImage() : metadata_(kRGB, "", 0, 0, 0, 0, 0), data_size_(0), data_(NULL), owns_data_(true) {}

inline CompressedImage::CompressedImage(size_t data_size, uint8 *external_data)
    : metadata_(kRGB, "", 0, 0, 0, 0, 0), data_size_(data_size), data_(external_data), owns_data_(false) {
  DCHECK(external_data);
}

inline void CompressedImage::Duplicate(const CompressedImage &from) {
  if (&from == this && owns_data_) return;  // already an owned copy of itself
  DCHECK(from.data_);
  // copy first, release afterwards: `from` may be this very image over external storage
  uint8 *fresh = new uint8[from.data_size_];
  std::memcpy(fresh, from.data_, from.data_size_);
  const Metadata meta = from.metadata_;
  Adopt(meta, from.data_size_, fresh);
}

// 3 for kRGB / kBGR, 4 for kRGBA / kBGRA, 0 for anything else.
inline int GetNumFormatComponents(CompressedImage::Format format) {
  const unsigned f = static_cast<unsigned>(format);
  return f < 2u ? 3 : (f < 4u ? 4 : 0);
}

inline bool NeedsRedAndBlueSwapped(CompressedImage::Format format) {
  return format == CompressedImage::kBGR || format == CompressedImage::kBGRA;
}

}  // namespace image_codec_compression

#endif  // IMAGE_COMPRESSION_PUBLIC_COMPRESSED_IMAGE_H_
