// compressors.cc -- the image_codec_compression::{Dxtc,Etc,Pvrtc}Compressor classes of the B200 build.
//
// Host-side mirror of the reference's public methods: identical argument checks, metadata and ownership rules
//   DxtcCompressor::Compress / CompressAndPad   reference internal/dxtc_compressor.cc:735-750, 799-818
//   EtcCompressor::Compress / CompressAndPad    reference internal/etc_compressor.cc:747-758, 787-800
//   PvrtcCompressor::Compress                   reference internal/pvrtc_compressor.cc:636-667
//   SetUpCompressedImage                        reference internal/compressor4x4_helper.cc:22-43
// ...and then one call through the C ABI (include/icb200.h, icb_compress_host) where the reference ran its
// per-block CPU loop.  No CPU encoder exists in this library: if the CUDA call fails, Compress returns false.
#include <algorithm>
#include <cstring>
#include <string>

#include "icb200.h"
#include "image_compression/public/dxtc_compressor.h"
#include "image_compression/public/dxtc_to_etc_transcoder.h"
#include "image_compression/public/etc_compressor.h"
#include "image_compression/public/pvrtc_compressor.h"

namespace image_codec_compression {

namespace {

inline uint32 NumBlocks(uint32 pixels) { return (pixels + 3) / 4; }

// Gives `image` storage for a grid covering coded_height x coded_width and fills its metadata; mirrors
// SetUpCompressedImage, including the exact-size check for caller-owned storage.
bool PrepareOutput(const char *name, size_t block_size, CompressedImage::Format format, uint32 coded_height,
                   uint32 coded_width, uint32 padding_bytes_per_row, CompressedImage *image) {
  const uint32 rows = NumBlocks(coded_height), cols = NumBlocks(coded_width);
  const size_t size = static_cast<size_t>(rows) * cols * block_size;
  const CompressedImage::Metadata meta(format, name, coded_height, coded_width, 4 * rows, 4 * cols,
                                       padding_bytes_per_row);
  if (image->OwnsData()) {
    image->CreateOwnedData(meta, size);
  } else {
    if (image->GetDataSize() != size) return false;
    image->SetMetadata(meta);
  }
  return true;
}

bool Encode4x4(int codec, const char *name, size_t block_size, CompressedImage::Format format, uint32 height,
               uint32 width, uint32 padded_height, uint32 padded_width, uint32 padding, int strategy,
               const uint8 *buffer, CompressedImage *image) {
  // CompressAndPad reports the padded size as the "uncompressed" size (compressor4x4_helper.h:486-492).
  const uint32 coded_h = std::max(height, padded_height), coded_w = std::max(width, padded_width);
  if (!PrepareOutput(name, block_size, format, coded_h, coded_w, padding, image)) return false;
  return icb_compress_host(codec, static_cast<int>(format), height, width, padded_height, padded_width, padding,
                           strategy, buffer, image->GetMutableData(), image->GetDataSize()) == ICB_OK;
}

bool IsPowerOfTwo(uint32 x) { return x != 0 && (x & (x - 1)) == 0; }

// Compressor4x4Helper::Decompress (internal/compressor4x4_helper.h:218-262) on the GPU: the buffer is resized to
// uncompressed_height * uncompressed_width pixels and blocks are walked ceil(uncompressed_width / 4) per row, as
// the reference does.  The reference writes rows padding_bytes_per_row apart into a buffer that has no room for the
// padding (undefined behaviour when padding != 0); this build refuses that case instead.
bool Decode4x4(int codec, const CompressedImage &image, std::vector<uint8> *out) {
  const CompressedImage::Metadata &m = image.GetMetadata();
  if (m.padding_bytes_per_row != 0) return false;
  const size_t ncomp = codec == ICB_CODEC_DXT5 ? 4 : 3;
  out->resize(static_cast<size_t>(m.uncompressed_height) * m.uncompressed_width * ncomp);
  return icb_decompress_host(codec, static_cast<int>(m.format), m.uncompressed_height, m.uncompressed_width,
                             NumBlocks(m.uncompressed_width), image.GetData(), image.GetDataSize(), &out->at(0),
                             out->size()) == ICB_OK;
}

// ---- compressed-domain operations ---------------------------------------------
// Each mirrors the reference's host logic (validity checks, sizes, metadata) and hands the block arithmetic to the
// C ABI (icb_blockop_host).  CopySubimage is a pure row-wise memcpy of blocks -- no arithmetic -- and stays on the
// host, as in the reference (a device-resident form, icb_copy_subimage4x4, exists for streams that live in HBM).

// Compressor4x4Helper::Downsample (internal/compressor4x4_helper.h:264-391)
bool Downsample4x4(int codec, size_t block_size, int strategy, const CompressedImage &image, CompressedImage *out) {
  const CompressedImage::Metadata &m = image.GetMetadata();
  const uint32 rows = NumBlocks(m.uncompressed_height), cols = NumBlocks(m.uncompressed_width);
  if ((rows > 1 && rows % 2 != 0) || (cols > 1 && cols % 2 != 0)) return false;
  const uint32 dh = (m.uncompressed_height + 1) / 2, dw = (m.uncompressed_width + 1) / 2;
  if (!PrepareOutput(m.compressor_name.c_str(), block_size, m.format, dh, dw, 0, out)) return false;
  // the reference sets the output up first and only then refuses a 3-pixel single block (:335)
  if (rows == 1 && cols == 1 && (m.uncompressed_height == 3 || m.uncompressed_width == 3)) return false;
  const uint32 args[2] = {m.uncompressed_height, m.uncompressed_width};
  return icb_blockop_host(ICB_OP_DOWNSAMPLE, codec, strategy, args, image.GetData(), image.GetDataSize(),
                          out->GetMutableData(), out->GetDataSize()) == ICB_OK;
}

// Compressor4x4Helper::Pad (:393-477)
bool Pad4x4(int codec, size_t block_size, int strategy, const CompressedImage &image, uint32 padded_height,
            uint32 padded_width, CompressedImage *out) {
  const CompressedImage::Metadata &m = image.GetMetadata();
  if (m.compressed_height >= padded_height && m.compressed_width >= padded_width) {
    out->Duplicate(image);
    return true;
  }
  // The reference overruns its output buffer when one dimension shrinks below the input's block count while the
  // other grows; refuse before touching `out`.
  if (NumBlocks(padded_height) < NumBlocks(m.compressed_height) || NumBlocks(padded_width) < NumBlocks(m.compressed_width))
    return false;
  if (!PrepareOutput(m.compressor_name.c_str(), block_size, m.format, padded_height, padded_width, 0, out)) return false;
  const uint32 args[4] = {m.compressed_height, m.compressed_width, padded_height, padded_width};
  return icb_blockop_host(ICB_OP_PAD, codec, strategy, args, image.GetData(), image.GetDataSize(), out->GetMutableData(),
                          out->GetDataSize()) == ICB_OK;
}

// Compressor4x4Helper::CreateSolidImage (:522-545)
bool Solid4x4(int codec, const char *name, size_t block_size, CompressedImage::Format format, uint32 height,
              uint32 width, const uint8 *color, CompressedImage *out) {
  if (!PrepareOutput(name, block_size, format, height, width, 0, out)) return false;
  if (out->GetDataSize() == 0) return true;  // zero-sized image: nothing to fill (the reference loops zero times)
  const uint32 packed = color[0] | (color[1] << 8) | (color[2] << 16) |
                        (codec == ICB_CODEC_DXT5 ? static_cast<uint32>(color[3]) << 24 : 0u);
  const uint32 args[3] = {height, width, packed};
  return icb_blockop_host(ICB_OP_SOLID, codec, 0, args, NULL, 0, out->GetMutableData(), out->GetDataSize()) == ICB_OK;
}

// Compressor4x4Helper::CopySubimage (:547-592)
bool CopySubimage4x4(size_t block_size, const CompressedImage &image, uint32 start_row, uint32 start_column,
                     uint32 height, uint32 width, CompressedImage *out) {
  const CompressedImage::Metadata &m = image.GetMetadata();
  if (start_row % 4 != 0 || start_column % 4 != 0 || height % 4 != 0 || width % 4 != 0 ||
      start_row > m.compressed_height || start_column > m.compressed_width ||
      start_row + height > m.compressed_height || start_column + width > m.compressed_width)
    return false;
  if (!PrepareOutput(m.compressor_name.c_str(), block_size, m.format, height, width, 0, out)) return false;
  const size_t src_cols = NumBlocks(m.compressed_width), dst_cols = NumBlocks(width);
  const uint8 *src = image.GetData() + (static_cast<size_t>(start_row / 4) * src_cols + start_column / 4) * block_size;
  uint8 *dst = out->GetMutableData();
  for (uint32 r = 0; r < NumBlocks(height); ++r, src += src_cols * block_size, dst += dst_cols * block_size)
    std::memcpy(dst, src, dst_cols * block_size);
  return true;
}

}  // namespace

// ---------------------------------------------------------------------------------------------------------
// DXT1 / DXT5
// ---------------------------------------------------------------------------------------------------------

DxtcCompressor::DxtcCompressor() {}
DxtcCompressor::~DxtcCompressor() {}

bool DxtcCompressor::SupportsFormat(CompressedImage::Format) const { return true; }

size_t DxtcCompressor::ComputeCompressedDataSize(CompressedImage::Format format, uint32 height, uint32 width) {
  if (height == 0 || width == 0) return 0;
  const size_t block = GetNumFormatComponents(format) == 3 ? 8 : 16;
  return static_cast<size_t>(std::max(1u, NumBlocks(height))) * std::max(1u, NumBlocks(width)) * block;
}

bool DxtcCompressor::IsValidCompressedImage(const CompressedImage &image) {
  const CompressedImage::Metadata &m = image.GetMetadata();
  return m.compressor_name == "dxtc" && m.uncompressed_height > 0 && m.uncompressed_width > 0 &&
         m.compressed_height >= m.uncompressed_height && m.compressed_width >= m.uncompressed_width &&
         image.GetDataSize() == ComputeCompressedDataSize(m.format, m.compressed_height, m.compressed_width);
}

bool DxtcCompressor::Compress(CompressedImage::Format format, uint32 height, uint32 width,
                              uint32 padding_bytes_per_row, const uint8 *buffer, CompressedImage *image) {
  if (!buffer || !image || height == 0 || width == 0) return false;
  const bool dxt1 = GetNumFormatComponents(format) == 3;
  return Encode4x4(dxt1 ? ICB_CODEC_DXT1 : ICB_CODEC_DXT5, "dxtc", dxt1 ? 8 : 16, format, height, width, 0, 0,
                   padding_bytes_per_row, 0, buffer, image);
}

bool DxtcCompressor::CompressAndPad(CompressedImage::Format format, uint32 height, uint32 width,
                                    uint32 padded_height, uint32 padded_width, uint32 padding_bytes_per_row,
                                    const uint8 *buffer, CompressedImage *padded_image) {
  if (!buffer || !padded_image || height == 0 || width == 0) return false;
  const bool dxt1 = GetNumFormatComponents(format) == 3;
  return Encode4x4(dxt1 ? ICB_CODEC_DXT1 : ICB_CODEC_DXT5, "dxtc", dxt1 ? 8 : 16, format, height, width,
                   padded_height, padded_width, padding_bytes_per_row, 0, buffer, padded_image);
}

bool DxtcCompressor::Decompress(const CompressedImage &image, std::vector<uint8> *decompressed_buffer) {
  if (!IsValidCompressedImage(image) || !decompressed_buffer) return false;
  const bool dxt1 = GetNumFormatComponents(image.GetMetadata().format) == 3;
  return Decode4x4(dxt1 ? ICB_CODEC_DXT1 : ICB_CODEC_DXT5, image, decompressed_buffer);
}

bool DxtcCompressor::Downsample(const CompressedImage &image, CompressedImage *downsampled_image) {
  if (!IsValidCompressedImage(image) || !downsampled_image) return false;
  const bool dxt1 = GetNumFormatComponents(image.GetMetadata().format) == 3;
  return Downsample4x4(dxt1 ? ICB_CODEC_DXT1 : ICB_CODEC_DXT5, dxt1 ? 8 : 16, 0, image, downsampled_image);
}

bool DxtcCompressor::Pad(const CompressedImage &image, uint32 padded_height, uint32 padded_width,
                         CompressedImage *padded_image) {
  if (!IsValidCompressedImage(image) || !padded_image) return false;
  const bool dxt1 = GetNumFormatComponents(image.GetMetadata().format) == 3;
  return Pad4x4(dxt1 ? ICB_CODEC_DXT1 : ICB_CODEC_DXT5, dxt1 ? 8 : 16, 0, image, padded_height, padded_width, padded_image);
}

bool DxtcCompressor::CreateSolidImage(CompressedImage::Format format, uint32 height, uint32 width, const uint8 *color,
                                      CompressedImage *image) {
  if (!image) return false;
  const bool dxt1 = GetNumFormatComponents(format) == 3;
  return Solid4x4(dxt1 ? ICB_CODEC_DXT1 : ICB_CODEC_DXT5, "dxtc", dxt1 ? 8 : 16, format, height, width, color, image);
}

bool DxtcCompressor::CopySubimage(const CompressedImage &image, uint32 start_row, uint32 start_column, uint32 height,
                                  uint32 width, CompressedImage *subimage) {
  if (!IsValidCompressedImage(image) || !subimage) return false;
  return CopySubimage4x4(GetNumFormatComponents(image.GetMetadata().format) == 3 ? 8 : 16, image, start_row,
                         start_column, height, width, subimage);
}

// ---------------------------------------------------------------------------------------------------------
// ETC1
// ---------------------------------------------------------------------------------------------------------

EtcCompressor::EtcCompressor() : compression_strategy_(kSmallerError) {}
EtcCompressor::~EtcCompressor() {}

bool EtcCompressor::SupportsFormat(CompressedImage::Format format) const { return format == CompressedImage::kRGB; }

size_t EtcCompressor::ComputeCompressedDataSize(CompressedImage::Format format, uint32 height, uint32 width) {
  if (height == 0 || width == 0 || format != CompressedImage::kRGB) return 0;
  return static_cast<size_t>(std::max(1u, NumBlocks(height))) * std::max(1u, NumBlocks(width)) * 8;
}

bool EtcCompressor::IsValidCompressedImage(const CompressedImage &image) {
  const CompressedImage::Metadata &m = image.GetMetadata();
  return m.format == CompressedImage::kRGB && m.compressor_name == "etc" && m.uncompressed_height > 0 &&
         m.uncompressed_width > 0 && m.compressed_height >= m.uncompressed_height &&
         m.compressed_width >= m.uncompressed_width &&
         image.GetDataSize() == static_cast<size_t>(NumBlocks(m.compressed_height)) * NumBlocks(m.compressed_width) * 8;
}

bool EtcCompressor::Compress(CompressedImage::Format format, uint32 height, uint32 width,
                             uint32 padding_bytes_per_row, const uint8 *buffer, CompressedImage *image) {
  if (!buffer || !image || height == 0 || width == 0 || format != CompressedImage::kRGB) return false;
  return Encode4x4(ICB_CODEC_ETC1, "etc", 8, format, height, width, 0, 0, padding_bytes_per_row,
                   static_cast<int>(compression_strategy_), buffer, image);
}

bool EtcCompressor::CompressAndPad(CompressedImage::Format format, uint32 height, uint32 width, uint32 padded_height,
                                   uint32 padded_width, uint32 padding_bytes_per_row, const uint8 *buffer,
                                   CompressedImage *padded_image) {
  if (!buffer || !padded_image || height == 0 || width == 0 || format != CompressedImage::kRGB) return false;
  return Encode4x4(ICB_CODEC_ETC1, "etc", 8, format, height, width, padded_height, padded_width,
                   padding_bytes_per_row, static_cast<int>(compression_strategy_), buffer, padded_image);
}

bool EtcCompressor::Decompress(const CompressedImage &image, std::vector<uint8> *decompressed_buffer) {
  if (!IsValidCompressedImage(image) || !decompressed_buffer) return false;
  return Decode4x4(ICB_CODEC_ETC1, image, decompressed_buffer);
}

bool EtcCompressor::Downsample(const CompressedImage &image, CompressedImage *downsampled_image) {
  if (!IsValidCompressedImage(image) || !downsampled_image) return false;
  return Downsample4x4(ICB_CODEC_ETC1, 8, static_cast<int>(compression_strategy_), image, downsampled_image);
}

bool EtcCompressor::Pad(const CompressedImage &image, uint32 padded_height, uint32 padded_width,
                        CompressedImage *padded_image) {
  if (!IsValidCompressedImage(image) || !padded_image) return false;
  return Pad4x4(ICB_CODEC_ETC1, 8, static_cast<int>(compression_strategy_), image, padded_height, padded_width, padded_image);
}

bool EtcCompressor::CreateSolidImage(CompressedImage::Format format, uint32 height, uint32 width, const uint8 *color,
                                     CompressedImage *image) {
  if (!image || format != CompressedImage::kRGB) return false;
  return Solid4x4(ICB_CODEC_ETC1, "etc", 8, format, height, width, color, image);
}

bool EtcCompressor::CopySubimage(const CompressedImage &image, uint32 start_row, uint32 start_column, uint32 height,
                                 uint32 width, CompressedImage *subimage) {
  if (!IsValidCompressedImage(image) || !subimage) return false;
  return CopySubimage4x4(8, image, start_row, start_column, height, width, subimage);
}

// ---------------------------------------------------------------------------------------------------------
// PVRTC1 2bpp
// ---------------------------------------------------------------------------------------------------------

PvrtcCompressor::PvrtcCompressor() {}
PvrtcCompressor::~PvrtcCompressor() {}

bool PvrtcCompressor::SupportsFormat(CompressedImage::Format format) const { return format == CompressedImage::kRGBA; }

size_t PvrtcCompressor::ComputeCompressedDataSize(CompressedImage::Format, uint32 height, uint32 width) {
  return static_cast<size_t>(width * height / 4);  // 32-bit product, as in the reference
}

bool PvrtcCompressor::IsValidCompressedImage(const CompressedImage &image) {
  const CompressedImage::Metadata &m = image.GetMetadata();
  return m.format == CompressedImage::kRGBA && m.compressor_name == "pvrtc" && m.uncompressed_height >= 4 &&
         m.uncompressed_width >= 8 && m.compressed_width == m.compressed_height &&
         IsPowerOfTwo(m.uncompressed_height) && IsPowerOfTwo(m.uncompressed_width) &&
         m.compressed_height == m.uncompressed_height && m.compressed_width == m.uncompressed_width &&
         image.GetDataSize() == ComputeCompressedDataSize(m.format, m.uncompressed_height, m.uncompressed_width);
}

bool PvrtcCompressor::Compress(CompressedImage::Format format, uint32 height, uint32 width,
                               uint32 padding_bytes_per_row, const uint8 *buffer, CompressedImage *image) {
  if (!buffer || !image || height == 0 || width == 0) return false;
  if (!IsPowerOfTwo(width) || !IsPowerOfTwo(height) || width != height) return false;
  if (padding_bytes_per_row != 0) return false;
  if (width % 8 != 0 || height % 4 != 0) return false;
  // `format` is recorded but not checked: the pixels are always read as RGBA8888 (pvrtc_compressor.cc:664).
  const size_t size = ComputeCompressedDataSize(format, height, width);
  const CompressedImage::Metadata meta(format, "pvrtc", height, width, height, width, 0);
  if (image->OwnsData()) {
    image->CreateOwnedData(meta, size);
  } else {
    if (image->GetDataSize() != size) return false;
    image->SetMetadata(meta);
  }
  return icb_compress_host(ICB_CODEC_PVRTC2, ICB_RGBA, height, width, 0, 0, 0, 0, buffer, image->GetMutableData(),
                           size) == ICB_OK;
}

bool PvrtcCompressor::CompressAndPad(CompressedImage::Format, uint32, uint32, uint32, uint32, uint32, const uint8 *,
                                     CompressedImage *) {
  return false;
}
bool PvrtcCompressor::Decompress(const CompressedImage &, std::vector<uint8> *) { return false; }
bool PvrtcCompressor::Downsample(const CompressedImage &, CompressedImage *) { return false; }
bool PvrtcCompressor::Pad(const CompressedImage &, uint32, uint32, CompressedImage *) { return false; }
bool PvrtcCompressor::CreateSolidImage(CompressedImage::Format, uint32, uint32, const uint8 *, CompressedImage *) {
  return false;
}
bool PvrtcCompressor::CopySubimage(const CompressedImage &, uint32, uint32, uint32, uint32, CompressedImage *) {
  return false;
}

// internal/dxtc_to_etc_transcoder.cc:29-40: every 8 bytes of the image are read as a DXT1 block and overwritten with
// the ETC1 (heuristic strategy) encoding of its 16 decoded pixels.  Like the reference, the metadata is left alone.
// The reference returns void; this build returns whether the GPU call succeeded (the data is untouched if not).
bool TranscodeDxt1ToEtc1Checked(CompressedImage *image) {
  if (!image || !image->GetMutableData()) return false;
  const size_t bytes = image->GetDataSize() / 8 * 8;
  if (bytes == 0) return true;
  return icb_blockop_host(ICB_OP_TRANSCODE, ICB_CODEC_DXT1, ICB_ETC_HEURISTIC, NULL, image->GetData(), bytes,
                          image->GetMutableData(), bytes) == ICB_OK;
}

void TranscodeDxt1ToEtc1(CompressedImage *image) { (void)TranscodeDxt1ToEtc1Checked(image); }

}  // namespace image_codec_compression
