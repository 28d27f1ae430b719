// api_test_shim.cc -- extern "C" doorway used by tests/ to drive the C++ Compressor classes of this build exactly
// the way oracle/ref_shim.cc drives the reference's, so the two can be compared call for call.
#include <cstring>
#include <vector>

#include "image_compression/public/dxtc_compressor.h"
#include "image_compression/public/dxtc_to_etc_transcoder.h"
#include "image_compression/public/etc_compressor.h"
#include "image_compression/public/pvrtc_compressor.h"

using namespace image_codec_compression;  // NOLINT

namespace {
long Run(Compressor *c, int format, unsigned h, unsigned w, int pad_mode, unsigned ph, unsigned pw, unsigned padding,
         const unsigned char *src, unsigned char *dst, size_t dst_cap, unsigned *meta, int external) {
  const CompressedImage::Format f = static_cast<CompressedImage::Format>(format);
  CompressedImage owned;
  CompressedImage outside(dst_cap, dst);
  CompressedImage *image = external ? &outside : &owned;
  const bool ok = pad_mode ? c->CompressAndPad(f, h, w, ph, pw, padding, src, image)
                           : c->Compress(f, h, w, padding, src, image);
  if (!ok) return 0;
  if (!external) {
    if (image->GetDataSize() > dst_cap) return -1;
    std::memcpy(dst, image->GetData(), image->GetDataSize());
  }
  if (meta) {
    const CompressedImage::Metadata &m = image->GetMetadata();
    meta[0] = m.format; meta[1] = m.uncompressed_height; meta[2] = m.uncompressed_width;
    meta[3] = m.compressed_height; meta[4] = m.compressed_width; meta[5] = m.padding_bytes_per_row;
    meta[6] = static_cast<unsigned>(m.compressor_name.size());
  }
  return static_cast<long>(image->GetDataSize());
}
}  // namespace

extern "C" {
__attribute__((visibility("default"))) long icapi_dxt(int format, unsigned h, unsigned w, int pad_mode, unsigned ph,
                                                      unsigned pw, unsigned padding, const unsigned char *src,
                                                      unsigned char *dst, size_t dst_cap, unsigned *meta, int external) {
  DxtcCompressor c;
  return Run(&c, format, h, w, pad_mode, ph, pw, padding, src, dst, dst_cap, meta, external);
}
__attribute__((visibility("default"))) long icapi_etc(int strategy, int format, unsigned h, unsigned w, int pad_mode,
                                                      unsigned ph, unsigned pw, unsigned padding,
                                                      const unsigned char *src, unsigned char *dst, size_t dst_cap,
                                                      unsigned *meta, int external) {
  EtcCompressor c;
  c.SetCompressionStrategy(static_cast<EtcCompressor::CompressionStrategy>(strategy));
  return Run(&c, format, h, w, pad_mode, ph, pw, padding, src, dst, dst_cap, meta, external);
}
__attribute__((visibility("default"))) long icapi_pvrtc(int format, unsigned h, unsigned w, unsigned padding,
                                                        const unsigned char *src, unsigned char *dst, size_t dst_cap,
                                                        unsigned *meta, int external) {
  PvrtcCompressor c;
  return Run(&c, format, h, w, 0, 0, 0, padding, src, dst, dst_cap, meta, external);
}
// Compress then Decompress through the classes; returns decompressed bytes (0 on failure).
__attribute__((visibility("default"))) long icapi_roundtrip(int codec, int strategy, int format, unsigned h, unsigned w,
                                                            const unsigned char *src, unsigned char *dst, size_t dst_cap) {
  DxtcCompressor dxt;
  EtcCompressor etc;
  etc.SetCompressionStrategy(static_cast<EtcCompressor::CompressionStrategy>(strategy));
  Compressor *c = codec == 2 ? static_cast<Compressor *>(&etc) : static_cast<Compressor *>(&dxt);
  CompressedImage image;
  if (!c->Compress(static_cast<CompressedImage::Format>(format), h, w, 0, src, &image)) return 0;
  std::vector<uint8> out;
  if (!c->Decompress(image, &out)) return 0;
  if (out.size() > dst_cap) return -1;
  std::memcpy(dst, out.data(), out.size());
  return static_cast<long>(out.size());
}
// Compressed-domain operations through the classes, driven exactly like oracle/ref_shim.cc's icref_block_op:
// the input CompressedImage is made by compressing a dummy h x w image into external storage (which sets the
// metadata) and overwriting the storage with `blocks`.
// op: 0 Downsample, 1 Pad(a, b), 2 CopySubimage(a, b, c, d), 3 CreateSolidImage(colour = blocks[0..3]), 4 Transcode
__attribute__((visibility("default"))) long icapi_block_op(int op, int codec, int strategy, int format, unsigned h,
                                                           unsigned w, unsigned a, unsigned b, unsigned c, unsigned d,
                                                           const unsigned char *blocks, size_t nbytes,
                                                           unsigned char *dst, size_t dst_cap, unsigned *meta) {
  DxtcCompressor dxt;
  EtcCompressor etc;
  etc.SetCompressionStrategy(static_cast<EtcCompressor::CompressionStrategy>(strategy));
  Compressor *comp = codec == 2 ? static_cast<Compressor *>(&etc) : static_cast<Compressor *>(&dxt);
  const CompressedImage::Format f = static_cast<CompressedImage::Format>(format);
  CompressedImage out;
  bool ok = false;
  if (op == 3) {
    ok = comp->CreateSolidImage(f, h, w, blocks, &out);
  } else {
    std::vector<unsigned char> storage(nbytes), dummy(static_cast<size_t>(h) * w * 4, 0);
    CompressedImage in(nbytes, storage.data());
    if (!comp->Compress(f, h, w, 0, dummy.data(), &in)) return 0;
    std::memcpy(storage.data(), blocks, nbytes);
    if (op == 0) ok = comp->Downsample(in, &out);
    if (op == 1) ok = comp->Pad(in, a, b, &out);
    if (op == 2) ok = comp->CopySubimage(in, a, b, c, d, &out);
    if (op == 4) {
      TranscodeDxt1ToEtc1(&in);  // void, as in the reference; a failed call leaves the DXT1 blocks in place,
      ok = true;                 // which the caller's comparison with the expected ETC1 blocks then reports
      out.Duplicate(in);
    }
  }
  if (!ok) return 0;
  if (out.GetDataSize() > dst_cap) return -1;
  std::memcpy(dst, out.GetData(), out.GetDataSize());
  if (meta) {
    const CompressedImage::Metadata &m = out.GetMetadata();
    meta[0] = m.format; meta[1] = m.uncompressed_height; meta[2] = m.uncompressed_width;
    meta[3] = m.compressed_height; meta[4] = m.compressed_width; meta[5] = m.padding_bytes_per_row;
    meta[6] = static_cast<unsigned>(m.compressor_name.size());
  }
  return static_cast<long>(out.GetDataSize());
}
// Calls every one of the ten Compressor virtuals through a base-class pointer, in vtable order, on one small image.
// Results are appended to dst as segments; seg_sizes[k] receives the byte count of call k's result (for the three
// calls that return no image: one byte holding the bool, or the eight bytes of the size).  Returns a bit mask of the calls
// that succeeded (0x3ff when all ten did), or -1 when dst is too small.  tests/test_cpp_api.py builds this file twice
// -- against this build's headers and against the reference's own headers -- and requires identical output from
// both: that is the check that the vtable layout (declaration order in compressor.h) is link-compatible.
__attribute__((visibility("default"))) long icapi_all_virtuals(int codec, int format, unsigned h, unsigned w,
                                                               const unsigned char *src, unsigned char *dst,
                                                               size_t dst_cap, unsigned *seg_sizes) {
  DxtcCompressor dxt;
  EtcCompressor etc;
  Compressor *c = codec == 2 ? static_cast<Compressor *>(&etc) : static_cast<Compressor *>(&dxt);
  const CompressedImage::Format f = static_cast<CompressedImage::Format>(format);
  size_t used = 0;
  long mask = 0;
  bool overflow = false;
  auto put = [&](int k, bool ok, const void *data, size_t n) {
    seg_sizes[k] = 0;
    if (!ok) return;
    if (used + n > dst_cap) { overflow = true; return; }
    std::memcpy(dst + used, data, n);
    used += n;
    seg_sizes[k] = static_cast<unsigned>(n);
    mask |= 1L << k;
  };
  auto put_image = [&](int k, bool ok, const CompressedImage &im) {
    put(k, ok && im.GetData() != NULL, im.GetData(), ok ? im.GetDataSize() : 0);
  };
  const unsigned char yes = c->SupportsFormat(f) ? 1 : 0;
  put(0, true, &yes, 1);
  CompressedImage image;
  const bool compressed = c->Compress(f, h, w, 0, src, &image);
  const unsigned char valid = compressed && c->IsValidCompressedImage(image) ? 1 : 0;
  put(1, true, &valid, 1);
  const size_t size = c->ComputeCompressedDataSize(f, h, w);
  put(2, true, &size, sizeof(size));
  put_image(3, compressed, image);
  if (compressed) {
    std::vector<uint8> pixels;
    const bool ok4 = c->Decompress(image, &pixels);
    put(4, ok4, pixels.data(), ok4 ? pixels.size() : 0);
    CompressedImage half, padded, sub;
    put_image(5, c->Downsample(image, &half), half);
    put_image(6, c->Pad(image, h + 8, w + 12, &padded), padded);
    CompressedImage both;
    put_image(7, c->CompressAndPad(f, h, w, h + 8, w + 12, 0, src, &both), both);
    CompressedImage solid;
    put_image(8, c->CreateSolidImage(f, h, w, src, &solid), solid);
    put_image(9, c->CopySubimage(image, 4, 4, 8, 8, &sub), sub);
  }
  return overflow ? -1 : mask;
}
__attribute__((visibility("default"))) size_t icapi_size(int codec, int format, unsigned h, unsigned w) {
  const CompressedImage::Format f = static_cast<CompressedImage::Format>(format);
  if (codec == 0) return DxtcCompressor().ComputeCompressedDataSize(f, h, w);
  if (codec == 1) return EtcCompressor().ComputeCompressedDataSize(f, h, w);
  return PvrtcCompressor().ComputeCompressedDataSize(f, h, w);
}
}
