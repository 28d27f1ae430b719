// ref_shim.cc -- TEST INFRASTRUCTURE.  extern "C" doorway into the UNMODIFIED reference library.
//
// oracle/Makefile compiles this file together with the reference's own sources, taken where they lie under
// /root/reference (never copied into the repo), into oracle/_ref/libicref.so.  tests/ use it to pin the C
// restatement (texblock_oracle.c) and tools/gen_golden.py uses it to produce tests/golden/.  bench.py may time it
// as the "reference" CPU baseline.  The product never links it.
#include <cstring>
#include <vector>

#include "image_compression/public/compressed_image.h"
#include "image_compression/public/dxtc_compressor.h"
#include "image_compression/public/dxtc_to_etc_transcoder.h"
#include "image_compression/public/etc_compressor.h"
#include "image_compression/public/pvrtc_compressor.h"

using image_codec_compression::CompressedImage;
using image_codec_compression::Compressor;
using image_codec_compression::DxtcCompressor;
using image_codec_compression::EtcCompressor;
using image_codec_compression::PvrtcCompressor;

namespace {

// Runs Compress / CompressAndPad into caller storage.  Returns bytes produced, 0 when the reference said false,
// and reports the metadata the reference filled in (7 uint32: format, uh, uw, ch, cw, padding, name length).
long run(Compressor *c, int format, unsigned h, unsigned w, int pad_mode, unsigned ph, unsigned pw, unsigned padding,
         const unsigned char *src, unsigned char *dst, size_t dst_cap, unsigned *meta) {
  CompressedImage image;
  const CompressedImage::Format f = static_cast<CompressedImage::Format>(format);
  const bool ok = pad_mode ? c->CompressAndPad(f, h, w, ph, pw, padding, src, &image)
                           : c->Compress(f, h, w, padding, src, &image);
  if (!ok) return 0;
  if (image.GetDataSize() > dst_cap) return -1;
  std::memcpy(dst, image.GetData(), image.GetDataSize());
  if (meta) {
    const CompressedImage::Metadata &m = image.GetMetadata();
    meta[0] = m.format;
    meta[1] = m.uncompressed_height;
    meta[2] = m.uncompressed_width;
    meta[3] = m.compressed_height;
    meta[4] = m.compressed_width;
    meta[5] = m.padding_bytes_per_row;
    meta[6] = static_cast<unsigned>(m.compressor_name.size());
  }
  return static_cast<long>(image.GetDataSize());
}

}  // namespace

extern "C" {

long icref_dxt(int format, unsigned h, unsigned w, int pad_mode, unsigned ph, unsigned pw, unsigned padding,
               const unsigned char *src, unsigned char *dst, size_t dst_cap, unsigned *meta) {
  DxtcCompressor c;
  return run(&c, format, h, w, pad_mode, ph, pw, padding, src, dst, dst_cap, meta);
}

long icref_etc(int strategy, int format, unsigned h, unsigned w, int pad_mode, unsigned ph, unsigned pw,
               unsigned padding, const unsigned char *src, unsigned char *dst, size_t dst_cap, unsigned *meta) {
  EtcCompressor c;
  c.SetCompressionStrategy(static_cast<EtcCompressor::CompressionStrategy>(strategy));
  return run(&c, format, h, w, pad_mode, ph, pw, padding, src, dst, dst_cap, meta);
}

long icref_pvrtc(int format, unsigned h, unsigned w, unsigned padding, const unsigned char *src, unsigned char *dst,
                 size_t dst_cap, unsigned *meta) {
  PvrtcCompressor c;
  return run(&c, format, h, w, 0, 0, 0, padding, src, dst, dst_cap, meta);
}

// Compress straight into external storage (public/compressed_image.h:94-100): used for the multi-threaded
// row-stripe CPU baseline and to test the size-mismatch error path.
int icref_dxt_external(int format, unsigned h, unsigned w, unsigned padding, const unsigned char *src,
                       unsigned char *dst, size_t dst_size) {
  DxtcCompressor c;
  CompressedImage image(dst_size, dst);
  return c.Compress(static_cast<CompressedImage::Format>(format), h, w, padding, src, &image) ? 1 : 0;
}

int icref_etc_external(int strategy, unsigned h, unsigned w, unsigned padding, const unsigned char *src,
                       unsigned char *dst, size_t dst_size) {
  EtcCompressor c;
  c.SetCompressionStrategy(static_cast<EtcCompressor::CompressionStrategy>(strategy));
  CompressedImage image(dst_size, dst);
  return c.Compress(CompressedImage::kRGB, h, w, padding, src, &image) ? 1 : 0;
}

// Decompress a block stream through the reference's Decompress().  The CompressedImage is produced by compressing
// a dummy image of the right shape into external storage (which sets the metadata) and then overwriting the
// storage with `blocks`.  Returns bytes written to dst, 0 on failure.
long icref_decompress(int codec, int strategy, int format, unsigned h, unsigned w, const unsigned char *blocks,
                      size_t nbytes, unsigned char *dst, size_t dst_cap) {
  std::vector<unsigned char> storage(nbytes), dummy(static_cast<size_t>(h) * w * 4, 0);
  CompressedImage image(nbytes, storage.data());
  DxtcCompressor dxt;
  EtcCompressor etc;
  etc.SetCompressionStrategy(static_cast<EtcCompressor::CompressionStrategy>(strategy));
  Compressor *c = codec == 2 ? static_cast<Compressor *>(&etc) : static_cast<Compressor *>(&dxt);
  if (!c->Compress(static_cast<CompressedImage::Format>(format), h, w, 0, dummy.data(), &image)) return 0;
  std::memcpy(storage.data(), blocks, nbytes);
  std::vector<uint8> out;
  if (!c->Decompress(image, &out)) return 0;
  if (out.size() > dst_cap) return -1;
  std::memcpy(dst, out.data(), out.size());
  return static_cast<long>(out.size());
}

namespace {
// Wraps a block stream in a CompressedImage with the metadata Compress() would have produced for an h x w image.
struct Wrapped {
  std::vector<unsigned char> storage;
  CompressedImage image;
  Wrapped(Compressor *c, int format, unsigned h, unsigned w, const unsigned char *blocks, size_t nbytes)
      : storage(nbytes), image(nbytes, storage.data()) {
    std::vector<unsigned char> dummy(static_cast<size_t>(h) * w * 4, 0);
    ok = c->Compress(static_cast<CompressedImage::Format>(format), h, w, 0, dummy.data(), &image);
    if (ok) std::memcpy(storage.data(), blocks, nbytes);
  }
  bool ok;
};
long emit(const CompressedImage &img, unsigned char *dst, size_t dst_cap, unsigned *meta) {
  if (img.GetDataSize() > dst_cap) return -1;
  std::memcpy(dst, img.GetData(), img.GetDataSize());
  if (meta) {
    const CompressedImage::Metadata &m = img.GetMetadata();
    meta[0] = m.format; meta[1] = m.uncompressed_height; meta[2] = m.uncompressed_width;
    meta[3] = m.compressed_height; meta[4] = m.compressed_width; meta[5] = m.padding_bytes_per_row;
    meta[6] = static_cast<unsigned>(m.compressor_name.size());
  }
  return static_cast<long>(img.GetDataSize());
}
Compressor *pick(int codec, int strategy, DxtcCompressor *d, EtcCompressor *e) {
  e->SetCompressionStrategy(static_cast<EtcCompressor::CompressionStrategy>(strategy));
  return codec == 2 ? static_cast<Compressor *>(e) : static_cast<Compressor *>(d);
}
}  // namespace

// op: 0 Downsample, 1 Pad(a, b = padded height, width), 2 CopySubimage(a,b,c,d = row, col, height, width)
long icref_block_op(int op, int codec, int strategy, int format, unsigned h, unsigned w, unsigned a, unsigned b,
                    unsigned c, unsigned d, const unsigned char *blocks, size_t nbytes, unsigned char *dst,
                    size_t dst_cap, unsigned *meta) {
  DxtcCompressor dxt;
  EtcCompressor etc;
  Compressor *comp = pick(codec, strategy, &dxt, &etc);
  Wrapped in(comp, format, h, w, blocks, nbytes);
  if (!in.ok) return 0;
  CompressedImage out;
  bool ok = false;
  if (op == 0) ok = comp->Downsample(in.image, &out);
  if (op == 1) ok = comp->Pad(in.image, a, b, &out);
  if (op == 2) ok = comp->CopySubimage(in.image, a, b, c, d, &out);
  if (!ok) return 0;
  return emit(out, dst, dst_cap, meta);
}

long icref_solid(int codec, int format, unsigned h, unsigned w, const unsigned char *color, unsigned char *dst,
                 size_t dst_cap, unsigned *meta) {
  DxtcCompressor dxt;
  EtcCompressor etc;
  Compressor *comp = pick(codec, 2, &dxt, &etc);
  CompressedImage out;
  if (!comp->CreateSolidImage(static_cast<CompressedImage::Format>(format), h, w, color, &out)) return 0;
  return emit(out, dst, dst_cap, meta);
}

void icref_transcode(unsigned char *blocks, size_t nbytes) {
  CompressedImage image(nbytes, blocks);
  image_codec_compression::TranscodeDxt1ToEtc1(&image);
}

size_t icref_size(int codec, int format, unsigned h, unsigned w) {
  if (codec == 0) return DxtcCompressor().ComputeCompressedDataSize(static_cast<CompressedImage::Format>(format), h, w);
  if (codec == 1) return EtcCompressor().ComputeCompressedDataSize(static_cast<CompressedImage::Format>(format), h, w);
  return PvrtcCompressor().ComputeCompressedDataSize(static_cast<CompressedImage::Format>(format), h, w);
}

}  // extern "C"
