// In-place DXT1 -> ETC1 transcode (reference: image_compression/public/dxtc_to_etc_transcoder.h:24).
// Runs on the GPU (icb_transcode_dxt1_to_etc1: DXT1 decode -> ETC1 heuristic encode, one thread per block).  The
// reference returns void; this build returns false if the CUDA call failed (the image is untouched then).  The
// mangled name does not include the return type, so callers compiled against the reference header still link.
#ifndef IMAGE_COMPRESSION_PUBLIC_DXTC_TO_ETC_TRANSCODER_H_
#define IMAGE_COMPRESSION_PUBLIC_DXTC_TO_ETC_TRANSCODER_H_

#include "image_compression/public/compressed_image.h"

namespace image_codec_compression {

bool TranscodeDxt1ToEtc1(CompressedImage *image);

}  // namespace image_codec_compression

#endif  // IMAGE_COMPRESSION_PUBLIC_DXTC_TO_ETC_TRANSCODER_H_
