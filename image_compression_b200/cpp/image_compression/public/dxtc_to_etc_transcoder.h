// In-place DXT1 -> ETC1 transcode (reference: image_compression/public/dxtc_to_etc_transcoder.h:24).
// Runs on the GPU (icb_transcode_dxt1_to_etc1: DXT1 decode -> ETC1 heuristic encode, one thread per block).  Same
// signature as the reference (void).  The reference cannot fail; this build can (no CUDA device): the image is left
// untouched then and icb_last_error() (include/icb200.h) has the reason.  TranscodeDxt1ToEtc1Checked is an extension
// of this build that reports it.
#ifndef IMAGE_COMPRESSION_PUBLIC_DXTC_TO_ETC_TRANSCODER_H_
#define IMAGE_COMPRESSION_PUBLIC_DXTC_TO_ETC_TRANSCODER_H_

#include "image_compression/public/compressed_image.h"

namespace image_codec_compression {

void TranscodeDxt1ToEtc1(CompressedImage *image);
bool TranscodeDxt1ToEtc1Checked(CompressedImage *image);

}  // namespace image_codec_compression

#endif  // IMAGE_COMPRESSION_PUBLIC_DXTC_TO_ETC_TRANSCODER_H_
