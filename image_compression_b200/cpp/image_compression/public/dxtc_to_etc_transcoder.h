// In-place DXT1 -> ETC1 transcode (reference: image_compression/public/dxtc_to_etc_transcoder.h:24).
// Not on the GPU compress path; declared for source compatibility.  See DESIGN.md "out of scope this round".
#ifndef IMAGE_COMPRESSION_PUBLIC_DXTC_TO_ETC_TRANSCODER_H_
#define IMAGE_COMPRESSION_PUBLIC_DXTC_TO_ETC_TRANSCODER_H_

#include "image_compression/public/compressed_image.h"

namespace image_codec_compression {

bool TranscodeDxt1ToEtc1(CompressedImage *image);

}  // namespace image_codec_compression

#endif  // IMAGE_COMPRESSION_PUBLIC_DXTC_TO_ETC_TRANSCODER_H_
