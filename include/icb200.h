/*
 * icb200.h -- C ABI of the B200 (sm_100a) texture-block encoder: the drop-in boundary for the compress path of
 * google/image-compression.
 *
 * What each entry point replaces (paths relative to /root/reference/image_compression/):
 *
 *   icb_dxt1_encode_rgb8 / icb_dxt5_encode_rgba8
 *       the block loop + encoders behind DxtcCompressor::Compress / CompressAndPad
 *       (internal/dxtc_compressor.cc:735-750, 799-818 -> internal/compressor4x4_helper.h:175-216, 479-520)
 *   icb_dxt1_encode_rgba8
 *       same DXT1 encoder fed from a 4-byte-per-pixel source with alpha ignored -- an extension needed by the
 *       headline metric ("DXT1 of RGBA8"); the reference can only reach it by stripping alpha first
 *       
 *   icb_etc1_encode_rgb8
 *       EtcCompressor::Compress / CompressAndPad (internal/etc_compressor.cc:747-758, 787-800)
 *   icb_pvrtc2_encode_rgba8 / icb_pvrtc2_encode_stripe
 *       CompressPVRTC_RGBA_2BPP behind PvrtcCompressor::Compress (internal/pvrtc_compressor.cc:586-597, 636-667)
 *   icb_compress_host
 *       what XxxCompressor::Compress hands its buffer to: host pixels in, host blocks out
 *   icb_compressed_size
 *       ComputeCompressedDataSize (dxtc_compressor.cc:725-733, etc_compressor.cc:734-745, pvrtc_compressor.cc:631-634)
 *
 * Conventions: plain pointers and sizes only; every function returns ICB_OK (0) or a negative icb_status and
 * never throws; (height, width) argument order as in the reference API; device pointers are ordinary CUDA device
 * pointers on the current device; `stream` is a cudaStream_t passed as void* (NULL = default stream).  Device
 * entry points are asynchronous with respect to the host.  There is no CPU fallback: without a usable CUDA
 * device every compute entry point fails with ICB_ERR_CUDA.
 */
#ifndef ICB200_H_
#define ICB200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define ICB_API __attribute__((visibility("default")))
#else
#define ICB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define ICB_ABI_VERSION 2

typedef enum icb_status {
  ICB_OK = 0,
  ICB_ERR_INVALID = -1,     /* null pointer, zero dimension, unsupported format for the codec ...          */
  ICB_ERR_CUDA = -2,        /* a CUDA call failed; icb_last_error() has the text                            */
  ICB_ERR_SIZE = -3,        /* destination buffer size does not match the required size                     */
  ICB_ERR_UNSUPPORTED = -4  /* shape the codec rejects (PVRTC: non power-of-two / non-square / padded rows)  */
} icb_status;

typedef enum icb_codec { ICB_CODEC_DXT1 = 0, ICB_CODEC_DXT5 = 1, ICB_CODEC_ETC1 = 2, ICB_CODEC_PVRTC2 = 3 } icb_codec;

/* Same numbering as CompressedImage::Format (public/compressed_image.h:35-40). */
typedef enum icb_format { ICB_RGB = 0, ICB_BGR = 1, ICB_RGBA = 2, ICB_BGRA = 3 } icb_format;

/* Same numbering as EtcCompressor::CompressionStrategy (public/etc_compressor.h:57-62). */
typedef enum icb_etc_strategy {
  ICB_ETC_SPLIT_HORIZONTALLY = 0,
  ICB_ETC_SPLIT_VERTICALLY = 1,
  ICB_ETC_SMALLER_ERROR = 2,
  ICB_ETC_HEURISTIC = 3
} icb_etc_strategy;

ICB_API int icb_abi_version(void);
/* Thread-local description of the last failure on the calling thread ("" if none). */
ICB_API const char *icb_last_error(void);
/* Number of visible CUDA devices, or a negative icb_status. */
ICB_API int icb_device_count(void);

/* Bytes of packed blocks for an image whose block grid covers coded_height x coded_width pixels. */
ICB_API size_t icb_compressed_size(int codec, uint32_t coded_height, uint32_t coded_width);

/*
 * Device-resident encoders.  d_src: pixel (0,0) of a height x width image with src_pitch bytes per row.
 * The block grid covers coded_height x coded_width pixels (>= height, width; equal for Compress, the padded size
 * for CompressAndPad); windows beyond the image replicate its last row / column.  d_dst receives
 * ceil(coded_height/4) * ceil(coded_width/4) blocks in raster order.  swap_rb selects kBGR / kBGRA.
 * Alignment: d_dst (here and in every entry point that writes blocks) must be aligned to the block size, 8 bytes
 * (DXT1, ETC1, PVRTC) or 16 (DXT5); block streams that are READ (decoders, block operations) to 8 bytes;
 * ICB_ERR_INVALID otherwise.  d_src may have any alignment and pitch (16-byte aligned base and pitch take the TMA path).
 */
ICB_API int icb_dxt1_encode_rgb8(const void *d_src, uint32_t height, uint32_t width, size_t src_pitch, uint32_t coded_height,
                         uint32_t coded_width, int swap_rb, void *d_dst, void *stream);
ICB_API int icb_dxt1_encode_rgba8(const void *d_src, uint32_t height, uint32_t width, size_t src_pitch, uint32_t coded_height,
                          uint32_t coded_width, int swap_rb, void *d_dst, void *stream);
ICB_API int icb_dxt5_encode_rgba8(const void *d_src, uint32_t height, uint32_t width, size_t src_pitch, uint32_t coded_height,
                          uint32_t coded_width, int swap_rb, void *d_dst, void *stream);
ICB_API int icb_etc1_encode_rgb8(const void *d_src, uint32_t height, uint32_t width, size_t src_pitch, uint32_t coded_height,
                         uint32_t coded_width, int strategy, void *d_dst, void *stream);

/*
 * Row-stripe form of the three encoders above, for sharding one image over several GPUs: encodes only block rows
 * [block_row_begin, block_row_end) of the grid.  d_src still points at pixel (0,0) of the WHOLE image (rows the
 * stripe does not touch need not be resident); d_dst points at the first block of the stripe.
 */
ICB_API int icb_encode4x4_stripe(int codec, int format_components, const void *d_src, uint32_t height, uint32_t width,
                         size_t src_pitch, uint32_t coded_height, uint32_t coded_width, int swap_rb, int etc_strategy,
                         uint32_t block_row_begin, uint32_t block_row_end, void *d_dst, void *stream);

/*
 * PVRTC1 2 bpp.  width == height, power of two, >= 8; rows contiguous.  d_dst receives width*height/4 bytes
 * (blocks in Z-order).  d_scratch: icb_pvrtc2_scratch_size() bytes of device memory, or NULL to let the library
 * allocate it stream-ordered.
 */
ICB_API size_t icb_pvrtc2_scratch_size(uint32_t height, uint32_t width);
ICB_API int icb_pvrtc2_encode_rgba8(const void *d_src, uint32_t height, uint32_t width, void *d_dst, void *d_scratch,
                            void *stream);
/*
 * Row-stripe form for sharding one image over several GPUs: encodes block rows
 * [block_row_begin, block_row_end) (8x4 blocks: height/4 block rows).  PVRTC is not block-local -- a pixel blends the
 * A/B colours of the neighbouring blocks, toroidally -- so d_rows holds the stripe's pixel rows PLUS one block row (4
 * pixel rows) of halo above and below: image rows 4*(begin-1) .. 4*(end+1)-1, wrapped modulo height, in that order,
 * contiguous, width*4 bytes each.  The halo blocks' colours are recomputed locally; nothing is exchanged between
 * ranks.  d_first_pixel points at a device copy of image pixel (0,0) (4 bytes), which the reference's extreme search
 * can pick in any block (quirk P1, internal/pvrtc_compressor.cc:255-329).  d_dst is the WHOLE image's block buffer
 * (width*height/4 bytes, Z-order): the stripe's blocks are stored at their final positions, so with a peer-mapped
 * d_dst (icb_ipc_open) the ranks assemble the image without a gather.  d_scratch as above (same size), or NULL.
 * The stripe plus its halo must not exceed the image: end - begin + 2 <= height/4.
 */
ICB_API int icb_pvrtc2_encode_stripe(const void *d_rows, const void *d_first_pixel, uint32_t height, uint32_t width,
                             uint32_t block_row_begin, uint32_t block_row_end, void *d_dst, void *d_scratch, void *stream);

/*
 * Host-buffer path: validates like the reference's Compress / CompressAndPad, stages src to the device in
 * chunks overlapped with the kernels, and copies the blocks back.  padded_height / padded_width: 0 for Compress.
 * dst_size must equal the required size exactly (ICB_ERR_SIZE otherwise), mirroring SetUpCompressedImage
 * (internal/compressor4x4_helper.cc:22-43).  Blocking.  Pinned src/dst (icb_host_alloc) are DMA'd directly; ordinary
 * pageable memory is staged through pinned buffers by a small pool of copy threads (ICB_STAGING_THREADS=n overrides
 * their number, 0 leaves pageable copies to the CUDA driver).  ICB_HOST_DEVICES=N|all spreads one call over N GPUs of
 * the box (the current device and the next ordinals): chunk c is uploaded to, encoded on and downloaded from device
 * c mod N, each over its own PCIe link, and every device writes its blocks to their place in dst -- no gather.  The
 * caller's current device is unchanged on return.
 */
ICB_API int icb_compress_host(int codec, int format, uint32_t height, uint32_t width, uint32_t padded_height,
                      uint32_t padded_width, uint32_t padding_bytes_per_row, int etc_strategy, const void *src,
                      void *dst, size_t dst_size);

/*
 * How many GPUs icb_compress_host spreads one call over, set by program instead of by ICB_HOST_DEVICES (what a C++
 * caller of Compress() uses: the classes call icb_compress_host).  n >= 1: that many (the current device and the next
 * ordinals); 0: all visible devices; < 0: back to the environment variable / one device.  Process-wide.  Returns the
 * previous setting (0x7fffffff if there was none) or a negative icb_status.
 */
ICB_API int icb_set_host_devices(int n);

/*
 * Host-path resources.  The host-buffer entry points (icb_compress_host, icb_decompress_host, icb_blockop_host) work
 * through "pipes" -- three streams, events, device buffers for the image and its blocks, pinned staging rings -- that
 * are leased from a process-wide pool for the duration of one call and handed back afterwards; concurrent calls get a
 * pipe each, and a host thread that exits leaves nothing behind.  A pooled pipe's buffers only ever grow (a pipe that
 * has encoded an 8192x8192 RGBA8 image keeps ~300 MB of device memory plus up to 96 MB of pinned memory), so the pool
 * holds `peak concurrent calls` pipes per device until icb_trim() frees the idle ones (and trims the library's
 * stream-ordered scratch pools).  icb_trim returns the bytes it released; icb_idle_pipes the number of pooled pipes.
 */
ICB_API size_t icb_trim(void);
ICB_API size_t icb_idle_pipes(void);

/*
 * One image over several GPUs from ONE process -- the C-ABI form of the row-stripe
 * sharding that bench.py / image_compression_b200.sharding do with one process per GPU.  Replaces the same loop as
 * the single-GPU encoders, Compressor4x4Helper::Compress (internal/compressor4x4_helper.h:202-214), cut into stripes
 * of whole block rows.
 *
 * icb_ctx_create        n_dev devices (dev_ids NULL = ordinals 0..n_dev-1; n_dev 0 = all visible).  dev_ids[0] is the
 *                       ROOT: the device that ends up holding the packed block stream.  Creates one stream per
 *                       non-root device and enables peer access from every device to the root
 *                       (cudaDeviceEnablePeerAccess -- one process, so no IPC handles are needed).
 * icb_stripe_partition  block rows of stripe r = [splits[r], splits[r+1]), n+1 entries.  root_share_permille < 0: even
 *                       split.  Otherwise the root's stripe is at least that share of the rows and the rest is split
 *                       evenly: with the stream delivered to the root, the job is bounded by the root's NVLink ingress
 *                       from N = 4 up, and the fastest partition lets the root encode while the others' blocks arrive
 *                       (icb_root_share_permille gives the share that balances the two on B200; it returns -1 where
 *                       an even split is already compute-bound, e.g. ETC1).  Stripes are multiples of 4 block rows.
 * icb_encode_sharded    d_src_stripes[r]: device pointer, resident on device r of the context, to the first pixel row
 *                       (row 4*splits[r]) of stripe r, src_pitch bytes per row.  d_dst: on the root, the WHOLE
 *                       stream (icb_compressed_size bytes).  Every device encodes its stripe and stores the blocks
 *                       straight into d_dst -- peer stores over NVLink, no gather pass.  `stream` is a stream of the
 *                       root device: the other devices' work is ordered after what it holds (event fork) and it
 *                       waits for all of them (event join), so the call is asynchronous and stream-ordered like the
 *                       single-GPU encoders.  The caller's current device is unchanged on return.
 * icb_ctx_compress_host icb_compress_host over exactly the context's devices (each uploads, encodes and downloads its
 *                       own chunks over its own PCIe link; nothing is gathered).
 * Threading: icb_encode_sharded orders its devices through events owned by the context, so calls on ONE context must
 * not overlap (use one context per calling thread; contexts are cheap); icb_ctx_compress_host leases its streams from
 * the process-wide pool and may be called concurrently.
 */
typedef struct icb_ctx icb_ctx;
ICB_API int icb_ctx_create(int n_dev, const int *dev_ids, icb_ctx **ctx);
ICB_API int icb_ctx_destroy(icb_ctx *ctx);
ICB_API int icb_ctx_device_count(const icb_ctx *ctx);
ICB_API int icb_ctx_device(const icb_ctx *ctx, int k);
ICB_API int icb_ctx_peer_stores(const icb_ctx *ctx);
ICB_API int icb_stripe_partition(int n, uint32_t grid_rows, int root_share_permille, uint32_t *splits);
ICB_API int icb_root_share_permille(int codec, int format_components, int n);
ICB_API int icb_encode_sharded(icb_ctx *ctx, int codec, int format, uint32_t height, uint32_t width, size_t src_pitch,
                               int etc_strategy, const uint32_t *splits, const void *const *d_src_stripes, void *d_dst,
                               void *stream);
ICB_API int icb_ctx_compress_host(icb_ctx *ctx, int codec, int format, uint32_t height, uint32_t width,
                                  uint32_t padded_height, uint32_t padded_width, uint32_t padding_bytes_per_row,
                                  int etc_strategy, const void *src, void *dst, size_t dst_size);

/*
 * Block decoders (DXT1 -> RGB888, DXT5 -> RGBA8888, ETC1 -> RGB888), the step after the compress path:
 * Compressor4x4Helper::Decompress (internal/compressor4x4_helper.h:218-262).  d_blocks holds
 * ceil(height/4) * block_cols blocks in raster order (the reference walks ceil(uncompressed_width/4) blocks per
 * row; pass that); pixels outside height x width are dropped.  d_dst rows are dst_pitch bytes apart.
 */
ICB_API int icb_decode4x4(int codec, const void *d_blocks, uint32_t height, uint32_t width, uint32_t block_cols,
                          int swap_rb, void *d_dst, size_t dst_pitch, void *stream);
/* Host-buffer form: blocks in, height*width*(3|4) bytes of pixels out (rows contiguous).  Blocking. */
ICB_API int icb_decompress_host(int codec, int format, uint32_t height, uint32_t width, uint32_t block_cols,
                                const void *blocks, size_t blocks_size, void *dst, size_t dst_size);

/*
 * Compressed-domain operations on DXT1 / DXT5 / ETC1 block streams (codec = ICB_CODEC_DXT1 | DXT5 | ETC1), the
 * callers either side of the compress path.  Device-resident and asynchronous like
 * the encoders; a stream of blocks that lives in HBM gets its mip chain, padding and ETC1 twin without visiting
 * the host.  etc_strategy is only read for ICB_CODEC_ETC1 (the re-encode step).
 *
 * icb_downsample4x4      Compressor4x4Helper::Downsample (internal/compressor4x4_helper.h:264-391, 594-636):
 *                        d_blocks holds ceil(height/4) * ceil(width/4) blocks of a height x width image; d_dst
 *                        receives the blocks of the ceil(height/2) x ceil(width/2) image.  ICB_ERR_UNSUPPORTED where
 *                        the reference returns false (an odd block count > 1 in a dimension; a 3-pixel dimension of a
 *                        single-block image).
 * icb_pad4x4             Compressor4x4Helper::Pad (:393-477) with the codecs' pad blocks
 *                        (internal/dxtc_compressor.cc:594-696, internal/etc_compressor.cc:645-698).  compressed_* is
 *                        the size of the input grid in pixels (a multiple of 4), padded_* the size asked for; d_dst
 *                        receives ceil(padded_height/4) * ceil(padded_width/4) blocks.  When both padded sizes are <=
 *                        the compressed ones the blocks are copied unchanged (the reference's Duplicate).  The
 *                        reference overruns its output when one padded dimension has FEWER blocks than the input and
 *                        the other more; that case is refused with ICB_ERR_UNSUPPORTED.
 * icb_copy_subimage4x4   Compressor4x4Helper::CopySubimage (:547-592): all four values multiples of 4 and inside the
 *                        compressed_height x compressed_width grid, else ICB_ERR_INVALID.
 * icb_fill_solid4x4      CreateSolidImage (internal/dxtc_compressor.cc:820-840, internal/etc_compressor.cc:595-617,
 *                        772-785): colour = 3 (DXT1, ETC1) or 4 (DXT5) bytes in the order the caller would pass them
 *                        to the reference; fills ceil(height/4) * ceil(width/4) identical blocks.
 * icb_transcode_dxt1_to_etc1  TranscodeDxt1ToEtc1 (internal/dxtc_to_etc_transcoder.cc:29-40): in place, num_blocks
 *                        8-byte blocks, ETC1 heuristic strategy.
 */
ICB_API int icb_downsample4x4(int codec, int etc_strategy, const void *d_blocks, uint32_t height, uint32_t width,
                              void *d_dst, void *stream);
ICB_API int icb_pad4x4(int codec, int etc_strategy, const void *d_blocks, uint32_t compressed_height,
                       uint32_t compressed_width, uint32_t padded_height, uint32_t padded_width, void *d_dst,
                       void *stream);
ICB_API int icb_copy_subimage4x4(int codec, const void *d_blocks, uint32_t compressed_height, uint32_t compressed_width,
                                 uint32_t start_row, uint32_t start_column, uint32_t height, uint32_t width, void *d_dst,
                                 void *stream);
ICB_API int icb_fill_solid4x4(int codec, const uint8_t *colour, uint32_t height, uint32_t width, void *d_dst,
                              void *stream);
ICB_API int icb_transcode_dxt1_to_etc1(void *d_blocks, size_t num_blocks, void *stream);

/*
 * Host-buffer forms of the operations above (what Compressor::Downsample / Pad / CreateSolidImage and
 * TranscodeDxt1ToEtc1 hand their buffers to).  Blocking.  dst_size must equal the size of the result exactly.
 */
typedef enum icb_block_op { ICB_OP_DOWNSAMPLE = 0, ICB_OP_PAD = 1, ICB_OP_COPY_SUBIMAGE = 2, ICB_OP_SOLID = 3, ICB_OP_TRANSCODE = 4 } icb_block_op;
/*
 * args by op:  DOWNSAMPLE {height, width}            PAD {compressed_height, compressed_width, padded_height, padded_width}
 *              COPY_SUBIMAGE {compressed_height, compressed_width, start_row, start_column, height, width}
 *              SOLID {height, width, colour bytes packed little-endian}          TRANSCODE {} (src may equal dst)
 */
ICB_API int icb_blockop_host(int op, int codec, int etc_strategy, const uint32_t *args, const void *src, size_t src_size,
                             void *dst, size_t dst_size);

/*
 * Peer-mapped output for the multi-GPU stripe path ("B200-native alternative" to the NCCL gather
 * of the packed block stream): the rank that wants the whole stream allocates it with icb_device_alloc and exports
 * it; every other rank (one process per GPU) opens the handle and passes `mapped + byte offset of its stripe` as d_dst
 * of icb_encode4x4_stripe.  The encoder's block stores then travel over NVLink straight into the owner's buffer --
 * the gather is fused into the kernel's epilogue and overlaps the encode -- and the stream is complete on the owner
 * once every rank's stream has been synchronised.  icb_device_alloc is a plain cudaMalloc (pool sub-allocations
 * cannot be exported); the handle is ICB_IPC_HANDLE_BYTES opaque bytes to ship through any channel.
 */
#define ICB_IPC_HANDLE_BYTES 64
ICB_API int icb_device_alloc(size_t bytes, void **d_ptr);
ICB_API int icb_device_free(void *d_ptr);
ICB_API int icb_ipc_export(const void *d_ptr, void *handle);
ICB_API int icb_ipc_open(const void *handle, void **d_ptr);
ICB_API int icb_ipc_close(void *d_ptr);

/* Page-locked host memory for icb_compress_host callers.  icb_host_register page-locks a buffer the caller already owns
 * (e.g. the std::vector a C++ caller of Compress() keeps its frames in) for as long as it stays registered, which puts
 * every later call that reads or writes it on the pinned path (8192x8192 RGBA8 -> DXT1: 5.0 ms instead of 6.9 ms through
 * the staging copies); registration itself costs about a millisecond per 100 MB, so it pays for buffers that are reused. */
ICB_API void *icb_host_alloc(size_t bytes);
ICB_API void icb_host_free(void *p);
ICB_API int icb_host_register(void *p, size_t bytes);
ICB_API int icb_host_unregister(void *p);

/*
 * Synthetic input stream S(seed): fills d_dst[0..bytes) with bytes byte_offset.. of the
 * stream, so any stripe can be generated in place on its GPU.
 */
ICB_API int icb_fill_synthetic(void *d_dst, size_t bytes, uint64_t seed, uint64_t byte_offset, void *stream);

/* Number of kernels this library has launched in this process (bench.py reports it as gpu_launches). */
ICB_API uint64_t icb_launch_count(void);

/* Force (1) or forbid (0) the TMA fast path for testing; -1 restores automatic selection.  Returns previous. */
ICB_API int icb_set_tma_mode(int mode);

#ifdef __cplusplus
}
#endif
#endif /* ICB200_H_ */
