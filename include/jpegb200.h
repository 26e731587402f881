/*
 * jpegb200.h -- C ABI of libjpegb200.so, the B200 (sm_100a) implementation of the
 * natural_c grayscale JPEG encode hot path.
 *
 * Plain C: pointers and sizes only, no torch / CUDA types in any signature (a CUDA
 * stream is passed as void*).  Three groups of entry points:
 *
 *   1. The reference's own core-stage API, same names, same struct layouts, same
 *      ownership and NULL-on-error behaviour.  A maintainer drops
 *      natural_c/src/core/ *.c from the build, links -ljpegb200, and
 *      io/jpeg_handler.c:saveJPEGGrayscale keeps working unmodified.  Host pointers
 *      in, host pointers out; every stage runs as a CUDA kernel (there is no CPU
 *      fallback: without a usable GPU every call returns NULL).
 *   2. The fused entry the shimmed orchestrator calls (BMPImage -> scan bytes in one
 *      pass over device memory, no per-stage intermediates on the host).
 *   3. A device-resident encoder handle for batches and for MCU-row stripes of one
 *      huge image (multi-GPU), used by the bench / Python host layer.
 *
 * Reference citations are relative to natural_c/ of
 * strbac-damjan/jpeg-image-compression.
 */
#ifndef JPEGB200_H
#define JPEGB200_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------- */
/* 1. reference-compatible types (x86-64 layouts verified with sizeof/offsetof) */

#ifndef JPEGB200_NO_REFERENCE_TYPES
typedef struct BMPImage {            /* include/bmp_handler.h:37-41 (16 B) */
    int32_t  width;
    int32_t  height;
    uint8_t *data;                   /* top-down interleaved RGB, pitch 3*width */
} BMPImage;

typedef struct {                     /* include/converter.h:21-26 */
    int      width;                  /* padded to a multiple of 8 */
    int      height;
    uint8_t *data;
} YImage;

typedef struct {                     /* include/converter.h:13-18 */
    int     width;
    int     height;
    int8_t *data;
} CenteredYImage;

typedef struct {                     /* include/dct.h:14-19 */
    int    width;
    int    height;
    float *coefficients;             /* raster image layout */
} DCTImage;

typedef struct {                     /* include/quantization.h:10-14 */
    int      width;
    int      height;
    int16_t *data;
} QuantizedImage;

typedef struct {                     /* include/zigzag.h:7-12 (24 B) */
    int      numBlocksW;
    int      numBlocksH;
    int      totalBlocks;
    int16_t *data;                   /* [totalBlocks][64] */
} ZigZagData;

typedef struct {                     /* include/rle.h:8-14 (6 B: u8 @0, u16 @2, u8 @4) */
    uint8_t  symbol;
    uint16_t code;
    uint8_t  codeBits;
} RLESymbol;

typedef struct {                     /* include/rle.h:17-21 (24 B) */
    RLESymbol *data;
    size_t     count;
    size_t     capacity;
} RLEData;

typedef struct {                     /* include/huffman.h:9-13 (24 B) */
    uint8_t *data;
    size_t   size;
    size_t   capacity;
} JpegEncoderBuffer;
#endif

/* constant tables, value-identical to src/core/jpeg_tables.c:3-48
 * (declared in include/jpeg_tables.h:7-15) */
extern const unsigned char std_luminance_quant_tbl[64];
extern const unsigned char std_dc_luminance_nrcodes[16];
extern const unsigned char std_dc_luminance_values[12];
extern const unsigned char std_ac_luminance_nrcodes[16];
extern const unsigned char std_ac_luminance_values[162];

/* core stages -- each replaces the reference function of the same name.
 * NULL in (or ->data == NULL) => NULL out; results are malloc'd host structs that
 * the caller releases with the matching free*. */
YImage            *convertBMPToJPEGGrayscale(const BMPImage *image);          /* src/core/converter.c:4   (include/converter.h:28) */
CenteredYImage    *centerYImage(const YImage *source);                        /* src/core/converter.c:60  (include/converter.h:30) */
void               freeCenteredYImage(CenteredYImage *img);                   /* src/core/converter.c:92  */
void               computeDCTBlock(const int8_t inputBlock[8][8],
                                   float outputBlock[8][8]);                  /* src/core/dct.c:63        (include/dct.h:22) */
DCTImage          *performDCT(const CenteredYImage *image);                   /* src/core/dct.c:98        (include/dct.h:23) */
void               freeDCTImage(DCTImage *img);                               /* src/core/dct.c:153       */
QuantizedImage    *quantizeImage(const DCTImage *dctImg);                     /* src/core/quantization.c:3  (include/quantization.h:16) */
void               freeQuantizedImage(QuantizedImage *img);                   /* src/core/quantization.c:45 */
ZigZagData        *performZigZag(const QuantizedImage *qImg);                 /* src/core/zigzag.c:21     (include/zigzag.h:14) */
void               freeZigZagData(ZigZagData *zData);                         /* src/core/zigzag.c:70     */
RLEData           *performRLE(const ZigZagData *zigZagData);                  /* src/core/rle.c:51        (include/rle.h:23) */
void               freeRLEData(RLEData *rleData);                             /* src/core/rle.c:129       */
JpegEncoderBuffer *encodeHuffman(const RLEData *rleData, int totalBlocks);    /* src/core/huffman.c:121   (include/huffman.h:36) */
void               freeJpegEncoderBuffer(JpegEncoderBuffer *buffer);          /* src/core/huffman.c:195   */

/* host-side I/O kept in C (restated from behaviour, not accelerated) */
BMPImage *loadBMPImage(const char *filename);                                 /* src/io/bmp_handler.c:15  */
void      freeBMPImage(BMPImage *image);                                      /* src/io/bmp_handler.c:5   */
bool      saveJPEGGrayscale(const char *filename, const BMPImage *img);       /* src/io/jpeg_handler.c:119 (fused GPU path inside) */
void      freeYImage(YImage *img);                                            /* src/io/jpeg_handler.c:284 */

/* ------------------------------------------------------------------------- */
/* 2. fused entry: everything saveJPEGGrayscale does between fopen and the header
 *    writes (src/io/jpeg_handler.c:133-201) in one device pass.  Returns the stuffed
 *    scan bytes (what jpeg_handler.c:252 fwrites), or NULL on any failure. */
JpegEncoderBuffer *jpegb200_encode_scan(const BMPImage *image);

/* As above, and also returns the first block's 8x8 quantized coefficients in raster
 * order (the orchestrator prints them, src/io/jpeg_handler.c:168-175). */
JpegEncoderBuffer *jpegb200_encode_scan_dbg(const BMPImage *image, int16_t first_block[64]);

/* The 328 header bytes of src/io/jpeg_handler.c:220-233 for an image of w x h. */
size_t jpegb200_jfif_header(int width, int height, uint8_t out[328]);

/* ------------------------------------------------------------------------- */
/* 3. device-resident encoder */

typedef struct jpegb200_encoder jpegb200_encoder;

enum {
    JPEGB200_OK = 0,
    JPEGB200_ERR_CUDA = 1,          /* a CUDA call failed (see jpegb200_last_error) */
    JPEGB200_ERR_ARG = 2,
    JPEGB200_ERR_WORKSPACE = 3,     /* packed-bits workspace too small: raise bytes_per_block */
    JPEGB200_ERR_OUTPUT = 4,        /* caller's scan buffer too small */
    JPEGB200_ERR_INTERNAL = 5
};

/* dct_mode: 0 = transform on the tensor cores (exact integer limb sums of the reference's LUT products) + rigorous
 * guard band + exact re-evaluation of flagged coefficients (default, bit-exact by construction); 1 = every
 * coefficient in the reference's exact summation order (slow, used to cross-check mode 0); 2 = register butterfly
 * transform + guard band + exact re-evaluation (the round-1 transform, same output, kept for comparison). */
int          jpegb200_device_count(void);
const char  *jpegb200_last_error(void);
jpegb200_encoder *jpegb200_encoder_create(int device);
void         jpegb200_encoder_destroy(jpegb200_encoder *enc);
int          jpegb200_encoder_set_dct_mode(jpegb200_encoder *enc, int dct_mode);
/* Launch-shape hint: how many encoder handles the caller keeps busy on this GPU at the same time, each on its own
 * stream (default 1: every kernel is sized for the whole device, lowest latency per image).  With n > 1 each launch
 * is sized for about 1.2 / sqrt(n) of the SMs, so launches of different handles run side by side instead of queueing
 * behind one another's tails: higher aggregate throughput on small images (3840x2160, 8 handles: +18 %), longer
 * latency per image.  Output bytes do not depend on it. */
int          jpegb200_encoder_set_concurrency(jpegb200_encoder *enc, int handles);
/* Debugging aid.  With JPEGB200_GUARD=1 in the environment every workspace buffer of a handle is allocated at exactly
 * the size the launch needs, between two 4 KB guard bands, all filled with 0xA5.  This call synchronises the device and
 * reports how many guard bytes were overwritten and how many buffers carry guards (0 when the variable is not set). */
int          jpegb200_encoder_check_guards(jpegb200_encoder *enc, uint64_t *corrupted_bytes, int *guarded_buffers);
/* Workspace hint: expected packed bytes per 8x8 block, averaged over any 256 consecutive blocks
 * (default 24; up to 32 the entropy kernel runs with its small shared-memory bit windows, above
 * that with the worst-case ones: 184 covers every possible input).  It also sizes the per-image
 * slots of batch mode.  Too small a value is never silent: the encode reports
 * JPEGB200_ERR_WORKSPACE and can be re-run with a larger one. */
int          jpegb200_encoder_set_bytes_per_block(jpegb200_encoder *enc, int bytes_per_block);

/* Description of one launch: `count` images of identical geometry, image i starting
 * at d_rgb + i*image_stride (top-down interleaved RGB, row pitch 3*width bytes).  No alignment
 * is required of d_rgb, the stride or the width.  Pixel rows are fetched in whole 16-byte lines,
 * so up to 15 bytes before the first and after the last pixel byte may be read (never used):
 * they must be readable device memory, which holds for any pointer into a cudaMalloc'd (or
 * pooled) allocation because allocations start and end on at least 256-byte boundaries. */
typedef struct {
    const uint8_t *d_rgb;
    int32_t  width;
    int32_t  height;
    int32_t  count;
    uint64_t image_stride;          /* bytes between consecutive images (>= 3*w*h) */
} jpegb200_batch;

/* Encode a batch resident in device memory.  All arguments are device pointers:
 *   d_scan          stuffed scan bytes of all images, back to back
 *   d_scan_offsets  uint64[count+1], byte offset of image i in d_scan ([count] = total)
 * Asynchronous on `cuda_stream`; call jpegb200_encoder_status after synchronising. */
int jpegb200_encode_batch_device(jpegb200_encoder *enc, const jpegb200_batch *batch,
                                 uint8_t *d_scan, uint64_t scan_capacity,
                                 uint64_t *d_scan_offsets, void *cuda_stream);

/* Same, but the output holds complete JFIF files: for every image the 328-byte header that
 * saveJPEGGrayscale writes (src/io/jpeg_handler.c:220-233), its scan (:252) and the EOI marker (:262),
 * back to back; d_file_offsets is uint64[count+1].  One D2H copy then yields `count` finished .jpg
 * files with no per-file host formatting.  Capacity: scan bytes + 330 per image. */
int jpegb200_encode_batch_files_device(jpegb200_encoder *enc, const jpegb200_batch *batch,
                                       uint8_t *d_files, uint64_t capacity,
                                       uint64_t *d_file_offsets, void *cuda_stream);

/* Device error word (sticky until read; 0 = ok, else a JPEGB200_ERR_*); synchronises the stream. */
int jpegb200_encoder_status(jpegb200_encoder *enc, void *cuda_stream);

/* Counters.  blocks / kernel_launches / packed_bytes describe the last encode call;
 * flagged_coefficients accumulates since the previous jpegb200_encoder_stats call. */
typedef struct {
    uint64_t blocks;                /* 8x8 blocks processed                           */
    uint64_t flagged_coefficients;  /* coefficients re-evaluated in reference order    */
    uint64_t kernel_launches;       /* kernels launched by the last encode call        */
    uint64_t packed_bytes;          /* unstuffed stream bytes (all images)             */
} jpegb200_stats;
int jpegb200_encoder_stats(jpegb200_encoder *enc, jpegb200_stats *out);

/* Per-kernel device time: with profiling on, every kernel launch is bracketed by cudaEvents
 * on the launching stream.  kernel ids: 0 block kernel (K1), 1 merge + stuff (K2), 2 batch layout,
 * 3 batch compaction, 4 file framing, 5 strip entropy kernel (K1b).  Synchronises the device. */
int jpegb200_encoder_set_profiling(jpegb200_encoder *enc, int on);
int jpegb200_encoder_kernel_times(jpegb200_encoder *enc, double ms_total[8], uint64_t calls[8], int reset);

/* Host buffers in, host buffers out (the reference-facing call with an explicit handle):
 * H2D copy of the RGB payload, the three kernels, D2H copy of the stuffed scan (copied speculatively at the size of
 * the previous result, so that the call needs ONE stream synchronisation).  Pinned caller buffers make the copies
 * run at PCIe speed.  Synchronous. */
int jpegb200_encode_host(jpegb200_encoder *enc, const uint8_t *host_rgb, int width, int height,
                         uint8_t *host_scan, uint64_t host_capacity, uint64_t *host_scan_bytes,
                         void *cuda_stream);

/* Stage taps for parity tests (device -> host copies; synchronous):
 * zig-zag coefficients of the last launch widened to int16 [count*blocks][64], and
 * per-block Huffman bit costs uint32[count*blocks]. */
int jpegb200_encoder_read_coefficients(jpegb200_encoder *enc, int16_t *host_zz, uint64_t nblocks);
int jpegb200_encoder_read_block_bits(jpegb200_encoder *enc, uint32_t *host_bits, uint64_t nblocks);

/* BMP file image in host memory -> complete JPEG file in host memory (header + scan + EOI), the
 * opt-in fast path of SURVEY.md section 8f: the BMP's pixel array goes to the device as it is
 * (bottom-up, BGR, 4-byte padded rows) and the block kernel reads it in place, replacing the host
 * loader's per-pixel pass (src/io/bmp_handler.c:103-124).  Same acceptance rules as loadBMPImage
 * (src/io/bmp_handler.c:23-49,68-75,88).  Output bytes equal what saveJPEGGrayscale writes. */
int jpegb200_encode_bmp_to_jpeg_host(jpegb200_encoder *enc, const uint8_t *bmp, uint64_t bmp_bytes,
                                     uint8_t *jpeg_out, uint64_t jpeg_capacity, uint64_t *jpeg_bytes,
                                     int *width, int *height, void *cuda_stream);

/* Tuning aid: with JPEGB200_K2_TRACE=1 in the environment K2 records 8 phase timestamps (ns) per
 * tile; this copies them out ([ntiles][8]). */
int jpegb200_encoder_read_trace(jpegb200_encoder *enc, uint64_t *host, uint64_t ntiles);
/* Same for the block kernel (JPEGB200_K1_TRACE=1): 8 timestamps per persistent warp, followed (pass
 * 2 x the warp count) by the completion times of each warp's first 8 strips. */
int jpegb200_encoder_read_k1_trace(jpegb200_encoder *enc, uint64_t *host, uint64_t nwarps);
/* Grid and warps per CTA of the last block-kernel launch (sizes the trace above). */
int jpegb200_encoder_launch_shape(jpegb200_encoder *enc, int *k1_grid, int *k1_warps_per_cta);

/* ---- MCU-row stripes of one image across several GPUs ----------------------
 * Rank r owns block rows [row0, row0+rows) of an image.  d_rgb points at the stripe's first
 * pixel row; stripe_height = pixel rows of the stripe (a multiple of 8 except for the image's
 * last stripe, which carries the ragged bottom); halo_rows = pixel rows of the NEXT stripe's
 * first block row that follow in the same buffer (min(8, rows left); 0 for the last stripe).
 * The halo lets a rank finish the byte its stream ends in without receiving bits from its
 * neighbour.  As for jpegb200_batch, pixel rows are fetched in whole 16-byte lines: up to 15 bytes before the
 * first and after the last pixel byte of the stripe buffer may be read (never used) and must be readable device
 * memory (true for any pointer into a cudaMalloc'd or pooled allocation).  Two phases with one tiny exchange between them (done by the caller over NCCL;
 * see INTEGRATION.md and jpeg_image_compression_b200/stripes.py):
 *
 *   analyze : fused block kernel over the stripe (+ halo) -> {first_dc, last_dc, bits_pred0}
 *             bits_pred0 = the stripe's bit count if its first block were predicted from DC 0
 *   -- all-gather the summaries: every rank derives each stripe's true predictor (previous
 *      stripe's last_dc), true bit count and therefore its global bit offset --
 *   encode  : scan + pack + stuff at the true global bit phase.  The stripe emits exactly the
 *             stuffed bytes whose first bit lies inside its bit range; the concatenation of all
 *             ranks' bytes is the image's scan, byte-identical to the single-GPU encode.     */
typedef struct {
    int16_t  first_dc;
    int16_t  last_dc;
    uint32_t reserved;              /* device variants: 1 = the rank owns block rows, 0 = empty slot */
    uint64_t bits_pred0;
} jpegb200_stripe_summary;

int jpegb200_stripe_analyze(jpegb200_encoder *enc, const uint8_t *d_rgb, int width, int stripe_height,
                            int halo_rows, jpegb200_stripe_summary *host_out, void *cuda_stream);
int jpegb200_stripe_encode(jpegb200_encoder *enc, int16_t dc_predictor, uint64_t bit_begin,
                           uint8_t *d_scan, uint64_t scan_capacity, uint64_t *host_scan_bytes,
                           void *cuda_stream);

/* Device-resident variants: nothing returns to the host between analyze, exchange and encode, so the three steps
 * can be enqueued back to back on one stream (and captured in a CUDA graph).
 *   analyze_device : as above, the summary is reduced on the device and written to d_out (device memory, 16 bytes)
 *   -- all-gather the summaries into d_all[world] on the device (NCCL); a rank without block rows contributes an
 *      all-zero slot --
 *   encode_device  : a one-warp kernel derives this rank's predictor and global bit offset from d_all (the device
 *                    version of stripes.resolve_offsets), then scan + merge + stuff.  d_scan_info[0] = 0,
 *                    d_scan_info[1] = stuffed bytes this rank produced (device memory). */
int jpegb200_stripe_analyze_device(jpegb200_encoder *enc, const uint8_t *d_rgb, int width, int stripe_height,
                                   int halo_rows, jpegb200_stripe_summary *d_out, void *cuda_stream);
int jpegb200_stripe_encode_device(jpegb200_encoder *enc, const jpegb200_stripe_summary *d_all, int world, int rank,
                                  uint8_t *d_scan, uint64_t scan_capacity, uint64_t *d_scan_info, void *cuda_stream);

/* Synthetic workload generator on the device:
 * fills count images of w x h RGB, image i uses seed0 + i. */
int jpegb200_synth_rgb_device(uint8_t *d_rgb, int width, int height, int count,
                              uint64_t image_stride, uint32_t seed0, int amp, void *cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* JPEGB200_H */
