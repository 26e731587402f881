/*
 * oracle/jpeg_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement of the reference natural_c grayscale JPEG encode path
 * (strbac-damjan/jpeg-image-compression, natural_c/src/core).  It exists to
 * check the CUDA path; only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may call it.  The product library
 * (libjpegb200.so) never links or loads it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below
 * against the reference's own objects (oracle/_ref/libnaturalc_ref.so, built by
 * oracle/Makefile from the sources under /root/reference where they lie) on the
 * reference's four asset BMPs and on synthetic edge cases, and against the
 * golden fixtures in tests/golden/ that were minted from that reference build
 * (tests/golden/make_golden.py).
 *
 * All citations are relative to /root/reference/natural_c/.
 */
#ifndef JPEG_ORACLE_H
#define JPEG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* One run-length symbol; same field order and 6-byte footprint as the
 * reference's RLESymbol (include/rle.h:8-14): u8 @0, u16 @2, u8 @4. */
typedef struct {
    uint8_t  symbol;
    uint16_t amplitude;
    uint8_t  nbits;
} orc_symbol;

/* the five constant tables (src/core/jpeg_tables.c:3-48) */
extern const uint8_t orc_quant_luma[64];
extern const uint8_t orc_dc_counts[16];
extern const uint8_t orc_dc_values[12];
extern const uint8_t orc_ac_counts[16];
extern const uint8_t orc_ac_values[162];
extern const uint8_t orc_zigzag_order[64];

int    orc_pad8(int n);                                          /* converter.c:15-16 */
void   orc_luma_pad(const uint8_t *rgb, int w, int h, uint8_t *y);/* converter.c:28-55 */
void   orc_level_shift(const uint8_t *y, size_t n, int8_t *out);  /* converter.c:83-87 */
void   orc_fdct_block(const int8_t in[64], float out[64]);        /* dct.c:63-96      */
void   orc_fdct_image(const int8_t *img, int wp, int hp, float *coef); /* dct.c:119-147 */
void   orc_quantize(const float *coef, int wp, int hp, int16_t *q);    /* quantization.c:20-38 */
void   orc_zigzag(const int16_t *q, int wp, int hp, int16_t *zz);      /* zigzag.c:43-65 */
size_t orc_rle(const int16_t *zz, size_t nblocks, orc_symbol *out);    /* rle.c:59-124 (out==NULL: count only) */
size_t orc_huffman(const orc_symbol *sym, size_t nsym, size_t nblocks,
                   uint8_t *out);                                  /* huffman.c:121-193 (out==NULL: size only) */

/* Bit cost of one zig-zag block given the DC predictor (rle.c + huffman.c
 * combined); lets tests check the device's per-block code lengths. */
uint32_t orc_block_bits(const int16_t zz[64], int16_t prev_dc);

/* RGB (top-down, interleaved, pitch 3*w) -> zig-zag coefficients.
 * zz must hold (wp/8)*(hp/8)*64 int16.  Returns number of blocks. */
size_t orc_coefficients(const uint8_t *rgb, int w, int h, int16_t *zz);

/* Whole hot path: RGB -> stuffed scan bytes (what saveJPEGGrayscale fwrites at
 * io/jpeg_handler.c:252).  *out is malloc'd; caller frees with orc_free. */
size_t orc_encode_scan(const uint8_t *rgb, int w, int h, uint8_t **out);
void   orc_free(void *p);

/* The 328 header bytes written by io/jpeg_handler.c:220-233 for (w,h). */
size_t orc_jfif_header(int w, int h, uint8_t out[328]);

/* Synthetic workload generator of SURVEY.md section 8(d) (integer only). */
void   orc_synth_rgb(int w, int h, uint32_t seed, int amp, uint8_t *rgb);

/* Streaming encode of the synthetic image (w, h, seed, amp) of ANY size, including those beyond the
 * reference's int-indexed buffers (> 715 Mpixel): bands of `band_block_rows` block rows
 * go through stages 1-5 on `threads` threads, the entropy stages consume them in order with the DC
 * predictor and bit accumulator carried along.  Equals orc_encode_scan(orc_synth_rgb(...)) wherever
 * that is computable (checked in tests/test_oracle.py).  *out is malloc'd; free with orc_free. */
size_t orc_encode_scan_synth_banded(int w, int h, uint32_t seed, int amp, int band_block_rows, int threads,
                                    uint8_t **out);

#ifdef __cplusplus
}
#endif
#endif
