/* jfif_io.c -- host-side JFIF container writer and the orchestrator shim.
 *
 * saveJPEGGrayscale keeps the reference's contract (natural_c/src/io/jpeg_handler.c:119-282:
 * same stdout lines, same file layout: 328 header bytes + scan + FFD9, bool result) but the
 * seven core-stage calls of jpeg_handler.c:133-201 are replaced by ONE call into the fused GPU
 * path, jpegb200_encode_scan_dbg.  No CPU fallback: if the GPU path fails the function prints
 * the reference's first error line and returns false. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jpegb200.h"

static const unsigned char zigzag_order[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

static unsigned char *be16(unsigned char *p, unsigned v)
{
    p[0] = (unsigned char)(v >> 8);
    p[1] = (unsigned char)v;
    return p + 2;
}

/* SOI+APP0 (20) | DQT (69) | SOF0 (13) | DHT-DC (33) | DHT-AC (183) | SOS (10) = 328 bytes,
 * jpeg_handler.c:7-110 in the order of :220-233. */
size_t jpegb200_jfif_header(int width, int height, uint8_t out[328])
{
    unsigned char *p = out;
    p = be16(p, 0xFFD8); p = be16(p, 0xFFE0); p = be16(p, 16);
    memcpy(p, "JFIF", 5); p += 5;
    p = be16(p, 0x0101);
    *p++ = 1;                                            /* units: dpi */
    p = be16(p, 96); p = be16(p, 96);
    *p++ = 0; *p++ = 0;                                  /* no thumbnail */

    p = be16(p, 0xFFDB); p = be16(p, 67);
    *p++ = 0x00;
    for (int i = 0; i < 64; ++i) *p++ = std_luminance_quant_tbl[zigzag_order[i]];

    p = be16(p, 0xFFC0); p = be16(p, 11);
    *p++ = 8;
    p = be16(p, (uint16_t)height);                       /* original, unpadded dimensions (:226) */
    p = be16(p, (uint16_t)width);
    *p++ = 1; *p++ = 1; *p++ = 0x11; *p++ = 0;

    p = be16(p, 0xFFC4); p = be16(p, 31);
    *p++ = 0x00;
    memcpy(p, std_dc_luminance_nrcodes, 16); p += 16;
    memcpy(p, std_dc_luminance_values, 12); p += 12;

    p = be16(p, 0xFFC4); p = be16(p, 181);
    *p++ = 0x10;
    memcpy(p, std_ac_luminance_nrcodes, 16); p += 16;
    memcpy(p, std_ac_luminance_values, 162); p += 162;

    p = be16(p, 0xFFDA); p = be16(p, 8);
    *p++ = 1; *p++ = 1; *p++ = 0x00; *p++ = 0; *p++ = 63; *p++ = 0;
    return (size_t)(p - out);
}

bool saveJPEGGrayscale(const char *filename, const BMPImage *img)
{
    FILE *file = fopen(filename, "wb");
    if (!file) {
        perror("Error opening output file");
        return false;
    }
    printf("Starting JPEG compression pipeline...\n");

    int16_t first[64];
    JpegEncoderBuffer *scan = jpegb200_encode_scan_dbg(img, first);
    if (!scan) {
        /* the reference reports the first stage that failed (jpeg_handler.c:135) */
        printf("Error: Failed to convert BMP to Grayscale (YImage).\n");
        fclose(file);
        return false;
    }
    printf("Natural C quant (First Block):\n");          /* jpeg_handler.c:168-175 */
    for (int r = 0; r < 8; ++r) {
        for (int c = 0; c < 8; ++c) printf("%d ", first[r * 8 + c]);
        printf("\n");
    }
    printf("Pipeline finished. Writing to file...\n");

    uint8_t header[328];
    const size_t hlen = jpegb200_jfif_header(img->width, img->height, header);
    if (fwrite(header, 1, hlen, file) != hlen) {
        printf("Error: Failed to write JPEG headers to file.\n");
        fclose(file);
        freeJpegEncoderBuffer(scan);
        return false;
    }
    bool ok = true;
    const size_t written = fwrite(scan->data, 1, scan->size, file);
    if (written != scan->size) {
        printf("Error: Failed to write bitstream data. Wrote %zu of %zu bytes.\n", written, scan->size);
        ok = false;
    } else {
        printf("Bitstream written: %zu bytes.\n", written);
    }
    const unsigned char eoi[2] = {0xFF, 0xD9};
    fwrite(eoi, 1, 2, file);
    fclose(file);
    freeJpegEncoderBuffer(scan);
    if (ok) printf("Compression successful. File saved: %s\n", filename);
    return ok;
}

void freeYImage(YImage *img)                             /* jpeg_handler.c:284-294 */
{
    if (!img) return;
    free(img->data);
    free(img);
}
