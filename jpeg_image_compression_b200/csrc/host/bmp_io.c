/* bmp_io.c -- host-side BMP reader (kept in C, not accelerated; north_star).
 *
 * Behavioural restatement of the reference loader natural_c/src/io/bmp_handler.c:15-129:
 * accepts 'BM' files with a 24-bit BI_RGB DIB whose first 40 header bytes are a
 * BITMAPINFOHEADER prefix (V4/V5 headers work because pixel data is located through
 * bfOffBits), returns top-down interleaved RGB with no row padding, and prints the
 * reference's diagnostics on stderr.  Same failure behaviour: NULL. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "jpegb200.h"

static uint32_t rd32(const unsigned char *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
static uint16_t rd16(const unsigned char *p) { return (uint16_t)(p[0] | (p[1] << 8)); }

void freeBMPImage(BMPImage *image)                       /* bmp_handler.c:5-12 */
{
    if (!image) return;
    free(image->data);
    free(image);
}

BMPImage *loadBMPImage(const char *filename)
{
    FILE *f = fopen(filename, "rb");
    if (!f) {
        fprintf(stderr, "Error: Unable to open file: %s\n", filename);
        return NULL;
    }
    unsigned char fh[14], ih[40];
    if (fread(fh, sizeof fh, 1, f) != 1) {
        fprintf(stderr, "Error: Failed to read BMP file header.\n");
        fclose(f);
        return NULL;
    }
    if (rd16(fh) != 0x4D42) {                            /* bmp_handler.c:30 */
        fprintf(stderr, "Error: File is not a valid BMP file.\n");
        fclose(f);
        return NULL;
    }
    if (fread(ih, sizeof ih, 1, f) != 1) {
        fprintf(stderr, "Error: Failed to read BMP info header.\n");
        fclose(f);
        return NULL;
    }
    if (rd16(ih + 14) != 24) {                           /* bmp_handler.c:44 */
        fprintf(stderr, "Error: Only 24-bit BMP images are supported.\n");
        fclose(f);
        return NULL;
    }
    if (rd32(ih + 16) != 0) {                            /* bmp_handler.c:49 */
        fprintf(stderr, "Error: Compressed BMP images are not supported.\n");
        fclose(f);
        return NULL;
    }
    BMPImage *img = (BMPImage *)malloc(sizeof *img);
    if (!img) {
        fprintf(stderr, "Error: Memory allocation failed for BMPImage struct.\n");
        fclose(f);
        return NULL;
    }
    img->width = (int32_t)rd32(ih + 4);
    img->height = (int32_t)rd32(ih + 8);
    int bottom_up = 1;                                   /* bmp_handler.c:68-72 */
    if (img->height < 0) {
        img->height = -img->height;
        bottom_up = 0;
    }
    const size_t w = (size_t)(img->width > 0 ? img->width : 0), h = (size_t)img->height;
    const size_t pitch = (w * 3u + 3u) & ~(size_t)3u;    /* bmp_handler.c:75 */
    const size_t nbytes = w * h * 3u;                    /* 64-bit sizes: no overflow above 715 Mpx */
    img->data = (uint8_t *)malloc(nbytes > 0 ? nbytes : 1u);
    if (!img->data) {
        fprintf(stderr, "Error: Memory allocation failed for pixel data.\n");
        free(img);
        fclose(f);
        return NULL;
    }
    if (fseek(f, (long)rd32(fh + 10), SEEK_SET) != 0) {  /* bmp_handler.c:88 */
        fprintf(stderr, "Error: Unable to seek to bitmap data.\n");
        freeBMPImage(img);
        fclose(f);
        return NULL;
    }
    unsigned char *row = (unsigned char *)malloc(pitch ? pitch : 1u);
    if (!row) {
        fprintf(stderr, "Error: Memory allocation failed for row buffer.\n");
        freeBMPImage(img);
        fclose(f);
        return NULL;
    }
    for (size_t i = 0; i < h; ++i) {
        if (fread(row, 1, pitch, f) != pitch) {
            fprintf(stderr, "Error: Insufficient data reading row %d\n", (int)i);
            free(row);
            freeBMPImage(img);
            fclose(f);
            return NULL;
        }
        uint8_t *dst = img->data + (bottom_up ? h - 1 - i : i) * w * 3u;
        for (size_t x = 0; x < w; ++x) {                 /* BGR -> RGB, bmp_handler.c:115-122 */
            dst[3 * x + 0] = row[3 * x + 2];
            dst[3 * x + 1] = row[3 * x + 1];
            dst[3 * x + 2] = row[3 * x + 0];
        }
    }
    free(row);
    fclose(f);
    return img;
}
