/*
 * oracle/jpeg_oracle.c -- TEST INFRASTRUCTURE ONLY (see jpeg_oracle.h).
 *
 * Plain-C restatement of the reference natural_c encode path.  Every function
 * names the reference lines whose arithmetic it follows.  Floating point is
 * written so that gcc emits exactly the reference's operation order
 * (cvtsi2ss, mulss, mulss, addss per term); build with -ffp-contract=off and
 * never with -ffast-math.
 *
 * Citations are relative to /root/reference/natural_c/.
 */
#include "jpeg_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- tables (src/core/jpeg_tables.c:3-48, ITU-T T.81 Annex K) ------------ */

const uint8_t orc_quant_luma[64] = {
    16, 11, 10, 16,  24,  40,  51,  61,   12, 12, 14, 19,  26,  58,  60,  55,
    14, 13, 16, 24,  40,  57,  69,  56,   14, 17, 22, 29,  51,  87,  80,  62,
    18, 22, 37, 56,  68, 109, 103,  77,   24, 35, 55, 64,  81, 104, 113,  92,
    49, 64, 78, 87, 103, 121, 120, 101,   72, 92, 95, 98, 112, 100, 103,  99};

const uint8_t orc_dc_counts[16] = {0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
const uint8_t orc_dc_values[12] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11};
const uint8_t orc_ac_counts[16] = {0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7D};
const uint8_t orc_ac_values[162] = {
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61,
    0x07, 0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xA1, 0x08, 0x23, 0x42, 0xB1, 0xC1, 0x15, 0x52,
    0xD1, 0xF0, 0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0A, 0x16, 0x17, 0x18, 0x19, 0x1A, 0x25,
    0x26, 0x27, 0x28, 0x29, 0x2A, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3A, 0x43, 0x44, 0x45,
    0x46, 0x47, 0x48, 0x49, 0x4A, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5A, 0x63, 0x64,
    0x65, 0x66, 0x67, 0x68, 0x69, 0x6A, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7A, 0x83,
    0x84, 0x85, 0x86, 0x87, 0x88, 0x89, 0x8A, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99,
    0x9A, 0xA2, 0xA3, 0xA4, 0xA5, 0xA6, 0xA7, 0xA8, 0xA9, 0xAA, 0xB2, 0xB3, 0xB4, 0xB5, 0xB6,
    0xB7, 0xB8, 0xB9, 0xBA, 0xC2, 0xC3, 0xC4, 0xC5, 0xC6, 0xC7, 0xC8, 0xC9, 0xCA, 0xD2, 0xD3,
    0xD4, 0xD5, 0xD6, 0xD7, 0xD8, 0xD9, 0xDA, 0xE1, 0xE2, 0xE3, 0xE4, 0xE5, 0xE6, 0xE7, 0xE8,
    0xE9, 0xEA, 0xF1, 0xF2, 0xF3, 0xF4, 0xF5, 0xF6, 0xF7, 0xF8, 0xF9, 0xFA};

/* zig-zag visiting order, raster index of the i-th visited coefficient
 * (src/core/zigzag.c:7-15; the same map is used for DQT, io/jpeg_handler.c:25-34) */
const uint8_t orc_zigzag_order[64] = {
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

/* 6-decimal cosine table cos((2s+1) f pi/16) as literally tabulated by the
 * reference, indexed [spatial][frequency] (src/core/dct.c:9-18).  The last
 * digits are NOT symmetric (-0.382684 / 0.195091 / -0.923879 in rows 4-7);
 * that asymmetry is part of the reference's output and is kept. */
static const float k_cos[8][8] = {
    {1.000000f, 0.980785f, 0.923880f, 0.831470f, 0.707107f, 0.555570f, 0.382683f, 0.195090f},
    {1.000000f, 0.831470f, 0.382683f, -0.195090f, -0.707107f, -0.980785f, -0.923880f, -0.555570f},
    {1.000000f, 0.555570f, -0.382683f, -0.980785f, -0.707107f, 0.195090f, 0.923880f, 0.831470f},
    {1.000000f, 0.195090f, -0.923880f, -0.555570f, 0.707107f, 0.831470f, -0.382683f, -0.980785f},
    {1.000000f, -0.195090f, -0.923880f, 0.555570f, 0.707107f, -0.831470f, -0.382684f, 0.980785f},
    {1.000000f, -0.555570f, -0.382684f, 0.980785f, -0.707107f, -0.195090f, 0.923880f, -0.831470f},
    {1.000000f, -0.831470f, 0.382684f, 0.195091f, -0.707107f, 0.980785f, -0.923879f, 0.555570f},
    {1.000000f, -0.980785f, 0.923880f, -0.831470f, 0.707107f, -0.555570f, 0.382684f, -0.195090f}};

/* C(0)=0.707107, C(k>0)=1 (src/core/dct.c:4-6) */
static float norm_factor(int k) { return k == 0 ? 0.707107f : 1.000000f; }

/* ---- stage 1: luma + edge-replicated padding (converter.c:15-16,28-55) ---- */

int orc_pad8(int n) { return (n + 7) & ~7; }

void orc_luma_pad(const uint8_t *rgb, int w, int h, uint8_t *y)
{
    const int wp = orc_pad8(w), hp = orc_pad8(h);
    for (int row = 0; row < hp; ++row) {
        const int sr = row < h ? row : h - 1;              /* converter.c:31 */
        const uint8_t *src = rgb + (size_t)sr * (size_t)w * 3u;
        uint8_t *dst = y + (size_t)row * (size_t)wp;
        for (int col = 0; col < wp; ++col) {
            const int sc = col < w ? col : w - 1;          /* converter.c:36 */
            const uint8_t *px = src + (size_t)sc * 3u;
            const uint32_t acc = 77u * px[0] + 150u * px[1] + 29u * px[2];
            dst[col] = (uint8_t)(acc >> 8);                /* converter.c:51-53 */
        }
    }
}

/* ---- stage 2: level shift (converter.c:83-87) ---------------------------- */

void orc_level_shift(const uint8_t *y, size_t n, int8_t *out)
{
    for (size_t i = 0; i < n; ++i) out[i] = (int8_t)((int)y[i] - 128);
}

/* ---- stage 3: direct 8x8 DCT-II (dct.c:63-96) ---------------------------- */

void orc_fdct_block(const int8_t in[64], float out[64])
{
    for (int u = 0; u < 8; ++u) {           /* u: vertical frequency, pairs with row index  */
        for (int v = 0; v < 8; ++v) {       /* v: horizontal frequency, pairs with col index */
            float acc = 0.0f;
            for (int r = 0; r < 8; ++r) {
                for (int c = 0; c < 8; ++c) {
                    /* dct.c:75-83: (pixel * cos[r][u]) * cos[c][v], then add; no FMA */
                    float t = (float)in[r * 8 + c];
                    t = t * k_cos[r][u];
                    t = t * k_cos[c][v];
                    acc = acc + t;
                }
            }
            /* dct.c:93: ((0.25f * cu) * cv) * sum */
            float scale = 0.25f * norm_factor(u);
            scale = scale * norm_factor(v);
            out[u * 8 + v] = scale * acc;
        }
    }
}

/* Block tiling; coefficients land in raster image layout (dct.c:119-147). */
void orc_fdct_image(const int8_t *img, int wp, int hp, float *coef)
{
    int8_t blk[64];
    float  res[64];
    for (int by = 0; by + 8 <= hp; by += 8) {
        for (int bx = 0; bx + 8 <= wp; bx += 8) {
            for (int r = 0; r < 8; ++r)
                memcpy(blk + r * 8, img + (size_t)(by + r) * (size_t)wp + (size_t)bx, 8);
            orc_fdct_block(blk, res);
            for (int r = 0; r < 8; ++r)
                memcpy(coef + (size_t)(by + r) * (size_t)wp + (size_t)bx, res + r * 8,
                       8 * sizeof(float));
        }
    }
}

/* ---- stage 4: quantization (quantization.c:20-38) ------------------------ */

void orc_quantize(const float *coef, int wp, int hp, int16_t *q)
{
    for (int row = 0; row < hp; ++row) {
        for (int col = 0; col < wp; ++col) {
            const size_t i = (size_t)row * (size_t)wp + (size_t)col;
            const float step = (float)orc_quant_luma[(row & 7) * 8 + (col & 7)];
            q[i] = (int16_t)roundf(coef[i] / step);        /* quantization.c:34-36 */
        }
    }
}

/* ---- stage 5: zig-zag to block-major (zigzag.c:43-65) -------------------- */

void orc_zigzag(const int16_t *q, int wp, int hp, int16_t *zz)
{
    size_t b = 0;
    for (int by = 0; by < hp; by += 8) {
        for (int bx = 0; bx < wp; bx += 8, ++b) {
            for (int i = 0; i < 64; ++i) {
                const int pos = orc_zigzag_order[i];
                zz[b * 64 + (size_t)i] =
                    q[(size_t)(by + (pos >> 3)) * (size_t)wp + (size_t)(bx + (pos & 7))];
            }
        }
    }
}

/* ---- stage 6: DC differences + AC run lengths (rle.c:9-35,59-124) -------- */

static uint8_t magnitude_class(int16_t v)                  /* rle.c:9-22 */
{
    int a = v < 0 ? -v : v;
    uint8_t n = 0;
    while (a) { ++n; a >>= 1; }
    return n;
}

static uint16_t amplitude_bits(int16_t v)                  /* rle.c:24-35 */
{
    return v > 0 ? (uint16_t)v : (uint16_t)(v - 1);        /* note: v==0 gives 0xFFFF */
}

static void emit(orc_symbol *out, size_t *n, uint8_t sym, uint16_t amp, uint8_t nbits)
{
    if (out) {
        out[*n].symbol = sym;
        out[*n].amplitude = amp;
        out[*n].nbits = nbits;
    }
    ++*n;
}

/* the chain of rle.c:59-124 over `nblocks` blocks, starting from predictor *pred_io (0 at the start
 * of an image, rle.c:59) and leaving the last block's DC there */
static size_t rle_chain(const int16_t *zz, size_t nblocks, orc_symbol *out, int16_t *pred_io)
{
    size_t n = 0;
    int16_t pred = *pred_io;
    for (size_t b = 0; b < nblocks; ++b) {
        const int16_t *c = zz + b * 64;
        const int16_t diff = (int16_t)(c[0] - pred);       /* rle.c:68-70 */
        pred = c[0];
        const uint8_t dcs = magnitude_class(diff);
        emit(out, &n, dcs, amplitude_bits(diff), dcs);     /* rle.c:72-76 */

        int last = 0;                                      /* rle.c:83-89 */
        for (int k = 63; k > 0; --k)
            if (c[k] != 0) { last = k; break; }

        int run = 0;
        for (int k = 1; k <= last; ++k) {                  /* rle.c:92-117 */
            if (c[k] == 0) { ++run; continue; }
            for (; run >= 16; run -= 16) emit(out, &n, 0xF0, 0, 0);   /* ZRL, rle.c:99-103 */
            const uint8_t s = magnitude_class(c[k]);
            emit(out, &n, (uint8_t)((run << 4) | s), amplitude_bits(c[k]), s);
            run = 0;
        }
        if (last < 63) emit(out, &n, 0x00, 0, 0);          /* EOB, rle.c:121-123 */
    }
    *pred_io = pred;
    return n;
}

size_t orc_rle(const int16_t *zz, size_t nblocks, orc_symbol *out)
{
    int16_t pred = 0;                                      /* rle.c:59: one chain per image */
    return rle_chain(zz, nblocks, out, &pred);
}

/* ---- stage 7: canonical Huffman + bit packing (huffman.c:26-193) --------- */

typedef struct { uint16_t code; uint8_t len; } hcode;

static void canonical_codes(const uint8_t counts[16], const uint8_t *values, hcode *tab)
{                                                          /* huffman.c:89-104 */
    uint16_t next = 0;
    int vi = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < counts[len - 1]; ++i) {
            tab[values[vi]].code = next++;
            tab[values[vi]].len = (uint8_t)len;
            ++vi;
        }
        next = (uint16_t)(next << 1);
    }
}

static void huffman_tables(hcode dc[16], hcode ac[256])    /* huffman.c:106-117 */
{
    memset(dc, 0, 16 * sizeof(hcode));
    memset(ac, 0, 256 * sizeof(hcode));
    canonical_codes(orc_dc_counts, orc_dc_values, dc);
    canonical_codes(orc_ac_counts, orc_ac_values, ac);
}

typedef struct {
    uint8_t *dst;      /* NULL: count only */
    size_t   size;
    uint64_t acc;      /* low `fill` bits are pending, MSB first */
    int      fill;
} bitsink;

static void sink_byte(bitsink *s, uint8_t b)               /* huffman.c:26-32 */
{
    if (s->dst) s->dst[s->size] = b;
    ++s->size;
    if (b == 0xFF) {
        if (s->dst) s->dst[s->size] = 0x00;
        ++s->size;
    }
}

static void sink_bits(bitsink *s, uint16_t value, uint8_t n)   /* huffman.c:35-62 */
{
    if (n == 0) return;
    value &= (uint16_t)((1u << n) - 1u);                   /* huffman.c:39 */
    s->acc = (s->acc << n) | value;
    s->fill += n;
    while (s->fill >= 8) {
        s->fill -= 8;
        sink_byte(s, (uint8_t)(s->acc >> s->fill));
    }
}

static void sink_flush(bitsink *s)                         /* huffman.c:65-81: ZERO padding */
{
    if (s->fill > 0) sink_byte(s, (uint8_t)(s->acc << (8 - s->fill)));
    s->fill = 0;
}

/* huffman.c:139-191 for a run of blocks, appending to an open bit sink (no flush) */
static void huffman_blocks(bitsink *sink, const hcode dc[16], const hcode ac[256], const orc_symbol *sym, size_t nsym,
                           size_t nblocks)
{
    bitsink s = *sink;
    size_t i = 0;
    for (size_t b = 0; b < nblocks && i < nsym; ++b) {     /* huffman.c:139-141 */
        const orc_symbol d = sym[i++];                     /* huffman.c:145-153 */
        sink_bits(&s, dc[d.symbol].code, dc[d.symbol].len);
        sink_bits(&s, d.amplitude, d.nbits);
        int done = 1;
        while (done < 64 && i < nsym) {                    /* huffman.c:158-188 */
            const orc_symbol a = sym[i++];
            sink_bits(&s, ac[a.symbol].code, ac[a.symbol].len);
            if (a.nbits) sink_bits(&s, a.amplitude, a.nbits);
            if (a.symbol == 0x00) break;
            done += (a.symbol == 0xF0) ? 16 : ((a.symbol >> 4) & 15) + 1;
        }
    }
    *sink = s;
}

size_t orc_huffman(const orc_symbol *sym, size_t nsym, size_t nblocks, uint8_t *out)
{
    hcode dc[16], ac[256];
    huffman_tables(dc, ac);
    bitsink s = {out, 0, 0, 0};
    huffman_blocks(&s, dc, ac, sym, nsym, nblocks);
    sink_flush(&s);
    return s.size;
}

uint32_t orc_block_bits(const int16_t zz[64], int16_t prev_dc)
{
    hcode dc[16], ac[256];
    huffman_tables(dc, ac);
    orc_symbol tmp[80];
    int16_t blk[64];
    memcpy(blk, zz, sizeof(blk));
    blk[0] = (int16_t)(zz[0] - prev_dc);                   /* predictor 0 => diff */
    const size_t n = orc_rle(blk, 1, tmp);
    uint32_t bits = (uint32_t)dc[tmp[0].symbol].len + tmp[0].nbits;
    for (size_t i = 1; i < n; ++i) bits += (uint32_t)ac[tmp[i].symbol].len + tmp[i].nbits;
    return bits;
}

/* ---- composed paths ------------------------------------------------------- */

size_t orc_coefficients(const uint8_t *rgb, int w, int h, int16_t *zz)
{
    const int wp = orc_pad8(w), hp = orc_pad8(h);
    const size_t n = (size_t)wp * (size_t)hp;
    uint8_t *y = (uint8_t *)malloc(n);
    int8_t *cy = (int8_t *)malloc(n);
    float *f = (float *)malloc(n * sizeof(float));
    int16_t *q = (int16_t *)malloc(n * sizeof(int16_t));
    if (!y || !cy || !f || !q) { free(y); free(cy); free(f); free(q); return 0; }
    orc_luma_pad(rgb, w, h, y);
    orc_level_shift(y, n, cy);
    orc_fdct_image(cy, wp, hp, f);
    orc_quantize(f, wp, hp, q);
    orc_zigzag(q, wp, hp, zz);
    free(y); free(cy); free(f); free(q);
    return n / 64;
}

size_t orc_encode_scan(const uint8_t *rgb, int w, int h, uint8_t **out)
{
    const size_t nblocks = (size_t)(orc_pad8(w) / 8) * (size_t)(orc_pad8(h) / 8);
    *out = NULL;
    int16_t *zz = (int16_t *)malloc(nblocks * 64 * sizeof(int16_t));
    if (!zz) return 0;
    if (orc_coefficients(rgb, w, h, zz) != nblocks) { free(zz); return 0; }
    const size_t nsym = orc_rle(zz, nblocks, NULL);
    orc_symbol *sym = (orc_symbol *)malloc((nsym ? nsym : 1) * sizeof(orc_symbol));
    if (!sym) { free(zz); return 0; }
    orc_rle(zz, nblocks, sym);
    free(zz);
    const size_t nbytes = orc_huffman(sym, nsym, nblocks, NULL);
    uint8_t *buf = (uint8_t *)malloc(nbytes ? nbytes : 1);
    if (buf) orc_huffman(sym, nsym, nblocks, buf);
    free(sym);
    *out = buf;
    return buf ? nbytes : 0;
}

void orc_free(void *p) { free(p); }

/* ---- JFIF header bytes (io/jpeg_handler.c:7-110,220-233) ------------------ */

static uint8_t *put16(uint8_t *p, unsigned v) { p[0] = (uint8_t)(v >> 8); p[1] = (uint8_t)v; return p + 2; }

size_t orc_jfif_header(int w, int h, uint8_t out[328])
{
    uint8_t *p = out;
    /* SOI + APP0: JFIF 1.01, units=1 (dpi), 96x96, no thumbnail (jpeg_handler.c:7-22) */
    p = put16(p, 0xFFD8); p = put16(p, 0xFFE0); p = put16(p, 16);
    memcpy(p, "JFIF", 5); p += 5;
    p = put16(p, 0x0101); *p++ = 1; p = put16(p, 96); p = put16(p, 96); *p++ = 0; *p++ = 0;
    /* DQT, table 0, 8-bit, zig-zag order (jpeg_handler.c:36-49) */
    p = put16(p, 0xFFDB); p = put16(p, 67); *p++ = 0;
    for (int i = 0; i < 64; ++i) *p++ = orc_quant_luma[orc_zigzag_order[i]];
    /* SOF0: 8-bit, ORIGINAL (unpadded) h,w truncated to u16, 1 component (jpeg_handler.c:52-67,226) */
    p = put16(p, 0xFFC0); p = put16(p, 11); *p++ = 8;
    p = put16(p, (uint16_t)h); p = put16(p, (uint16_t)w);
    *p++ = 1; *p++ = 1; *p++ = 0x11; *p++ = 0;
    /* DHT DC then DHT AC (jpeg_handler.c:70-93) */
    p = put16(p, 0xFFC4); p = put16(p, 31); *p++ = 0x00;
    memcpy(p, orc_dc_counts, 16); p += 16; memcpy(p, orc_dc_values, 12); p += 12;
    p = put16(p, 0xFFC4); p = put16(p, 181); *p++ = 0x10;
    memcpy(p, orc_ac_counts, 16); p += 16; memcpy(p, orc_ac_values, 162); p += 162;
    /* SOS: length 8, comp 1, tables 0/0, Ss=0 Se=63 Ah/Al=0 (jpeg_handler.c:96-110) */
    p = put16(p, 0xFFDA); p = put16(p, 8); *p++ = 1; *p++ = 1; *p++ = 0x00;
    *p++ = 0; *p++ = 63; *p++ = 0;
    return (size_t)(p - out);
}

/* ---- synthetic workload ---------------------------- */

static uint32_t tri(uint32_t t, uint32_t period)
{
    const uint32_t m = t % (2u * period);
    const uint32_t d = m > period ? m - period : period - m;
    return d * 255u / period;
}

/* rows [y0, y0 + rows) of the w-wide synthetic image, written from rgb[0] on */
static void synth_rows(int w, int y0, int rows, uint32_t seed, int amp, uint8_t *rgb)
{
    static const int offset[3] = {10, 0, -10};
    const uint32_t span = (uint32_t)(2 * amp + 1);
    for (int y = 0; y < rows; ++y) {
        for (int x = 0; x < w; ++x) {
            const uint32_t ux = (uint32_t)x, uy = (uint32_t)(y0 + y);
            const uint32_t base = (tri(ux + 2u * uy, 419u) + tri(3u * ux + (1u << 20) - uy, 1021u) +
                                   tri(uy, 173u) + tri(ux, 67u)) / 4u;
            for (int c = 0; c < 3; ++c) {
                uint32_t hsh = (ux * 0x9E3779B1u) ^ (uy * 0x85EBCA77u) ^
                               (((uint32_t)c * 0xC2B2AE3Du) ^ (seed * 0x27D4EB2Fu));
                hsh ^= hsh >> 15; hsh *= 0x2C1B3C6Du;
                hsh ^= hsh >> 12; hsh *= 0x297A2D39u;
                hsh ^= hsh >> 15;
                const int noise = (int)((hsh >> 24) % span) - amp;
                int v = (int)base + offset[c] + noise;
                v = v < 0 ? 0 : (v > 255 ? 255 : v);
                rgb[((size_t)y * (size_t)w + (size_t)x) * 3u + (size_t)c] = (uint8_t)v;
            }
        }
    }
}

void orc_synth_rgb(int w, int h, uint32_t seed, int amp, uint8_t *rgb) { synth_rows(w, 0, h, seed, amp, rgb); }

/* ---- streaming encode of a synthetic image of any size --------------------
 * For images beyond the reference's int-indexed buffers (> 715 Mpixel: bmp_handler.c:78,119,
 * converter.c:39 overflow) the whole-image functions above would need tens of GB.
 * This walks the image in bands of whole block rows: a band's pixels are generated and taken
 * through stages 1-5 (position-independent), `threads` bands at a time in parallel; the entropy
 * stages then consume the bands in order with the DC predictor (rle.c:59-70) and the bit
 * accumulator (huffman.c:35-62) carried across bands.  Same arithmetic, same order, size_t
 * indices.  tests/test_oracle.py checks it against orc_encode_scan and the reference's hashes. */
#include <pthread.h>

typedef struct {
    int w, y0, rows;
    uint32_t seed;
    int amp;
    int16_t *zz;          /* out: (wp/8)*(rows_p/8)*64 coefficients */
    size_t nblocks;
    int failed;
} band_job;

static void *band_worker(void *arg)
{
    band_job *j = (band_job *)arg;
    uint8_t *rgb = (uint8_t *)malloc((size_t)j->w * (size_t)j->rows * 3u);
    if (!rgb) { j->failed = 1; return NULL; }
    synth_rows(j->w, j->y0, j->rows, j->seed, j->amp, rgb);
    j->failed = orc_coefficients(rgb, j->w, j->rows, j->zz) != j->nblocks;   /* ragged last band: rows replicate, converter.c:31 */
    free(rgb);
    return NULL;
}

size_t orc_encode_scan_synth_banded(int w, int h, uint32_t seed, int amp, int band_block_rows, int threads, uint8_t **out)
{
    *out = NULL;
    if (w <= 0 || h <= 0 || band_block_rows <= 0 || threads <= 0 || threads > 256) return 0;
    const int bw = orc_pad8(w) / 8, band_px = band_block_rows * 8;
    const int nbands = (h + band_px - 1) / band_px;
    const size_t band_blocks = (size_t)bw * (size_t)band_block_rows;
    hcode dc[16], ac[256];
    huffman_tables(dc, ac);
    size_t cap = (size_t)bw * (size_t)(orc_pad8(h) / 8) * 24u + 4096u;   /* grown on demand */
    uint8_t *buf = (uint8_t *)malloc(cap);
    int16_t *zz = (int16_t *)malloc((size_t)threads * band_blocks * 64u * sizeof(int16_t));
    orc_symbol *sym = (orc_symbol *)malloc(band_blocks * 64u * sizeof(orc_symbol));
    band_job *jobs = (band_job *)calloc((size_t)threads, sizeof(band_job));
    pthread_t *tid = (pthread_t *)calloc((size_t)threads, sizeof(pthread_t));
    int ok = buf && zz && sym && jobs && tid;
    bitsink s = {buf, 0, 0, 0};
    int16_t pred = 0;
    for (int b0 = 0; ok && b0 < nbands; b0 += threads) {
        const int n = nbands - b0 < threads ? nbands - b0 : threads;
        for (int i = 0; i < n; ++i) {
            band_job *j = &jobs[i];
            j->w = w; j->y0 = (b0 + i) * band_px;
            j->rows = h - j->y0 < band_px ? h - j->y0 : band_px;
            j->seed = seed; j->amp = amp;
            j->zz = zz + (size_t)i * band_blocks * 64u;
            j->nblocks = (size_t)bw * (size_t)(orc_pad8(j->rows) / 8);
            j->failed = 0;
            if (pthread_create(&tid[i], NULL, band_worker, j) != 0) { j->failed = 1; tid[i] = 0; band_worker(j); }
        }
        for (int i = 0; i < n; ++i) {
            if (tid[i]) pthread_join(tid[i], NULL);
            tid[i] = 0;
        }
        for (int i = 0; ok && i < n; ++i) {
            const band_job *j = &jobs[i];
            if (j->failed) { ok = 0; break; }
            const size_t nsym = rle_chain(j->zz, j->nblocks, sym, &pred);
            if (s.size + nsym * 8u + 64u > cap) {          /* <= 27 bits per symbol, doubled by stuffing at worst */
                cap = (s.size + nsym * 8u + 64u) * 2u;
                uint8_t *nb = (uint8_t *)realloc(buf, cap);
                if (!nb) { ok = 0; break; }
                buf = nb;
                s.dst = buf;
            }
            huffman_blocks(&s, dc, ac, sym, nsym, j->nblocks);
        }
    }
    if (ok) sink_flush(&s);
    free(zz); free(sym); free(jobs); free(tid);
    if (!ok) { free(buf); return 0; }
    *out = buf;
    return s.size;
}
