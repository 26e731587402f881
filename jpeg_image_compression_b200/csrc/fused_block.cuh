// fused_block.cuh -- K1, the fused block kernel.
//
// One pass over the RGB payload: luma (converter.c:51) + edge replication
// (converter.c:31,36) + level shift (converter.c:84-86) + 8x8 forward DCT (dct.c:63-96)
// + quantization (quantization.c:34-36) + zig-zag (zigzag.c:51-60) + the AC part of the
// block's Huffman bit cost (rle.c:83-123 with huffman.c code lengths).
//
// Work unit: a STRIP of 32 consecutive 8x8 blocks in one block row (256 x 8 pixels,
// 6 KB of RGB).  One warp owns a strip:
//   1. TMA into shared memory, mbarrier-completed: for 16-byte aligned inputs ONE tensor-map copy
//      (cp.async.bulk.tensor.3d / UTMALDG) of the 8 x 768-byte tile, otherwise one bulk copy
//      (cp.async.bulk / UBLKCP) per pixel row at the rows' natural 16-byte phase (any width / base
//      alignment / pitch);
//   2. cooperative luma pass: 4 pixels per lane-step (funnel-shift realign, PRMT, DP4A),
//      bytes written to a 256 x 8 Y tile;
//   3. the next strip's bulk copies are issued (the raw tile is free again) so their HBM
//      latency overlaps the arithmetic below;
//   4. one lane per 8x8 block, block entirely in registers: magic-number u8->f32,
//      scaled even/odd butterfly DCT (rows scalar, columns two at a time with packed fp32),
//      quantization by one FFMA per bound (see below), zig-zag packing to int8 with PRMT, AC bit
//      cost by table look-up in a loop over the bytes parked in shared memory (the hot loop has
//      to fit the 32 KB instruction cache), 4 x 128-bit stores of the 64 coefficients.
//
// Bit-exactness.  The reference sums 64 products sequentially in fp32 with two unfused
// multiplies per term; replaying that costs ~136 flop/pixel.  Instead the butterfly
// value T (ideal cosines, FMA) is bracketed: the reference's sum s_ref satisfies
// |s_ref - g*T| <= kGamma * A with A = sum|p| over the block (DESIGN.md section 3 derives the
// bound: 65.1u*A for the reference's own roundings, 16u*A for the butterfly's, 15.7u*A
// for the 6-decimal LUT vs ideal cosines, 6u*A for the scale/divide roundings, 102.8u*A in
// total with u = 2^-24; the kernel takes the minimum with a mean-separated refinement,
// u*(17.5*A + 103*Ac + 500), Ac = sum|p - mean|).  Quantization is monotone in s, so if
// rounding (T-E)*rk and (T+E)*rk give the same integer that integer IS the reference's;
// otherwise (about 2e-4 of the coefficients) the lane re-evaluates that one coefficient in
// the reference's exact operation order (exact_quantized below).  The DC term is an exact integer sum in both
// formulations and is quantized with the reference's own operation sequence.
#pragma once

#include <cuda.h>          // CUtensorMap (type only; the encoder function is looked up at run time)

#include "common.cuh"
#include "scan_pack.cuh"

namespace jb {

constexpr int K1_WARPS = 8;
constexpr int K1_THREADS = K1_WARPS * 32;
constexpr int K1_CTAS_PER_SM = 2;
constexpr int RAW_PITCH = 784;                       // 768 payload + 16 bytes of alignment slack
constexpr int RAW_BYTES = 8 * RAW_PITCH;             // 6272
constexpr int Y_PITCH = 256;
constexpr int Y_BYTES = 8 * Y_PITCH;                 // 2048
constexpr int ZZ_PITCH = 17;                          // words per lane: 16 coefficient words + 1 (bank-conflict-free)
constexpr int ZZ_BYTES = 32 * ZZ_PITCH * 4;          // 2176
constexpr int HALO_BYTES = 256;                      // the raster-predecessor block's 8 x 24 RGB bytes (192), padded
constexpr int K1_WARP_SMEM = RAW_BYTES + Y_BYTES + ZZ_BYTES + HALO_BYTES;    // 10752 = 84 x 128
constexpr int ACLUT_BYTES = 16400;                   // 63 rows x 260 bytes + EOB length, then 16 DC lengths
constexpr int K1_TABLE_BYTES = 16512;                // table area rounded up to 128 bytes: tensor-map copies need that alignment
constexpr int K1_SMEM = K1_TABLE_BYTES + K1_WARPS * K1_WARP_SMEM;   // 102528: two CTAs per SM
constexpr int TMAP_ROW_BYTES = 768;                  // a tensor-map box is dense: 8 rows x 768 bytes

// Everything a warp needs to know about one strip; computed once per strip (32-bit math).
struct StripCtx {
    const uint8_t *row0;     // first byte of the strip's first pixel row
    uint64_t block0;         // global index of the strip's first block
    int64_t pitch;           // bytes from a pixel row to the next (signed: bottom-up BMP rows are walked backwards)
    int rmax;                // last real pixel row of the strip, relative (rows beyond replicate it)
    int npx;                 // real pixels per row in the strip (<= 256)
    int vb;                  // 8x8 blocks in the strip (<= 32)
    uint32_t mispack;        // 16-byte phase of each of the 8 row pointers, 4 bits per row
    int tx, ty, tz;          // tensor-map coordinates of the tile: word column, pixel row, image
    // the block that precedes the strip in raster order (its DC is the strip's first predictor)
    const uint8_t *halo;     // its top-left pixel; nullptr for the first strip of an image
    int halo_rmax, halo_cmax;// last real row / column inside that block (edges replicate)
};

// position of a strip: image, block row, strip inside the block row.  Found by division once per warp and
// then advanced incrementally (a warp's strips are `nwarps` apart).
struct StripPos {
    uint32_t img, brow, sx;
};

__device__ __forceinline__ StripPos strip_pos(const Geom &g, uint32_t s)
{
    const uint32_t per_image = (uint32_t)g.bh * (uint32_t)g.spr;
    StripPos p;
    p.img = s / per_image;
    const uint32_t rem = s - p.img * per_image;
    p.brow = rem / (uint32_t)g.spr;
    p.sx = rem - p.brow * (uint32_t)g.spr;
    return p;
}

__device__ __forceinline__ void strip_advance(const Geom &g, StripPos &p, uint32_t dq, uint32_t dr)
{
    p.sx += dr;
    p.brow += dq;
    if (p.sx >= (uint32_t)g.spr) {
        p.sx -= (uint32_t)g.spr;
        ++p.brow;
    }
    while (p.brow >= (uint32_t)g.bh) {
        p.brow -= (uint32_t)g.bh;
        ++p.img;
    }
}

__device__ __forceinline__ StripCtx strip_ctx(const Geom &g, const StripPos &p)
{
    const uint32_t img = p.img, brow = p.brow, sx = p.sx;
    StripCtx c;
    c.pitch = g.row_pitch;
    c.row0 = g.rgb + (uint64_t)img * g.image_stride + (int64_t)(brow * 8u) * c.pitch + (uint64_t)sx * 768u;
    c.block0 = (uint64_t)img * g.blocks_per_image + (uint64_t)brow * (uint32_t)g.bw + sx * 32u;
    c.rmax = min(7, g.h - 1 - (int)(brow * 8u));                          // converter.c:31
    c.npx = min(256, g.w - (int)(sx * 256u));
    c.vb = min(32, g.bw - (int)(sx * 32u));
    c.mispack = 0;                                  // filled in by strip_issue_loads
    c.tx = (int)(sx * 192u);
    c.ty = (int)(brow * 8u);
    c.tz = (int)img;
    if (sx > 0) {                                   // previous block is in the same block row
        c.halo = c.row0 - 24;
        c.halo_rmax = c.rmax;
        c.halo_cmax = 7;
    } else if (brow > 0) {                          // last block of the previous block row
        const int hx0 = (g.bw - 1) * 8;
        c.halo = c.row0 - 8 * c.pitch + (int64_t)hx0 * 3;
        c.halo_rmax = 7;
        c.halo_cmax = g.w - 1 - hx0;
    } else {
        c.halo = nullptr;
        c.halo_rmax = c.halo_cmax = 0;
    }
    return c;
}

// stage the strip's 8 pixel rows: one TMA bulk copy per row (cp.async.bulk -> UBLKCP), lane r issuing
// row r, all completed through the warp's mbarrier.  Each copy starts at the row's 16-byte-aligned
// address and covers whole 16-byte chunks, so any width / base alignment works.  Also records the
// rows' 16-byte phases in c.mispack.  Must be called by the whole warp.
__device__ __forceinline__ void strip_issue_loads(StripCtx &c, uint8_t *raw, uint8_t *hal, uint64_t *bar, int lane,
                                                  const CUtensorMap *tmap)
{
    if (tmap) {
        // aligned input: ONE tensor-map copy of the 8 x 768-byte box (rows and columns beyond the image are
        // zero-filled and never used: the luma pass clamps the row and stops at the last real pixel)
        c.mispack = 0;
        if (lane == 0) {
            fence_proxy_async();
            mbar_expect_tx(bar, 8 * TMAP_ROW_BYTES);
            tensor_g2s_3d(raw, tmap, c.tx, c.ty, c.tz, bar);
        }
        // the raster-predecessor block (8 rows x 24 bytes, 8-byte aligned here) rides along as 24 asynchronous
        // 8-byte copies whose completion is folded into the same mbarrier phase (every lane arrives once)
        if (c.halo != nullptr && lane < 24) {
            const int r = lane / 3, part = lane - 3 * r;
            cp_async8(hal + lane * 8, c.halo + (int64_t)min(r, c.halo_rmax) * c.pitch + 8 * part);
        }
        cp_async_mbar_arrive_noinc(bar);
        return;
    }
    const int r = lane & 7;
    const uint8_t *row = c.row0 + (int64_t)min(r, c.rmax) * c.pitch;
    const uint32_t mis = (uint32_t)(uintptr_t)row & 15u;
    const uint32_t bytes = lane < 8 ? (mis + 3u * (uint32_t)c.npx + 15u) & ~15u : 0u;     // <= 784
    c.mispack = __reduce_or_sync(0xffffffffu, lane < 8 ? mis << (4 * r) : 0u);
    const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
    fence_proxy_async();                                 // the tile was just read through the generic proxy
    if (lane == 0) mbar_expect_tx(bar, total);
    __syncwarp();
    if (lane < 8) bulk_g2s(raw + r * RAW_PITCH, row - mis, bytes, bar);
}

// luma of 4 consecutive pixels held in 3 words: Y = (77R + 150G + 29B) >> 8  (converter.c:51)
__device__ __forceinline__ uint32_t luma4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t wt_lo, uint32_t wt_hi)
{
    const uint32_t y0 = __dp4a(w0, wt_lo, 0u);
    const uint32_t y1 = __dp4a(__byte_perm(w0, w1, 0x0543u), wt_lo, 0u);
    const uint32_t y2 = __dp4a(__byte_perm(w1, w2, 0x0432u), wt_lo, 0u);
    const uint32_t y3 = __dp4a(w2, wt_hi, 0u);
    return __byte_perm(__byte_perm(y0, y1, 0x0051u), __byte_perm(y2, y3, 0x0051u), 0x5410u);
}

__device__ __forceinline__ float u8_to_centered(uint32_t word, int byte)
{
    // 0x4B0000xx is 2^23 + xx; subtracting 2^23 + 128 is exact: (float)(xx - 128), converter.c:84
    const uint32_t bits = __byte_perm(word, 0x4B000000u, 0x7650u + (uint32_t)byte);
    return __uint_as_float(bits) - 8388736.0f;
}

// Reference-order evaluation of ONE quantized coefficient (dct.c:65-93,
// quantization.c:34-36): sequential fp32 sum, two unfused multiplies per term, IEEE
// divide, round half away from zero.  yblk: the block's Y bytes (pitch Y_PITCH).
__device__ __noinline__ int exact_quantized(const uint8_t *yblk, int u, int v)
{
    float cv[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) cv[c] = c_ref_cos[c * 8 + v];
    float acc = 0.0f;
#pragma unroll 1
    for (int r = 0; r < 8; ++r) {
        const float cu = c_ref_cos[r * 8 + u];
        const uint2 w = *reinterpret_cast<const uint2 *>(yblk + r * Y_PITCH);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            float t = __fmul_rn(u8_to_centered(c < 4 ? w.x : w.y, c & 3), cu);
            t = __fmul_rn(t, cv[c]);
            acc = __fadd_rn(acc, t);
        }
    }
    const float f = __fmul_rn(c_ref_scale[u * 8 + v], acc);
    return (int)roundf(__fdiv_rn(f, c_quant_f[u * 8 + v]));
}

// Scaled 8-point DCT-II butterfly.  Outputs T_k = S_k / g_k with S_k = sum x_i cos((2i+1)k pi/16),
// g = {1,1,cos(pi/8),1,cos(pi/4),1,cos(pi/8),1}; the g factors are folded into c_rk.
__device__ __forceinline__ void dct8(float &x0, float &x1, float &x2, float &x3, float &x4, float &x5,
                                     float &x6, float &x7)
{
    constexpr float C1 = 0.98078528040323044913f, C3 = 0.83146961230254523708f;
    constexpr float C5 = 0.55557023301960222474f, C7 = 0.19509032201612826785f;
    constexpr float TAN = 0.41421356237309504880f;              // tan(pi/8)
    const float a0 = x0 + x7, a1 = x1 + x6, a2 = x2 + x5, a3 = x3 + x4;
    const float b0 = x0 - x7, b1 = x1 - x6, b2 = x2 - x5, b3 = x3 - x4;
    const float e0 = a0 + a3, e1 = a1 + a2, d0 = a0 - a3, d1 = a1 - a2;
    x0 = e0 + e1;
    x4 = e0 - e1;
    x2 = fmaf(d1, TAN, d0);
    x6 = fmaf(d0, TAN, -d1);
    x1 = fmaf(b3, C7, fmaf(b2, C5, fmaf(b1, C3, b0 * C1)));
    x3 = fmaf(b3, -C5, fmaf(b2, -C1, fmaf(b1, -C7, b0 * C3)));
    x5 = fmaf(b3, C3, fmaf(b2, C7, fmaf(b1, -C1, b0 * C5)));
    x7 = fmaf(b3, -C1, fmaf(b2, C3, fmaf(b1, -C5, b0 * C7)));
}

// ---- the same butterfly on two columns at once (Blackwell packed fp32: FADD2 / FMUL2 / FFMA2) ----------
// Element for element the operation sequence of dct8 (fma.rn.f32x2 etc. round each half like the scalar
// instruction), so the guard-band analysis is unchanged; it halves the instruction count of the column
// pass.  A pair lives in an aligned register pair, packing and unpacking are free.  The constants
// come from constant memory (ptxas keeps them in uniform registers).
typedef unsigned long long f32x2;
__constant__ float2 c_dct2[9] = {
    {0.98078528040323044913f, 0.98078528040323044913f},   // 0: C1
    {0.83146961230254523708f, 0.83146961230254523708f},   // 1: C3
    {0.55557023301960222474f, 0.55557023301960222474f},   // 2: C5
    {0.19509032201612826785f, 0.19509032201612826785f},   // 3: C7
    {0.41421356237309504880f, 0.41421356237309504880f},   // 4: tan(pi/8)
    {-0.98078528040323044913f, -0.98078528040323044913f}, // 5: -C1
    {-0.55557023301960222474f, -0.55557023301960222474f}, // 6: -C5
    {-0.19509032201612826785f, -0.19509032201612826785f}, // 7: -C7
    {-8388736.0f, -8388736.0f}};                           // 8: -(2^23 + 128), see u8_to_centered
__device__ __forceinline__ f32x2 pack2(float lo, float hi)
{
    f32x2 p;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "f"(lo), "f"(hi));
    return p;
}
__device__ __forceinline__ void unpack2(f32x2 p, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b)
{
    f32x2 d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c)
{
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ f32x2 kdct2(int i) { return *reinterpret_cast<const f32x2 *>(&c_dct2[i]); }

__device__ __forceinline__ void dct8_2(f32x2 &x0, f32x2 &x1, f32x2 &x2, f32x2 &x3, f32x2 &x4, f32x2 &x5, f32x2 &x6,
                                       f32x2 &x7)
{
    const f32x2 C1 = kdct2(0), C3 = kdct2(1), C5 = kdct2(2), C7 = kdct2(3), TAN = kdct2(4);
    const f32x2 NC1 = kdct2(5), NC5 = kdct2(6), NC7 = kdct2(7);
    const f32x2 a0 = add2(x0, x7), a1 = add2(x1, x6), a2 = add2(x2, x5), a3 = add2(x3, x4);
    const f32x2 b0 = sub2(x0, x7), b1 = sub2(x1, x6), b2 = sub2(x2, x5), b3 = sub2(x3, x4);
    const f32x2 e0 = add2(a0, a3), e1 = add2(a1, a2), d0 = sub2(a0, a3), d1 = sub2(a1, a2);
    const f32x2 nd1 = sub2(a2, a1);                               // -d1, exactly
    x0 = add2(e0, e1);
    x4 = sub2(e0, e1);
    x2 = fma2(d1, TAN, d0);
    x6 = fma2(d0, TAN, nd1);
    x1 = fma2(b3, C7, fma2(b2, C5, fma2(b1, C3, mul2(b0, C1))));
    x3 = fma2(b3, NC5, fma2(b2, NC1, fma2(b1, NC7, mul2(b0, C3))));
    x5 = fma2(b3, C3, fma2(b2, C7, fma2(b1, NC1, mul2(b0, C5))));
    x7 = fma2(b3, NC1, fma2(b2, C3, fma2(b1, NC5, mul2(b0, C7))));
}

// scale class of a frequency index: 0 -> g=1, 1 -> g=cos(pi/8), 2 -> g=cos(pi/4)
__host__ __device__ constexpr int gclass(int k) { return (k == 2 || k == 6) ? 1 : (k == 4 ? 2 : 0); }

__global__ void __launch_bounds__(K1_THREADS, K1_CTAS_PER_SM)
k_fused_blocks(const Geom g, int8_t *__restrict__ coef, uint32_t *__restrict__ blkinfo,
               StripRec *__restrict__ strips, uint32_t *__restrict__ strip_bits, const uint8_t *__restrict__ tables,
               unsigned long long *__restrict__ flagged_counter, const int exact_mode,
               uint64_t *__restrict__ lookback_state, const uint64_t lookback_words,
               unsigned long long *__restrict__ trace, const __grid_constant__ CUtensorMap tmap_param)
{
#ifdef JPEGB200_TRACE   // tracing build only (make trace -> libjpegb200_trace.so)
#define K1_TRACE(slot) do { if (trace && lane == 0) trace[(uint64_t)(blockIdx.x * K1_WARPS + warp) * 8 + (slot)] = globaltimer_ns(); } while (0)
#else
#define K1_TRACE(slot) do { } while (0)
#endif
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t *aclut = smem;
    const uint8_t *s_dclen = smem + TBL_DC_LEN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *raw = smem + K1_TABLE_BYTES + warp * K1_WARP_SMEM;
    const CUtensorMap *tmap = g.use_tmap ? &tmap_param : nullptr;
    const uint32_t raw_pitch = g.use_tmap ? TMAP_ROW_BYTES : RAW_PITCH;
    uint8_t *ybuf = raw + RAW_BYTES;
    uint32_t *zs = reinterpret_cast<uint32_t *>(ybuf + Y_BYTES) + lane * ZZ_PITCH;   // this lane's 64 coefficient bytes
    uint8_t *hal = ybuf + Y_BYTES + ZZ_BYTES;                    // predecessor block, tensor-map path only

    K1_TRACE(0);
    // Static schedule: persistent warp i takes strips i, i + nwarps, ...; the 8 warps of a CTA work on 8
    // consecutive strips (contiguous 6 KB runs of the same pixel rows).  Measured alternatives: interleaving
    // warp indices across CTAs balances a partial last round over the SMs but loses that locality (+4 %
    // time); handing strips out by ticket cannot balance a ~2-round image either, because the next strip
    // must be known a whole strip ahead for the prefetch.
    const uint32_t nwarps = gridDim.x * K1_WARPS;
    const uint32_t total = (uint32_t)g.total_strips;
    uint32_t s = blockIdx.x * K1_WARPS + warp;
    __shared__ __align__(8) uint64_t s_bar[K1_WARPS + 1];      // one mbarrier per warp (pixel tiles) + one for the table
    uint64_t *bar = &s_bar[warp];
    if (lane == 0) mbar_init(bar, g.use_tmap ? 33 : 1);          // tensor-map path: + one asynchronous arrival per lane
    if (threadIdx.x == 0) mbar_init(&s_bar[K1_WARPS], 1);
    mbar_fence_init();
    __syncthreads();
    uint32_t phase = 0;
    StripCtx cur;
    StripPos pos;
    const uint32_t dq = nwarps / (uint32_t)g.spr, dr = nwarps - dq * (uint32_t)g.spr;
    if (s < total) {
        pos = strip_pos(g, s);
        cur = strip_ctx(g, pos);
        strip_issue_loads(cur, raw, hal, bar, lane, tmap);                  // first strip's pixels are in flight ...
    }
    if (threadIdx.x == 0) {                                      // ... while the bit-cost table is staged (one 16 KB bulk copy)
        mbar_expect_tx(&s_bar[K1_WARPS], ACLUT_BYTES);
        bulk_g2s(aclut, tables + TBL_ACLUT, ACLUT_BYTES, &s_bar[K1_WARPS]);
    }
    // reset the look-back state of the entropy kernel (K2) that follows in the stream: keeps a
    // whole encode at two launches and CUDA-graph replayable
    for (uint64_t i = (uint64_t)blockIdx.x * K1_THREADS + threadIdx.x; i < lookback_words;
         i += (uint64_t)gridDim.x * K1_THREADS)
        lookback_state[i] = 0;
    bool table_ready = false;                                    // waited for at its first use
    K1_TRACE(1);

    uint32_t nflag = 0;
    const uint32_t xforce = exact_mode ? 0x01010101u : 0u;
    float magic;                                     // 1.5 * 2^23 held in a register: leaves the FFMA's
    asm("mov.f32 %0, 0f4B400000;" : "=f"(magic));     // constant-bank slot to the quantizer constant c_rk

    for (; s < total; s += nwarps) {
        // luma sum of the raster-predecessor block: lane l covers row l/4, columns 2*(l%4) and +1;
        // the loads are issued now and consumed after the transform
        uint32_t halo_y = 0;
        if (cur.halo && !tmap) {
            const uint8_t *hp = cur.halo + (int64_t)min(lane >> 2, cur.halo_rmax) * cur.pitch;
            const uint8_t *p0 = hp + 3 * min(2 * (lane & 3), cur.halo_cmax), *p1 = hp + 3 * min(2 * (lane & 3) + 1, cur.halo_cmax);
            const uint32_t k0 = g.wt_lo & 0xFFu, k1 = (g.wt_lo >> 8) & 0xFFu, k2 = (g.wt_lo >> 16) & 0xFFu;
            halo_y = ((k0 * p0[0] + k1 * p0[1] + k2 * p0[2]) >> 8) + ((k0 * p1[0] + k1 * p1[1] + k2 * p1[2]) >> 8);
        }
        mbar_wait(bar, phase);
        phase ^= 1u;
        if (cur.halo && tmap) {                          // same sum from the block staged in shared memory
            const uint8_t *hp = hal + 24 * (lane >> 2);
            const uint8_t *p0 = hp + 3 * min(2 * (lane & 3), cur.halo_cmax), *p1 = hp + 3 * min(2 * (lane & 3) + 1, cur.halo_cmax);
            const uint32_t k0 = g.wt_lo & 0xFFu, k1 = (g.wt_lo >> 8) & 0xFFu, k2 = (g.wt_lo >> 16) & 0xFFu;
            halo_y = ((k0 * p0[0] + k1 * p0[1] + k2 * p0[2]) >> 8) + ((k0 * p1[0] + k1 * p1[1] + k2 * p1[2]) >> 8);
        }
        if (!table_ready) K1_TRACE(2);

        // ---- luma pass -> 256 x 8 Y tile ----------------------------------------------------
        if (cur.npx == 256 && (cur.mispack & 0x33333333u) == 0u) {
            // full strip, word-aligned rows: 3 LDS + DP4A/PRMT per 4 pixels.  Two rows per trip, not fully
            // unrolled: the kernel's hot loop has to stay inside the 32 KB instruction cache (L1.5)
#pragma unroll 2
            for (int r = 0; r < 8; ++r) {
                const uint32_t *rw = reinterpret_cast<const uint32_t *>(raw + min(r, cur.rmax) * raw_pitch + ((cur.mispack >> (4 * r)) & 12u)) + 3 * lane;
                uint32_t *yo = reinterpret_cast<uint32_t *>(ybuf + r * Y_PITCH);
                yo[lane] = luma4(rw[0], rw[1], rw[2], g.wt_lo, g.wt_hi);
                yo[lane + 32] = luma4(rw[96], rw[97], rw[98], g.wt_lo, g.wt_hi);
            }
        } else {
            // any width / any base alignment: funnel-shift the row to word alignment
            const int ngroups = (cur.npx + 3) >> 2;
#pragma unroll 1
            for (int item = lane; item < 512; item += 32) {
                const int r = item >> 6, gc = item & 63;
                if (gc < ngroups) {
                    const uint32_t mis = (cur.mispack >> (4 * r)) & 15u;
                    const uint32_t *rw = reinterpret_cast<const uint32_t *>(raw + min(r, cur.rmax) * raw_pitch) + (mis >> 2) + 3 * gc;
                    const uint32_t sh = (mis & 3u) * 8u;
                    const uint32_t q0 = rw[0], q1 = rw[1], q2 = rw[2], q3 = rw[3];
                    reinterpret_cast<uint32_t *>(ybuf + r * Y_PITCH)[gc] =
                        luma4(__funnelshift_r(q0, q1, sh), __funnelshift_r(q1, q2, sh), __funnelshift_r(q2, q3, sh), g.wt_lo, g.wt_hi);
                }
            }
        }
        __syncwarp();

        // raw tile is free: prefetch the next strip while this one is transformed
        const StripCtx me = cur;
        if (s + nwarps < total) {
            strip_advance(g, pos, dq, dr);
            cur = strip_ctx(g, pos);
            strip_issue_loads(cur, raw, hal, bar, lane, tmap);
        }

        // right-edge replication inside the last real block (converter.c:36)
        const int padpx = me.vb * 8 - me.npx;
        if (padpx > 0) {
            if (lane < padpx) {
#pragma unroll
                for (int r = 0; r < 8; ++r) ybuf[r * Y_PITCH + me.npx + lane] = ybuf[r * Y_PITCH + me.npx - 1];
            }
            __syncwarp();
        }

        uint32_t my_bits = 0, my_last = 0;                     // this lane's block: bit cost, last non-zero AC index
        int my_dc = 0;
        if (lane < me.vb) {
            const uint8_t *yblk = ybuf + lane * 8;
            float x[8][8];
            uint32_t absdev = 0;                                   // A = sum |Y - 128|
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const uint2 v = *reinterpret_cast<const uint2 *>(yblk + r * Y_PITCH);
                absdev = __vsadu4(v.x, 0x80808080u) + absdev;
                absdev = __vsadu4(v.y, 0x80808080u) + absdev;
                // u8 -> centred f32 (converter.c:84): byte into the mantissa of 2^23, then one packed subtraction
                // per two pixels (0x4B0000xx is 2^23 + xx; subtracting 2^23 + 128 is exact)
#pragma unroll
                for (int c = 0; c < 4; c += 2) {
                    const f32x2 lo = pack2(__uint_as_float(__byte_perm(v.x, 0x4B000000u, 0x7650u + c)),
                                           __uint_as_float(__byte_perm(v.x, 0x4B000000u, 0x7651u + c)));
                    const f32x2 hi = pack2(__uint_as_float(__byte_perm(v.y, 0x4B000000u, 0x7650u + c)),
                                           __uint_as_float(__byte_perm(v.y, 0x4B000000u, 0x7651u + c)));
                    unpack2(add2(lo, kdct2(8)), x[r][c], x[r][c + 1]);
                    unpack2(add2(hi, kdct2(8)), x[r][c + 4], x[r][c + 5]);
                }
                dct8(x[r][0], x[r][1], x[r][2], x[r][3], x[r][4], x[r][5], x[r][6], x[r][7]);   // along the row
            }
            // down the columns, two adjacent columns per instruction
#pragma unroll
            for (int c = 0; c < 8; c += 2) {
                f32x2 p[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) p[r] = pack2(x[r][c], x[r][c + 1]);
                dct8_2(p[0], p[1], p[2], p[3], p[4], p[5], p[6], p[7]);
#pragma unroll
                for (int r = 0; r < 8; ++r) unpack2(p[r], x[r][c], x[r][c + 1]);
            }

            // Ac = sum |Y - m|, m = the block's rounded mean luma (from the exact DC sum)
            uint32_t absmean = 0;
            {
                const int mean = (int)rintf(x[0][0] * 0.015625f) + 128;
                const uint32_t m4 = (uint32_t)mean * 0x01010101u;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const uint2 v = *reinterpret_cast<const uint2 *>(yblk + r * Y_PITCH);
                    absmean = __vsadu4(v.x, m4) + absmean;
                    absmean = __vsadu4(v.y, m4) + absmean;
                }
            }
            // guard half-widths in T units for the 6 scale-class pairs
            const float eb = fminf((float)absdev * kGamma, fmaf((float)absdev, kGammaA, fmaf((float)absmean, kGammaC, kGamma0)));
            float ecls[3][3];
            {
                constexpr float IG[3] = {1.0f, 1.0823922002923940f, 1.4142135623730951f};   // 1/g, rounded up
#pragma unroll
                for (int a = 0; a < 3; ++a)
#pragma unroll
                    for (int b = 0; b < 3; ++b) ecls[a][b] = eb * (IG[a] * IG[b] * 1.000001f);
            }

            // quantize in zig-zag order, pack to int8 (low byte of the magic-biased float)
            uint32_t zw[16], xw[16];
#pragma unroll
            for (int w = 0; w < 16; ++w) {
                uint32_t hb[4], lb[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    constexpr uint8_t ZZ[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                                12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                                35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                                58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
                    const int pos = ZZ[4 * w + j], u = pos >> 3, v = pos & 7;
                    const float t = x[u][v];
                    const float e = ecls[gclass(u)][gclass(v)];
                    const float rk = c_rk[pos];
                    hb[j] = __float_as_uint(fmaf(__fadd_rn(t, e), rk, magic));
                    lb[j] = __float_as_uint(fmaf(__fadd_rn(t, -e), rk, magic));
                }
                const uint32_t h = __byte_perm(__byte_perm(hb[0], hb[1], 0x0040u), __byte_perm(hb[2], hb[3], 0x0040u), 0x5410u);
                const uint32_t l = __byte_perm(__byte_perm(lb[0], lb[1], 0x0040u), __byte_perm(lb[2], lb[3], 0x0040u), 0x5410u);
                zw[w] = h;
                xw[w] = (h ^ l) | xforce;
            }

            // DC: exact integer sum in both formulations; reference operation sequence
            // fl(k00 * S) / 16 then roundf (dct.c:93, quantization.c:36)
            {
                const float f = __fmul_rn(c_ref_scale[0], x[0][0]);
                const int dcq = (int)roundf(f * 0.0625f);
                zw[0] = (zw[0] & 0xFFFFFF00u) | ((uint32_t)dcq & 0xFFu);
                xw[0] &= 0xFFFFFF00u;
            }

            uint32_t anyx = 0;
#pragma unroll
            for (int w = 0; w < 16; ++w) anyx |= xw[w];
            if (anyx) {
                // rare (about 1 block in 200): the bracket straddles a rounding boundary for some
                // coefficient(s); re-evaluate exactly those in the reference's operation order
                uint64_t fm = 0;
#pragma unroll
                for (int w = 0; w < 16; ++w) {
                    fm |= (uint64_t)nonzero_nibble(xw[w]) << (4 * w);
                }
#pragma unroll 1
                while (fm) {
                    const int k = __ffsll((long long)fm) - 1;
                    fm &= fm - 1;
                    const int pos = c_zigzag[k];
                    const uint32_t q = (uint32_t)exact_quantized(yblk, pos >> 3, pos & 7) & 0xFFu;
                    const uint32_t sh = 8u * (uint32_t)(k & 3), keep = ~(0xFFu << sh), ins = q << sh;
                    const int kw = k >> 2;
#pragma unroll
                    for (int w = 0; w < 16; ++w)
                        if (w == kw) zw[w] = (zw[w] & keep) | ins;
                    ++nflag;
                }
            }

            if (!table_ready) {
                K1_TRACE(3);
                mbar_wait(&s_bar[K1_WARPS], 0);                    // the bit-cost table, staged during the first transform
                K1_TRACE(4);
            }
            // AC bit cost: code length + amplitude bits per non-zero coefficient, ZRLs, EOB
            // The walk is a real loop over bytes parked in shared memory, not 63 unrolled copies reading
            // registers: the kernel's hot loop has to fit the SM's 32 KB instruction cache, otherwise the 16
            // warps (each at its own place in the code) saturate the GPC-level instruction cache.
#pragma unroll
            for (int w = 0; w < 16; ++w) zs[w] = zw[w];
            const uint8_t *zb = reinterpret_cast<const uint8_t *>(zs);
            uint32_t bits = 0, lastk = 0;                          // lastk = ACLUT_STRIDE * (index of the last non-zero)
            const uint8_t *lut = aclut;                            // row of zero run 0 for position k
#pragma unroll 21                     // 3 trips: measured best between code size (instruction cache) and loop overhead
            for (int k = 1; k < 64; ++k) {
                const uint32_t byte = zb[k];
                bits += lut[byte - lastk];
                lut += ACLUT_STRIDE;
                if (byte != 0) lastk = (uint32_t)(lut - aclut);   // = k * ACLUT_STRIDE
            }
            my_last = (lastk * 253u) >> 16;                        // lastk / 260 for lastk <= 63*260
            if (my_last != 63u) bits += aclut[ACLUT_ROWS * ACLUT_STRIDE];   // EOB code length (rle.c:121-123)
            my_bits = bits;
            my_dc = (int)(int8_t)(zw[0] & 0xFFu);

            uint4 *dst = reinterpret_cast<uint4 *>(coef + (me.block0 + (uint32_t)lane) * 64);
            dst[0] = make_uint4(zw[0], zw[1], zw[2], zw[3]);
            dst[1] = make_uint4(zw[4], zw[5], zw[6], zw[7]);
            dst[2] = make_uint4(zw[8], zw[9], zw[10], zw[11]);
            dst[3] = make_uint4(zw[12], zw[13], zw[14], zw[15]);
        }
        if (!table_ready) {
            mbar_wait(&s_bar[K1_WARPS], 0);                        // lanes without a block; immediate for the others
            table_ready = true;
            K1_TRACE(5);
        }
        // DC-difference costs (rle.c:68-76).  The strip's first block is predicted from the block
        // before the strip, whose quantized DC follows from its 64 luma values alone (an exact integer
        // sum, same closed form as above).  The very first strip of an image is charged by K2 (its
        // predictor is 0, or the previous stripe's last DC in multi-GPU runs).
        {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) halo_y += __shfl_xor_sync(0xffffffffu, halo_y, o);
            int prev_dc = __shfl_up_sync(0xffffffffu, my_dc, 1);
            if (lane == 0) {
                const float f = __fmul_rn(c_ref_scale[0], (float)((int)halo_y - 8192));
                prev_dc = (int)roundf(f * 0.0625f);
            }
            if (lane < me.vb && (lane > 0 || me.halo)) my_bits += s_dclen[magnitude_class(my_dc - prev_dc)];
            uint32_t incl = my_bits;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += n;
            }
            const int first_dc = __shfl_sync(0xffffffffu, my_dc, 0);
            if (lane < me.vb) blkinfo[me.block0 + (uint32_t)lane] = blk_pack(incl - my_bits, my_last);
            if (lane == me.vb - 1) {
                strips[s] = StripRec{incl, (int16_t)first_dc, (int16_t)my_dc};
                strip_bits[s] = incl;
            }
        }
        __syncwarp();
#ifdef JPEGB200_TRACE
        // second half of the trace buffer: completion times of this warp's first 8 strips
        if (trace && lane == 0) {
            const uint32_t it = (s - (blockIdx.x * K1_WARPS + warp)) / nwarps;
            if (it < 8) trace[(uint64_t)(nwarps + blockIdx.x * K1_WARPS + warp) * 8 + it] = globaltimer_ns();
        }
#endif
    }

    K1_TRACE(6);
#ifdef JPEGB200_TRACE
    {   // slot 7: strips done * 1000 + coefficients this warp re-evaluated on the exact path
        uint32_t nf = nflag;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nf += __shfl_xor_sync(0xffffffffu, nf, o);
        if (trace && lane == 0) trace[(uint64_t)(blockIdx.x * K1_WARPS + warp) * 8 + 7] = nf;
    }
#endif
#undef K1_TRACE
    if (flagged_counter) {
        nflag += __shfl_xor_sync(0xffffffffu, nflag, 16);
        nflag += __shfl_xor_sync(0xffffffffu, nflag, 8);
        nflag += __shfl_xor_sync(0xffffffffu, nflag, 4);
        nflag += __shfl_xor_sync(0xffffffffu, nflag, 2);
        nflag += __shfl_xor_sync(0xffffffffu, nflag, 1);
        if (lane == 0 && nflag) atomicAdd(flagged_counter, (unsigned long long)nflag);
    }
}

}  // namespace jb
