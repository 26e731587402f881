// fused_block.cuh -- K1, the fused block kernel.
//
// One pass over the RGB payload: luma (converter.c:51) + edge replication (converter.c:31,36) + level shift
// (converter.c:84-86) + 8x8 forward DCT (dct.c:63-96) + quantization (quantization.c:34-36) + zig-zag
// (zigzag.c:51-60).  Output: the quantized coefficients as int8 in zig-zag order, 32 + 32 bytes per block in two planes.
//
// Work unit: a STRIP of 32 consecutive 8x8 blocks in one block row (256 x 8 pixels, 6 KB of RGB).  One warp owns
// a strip, one lane a block:
//   1. TMA into shared memory, mbarrier-completed: for 16-byte aligned inputs ONE tensor-map copy
//      (cp.async.bulk.tensor.3d / UTMALDG) of the 8 x 768-byte tile, otherwise one bulk copy (UBLKCP) per pixel row;
//   2. luma pass, one lane per block: per pixel row three LDS.64 (the block's 8 pixels), eight DP4A whose accumulator
//      carries the fp16 exponent of 1024, so that one PRMT per pixel pair gives the fp16 pair {1024 + Y} and one HADD2
//      the level-shifted values; bytes written to a 256 x 8 Y tile, fp16 values (TC) with one STS.128 per row
//      into the MMA's A operand (ragged or oddly aligned strips: cooperative 4-pixels-per-lane path);
//   3. transform + quantization, one of
//      TC = true  (default): the 64-term sums of all 64 coefficients run on the tensor cores.  A tile of 128
//                 blocks (4 strips = 4 warps) is one tcgen05.mma chain D[128 x 128] = A[128 x 64] * B^T: A = the
//                 level-shifted luma values as fp16 integers in the UMMA operand layout in shared memory; no
//                 barrier: every warp bumps the group's arrival counter, the last one to arrive issues the MMA,
//                 and all of them request the next strip's pixels and do the block statistics while it runs;
//                 B = the reference's own LUT products cos[r][u]*cos[c][v] in 22-bit
//                 fixed point, split into two 11-bit integer limbs held as fp16 (shared memory, UMMA K-major
//                 layout).  Every product and partial sum is an integer below 2^24, so the fp32 accumulators in
//                 TMEM hold the EXACT integer sums; tcgen05.ld hands each lane its block's 128 limb sums, 16
//                 coefficients at a time, already in zig-zag order.  Everything after that is packed fp32
//                 (FFMA2 / FADD2): limb combine, guard bracket, quantization.
//      TC = false: scaled even/odd butterfly in registers (rows scalar, columns packed fp32), as in round 1.
//      In both cases the value is bracketed (see below) and the few coefficients whose bracket straddles a rounding
//      boundary are re-evaluated in the reference's exact operation order;
//   4. 128-bit stores of the block's coefficient bytes: positions 0..31 always, 32..63 only when any is non-zero.
// The entropy stage is a kernel of its own (strip_entropy.cuh): its sparse symbol walks are latency-bound and want
// two to three times the occupancy this register-heavy kernel can have (measured: DESIGN.md section 3).
//
// Bit-exactness.  The reference sums 64 products sequentially in fp32 with two unfused multiplies per term;
// replaying that costs ~136 flop/pixel.  Instead the fast value is bracketed: |s_ref - fast| <= gamma(A, Ac) with
// A = sum|p| over the block and Ac = sum|p - mean| (DESIGN.md section 3 derives the constants for both transforms).
// Quantization is monotone in s, so if rounding (t-E)*rk and (t+E)*rk gives the same integer that integer IS the
// reference's; otherwise the lane re-evaluates that one coefficient in the reference's exact operation order
// (exact_quantized).  The DC term is an exact integer sum in all formulations and is quantized with the reference's
// own operation sequence.
#pragma once

#include <cuda.h>          // CUtensorMap (type only; the encoder function is looked up at run time)

#include <cuda_fp16.h>

#include "common.cuh"

namespace jb {

constexpr int RAW_PITCH = 784;                       // 768 payload + 16 bytes of alignment slack
constexpr int RAW_BYTES = 8 * RAW_PITCH;             // 6272
constexpr int Y_PITCH = 256;
constexpr int Y_BYTES = 8 * Y_PITCH;                 // 2048
constexpr int TMAP_ROW_BYTES = 768;                  // a tensor-map box is dense: 8 rows x 768 bytes
constexpr int K1_BMAT_BYTES = 16384;                 // limb matrix of the tensor-core transform

// Launch shape and shared-memory carve-up of the two instantiations.
//   TC: one CTA of 16 warps per SM, warps in groups of four (a 128-block MMA tile: 16 KB of A in shared memory and
//       128 accumulator columns of tensor memory per group -- 4 x 128 = all 512 columns)
template <bool TC>
struct K1Cfg {
    static constexpr int WARPS = TC ? 16 : 8;
    static constexpr int THREADS = WARPS * 32;
    static constexpr int CTAS_PER_SM = TC ? 1 : 2;
    static constexpr int GROUPS = WARPS / 4;
    static constexpr int WARP_SMEM = RAW_BYTES + Y_BYTES;             // 8320 = 65 x 128
    static constexpr int A_TILE_BYTES = 16384;                       // fp16 [128 blocks][64] in UMMA K-major layout
    static constexpr int TABLE_BYTES = TC ? K1_BMAT_BYTES + GROUPS * A_TILE_BYTES : 0;
    static constexpr int SMEM = TABLE_BYTES + WARPS * WARP_SMEM;
    static constexpr int TMEM_COLS = 512;
    static_assert(WARP_SMEM % 128 == 0, "per-warp region must keep the tensor-map tiles 128-byte aligned");
};
constexpr int TC_GROUP_COLS = 128;                   // accumulator columns per group: 2 limbs x 64 coefficients
constexpr uint32_t TC_IDESC = (1u << 4)              // D: fp32
                              | (0u << 7) | (0u << 10)      // A, B: fp16
                              | (0u << 15) | (0u << 16)     // both K-major
                              | ((128u >> 3) << 17)         // N = 128
                              | ((128u >> 4) << 24);        // M = 128

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
};

// position of a strip: image, block row, strip inside the block row.  Found by division once per warp and
// then advanced incrementally (a warp's strips are a fixed distance apart).
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
    return c;
}

// stage the strip's 8 pixel rows: one tensor-map copy, or one TMA bulk copy per row (lane r issuing row r), completed
// through the warp's mbarrier.  Each row copy starts at the row's 16-byte-aligned address and covers whole 16-byte
// chunks, so any width / base alignment works.  Also records the rows' 16-byte phases in c.mispack.  Must be called
// by the whole warp.
__device__ __forceinline__ void strip_issue_loads(StripCtx &c, uint8_t *raw, uint64_t *bar, int lane, const CUtensorMap *tmap, bool fenced = false)
{
    if (tmap) {
        // aligned input: ONE tensor-map copy of the 8 x 768-byte box (rows and columns beyond the image are
        // zero-filled and never used: the luma pass clamps the row and stops at the last real pixel)
        c.mispack = 0;
        if (lane == 0) {
            if (!fenced) fence_proxy_async();
            mbar_expect_tx(bar, 8 * TMAP_ROW_BYTES);
            tensor_g2s_3d(raw, tmap, c.tx, c.ty, c.tz, bar);
        }
        return;
    }
    const int r = lane & 7;
    const uint8_t *row = c.row0 + (int64_t)min(r, c.rmax) * c.pitch;
    const uint32_t mis = (uint32_t)(uintptr_t)row & 15u;
    const uint32_t bytes = lane < 8 ? (mis + 3u * (uint32_t)c.npx + 15u) & ~15u : 0u;     // <= 784
    c.mispack = __reduce_or_sync(0xffffffffu, lane < 8 ? mis << (4 * r) : 0u);
    const uint32_t total = __reduce_add_sync(0xffffffffu, bytes);
    if (!fenced) fence_proxy_async();                    // the tile was just read through the generic proxy
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

// luma of two pixels (each in the weighted bytes of one word) as the fp16 pair {1024 + Y0, 1024 + Y1}: the DP4A
// accumulator 0x64000000 puts 0x64 above the 16-bit sum whose high byte is Y
__device__ __forceinline__ uint32_t luma2_half(uint32_t a, uint32_t b, uint32_t wa, uint32_t wb)
{
    return __byte_perm(__dp4a(a, wa, 0x64000000u), __dp4a(b, wb, 0x64000000u), 0x7531u);
}

__device__ __forceinline__ uint32_t half2_sub1152(uint32_t h)
{
    uint32_t r;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(h), "r"(0xE480E480u));
    return r;
}

// four luma bytes -> four level-shifted fp16 values (converter.c:84): 0x64yy is the fp16 1024 + yy, and subtracting
// 1152 is exact
__device__ __forceinline__ uint2 y4_to_half4(uint32_t y)
{
    const uint32_t lo = __byte_perm(y, 0x64646464u, 0x4140u), hi = __byte_perm(y, 0x64646464u, 0x4342u);
    uint2 r;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r.x) : "r"(lo), "r"(0xE480E480u));
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(r.y) : "r"(hi), "r"(0xE480E480u));
    return r;
}

__device__ __forceinline__ float u8_to_centered(uint32_t word, int byte)
{
    // 0x4B0000xx is 2^23 + xx; subtracting 2^23 + 128 is exact: (float)(xx - 128), converter.c:84
    const uint32_t bits = __byte_perm(word, 0x4B000000u, 0x7650u + (uint32_t)byte);
    return __uint_as_float(bits) - 8388736.0f;
}

// Reference-order evaluation of ONE quantized coefficient (dct.c:65-93, quantization.c:34-36): sequential fp32 sum of
// 64 products, two unfused multiplies per term, IEEE divide, round half away from zero.  Warp-cooperative: every lane
// forms two of the 64 products fl(fl(p * cos[r][u]) * cos[c][v]) -- they are independent -- and parks them in the
// warp's scratch; only the ordered summation is serial, done by the lane that owns the block (the other lanes of
// the warp, and through the tile barrier its three sibling warps, wait for it: a single-lane evaluation of all 64
// terms took four times as long).  yblk: the block's Y bytes (pitch Y_PITCH); scratch: 64 floats, 16-byte aligned.
// Must be called by the whole warp with uniform (yblk, u, v); returns the quantized value (valid on every lane).
__device__ __forceinline__ int exact_quantized_coop(const uint8_t *yblk, int u, int v, const float *s_cos, float *scratch, int lane)
{
    const int r = lane >> 2, c = 2 * (lane & 3);
    const float cu = s_cos[r * 8 + u];
    const uint32_t two = *reinterpret_cast<const uint16_t *>(yblk + r * Y_PITCH + c);
    const float p0 = (float)((int)(two & 0xFFu) - 128), p1 = (float)((int)(two >> 8) - 128);     // converter.c:84
    const float t0 = __fmul_rn(__fmul_rn(p0, cu), s_cos[c * 8 + v]);
    const float t1 = __fmul_rn(__fmul_rn(p1, cu), s_cos[(c + 1) * 8 + v]);
    *reinterpret_cast<float2 *>(scratch + 2 * lane) = make_float2(t0, t1);
    __syncwarp();
    float acc = 0.0f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const float4 t = *reinterpret_cast<const float4 *>(scratch + 4 * i);       // same address on every lane: broadcast
        acc = __fadd_rn(acc, t.x);
        acc = __fadd_rn(acc, t.y);
        acc = __fadd_rn(acc, t.z);
        acc = __fadd_rn(acc, t.w);
    }
    __syncwarp();
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

// ---- Blackwell packed fp32 (FADD2 / FMUL2 / FFMA2): two fp32 lanes per instruction, each rounded like the scalar
// instruction.  A pair lives in an aligned register pair, packing and unpacking are free. ------------------------
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
__device__ __forceinline__ f32x2 pack2u(uint32_t lo, uint32_t hi)
{
    f32x2 p;
    asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "r"(lo), "r"(hi));
    return p;
}
__device__ __forceinline__ void unpack2(f32x2 p, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p)); }
__device__ __forceinline__ void unpack2u(f32x2 p, uint32_t &lo, uint32_t &hi) { asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(p)); }
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

template <bool TC>
__global__ void __launch_bounds__(K1Cfg<TC>::THREADS, K1Cfg<TC>::CTAS_PER_SM)
k_fused_blocks(const Geom g, int8_t *__restrict__ coef, uint32_t *__restrict__ himask, const uint8_t *__restrict__ tables, unsigned long long *__restrict__ flagged_counter,
               const int exact_mode, uint64_t *__restrict__ lookback_state, const uint64_t lookback_words,
               unsigned long long *__restrict__ trace, const __grid_constant__ CUtensorMap tmap_param)
{
    using Cfg = K1Cfg<TC>;
    constexpr int WARPS = Cfg::WARPS;
#ifdef JPEGB200_TRACE   // tracing build only (make trace -> libjpegb200_trace.so)
#define K1_TRACE(slot) do { if (trace && lane == 0) trace[(uint64_t)(blockIdx.x * WARPS + warp) * 8 + (slot)] = globaltimer_ns(); } while (0)
// phases of the warp's 4th tile (steady state): third region of the trace buffer
#define K1_TRACE_TILE(slot) do { if (trace && lane == 0 && trace_it == 3u) trace[(uint64_t)(2u * stride + blockIdx.x * WARPS + warp) * 8 + (slot)] = globaltimer_ns(); } while (0)
#define K1_TRACE_TILE2(slot) do { if (trace && lane == 0 && trace_it == 3u) trace[(uint64_t)(3u * stride + blockIdx.x * WARPS + warp) * 8 + (slot)] = globaltimer_ns(); } while (0)
#else
#define K1_TRACE(slot) do { } while (0)
#define K1_TRACE_TILE(slot) do { } while (0)
#define K1_TRACE_TILE2(slot) do { } while (0)
#endif
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *raw = smem + Cfg::TABLE_BYTES + warp * Cfg::WARP_SMEM;
    const CUtensorMap *tmap = g.use_tmap ? &tmap_param : nullptr;
    const uint32_t raw_pitch = g.use_tmap ? TMAP_ROW_BYTES : RAW_PITCH;
    uint8_t *ybuf = raw + RAW_BYTES;
    __shared__ __align__(8) uint64_t s_bar[WARPS + 1 + 4];       // per-warp tile barriers, the table barrier, per-group MMA barriers
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_arrive[4];                             // per group: warps that have finished the tile's luma pass (wraps at 4)
    __shared__ float s_cos[64];                                  // the reference's cosine LUT (exact re-evaluation)
    __shared__ __align__(16) float s_scratch[WARPS][64];         // per warp: the 64 products of one re-evaluated coefficient

    K1_TRACE(0);
    // Static schedule.  TC: a group of four warps takes tiles of four consecutive strips, tile = group index, + number of
    // groups, ...; otherwise persistent warp i takes strips i, i + nwarps, ....  Either way the warps of a CTA work on
    // consecutive strips (contiguous 6 KB runs of the same pixel rows).
    const uint32_t total = (uint32_t)g.total_strips;
    const uint32_t stride = gridDim.x * WARPS;                   // distance between a warp's strips
    uint32_t s = blockIdx.x * WARPS + warp;
    uint32_t tile_s0 = TC ? (s & ~3u) : s;                       // group-uniform loop variable
    uint64_t *bar = &s_bar[warp];
    uint64_t *table_bar = &s_bar[WARPS];
    if (lane == 0) mbar_init(bar, 1);
    if (threadIdx.x == 0) {
        mbar_init(table_bar, 1);
        if (TC)
            for (int i = 0; i < Cfg::GROUPS; ++i) mbar_init(&s_bar[WARPS + 1 + i], 1);
    }
    if (threadIdx.x < 64) s_cos[threadIdx.x] = c_ref_cos[threadIdx.x];
    if (threadIdx.x < 4) s_arrive[threadIdx.x] = 0u;
    mbar_fence_init();
    __syncwarp();
    // The first strip's pixels are requested before the block-wide set-up below (TMEM allocation, barrier): a warp only
    // needs its own mbarrier, which its lane 0 has just initialised and fenced.
    uint32_t phase = 0;
    StripCtx cur;
    StripPos pos;
    const uint32_t dq = stride / (uint32_t)g.spr, dr = stride - dq * (uint32_t)g.spr;
    if (s < total) {
        pos = strip_pos(g, s);
        cur = strip_ctx(g, pos);
        strip_issue_loads(cur, raw, bar, lane, tmap);
    }
    if (TC && threadIdx.x == 0) {                                // ... and the limb matrix (one 16 KB TMA bulk copy)
        mbar_expect_tx(table_bar, K1_BMAT_BYTES);
        bulk_g2s(smem, tables + TBL_BMAT, K1_BMAT_BYTES, table_bar);
    }
    if (TC && warp == 0) tmem_alloc(&s_tmem, Cfg::TMEM_COLS);
    if (TC) tc_fence_before_sync();
    __syncthreads();
    if (TC) tc_fence_after_sync();
    // reset the look-back state of the merge kernel (K2) that follows in the stream: keeps a
    // whole encode CUDA-graph replayable without a memset node
    for (uint64_t i = (uint64_t)blockIdx.x * Cfg::THREADS + threadIdx.x; i < lookback_words;
         i += (uint64_t)gridDim.x * Cfg::THREADS)
        lookback_state[i] = 0;
    K1_TRACE(1);

    // tensor memory of this warp's group: accumulators at column 128 * group; the warp owns lanes 32 * (warp % 4).
    // A operand of the group: 16 KB after the limb matrix, this warp's 32 blocks = 4 KB of it.  Layout (UMMA K-major,
    // no swizzle): block b, pixel row y, column x -> (b/8)*1024 + y*128 + (b%8)*16 + 2x, i.e. a core matrix (8 blocks x
    // 8 pixels of one row) is 128 contiguous bytes, LBO (next pixel row) = 128, SBO (next 8 blocks) = 1024.
    const int group = warp >> 2, quad = warp & 3;
    const uint32_t tmem_d = TC ? s_tmem + (uint32_t)(group * TC_GROUP_COLS) : 0u;
    const uint32_t tmem_lane = (uint32_t)(quad * 32) << 16;
    uint8_t *a_warp = smem + K1_BMAT_BYTES + group * Cfg::A_TILE_BYTES + quad * 4096;
    uint64_t *mma_bar = &s_bar[WARPS + 1 + group];
    uint32_t mma_phase = 0;

    uint32_t nflag = 0;
    const uint32_t xforce = exact_mode ? 0x01010101u : 0u;
    float magic;                                     // 1.5 * 2^23 held in a register: leaves the FFMA's
    asm("mov.f32 %0, 0f4B400000;" : "=f"(magic));     // constant-bank slot to the quantizer constant

#ifdef JPEGB200_TRACE
    uint32_t trace_it = 0xFFFFFFFFu;
#endif
    for (; tile_s0 < total; tile_s0 += stride, s += stride) {
#ifdef JPEGB200_TRACE
        ++trace_it;
#endif
        const bool valid = s < total;                            // TC: the last tile may have idle warps
        StripCtx me = cur;
        uint32_t yw[16];                                         // the lane's 64 luma bytes (TC)
        uint32_t absdev = 0, absmean = 0, sumy = 0;              // A = sum |Y - 128|, Ac = sum |Y - mean|, sum Y
        float x00 = 0.0f;                                        // sum (Y - 128), exact
        const bool fast = cur.npx == 256 && (cur.mispack & 0x77777777u) == 0u;   // (warp-uniform)
        if (valid) {
            K1_TRACE_TILE(0);
            mbar_wait(bar, phase);
            phase ^= 1u;
            K1_TRACE(2);
            K1_TRACE_TILE(1);

            // ---- luma pass -> 256 x 8 Y tile ----------------------------------------------------
            if (fast) {
                // full strip, 8-byte aligned rows: lane = block.  Per row 24 bytes (three LDS.64) -> 8 DP4A (two of the
                // three inputs re-aligned by PRMT; accumulator 0x64000000, so byte 3 of every sum is the fp16 exponent
                // of 1024 and byte 1 is Y) -> 4 PRMT to fp16 pairs 1024 + Y -> 4 HADD2 (-1152) -> one STS.128 into the
                // A operand; the Y bytes for the statistics and the re-evaluation path cost 2 more PRMT.  Two rows per
                // trip, not fully unrolled: the kernel's hot loop has to stay inside the 32 KB instruction cache (L1.5)
#pragma unroll 2
                for (int r = 0; r < 8; ++r) {
                    const uint2 *rw = reinterpret_cast<const uint2 *>(raw + min(r, cur.rmax) * raw_pitch + ((cur.mispack >> (4 * r)) & 8u)) + 3 * lane;
                    const uint2 q0 = rw[0], q1 = rw[1], q2 = rw[2];
                    const uint32_t p01 = luma2_half(q0.x, __byte_perm(q0.x, q0.y, 0x0543u), g.wt_lo, g.wt_lo);
                    const uint32_t p23 = luma2_half(__byte_perm(q0.y, q1.x, 0x0432u), q1.x, g.wt_lo, g.wt_hi);
                    const uint32_t p45 = luma2_half(q1.y, __byte_perm(q1.y, q2.x, 0x0543u), g.wt_lo, g.wt_lo);
                    const uint32_t p67 = luma2_half(__byte_perm(q2.x, q2.y, 0x0432u), q2.y, g.wt_lo, g.wt_hi);
                    const uint2 yv = make_uint2(__byte_perm(p01, p23, 0x6420u), __byte_perm(p45, p67, 0x6420u));
                    *reinterpret_cast<uint2 *>(ybuf + r * Y_PITCH + 8 * lane) = yv;
                    if (TC) {
                        absdev = __vsadu4(yv.x, 0x80808080u) + absdev;
                        absdev = __vsadu4(yv.y, 0x80808080u) + absdev;
                        sumy = __vsadu4(yv.x, 0u) + sumy;
                        sumy = __vsadu4(yv.y, 0u) + sumy;
                        *reinterpret_cast<uint4 *>(a_warp + ((lane >> 3) << 10) + (r << 7) + ((lane & 7) << 4)) =
                            make_uint4(half2_sub1152(p01), half2_sub1152(p23), half2_sub1152(p45), half2_sub1152(p67));
                    }
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
                        const uint32_t yv = luma4(__funnelshift_r(q0, q1, sh), __funnelshift_r(q1, q2, sh), __funnelshift_r(q2, q3, sh), g.wt_lo, g.wt_hi);
                        reinterpret_cast<uint32_t *>(ybuf + r * Y_PITCH)[gc] = yv;
                        if (TC)
                            *reinterpret_cast<uint2 *>(a_warp + ((gc >> 4) << 10) + (r << 7) + (((gc >> 1) & 7) << 4) + ((gc & 1) << 3)) = y4_to_half4(yv);
                    }
                }
            }
            __syncwarp();
            K1_TRACE_TILE(2);

            // raw tile is free: prefetch the next strip while this one is transformed (TC: right after the arrival below,
            // so that the group's MMA is not held up by the address arithmetic)
            if (!TC && s + stride < total) {
                strip_advance(g, pos, dq, dr);
                cur = strip_ctx(g, pos);
                strip_issue_loads(cur, raw, bar, lane, tmap);
            }

            // right-edge replication inside the last real block (converter.c:36)
            const int padpx = me.vb * 8 - me.npx;
            if (padpx > 0) {
                if (lane < padpx) {
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        const uint8_t yv = ybuf[r * Y_PITCH + me.npx - 1];
                        ybuf[r * Y_PITCH + me.npx + lane] = yv;
                        if (TC) {
                            const int px = me.npx + lane, b = px >> 3;
                            *reinterpret_cast<__half *>(a_warp + ((b >> 3) << 10) + (r << 7) + ((b & 7) << 4) + ((px & 7) << 1)) = __int2half_rn((int)yv - 128);
                        }
                    }
                }
                __syncwarp();
            }
        }
        const uint8_t *yblk = ybuf + lane * 8;

        uint32_t zw[16], xw[16];                                 // quantized coefficients (zig-zag, int8) / bracket disagreement
        if (TC) {
            if (valid) fence_proxy_async();                      // the luma pass wrote this warp's quarter of A (generic proxy -> MMA)
            K1_TRACE_TILE2(0);
            tc_fence_before_sync();
            // The tile's four A quarters are in shared memory (and the group's loads of the previous accumulator are
            // done) once all four warps have passed this point.  No warp waits here: each bumps the group's arrival
            // counter (release), and the one that finds the other three already there (acquire) issues the MMA.
            __syncwarp();
            K1_TRACE_TILE2(1);
            uint32_t arrived = 0;
            if (lane == 0)
                asm volatile("atom.acq_rel.cta.shared::cta.inc.u32 %0, [%1], 3;" : "=r"(arrived) : "r"(smem_u32(&s_arrive[group])) : "memory");
            K1_TRACE_TILE2(2);
            if (arrived == 3u) {
                mbar_wait(table_bar, 0);                       // the limb matrix (immediate after the first tile)
                tc_fence_after_sync();
                const uint32_t b_sa = smem_u32(smem), a_sa = b_sa + K1_BMAT_BYTES + (uint32_t)(group * Cfg::A_TILE_BYTES);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)                   // K = 64 in four steps of 16 (two pixel rows); B: 4 KB per step, LBO 128, SBO 256
                    umma_f16_ss(tmem_d, umma_desc(a_sa + 256u * ks, 128u, 1024u), umma_desc(b_sa + 4096u * ks, 128u, 256u), TC_IDESC, ks > 0 ? 1u : 0u);
                umma_commit(mma_bar);
            }
            K1_TRACE_TILE2(3);
            if (valid && s + stride < total) {                   // prefetch: the fence above also covers the reads of the raw tile
                strip_advance(g, pos, dq, dr);
                cur = strip_ctx(g, pos);
                strip_issue_loads(cur, raw, bar, lane, tmap, /*fenced=*/true);
            }
            K1_TRACE_TILE(3);
            if (valid) {
                // block statistics from the Y bytes while the MMA runs: A = sum |Y - 128| and sum Y (the full-strip luma
                // pass has accumulated them already), Ac = sum |Y - mean|
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    const uint2 v = *reinterpret_cast<const uint2 *>(yblk + r * Y_PITCH);
                    yw[2 * r] = v.x;
                    yw[2 * r + 1] = v.y;
                }
                if (!fast) {
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        absdev = __vsadu4(yw[j], 0x80808080u) + absdev;
                        sumy = __vsadu4(yw[j], 0u) + sumy;
                    }
                }
                x00 = (float)((int)sumy - 8192);
                {
                    const int mean = (int)rintf(x00 * 0.015625f) + 128;
                    const uint32_t m4 = (uint32_t)mean * 0x01010101u;
#pragma unroll
                    for (int j = 0; j < 16; ++j) absmean = __vsadu4(yw[j], m4) + absmean;
                }
                K1_TRACE_TILE(4);
                mbar_wait(mma_bar, mma_phase);
                tc_fence_after_sync();
                K1_TRACE_TILE(5);
                // guard half-width in fixed-point units
                const float eb = fminf((float)absdev * kTcGamma, fmaf((float)absdev, kTcGammaA, fmaf((float)absmean, kTcGammaC, kTcGamma0)));
                const f32x2 e2 = pack2(eb, eb), k2048 = pack2(2048.0f, 2048.0f), m2 = pack2(magic, magic);
#pragma unroll
                for (int c = 0; c < 4; ++c) {                    // 16 coefficients = 4 zig-zag words per TMEM load
                    uint32_t r[32];
                    JB_TMEM_LD32(tmem_d + tmem_lane + 32u * c, r);
                    tmem_wait_ld();
                    if (c == 0) K1_TRACE_TILE2(4);
                    if (c == 1) K1_TRACE_TILE2(5);
#pragma unroll
                    for (int wi = 0; wi < 4; ++wi) {
                        // columns 8wi..: S0(a) S0(b) S1(a) S1(b) S0(c) S0(d) S1(c) S1(d), (a..d) = zig-zag positions 4w..4w+3
                        uint32_t hb[4], lb[4];
#pragma unroll
                        for (int p = 0; p < 2; ++p) {
                            const f32x2 s0 = pack2u(r[8 * wi + 4 * p], r[8 * wi + 4 * p + 1]);
                            const f32x2 s1 = pack2u(r[8 * wi + 4 * p + 2], r[8 * wi + 4 * p + 3]);
                            const f32x2 t = fma2(s0, k2048, s1);                       // the 2^21-scaled sum of products
                            const f32x2 rk = *reinterpret_cast<const f32x2 *>(&c_rk_tc[8 * c + 2 * wi + p]);
                            unpack2u(fma2(add2(t, e2), rk, m2), hb[2 * p], hb[2 * p + 1]);
                            unpack2u(fma2(sub2(t, e2), rk, m2), lb[2 * p], lb[2 * p + 1]);
                        }
                        const uint32_t h = __byte_perm(__byte_perm(hb[0], hb[1], 0x0040u), __byte_perm(hb[2], hb[3], 0x0040u), 0x5410u);
                        const uint32_t l = __byte_perm(__byte_perm(lb[0], lb[1], 0x0040u), __byte_perm(lb[2], lb[3], 0x0040u), 0x5410u);
                        zw[4 * c + wi] = h;
                        xw[4 * c + wi] = (h ^ l) | xforce;
                    }
                }
                tc_fence_before_sync();                          // orders these loads before the next tile's MMA (through the arrival counter)
                K1_TRACE_TILE(6);
            }
            mma_phase ^= 1u;
        } else if (valid) {
            float x[8][8];
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
            x00 = x[0][0];
            {
                const int mean = (int)rintf(x00 * 0.015625f) + 128;
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
#pragma unroll
            for (int w = 0; w < 16; ++w) {
                uint32_t hb[4], lb[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    constexpr uint8_t ZZ[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                                12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                                35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                                58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
                    const int zpos = ZZ[4 * w + j], u = zpos >> 3, v = zpos & 7;
                    const float t = x[u][v];
                    const float e = ecls[gclass(u)][gclass(v)];
                    const float rk = c_rk[zpos];
                    hb[j] = __float_as_uint(fmaf(__fadd_rn(t, e), rk, magic));
                    lb[j] = __float_as_uint(fmaf(__fadd_rn(t, -e), rk, magic));
                }
                const uint32_t h = __byte_perm(__byte_perm(hb[0], hb[1], 0x0040u), __byte_perm(hb[2], hb[3], 0x0040u), 0x5410u);
                const uint32_t l = __byte_perm(__byte_perm(lb[0], lb[1], 0x0040u), __byte_perm(lb[2], lb[3], 0x0040u), 0x5410u);
                zw[w] = h;
                xw[w] = (h ^ l) | xforce;
            }
        }
        if (!valid) continue;                                    // (TC only) idle warp of the last tile: barriers done

        uint32_t anyx = 0;
        if (lane < me.vb) {
            // DC: exact integer sum in all formulations; reference operation sequence
            // fl(k00 * S) / 16 then roundf (dct.c:93, quantization.c:36)
            const float f = __fmul_rn(c_ref_scale[0], x00);
            const int dcq = (int)roundf(f * 0.0625f);
            zw[0] = (zw[0] & 0xFFFFFF00u) | ((uint32_t)dcq & 0xFFu);
            xw[0] &= 0xFFFFFF00u;
#pragma unroll
            for (int w = 0; w < 16; ++w) anyx |= xw[w];
        }
        // rare (about 1 block in 200): the bracket straddles a rounding boundary for some coefficient(s); those are
        // re-evaluated in the reference's operation order, one after the other, by the whole warp
#ifdef JB_EXPERIMENT_NOFLAG      // timing experiment only (wrong results): how much do the re-evaluations cost?
        anyx = 0;
#endif
        uint32_t pending = __ballot_sync(0xffffffffu, anyx != 0u);
        while (pending) {
            const int owner = __ffs((int)pending) - 1;
            pending &= pending - 1;
            uint32_t fm_lo = 0, fm_hi = 0;                       // the owner's 64-bit mask of flagged positions
            if (lane == owner) {
#pragma unroll
                for (int w = 0; w < 8; ++w) {
                    fm_lo |= nonzero_nibble(xw[w]) << (4 * w);
                    fm_hi |= nonzero_nibble(xw[w + 8]) << (4 * w);
                }
            }
            fm_lo = __shfl_sync(0xffffffffu, fm_lo, owner);
            fm_hi = __shfl_sync(0xffffffffu, fm_hi, owner);
            const uint8_t *oblk = ybuf + owner * 8;
#pragma unroll 1
            while (fm_lo | fm_hi) {
                const int k = fm_lo ? __ffs((int)fm_lo) - 1 : 32 + __ffs((int)fm_hi) - 1;
                if (fm_lo) fm_lo &= fm_lo - 1;
                else fm_hi &= fm_hi - 1;
                const int zpos = c_zigzag[k];
                const uint32_t q = (uint32_t)exact_quantized_coop(oblk, zpos >> 3, zpos & 7, s_cos, s_scratch[warp], lane) & 0xFFu;
                if (lane == owner) {
                    const uint32_t sh = 8u * (uint32_t)(k & 3), keep = ~(0xFFu << sh), ins = q << sh;
                    const int kw = k >> 2;
#pragma unroll
                    for (int w = 0; w < 16; ++w)
                        if (w == kw) zw[w] = (zw[w] & keep) | ins;
                    ++nflag;
                }
            }
        }
        // Coefficients 0..31 always (32 bytes per lane, 1 KB per warp, contiguous); 32..63 into their own plane and only
        // for the blocks that have any (one block in 500 at amp=20, one in four on a busy photograph): the strip's
        // mask tells the entropy kernel which lanes to read them for.
        {
            const uint32_t hi_any = (zw[8] | zw[9] | zw[10] | zw[11] | zw[12] | zw[13] | zw[14] | zw[15]) != 0u && lane < me.vb;
            const uint32_t mask = __ballot_sync(0xffffffffu, hi_any);
            if (lane == 0) himask[s] = mask;
            if (lane < me.vb) {
                uint4 *dst = reinterpret_cast<uint4 *>(coef + (me.block0 + (uint32_t)lane) * 32);
#ifdef JB_EXPERIMENT_NOSTORE
                if (zw[5] == 0x12345678u)
#endif
                {
                    dst[0] = make_uint4(zw[0], zw[1], zw[2], zw[3]);
                    dst[1] = make_uint4(zw[4], zw[5], zw[6], zw[7]);
                }
                if (hi_any) {
                    // second plane: after the first one's 32 bytes for every block of the launch
                    uint4 *dsth = reinterpret_cast<uint4 *>(coef + (g.blocks_per_image * (uint64_t)g.count + me.block0 + (uint32_t)lane) * 32);
                    dsth[0] = make_uint4(zw[8], zw[9], zw[10], zw[11]);
                    dsth[1] = make_uint4(zw[12], zw[13], zw[14], zw[15]);
                }
            }
        }
        __syncwarp();                                            // the Y tile is rewritten by the next strip's luma pass
        K1_TRACE_TILE(7);
#ifdef JPEGB200_TRACE
        // second half of the trace buffer: completion times of this warp's first 8 strips
        if (trace && lane == 0) {
            const uint32_t it = (s - (blockIdx.x * WARPS + warp)) / stride;
            if (it < 8) trace[(uint64_t)(stride + blockIdx.x * WARPS + warp) * 8 + it] = globaltimer_ns();
        }
#endif
    }

    K1_TRACE(6);
#ifdef JPEGB200_TRACE
    {   // slot 7: coefficients this warp re-evaluated on the exact path
        uint32_t nf = nflag;
#pragma unroll
        for (int ofs = 16; ofs > 0; ofs >>= 1) nf += __shfl_xor_sync(0xffffffffu, nf, ofs);
        if (trace && lane == 0) trace[(uint64_t)(blockIdx.x * WARPS + warp) * 8 + 7] = nf;
    }
#endif
#undef K1_TRACE
#undef K1_TRACE_TILE
#undef K1_TRACE_TILE2
    if (flagged_counter) {
        nflag += __shfl_xor_sync(0xffffffffu, nflag, 16);
        nflag += __shfl_xor_sync(0xffffffffu, nflag, 8);
        nflag += __shfl_xor_sync(0xffffffffu, nflag, 4);
        nflag += __shfl_xor_sync(0xffffffffu, nflag, 2);
        nflag += __shfl_xor_sync(0xffffffffu, nflag, 1);
        if (lane == 0 && nflag) atomicAdd(flagged_counter, (unsigned long long)nflag);
    }
    if (TC) {
        tc_fence_before_sync();
        __syncthreads();
        if (warp == 0) tmem_dealloc(s_tmem, Cfg::TMEM_COLS);
    }
}

}  // namespace jb
