// common.cuh -- shared device declarations for libjpegb200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace jb {

// ---- constant tables (filled once per device by upload_tables; this header is
// included by exactly one translation unit, jpegb200.cu) -----------------------
// Reference 6-decimal cosine LUT [spatial][frequency] (natural_c/src/core/dct.c:9-18)
// and C() factors (dct.c:4-6): used ONLY by the exact (reference-order) evaluation.
__constant__ float c_ref_cos[64];
__constant__ float c_ref_scale[64];   // fl(fl(0.25f*C[u])*C[v])            (dct.c:93)
__constant__ float c_quant_f[64];     // (float)std_luminance_quant_tbl[i]   (quantization.c:35)
// Fast path: q' = T[u][v] * c_rk[u*8+v], with T the scaled butterfly output.
__constant__ float c_rk[64];
// Tensor-core path: q' = t * c_rk_tc[k] with t the 2^21-scaled fixed-point sum of zig-zag position k
// (two consecutive positions are read as one packed-fp32 operand)
__constant__ float2 c_rk_tc[32];
__constant__ uint8_t c_zigzag[64];    // raster index of zig-zag position k  (zigzag.c:7-15)
__constant__ uint8_t c_dc_len[16];    // DC Huffman code length + size, per size class
__constant__ uint32_t c_dc_code[16];  // (code << 8) | len, per size class
__constant__ uint32_t c_ac_code[256]; // (code << 8) | len, per (run<<4|size) symbol

// Guard half-width of the quantization bracket, in units of the un-scaled DCT sum s:
//   |s_ref - g*T|  <=  min( kGamma * A ,  kGammaA * A + kGammaC * Ac + kGamma0 )
// A = sum|p| over the block, Ac = sum|p - c| with c = round(mean p).  The second form separates the part of
// the reference's rounding error that scales with the block's DC level (only its running partial sums see
// it) from everything that only sees the deviations from the mean.  With u = 2^-24:
//   kGamma = 6.5e-6 (>= 102.8 u), kGammaA >= 16.97 u, kGammaC >= 100.67 u, kGamma0 >= 479 u.
constexpr float kGamma = 6.5e-6f;
constexpr float kGammaA = 1.0431e-6f;     // 17.5 u
constexpr float kGammaC = 6.1394e-6f;     // 103 u
constexpr float kGamma0 = 2.9803e-5f;     // 500 u
constexpr float kMagic = 12582912.0f;        // 1.5 * 2^23: fmaf(x, r, kMagic) rounds x*r to nearest-even integer

// Tensor-core transform (fused_block.cuh, TC = true): the 64-term sums are taken over the REFERENCE's own LUT
// products, quantized to 22-bit fixed point (|W/2^21 - w| <= 2^-22 = 4u) and accumulated exactly, so the
// "LUT vs ideal cosine" and "butterfly rounding" terms of the bound disappear:
//   |s_ref - t/2^21|  <=  min( kTcGamma * A ,  kTcGammaA * A + kTcGammaC * Ac + kTcGamma0 )
//   kTcGamma >= 76.2 u,  kTcGammaA >= 15.4 u,  kTcGammaC >= 74.2 u,  kTcGamma0 >= 430 u
// The constants below are already multiplied by 2^21 (the kernel works in fixed-point units).
constexpr float kTcScale = 2097152.0f;        // 2^21
constexpr float kTcGamma = 4.5896e-6f * kTcScale;      // 77 u
constexpr float kTcGammaA = 9.2387e-7f * kTcScale;     // 15.5 u
constexpr float kTcGammaC = 4.4703e-6f * kTcScale;     // 75 u
constexpr float kTcGamma0 = 2.6226e-5f * kTcScale;     // 440 u

// one record per 32-block strip, written by K1
struct __align__(8) StripRec {
    uint32_t bits;      // bits of the strip's entropy-coded stream (an image's first strip: without its first DC symbol)
    int16_t first_dc;   // quantized DC of the strip's first block
    int16_t last_dc;    // quantized DC of the strip's last block
};

// device table block (one allocation per encoder)
constexpr int TBL_SYM = 0;                       // uint32 [16][256] ready-made AC symbols (see build_tables)
constexpr int TBL_DC_CODE = TBL_SYM + 16384;     // uint32 [16]   (code << 8) | len per DC size class
constexpr int TBL_BMAT = TBL_DC_CODE + 128;      // fp16 [128 x 64] limb matrix of the tensor-core DCT, UMMA K-major layout
constexpr int TBL_BYTES = TBL_BMAT + 16384;

// error word bits (device -> host)
enum : uint32_t {
    ERRBIT_WORKSPACE = 1u << 0,
    ERRBIT_OUTPUT = 1u << 1,
    ERRBIT_LOOKBACK = 1u << 2,
};

// geometry of one launch: `count` images of w x h, bw x bh blocks each
struct Geom {
    const uint8_t *rgb;      // first pixel of the TOP row of image 0
    uint64_t image_stride;
    int64_t row_pitch;       // bytes from one pixel row to the next one below it: 3*w for BMPImage.data; negative
                             // (and padded to 4) when the rows of a bottom-up BMP file are read in place
    uint32_t wt_lo, wt_hi;   // DP4A luma weights for a pixel in bytes 0..2 / 1..3 of a word: (77,150,29) in the
                             // pixel's channel order (RGB for BMPImage.data, BGR for raw BMP rows)
    int w, h, bw, bh;
    int spr;                 // 32-block strips per block row
    int count;
    uint64_t blocks_per_image;
    uint64_t total_strips;
    int use_tmap;            // 1: pixel tiles are fetched with one tensor-map TMA copy per strip (16-byte aligned
                             // base / pitch / stride); 0: one bulk copy per pixel row (any alignment)
};

__device__ __forceinline__ uint32_t smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}

__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gmem_src)      // both 8-byte aligned
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
// the executing thread arrives on the mbarrier once all its earlier cp.async copies have landed (the
// barrier's arrival count has to include these arrivals)
__device__ __forceinline__ void cp_async_mbar_arrive_noinc(uint64_t *bar)
{
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMA bulk copies (cp.async.bulk -> UBLKCP) completed through an mbarrier ------------------
// nanosecond wall clock shared by all SMs (tuning aids only)
__device__ __forceinline__ unsigned long long globaltimer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t arrivals)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(arrivals) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory"); }
// order this thread's earlier generic-proxy accesses to shared memory before later async-proxy (TMA) writes
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared bulk copy; src and dst 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_g2s(void *smem_dst, const void *gmem_src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// one box of a 3-D tensor map (x: 32-bit words in a row, y: pixel rows, z: images) -> shared memory (128-byte aligned)
__device__ __forceinline__ void tensor_g2s_3d(void *smem_dst, const void *tmap, int x, int y, int z, uint64_t *bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
            smem_u32(smem_dst)),
        "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "MBAR_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra MBAR_DONE;\n"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// 4-bit mask of the non-zero bytes of a word (bit j <-> byte j)
__device__ __forceinline__ uint32_t nonzero_nibble(uint32_t w)
{
    const uint32_t t = (((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) & 0x80808080u;   // 0x80 per non-zero byte
    return (t * 0x00204081u) >> 28;                                             // gather bits 7,15,23,31
}

// ---- tcgen05 / tensor memory (5th-generation tensor cores) ---------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t *smem_result, uint32_t columns)      // one whole warp
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_result)), "r"(columns) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t columns)           // the allocating warp
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "r"(columns) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }
// shared-memory matrix descriptor, K-major, no swizzle: core matrices (8 rows x 16 bytes, contiguous) LBO bytes apart
// along K and SBO bytes apart along M/N; descriptor version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           (1ull << 46);
}
// D[tmem] (+)= A[tmem] * B[smem]^T, fp16 operands, fp32 accumulate; issued by ONE thread
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T
__device__ __forceinline__ void umma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate)
{
    asm volatile(
        "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// the mbarrier receives one arrival once all MMAs issued so far by this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t *bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
#define JB_TMEM_ST32(taddr, r)                                                                                         \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16," \
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),                   \
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),   \
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),       \
                 "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),      \
                 "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                   \
                 : "memory")
#define JB_TMEM_LD32(taddr, r)                                                                                         \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"       \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"                          \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),     \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),           \
                   "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),         \
                   "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),         \
                   "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                              \
                 : "r"(taddr)                                                                                          \
                 : "memory")

// decoupled look-back tile state: [63:42] epoch, [41:40] status, [39:0] value
constexpr uint64_t LB_VALUE_MASK = (1ull << 40) - 1;
constexpr int LB_STATUS_SHIFT = 40;
constexpr int LB_EPOCH_SHIFT = 42;
constexpr uint32_t LB_AGGREGATE = 1, LB_PREFIX = 2;

__device__ __forceinline__ uint64_t lb_pack(uint32_t epoch, uint32_t status, uint64_t value)
{
    return ((uint64_t)epoch << LB_EPOCH_SHIFT) | ((uint64_t)status << LB_STATUS_SHIFT) | (value & LB_VALUE_MASK);
}

__device__ __forceinline__ uint64_t ld_volatile_u64(const uint64_t *p)
{
    uint64_t v;
    asm volatile("ld.volatile.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_volatile_u64(uint64_t *p, uint64_t v)
{
    asm volatile("st.volatile.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}

// Exclusive prefix of the aggregates of tiles [0, tile) of one segment, executed by
// ONE WARP (all 32 lanes).  `state` points at the segment's tile 0.  Spins are bounded;
// on timeout sets ERRBIT_LOOKBACK and returns what it has (never hangs the GPU).
__device__ __forceinline__ uint64_t lookback_exclusive(const uint64_t *state, int tile, uint32_t epoch,
                                                       uint32_t *err)
{
    const int lane = threadIdx.x & 31;
    uint64_t excl = 0;
    int base = tile - 1;
    while (base >= 0) {
        const int j = base - lane;
        uint64_t v = 0;
        uint32_t status = LB_PREFIX;           // virtual tile before tile 0: prefix 0
        if (j >= 0) {
            int spins = 0;
            for (;;) {
                v = ld_volatile_u64(state + j);
                status = (uint32_t)(v >> LB_STATUS_SHIFT) & 3u;
                if ((uint32_t)(v >> LB_EPOCH_SHIFT) == epoch && status != 0) break;
                if (++spins > (1 << 22)) { atomicOr(err, ERRBIT_LOOKBACK); status = LB_PREFIX; v = 0; break; }
                __nanosleep(20);
            }
            v &= LB_VALUE_MASK;
        }
        const uint32_t pmask = __ballot_sync(0xffffffffu, status == LB_PREFIX);
        const int first = pmask ? (__ffs(pmask) - 1) : 31;   // nearest tile holding an inclusive prefix
        uint64_t contrib = (lane <= first) ? v : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
        excl += contrib;
        if (pmask) break;
        base -= 32;
    }
    return excl;
}

// Grouped look-back (whole CTA).  Tiles of one segment publish their aggregate in agg[tile]
// (bit 63 = valid) as soon as they know it; the last tile of every group of LB_GROUP tiles also
// publishes its inclusive prefix in incl[group].  A tile's exclusive prefix is then
//     incl[group - 1]  +  sum of agg[j] for the earlier tiles j of its own group,
// and the (up to 1023) aggregate loads are independent: every thread issues all of its loads
// back to back, so the whole look-back costs about one L2 round trip instead of a polling chain
// (no thread ever re-reads a word that was already valid).  Only group boundaries form a serial
// chain, one hop per 1024 tiles.  Spins are bounded; on timeout ERRBIT_LOOKBACK is set.
constexpr int LB_GROUP = 1024;
constexpr uint64_t LB_VALID = 1ull << 63;

__device__ __forceinline__ uint64_t lb_wait(const uint64_t *p, uint32_t *err)
{
    uint64_t v = ld_volatile_u64(p);
    int spins = 0;
    while (!(v & LB_VALID)) {
        if (++spins > (1 << 22)) { atomicOr(err, ERRBIT_LOOKBACK); return 0; }
        __nanosleep(100);
        v = ld_volatile_u64(p);
    }
    return v & ~LB_VALID;
}

// s_scratch: 9 words of shared memory.  Must be called by the whole CTA (blockDim.x <= 256).
__device__ __forceinline__ uint64_t lookback_grouped(const uint64_t *agg, const uint64_t *incl, int tile, uint32_t *err,
                                                     uint64_t *s_scratch)
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthreads = blockDim.x, nwarp = nthreads >> 5;
    const int g0 = (tile / LB_GROUP) * LB_GROUP;
    uint32_t sum = 0;                      // aggregates of one group fit 32 bits; the checkpoint is added as 64 bits
    uint64_t base = 0;
    if (tid == 0 && g0 > 0) base = lb_wait(incl + (tile / LB_GROUP) - 1, err);
    constexpr int BATCH = 4;                 // 4 x 256 threads = one whole group per pass
    for (int j0 = g0 + tid; j0 < tile; j0 += nthreads * BATCH) {
        uint64_t v[BATCH];
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
            const int j = j0 + i * nthreads;
            v[i] = j < tile ? ld_volatile_u64(agg + j) : LB_VALID;
        }
#pragma unroll
        for (int i = 0; i < BATCH; ++i) {
            const int j = j0 + i * nthreads;
            if (!(v[i] & LB_VALID)) v[i] = lb_wait(agg + j, err);
            sum += (uint32_t)v[i];
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    __syncthreads();                       // s_scratch may still be read from a previous call
    if (lane == 0) s_scratch[warp] = sum;
    if (tid == 0) s_scratch[8] = base;
    __syncthreads();
    uint64_t excl = s_scratch[8];
    for (int w = 0; w < nwarp; ++w) excl += s_scratch[w];
    return excl;
}

}  // namespace jb
