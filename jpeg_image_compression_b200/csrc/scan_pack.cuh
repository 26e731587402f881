// scan_pack.cuh -- K2, the fused entropy kernel: bit-offset scan + Huffman bit packing +
// 0xFF byte stuffing in ONE launch (plus the two small batch-mode helpers).
//
// The reference's entropy stage is one serial chain: DC prediction across all blocks
// (rle.c:59-70) and one contiguous MSB-first bit stream (huffman.c:35-62) with a zero byte
// stuffed after every 0xFF (huffman.c:26-32) and a zero-padded last byte (huffman.c:65-81).
//
// K1 leaves, per 32-block strip, a record {bits, first DC, last DC} (the bit count already includes
// every DC-difference symbol) and per block its bit offset inside the strip, so everything
// cross-block collapses to two prefix sums:
//
//   tile = 8 consecutive strips (<= 256 blocks), one CTA, warp w <-> strip, lane <-> block
//   1. the lane's coefficients, and the bit counts of the earlier strips, are requested up front; the
//      tile's bit offset is their plain sum (wait-free; one checkpoint word per 1024 tiles);
//   2. every lane walks the non-zero coefficients of its block (one table look-up per symbol) and
//      appends code+amplitude bits to a register accumulator that is OR-reduced word-wise into a
//      zeroed shared-memory window;
//   3. bytes are owned by the tile that holds their first bit: the last, partial byte is
//      completed by encoding the next tile's first block(s) clipped at the byte boundary, the
//      first partial byte is skipped -- no bits ever cross CTAs through global memory;
//   4. 0xFF bytes of the tile's byte range are counted and published; a look-back over the earlier
//      tiles' counts gives the number of stuffed zeros before the tile, and the stuffed bytes go
//      straight to the output.  A persistent CTA resolves that look-back one tile late (after
//      packing its next tile into a second window), when it no longer has to wait.
// CTAs take tiles in increasing order: blockIdx.x when all tiles fit in one wave, else by ticket.
#pragma once

#include "common.cuh"

namespace jb {

constexpr int K2_WARPS = 8;                                    // strips per tile
constexpr int K2_THREADS = K2_WARPS * 32;
constexpr int K2_MAX_BLOCK_BITS = 1472;                       // >= 14 + 63*23 = 1463: any tile fits
constexpr int K2_SMALL_BLOCK_BITS = 256;                      // default window: 32 bytes per block (4 bits per pixel) on average
constexpr int k2_win_words(int block_bits) { return (K2_THREADS * block_bits) / 32 + 8; }

// one record per strip, written by K1
struct __align__(8) StripRec {
    uint32_t bits;      // bit cost of the strip's blocks (an image's first strip: without its first DC symbol)
    int16_t first_dc;   // quantized DC of the strip's first block
    int16_t last_dc;    // quantized DC of the strip's last block
};

// per-block word written by K1: [15:0] bit offset inside the strip (same convention as
// StripRec.bits), [21:16] index of the last non-zero AC coefficient (0 = none)
__host__ __device__ __forceinline__ uint32_t blk_pack(uint32_t off, uint32_t last) { return off | (last << 16); }

// device table block (one allocation per encoder): bit-cost LUT for K1, code tables for K2
constexpr int ACLUT_ROWS = 63;                                 // zero run 0..62
constexpr int ACLUT_STRIDE = 260;                              // 256 + 4: rows start in different banks
constexpr int TBL_ACLUT = 0;                                   // uint8  [63][260] + EOB length at [16380]
constexpr int TBL_DC_LEN = 16384;                              // uint8  [16]   DC code length + size  (K1 stages [0, 16400))
constexpr int TBL_AC_CODE = 16400;                             // uint32 [256]  (code << 8) | len per (run<<4|size)
constexpr int TBL_DC_CODE = TBL_AC_CODE + 1024;                // uint32 [16]   (code << 8) | len per size class
constexpr int TBL_SYM = TBL_DC_CODE + 64;                      // uint32 [16][256] ready-made AC symbols, see encode_block
constexpr int TBL_BYTES = TBL_SYM + 16384 + 32;

constexpr int K2_STAGE_STRIDE = 17;                            // words per lane in the coefficient staging area
constexpr int K2_SYM_WORDS = 16 * 256;                         // AC symbol table staged in shared memory
constexpr int k2_smem(int block_bits) { return (2 * k2_win_words(block_bits) + K2_THREADS * K2_STAGE_STRIDE + K2_SYM_WORDS) * 4; }

struct PackArgs {
    const uint8_t *tables;         // device table block (TBL_* offsets)
    const int8_t *coef;            // [count*nb_avail][64] zig-zag int8
    const uint32_t *blkinfo;       // [count*nb_avail]
    const StripRec *strips;        // [count*strips_avail]
    const uint32_t *strip_bits;    // [count*strips_avail] compact copy of StripRec.bits (read four at a time)
    uint64_t *bit_incl;            // bit-offset checkpoints, one per 1024-tile group: [count*groups]
    uint64_t *ff_agg, *ff_incl;    // grouped look-back state of the stuffed-zero counts: [count*tiles], [count*groups]
    unsigned long long *tile_counter;   // next tile to hand out (cleared together with the look-back state)
    int dynamic_tiles;             // 0: one tile per CTA (tile = blockIdx.x); 1: persistent CTAs draw tiles from tile_counter
    uint8_t *out;                  // stuffed bytes: caller's buffer (count==1) or per-image slots (batch)
    uint64_t out_capacity;         // bytes available per image at `out`
    uint64_t out_slot;             // byte distance between images at `out` (batch), 0 for count==1
    uint64_t *image_bytes;         // [count] stuffed size of each image
    uint64_t *image_bits;          // [count] total bits of each image (owned blocks)
    uint64_t *scan_offsets;        // count==1: scan_offsets[0..1] written here; batch: written by k_layout
    uint32_t *err;
    uint32_t strips_owned;         // strips per image that belong to the stream
    uint32_t strips_avail;         // >= strips_owned: strips K1 produced (stripe halo)
    uint32_t spr, bw;              // strips per block row, blocks per block row
    uint32_t nb_avail;             // blocks per image K1 produced
    int tiles;                     // tiles per image (over owned strips)
    int count;
    int16_t dc_pred0;              // DC predictor of the image's first block (0; stripes: previous stripe's last DC)
    uint32_t bit_phase;            // bit offset of the first bit inside byte 0 (0; stripes: global phase & 7)
    unsigned long long *trace;     // optional [tiles*count][8] phase timestamps (ns), tuning aid; nullptr in production
};

__device__ __forceinline__ int magnitude_class(int v)          // rle.c:9-22
{
    const int a = v < 0 ? -v : v;
    return 32 - __clz(a);
}

__device__ __forceinline__ uint32_t strip_blocks(uint32_t strip_in_image, uint32_t spr, uint32_t bw)
{
    const uint32_t sx = strip_in_image % spr;
    return min(32u, bw - sx * 32u);
}

// shared-memory accesses by 32-bit shared-space address: keeps the symbol loop free of the
// generic-to-shared address arithmetic the compiler otherwise repeats at every access
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ uint32_t lds_u8(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void red_or_shared(uint32_t saddr, uint32_t v, bool enable)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p red.shared.or.b32 [%0], %1;\n}"
                 :: "r"(saddr), "r"(v), "r"((uint32_t)enable) : "memory");
}

// MSB-first bit appender with a 32-bit register accumulator, flushed word-wise into the zeroed
// shared-memory window with OR-reductions (the first and last word of a block are shared with its
// neighbours).  Branch-free: the flush is predicated.
struct BitWriter {
    uint32_t waddr;       // shared-space address of the window word being filled
    uint32_t acc, fill;
    __device__ __forceinline__ void start(uint32_t win_saddr, uint32_t relbit)
    {
        waddr = win_saddr + ((relbit >> 5) << 2);
        fill = relbit & 31u;
        acc = 0;
    }
    __device__ __forceinline__ void put(uint32_t vl, uint32_t n)           // n in 1..27 bits, left-aligned in vl
    {
        acc |= vl >> fill;
        fill += n;
        const bool full = fill >= 32u;
        red_or_shared(waddr, acc, full);
        fill &= 31u;
        waddr += full ? 4u : 0u;
        acc = full ? vl << (n - fill) : acc;                               // the bits that did not fit (none if fill == 0)
    }
    __device__ __forceinline__ void finish() { red_or_shared(waddr, acc, fill != 0u); }
};

// append n left-aligned bits with clipping at bit `limit` (window-relative); used only for the halo blocks
__device__ __forceinline__ void put_clipped(uint32_t win_saddr, uint32_t &pos, uint32_t limit, uint32_t vl, uint32_t n)
{
    if (pos >= limit) return;
    if (pos + n > limit) n = limit - pos;
    vl &= 0xFFFFFFFFu << (32u - n);                              // n >= 1
    const uint64_t x = ((uint64_t)vl << 32) >> (pos & 31u);
    const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
    const uint32_t wa = win_saddr + ((pos >> 5) << 2);
    red_or_shared(wa, hi, hi != 0u);
    red_or_shared(wa + 4u, lo, lo != 0u);
    pos += n;
}

// 4-bit mask of the non-zero bytes of a word (bit j <-> byte j)
__device__ __forceinline__ uint32_t nonzero_nibble(uint32_t w)
{
    const uint32_t t = (((w & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | w) & 0x80808080u;   // 0x80 per non-zero byte
    return (t * 0x00204081u) >> 28;                                             // gather bits 7,15,23,31
}

// Walk one block's symbols (rle.c:59-124) and hand (value, nbits) pairs (huffman.c:145-173) to
// emit(), which returns false to stop early.  sw: shared-space address of the block's 16 coefficient
// words (zig-zag order, int8); sym, dc: shared-space addresses of the symbol tables.
// The lane first builds the 63-bit map of its non-zero AC coefficients and then visits only
// those: the loop trip count is the lane's symbol count, so a warp runs max-over-lanes symbols
// instead of one divergent branch per coefficient position.  Each visit is one table look-up:
// sym[run & 15][value & 255] = (Huffman code << size | amplitude bits), left-aligned in the word, with
// the total length in the low 5 bits, i.e. huffman.c:164-173 applied to the symbol rle.c:106-113 would
// have produced.  emit() receives (left-aligned bits, count).
// mlo/mhi: the block's non-zero map (bit k <-> zig-zag position k, DC excluded), built by the caller
// from the coefficient words while they were in registers.
template <typename Emit>
__device__ __forceinline__ void encode_block(uint32_t sw, int my_dc, int prev_dc, int last, uint32_t mlo, uint32_t mhi,
                                             uint32_t sym, uint32_t dc, Emit emit)
{
    {
        const int diff = my_dc - prev_dc;                                              // rle.c:68-70
        const int sz = magnitude_class(diff);
        const uint32_t hc = lds_u32(dc + 4u * (uint32_t)sz);
        const uint32_t amp = (uint32_t)(diff > 0 ? diff : diff - 1) & ((1u << sz) - 1u);   // rle.c:24-35, huffman.c:39
        const uint32_t n = (hc & 0xFFu) + (uint32_t)sz;
        if (!emit((((hc >> 8) << sz) | amp) << (32u - n), n)) return;
    }
    int prev = 0;                                                 // position of the previous non-zero (0 = DC)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t m = half ? mhi : mlo;
#pragma unroll 1
        while (m) {
            const int k = 32 * half + __ffs((int)m) - 1;
            m &= m - 1;
            const uint32_t byte = lds_u8(sw + (uint32_t)k);
            int run = k - prev - 1;
            prev = k;
            if (run >= 16) {                                                           // ZRL, rle.c:99-103
                const uint32_t z = lds_u32(sym + 4u * 0x80u);   // slot (run 0, value -128: cannot occur) holds the ZRL code
                do {
                    if (!emit(z & ~31u, z & 31u)) return;
                    run -= 16;
                } while (run >= 16);
            }
            const uint32_t e = lds_u32(sym + 4u * (((uint32_t)run << 8) | byte));
            if (!emit(e & ~31u, e & 31u)) return;
        }
    }
    if (last < 63) {                                                                   // EOB, rle.c:121-123
        const uint32_t e = lds_u32(sym);                      // slot (run 0, value 0) holds the EOB code
        emit(e & ~31u, e & 31u);
    }
}

__device__ __forceinline__ uint32_t count_ff_bytes(uint32_t w)
{
    uint32_t x = w & (w >> 4);
    x &= x >> 2;
    x &= x >> 1;
    return __popc(x & 0x01010101u);
}

// window word i restricted to the owned window bytes [wb0, wb1) (bytes are MSB first inside a word)
__device__ __forceinline__ uint32_t masked_word(const uint32_t *win, uint32_t i, uint32_t wb0, uint32_t wb1)
{
    uint32_t v = win[i];
    const uint32_t lo = i * 4, hi = lo + 4;
    if (lo < wb0) v &= 0xFFFFFFFFu >> (8 * (wb0 - lo));
    if (hi > wb1) v &= wb1 > lo ? 0xFFFFFFFFu << (8 * (hi - wb1)) : 0u;
    return v;
}

// BLOCK_BITS: window capacity per block.  The default instantiation (256: four bits per pixel on
// average over a tile) keeps shared memory small (4 CTAs per SM); a tile that does not fit raises
// ERRBIT_WORKSPACE and the caller re-runs with the worst-case instantiation (1472), selected through
// jpegb200_encoder_set_bytes_per_block.
// phase timestamps exist only in the tracing build of the library (make trace -> libjpegb200_trace.so)
#ifdef JPEGB200_TRACE
#define K2_TRACE(tile_id, slot) do { if (a.trace && tid == 0) a.trace[(tile_id) * 8 + (slot)] = globaltimer_ns(); } while (0)
#else
#define K2_TRACE(tile_id, slot) do { } while (0)
#endif

// A packed tile whose bytes are not written yet.  A CTA packs tile i+1 before it resolves the
// stuffed-zero look-back of tile i and writes it out: by then every predecessor has published its
// count, so the look-back never waits in steady state (two bit windows, used alternately).
struct PendingTile {
    uint64_t t;          // ticket (img * tiles + tile); ~0 = none
    uint64_t w0;         // stream word index of window word 0
    uint64_t B0, B1;     // owned stream bytes [B0, B1)
    uint32_t tile_ff;    // 0xFF bytes among them
};

template <int BLOCK_BITS>
__global__ void __launch_bounds__(K2_THREADS)
k_scan_pack_stuff(const PackArgs a)
{
    constexpr int WIN_WORDS = k2_win_words(BLOCK_BITS);
    extern __shared__ __align__(16) uint32_t k2_smem_words[];   // two bit windows, the staging area, the symbol table
    uint32_t *stage = k2_smem_words + 2 * WIN_WORDS + (threadIdx.x * K2_STAGE_STRIDE);   // this lane's 16 coefficient words
    uint32_t *s_sym = k2_smem_words + 2 * WIN_WORDS + K2_THREADS * K2_STAGE_STRIDE;
    __shared__ uint32_t s_dc[16];
    __shared__ uint32_t s_strip_base[K2_WARPS];     // bit offset of each strip inside the tile
    __shared__ uint32_t s_warp[K2_WARPS], s_tile_bits;
    __shared__ uint64_t s_scratch[9];
    __shared__ uint32_t s_halo[20];                 // the block after the tile: 16 coefficient words, blkinfo, non-zero map (2), predictor
    __shared__ __align__(8) uint64_t s_bar;         // completion of the symbol table's bulk copy
    __shared__ unsigned long long s_next_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t smem_sa = smem_u32(k2_smem_words);                         // shared-space addresses for the symbol loop
    asm volatile("mov.b32 %0, %0;" : "+r"(smem_sa));                    // opaque: computed once, not rematerialised at every use
    const uint32_t stage_sa = smem_sa + (uint32_t)(2 * WIN_WORDS + tid * K2_STAGE_STRIDE) * 4u;
    const uint32_t sym_sa = smem_sa + (uint32_t)(2 * WIN_WORDS + K2_THREADS * K2_STAGE_STRIDE) * 4u;
    const uint32_t dc_sa = smem_u32(s_dc);
    const uint32_t halo_sa = smem_u32(s_halo);
    if (tid == 0) {
        s_next_tile = a.dynamic_tiles ? atomicAdd(a.tile_counter, 1ull) : (unsigned long long)blockIdx.x;
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&s_bar, K2_SYM_WORDS * 4);                 // 16 KB, one TMA bulk copy; awaited before the first pack
        bulk_g2s(s_sym, a.tables + TBL_SYM, K2_SYM_WORDS * 4, &s_bar);
    }
    if (tid < 16) s_dc[tid] = reinterpret_cast<const uint32_t *>(a.tables + TBL_DC_CODE)[tid];
    bool sym_ready = false;
    const uint64_t origin = ((uint64_t)a.bit_phase + 7) >> 3;     // first stream byte this image/stripe owns
    const int groups = (a.tiles + LB_GROUP - 1) / LB_GROUP;
    const uint64_t total_tiles = (uint64_t)a.tiles * (uint64_t)a.count;

    // Tiles wait on their predecessors (look-backs), so a tile must never be owned by a CTA that is not
    // running yet.  If all tiles fit in one wave the launch has one CTA per tile (tile = blockIdx.x: CTAs
    // are dispatched in index order).  Otherwise persistent CTAs draw tiles from an atomic counter,
    // strictly in increasing order and only once the CTA runs: every predecessor of a tile is then finished
    // or in the hands of a running CTA that publishes its counts without waiting on a later tile -- no
    // deadlock even when only part of the grid is resident (e.g. next to another stream's kernels).
    PendingTile pend;
    pend.t = ~0ull;
    int cur = 0;
    for (;;) {
        __syncthreads();
        const uint64_t t = s_next_tile;
        const bool have_tile = t < total_tiles;
        PendingTile mine;
        mine.t = ~0ull;
        unsigned long long ticket = ~0ull;
        if (have_tile) {
            uint32_t *win = k2_smem_words + cur * WIN_WORDS;
            const uint32_t win_sa = smem_sa + (uint32_t)(cur * WIN_WORDS) * 4u;
            const int img = a.count == 1 ? 0 : (int)((uint32_t)t / (uint32_t)a.tiles);   // tickets fit 32 bits
            const int tile = (int)((uint32_t)t - (uint32_t)img * (uint32_t)a.tiles);
            const StripRec *recs = a.strips + (uint64_t)img * a.strips_avail;
            const uint32_t strip0 = (uint32_t)tile * K2_WARPS;
            const uint32_t nstrips = min((uint32_t)K2_WARPS, a.strips_owned - strip0);
            const bool last_tile = tile == a.tiles - 1;
            uint64_t *bit_incl = a.bit_incl + (uint64_t)img * groups;

            K2_TRACE(t, 0);
            // ---- 0. fetch this lane's block before any waiting (loads do not depend on the offsets) ----
            const uint64_t img_block0 = (uint64_t)img * a.nb_avail;
            const uint32_t my_strip = strip0 + warp;
            const bool have = (uint32_t)warp < nstrips && (uint32_t)lane < strip_blocks(my_strip, a.spr, a.bw);
            uint32_t info = 0, mlo = 0, mhi = 0;                     // non-zero map of the block's AC coefficients
            int my_dc = 0;
            uint4 q[4];
            if (have) {
                const uint32_t brow = my_strip / a.spr, sx = my_strip - brow * a.spr;
                const uint64_t b = img_block0 + (uint64_t)brow * a.bw + sx * 32u + lane;
                const uint4 *src = reinterpret_cast<const uint4 *>(a.coef + b * 64);
#pragma unroll
                for (int i = 0; i < 4; ++i) q[i] = src[i];
                info = a.blkinfo[b];
            }
            int prev_dc = 0;
            if (have && lane == 0) prev_dc = my_strip == 0 ? (int)a.dc_pred0 : (int)recs[my_strip - 1].last_dc;
            // the block that follows the tile in raster order (step 3 needs its first few bits): the last warp
            // requests its 16 coefficient words and its blkinfo now, together with everything else
            uint32_t halo_word = 0;
            if (warp == K2_WARPS - 1 && lane < 17 && strip0 + nstrips < a.strips_avail) {
                const uint32_t hst = strip0 + nstrips, hbrow = hst / a.spr, hsx = hst - hbrow * a.spr;
                const uint64_t hb = img_block0 + (uint64_t)hbrow * a.bw + hsx * 32u;
                halo_word = lane < 16 ? reinterpret_cast<const uint32_t *>(a.coef + hb * 64)[lane] : a.blkinfo[hb];
            }

            // ---- 1a. tile bit offset, wait-free: the loads go out together with the coefficient loads -----
            // K1 left complete per-strip bit counts (only the image's very first DC symbol is missing: its
            // predictor is a run-time argument), so the offset is a plain sum over the earlier strips of the
            // tile's 1024-tile group (read from the compact copy of the counts, four per load) plus the
            // group's checkpoint.
            const uint32_t fix0 = c_dc_len[magnitude_class((int)recs[0].first_dc - (int)a.dc_pred0)];
            const int g0 = (tile / LB_GROUP) * LB_GROUP;
            uint32_t part = 0;                                       // a 1024-tile group holds < 2^29 bits
            {
                const uint64_t base = (uint64_t)img * a.strips_avail;            // index of the image's strip 0
                const uint32_t *sbits = a.strip_bits + base;
                const uint32_t lo = (uint32_t)g0 * K2_WARPS, hi = strip0;
                const uint32_t head = min(hi, lo + ((4u - (uint32_t)((base + lo) & 3u)) & 3u));   // up to 16-byte alignment
                if (lo + tid < head) part += sbits[lo + tid];
                const uint32_t nvec = (hi - head) >> 2, tail = head + 4u * nvec;
                const uint4 *v4 = reinterpret_cast<const uint4 *>(sbits + head);
                for (uint32_t i = tid; i < nvec; i += K2_THREADS) {
                    const uint4 v = v4[i];
                    part += (v.x + v.y) + (v.z + v.w);
                }
                if (tail + tid < hi) part += sbits[tail + tid];
            }
            uint32_t tile_tot = 0;                                   // warp 0: this lane's strip of the tile
            if (warp == 0 && (uint32_t)lane < nstrips) {
                const StripRec rec = recs[strip0 + lane];
                tile_tot = rec.bits;
                if ((uint32_t)lane == nstrips - 1) s_halo[19] = (uint32_t)(int)rec.last_dc;   // predictor of the successor block
            }

            if (have) {
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    stage[4 * i] = q[i].x; stage[4 * i + 1] = q[i].y; stage[4 * i + 2] = q[i].z; stage[4 * i + 3] = q[i].w;
                    const uint32_t m16 = nonzero_nibble(q[i].x) | (nonzero_nibble(q[i].y) << 4) | (nonzero_nibble(q[i].z) << 8) |
                                         (nonzero_nibble(q[i].w) << 12);
                    if (i < 2) mlo |= m16 << (16 * i);
                    else mhi |= m16 << (16 * (i - 2));
                }
                mlo &= ~1u;                                          // position 0 is the DC
                my_dc = (int)(int8_t)(q[0].x & 0xFFu);
            }
            if (warp == K2_WARPS - 1) {                              // park the successor block and its non-zero map
                const uint32_t nib = lane < 16 ? nonzero_nibble(halo_word) : 0u;
                const uint32_t hlo = __reduce_or_sync(0xffffffffu, lane < 8 ? nib << (4 * lane) : 0u) & ~1u;
                const uint32_t hhi = __reduce_or_sync(0xffffffffu, lane >= 8 ? nib << (4 * (lane & 7)) : 0u);
                if (lane < 17) s_halo[lane] = halo_word;
                if (lane == 17) s_halo[17] = hlo;
                if (lane == 18) s_halo[18] = hhi;
            }
            // predictor: previous block in raster order (rle.c:59-70) = the previous lane's block, or the
            // previous strip's last block for lane 0 (fetched above)
            {
                const int up = __shfl_up_sync(0xffffffffu, my_dc, 1);
                if (lane != 0) prev_dc = up;
            }

            K2_TRACE(t, 1);
            // ---- 1b. finish the tile bit offset ----------------------------------------------------------
            if (tid == 0) s_scratch[K2_WARPS] = g0 > 0 ? lb_wait(bit_incl + tile / LB_GROUP - 1, a.err) : (tile > 0 ? fix0 : 0u);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0) s_scratch[warp] = part;
            if (warp == 0) {
                uint32_t tot = tile_tot;
                if (tile == 0 && lane == 0) tot += fix0;
                uint32_t incl = tot;
#pragma unroll
                for (int o = 1; o < K2_WARPS; o <<= 1) {
                    const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += n;
                }
                if (lane == K2_WARPS - 1) s_tile_bits = incl;
                if (lane < K2_WARPS) s_strip_base[lane] = incl - tot;
            }
            __syncthreads();
            uint64_t bit_excl = s_scratch[K2_WARPS];                 // the group's checkpoint (64-bit)
#pragma unroll
            for (int w = 0; w < K2_WARPS; ++w) bit_excl += s_scratch[w];
            const uint64_t begin = bit_excl + a.bit_phase, end = begin + s_tile_bits;
            if (tid == 0) {
                if (tile % LB_GROUP == LB_GROUP - 1) st_volatile_u64(bit_incl + tile / LB_GROUP, LB_VALID | (bit_excl + s_tile_bits));
                if (last_tile) a.image_bits[img] = bit_excl + s_tile_bits;
            }
            K2_TRACE(t, 2);
            const uint64_t w0 = begin >> 5;
            uint32_t nwords = (uint32_t)(((end + 31) >> 5) - w0);
            const bool fits = nwords + 2 <= (uint32_t)WIN_WORDS;
            if (!fits) {                                              // dense tile: needs the large-window instantiation
                if (tid == 0) atomicOr(a.err, ERRBIT_WORKSPACE);
                nwords = 0;
            }
            for (uint32_t i = tid; i < nwords + 2; i += K2_THREADS) win[i] = 0;
            if (!sym_ready) {
                mbar_wait(&s_bar, 0);
                sym_ready = true;
            }
            __syncthreads();

            K2_TRACE(t, 3);
            // ---- 2. pack this tile's blocks ----------------------------------------------------------
            if (have && fits) {
                // K1's strip-local offsets are complete except in the image's first strip (fix0)
                const uint64_t off = begin + s_strip_base[warp] + (info & 0xFFFFu) + (my_strip == 0 && lane ? fix0 : 0u);
                BitWriter bw;
                bw.start(win_sa, (uint32_t)(off - (w0 << 5)));
                encode_block(stage_sa, my_dc, prev_dc, (int)((info >> 16) & 63u), mlo, mhi, sym_sa, dc_sa,
                             [&](uint32_t v, uint32_t n) { bw.put(v, n); return true; });
                bw.finish();
            }

            // ---- 3. complete the last owned byte with the next tile's leading bits (at most 7) ---------
            // Bytes are owned by the tile that holds their first bit.  Runs concurrently with the packing
            // above: it only ORs into bits at or after `end`.
            const uint64_t limit = (end + 7) & ~7ull;                // first bit NOT owned by this tile
            if (tid == 0 && fits && (end & 7u) && !(last_tile && a.strips_avail == a.strips_owned)) {
                uint32_t pos = (uint32_t)(end - (w0 << 5));
                const uint32_t lim = (uint32_t)(limit - (w0 << 5));
                uint32_t st = strip0 + nstrips;                      // raster successor of the tile's last block
                int hprev = (int)s_halo[19];
                auto emit_clipped = [&](uint32_t v, uint32_t nb) {
                    put_clipped(win_sa, pos, lim, v, nb);
                    return pos < lim;
                };
                if (st < a.strips_avail) {
                    // the first successor block was staged in s_halo during step 0
                    const int hdc = (int)(int8_t)(s_halo[0] & 0xFFu);
                    encode_block(halo_sa, hdc, hprev, (int)((s_halo[16] >> 16) & 63u), s_halo[17], s_halo[18], sym_sa, dc_sa,
                                 emit_clipped);
                    hprev = hdc;
                    uint32_t lb = 1;
                    if (lb >= strip_blocks(st, a.spr, a.bw)) { lb = 0; ++st; }
                    // rare: that block was shorter than the missing bits (a block can be as short as 6 bits)
                    if (pos < lim && st < a.strips_avail) {
                        const uint32_t brow = st / a.spr, sx = st - brow * a.spr;
                        const uint64_t b = img_block0 + (uint64_t)brow * a.bw + sx * 32u + lb;
                        const uint32_t *src = reinterpret_cast<const uint32_t *>(a.coef + b * 64);
                        uint32_t hlo = 0, hhi = 0;
#pragma unroll 1
                        for (int i = 0; i < 16; ++i) {
                            const uint32_t q = src[i];
                            stage[i] = q;
                            if (i < 8) hlo |= nonzero_nibble(q) << (4 * i);
                            else hhi |= nonzero_nibble(q) << (4 * (i - 8));
                        }
                        hlo &= ~1u;
                        const uint32_t hinfo = a.blkinfo[b];
                        encode_block(stage_sa, (int)(int8_t)(src[0] & 0xFFu), hprev, (int)((hinfo >> 16) & 63u), hlo, hhi, sym_sa,
                                     dc_sa, emit_clipped);
                    }
                }
            }
            __syncthreads();

            K2_TRACE(t, 4);
            // ---- 4. count the 0xFF bytes this tile owns and publish the count ---------------------------
            // owned bytes: [B0, B1) of the image's stream; window byte index = stream byte - 4*w0
            mine.t = t;
            mine.w0 = w0;
            mine.B0 = (begin + 7) >> 3;
            mine.B1 = fits ? (end + 7) >> 3 : mine.B0;
            const uint32_t wb0 = (uint32_t)(mine.B0 - 4 * w0), wb1 = (uint32_t)(mine.B1 - 4 * w0);   // window byte range
            const uint32_t wfirst = wb0 >> 2, wlast = (wb1 + 3) >> 2;                                // window word range
            uint32_t ffs = 0;
            for (uint32_t i = wfirst + tid; i < wlast; i += K2_THREADS) ffs += count_ff_bytes(masked_word(win, i, wb0, wb1));
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ffs += __shfl_xor_sync(0xffffffffu, ffs, o);
            if (lane == 0) s_warp[warp] = ffs;
            __syncthreads();
            mine.tile_ff = 0;
#pragma unroll
            for (int w = 0; w < K2_WARPS; ++w) mine.tile_ff += s_warp[w];
            if (tid == 0) {
                st_volatile_u64(a.ff_agg + t, LB_VALID | mine.tile_ff);
                // next ticket; its round trip overlaps the write-out below (static launches: one tile per CTA)
                ticket = a.dynamic_tiles ? atomicAdd(a.tile_counter, 1ull) : ~0ull;
            }
            K2_TRACE(t, 5);
        }

        // ---- 5. the previously packed tile: stuffed-zero look-back, then write its bytes ---------------
        // (after the last ticket, or with one tile per CTA, this is an extra round that only writes)
        const PendingTile w = pend;
        if (w.t != ~0ull) {
            const uint32_t *win = k2_smem_words + (cur ^ 1) * WIN_WORDS;
            const int img = a.count == 1 ? 0 : (int)((uint32_t)w.t / (uint32_t)a.tiles);
            const int tile = (int)((uint32_t)w.t - (uint32_t)img * (uint32_t)a.tiles);
            uint64_t *ff_incl = a.ff_incl + (uint64_t)img * groups;
            const uint64_t ff_excl = lookback_grouped(a.ff_agg + (uint64_t)img * a.tiles, ff_incl, tile, a.err, s_scratch);
            if (tid == 0) {
                if (tile % LB_GROUP == LB_GROUP - 1) st_volatile_u64(ff_incl + tile / LB_GROUP, LB_VALID | (ff_excl + w.tile_ff));
                if (tile == a.tiles - 1) {
                    const uint64_t size = w.B1 - origin + ff_excl + w.tile_ff;
                    a.image_bytes[img] = size;
                    if (a.count == 1) {
                        a.scan_offsets[0] = 0;
                        a.scan_offsets[1] = size;
                    }
                    if (size > a.out_capacity) atomicOr(a.err, a.count == 1 ? ERRBIT_OUTPUT : ERRBIT_WORKSPACE);
                }
            }
            K2_TRACE(w.t, 6);
            const uint32_t wb0 = (uint32_t)(w.B0 - 4 * w.w0), wb1 = (uint32_t)(w.B1 - 4 * w.w0);
            const uint32_t wfirst = wb0 >> 2, wlast = (wb1 + 3) >> 2;
            uint8_t *out = a.out + (uint64_t)img * a.out_slot;
            const uint64_t out_base = (w.B0 - origin) + ff_excl;      // output index of window byte wb0
            // two consecutive window words per thread and round: a typical tile (about 340 words) is one round
            uint32_t carry = 0;                                       // stuffed zeros of the earlier rounds (uniform)
            for (uint32_t i0 = wfirst; i0 < wlast; i0 += 2 * K2_THREADS) {
                if (i0 != wfirst) __syncthreads();                    // s_warp is rewritten in every round
                const uint32_t i = i0 + 2 * tid;
                const uint32_t v0 = i < wlast ? masked_word(win, i, wb0, wb1) : 0u;
                const uint32_t v1 = i + 1 < wlast ? masked_word(win, i + 1, wb0, wb1) : 0u;
                const uint32_t cnt = count_ff_bytes(v0) + count_ff_bytes(v1);
                uint32_t incl = cnt;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += n;
                }
                if (lane == 31) s_warp[warp] = incl;
                __syncthreads();
                uint32_t before = carry + incl - cnt;
                uint32_t round_total = 0;
#pragma unroll
                for (int ww = 0; ww < K2_WARPS; ++ww) {
                    const uint32_t ws = s_warp[ww];
                    if (ww < warp) before += ws;
                    round_total += ws;
                }
                carry += round_total;
                if (i < wlast) {
                    uint64_t pos = out_base + before + ((uint64_t)i * 4 > wb0 ? (uint64_t)i * 4 - wb0 : 0);
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        if (i + j < wlast) {
                            const uint32_t raw = win[i + j];
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint32_t wb = (i + j) * 4 + k;
                                if (wb >= wb0 && wb < wb1) {
                                    const uint8_t byte = (uint8_t)(raw >> (24 - 8 * k));
                                    if (pos < a.out_capacity) out[pos] = byte;
                                    ++pos;
                                    if (byte == 0xFF) {                 // huffman.c:29-31
                                        if (pos < a.out_capacity) out[pos] = 0x00;
                                        ++pos;
                                    }
                                }
                            }
                        }
                    }
                }
            }
            K2_TRACE(w.t, 7);
        }
        if (!have_tile) break;
        pend = mine;
        cur ^= 1;
        if (tid == 0) s_next_tile = ticket;
    }
}

// ---------------------------------------------------------------------------------
// batch mode: exclusive scan of the stuffed image sizes -> scan_offsets[count+1]  (one CTA)
// `extra`: bytes that frame every image in the output besides its scan (0, or 330 in files mode: the
// 328-byte JFIF header and the 2-byte EOI marker)
__global__ void __launch_bounds__(1024)
k_layout(const uint64_t *__restrict__ image_bytes, uint64_t *__restrict__ scan_offsets, const int count,
         const uint64_t scan_capacity, uint32_t *err, const uint64_t extra)
{
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < count; base += 1024) {
        const int i = base + tid;
        const uint64_t v = i < count ? image_bytes[i] + extra : 0;
        uint64_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint64_t n = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += n;
        }
        if (lane == 31) warp_sums[warp] = incl;
        __syncthreads();
        uint64_t wex = 0, tot = 0;
        for (int w = 0; w < 32; ++w) {
            const uint64_t ws = warp_sums[w];
            if (w < warp) wex += ws;
            tot += ws;
        }
        const uint64_t excl = carry_s + wex + incl - v;
        if (i < count) {
            scan_offsets[i] = excl;
            if (i == count - 1) {
                scan_offsets[count] = excl + v;
                if (excl + v > scan_capacity) atomicOr(err, ERRBIT_OUTPUT);
            }
        }
        __syncthreads();
        if (tid == 0) carry_s += tot;
        __syncthreads();
    }
}

// batch mode: move every image's stuffed bytes from its (16-byte aligned) slot to its final, arbitrarily
// aligned offset.  Destination-aligned 32-bit stores; the source is read as aligned words and realigned
// with a funnel shift; the few head / tail bytes are copied one by one.
__global__ void __launch_bounds__(256)
k_compact(const uint8_t *__restrict__ slots, const uint64_t slot_stride, const uint64_t *__restrict__ image_bytes,
          const uint64_t *__restrict__ scan_offsets, uint8_t *__restrict__ scan, const uint64_t scan_capacity,
          const uint64_t lead)
{
    const int img = blockIdx.y;
    const uint64_t n = image_bytes[img], dst0 = scan_offsets[img] + lead;   // lead: room for the file header
    if (scan_offsets[img + 1] > scan_capacity) return;           // flagged by k_layout
    const uint8_t *src = slots + (uint64_t)img * slot_stride;
    uint8_t *dst = scan + dst0;
    const uint64_t head = min(n, (uint64_t)((4u - (uint32_t)((uintptr_t)dst & 3u)) & 3u));   // bytes up to dst alignment
    const uint64_t nwords = (n - head) >> 2, tail0 = head + nwords * 4;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (uint64_t)gridDim.x * blockDim.x;
    if (tid < head) dst[tid] = src[tid];
    if (tid < n - tail0) dst[tail0 + tid] = src[tail0 + tid];
    const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src);
    uint32_t *d32 = reinterpret_cast<uint32_t *>(dst + head);
    const uint32_t sh = (uint32_t)head * 8u;                     // src word phase relative to dst words (head < 4)
    for (uint64_t j = tid; j < nwords; j += nthreads) d32[j] = __funnelshift_r(s32[j], s32[j + 1], sh);
}

// files mode: frame every image's scan with the JFIF header (jpeg_handler.c:220-233; the 328 bytes are
// the same for all images of a batch and travel as a kernel argument) and the EOI marker
// (jpeg_handler.c:262).  One CTA per image.  single: the count == 1 path, where K2 wrote the scan in
// place at offset 328 and left scan_offsets = {0, scan bytes}.
struct JfifHeaderBytes {
    uint8_t b[328];
};

__global__ void __launch_bounds__(128)
k_frame_files(const JfifHeaderBytes header, const uint64_t *__restrict__ image_bytes, uint64_t *__restrict__ file_offsets,
              uint8_t *__restrict__ files, const uint64_t capacity, const int single, uint32_t *err)
{
    const int img = blockIdx.x;
    const uint64_t n = image_bytes[img];
    const uint64_t begin = single ? 0 : file_offsets[img], end = single ? n + 330 : file_offsets[img + 1];
    if (single && threadIdx.x == 0) {
        file_offsets[1] = end;
        if (end > capacity) atomicOr(err, ERRBIT_OUTPUT);
    }
    if (end > capacity) return;
    for (int i = threadIdx.x; i < 328; i += blockDim.x) files[begin + i] = header.b[i];
    if (threadIdx.x == 0) {
        files[begin + 328 + n] = 0xFF;
        files[begin + 329 + n] = 0xD9;
    }
}

}  // namespace jb
