// scan_pack.cuh -- K2, the stream merge kernel: bit-offset scan + shift-merge of the strips' bit streams + 0xFF byte
// stuffing in ONE launch (plus the small batch-mode helpers).
//
// The reference's entropy stage is one serial chain: DC prediction across all blocks (rle.c:59-70) and one
// contiguous MSB-first bit stream (huffman.c:35-62) with a zero byte stuffed after every 0xFF (huffman.c:26-32)
// and a zero-padded last byte (huffman.c:65-81).
//
// K1b leaves, per 32-block strip, its complete entropy-coded bits (left-aligned in a slot of its own; every
// DC-difference symbol included, except the image's very first one) and a record {bits, first DC, last DC}, so
// everything that crosses strips collapses to two prefix sums and a shift:
//
//   tile = 16 consecutive strips, one CTA of 4 warps
//   1. the bit counts of the earlier strips are requested up front; the tile's bit offset is their plain sum
//      (wait-free; one checkpoint word per 1024 tiles);
//   2. a warp takes every fourth strip of the tile: its lanes load the strip's stream words (coalesced), funnel-shift
//      them to the global bit phase and OR-reduce them into the tile's pre-zeroed window in shared memory;
//   3. bytes are owned by the tile that holds their first bit: the last, partial byte is completed with the
//      leading bits of the following strip(s), the first partial byte is skipped -- no bits ever cross CTAs
//      through global memory;
//   4. 0xFF bytes of the tile's byte range are counted and published; a look-back over the earlier tiles' counts
//      gives the number of stuffed zeros before the tile, and the stuffed bytes go straight to the output.  A
//      persistent CTA resolves that look-back one tile late (after assembling its next tile into a second
//      window), when it no longer has to wait.
// CTAs draw tiles from an atomic ticket counter, in increasing order: a tile is only ever owned by a running CTA,
// so every look-back is eventually satisfied even when only part of the grid is resident.
#pragma once

#include "common.cuh"

namespace jb {

constexpr int K2_WARPS = 4;
constexpr int K2_THREADS = K2_WARPS * 32;
constexpr int K2_TILE_STRIPS = 16;                             // strips per tile
constexpr int K2_SEGS = K2_TILE_STRIPS + 1;                    // + the image's first DC symbol (the successor strips are handled apart)
constexpr int K2_ROUND_WORDS = 4 * K2_THREADS;                 // write-out: four window words per thread and round
constexpr int K2_STAGE_BYTES = 2 * 4 * K2_ROUND_WORDS + 32;    // stuffed bytes of one round (worst case: every byte 0xFF) + alignment phase
__host__ __device__ constexpr int k2_win_words(int slot_bytes) { return K2_TILE_STRIPS * slot_bytes / 4 + 8; }
__host__ __device__ constexpr int k2_smem(int slot_bytes) { return 2 * k2_win_words(slot_bytes) * 4 + K2_STAGE_BYTES; }

// stripes: boundary summary of one stripe (= jpegb200_stripe_summary).  bits_pred0 = the stripe's bit count if its
// first block were predicted from DC 0 (the strips' counts lack exactly that one symbol).
struct StripeSummaryDev {
    int16_t first_dc;
    int16_t last_dc;
    uint32_t valid;              // 1: the rank owns block rows; 0: empty slot
    uint64_t bits_pred0;
};

struct PackArgs {
    const uint8_t *tables;         // device table block (TBL_* offsets)
    const uint8_t *streams;        // [count*strips_avail][slot_bytes] strip bit streams (K1)
    uint32_t slot_bytes;
    const StripRec *strips;        // [count*strips_avail]
    const uint32_t *strip_bits;    // [count*strips_avail] compact copy of StripRec.bits (read four at a time)
    uint64_t *bit_incl;            // bit-offset checkpoints, one per 1024-tile group: [count*groups]
    uint64_t *ff_agg, *ff_incl;    // grouped look-back state of the stuffed-zero counts: [count*tiles], [count*groups]
    unsigned long long *tile_counter;   // next tile to hand out (cleared together with the look-back state)
    uint8_t *out;                  // stuffed bytes: caller's buffer (count==1) or per-image slots (batch)
    uint64_t out_capacity;         // bytes available per image at `out`
    uint64_t out_slot;             // byte distance between images at `out` (batch), 0 for count==1
    uint64_t *image_bytes;         // [count] stuffed size of each image
    uint64_t *image_bits;          // [count] total bits of each image (owned blocks)
    uint64_t *scan_offsets;        // count==1: scan_offsets[0..1] written here; batch: written by k_layout
    uint32_t *err;
    uint32_t strips_owned;         // strips per image that belong to the stream
    uint32_t strips_avail;         // >= strips_owned: strips K1 produced (stripe halo)
    int tiles;                     // tiles per image (over owned strips)
    int count;
    int16_t dc_pred0;              // DC predictor of the image's first block (0; stripes: previous stripe's last DC)
    uint32_t bit_phase;            // bit offset of the first bit inside byte 0 (0; stripes: global phase & 7)
    const StripeSummaryDev *dyn_all;   // stripes, device-resident exchange: all ranks' summaries (device memory); if set,
    int dyn_rank;                      // dc_pred0 / bit_phase are derived from them in the kernel's prologue
    unsigned long long *trace;     // optional [tiles*count][8] phase timestamps (ns), tuning aid; nullptr in production
};

__device__ __forceinline__ int magnitude_class_k2(int v)       // rle.c:9-22
{
    const int a = v < 0 ? -v : v;
    return 32 - __clz(a);
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

// the four window words i..i+3 restricted to the owned window bytes [wb0, wb1), and the number of 0xFF bytes among them
__device__ __forceinline__ uint32_t masked_quad(const uint32_t *win, uint32_t i, uint32_t wb0, uint32_t wb1, uint32_t v[4])
{
    const uint4 q = *reinterpret_cast<const uint4 *>(win + i);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
    if (4u * i < wb0 || 4u * i + 16u > wb1) {                  // only a tile's first and last group are ragged
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t lo = 4u * (i + j), hi = lo + 4u;
            if (lo < wb0) v[j] &= wb0 - lo >= 4u ? 0u : 0xFFFFFFFFu >> (8u * (wb0 - lo));
            if (hi > wb1) v[j] &= wb1 > lo ? 0xFFFFFFFFu << (8u * (hi - wb1)) : 0u;
        }
    }
    return count_ff_bytes(v[0]) + count_ff_bytes(v[1]) + count_ff_bytes(v[2]) + count_ff_bytes(v[3]);
}

// OR `n` left-aligned bits into the zeroed window at bit position `pos` (one thread; used for the few odd segments)
__device__ __forceinline__ void window_put(uint32_t *win, uint32_t pos, uint32_t vl)
{
    const uint32_t sh = pos & 31u;
    const uint32_t hi = vl >> sh, lo = __funnelshift_r(0u, vl, sh);
    if (hi) atomicOr(win + (pos >> 5), hi);
    if (lo) atomicOr(win + (pos >> 5) + 1, lo);
}

// 32 bits of a segment's stream starting at bit `off` of the segment (off may be negative: the segment starts inside
// the word), restricted to the segment's `len` bits.  src: the segment's words (bit 31 of word 0 is its first bit).
__device__ __forceinline__ uint32_t segment_bits(const uint32_t *src, int off, uint32_t len)
{
    if (off >= (int)len || off <= -32) return 0u;
    if (off >= 0) {
        const uint32_t wi = (uint32_t)off >> 5, sh = (uint32_t)off & 31u, nw = (len + 31u) >> 5;
        const uint32_t lo = src[wi], hi = wi + 1u < nw ? src[wi + 1u] : 0u;
        uint32_t v = __funnelshift_l(hi, lo, sh);
        const uint32_t valid = len - (uint32_t)off;
        if (valid < 32u) v &= 0xFFFFFFFFu << (32u - valid);
        return v;
    }
    uint32_t lo = src[0];
    if (len < 32u) lo &= 0xFFFFFFFFu << (32u - len);
    return lo >> (uint32_t)(-off);
}

// phase timestamps exist only in the tracing build of the library (make trace -> libjpegb200_trace.so)
#ifdef JPEGB200_TRACE
#define K2_TRACE(tile_id, slot) do { if (a.trace && tid == 0) a.trace[(tile_id) * 8 + (slot)] = globaltimer_ns(); } while (0)
#else
#define K2_TRACE(tile_id, slot) do { } while (0)
#endif

// An assembled tile whose bytes are not written yet.  A CTA assembles tile i+1 before it resolves the
// stuffed-zero look-back of tile i and writes it out: by then every predecessor has published its
// count, so the look-back never waits in steady state (two windows, used alternately).
struct PendingTile {
    uint64_t t;          // ticket (img * tiles + tile); ~0 = none
    uint64_t w0;         // stream word index of window word 0
    uint64_t B0, B1;     // owned stream bytes [B0, B1)
    uint32_t tile_ff;    // 0xFF bytes among them
    uint32_t nwords;     // window words in use (zeroed again after the write-out)
};

// The windows are sized at launch (dynamic shared memory): K2_TILE_STRIPS x the strip slot size.  An image that needs
// bigger slots raises ERRBIT_WORKSPACE in K1 and the caller re-runs both kernels with a larger bytes_per_block.
__global__ void __launch_bounds__(K2_THREADS, 7)
k_merge_stuff(const PackArgs a, const int win_words)
{
    extern __shared__ __align__(16) uint32_t k2_smem_words[];   // two windows, then the output staging area
    uint8_t *stage = reinterpret_cast<uint8_t *>(k2_smem_words + 2 * win_words);
    __shared__ uint32_t s_seg_end[32];              // inclusive prefix of the segment lengths (bits, relative to the tile's begin), padded with ~0
    __shared__ uint32_t s_seg_raw[2];               // bit counts of the two strips after the tile
    __shared__ uint32_t s_pseudo[2];                // the image's first DC symbol, left-aligned; its length
    __shared__ uint32_t s_warp[K2_WARPS], s_warp_b[K2_WARPS];
    __shared__ uint64_t s_scratch[9];
    __shared__ unsigned long long s_next_tile;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_next_tile = atomicAdd(a.tile_counter, 1ull);
    for (int i = tid; i < win_words / 2; i += K2_THREADS) reinterpret_cast<uint4 *>(k2_smem_words)[i] = make_uint4(0u, 0u, 0u, 0u);   // both windows
    int32_t dc_pred0 = (int32_t)a.dc_pred0;
    uint32_t bit_phase = a.bit_phase;
    if (a.dyn_all != nullptr) {
        // stripes: this rank's DC predictor and global bit offset from all ranks' summaries (the device version of
        // stripes.resolve_offsets: the true cost of a stripe's first DC symbol replaces the predictor-0 cost)
        uint64_t bit = 0;
        int pred = 0;
        for (int r = 0; r < a.dyn_rank; ++r) {
            const StripeSummaryDev sr = a.dyn_all[r];
            if (!sr.valid) continue;
            bit += sr.bits_pred0 - c_dc_len[magnitude_class_k2((int)sr.first_dc)] + c_dc_len[magnitude_class_k2((int)sr.first_dc - pred)];
            pred = sr.last_dc;
        }
        dc_pred0 = pred;
        bit_phase = (uint32_t)(bit & 7u);
    }
    const uint64_t origin = ((uint64_t)bit_phase + 7) >> 3;       // first stream byte this image/stripe owns
    const int groups = (a.tiles + LB_GROUP - 1) / LB_GROUP;
    const uint64_t total_tiles = (uint64_t)a.tiles * (uint64_t)a.count;
    const uint32_t cap_bits = a.slot_bytes * 8u;

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
            uint32_t *win = k2_smem_words + cur * win_words;
            const int img = a.count == 1 ? 0 : (int)((uint32_t)t / (uint32_t)a.tiles);   // tickets fit 32 bits
            const int tile = (int)((uint32_t)t - (uint32_t)img * (uint32_t)a.tiles);
            const uint64_t strip_base = (uint64_t)img * a.strips_avail;                  // index of the image's strip 0
            const uint32_t *sbits = a.strip_bits + strip_base;
            const uint32_t strip0 = (uint32_t)tile * K2_TILE_STRIPS;
            const uint32_t nstrips = min((uint32_t)K2_TILE_STRIPS, a.strips_owned - strip0);
            const bool last_tile = tile == a.tiles - 1;
            uint64_t *bit_incl = a.bit_incl + (uint64_t)img * groups;

            K2_TRACE(t, 0);
            // ---- 1. segment table and tile bit offset (wait-free) -------------------------------------------------
            // segments: 0 = the image's first DC symbol (tile 0 only), 1..nstrips = the tile's strips; up to two
            // following strips complete the tile's last byte
            const int g0 = (tile / LB_GROUP) * LB_GROUP;
            uint32_t part = 0;                                       // a 1024-tile group holds < 2^30 bits
            {
                const uint32_t lo = (uint32_t)g0 * K2_TILE_STRIPS, hi = strip0;
                const uint32_t head = min(hi, lo + ((4u - (uint32_t)((strip_base + lo) & 3u)) & 3u));   // up to 16-byte alignment
                if (lo + tid < head) part += sbits[lo + tid];
                const uint32_t nvec = (hi - head) >> 2, tail = head + 4u * nvec;
                const uint4 *v4 = reinterpret_cast<const uint4 *>(sbits + head);
                for (uint32_t i = tid; i < nvec; i += K2_THREADS) {
                    const uint4 v = v4[i];
                    part += (v.x + v.y) + (v.z + v.w);
                }
                if (tail + tid < hi) part += sbits[tail + tid];
            }
            // the first 64 words of this warp's strips are requested now, together with the bit counts: one global
            // round trip for both (words beyond a strip's end are ignored below; a slot holds at least 64 words)
            const uint8_t *tile_streams = a.streams + (strip_base + strip0) * (uint64_t)a.slot_bytes;
            uint32_t early[K2_TILE_STRIPS / K2_WARPS][2];
#pragma unroll
            for (int q = 0; q < K2_TILE_STRIPS / K2_WARPS; ++q) {
                const uint32_t kk = warp + q * K2_WARPS;
                early[q][0] = early[q][1] = 0;
                if (kk < nstrips) {
                    const uint32_t *src = reinterpret_cast<const uint32_t *>(tile_streams + kk * (uint64_t)a.slot_bytes);
                    early[q][0] = src[lane];
                    early[q][1] = src[lane + 32];
                }
            }
            uint32_t fix0 = 0;                                       // bits of the image's first DC symbol (rle.c:68-76)
            if (warp == 0) {
                uint32_t len = 0;
                if (lane == 0) {
                    const int diff = (int)a.strips[strip_base].first_dc - dc_pred0;
                    const int sz = magnitude_class_k2(diff);
                    const uint32_t hc = c_dc_code[sz];
                    const uint32_t amp = (uint32_t)(diff > 0 ? diff : diff - 1) & ((1u << sz) - 1u);   // rle.c:24-35, huffman.c:39
                    fix0 = (hc & 0xFFu) + (uint32_t)sz;
                    s_pseudo[0] = (((hc >> 8) << sz) | amp) << (32u - fix0);
                    s_pseudo[1] = fix0;
                    len = tile == 0 ? fix0 : 0u;
                } else if ((uint32_t)lane <= nstrips) {
                    len = min(sbits[strip0 + lane - 1], cap_bits);
                } else if ((uint32_t)lane <= nstrips + 2u) {
                    const uint32_t st = strip0 + lane - 1;           // a strip after the tile
                    s_seg_raw[lane - nstrips - 1] = st < a.strips_avail ? min(sbits[st], cap_bits) : 0u;
                }
                uint32_t incl = len;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += n;
                }
                s_seg_end[lane] = (uint32_t)lane <= nstrips ? incl : 0xFFFFFFFFu;
            }
            if (tid == 0) s_scratch[K2_WARPS] = g0 > 0 ? lb_wait(bit_incl + tile / LB_GROUP - 1, a.err) : (tile > 0 ? fix0 : 0u);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0) s_scratch[warp] = part;
            __syncthreads();
            K2_TRACE(t, 1);
            uint64_t bit_excl = s_scratch[K2_WARPS];                 // the group's checkpoint (64-bit)
#pragma unroll
            for (int w = 0; w < K2_WARPS; ++w) bit_excl += s_scratch[w];
            const uint32_t tile_bits = s_seg_end[nstrips];
            const uint64_t begin = bit_excl + bit_phase, end = begin + tile_bits;
            if (tid == 0) {
                if (tile % LB_GROUP == LB_GROUP - 1) st_volatile_u64(bit_incl + tile / LB_GROUP, LB_VALID | (bit_excl + tile_bits));
                if (last_tile) a.image_bits[img] = bit_excl + tile_bits;
            }
            // bits that complete the last owned byte: taken from the strips after the tile (a strip can be as
            // short as 6 bits, so up to two); the image's last tile is zero-padded instead (huffman.c:65-81)
            const uint32_t need = (8u - (uint32_t)(end & 7u)) & 7u;
            const uint32_t succ1 = min(need, s_seg_raw[0]), succ2 = min(need - succ1, s_seg_raw[1]);
            const uint64_t w0 = begin >> 5;
            const uint64_t limit = (end + 7) & ~7ull;                // first bit NOT owned by this tile
            uint32_t nwords = (uint32_t)(((limit + 31) >> 5) - w0);  // covers the completed last byte
            const bool fits = nwords + 6 <= (uint32_t)win_words;
            if (!fits) {                                              // cannot happen with consistent slot sizes; never overrun
                if (tid == 0) atomicOr(a.err, ERRBIT_WORKSPACE);
                nwords = 0;
            }
            K2_TRACE(t, 2);

            // ---- 2. assemble the tile's part of the stream: every strip is shifted to the global bit phase -------------
            // A warp takes every fourth strip; a lane loads stream word i (coalesced), gets word i-1 from its neighbour
            // and OR-reduces ((word i-1 : word i) >> phase) into the zeroed window.  The strips' streams are zero beyond
            // their last bit (K1b), so nothing has to be masked.
            if (fits) {
                const uint32_t rel_begin = (uint32_t)(begin - (w0 << 5));         // bit position of the tile's first bit in the window
#pragma unroll
                for (int q = 0; q < K2_TILE_STRIPS / K2_WARPS; ++q) {
                    const uint32_t kk = warp + q * K2_WARPS;
                    if (kk >= nstrips) break;
                    const uint32_t seg_begin = s_seg_end[kk], len = s_seg_end[kk + 1] - seg_begin;
                    const uint32_t dstbit = rel_begin + seg_begin, sh = dstbit & 31u, nw = (len + 31u) >> 5;
                    const uint32_t *src = reinterpret_cast<const uint32_t *>(tile_streams + kk * (uint64_t)a.slot_bytes);
                    uint32_t *dst = win + (dstbit >> 5);
                    uint32_t carry = 0;                                           // word i-1 for lane 0
                    for (uint32_t i0 = 0; i0 < nw + 1; i0 += 32) {
                        const uint32_t i = i0 + lane;
                        uint32_t cur_w = i0 == 0 ? early[q][0] : i0 == 32 ? early[q][1] : (i < nw ? src[i] : 0u);
                        if (i >= nw) cur_w = 0u;
                        uint32_t prev_w = __shfl_up_sync(0xffffffffu, cur_w, 1);
                        if (lane == 0) prev_w = carry;
                        carry = __shfl_sync(0xffffffffu, cur_w, 31);
                        const uint32_t v = __funnelshift_r(cur_w, prev_w, sh);    // (word i-1 : word i) >> phase
                        if (v != 0u && i < nw + 1) atomicOr(dst + i, v);
                    }
                }
                if (tid == 0) {
                    if (tile == 0) window_put(win, rel_begin, s_pseudo[0]);       // the image's first DC symbol
                    // the (at most two) strips after the tile, clipped to the bits that complete the last byte
                    if (succ1) {
                        const uint32_t w1 = *reinterpret_cast<const uint32_t *>(tile_streams + nstrips * (uint64_t)a.slot_bytes);
                        window_put(win, rel_begin + tile_bits, w1 & (0xFFFFFFFFu << (32u - succ1)));
                        if (succ2) {
                            const uint32_t w2 = *reinterpret_cast<const uint32_t *>(tile_streams + (nstrips + 1) * (uint64_t)a.slot_bytes);
                            window_put(win, rel_begin + tile_bits + succ1, w2 & (0xFFFFFFFFu << (32u - succ2)));
                        }
                    }
                }
            }
            __syncthreads();

            K2_TRACE(t, 4);
            // ---- 3. count the 0xFF bytes this tile owns and publish the count ---------------------------
            // owned bytes: [B0, B1) of the image's stream; window byte index = stream byte - 4*w0
            mine.t = t;
            mine.w0 = w0;
            mine.B0 = (begin + 7) >> 3;
            mine.B1 = fits ? (end + 7) >> 3 : mine.B0;
            const uint32_t wb0 = (uint32_t)(mine.B0 - 4 * w0), wb1 = (uint32_t)(mine.B1 - 4 * w0);   // window byte range
            const uint32_t wlast = (wb1 + 3) >> 2;                                                   // window word range
            mine.nwords = nwords;
            uint32_t ffs = 0;
            for (uint32_t i = ((wb0 >> 2) & ~3u) + 4u * tid; i < wlast; i += 4u * K2_THREADS) {
                uint32_t v[4];
                ffs += masked_quad(win, i, wb0, wb1, v);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ffs += __shfl_xor_sync(0xffffffffu, ffs, o);
            if (lane == 0) s_warp[warp] = ffs;
            __syncthreads();
            mine.tile_ff = 0;
#pragma unroll
            for (int w = 0; w < K2_WARPS; ++w) mine.tile_ff += s_warp[w];
            if (tid == 0) {
                st_volatile_u64(a.ff_agg + t, LB_VALID | mine.tile_ff);
                ticket = atomicAdd(a.tile_counter, 1ull);        // next ticket; its round trip overlaps the write-out below
            }
            K2_TRACE(t, 5);
        }

        // ---- 4. the previously assembled tile: stuffed-zero look-back, then write its bytes -------------------
        // (after the last ticket this is an extra round that only writes)
        const PendingTile w = pend;
        if (w.t != ~0ull) {
            const uint32_t *win = k2_smem_words + (cur ^ 1) * win_words;
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
            const uint32_t wlast = (wb1 + 3) >> 2;
            uint8_t *out = a.out + (uint64_t)img * a.out_slot;
            uint64_t round_pos = (w.B0 - origin) + ff_excl;           // output index of the round's first byte
            // Rounds of four consecutive window words (16 stream bytes) per thread and quad; a typical tile (about 680
            // words, a few 0xFF bytes) is ONE round of two quads per thread, a tile whose stuffed bytes might not fit the
            // staging area takes rounds of one quad.  The stuffed bytes (huffman.c:26-32) are laid out in the staging area
            // at the output's 16-byte phase and leave as aligned 128-bit stores; only the ragged first / last 16-byte
            // unit of a round is written byte by byte.
            const uint32_t i_start = (wb0 >> 2) & ~3u;
            const uint32_t Q = (wlast - i_start <= 2u * K2_ROUND_WORDS &&
                                (uint32_t)(w.B1 - w.B0) + w.tile_ff + 64u <= (uint32_t)(K2_STAGE_BYTES - 32)) ? 2u : 1u;
            for (uint32_t i0 = i_start; i0 < wlast; i0 += Q * K2_ROUND_WORDS) {
                uint32_t v[2][4], own[2], incl[2], lo[2], hi[2];
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    own[q] = 0;
                    lo[q] = hi[q] = 0;
                    if ((uint32_t)q < Q) {
                        const uint32_t i = i0 + 4u * ((uint32_t)tid + K2_THREADS * q);
                        const uint32_t cnt = masked_quad(win, i, wb0, wb1, v[q]);     // beyond wlast: masked off
                        // owned bytes of this quad: window bytes [lo, hi)
                        lo[q] = min(max(4u * i, wb0), wb1);
                        hi[q] = max(min(4u * i + 16u, wb1), lo[q]);
                        own[q] = (cnt << 16) | (hi[q] - lo[q]);                       // both prefix sums at once (<= 2^13 each per round)
                    }
                    incl[q] = own[q];
                }
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t n0 = __shfl_up_sync(0xffffffffu, incl[0], o), n1 = __shfl_up_sync(0xffffffffu, incl[1], o);
                    if (lane >= o) {
                        incl[0] += n0;
                        incl[1] += n1;
                    }
                }
                if (lane == 31) {
                    s_warp[warp] = incl[0];
                    s_warp_b[warp] = incl[1];
                }
                __syncthreads();
                uint32_t before[2] = {incl[0] - own[0], incl[1] - own[1]}, total0 = 0, total1 = 0;
#pragma unroll
                for (int ww = 0; ww < K2_WARPS; ++ww) {
                    const uint32_t wa = s_warp[ww], wb_ = s_warp_b[ww];
                    if (ww < warp) {
                        before[0] += wa;
                        before[1] += wb_;
                    }
                    total0 += wa;
                    total1 += wb_;
                }
                before[1] += total0;                                              // the second quads follow all first quads
                const uint32_t round_total = total0 + total1;
                const uint32_t round_bytes = (round_total & 0xFFFFu) + (round_total >> 16);   // stuffed bytes of this round
                const uint32_t phase = (uint32_t)((uintptr_t)(out + round_pos) & 15u);
#pragma unroll
                for (int q = 0; q < 2; ++q) {
                    if ((uint32_t)q < Q) {
                        const uint32_t i = i0 + 4u * ((uint32_t)tid + K2_THREADS * q);
                        uint32_t rel = phase + (before[q] & 0xFFFFu) + (before[q] >> 16);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
#pragma unroll
                            for (int kb = 0; kb < 4; ++kb) {
                                const uint32_t b = 4u * (i + j) + kb;
                                if (b >= lo[q] && b < hi[q]) {
                                    const uint32_t byte = (v[q][j] >> (24 - 8 * kb)) & 0xFFu;
                                    stage[rel++] = (uint8_t)byte;
                                    if (byte == 0xFFu) stage[rel++] = 0;          // huffman.c:29-31
                                }
                            }
                        }
                    }
                }
                __syncthreads();
                {
                    uint8_t *dst = out + round_pos - phase;                       // 16-byte aligned
                    const uint64_t room = a.out_capacity > round_pos ? a.out_capacity - round_pos : 0;   // bytes of this round that fit
                    const uint32_t nbytes = (uint32_t)min((uint64_t)round_bytes, room);
                    const uint32_t first = phase, last = phase + nbytes;          // staging range to copy
                    for (uint32_t u = tid; 16u * u < last; u += K2_THREADS) {
                        const uint32_t ub = 16u * u;
                        if (ub >= first && ub + 16u <= last) {
                            reinterpret_cast<uint4 *>(dst)[u] = reinterpret_cast<const uint4 *>(stage)[u];
                        } else {
                            for (uint32_t x = max(ub, first); x < min(ub + 16u, last); ++x) dst[x] = stage[x];
                        }
                    }
                }
                round_pos += round_bytes;
                __syncthreads();                                                  // staging area and the warp sums are reused
            }
            {   // the window is handed back zeroed
                uint4 *wq = reinterpret_cast<uint4 *>(k2_smem_words + (cur ^ 1) * win_words);
                for (uint32_t i = tid; 4u * i < w.nwords + 4u; i += K2_THREADS) wq[i] = make_uint4(0u, 0u, 0u, 0u);
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
// aligned offset.  Destination-aligned 128-bit stores; the source is read as aligned words and realigned
// with funnel shifts; the few head / tail bytes are copied one by one.
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
    const uint64_t head = min(n, (uint64_t)((16u - (uint32_t)((uintptr_t)dst & 15u)) & 15u));   // bytes up to dst alignment
    const uint64_t nvec = (n - head) >> 4, tail0 = head + nvec * 16;
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nthreads = (uint64_t)gridDim.x * blockDim.x;
    if (tid < head) dst[tid] = src[tid];
    if (tid < n - tail0) dst[tail0 + tid] = src[tail0 + tid];
    const uint32_t *s32 = reinterpret_cast<const uint32_t *>(src) + (head >> 2);   // source word that holds byte `head`
    uint4 *d128 = reinterpret_cast<uint4 *>(dst + head);
    const uint32_t sh = (uint32_t)(head & 3u) * 8u;              // byte phase of the source inside its word
    for (uint64_t j = tid; j < nvec; j += nthreads) {
        const uint32_t *p = s32 + 4 * j;
        const uint32_t w0 = p[0], w1 = p[1], w2 = p[2], w3 = p[3], w4 = sh ? p[4] : 0u;
        d128[j] = make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh),
                             __funnelshift_r(w3, w4, sh));
    }
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
