// entropy.cuh -- K2 (bit-offset scan), K3 (scatter bit-pack), K4 (0xFF stuffing).
//
// The reference's entropy stage is one serial chain: DC prediction across all blocks
// (rle.c:59-70) and one contiguous MSB-first bit stream (huffman.c:35-62) with a zero
// byte stuffed after every 0xFF (huffman.c:26-32) and a zero-padded last byte
// (huffman.c:65-81).  Here:
//   K2  per-block bit cost = AC cost (from K1) + DC-difference cost, then an exclusive
//       prefix sum over the image's blocks: one pass, decoupled look-back between
//       1024-block tiles, segmented per image (grid.y = image).
//   K3  one lane per block re-derives the block's symbols from its 64 int8
//       coefficients and ORs code+amplitude bits into a shared-memory window at the
//       block's bit offset; the window is then written out coalesced (boundary words
//       with atomicOr, they are shared with the neighbouring tiles).
//   K4  counts 0xFF bytes per 4 KiB chunk, look-back prefix over the chunks of the
//       image, and writes the stuffed stream.
#pragma once

#include "common.cuh"

namespace jb {

constexpr int K2_THREADS = 256;
constexpr int K2_BLOCKS_PER_THREAD = 4;
constexpr int K2_TILE = K2_THREADS * K2_BLOCKS_PER_THREAD;     // 1024 blocks per scan tile
constexpr int K3_THREADS = 256;                                // blocks per pack tile
constexpr int K3_MAX_BLOCK_BITS = 1472;                        // >= 14 + 63*23 = 1463
constexpr int K3_SMEM_WORDS = (K3_THREADS * K3_MAX_BLOCK_BITS) / 32 + 4;
constexpr int K4_THREADS = 256;
constexpr int K4_CHUNK = K4_THREADS * 16;                      // 4096 bytes

// per-launch entropy parameters
struct EntropyArgs {
    const int8_t *coef;            // [count*nb][64] zig-zag int8
    const uint32_t *blockinfo;     // [count*nb]  (ac_bits << 16) | (uint16)dc
    uint32_t *blockoff;            // [count*nb]  bit offset within the block's scan tile
    uint64_t *tilebase;            // [count*tiles] bit offset of the tile within its image
    uint64_t *scan_state;          // [count*tiles] look-back state (K2)
    uint64_t *image_bits;          // [count] total bits per image
    uint64_t *image_base;          // [count] byte offset of the image's packed bits in `packed` (16-B aligned)
    uint32_t *packed;              // unstuffed stream words
    uint64_t packed_capacity;      // bytes
    uint64_t *stuff_state;         // [count*chunks_cap] look-back state (K4)
    uint64_t *image_ff;            // [count] number of 0xFF bytes per image
    uint8_t *scan;                 // output: stuffed bytes
    uint64_t scan_capacity;
    uint64_t *scan_offsets;        // [count+1]
    uint32_t *err;
    uint64_t nb;                   // blocks per image
    int tiles;                     // scan tiles per image
    int chunks_cap;                // K4 chunks per image the grid covers
    int count;
    uint32_t epoch;
    int16_t dc_pred0;              // DC predictor of each image's first block (0; stripes: previous stripe's last DC)
    uint32_t bit_phase;            // bit offset of the first bit inside packed byte 0 (0; stripes: global phase)
};

__device__ __forceinline__ int magnitude_class(int v)          // rle.c:9-22
{
    const int a = v < 0 ? -v : v;
    return 32 - __clz(a);
}

// ---------------------------------------------------------------------------------
// K2
__global__ void __launch_bounds__(K2_THREADS)
k_bit_scan(const EntropyArgs a)
{
    __shared__ uint32_t warp_sums[K2_THREADS / 32];
    __shared__ uint64_t tile_excl;
    const int tile = blockIdx.x, img = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t img_b0 = (uint64_t)img * a.nb;
    const uint64_t local0 = (uint64_t)tile * K2_TILE + (uint64_t)tid * K2_BLOCKS_PER_THREAD;

    uint32_t bits[K2_BLOCKS_PER_THREAD];
    uint32_t sum = 0;
    int prev_dc = 0;
    if (local0 < a.nb)
        prev_dc = local0 == 0 ? (int)a.dc_pred0 : (int)(int16_t)(a.blockinfo[img_b0 + local0 - 1] & 0xFFFFu);
#pragma unroll
    for (int j = 0; j < K2_BLOCKS_PER_THREAD; ++j) {
        bits[j] = 0;
        if (local0 + j < a.nb) {
            const uint32_t info = a.blockinfo[img_b0 + local0 + j];
            const int dc = (int)(int16_t)(info & 0xFFFFu);
            bits[j] = (info >> 16) + c_dc_len[magnitude_class(dc - prev_dc)];   // rle.c:68-76
            prev_dc = dc;
        }
        sum += bits[j];
    }
    // block-wide exclusive scan of per-thread sums
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t warp_excl = 0, total = 0;
#pragma unroll
    for (int w = 0; w < K2_THREADS / 32; ++w) {
        const uint32_t ws = warp_sums[w];
        if (w < warp) warp_excl += ws;
        total += ws;
    }
    uint64_t *state = a.scan_state + (uint64_t)img * a.tiles;
    if (warp == 0) {
        if (lane == 0)
            st_volatile_u64(state + tile, lb_pack(a.epoch, tile == 0 ? LB_PREFIX : LB_AGGREGATE, total));
        const uint64_t excl = lookback_exclusive(state, tile, a.epoch, a.err);
        if (lane == 0) {
            if (tile != 0) st_volatile_u64(state + tile, lb_pack(a.epoch, LB_PREFIX, excl + total));
            tile_excl = excl;
            a.tilebase[(uint64_t)img * a.tiles + tile] = excl;
            if (tile == a.tiles - 1) a.image_bits[img] = excl + total;
        }
    }
    uint32_t off = warp_excl + incl - sum;
#pragma unroll
    for (int j = 0; j < K2_BLOCKS_PER_THREAD; ++j) {
        if (local0 + j < a.nb) a.blockoff[img_b0 + local0 + j] = off;
        off += bits[j];
    }
    // single image: the packed stream starts at byte 0, so this kernel can also clear the
    // words K3 tiles share (each pack tile's first word) and the stream's last word.
    if (a.count == 1) {
        __syncthreads();
        const uint64_t base = tile_excl + a.bit_phase;
        uint32_t o2 = warp_excl + incl - sum;
#pragma unroll
        for (int j = 0; j < K2_BLOCKS_PER_THREAD; ++j) {
            const uint64_t lb = local0 + j;
            if (lb < a.nb) {
                const uint64_t w = (base + o2) >> 5;
                if ((lb % K3_THREADS) == 0 && (w + 1) * 4 <= a.packed_capacity) a.packed[w] = 0;
                if (lb == a.nb - 1) {
                    const uint64_t wl = (base + o2 + bits[j] - 1) >> 5;
                    if ((wl + 1) * 4 <= a.packed_capacity) a.packed[wl] = 0;
                }
            }
            o2 += bits[j];
        }
        if (tile == 0 && tid == 0 && a.packed_capacity >= 4) a.packed[0] = 0;
    }
}

// ---------------------------------------------------------------------------------
// image layout: exclusive scan of f(in[i]) over images, one CTA.
//   mode 0: in = image_bits  -> out = image_base  (packed bytes, 16-byte aligned slots)
//   mode 1: in = image_bits/image_ff -> out = scan_offsets[count+1] (stuffed sizes)
__global__ void __launch_bounds__(1024)
k_image_layout(const EntropyArgs a, const int mode)
{
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (int base = 0; base < a.count; base += 1024) {
        const int i = base + tid;
        uint64_t v = 0;
        if (i < a.count) {
            const uint64_t nbytes = (a.image_bits[i] + a.bit_phase + 7) >> 3;
            v = mode == 0 ? ((nbytes + 15) & ~15ull) + 16 : nbytes + a.image_ff[i];
        }
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
        const uint64_t carry = carry_s;
        const uint64_t excl = carry + wex + incl - v;
        if (i < a.count) {
            if (mode == 0) {
                a.image_base[i] = excl;
                if (excl + v > a.packed_capacity) atomicOr(a.err, ERRBIT_WORKSPACE);
            } else {
                a.scan_offsets[i] = excl;
                if (i == a.count - 1) {
                    a.scan_offsets[a.count] = excl + v;
                    if (excl + v > a.scan_capacity) atomicOr(a.err, ERRBIT_OUTPUT);
                }
            }
        }
        __syncthreads();
        if (tid == 0) carry_s = carry + tot;
        __syncthreads();
    }
}

// batch mode: clear the words that pack tiles share (tile-start words, last word)
__global__ void k_zero_shared_words(const EntropyArgs a)
{
    const uint64_t ptiles = (a.nb + K3_THREADS - 1) / K3_THREADS;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= ptiles * (uint64_t)a.count) return;
    const uint64_t img = i / ptiles, pt = i - img * ptiles;
    const uint64_t lb = pt * K3_THREADS;
    const uint64_t base_bytes = a.image_base[img];
    if (base_bytes + (((a.image_bits[img] + a.bit_phase + 7) >> 3) + 3) > a.packed_capacity) return;   // flagged by layout
    uint32_t *out = a.packed + (base_bytes >> 2);
    const uint64_t off = a.tilebase[img * a.tiles + lb / K2_TILE] + a.blockoff[img * a.nb + lb] + a.bit_phase;
    out[off >> 5] = 0;
    if (pt == 0) out[0] = 0;
    if (pt == ptiles - 1) out[(a.image_bits[img] + a.bit_phase - 1) >> 5] = 0;
}

// ---------------------------------------------------------------------------------
// K3
struct BitCursor {
    uint32_t *win;      // shared-memory window (zeroed), word 0 = stream word w0
    uint32_t pos;       // bit position relative to the window start
};

// append the low n bits of v (n <= 32, MSB first)
__device__ __forceinline__ void put_bits(BitCursor &c, uint32_t v, uint32_t n)
{
    const uint32_t wi = c.pos >> 5, bo = c.pos & 31u;
    const uint64_t x = ((uint64_t)v << (64u - n)) >> bo;      // value left-aligned at bit `bo` of a 64-bit lane
    const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
    if (hi) atomicOr(c.win + wi, hi);
    if (lo) atomicOr(c.win + wi + 1, lo);
    c.pos += n;
}

__global__ void __launch_bounds__(K3_THREADS)
k_pack(const EntropyArgs a)
{
    extern __shared__ __align__(16) uint32_t win[];
    __shared__ uint64_t s_begin, s_end;
    const int tid = threadIdx.x, img = blockIdx.y;
    const uint64_t lb = (uint64_t)blockIdx.x * K3_THREADS + tid;
    const bool valid = lb < a.nb;
    const uint64_t b = (uint64_t)img * a.nb + lb;
    const uint64_t total_bits = a.image_bits[img];
    const uint64_t base_bytes = a.count == 1 ? 0 : a.image_base[img];
    if (base_bytes + (((total_bits + a.bit_phase + 7) >> 3) + 3) > a.packed_capacity) {
        if (tid == 0) atomicOr(a.err, ERRBIT_WORKSPACE);
        return;
    }
    uint64_t off = 0;
    if (valid) off = a.tilebase[(uint64_t)img * a.tiles + lb / K2_TILE] + a.blockoff[b] + a.bit_phase;
    if (tid == 0) s_begin = off;
    const uint64_t last_lb = min((uint64_t)(blockIdx.x + 1) * K3_THREADS, a.nb) - 1;
    if (lb == last_lb) {
        const uint64_t nxt = last_lb + 1;
        s_end = nxt < a.nb ? a.tilebase[(uint64_t)img * a.tiles + nxt / K2_TILE] + a.blockoff[(uint64_t)img * a.nb + nxt] + a.bit_phase
                           : total_bits + a.bit_phase;
    }
    __syncthreads();
    const uint64_t w0 = s_begin >> 5;
    const uint32_t nwords = (uint32_t)(((s_end + 31) >> 5) - w0);
    for (uint32_t i = tid; i < nwords + 1; i += K3_THREADS) win[i] = 0;
    __syncthreads();

    if (valid) {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.coef + b * 64);
        uint32_t zw[16];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint4 q = src[i];
            zw[4 * i] = q.x; zw[4 * i + 1] = q.y; zw[4 * i + 2] = q.z; zw[4 * i + 3] = q.w;
        }
        BitCursor cur{win, (uint32_t)(off - (w0 << 5))};
        // DC difference (rle.c:68-76, huffman.c:145-153)
        const int prev_dc = lb == 0 ? (int)a.dc_pred0 : (int)(int16_t)(a.blockinfo[b - 1] & 0xFFFFu);
        const int dc = (int)(int8_t)(zw[0] & 0xFFu);
        {
            const int diff = dc - prev_dc;
            const int sz = magnitude_class(diff);
            const uint32_t hc = c_dc_code[sz];
            const uint32_t amp = (uint32_t)(diff > 0 ? diff : diff - 1) & ((1u << sz) - 1u);    // rle.c:24-35, huffman.c:39
            put_bits(cur, ((hc >> 8) << sz) | amp, (hc & 0xFFu) + sz);
        }
        // AC run lengths (rle.c:83-123, huffman.c:158-188)
        int run = 0;
#pragma unroll 1
        for (int k = 1; k < 64; ++k) {
            const int v = (int)(int8_t)((zw[k >> 2] >> (8 * (k & 3))) & 0xFFu);
            if (v == 0) { ++run; continue; }
            while (run >= 16) {                                 // ZRL (rle.c:99-103)
                const uint32_t z = c_ac_code[0xF0];
                put_bits(cur, z >> 8, z & 0xFFu);
                run -= 16;
            }
            const int sz = magnitude_class(v);
            const uint32_t hc = c_ac_code[(run << 4) | sz];
            const uint32_t amp = (uint32_t)(v > 0 ? v : v - 1) & ((1u << sz) - 1u);
            put_bits(cur, ((hc >> 8) << sz) | amp, (hc & 0xFFu) + sz);
            run = 0;
        }
        if (run > 0) {                                          // EOB (rle.c:121-123)
            const uint32_t e = c_ac_code[0x00];
            put_bits(cur, e >> 8, e & 0xFFu);
        }
    }
    __syncthreads();
    // stream bytes are MSB-first: byte-swap the big-endian window words for memory order
    uint32_t *out = a.packed + (base_bytes >> 2) + w0;
    for (uint32_t i = tid; i < nwords; i += K3_THREADS) {
        const uint32_t v = __byte_perm(win[i], 0u, 0x0123u);
        // a word is shared with the neighbouring tile iff the tile boundary falls inside it;
        // exactly those words were cleared beforehand (K2 / k_zero_shared_words)
        const bool shared = (i == 0 && (s_begin & 31u)) || (i == nwords - 1 && (s_end & 31u));
        if (shared) { if (v) atomicOr(out + i, v); }
        else out[i] = v;
    }
}

// ---------------------------------------------------------------------------------
// K4
__device__ __forceinline__ uint32_t count_ff_bytes(uint32_t w)
{
    uint32_t x = w & (w >> 4);
    x &= x >> 2;
    x &= x >> 1;
    return __popc(x & 0x01010101u);
}

// mode 0: count + look-back + write in one pass (single image / stripe; scan_offsets
//         filled here).  mode 1: count only (publishes inclusive prefixes, image_ff).
//         mode 2: write using the prefixes of a completed mode-1 pass and scan_offsets.
// Stripe controls: bytes [byte_begin, nbytes) of the packed stream are emitted and
// `or_last` is ORed into the last byte first.
struct StuffArgs {
    int mode;
    uint32_t byte_begin;
    uint32_t or_last;
};

__global__ void __launch_bounds__(K4_THREADS)
k_stuff(const EntropyArgs a, const StuffArgs sa)
{
    __shared__ uint32_t warp_sums[K4_THREADS / 32];
    __shared__ uint64_t chunk_excl;
    const int chunk = blockIdx.x, img = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t nbytes = (a.image_bits[img] + a.bit_phase + 7) >> 3;
    const uint64_t c0 = (uint64_t)chunk * K4_CHUNK;
    if (c0 >= nbytes) return;
    const uint64_t base_bytes = a.count == 1 ? 0 : a.image_base[img];
    if (base_bytes + nbytes + 3 > a.packed_capacity) return;        // flagged earlier
    const uint64_t t0 = c0 + (uint64_t)tid * 16;

    uint4 d = make_uint4(0, 0, 0, 0);
    if (t0 < nbytes) d = *reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(a.packed) + base_bytes + t0);
    uint32_t w[4] = {d.x, d.y, d.z, d.w};
    // bytes outside [byte_begin, nbytes) are not part of this stream: blank them
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint64_t p = t0 + 4 * j;
        if (p + 4 > nbytes || p < sa.byte_begin) {
            uint32_t keep = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (p + k < nbytes && p + k >= sa.byte_begin) keep |= 0xFFu << (8 * k);
            w[j] &= keep;
        }
        if (sa.or_last && nbytes - 1 >= p && nbytes - 1 < p + 4) w[j] |= (sa.or_last & 0xFFu) << (8 * (uint32_t)(nbytes - 1 - p));
    }
    const uint32_t cnt = count_ff_bytes(w[0]) + count_ff_bytes(w[1]) + count_ff_bytes(w[2]) + count_ff_bytes(w[3]);

    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += n;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t warp_excl = 0, total = 0;
#pragma unroll
    for (int k = 0; k < K4_THREADS / 32; ++k) {
        const uint32_t ws = warp_sums[k];
        if (k < warp) warp_excl += ws;
        total += ws;
    }
    uint64_t *state = a.stuff_state + (uint64_t)img * a.chunks_cap;
    const bool last_chunk = c0 + K4_CHUNK >= nbytes;
    if (warp == 0) {
        uint64_t excl;
        if (sa.mode == 2) {
            excl = chunk == 0 ? 0 : (ld_volatile_u64(state + chunk - 1) & LB_VALUE_MASK);
        } else {
            if (lane == 0)
                st_volatile_u64(state + chunk, lb_pack(a.epoch, chunk == 0 ? LB_PREFIX : LB_AGGREGATE, total));
            excl = lookback_exclusive(state, chunk, a.epoch, a.err);
            if (lane == 0 && chunk != 0) st_volatile_u64(state + chunk, lb_pack(a.epoch, LB_PREFIX, excl + total));
        }
        if (lane == 0) {
            chunk_excl = excl;
            if (last_chunk && sa.mode != 2) {
                a.image_ff[img] = excl + total;
                if (sa.mode == 0) {
                    const uint64_t sz = nbytes - sa.byte_begin + excl + total;
                    a.scan_offsets[0] = 0;
                    a.scan_offsets[1] = sz;
                    if (sz > a.scan_capacity) atomicOr(a.err, ERRBIT_OUTPUT);
                }
            }
        }
    }
    if (sa.mode == 1) return;
    __syncthreads();
    if (t0 >= nbytes) return;
    const uint64_t out_base = sa.mode == 0 ? 0 : a.scan_offsets[img];
    uint64_t pos = out_base + chunk_excl + warp_excl + (incl - cnt) + (t0 > sa.byte_begin ? t0 - sa.byte_begin : 0);
    const uint64_t limit = a.scan_capacity;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint64_t p = t0 + 4 * j + k;
            if (p < nbytes && p >= sa.byte_begin) {
                const uint8_t byte = (uint8_t)(w[j] >> (8 * k));
                if (pos < limit) a.scan[pos] = byte;
                ++pos;
                if (byte == 0xFF) {                            // huffman.c:29-31
                    if (pos < limit) a.scan[pos] = 0x00;
                    ++pos;
                }
            }
        }
    }
}

}  // namespace jb
