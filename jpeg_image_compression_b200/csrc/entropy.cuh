// entropy.cuh -- standalone 0xFF byte-stuffing pass over an already packed bit stream
// (huffman.c:26-32,65-81).  Used by the stage-level encodeHuffman (stage_api.inl); the
// production path stuffs inside the fused kernel of scan_pack.cuh.
#pragma once

#include "common.cuh"
#include "scan_pack.cuh"

namespace jb {

constexpr int K4_THREADS = 256;
constexpr int K4_CHUNK = K4_THREADS * 16;                      // 4096 bytes per CTA

struct EntropyArgs {
    uint64_t *image_bits;          // [count] total bits per image
    uint64_t *image_base;          // [count] byte offset of the image's packed bits in `packed` (count > 1)
    uint32_t *packed;              // unstuffed stream, memory byte order
    uint64_t packed_capacity;      // bytes
    uint64_t *stuff_state;         // [count*chunks_cap] look-back state
    uint64_t *image_ff;            // [count] number of 0xFF bytes per image
    uint8_t *scan;                 // output: stuffed bytes
    uint64_t scan_capacity;
    uint64_t *scan_offsets;        // [count+1]
    uint32_t *err;
    int chunks_cap;
    int count;
    uint32_t epoch;
    uint32_t bit_phase;
};

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
