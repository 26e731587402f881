// entropy.cuh -- standalone 0xFF byte-stuffing pass over an already packed bit stream
// (huffman.c:26-32,65-81).  Used only by the stage-level encodeHuffman (stage_api.inl); the
// production path stuffs inside the merge kernel of scan_pack.cuh.
#pragma once

#include "common.cuh"
#include "scan_pack.cuh"

namespace jb {

constexpr int K4_THREADS = 256;
constexpr int K4_CHUNK = K4_THREADS * 16;                      // 4096 bytes per CTA

struct EntropyArgs {
    uint64_t *image_bits;          // [1] total bits of the stream
    uint32_t *packed;              // unstuffed stream, memory byte order
    uint64_t packed_capacity;      // bytes
    uint64_t *stuff_state;         // [chunks] look-back state
    uint8_t *scan;                 // output: stuffed bytes
    uint64_t scan_capacity;
    uint64_t *scan_offsets;        // [2]: {0, stuffed size}
    uint32_t *err;
    uint32_t epoch;
};

// count + decoupled look-back + write in one pass: one CTA per 4096 packed bytes
__global__ void __launch_bounds__(K4_THREADS)
k_stuff(const EntropyArgs a)
{
    __shared__ uint32_t warp_sums[K4_THREADS / 32];
    __shared__ uint64_t chunk_excl;
    const int chunk = blockIdx.x;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t nbytes = (a.image_bits[0] + 7) >> 3;
    const uint64_t c0 = (uint64_t)chunk * K4_CHUNK;
    if (c0 >= nbytes) return;
    if (nbytes + 3 > a.packed_capacity) return;                  // flagged earlier
    const uint64_t t0 = c0 + (uint64_t)tid * 16;

    uint4 d = make_uint4(0, 0, 0, 0);
    if (t0 < nbytes) d = *reinterpret_cast<const uint4 *>(reinterpret_cast<const uint8_t *>(a.packed) + t0);
    uint32_t w[4] = {d.x, d.y, d.z, d.w};
    // bytes at or beyond nbytes are not part of the stream: blank them
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint64_t p = t0 + 4 * j;
        if (p + 4 > nbytes) {
            uint32_t keep = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k)
                if (p + k < nbytes) keep |= 0xFFu << (8 * k);
            w[j] &= keep;
        }
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
    const bool last_chunk = c0 + K4_CHUNK >= nbytes;
    if (warp == 0) {
        if (lane == 0)
            st_volatile_u64(a.stuff_state + chunk, lb_pack(a.epoch, chunk == 0 ? LB_PREFIX : LB_AGGREGATE, total));
        const uint64_t excl = lookback_exclusive(a.stuff_state, chunk, a.epoch, a.err);
        if (lane == 0) {
            if (chunk != 0) st_volatile_u64(a.stuff_state + chunk, lb_pack(a.epoch, LB_PREFIX, excl + total));
            chunk_excl = excl;
            if (last_chunk) {
                const uint64_t sz = nbytes + excl + total;
                a.scan_offsets[0] = 0;
                a.scan_offsets[1] = sz;
                if (sz > a.scan_capacity) atomicOr(a.err, ERRBIT_OUTPUT);
            }
        }
    }
    __syncthreads();
    if (t0 >= nbytes) return;
    uint64_t pos = chunk_excl + warp_excl + (incl - cnt) + t0;
    const uint64_t limit = a.scan_capacity;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint64_t p = t0 + 4 * j + k;
            if (p < nbytes) {
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
