// strip_entropy.cuh -- K1b, the strip entropy kernel: DC-difference / run-length symbols (rle.c:51-127) and their
// Huffman codes + amplitude bits (huffman.c:121-193) for every 32-block strip, written as a strip-local bit stream.
//
// One warp per strip, one lane per 8x8 block, many warps per SM: the symbol walks are sparse, data-dependent and
// latency-bound, so this kernel is kept light (no transform, few registers) and runs at two to three times the
// occupancy of the block kernel.
//   1. the lane's quantized coefficients (int8, zig-zag order: 2 x 128-bit loads for positions 0..31, 2 more only for
//      the lanes K1 marked as reaching beyond; the next strip's are requested
//      before this strip is processed) are parked in shared memory and reduced to a 63-bit non-zero map;
//   2. ONE table walk over the non-zero coefficients: each visit is one look-up of the ready-made symbol word
//      (code + amplitude bits, length in the low 5 bits); the length is added to the block's bit cost and the word
//      is cached;
//   3. warp scan of the bit costs -> bit offset of every block inside the strip;
//   4. the cached words are appended to the strip's bit window in shared memory (register accumulator,
//      word-wise OR-reduction), and the window goes to the strip's slot in global memory with 128-bit stores,
//      together with the strip record {bits, first DC, last DC}.
// The DC predictor of a strip's first block is the last block of the previous strip (rle.c:59-70): one byte load.
// The very first block of an image is coded by K2 (its predictor is 0, or the previous stripe's last DC in
// multi-GPU runs).  K2 then only shifts the strips' streams to their global bit phase and stuffs.
#pragma once

#include "common.cuh"

namespace jb {

constexpr int STREAM_SMALL_BYTES = 1024;             // strip stream capacity of the default instantiation (32 B/block)
constexpr int STREAM_BIG_BYTES = 5888;               // worst case: 32 blocks x 1463 bits = 5852 bytes
constexpr int K1B_SYM_BYTES = 16384;                 // AC symbol table staged in shared memory

template <bool BIGWIN>
struct K1bCfg {
    static constexpr int WARPS = BIGWIN ? 8 : 16;                    // the 16 KB symbol table is shared by the CTA's warps
    static constexpr int THREADS = WARPS * 32;
    static constexpr int CTAS_PER_SM = BIGWIN ? 1 : 2;               // 32 warps per SM at <= 64 registers
    static constexpr int STREAM_BYTES = BIGWIN ? STREAM_BIG_BYTES : STREAM_SMALL_BYTES;
    static constexpr int WIN_BYTES = STREAM_BYTES + 128;             // + slack: the bit writer may touch one word past the end
    static constexpr int ZS_PITCH = 17;                              // words per lane of the coefficient staging area (conflict-free byte reads)
    static constexpr int ZS_BYTES = 32 * ZS_PITCH * 4;               // 2176
    static constexpr int CACHE_BYTES = 2048;                         // 16 cached symbol words per lane
    static constexpr int WARP_SMEM = ZS_BYTES + CACHE_BYTES + WIN_BYTES;
    static constexpr int SMEM = K1B_SYM_BYTES + WARPS * WARP_SMEM;
};

// ---- entropy coding helpers ------------------------------------------------------------------------------

__device__ __forceinline__ int magnitude_class(int v)          // rle.c:9-22
{
    const int a = v < 0 ? -v : v;
    return 32 - __clz(a);
}

// shared-memory accesses by 32-bit shared-space address: keeps the symbol loops free of the
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
__device__ __forceinline__ void sts_u32(uint32_t saddr, uint32_t v)
{
    asm volatile("st.shared.u32 [%0], %1;" :: "r"(saddr), "r"(v) : "memory");
}
__device__ __forceinline__ void red_or_shared(uint32_t saddr, uint32_t v, bool enable)
{
    asm volatile("{\n.reg .pred p;\nsetp.ne.u32 p, %2, 0;\n@p red.shared.or.b32 [%0], %1;\n}"
                 :: "r"(saddr), "r"(v), "r"((uint32_t)enable) : "memory");
}

// MSB-first bit appender: every symbol is OR-reduced into the zeroed shared-memory window at its absolute bit
// position (two word-wise reductions; the second one is a no-op when the symbol does not cross a word boundary).
// No register accumulator, hence no loop-carried dependency besides the position and no conditional flush.
struct BitWriter {
    uint32_t win;         // shared-space address of the window
    uint32_t pos;         // bit position of the next symbol
    __device__ __forceinline__ void start(uint32_t win_saddr, uint32_t relbit)
    {
        win = win_saddr;
        pos = relbit;
    }
    __device__ __forceinline__ void put(uint32_t vl, uint32_t n)           // n in 1..27 bits, left-aligned in vl (low bits zero)
    {
        const uint32_t sh = pos & 31u, waddr = win + ((pos >> 5) << 2);
        const uint32_t hi = vl >> sh, lo = __funnelshift_r(0u, vl, sh);       // (vl : 0) >> sh, low word = the bits that spill over
        asm volatile("red.shared.or.b32 [%0], %1;\n\tred.shared.or.b32 [%0+4], %2;" :: "r"(waddr), "r"(hi), "r"(lo) : "memory");
        pos += n;
    }
    __device__ __forceinline__ void finish() {}
};

constexpr int SYM_CACHE = 16;        // cached symbols per block (lane-interleaved words: 16 x 32 x 4 = 2048 bytes per warp)

// Walk 1 -- the ONLY table walk: the lane visits the non-zero AC coefficients of its block (map mlo/mhi, bit k <->
// zig-zag position k; rle.c:83-123).  Each visit is one look-up in the symbol table sym[run & 15][value & 255] =
// (Huffman code << size | amplitude bits) left-aligned with the total length in the low 5 bits, i.e. huffman.c:164-173
// applied to the symbol rle.c:106-113 would have produced; slot [0][0] carries EOB and [0][0x80] ZRL.  The entry's
// length is added to the block's bit cost and the entry itself is parked in the lane's symbol cache (entry j of lane l
// at word 32 j + l: conflict-free whatever j the lanes are at), so that the emit walk is a plain stream of cached
// words.  Caching stops at the first coefficient that needs a ZRL (zero run >= 16, rle.c:99-103) or does not fit the
// cache any more; that coefficient and everything after it (mask + previous position) is left to emit_tail.
struct SymWalk {
    uint32_t bits;        // AC bit cost incl. ZRLs and EOB
    uint32_t cend;        // shared-space address one past the last cached entry
    uint32_t rest_lo, rest_hi;   // non-zero positions NOT cached
    int rest_prev;        // position of the last cached non-zero; -1: everything (incl. EOB) is cached
    int last;             // position of the block's last non-zero
};

__device__ __forceinline__ SymWalk cache_symbols(uint32_t zs, uint32_t mlo, uint32_t mhi, uint32_t sym, uint32_t cptr, uint32_t climit)
{
    SymWalk w;
    const uint32_t eob = lds_u32(sym);
    uint32_t bits = 0;
    uint32_t rest_lo = 0, rest_hi = 0;
    int rest_prev = -1, prev = 0;
    // fast loops: one per half of the map; they leave through `spill` at the first coefficient that is not cached
    {
        uint32_t m = mlo;
#pragma unroll 1
        while (m) {
            const int k = __ffs((int)m) - 1;
            const uint32_t byte = lds_u8(zs + (uint32_t)k);
            const uint32_t run = (uint32_t)(k - prev - 1);
            if (run >= 16u || cptr == climit) { rest_lo = m; rest_hi = mhi; rest_prev = prev; goto spill; }
            const uint32_t e = lds_u32(sym + (((run << 8) | byte) << 2));
            bits += e & 31u;
            sts_u32(cptr, e);
            cptr += 128u;
            m &= m - 1;
            prev = k;
        }
        m = mhi;
#pragma unroll 1
        while (m) {
            const int k = 32 + __ffs((int)m) - 1;
            const uint32_t byte = lds_u8(zs + (uint32_t)k);
            const uint32_t run = (uint32_t)(k - prev - 1);
            if (run >= 16u || cptr == climit) { rest_hi = m; rest_prev = prev; goto spill; }
            const uint32_t e = lds_u32(sym + (((run << 8) | byte) << 2));
            bits += e & 31u;
            sts_u32(cptr, e);
            cptr += 128u;
            m &= m - 1;
            prev = k;
        }
    }
    if (prev < 63) {                                                                   // EOB, rle.c:121-123
        bits += eob & 31u;
        if (cptr != climit) {
            sts_u32(cptr, eob);
            cptr += 128u;
        } else {
            rest_prev = prev;                                                          // only the EOB is left for emit_tail
        }
    }
    w.last = prev;
    goto done;
spill:
    {   // rare: cost of the coefficients that are not cached (emit_tail walks them again)
        const uint32_t zrl_len = lds_u32(sym + 4u * 0x80u) & 31u;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t m = half ? rest_hi : rest_lo;
#pragma unroll 1
            while (m) {
                const int k = 32 * half + __ffs((int)m) - 1;
                m &= m - 1;
                const uint32_t byte = lds_u8(zs + (uint32_t)k);
                const uint32_t run = (uint32_t)(k - prev - 1);
                prev = k;
                bits += (run >> 4) * zrl_len + (lds_u32(sym + ((((run & 15u) << 8) | byte) << 2)) & 31u);   // ZRLs: rle.c:99-103
            }
        }
        if (prev < 63) bits += eob & 31u;
        w.last = prev;
    }
done:
    w.bits = bits;
    w.cend = cptr;
    w.rest_lo = rest_lo;
    w.rest_hi = rest_hi;
    w.rest_prev = rest_prev;
    return w;
}

// the symbols that did not fit the cache: positions in (rest_lo, rest_hi) after `prev`, then EOB
__device__ __noinline__ BitWriter emit_tail(BitWriter bw, uint32_t zs, uint32_t rest_lo, uint32_t rest_hi, int prev, uint32_t sym)
{
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        uint32_t m = half ? rest_hi : rest_lo;
#pragma unroll 1
        while (m) {
            const int k = 32 * half + __ffs((int)m) - 1;
            m &= m - 1;
            const uint32_t byte = lds_u8(zs + (uint32_t)k);
            int run = k - prev - 1;
            prev = k;
            if (run >= 16) {                                                           // ZRL, rle.c:99-103
                const uint32_t z = lds_u32(sym + 4u * 0x80u);
                do {
                    bw.put(z & ~31u, z & 31u);
                    run -= 16;
                } while (run >= 16);
            }
            const uint32_t e = lds_u32(sym + 4u * (((uint32_t)run << 8) | byte));
            bw.put(e & ~31u, e & 31u);
        }
    }
    if (prev < 63) {                                                                   // EOB, rle.c:121-123
        const uint32_t e = lds_u32(sym);
        bw.put(e & ~31u, e & 31u);
    }
    return bw;
}


// geometry + buffers of the strip entropy kernel
struct StripArgs {
    const int8_t *coef;            // [blocks][32] zig-zag positions 0..31, int8 (K1)
    const int8_t *coef_hi;         // [blocks][32] positions 32..63, valid for the lanes in himask[strip]
    const uint32_t *himask;        // [total_strips]
    const uint8_t *tables;         // device table block
    StripRec *strips;              // [total_strips]
    uint32_t *strip_bits;          // [total_strips] compact copy of StripRec.bits
    uint8_t *streams;              // [total_strips][slot_bytes]: the strip's bits, MSB first in 32-bit words
    uint32_t slot_bytes;           // multiple of 16, <= the instantiation's STREAM_BYTES
    uint32_t *dbg_blkinfo;         // optional stage tap: bit offset inside the strip | last non-zero << 16
    uint32_t *err;
    uint32_t total_strips;
    uint32_t spr, bw, bh;          // strips per block row, blocks per block row, block rows per image
    uint64_t blocks_per_image;
    // stripes: the CTA that finishes last reduces the stripe's boundary summary {first DC, last DC, bits if the first
    // block were predicted from 0} over the `summary_strips` owned strips into *summary (device memory, 16 bytes)
    void *summary;                 // nullptr: no summary
    uint32_t summary_strips;
    unsigned int *done_counter;    // zeroed before the launch (part of the look-back state K1 clears)
};

template <bool BIGWIN>
__global__ void __launch_bounds__(K1bCfg<BIGWIN>::THREADS, K1bCfg<BIGWIN>::CTAS_PER_SM)
k_strip_entropy(const StripArgs a)
{
    using Cfg = K1bCfg<BIGWIN>;
    extern __shared__ __align__(128) uint8_t smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t *wbase = smem + K1B_SYM_BYTES + warp * Cfg::WARP_SMEM;
    uint32_t *zs = reinterpret_cast<uint32_t *>(wbase) + lane * Cfg::ZS_PITCH;          // this lane's 64 coefficient bytes
    uint32_t *win = reinterpret_cast<uint32_t *>(wbase + Cfg::ZS_BYTES + Cfg::CACHE_BYTES);   // the strip's bit window
    uint32_t smem_sa = smem_u32(smem);
    asm volatile("mov.b32 %0, %0;" : "+r"(smem_sa));             // opaque: computed once, not rematerialised at every use
    const uint32_t sym_sa = smem_sa;                             // symbol table at offset 0
    const uint32_t zs_sa = smem_sa + (uint32_t)(reinterpret_cast<uint8_t *>(zs) - smem);
    const uint32_t cache_sa = smem_sa + (uint32_t)(wbase + Cfg::ZS_BYTES - smem) + 4u * (uint32_t)lane;
    const uint32_t win_sa = smem_sa + (uint32_t)(reinterpret_cast<uint8_t *>(win) - smem);
    __shared__ uint32_t s_dc[16];                                // DC codes: (code << 8) | len per size class
    __shared__ __align__(8) uint64_t s_bar;

    if (threadIdx.x == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&s_bar, K1B_SYM_BYTES);                   // 16 KB, one TMA bulk copy; awaited before the first walk
        bulk_g2s(smem, a.tables + TBL_SYM, K1B_SYM_BYTES, &s_bar);
    }
    if (threadIdx.x < 16) s_dc[threadIdx.x] = reinterpret_cast<const uint32_t *>(a.tables + TBL_DC_CODE)[threadIdx.x];
    __syncthreads();

    const uint32_t stride = gridDim.x * Cfg::WARPS;
    const uint32_t per_image = a.bh * a.spr;
    const uint32_t cap_bits = a.slot_bytes * 8u;
    bool table_ready = false;

    // strip -> (first block, blocks, first-of-image).  The position is found by division once and then advanced
    // incrementally (a warp's strips are `stride` apart).
    uint32_t s = blockIdx.x * Cfg::WARPS + warp;
    uint32_t p_img = s / per_image, p_brow, p_sx;
    {
        const uint32_t rem = s - p_img * per_image;
        p_brow = rem / a.spr;
        p_sx = rem - p_brow * a.spr;
    }
    const uint32_t dq = stride / a.spr, dr = stride - dq * a.spr;
    auto locate = [&](uint64_t &block0, uint32_t &vb, bool &first) {
        block0 = (uint64_t)p_img * a.blocks_per_image + (uint64_t)p_brow * a.bw + p_sx * 32u;
        vb = min(32u, a.bw - p_sx * 32u);
        first = (p_brow | p_sx) == 0u;
    };
    auto advance = [&]() {
        p_sx += dr;
        p_brow += dq;
        if (p_sx >= a.spr) {
            p_sx -= a.spr;
            ++p_brow;
        }
        while (p_brow >= a.bh) {
            p_brow -= a.bh;
            ++p_img;
        }
    };

    uint64_t block0 = 0;
    uint32_t vb = 0;
    bool first = false;
    uint4 q[4];
    int pred = 0;                                                // lane 0: quantized DC of the block before the strip
    uint32_t hmask = 0;                                          // lanes of the strip whose positions 32..63 are not all zero
    auto fetch = [&](uint32_t strip) {                           // the strip's coefficients -> q (positions 32..63 where present)
        hmask = __ldg(a.himask + strip);
        if ((uint32_t)lane < vb) {
            const uint4 *src = reinterpret_cast<const uint4 *>(a.coef + (block0 + lane) * 32);
            q[0] = src[0];
            q[1] = src[1];
            if ((hmask >> lane) & 1u) {
                const uint4 *srch = reinterpret_cast<const uint4 *>(a.coef_hi + (block0 + lane) * 32);
                q[2] = srch[0];
                q[3] = srch[1];
            } else {
                q[2] = q[3] = make_uint4(0u, 0u, 0u, 0u);
            }
        }
        if (lane == 0 && !first) pred = (int)a.coef[(block0 - 1) * 32];
    };
    if (s < a.total_strips) {
        locate(block0, vb, first);
        fetch(s);
    }
    for (; s < a.total_strips; s += stride) {
        const uint64_t my_block0 = block0;
        const uint32_t my_vb = vb;
        const bool my_first = first;
        uint32_t mlo = 0, mhi = 0;                               // non-zero map of the block's AC coefficients
        int my_dc = 0;
        const int prev_strip_dc = pred;
        if ((uint32_t)lane < my_vb) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                zs[4 * i] = q[i].x; zs[4 * i + 1] = q[i].y; zs[4 * i + 2] = q[i].z; zs[4 * i + 3] = q[i].w;
                mlo |= (nonzero_nibble(q[i].x) | (nonzero_nibble(q[i].y) << 4) | (nonzero_nibble(q[i].z) << 8) | (nonzero_nibble(q[i].w) << 12)) << (16 * i);
            }
            mlo &= ~1u;                                          // position 0 is the DC
            my_dc = (int)(int8_t)(q[0].x & 0xFFu);
            if (hmask) {                                         // (warp-uniform) some block of the strip reaches beyond position 31
#pragma unroll
                for (int i = 2; i < 4; ++i) {
                    zs[4 * i] = q[i].x; zs[4 * i + 1] = q[i].y; zs[4 * i + 2] = q[i].z; zs[4 * i + 3] = q[i].w;
                    mhi |= (nonzero_nibble(q[i].x) | (nonzero_nibble(q[i].y) << 4) | (nonzero_nibble(q[i].z) << 8) | (nonzero_nibble(q[i].w) << 12)) << (16 * (i - 2));
                }
            }
        }
        // request the next strip's coefficients now: their latency hides behind the walks below
        if (s + stride < a.total_strips) {
            advance();
            locate(block0, vb, first);
            fetch(s + stride);
        }
        if (!table_ready) {
            mbar_wait(&s_bar, 0);
            table_ready = true;
        }
        // DC differences (rle.c:68-76): the previous lane's block, or the block before the strip for lane 0
        bool with_dc;
        uint32_t dc_word = 0;                                    // the DC symbol: code + amplitude bits left-aligned | length
        {
            int prev_dc = __shfl_up_sync(0xffffffffu, my_dc, 1);
            if (lane == 0) prev_dc = prev_strip_dc;
            const int diff = my_dc - prev_dc;
            with_dc = (uint32_t)lane < my_vb && (lane > 0 || !my_first);
            if (with_dc) {
                const int sz = magnitude_class(diff);
                const uint32_t dc_entry = s_dc[sz];
                const uint32_t amp = (uint32_t)(diff > 0 ? diff : diff - 1) & ((1u << sz) - 1u);   // rle.c:24-35, huffman.c:39
                const uint32_t n = (dc_entry & 0xFFu) + (uint32_t)sz;                               // <= 9 + 11
                dc_word = ((((dc_entry >> 8) << sz) | amp) << (32u - n)) | n;
            }
        }
        __syncwarp();
        SymWalk sw;
        sw.bits = sw.rest_lo = sw.rest_hi = 0;
        sw.cend = cache_sa;
        sw.rest_prev = -1;
        sw.last = 0;
        uint32_t my_bits = 0;
        if ((uint32_t)lane < my_vb) {
            if (with_dc) sts_u32(cache_sa, dc_word);
            sw = cache_symbols(zs_sa, mlo, mhi, sym_sa, cache_sa + (with_dc ? 128u : 0u), cache_sa + 128u * SYM_CACHE);
            my_bits = sw.bits + (dc_word & 31u);
        }
        uint32_t incl = my_bits;
#pragma unroll
        for (int ofs = 1; ofs < 32; ofs <<= 1) {
            const uint32_t n = __shfl_up_sync(0xffffffffu, incl, ofs);
            if (lane >= ofs) incl += n;
        }
        const uint32_t strip_total = __shfl_sync(0xffffffffu, incl, 31);
        const bool fits = strip_total <= cap_bits;
        const uint32_t nquads = fits ? (strip_total + 127u) >> 7 : 0u;               // 16-byte units written to the slot
        // ---- emit the strip's bits into the zeroed window, then window -> the strip's slot ------------------------
        for (uint32_t i = lane; i < nquads + 1; i += 32) reinterpret_cast<uint4 *>(win)[i] = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        if (fits && (uint32_t)lane < my_vb) {
            BitWriter bw;
            bw.start(win_sa, incl - my_bits);
#pragma unroll 1
            for (uint32_t cp = cache_sa; cp != sw.cend;) {           // up to four symbols per trip: a quarter of the loop and reconvergence overhead
                const uint32_t left = (sw.cend - cp) >> 7;
                const uint32_t e0 = lds_u32(cp);
                const uint32_t e1 = left > 1u ? lds_u32(cp + 128u) : 0u;
                const uint32_t e2 = left > 2u ? lds_u32(cp + 256u) : 0u;
                const uint32_t e3 = left > 3u ? lds_u32(cp + 384u) : 0u;
                bw.put(e0 & ~31u, e0 & 31u);
                if (left > 1u) bw.put(e1 & ~31u, e1 & 31u);
                if (left > 2u) bw.put(e2 & ~31u, e2 & 31u);
                if (left > 3u) bw.put(e3 & ~31u, e3 & 31u);
                cp += 128u * min(left, 4u);
            }
            if (sw.rest_prev >= 0) bw = emit_tail(bw, zs_sa, sw.rest_lo, sw.rest_hi, sw.rest_prev, sym_sa);
            bw.finish();
        }
        __syncwarp();
        {
            uint4 *dst = reinterpret_cast<uint4 *>(a.streams + (uint64_t)s * a.slot_bytes);
            for (uint32_t i = lane; i < nquads; i += 32) dst[i] = reinterpret_cast<const uint4 *>(win)[i];
            const int first_dc = __shfl_sync(0xffffffffu, my_dc, 0);
            if ((uint32_t)lane == my_vb - 1) {
                a.strips[s] = StripRec{strip_total, (int16_t)first_dc, (int16_t)my_dc};
                a.strip_bits[s] = strip_total;
                if (!fits) atomicOr(a.err, ERRBIT_WORKSPACE);
            }
            if (a.dbg_blkinfo != nullptr) {                      // stage tap for the parity tests (uniform branch)
                if ((uint32_t)lane < my_vb) a.dbg_blkinfo[my_block0 + (uint32_t)lane] = (incl - my_bits) | ((uint32_t)sw.last << 16);
            }
        }
        __syncwarp();
    }

    if (a.summary != nullptr) {
        // last-CTA-done reduction (stripes): every CTA publishes its strip records, the one that arrives last sums them
        __shared__ bool s_last;
        __shared__ uint64_t s_part[Cfg::WARPS];
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) s_last = atomicAdd(a.done_counter, 1u) == gridDim.x - 1;
        __syncthreads();
        if (s_last) {
            __threadfence();
            uint64_t sum = 0;
            for (uint32_t i = threadIdx.x; i < a.summary_strips; i += Cfg::THREADS) sum += __ldcg(a.strip_bits + i);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            if (lane == 0) s_part[warp] = sum;
            __syncthreads();
            if (threadIdx.x == 0) {
                uint64_t bits = 0;
                for (int w = 0; w < Cfg::WARPS; ++w) bits += s_part[w];
                const volatile StripRec *recs = a.strips;           // written by other CTAs of this launch
                const int first_dc = recs[0].first_dc;
                const uint32_t dc_entry = s_dc[magnitude_class(first_dc)];
                bits += (dc_entry & 0xFFu) + (uint32_t)magnitude_class(first_dc);       // first DC symbol, predictor 0 (rle.c:68-76)
                struct { int16_t first_dc, last_dc; uint32_t valid; uint64_t bits_pred0; } out;
                out.first_dc = (int16_t)first_dc;
                out.last_dc = recs[a.summary_strips - 1].last_dc;
                out.valid = 1u;
                out.bits_pred0 = bits;
                *reinterpret_cast<decltype(out) *>(a.summary) = out;
            }
        }
    }
}

}  // namespace jb
