// stages.cuh -- one kernel per reference core stage, for the stage-level C API
// (convertBMPToJPEGGrayscale ... performZigZag).  These exist for API parity and
// stage-level tests; the production path is the fused kernel in fused_block.cuh.
// Every stage reproduces the reference's arithmetic exactly (the DCT in the reference's
// sequential summation order).
#pragma once

#include "common.cuh"

namespace jb {

// converter.c:28-55: Y = (77R + 150G + 29B) >> 8 with clamp-to-edge padding to wp x hp
__global__ void __launch_bounds__(256)
k_stage_luma(const uint8_t *__restrict__ rgb, const int w, const int h, const int wp, const int hp,
             uint8_t *__restrict__ y)
{
    const uint64_t n = (uint64_t)wp * (uint64_t)hp;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / (uint32_t)wp), col = (int)(i - (uint64_t)row * (uint32_t)wp);
        const int sr = min(row, h - 1), sc = min(col, w - 1);
        const uint8_t *px = rgb + ((uint64_t)sr * (uint64_t)w + (uint64_t)sc) * 3u;
        y[i] = (uint8_t)((77u * px[0] + 150u * px[1] + 29u * px[2]) >> 8);
    }
}

// converter.c:83-87
__global__ void __launch_bounds__(256)
k_stage_center(const uint8_t *__restrict__ y, const uint64_t n, int8_t *__restrict__ out)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        out[i] = (int8_t)((int)y[i] - 128);
}

// dct.c:63-96,119-147: 64 threads per 8x8 block, thread (u,v) evaluates its coefficient in
// the reference's order (row-major over the block, two unfused multiplies, one add).
__global__ void __launch_bounds__(256)
k_stage_dct(const int8_t *__restrict__ img, const int wp, const int hp, float *__restrict__ coef)
{
    __shared__ int8_t tile[4][64];
    const int bw = wp >> 3;
    const uint64_t nblocks = (uint64_t)bw * (uint64_t)(hp >> 3);
    const int sub = threadIdx.x >> 6, t = threadIdx.x & 63;
    const int r = t >> 3, c = t & 7;
    for (uint64_t b0 = (uint64_t)blockIdx.x * 4; b0 < nblocks; b0 += (uint64_t)gridDim.x * 4) {
        const uint64_t b = b0 + sub;
        const bool valid = b < nblocks;
        const uint64_t by = valid ? b / (uint32_t)bw : 0, bx = valid ? b - by * (uint32_t)bw : 0;
        const uint64_t pix = (by * 8 + r) * (uint64_t)wp + bx * 8 + c;
        if (valid) tile[sub][t] = img[pix];
        __syncthreads();
        if (valid) {
            const int u = r, v = c;
            float acc = 0.0f;
#pragma unroll
            for (int rr = 0; rr < 8; ++rr) {
                const float cu = c_ref_cos[rr * 8 + u];
#pragma unroll
                for (int cc = 0; cc < 8; ++cc) {
                    float p = (float)tile[sub][rr * 8 + cc];
                    p = __fmul_rn(p, cu);
                    p = __fmul_rn(p, c_ref_cos[cc * 8 + v]);
                    acc = __fadd_rn(acc, p);
                }
            }
            coef[pix] = __fmul_rn(c_ref_scale[u * 8 + v], acc);
        }
        __syncthreads();
    }
}

// quantization.c:20-38
__global__ void __launch_bounds__(256)
k_stage_quant(const float *__restrict__ coef, const int wp, const uint64_t n, int16_t *__restrict__ q)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / (uint32_t)wp), col = (int)(i - (uint64_t)row * (uint32_t)wp);
        q[i] = (int16_t)roundf(__fdiv_rn(coef[i], c_quant_f[(row & 7) * 8 + (col & 7)]));
    }
}

// zigzag.c:43-65
__global__ void __launch_bounds__(256)
k_stage_zigzag(const int16_t *__restrict__ q, const int wp, const int hp, int16_t *__restrict__ zz)
{
    const int bw = wp >> 3;
    const uint64_t n = (uint64_t)wp * (uint64_t)hp;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t b = i >> 6;
        const int k = (int)(i & 63), pos = c_zigzag[k];
        const uint64_t by = b / (uint32_t)bw, bx = b - by * (uint32_t)bw;
        zz[i] = q[(by * 8 + (pos >> 3)) * (uint64_t)wp + bx * 8 + (pos & 7)];
    }
}

}  // namespace jb

namespace jb {

// ---------------------------------------------------------------------------------
// generic exclusive scan u32 -> u64 (one pass, decoupled look-back over 1024-item tiles)
__global__ void __launch_bounds__(256)
k_scan_u32(const uint32_t *__restrict__ in, uint64_t *__restrict__ out, const uint64_t n, uint64_t *state,
           const uint32_t epoch, uint64_t *total_out, uint32_t *err)
{
    __shared__ uint32_t warp_sums[8];
    __shared__ uint64_t tile_excl;
    const int tile = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint64_t i0 = (uint64_t)tile * 1024 + (uint64_t)tid * 4;
    uint32_t v[4], sum = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        v[j] = i0 + j < n ? in[i0 + j] : 0u;
        sum += v[j];
    }
    uint32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    uint32_t wex = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        if (w < warp) wex += warp_sums[w];
        tot += warp_sums[w];
    }
    if (warp == 0) {
        if (lane == 0) st_volatile_u64(state + tile, lb_pack(epoch, tile == 0 ? LB_PREFIX : LB_AGGREGATE, tot));
        const uint64_t excl = lookback_exclusive(state, tile, epoch, err);
        if (lane == 0) {
            if (tile != 0) st_volatile_u64(state + tile, lb_pack(epoch, LB_PREFIX, excl + tot));
            tile_excl = excl;
            if (tile == (int)gridDim.x - 1 && total_out) *total_out = excl + tot;
        }
    }
    __syncthreads();
    uint64_t off = tile_excl + wex + incl - sum;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (i0 + j < n) out[i0 + j] = off;
        off += v[j];
    }
}

// ---------------------------------------------------------------------------------
// performRLE (rle.c:51-127) on int16 zig-zag blocks: count, scan, emit.
__device__ __forceinline__ int mag_class16(int v) { const int a = v < 0 ? -v : v; return 32 - __clz(a); }

__global__ void __launch_bounds__(256)
k_rle_count(const int16_t *__restrict__ zz, const uint64_t nblocks, uint32_t *__restrict__ counts)
{
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const int16_t *c = zz + b * 64;
    int last = 0;
    for (int k = 63; k > 0; --k)
        if (c[k] != 0) { last = k; break; }
    uint32_t n = 1;
    int run = 0;
    for (int k = 1; k <= last; ++k) {
        if (c[k] == 0) { ++run; continue; }
        n += 1u + (uint32_t)(run >> 4);
        run = 0;
    }
    if (last < 63) ++n;
    counts[b] = n;
}

struct SymbolOut { uint8_t *base; };
__device__ __forceinline__ void store_symbol(uint8_t *base, uint64_t i, uint32_t sym, uint32_t amp, uint32_t nbits)
{
    // RLESymbol: u8 symbol @0, (pad @1), u16 code @2, u8 codeBits @4, (pad @5)   include/rle.h:8-14
    uint16_t *p = reinterpret_cast<uint16_t *>(base + i * 6);
    p[0] = (uint16_t)(sym & 0xFFu);
    p[1] = (uint16_t)amp;
    p[2] = (uint16_t)(nbits & 0xFFu);
}

__global__ void __launch_bounds__(256)
k_rle_emit(const int16_t *__restrict__ zz, const uint64_t nblocks, const uint64_t *__restrict__ offsets,
           uint8_t *__restrict__ symbols)
{
    const uint64_t b = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nblocks) return;
    const int16_t *c = zz + b * 64;
    uint64_t i = offsets[b];
    const int prev = b == 0 ? 0 : (int)zz[(b - 1) * 64];
    const int16_t diff = (int16_t)((int)c[0] - prev);                        // rle.c:68-70
    const int dsz = mag_class16(diff);
    store_symbol(symbols, i++, (uint32_t)dsz, (uint16_t)(diff > 0 ? diff : diff - 1), (uint32_t)dsz);   // rle.c:24-35,72-76
    int last = 0;
    for (int k = 63; k > 0; --k)
        if (c[k] != 0) { last = k; break; }
    int run = 0;
    for (int k = 1; k <= last; ++k) {
        const int v = c[k];
        if (v == 0) { ++run; continue; }
        for (; run >= 16; run -= 16) store_symbol(symbols, i++, 0xF0u, 0u, 0u);   // rle.c:99-103
        const int sz = mag_class16(v);
        store_symbol(symbols, i++, (uint32_t)((run << 4) | sz) & 0xFFu, (uint16_t)(v > 0 ? v : v - 1), (uint32_t)sz);
        run = 0;
    }
    if (last < 63) store_symbol(symbols, i++, 0x00u, 0u, 0u);                // rle.c:121-123
}

// ---------------------------------------------------------------------------------
// encodeHuffman (huffman.c:121-193) from an arbitrary symbol array.  Which table a
// symbol uses depends on the walk state (expect-DC, or coefficients consumed so far), a
// 64-state automaton; chunks of 128 symbols are summarised as state->state maps, the maps
// are composed in order, then every chunk is walked from its now-known entry state.
constexpr int HS_CHUNK = 128;

__device__ __forceinline__ int hs_step(int state, uint32_t sym)
{
    if (state == 0) return 1;                               // DC consumed           (huffman.c:145-156)
    if (sym == 0x00u) return 0;                             // EOB                   (huffman.c:176-179)
    const int c = state + (sym == 0xF0u ? 16 : (int)((sym >> 4) & 15u) + 1);   // huffman.c:180-187
    return c >= 64 ? 0 : c;
}

__global__ void __launch_bounds__(256)
k_hs_maps(const uint8_t *__restrict__ symbols, const uint64_t nsym, uint8_t *__restrict__ maps)
{
    __shared__ uint8_t sym_s[4][HS_CHUNK];
    const int sub = threadIdx.x >> 6, s0 = threadIdx.x & 63;
    const uint64_t chunk = (uint64_t)blockIdx.x * 4 + sub;
    const uint64_t first = chunk * HS_CHUNK;
    for (int i = s0; i < HS_CHUNK; i += 64) sym_s[sub][i] = first + i < nsym ? symbols[(first + i) * 6] : 0;
    __syncthreads();
    if (first >= nsym) return;
    const int n = (int)min((uint64_t)HS_CHUNK, nsym - first);
    int st = s0;
    for (int i = 0; i < n; ++i) st = hs_step(st, sym_s[sub][i]);
    maps[chunk * 64 + s0] = (uint8_t)st;
}

__global__ void k_hs_chain(const uint8_t *__restrict__ maps, const uint64_t nchunks, uint8_t *__restrict__ entry)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    int st = 0;
    for (uint64_t c = 0; c < nchunks; ++c) {
        entry[c] = (uint8_t)st;
        st = maps[c * 64 + st];
    }
}

// mode 0: count block starts per chunk; mode 1: bit cost per chunk (blocks < total_blocks only);
// mode 2: emit bits at chunk_bits[chunk] into the zeroed `packed` words (atomicOr)
__global__ void __launch_bounds__(128)
k_hs_walk(const uint8_t *__restrict__ symbols, const uint64_t nsym, const uint8_t *__restrict__ entry,
          const uint64_t *__restrict__ chunk_blocks, const uint64_t *__restrict__ chunk_bits, const uint64_t total_blocks,
          uint32_t *__restrict__ out_counts, uint32_t *__restrict__ packed, const int mode)
{
    const uint64_t chunk = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t first = chunk * HS_CHUNK;
    if (first >= nsym) return;
    const int n = (int)min((uint64_t)HS_CHUNK, nsym - first);
    int st = entry[chunk];
    uint64_t blk = mode == 0 ? 0 : chunk_blocks[chunk];     // index of the NEXT block to start
    uint64_t pos = mode == 2 ? chunk_bits[chunk] : 0;
    uint32_t acc = 0;
    for (int i = 0; i < n; ++i) {
        const uint8_t *s = symbols + (first + i) * 6;
        const uint32_t sym = s[0];
        const bool is_dc = st == 0;
        if (is_dc) ++blk;
        st = hs_step(st, sym);
        if (mode == 0) { acc += is_dc ? 1u : 0u; continue; }
        if (blk > total_blocks) continue;                   // loop bound b < totalBlocks (huffman.c:139)
        const uint32_t amp_bits = s[4];
        const uint32_t hc = is_dc ? c_dc_code[sym & 15u] : c_ac_code[sym];
        const uint32_t hlen = hc & 0xFFu;
        const uint32_t alen = (is_dc || amp_bits > 0) ? amp_bits : 0u;
        if (mode == 1) { acc += hlen + alen; continue; }
        const uint32_t amp = (uint32_t)(*reinterpret_cast<const uint16_t *>(s + 2));
        for (int part = 0; part < 2; ++part) {              // putBits(code,len) then putBits(amplitude,size)
            const uint32_t nb = part == 0 ? hlen : alen;
            if (nb == 0) continue;                          // huffman.c:36
            const uint32_t val = (part == 0 ? (hc >> 8) : amp) & ((1u << nb) - 1u);   // huffman.c:39
            const uint64_t x = ((uint64_t)val << (64u - nb)) >> (pos & 31u);
            const uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
            if (hi) atomicOr(packed + (pos >> 5), __byte_perm(hi, 0u, 0x0123u));
            if (lo) atomicOr(packed + (pos >> 5) + 1, __byte_perm(lo, 0u, 0x0123u));
            pos += nb;
        }
    }
    if (mode != 2) out_counts[chunk] = acc;
}

}  // namespace jb
