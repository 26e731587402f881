// jpegb200.cu -- host side of libjpegb200.so: device tables, encoder handle, workspace,
// kernel orchestration and the C ABI declared in include/jpegb200.h.
//
// The single translation unit that contains every kernel (fused_block.cuh,
// entropy.cuh, stages.cuh, synth.cuh).  sm_100a only; no CPU fallback anywhere: if a
// CUDA call fails the entry point returns NULL / an error code.
#include "jpegb200.h"

#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "common.cuh"
#include "scan_pack.cuh"
#include "entropy.cuh"
#include "fused_block.cuh"
#include "strip_entropy.cuh"
#include "stages.cuh"
#include "synth.cuh"

namespace jb {

static thread_local std::string g_last_error;

static bool cuda_ok(cudaError_t e, const char *what)
{
    if (e == cudaSuccess) return true;
    g_last_error = std::string(what) + ": " + cudaGetErrorString(e);
    return false;
}
#define JB_CUDA(call)                                      \
    do {                                                   \
        if (!cuda_ok((call), #call)) return JPEGB200_ERR_CUDA; \
    } while (0)

// ---- table generation ----------------------------------------------------------

static const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                    12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                    35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                    58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// Reference cosine LUT, 6 decimals, asymmetric last digits preserved (dct.c:9-18).
static const float kRefCos[64] = {
    1.000000f, 0.980785f,  0.923880f,  0.831470f,  0.707107f,  0.555570f,  0.382683f,  0.195090f,
    1.000000f, 0.831470f,  0.382683f,  -0.195090f, -0.707107f, -0.980785f, -0.923880f, -0.555570f,
    1.000000f, 0.555570f,  -0.382683f, -0.980785f, -0.707107f, 0.195090f,  0.923880f,  0.831470f,
    1.000000f, 0.195090f,  -0.923880f, -0.555570f, 0.707107f,  0.831470f,  -0.382683f, -0.980785f,
    1.000000f, -0.195090f, -0.923880f, 0.555570f,  0.707107f,  -0.831470f, -0.382684f, 0.980785f,
    1.000000f, -0.555570f, -0.382684f, 0.980785f,  -0.707107f, -0.195090f, 0.923880f,  -0.831470f,
    1.000000f, -0.831470f, 0.382684f,  0.195091f,  -0.707107f, 0.980785f,  -0.923879f, 0.555570f,
    1.000000f, -0.980785f, 0.923880f,  -0.831470f, 0.707107f,  -0.555570f, 0.382684f,  -0.195090f};

struct HuffCode { uint16_t code; uint8_t len; };

// canonical code assignment from BITS/HUFFVAL (huffman.c:89-104)
static void canonical(const unsigned char *counts, const unsigned char *values, HuffCode *tab, int n)
{
    for (int i = 0; i < n; ++i) tab[i] = HuffCode{0, 0};
    uint16_t next = 0;
    int vi = 0;
    for (int len = 1; len <= 16; ++len) {
        for (int i = 0; i < counts[len - 1]; ++i, ++vi) tab[values[vi]] = HuffCode{next++, (uint8_t)len};
        next = (uint16_t)(next << 1);
    }
}

struct HostTables {
    float ref_scale[64], quant_f[64], rk[64];
    float rk_tc[64];                 // zig-zag order: ref_scale / (Q * 2^21), tensor-core path
    double tc_weight_sum_err;        // max over coefficients of |sum_i (W_i / 2^21 - w_i)|  
    uint8_t dc_len[16];
    uint32_t dc_code[16], ac_code[256];
    std::vector<uint8_t> block;      // device table block, TBL_* layout (common.cuh)
};

static int bit_length(int v) { int a = v < 0 ? -v : v, n = 0; while (a) { ++n; a >>= 1; } return n; }

// fp16 bit pattern of an integer |v| <= 2048 (exactly representable)
static uint16_t half_of_int(int v)
{
    if (v == 0) return 0;
    const uint16_t sign = v < 0 ? 0x8000u : 0u;
    const unsigned a = (unsigned)(v < 0 ? -v : v);
    int e = 0;
    while ((a >> (e + 1)) != 0) ++e;                                 // floor(log2 a) <= 11
    const unsigned mant = e <= 10 ? (a << (10 - e)) & 0x3FFu : (a >> (e - 10)) & 0x3FFu;
    return (uint16_t)(sign | ((unsigned)(e + 15) << 10) | mant);
}

static void build_tables(HostTables &t)
{
    const double g[3] = {1.0, std::cos(M_PI / 8), std::cos(M_PI / 4)};
    for (int u = 0; u < 8; ++u) {
        for (int v = 0; v < 8; ++v) {
            const float cu = u == 0 ? 0.707107f : 1.000000f, cv = v == 0 ? 0.707107f : 1.000000f;   // dct.c:4-6
            volatile float s1 = 0.25f * cu;                    // dct.c:93, each product rounded to fp32
            volatile float s2 = s1 * cv;
            t.ref_scale[u * 8 + v] = s2;
            t.quant_f[u * 8 + v] = (float)std_luminance_quant_tbl[u * 8 + v];
            t.rk[u * 8 + v] = (float)((double)s2 * g[gclass(u)] * g[gclass(v)] / (double)std_luminance_quant_tbl[u * 8 + v]);
        }
    }
    HuffCode dc[16], ac[256];
    canonical(std_dc_luminance_nrcodes, std_dc_luminance_values, dc, 16);
    canonical(std_ac_luminance_nrcodes, std_ac_luminance_values, ac, 256);
    for (int i = 0; i < 16; ++i) {
        t.dc_len[i] = (uint8_t)(dc[i].len + i);
        t.dc_code[i] = ((uint32_t)dc[i].code << 8) | dc[i].len;
    }
    for (int i = 0; i < 256; ++i) t.ac_code[i] = ((uint32_t)ac[i].code << 8) | ac[i].len;
    t.block.assign(TBL_BYTES, 0);
    // ready-made AC symbols: [run 0..15][int8 value] -> (code << size | amplitude) left-aligned | length in the low
    // 5 bits (rle.c:24-35,106-113 + huffman.c:39,164-173).  Value 0 never occurs as a coefficient, so slot
    // [0][0] carries EOB; ZRL lives in [0][0x80] (-128 cannot occur, |q| <= 95).
    {
        uint32_t *sym = reinterpret_cast<uint32_t *>(&t.block[TBL_SYM]);
        for (int run = 0; run < 16; ++run) {
            for (int b = 1; b < 256; ++b) {
                const int v = (int)(int8_t)b, sz = bit_length(v);
                if (sz > 7) continue;
                const HuffCode hc = ac[(run << 4) | sz];
                const uint32_t amp = (uint32_t)(v > 0 ? v : v - 1) & ((1u << sz) - 1u);
                const uint32_t len = (uint32_t)(hc.len + sz);                          // <= 16 + 7
                sym[run * 256 + b] = ((((uint32_t)hc.code << sz) | amp) << (32u - len)) | len;
            }
        }
        sym[0] = ((uint32_t)ac[0x00].code << (32 - ac[0x00].len)) | ac[0x00].len;      // EOB
        sym[0x80] = ((uint32_t)ac[0xF0].code << (32 - ac[0xF0].len)) | ac[0xF0].len;   // ZRL
    }
    memcpy(&t.block[TBL_DC_CODE], t.dc_code, sizeof(t.dc_code));
    // Tensor-core transform: B[n][k], n = (zig-zag position, limb), k = 8 * pixel row + pixel column.
    // Weight = the reference's own LUT product cos[r][u] * cos[c][v] (dct.c:79-84, exact in double) in 22-bit fixed
    // point, W = round(w * 2^21) = l0 * 2^11 + l1 with balanced 11-bit limbs: integers that fp16 holds exactly.
    // Column order inside a zig-zag word (4 positions a b c d): S0(a) S0(b) S1(a) S1(b) S0(c) S0(d) S1(c) S1(d), so that a
    // lane's tcgen05.ld delivers packed-fp32 operand pairs.  Memory layout = UMMA K-major, no swizzle:
    // [k/16][n/8][(k%16)/8][n%8][k%8] halves (core matrix = 8 rows x 16 bytes; LBO 128, SBO 256, 4 KB per K step).
    {
        uint16_t *bm = reinterpret_cast<uint16_t *>(&t.block[TBL_BMAT]);
        t.tc_weight_sum_err = 0.0;
        for (int k = 0; k < 64; ++k) {
            const int zpos = kZigzag[k], u = zpos >> 3, v = zpos & 7;
            t.rk_tc[k] = (float)((double)t.ref_scale[zpos] / ((double)std_luminance_quant_tbl[zpos] * 2097152.0));
            double sum_err = 0.0;
            for (int r = 0; r < 8; ++r) {
                for (int c = 0; c < 8; ++c) {
                    const double w = (double)kRefCos[r * 8 + u] * (double)kRefCos[c * 8 + v];
                    const long long W = llround(w * 2097152.0);
                    long long l1 = ((W + 1024) % 2048 + 2048) % 2048 - 1024;         // [-1024, 1023]
                    long long l0 = (W - l1) / 2048;                                   // [-1024, 1024]
                    sum_err += (double)W / 2097152.0 - w;
                    const int kidx = r * 8 + c, ks = kidx / 16, kc = (kidx % 16) / 8, ke = kidx % 8;
                    for (int limb = 0; limb < 2; ++limb) {
                        const int n = 8 * (k >> 2) + 4 * ((k & 3) >> 1) + 2 * limb + (k & 1);
                        bm[ks * 2048 + (n / 8) * 128 + kc * 64 + (n % 8) * 8 + ke] = half_of_int((int)(limb ? l1 : l0));
                    }
                }
            }
            if (k > 0) t.tc_weight_sum_err = std::max(t.tc_weight_sum_err, std::fabs(sum_err));
        }
    }
}

// ---- encoder handle ---------------------------------------------------------------

// Workspace buffer.  With JPEGB200_GUARD=1 in the environment (a debugging aid: compute-sanitizer is not available on
// every box) a buffer is allocated at exactly the requested size between two 4 KB guard bands, everything filled with
// 0xA5; jpegb200_encoder_check_guards counts the guard bytes that no longer hold the pattern.
constexpr size_t GUARD_BYTES = 4096;
struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
    void *base = nullptr;            // guard mode: start of the allocation (ptr = base + GUARD_BYTES)
    int reserve(size_t need)
    {
        if (need <= bytes) return JPEGB200_OK;
        release();
        const char *g = getenv("JPEGB200_GUARD");
        if (g && *g && *g != '0') {
            const size_t want = (need + 15) & ~(size_t)15;
            if (!cuda_ok(cudaMalloc(&base, want + 2 * GUARD_BYTES), "cudaMalloc(workspace)")) return JPEGB200_ERR_CUDA;
            if (!cuda_ok(cudaMemset(base, 0xA5, want + 2 * GUARD_BYTES), "cudaMemset(guard)")) return JPEGB200_ERR_CUDA;
            ptr = static_cast<uint8_t *>(base) + GUARD_BYTES;
            bytes = want;
            return JPEGB200_OK;
        }
        const size_t want = need + need / 8 + 256;
        if (!cuda_ok(cudaMalloc(&ptr, want), "cudaMalloc(workspace)")) return JPEGB200_ERR_CUDA;
        bytes = want;
        return JPEGB200_OK;
    }
    void release()
    {
        if (base) cudaFree(base);
        else if (ptr) cudaFree(ptr);
        ptr = base = nullptr;
        bytes = 0;
    }
    // guard mode: number of guard bytes overwritten (-1: CUDA error); 0 for an unguarded buffer
    long long corrupted() const
    {
        if (!base) return 0;
        std::vector<uint8_t> h(2 * GUARD_BYTES);
        if (cudaMemcpy(h.data(), base, GUARD_BYTES, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        if (cudaMemcpy(h.data() + GUARD_BYTES, static_cast<uint8_t *>(ptr) + bytes, GUARD_BYTES, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
        long long bad = 0;
        for (uint8_t b : h) bad += b != 0xA5;
        return bad;
    }
};

}  // namespace jb

struct jpegb200_encoder {
    int device = 0;
    int sm_count = 148;
    int dct_mode = 0;
    int concurrency = 1;             // handles the caller keeps busy on this device at the same time (grid sizing)
    int bytes_per_block = 24;
    jb::HostTables tables;
    jb::DeviceBuffer coef, himask, blkinfo, streams, strips, strip_bits, lookback, image_bits, image_bytes, slots, dtables, misc, host_in, host_scan, trace, trace1;
    uint64_t *pinned_status = nullptr;   // host-pinned {err, pad, offsets[2]}: one D2H + one sync per host call
    uint64_t last_scan_bytes = 0;        // size of the previous host-call result: the speculative D2H length
    uint32_t slot_bytes = 768;      // strip stream slot: 32 blocks x bytes_per_block
    bool want_taps = false;         // K1 also stores the stage taps (coefficients, per-block bit offsets)
    bool taps_valid = false;        // ... and did so in the last launch
    // last launch
    jb::PackArgs args{};
    jb::Geom geom{};
    uint64_t lookback_words = 0;
    uint64_t total_blocks = 0;      // blocks K1 produced (all images, incl. stripe halo)
    uint64_t launches = 0;
    int k2_ctas_per_sm[2] = {0, 0};   // [0] resident merge-kernel CTAs per SM, [1] the slot size that was computed for
    int k1_grid = 0, k1_warps = 0;   // shape of the last K1 launch
    bool stripe_ready = false;
    bool stripe_state_fresh = false;   // K2's look-back state is still the zeroed one K1 left (no memset before the first stripe encode)
    // optional per-kernel timing (cudaEvents on the launching stream)
    bool profiling = false;
    std::vector<cudaEvent_t> events;        // 2 per timed kernel of the last launch
    std::vector<int> event_kernel;          // kernel id of each event pair
    double kernel_ms[8] = {0};
    uint64_t kernel_calls[8] = {0};
};

namespace jb {

static std::mutex g_tables_mutex;
static bool g_tables_uploaded[64] = {false};

static int upload_tables(jpegb200_encoder *enc)
{
    HostTables &t = enc->tables;
    build_tables(t);
    {
        std::lock_guard<std::mutex> lock(g_tables_mutex);
        if (!g_tables_uploaded[enc->device & 63]) {
            JB_CUDA(cudaMemcpyToSymbol(c_ref_cos, kRefCos, sizeof(kRefCos)));
            JB_CUDA(cudaMemcpyToSymbol(c_ref_scale, t.ref_scale, sizeof(t.ref_scale)));
            JB_CUDA(cudaMemcpyToSymbol(c_quant_f, t.quant_f, sizeof(t.quant_f)));
            JB_CUDA(cudaMemcpyToSymbol(c_rk, t.rk, sizeof(t.rk)));
            JB_CUDA(cudaMemcpyToSymbol(c_rk_tc, t.rk_tc, sizeof(t.rk_tc)));
            JB_CUDA(cudaMemcpyToSymbol(c_zigzag, kZigzag, sizeof(kZigzag)));
            JB_CUDA(cudaMemcpyToSymbol(c_dc_len, t.dc_len, sizeof(t.dc_len)));
            JB_CUDA(cudaMemcpyToSymbol(c_dc_code, t.dc_code, sizeof(t.dc_code)));
            JB_CUDA(cudaMemcpyToSymbol(c_ac_code, t.ac_code, sizeof(t.ac_code)));
            JB_CUDA(cudaFuncSetAttribute(k_fused_blocks<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1Cfg<true>::SMEM));
            JB_CUDA(cudaFuncSetAttribute(k_fused_blocks<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1Cfg<false>::SMEM));
            JB_CUDA(cudaFuncSetAttribute(k_strip_entropy<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1bCfg<false>::SMEM));
            JB_CUDA(cudaFuncSetAttribute(k_strip_entropy<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, K1bCfg<true>::SMEM));
            JB_CUDA(cudaFuncSetAttribute(k_merge_stuff, cudaFuncAttributeMaxDynamicSharedMemorySize, k2_smem(STREAM_BIG_BYTES)));
            g_tables_uploaded[enc->device & 63] = true;
        }
    }
    int rc = enc->dtables.reserve(TBL_BYTES);
    if (rc) return rc;
    JB_CUDA(cudaMemcpy(enc->dtables.ptr, t.block.data(), TBL_BYTES, cudaMemcpyHostToDevice));
    rc = enc->misc.reserve(256);
    if (rc) return rc;
    JB_CUDA(cudaMemset(enc->misc.ptr, 0, 256));
    return JPEGB200_OK;
}

static uint32_t *misc_err(jpegb200_encoder *e) { return reinterpret_cast<uint32_t *>(e->misc.ptr); }
static unsigned long long *misc_flagged(jpegb200_encoder *e)
{
    return reinterpret_cast<unsigned long long *>(static_cast<uint8_t *>(e->misc.ptr) + 8);
}
static uint64_t *misc_offsets(jpegb200_encoder *e)
{
    return reinterpret_cast<uint64_t *>(static_cast<uint8_t *>(e->misc.ptr) + 16);
}

// Fill geometry + workspace for `count` images of w x h.  halo_rows > 0 (stripes): the last
// halo_rows pixel rows belong to the NEXT stripe; K1 transforms them (one extra block row) so
// that K2 can complete this stripe's last byte, but they are not part of this stripe's stream.
static int prepare(jpegb200_encoder *enc, const uint8_t *d_rgb, int w, int h, int count, uint64_t stride, int halo_rows,
                   int64_t row_pitch = 0, bool bgr = false)
{
    if (!enc || !d_rgb || w <= 0 || h <= 0 || count <= 0 || halo_rows < 0 || halo_rows > 8 || (halo_rows && (h & 7))) {
        g_last_error = "bad argument";
        return JPEGB200_ERR_ARG;
    }
    if (stride == 0) stride = (uint64_t)w * (uint64_t)(h + halo_rows) * 3u;
    JB_CUDA(cudaSetDevice(enc->device));
    Geom &g = enc->geom;
    g.rgb = d_rgb;
    g.image_stride = stride;
    g.row_pitch = row_pitch ? row_pitch : 3 * (int64_t)w;
    g.wt_lo = bgr ? 0x004D961Du : 0x001D964Du;          // bytes (77,150,29,0) resp. (29,150,77,0): converter.c:51
    g.wt_hi = g.wt_lo << 8;
    g.w = w;
    g.h = h + halo_rows;
    g.bw = (w + 7) / 8;
    g.bh = (h + 7) / 8 + (halo_rows ? 1 : 0);
    g.spr = (g.bw + 31) / 32;
    g.count = count;
    g.blocks_per_image = (uint64_t)g.bw * (uint64_t)g.bh;
    g.total_strips = (uint64_t)g.spr * (uint64_t)g.bh * (uint64_t)count;
    if (g.total_strips >= (1ull << 31)) {
        g_last_error = "launch too large: split the batch (strip index is 32-bit)";
        return JPEGB200_ERR_ARG;
    }
    const uint64_t tb = g.blocks_per_image * (uint64_t)count;
    const uint32_t strips_avail = (uint32_t)g.spr * (uint32_t)g.bh;
    const uint32_t strips_owned = (uint32_t)g.spr * (uint32_t)((h + 7) / 8);
    const uint64_t nb_owned = (uint64_t)g.bw * (uint64_t)((h + 7) / 8);
    enc->total_blocks = tb;
    const int tiles = (int)((strips_owned + K2_TILE_STRIPS - 1) / K2_TILE_STRIPS);
    int rc = 0;
    // strip stream slots: 32 blocks x bytes_per_block (the hint), at most the per-strip worst case
    enc->slot_bytes = (uint32_t)std::min(32 * enc->bytes_per_block, enc->bytes_per_block <= 32 ? STREAM_SMALL_BYTES : STREAM_BIG_BYTES);
    enc->slot_bytes = (enc->slot_bytes + 15u) & ~15u;
    if ((rc = enc->streams.reserve(g.total_strips * (uint64_t)enc->slot_bytes + 64))) return rc;
    if ((rc = enc->coef.reserve(tb * 64))) return rc;                 // two planes: positions 0..31, then 32..63 (sparse)
    if ((rc = enc->himask.reserve(g.total_strips * 4 + 16))) return rc;
    if (enc->want_taps && (rc = enc->blkinfo.reserve(tb * 4))) return rc;
    if ((rc = enc->strips.reserve(g.total_strips * sizeof(StripRec)))) return rc;
    if ((rc = enc->strip_bits.reserve(g.total_strips * 4 + 16))) return rc;
    // grouped look-back state of K2 (aggregate per tile + inclusive prefix per 1024-tile group), once
    // for the bit offsets and once for the stuffed-zero counts; contiguous, cleared by K1's prologue
    const int groups = (tiles + LB_GROUP - 1) / LB_GROUP;
    // ... followed by K2's tile counter in a 128-byte line of its own (it is hammered with atomics)
    const uint64_t state_words = (((uint64_t)tiles + 2ull * (uint64_t)groups) * (uint64_t)count + 15) & ~15ull;
    enc->lookback_words = state_words + 16;
    if ((rc = enc->lookback.reserve(enc->lookback_words * 8))) return rc;
    if ((rc = enc->image_bits.reserve((uint64_t)count * 8))) return rc;
    if ((rc = enc->image_bytes.reserve((uint64_t)count * 8))) return rc;
    // batch: every image is stuffed into its own slot, then compacted; stuffed size <= 2 x packed size
    const uint64_t slot = ((2 * (nb_owned * (uint64_t)enc->bytes_per_block + 64)) + 15) & ~15ull;
    if (count > 1 && (rc = enc->slots.reserve(slot * (uint64_t)count + 64))) return rc;

    PackArgs &a = enc->args;
    a.tables = static_cast<const uint8_t *>(enc->dtables.ptr);
    a.streams = static_cast<const uint8_t *>(enc->streams.ptr);
    a.slot_bytes = enc->slot_bytes;
    a.strips = static_cast<const StripRec *>(enc->strips.ptr);
    a.strip_bits = static_cast<const uint32_t *>(enc->strip_bits.ptr);
    a.bit_incl = static_cast<uint64_t *>(enc->lookback.ptr);
    a.ff_agg = a.bit_incl + (uint64_t)groups * (uint64_t)count;
    a.ff_incl = a.ff_agg + (uint64_t)tiles * (uint64_t)count;
    a.tile_counter = reinterpret_cast<unsigned long long *>(static_cast<uint64_t *>(enc->lookback.ptr) + state_words);
    a.out = count > 1 ? static_cast<uint8_t *>(enc->slots.ptr) : nullptr;
    a.out_capacity = count > 1 ? slot : 0;
    a.out_slot = count > 1 ? slot : 0;
    a.image_bytes = static_cast<uint64_t *>(enc->image_bytes.ptr);
    a.image_bits = static_cast<uint64_t *>(enc->image_bits.ptr);
    a.scan_offsets = nullptr;
    a.err = misc_err(enc);
    a.strips_owned = strips_owned;
    a.strips_avail = strips_avail;
    a.tiles = tiles;
    a.count = count;
    a.dc_pred0 = 0;
    a.bit_phase = 0;
    a.dyn_all = nullptr;
    a.dyn_rank = 0;
    a.trace = nullptr;
    if (getenv("JPEGB200_K2_TRACE")) {              // tuning aid: per-tile phase timestamps
        if ((rc = enc->trace.reserve((uint64_t)tiles * count * 64))) return rc;
        a.trace = static_cast<unsigned long long *>(enc->trace.ptr);
    }
    return JPEGB200_OK;
}

// ---- kernel launches (optionally bracketed by cudaEvents for per-kernel timing) -----------

enum KernelId { KID_BLOCK = 0, KID_ENTROPY = 1, KID_LAYOUT = 2, KID_COMPACT = 3, KID_FRAME = 4, KID_STRIP = 5 };

struct TimedLaunch {
    jpegb200_encoder *enc;
    cudaStream_t st;
    cudaEvent_t stop = nullptr;
    TimedLaunch(jpegb200_encoder *e, cudaStream_t s, int kid) : enc(e), st(s)
    {
        ++enc->launches;
        if (!enc->profiling) return;
        cudaEvent_t start = nullptr;
        if (cudaEventCreate(&start) != cudaSuccess || cudaEventCreate(&stop) != cudaSuccess) { stop = nullptr; return; }
        enc->events.push_back(start);
        enc->events.push_back(stop);
        enc->event_kernel.push_back(kid);
        cudaEventRecord(start, st);
    }
    ~TimedLaunch() { if (stop) cudaEventRecord(stop, st); }
};

// cuTensorMapEncodeTiled, looked up through the runtime (no link-time dependency on libcuda)
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                      const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static TensorMapEncodeFn tensor_map_encoder()
{
    static TensorMapEncodeFn fn = []() -> TensorMapEncodeFn {
        if (getenv("JPEGB200_NO_TMAP")) return nullptr;          // tuning aid: force the per-row copies
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<TensorMapEncodeFn>(p);
    }();
    return fn;
}

// Describe the pixel array as a 3-D tensor of 32-bit words (words per row, rows, images) so that K1 can fetch a
// strip's 8 x 768-byte tile with one TMA copy.  Needs 16-byte aligned base, pitch and image stride and top-down
// rows; otherwise (or without the driver entry point) K1 falls back to one bulk copy per pixel row.
static bool make_tensor_map(const Geom &g, CUtensorMap *tm)
{
    TensorMapEncodeFn enc = tensor_map_encoder();
    if (!enc || g.row_pitch <= 0 || (g.row_pitch & 15) || ((uintptr_t)g.rgb & 15) || (g.count > 1 && (g.image_stride & 15)))
        return false;
    if ((uint64_t)g.row_pitch * (uint64_t)g.h > g.image_stride && g.count > 1) return false;
    const cuuint64_t dims[3] = {(cuuint64_t)(g.row_pitch / 4), (cuuint64_t)g.h, (cuuint64_t)g.count};
    const cuuint64_t strides[2] = {(cuuint64_t)g.row_pitch, g.count > 1 ? (cuuint64_t)g.image_stride : (cuuint64_t)g.row_pitch * (cuuint64_t)g.h};
    const cuuint32_t box[3] = {TMAP_ROW_BYTES / 4, 8, 1}, estr[3] = {1, 1, 1};
    if (strides[1] & 15) return false;
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t *>(g.rgb), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// transform selection: dct_mode 0 = tensor-core DCT (default), 1 = every coefficient in reference order,
// 2 = register butterfly DCT (round-1 transform, kept for comparison); JPEGB200_DCT=butterfly forces 2 for mode 0
static bool use_tensor_dct(const jpegb200_encoder *enc)
{
    static const bool env_butterfly = []() { const char *e = getenv("JPEGB200_DCT"); return e && !strcmp(e, "butterfly"); }();
    return enc->dct_mode == 2 ? false : !env_butterfly;
}

// CTA slots of the device this handle sizes its persistent kernels for: all of them, or -- when the caller has said
// that `concurrency` handles are kept busy side by side (jpegb200_encoder_set_concurrency) -- about 1.2 / sqrt(n) of
// them, so that launches of different handles run next to one another instead of queueing behind each other's tails
// (measured at 3840x2160: 8 handles 407 -> 480 Gpixel/s, 16 handles -> 498; single-image latency 34 -> 53 / 68 us).
static uint64_t device_share(const jpegb200_encoder *enc, int ctas_per_sm)
{
    const uint64_t slots = (uint64_t)enc->sm_count * (uint64_t)std::max(1, ctas_per_sm);
    if (enc->concurrency <= 1) return slots;
    const double share = std::min(1.0, 1.2 / std::sqrt((double)enc->concurrency));
    return std::max<uint64_t>(1, (uint64_t)std::llround((double)slots * share));
}

// Grid of a kernel whose CTAs stride over `want` CTA-sized pieces of work when `cap` CTAs may run: with
// r = ceil(want / cap) pieces per CTA, ceil(want / r) CTAs finish at the same time as `cap` would (3840x2160: 127
// CTAs of two tiles per group instead of 148 of which 42 idle through the second round) and leave the other SMs to
// concurrent launches.
static int balanced_grid(uint64_t want, uint64_t cap)
{
    if (want <= cap) return (int)std::max<uint64_t>(want, 1);
    const uint64_t rounds = (want + cap - 1) / cap;
    return (int)((want + rounds - 1) / rounds);
}

template <bool TC>
static int launch_block_kernel_t(jpegb200_encoder *enc, cudaStream_t st, const CUtensorMap &tmap, bool stats)
{
    using Cfg = K1Cfg<TC>;
    const Geom &g = enc->geom;
    const uint64_t units = TC ? (g.total_strips + 3) / 4 : g.total_strips;            // 4-strip tiles / strips
    const uint64_t per_cta = TC ? Cfg::GROUPS : Cfg::WARPS;
    const uint64_t want = (units + per_cta - 1) / per_cta;
    int grid = balanced_grid(want, device_share(enc, Cfg::CTAS_PER_SM));
    if (const char *e = getenv("JPEGB200_K1_GRID")) grid = (int)std::min<uint64_t>(want, (uint64_t)std::max(1, atoi(e)));   // tuning aid
    enc->k1_grid = grid;
    enc->k1_warps = Cfg::WARPS;
    if (getenv("JPEGB200_K1_TRACE")) {              // tuning aid: per-warp timestamps
        if (int rc = enc->trace1.reserve((uint64_t)grid * Cfg::WARPS * 256)) return rc;     // [warps][8] phases, [warps][8] strip ends, 2 x [warps][8] phases of the 4th tile
        JB_CUDA(cudaMemsetAsync(enc->trace1.ptr, 0, (uint64_t)grid * Cfg::WARPS * 256, st));
    }
    {
        TimedLaunch t(enc, st, KID_BLOCK);
        k_fused_blocks<TC><<<grid, Cfg::THREADS, Cfg::SMEM, st>>>(g, static_cast<int8_t *>(enc->coef.ptr), static_cast<uint32_t *>(enc->himask.ptr), static_cast<const uint8_t *>(enc->dtables.ptr),
                                                                   stats ? misc_flagged(enc) : nullptr, enc->dct_mode == 1 ? 1 : 0,
                                                                   static_cast<uint64_t *>(enc->lookback.ptr), enc->lookback_words,
                                                                   static_cast<unsigned long long *>(enc->trace1.ptr), tmap);
    }
    JB_CUDA(cudaGetLastError());
    return JPEGB200_OK;
}

// K1: fused block kernel (persistent) -> int8 zig-zag coefficients
static int launch_block_kernel(jpegb200_encoder *enc, cudaStream_t st, bool stats = true)
{
    Geom &g = enc->geom;
    CUtensorMap tmap;
    memset(&tmap, 0, sizeof(tmap));
    g.use_tmap = make_tensor_map(g, &tmap) ? 1 : 0;
    return use_tensor_dct(enc) ? launch_block_kernel_t<true>(enc, st, tmap, stats) : launch_block_kernel_t<false>(enc, st, tmap, stats);
}

// K1b: strip entropy kernel (persistent) -> strip bit streams + strip records
static int launch_strip_entropy(jpegb200_encoder *enc, cudaStream_t st, bool taps, void *d_summary = nullptr)
{
    const Geom &g = enc->geom;
    StripArgs sa;
    sa.coef = static_cast<const int8_t *>(enc->coef.ptr);
    sa.coef_hi = sa.coef + enc->total_blocks * 32;
    sa.himask = static_cast<const uint32_t *>(enc->himask.ptr);
    sa.tables = static_cast<const uint8_t *>(enc->dtables.ptr);
    sa.strips = static_cast<StripRec *>(enc->strips.ptr);
    sa.strip_bits = static_cast<uint32_t *>(enc->strip_bits.ptr);
    sa.streams = static_cast<uint8_t *>(enc->streams.ptr);
    sa.slot_bytes = enc->slot_bytes;
    sa.dbg_blkinfo = taps ? static_cast<uint32_t *>(enc->blkinfo.ptr) : nullptr;
    sa.err = misc_err(enc);
    sa.total_strips = (uint32_t)g.total_strips;
    sa.spr = (uint32_t)g.spr;
    sa.bw = (uint32_t)g.bw;
    sa.bh = (uint32_t)g.bh;
    sa.blocks_per_image = g.blocks_per_image;
    sa.summary = d_summary;                         // stripes: reduced by the CTA that finishes last
    sa.summary_strips = enc->args.strips_owned;
    // a spare word of the line that holds K2's tile ticket: K1's prologue has zeroed it
    sa.done_counter = reinterpret_cast<unsigned int *>(static_cast<uint64_t *>(enc->lookback.ptr) + (enc->lookback_words - 8));
    enc->taps_valid = taps;
    const bool big = enc->slot_bytes > (uint32_t)STREAM_SMALL_BYTES;
    const uint64_t warps = big ? K1bCfg<true>::WARPS : K1bCfg<false>::WARPS;
    const uint64_t want = (g.total_strips + warps - 1) / warps;
    const int per_sm = big ? K1bCfg<true>::CTAS_PER_SM : K1bCfg<false>::CTAS_PER_SM;
    int grid = balanced_grid(want, device_share(enc, per_sm));
    if (const char *e = getenv("JPEGB200_K1B_GRID")) grid = (int)std::min<uint64_t>(want, (uint64_t)std::max(1, atoi(e)));   // tuning aid
    {
        TimedLaunch t(enc, st, KID_STRIP);
        if (big) k_strip_entropy<true><<<grid, K1bCfg<true>::THREADS, K1bCfg<true>::SMEM, st>>>(sa);
        else k_strip_entropy<false><<<grid, K1bCfg<false>::THREADS, K1bCfg<false>::SMEM, st>>>(sa);
    }
    JB_CUDA(cudaGetLastError());
    return JPEGB200_OK;
}

// K2: scan + shift-merge + stuff over tiles of 16 strips; as many CTAs as can be co-resident (at most one per tile),
// drawing tiles by ticket
static int launch_entropy(jpegb200_encoder *enc, cudaStream_t st)
{
    const PackArgs &a = enc->args;
    const int smem = k2_smem((int)enc->slot_bytes), win_words = k2_win_words((int)enc->slot_bytes);
    int &per_sm = enc->k2_ctas_per_sm[0];
    if (per_sm == 0 || enc->k2_ctas_per_sm[1] != (int)enc->slot_bytes) {     // occupancy at this launch's window size
        int n = 0;
        JB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_merge_stuff, K2_THREADS, smem));
        per_sm = n > 0 ? n : 1;
        enc->k2_ctas_per_sm[1] = (int)enc->slot_bytes;
    }
    const uint64_t total = (uint64_t)a.tiles * (uint64_t)a.count;
    // tiles are drawn by ticket, so any grid finishes together; a handle that shares the device keeps to two CTA slots per SM
    unsigned grid = (unsigned)std::min<uint64_t>(total, device_share(enc, enc->concurrency > 1 ? std::min(per_sm, 2) : per_sm));
    if (const char *e = getenv("JPEGB200_K2_GRID")) grid = std::max(1u, std::min(grid, (unsigned)atoi(e)));   // tuning aid
    {
        TimedLaunch t(enc, st, KID_ENTROPY);
        k_merge_stuff<<<grid, K2_THREADS, smem, st>>>(a, win_words);
    }
    JB_CUDA(cudaGetLastError());
    return JPEGB200_OK;
}

// fold finished event pairs into the per-kernel accumulators (requires the stream to be idle)
static void harvest_events(jpegb200_encoder *enc)
{
    for (size_t i = 0; i < enc->event_kernel.size(); ++i) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, enc->events[2 * i], enc->events[2 * i + 1]) == cudaSuccess) {
            enc->kernel_ms[enc->event_kernel[i]] += ms;
            enc->kernel_calls[enc->event_kernel[i]] += 1;
        }
        cudaEventDestroy(enc->events[2 * i]);
        cudaEventDestroy(enc->events[2 * i + 1]);
    }
    enc->events.clear();
    enc->event_kernel.clear();
}

// files: frame every scan with the JFIF header and EOI (d_scan then receives complete files)
static int encode_launch(jpegb200_encoder *enc, uint8_t *d_scan, uint64_t scan_capacity, uint64_t *d_scan_offsets,
                         cudaStream_t st, bool files = false)
{
    int rc = 0;
    enc->launches = 0;
    enc->stripe_ready = false;
    PackArgs &a = enc->args;
    a.scan_offsets = d_scan_offsets;
    const uint64_t lead = files ? 328 : 0, extra = files ? 330 : 0;
    if (a.count == 1) {
        a.out = d_scan + lead;
        a.out_capacity = scan_capacity > extra ? scan_capacity - extra : 0;
    }
    if ((rc = launch_block_kernel(enc, st))) return rc;
    if ((rc = launch_strip_entropy(enc, st, enc->want_taps))) return rc;
    if ((rc = launch_entropy(enc, st))) return rc;
    if (a.count > 1) {
        {
            TimedLaunch t(enc, st, KID_LAYOUT);
            k_layout<<<1, 1024, 0, st>>>(a.image_bytes, d_scan_offsets, a.count, scan_capacity, a.err, extra);
        }
        {
            TimedLaunch t(enc, st, KID_COMPACT);
            // 16 bytes per thread and iteration; a slot is sized for the worst case, typical images fill a ninth of it
            const unsigned gx = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((a.out_slot / 128 + 255) / 256, 32));
            k_compact<<<dim3(gx, (unsigned)a.count), 256, 0, st>>>(a.out, a.out_slot, a.image_bytes, d_scan_offsets, d_scan,
                                                                   scan_capacity, lead);
        }
        JB_CUDA(cudaGetLastError());
    }
    if (files) {
        JfifHeaderBytes hb;
        jpegb200_jfif_header(enc->geom.w, enc->geom.h, hb.b);
        TimedLaunch t(enc, st, KID_FRAME);
        k_frame_files<<<(unsigned)a.count, 128, 0, st>>>(hb, a.image_bytes, d_scan_offsets, d_scan, scan_capacity,
                                                         a.count == 1 ? 1 : 0, a.err);
        JB_CUDA(cudaGetLastError());
    }
    return JPEGB200_OK;
}

// Stage taps (zig-zag coefficients, per-block bit offsets) are not part of the production data flow any more: K1
// keeps the coefficients on chip.  The tap readers re-run K1 on the last launch's geometry with the taps enabled
// (the input must still be resident), unless that launch already stored them.
static int ensure_taps(jpegb200_encoder *enc)
{
    JB_CUDA(cudaSetDevice(enc->device));
    JB_CUDA(cudaDeviceSynchronize());
    if (enc->taps_valid) return JPEGB200_OK;
    if (!enc->coef.ptr || enc->total_blocks == 0) {
        g_last_error = "no previous launch";
        return JPEGB200_ERR_ARG;
    }
    int rc = 0;
    if ((rc = enc->blkinfo.reserve(enc->total_blocks * 4))) return rc;
    const uint64_t launches = enc->launches;
    rc = launch_strip_entropy(enc, nullptr, /*taps=*/true);      // same coefficients, same streams, plus the per-block offsets
    enc->launches = launches;
    if (rc) return rc;
    JB_CUDA(cudaDeviceSynchronize());
    return JPEGB200_OK;
}

}  // namespace jb

using namespace jb;

// ---- C ABI: device-resident encoder ---------------------------------------------------

extern "C" int jpegb200_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

extern "C" const char *jpegb200_last_error(void) { return g_last_error.c_str(); }

static std::vector<jb::DeviceBuffer *> all_buffers(jpegb200_encoder *enc)
{
    return {&enc->coef, &enc->himask, &enc->blkinfo, &enc->streams, &enc->strips, &enc->strip_bits, &enc->lookback, &enc->image_bits, &enc->image_bytes,
            &enc->slots, &enc->dtables, &enc->misc, &enc->host_in, &enc->host_scan, &enc->trace, &enc->trace1};
}

extern "C" jpegb200_encoder *jpegb200_encoder_create(int device)
{
    int n = 0;
    if (!cuda_ok(cudaGetDeviceCount(&n), "cudaGetDeviceCount") || device < 0 || device >= n) {
        if (g_last_error.empty()) g_last_error = "no such CUDA device";
        return nullptr;
    }
    if (!cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return nullptr;
    cudaDeviceProp prop;
    if (!cuda_ok(cudaGetDeviceProperties(&prop, device), "cudaGetDeviceProperties")) return nullptr;
    if (prop.major < 10) {
        g_last_error = "libjpegb200 is built for sm_100a (B200) only";
        return nullptr;
    }
    jpegb200_encoder *enc = new jpegb200_encoder();
    enc->device = device;
    enc->sm_count = prop.multiProcessorCount;
    if (upload_tables(enc) != JPEGB200_OK) {
        jpegb200_encoder_destroy(enc);
        return nullptr;
    }
    return enc;
}

extern "C" void jpegb200_encoder_destroy(jpegb200_encoder *enc)
{
    if (!enc) return;
    cudaSetDevice(enc->device);
    cudaDeviceSynchronize();
    harvest_events(enc);
    for (DeviceBuffer *b : all_buffers(enc)) b->release();
    if (enc->pinned_status) cudaFreeHost(enc->pinned_status);
    delete enc;
}

// Debugging aid (JPEGB200_GUARD=1): guard bytes around the workspace buffers that were overwritten, and how many
// buffers carry guards.  Synchronises the device.
extern "C" int jpegb200_encoder_check_guards(jpegb200_encoder *enc, uint64_t *corrupted_bytes, int *guarded_buffers)
{
    if (!enc || !corrupted_bytes || !guarded_buffers) return JPEGB200_ERR_ARG;
    JB_CUDA(cudaSetDevice(enc->device));
    JB_CUDA(cudaDeviceSynchronize());
    uint64_t bad = 0;
    int n = 0;
    for (jb::DeviceBuffer *b : all_buffers(enc)) {
        if (!b->base) continue;
        const long long c = b->corrupted();
        if (c < 0) return JPEGB200_ERR_CUDA;
        bad += (uint64_t)c;
        ++n;
    }
    *corrupted_bytes = bad;
    *guarded_buffers = n;
    return JPEGB200_OK;
}

extern "C" int jpegb200_encoder_set_concurrency(jpegb200_encoder *enc, int handles)
{
    if (!enc || handles < 1 || handles > 1024) return JPEGB200_ERR_ARG;
    enc->concurrency = handles;
    return JPEGB200_OK;
}

extern "C" int jpegb200_encoder_set_dct_mode(jpegb200_encoder *enc, int dct_mode)
{
    if (!enc || dct_mode < 0 || dct_mode > 2) return JPEGB200_ERR_ARG;
    enc->dct_mode = dct_mode;
    return JPEGB200_OK;
}

extern "C" int jpegb200_encoder_set_bytes_per_block(jpegb200_encoder *enc, int bytes_per_block)
{
    if (!enc || bytes_per_block < 1) return JPEGB200_ERR_ARG;
    enc->bytes_per_block = bytes_per_block > 184 ? 184 : bytes_per_block;   // 1463 bits is the per-block maximum
    return JPEGB200_OK;
}

extern "C" int jpegb200_encoder_set_profiling(jpegb200_encoder *enc, int on)
{
    if (!enc) return JPEGB200_ERR_ARG;
    enc->profiling = on != 0;
    return JPEGB200_OK;
}

extern "C" int jpegb200_encoder_kernel_times(jpegb200_encoder *enc, double ms_total[8], uint64_t calls[8], int reset)
{
    if (!enc || !ms_total || !calls) return JPEGB200_ERR_ARG;
    JB_CUDA(cudaSetDevice(enc->device));
    JB_CUDA(cudaDeviceSynchronize());
    harvest_events(enc);
    for (int i = 0; i < 8; ++i) {
        ms_total[i] = enc->kernel_ms[i];
        calls[i] = enc->kernel_calls[i];
        if (reset) { enc->kernel_ms[i] = 0; enc->kernel_calls[i] = 0; }
    }
    return JPEGB200_OK;
}

extern "C" int jpegb200_encode_batch_device(jpegb200_encoder *enc, const jpegb200_batch *batch, uint8_t *d_scan,
                                            uint64_t scan_capacity, uint64_t *d_scan_offsets, void *cuda_stream)
{
    if (!enc || !batch || !d_scan || !d_scan_offsets) {
        g_last_error = "bad argument";
        return JPEGB200_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    int rc = prepare(enc, batch->d_rgb, batch->width, batch->height, batch->count, batch->image_stride, 0);
    if (rc) return rc;
    return encode_launch(enc, d_scan, scan_capacity, d_scan_offsets, st);
}

// Same, but the output holds complete JFIF files (header + scan + EOI per image), back to back.
extern "C" int jpegb200_encode_batch_files_device(jpegb200_encoder *enc, const jpegb200_batch *batch, uint8_t *d_files,
                                                  uint64_t capacity, uint64_t *d_file_offsets, void *cuda_stream)
{
    if (!enc || !batch || !d_files || !d_file_offsets) {
        g_last_error = "bad argument";
        return JPEGB200_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    int rc = prepare(enc, batch->d_rgb, batch->width, batch->height, batch->count, batch->image_stride, 0);
    if (rc) return rc;
    return encode_launch(enc, d_files, capacity, d_file_offsets, st, true);
}

// Device error word (sticky until read): 0 = ok, else a JPEGB200_ERR_*.  Synchronises the stream.
extern "C" int jpegb200_encoder_status(jpegb200_encoder *enc, void *cuda_stream)
{
    if (!enc) return JPEGB200_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    uint32_t err = 0;
    JB_CUDA(cudaSetDevice(enc->device));
    JB_CUDA(cudaMemcpyAsync(&err, misc_err(enc), 4, cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaStreamSynchronize(st));
    if (err) JB_CUDA(cudaMemsetAsync(misc_err(enc), 0, 4, st));
    if (err & ERRBIT_LOOKBACK) { g_last_error = "look-back spin limit hit"; return JPEGB200_ERR_INTERNAL; }
    if (err & ERRBIT_WORKSPACE) { g_last_error = "packed-bits workspace too small"; return JPEGB200_ERR_WORKSPACE; }
    if (err & ERRBIT_OUTPUT) { g_last_error = "scan buffer too small"; return JPEGB200_ERR_OUTPUT; }
    return JPEGB200_OK;
}

// Counters: flagged_coefficients accumulates since the last call (read-and-reset).
extern "C" int jpegb200_encoder_stats(jpegb200_encoder *enc, jpegb200_stats *out)
{
    if (!enc || !out) return JPEGB200_ERR_ARG;
    JB_CUDA(cudaSetDevice(enc->device));
    JB_CUDA(cudaDeviceSynchronize());
    unsigned long long flagged = 0;
    JB_CUDA(cudaMemcpy(&flagged, misc_flagged(enc), 8, cudaMemcpyDeviceToHost));
    JB_CUDA(cudaMemset(misc_flagged(enc), 0, 8));
    out->blocks = enc->total_blocks;
    out->flagged_coefficients = flagged;
    out->kernel_launches = enc->launches;
    out->packed_bytes = 0;
    if (enc->args.count > 0 && enc->image_bits.ptr && !enc->stripe_ready) {
        std::vector<uint64_t> bits((size_t)enc->args.count);
        JB_CUDA(cudaMemcpy(bits.data(), enc->image_bits.ptr, bits.size() * 8, cudaMemcpyDeviceToHost));
        for (uint64_t b : bits) out->packed_bytes += (b + 7) / 8;
    }
    return JPEGB200_OK;
}

// Coefficients of blocks [first, first + n) of the last launch as int16 in zig-zag order, assembled from the two planes
// and the strips' masks (positions 32..63 of a block whose mask bit is clear were not stored: they are zero).
static int read_block_coefficients(jpegb200_encoder *enc, uint64_t first, uint64_t n, int16_t *host_zz)
{
    if (first + n > enc->total_blocks) return JPEGB200_ERR_ARG;
    const Geom &g = enc->geom;
    std::vector<int8_t> lo((size_t)n * 32), hi((size_t)n * 32);
    std::vector<uint32_t> mask((size_t)g.total_strips);
    const int8_t *base = static_cast<const int8_t *>(enc->coef.ptr);
    JB_CUDA(cudaMemcpy(lo.data(), base + first * 32, lo.size(), cudaMemcpyDeviceToHost));
    JB_CUDA(cudaMemcpy(hi.data(), base + enc->total_blocks * 32 + first * 32, hi.size(), cudaMemcpyDeviceToHost));
    JB_CUDA(cudaMemcpy(mask.data(), enc->himask.ptr, mask.size() * 4, cudaMemcpyDeviceToHost));
    for (uint64_t i = 0; i < n; ++i) {
        const uint64_t b = first + i, img = b / g.blocks_per_image, rem = b - img * g.blocks_per_image;
        const uint64_t brow = rem / (uint64_t)g.bw, bx = rem - brow * (uint64_t)g.bw;
        const uint64_t strip = (img * (uint64_t)g.bh + brow) * (uint64_t)g.spr + bx / 32;
        const bool has_hi = (mask[strip] >> (bx & 31)) & 1u;
        for (int k = 0; k < 32; ++k) {
            host_zz[i * 64 + k] = lo[i * 32 + k];
            host_zz[i * 64 + 32 + k] = has_hi ? hi[i * 32 + k] : 0;
        }
    }
    return JPEGB200_OK;
}

extern "C" int jpegb200_encoder_read_coefficients(jpegb200_encoder *enc, int16_t *host_zz, uint64_t nblocks)
{
    if (!enc || !host_zz || nblocks > enc->total_blocks) return JPEGB200_ERR_ARG;
    JB_CUDA(cudaSetDevice(enc->device));
    JB_CUDA(cudaDeviceSynchronize());
    int rc = read_block_coefficients(enc, 0, nblocks, host_zz);
    if (rc) return rc;
    return JPEGB200_OK;
}

// per-block bit cost, reconstructed from K1's strip-local offsets and strip records exactly the
// way K2 consumes them
// tuning aid (JPEGB200_K2_TRACE=1): per-tile phase timestamps of the last K2 launch, [tiles][8] ns
extern "C" int jpegb200_encoder_read_k1_trace(jpegb200_encoder *enc, uint64_t *host, uint64_t nwarps)
{
    if (!enc || !host || !enc->trace1.ptr || nwarps * 64 > enc->trace1.bytes) return JPEGB200_ERR_ARG;
    JB_CUDA(cudaSetDevice(enc->device));
    JB_CUDA(cudaDeviceSynchronize());
    JB_CUDA(cudaMemcpy(host, enc->trace1.ptr, nwarps * 64, cudaMemcpyDeviceToHost));
    return JPEGB200_OK;
}

// shape of the last block-kernel launch (tuning aids)
extern "C" int jpegb200_encoder_launch_shape(jpegb200_encoder *enc, int *k1_grid, int *k1_warps_per_cta)
{
    if (!enc || !k1_grid || !k1_warps_per_cta) return JPEGB200_ERR_ARG;
    *k1_grid = enc->k1_grid;
    *k1_warps_per_cta = enc->k1_warps;
    return JPEGB200_OK;
}

extern "C" int jpegb200_encoder_read_trace(jpegb200_encoder *enc, uint64_t *host, uint64_t ntiles)
{
    if (!enc || !host || !enc->trace.ptr || ntiles * 64 > enc->trace.bytes) return JPEGB200_ERR_ARG;
    JB_CUDA(cudaSetDevice(enc->device));
    JB_CUDA(cudaDeviceSynchronize());
    JB_CUDA(cudaMemcpy(host, enc->trace.ptr, ntiles * 64, cudaMemcpyDeviceToHost));
    return JPEGB200_OK;
}

extern "C" int jpegb200_encoder_read_block_bits(jpegb200_encoder *enc, uint32_t *host_bits, uint64_t nblocks)
{
    if (!enc || !host_bits || nblocks > enc->total_blocks) return JPEGB200_ERR_ARG;
    if (int rc = ensure_taps(enc)) return rc;
    const PackArgs &a = enc->args;
    const uint32_t spr = (uint32_t)enc->geom.spr, bw = (uint32_t)enc->geom.bw;
    const uint64_t nb_avail = enc->geom.blocks_per_image;
    std::vector<uint32_t> info((size_t)enc->total_blocks);
    std::vector<StripRec> recs((size_t)a.strips_avail * a.count);
    JB_CUDA(cudaMemcpy(info.data(), enc->blkinfo.ptr, info.size() * 4, cudaMemcpyDeviceToHost));
    JB_CUDA(cudaMemcpy(recs.data(), a.strips, recs.size() * sizeof(StripRec), cudaMemcpyDeviceToHost));
    uint64_t done = 0;
    for (int img = 0; img < a.count && done < nblocks; ++img) {
        for (uint32_t s = 0; s < a.strips_avail && done < nblocks; ++s) {
            const uint32_t brow = s / spr, sx = s % spr;
            const uint32_t vb = std::min<uint32_t>(32u, bw - sx * 32u);
            const uint64_t b0 = (uint64_t)img * nb_avail + (uint64_t)brow * bw + sx * 32u;
            const StripRec &r = recs[(size_t)img * a.strips_avail + s];
            // only the image's first DC symbol is not in K1's counts (its predictor is a run-time argument)
            const uint32_t fix = s == 0 ? enc->tables.dc_len[bit_length((int)r.first_dc - (int)a.dc_pred0)] : 0u;
            for (uint32_t l = 0; l < vb; ++l) {
                const uint32_t cur = info[b0 + l] & 0xFFFFu;
                const uint32_t nxt = l + 1 < vb ? (info[b0 + l + 1] & 0xFFFFu) : r.bits;
                if (b0 + l < nblocks) { host_bits[b0 + l] = nxt - cur + (l == 0 ? fix : 0u); ++done; }
            }
        }
    }
    return JPEGB200_OK;
}

// ---- C ABI: MCU-row stripes -------------------------------------------------------------

// misc layout (256 bytes): [0] uint32 err, [8] uint64 flagged counter, [16] uint64[2] scan offsets,
// [64] stripe summary (16 bytes)
static StripeSummaryDev *misc_summary(jpegb200_encoder *e) { return reinterpret_cast<StripeSummaryDev *>(static_cast<uint8_t *>(e->misc.ptr) + 64); }

extern "C" int jpegb200_stripe_analyze_device(jpegb200_encoder *enc, const uint8_t *d_rgb, int width, int stripe_height,
                                              int halo_rows, jpegb200_stripe_summary *d_out, void *cuda_stream)
{
    if (!d_out) return JPEGB200_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    int rc = prepare(enc, d_rgb, width, stripe_height, 1, 0, halo_rows);
    if (rc) return rc;
    enc->launches = 0;
    if ((rc = launch_block_kernel(enc, st))) return rc;
    if ((rc = launch_strip_entropy(enc, st, enc->want_taps, d_out))) return rc;
    enc->stripe_ready = true;
    enc->stripe_state_fresh = true;                 // K1's prologue has just cleared K2's look-back state
    return JPEGB200_OK;
}

extern "C" int jpegb200_stripe_analyze(jpegb200_encoder *enc, const uint8_t *d_rgb, int width, int stripe_height,
                                       int halo_rows, jpegb200_stripe_summary *host_out, void *cuda_stream)
{
    if (!enc || !host_out) return JPEGB200_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    int rc = jpegb200_stripe_analyze_device(enc, d_rgb, width, stripe_height, halo_rows,
                                            reinterpret_cast<jpegb200_stripe_summary *>(misc_summary(enc)), cuda_stream);
    if (rc) return rc;
    JB_CUDA(cudaMemcpyAsync(host_out, misc_summary(enc), sizeof(jpegb200_stripe_summary), cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaStreamSynchronize(st));
    return JPEGB200_OK;
}

// K2 for a stripe; d_all == nullptr: predictor and phase come from enc->args (host values), otherwise the kernel derives
// them from all ranks' summaries in device memory
static int stripe_encode_launch(jpegb200_encoder *enc, const StripeSummaryDev *d_all, int rank, uint8_t *d_scan, uint64_t scan_capacity,
                                uint64_t *d_scan_info, cudaStream_t st)
{
    PackArgs &a = enc->args;
    a.dyn_all = d_all;
    a.dyn_rank = rank;
    a.out = d_scan;
    a.out_capacity = scan_capacity;
    a.out_slot = 0;
    a.scan_offsets = d_scan_info;
    if (!enc->stripe_state_fresh)                   // a second encode after one analyze: clear K2's look-back state again
        JB_CUDA(cudaMemsetAsync(enc->lookback.ptr, 0, enc->lookback_words * 8, st));
    enc->stripe_state_fresh = false;
    return launch_entropy(enc, st);
}

extern "C" int jpegb200_stripe_encode_device(jpegb200_encoder *enc, const jpegb200_stripe_summary *d_all, int world, int rank,
                                             uint8_t *d_scan, uint64_t scan_capacity, uint64_t *d_scan_info, void *cuda_stream)
{
    if (!enc || !d_all || !d_scan || !d_scan_info || world <= 0 || rank < 0 || rank >= world || !enc->stripe_ready) {
        g_last_error = "jpegb200_stripe_encode_device: bad argument, or no preceding jpegb200_stripe_analyze";
        return JPEGB200_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    JB_CUDA(cudaSetDevice(enc->device));
    return stripe_encode_launch(enc, reinterpret_cast<const StripeSummaryDev *>(d_all), rank, d_scan, scan_capacity, d_scan_info, st);
}

extern "C" int jpegb200_stripe_encode(jpegb200_encoder *enc, int16_t dc_predictor, uint64_t bit_begin, uint8_t *d_scan,
                                      uint64_t scan_capacity, uint64_t *host_scan_bytes, void *cuda_stream)
{
    if (!enc || !d_scan || !host_scan_bytes || !enc->stripe_ready) {
        g_last_error = "jpegb200_stripe_encode without a preceding jpegb200_stripe_analyze";
        return JPEGB200_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    JB_CUDA(cudaSetDevice(enc->device));
    enc->args.dc_pred0 = dc_predictor;
    enc->args.bit_phase = (uint32_t)(bit_begin & 7u);
    int rc = stripe_encode_launch(enc, nullptr, 0, d_scan, scan_capacity, misc_offsets(enc), st);
    if (rc) return rc;
    uint64_t offs[2] = {0, 0};
    JB_CUDA(cudaMemcpyAsync(offs, misc_offsets(enc), 16, cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaStreamSynchronize(st));
    if ((rc = jpegb200_encoder_status(enc, cuda_stream))) return rc;
    *host_scan_bytes = offs[1];
    return JPEGB200_OK;
}

extern "C" int jpegb200_synth_rgb_device(uint8_t *d_rgb, int width, int height, int count, uint64_t image_stride,
                                         uint32_t seed0, int amp, void *cuda_stream)
{
    if (!d_rgb || width <= 0 || height <= 0 || count <= 0 || amp < 0) return JPEGB200_ERR_ARG;
    if (image_stride == 0) image_stride = (uint64_t)width * (uint64_t)height * 3u;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    const uint64_t per_image = (uint64_t)width * (uint64_t)height;
    const unsigned gx = (unsigned)std::min<uint64_t>((per_image + 255) / 256, 65535u * 16u);
    k_synth_rgb<<<dim3(gx, (unsigned)count), 256, 0, st>>>(d_rgb, width, height, image_stride, seed0, amp);
    JB_CUDA(cudaGetLastError());
    return JPEGB200_OK;
}

// ---- C ABI: host buffers in, host buffers out ----------------------------------------------
//
// The reference-facing call: RGB in host memory -> stuffed scan bytes in host memory.  The
// handle keeps its device staging buffers between calls; H2D and D2H are inside the call.
// If the caller's buffers are pinned (cudaHostAlloc / cudaHostRegister) the copies run at
// PCIe speed, otherwise the driver stages them.

// Common tail of the host entries: ONE stream synchronisation in the usual case.  The error word and the scan size
// (32 bytes at misc[0..32)) go to a pinned status block, and the scan bytes are copied speculatively in the same
// breath -- as many as the previous call produced plus a margin; only if this image turned out larger is the
// remainder fetched with a second copy.
static int fetch_scan_to_host(jpegb200_encoder *enc, const uint8_t *d_scan, uint64_t d_capacity, uint8_t *host_scan, uint64_t host_capacity,
                              uint64_t *host_scan_bytes, cudaStream_t st, void *cuda_stream)
{
    if (!enc->pinned_status) JB_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&enc->pinned_status), 64, cudaHostAllocDefault));
    uint64_t guess = enc->last_scan_bytes ? enc->last_scan_bytes + enc->last_scan_bytes / 8 + 4096 : d_capacity / 8;
    guess = std::min(std::min(guess, d_capacity), host_capacity);
    JB_CUDA(cudaMemcpyAsync(enc->pinned_status, enc->misc.ptr, 32, cudaMemcpyDeviceToHost, st));
    if (guess) JB_CUDA(cudaMemcpyAsync(host_scan, d_scan, guess, cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaStreamSynchronize(st));
    const uint32_t err = (uint32_t)enc->pinned_status[0];
    const uint64_t n = enc->pinned_status[3];
    if (err) return jpegb200_encoder_status(enc, cuda_stream);
    *host_scan_bytes = n;
    if (n > host_capacity) {
        g_last_error = "host scan buffer too small";
        return JPEGB200_ERR_OUTPUT;
    }
    enc->last_scan_bytes = n;
    if (n > guess) {
        JB_CUDA(cudaMemcpyAsync(host_scan + guess, d_scan + guess, n - guess, cudaMemcpyDeviceToHost, st));
        JB_CUDA(cudaStreamSynchronize(st));
    }
    return JPEGB200_OK;
}

extern "C" int jpegb200_encode_host(jpegb200_encoder *enc, const uint8_t *host_rgb, int width, int height,
                                    uint8_t *host_scan, uint64_t host_capacity, uint64_t *host_scan_bytes,
                                    void *cuda_stream)
{
    if (!enc || !host_rgb || !host_scan || !host_scan_bytes || width <= 0 || height <= 0) {
        g_last_error = "bad argument";
        return JPEGB200_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    JB_CUDA(cudaSetDevice(enc->device));
    const size_t nrgb = (size_t)width * (size_t)height * 3u;
    const uint64_t nb = (uint64_t)((width + 7) / 8) * (uint64_t)((height + 7) / 8);
    const uint64_t cap = 2 * (nb * (uint64_t)enc->bytes_per_block + 64);   // stuffed <= 2 x packed
    int rc = 0;
    if ((rc = enc->host_in.reserve(nrgb + 64))) return rc;
    if ((rc = enc->host_scan.reserve(cap + 64))) return rc;
    uint8_t *d_rgb = static_cast<uint8_t *>(enc->host_in.ptr);
    uint8_t *d_scan = static_cast<uint8_t *>(enc->host_scan.ptr);
    uint64_t *d_off = misc_offsets(enc);
    JB_CUDA(cudaMemcpyAsync(d_rgb, host_rgb, nrgb, cudaMemcpyHostToDevice, st));
    if ((rc = prepare(enc, d_rgb, width, height, 1, 0, 0))) return rc;
    if ((rc = encode_launch(enc, d_scan, cap, d_off, st))) return rc;
    return fetch_scan_to_host(enc, d_scan, cap, host_scan, host_capacity, host_scan_bytes, st, cuda_stream);
}

// ---- C ABI: BMP file image in, complete JPEG file out ---------------
//
// Opt-in fast path next to the reference-shaped one: the host does NOT run loadBMPImage's per-pixel
// BGR->RGB / vertical-flip pass (bmp_handler.c:103-124).  The file's pixel array is copied to the
// device as it is and K1 walks the rows in place (bottom-up => negative row pitch, 4-byte padded
// rows, BGR channel order => permuted DP4A weights).  Header checks and failure cases are those of
// loadBMPImage (bmp_handler.c:23-49,68-75,88,104).  The output is the whole file of
// saveJPEGGrayscale: 328 header bytes (jpeg_handler.c:220-233) + scan + FFD9.
namespace jb {
static uint32_t le32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
}

extern "C" int jpegb200_encode_bmp_to_jpeg_host(jpegb200_encoder *enc, const uint8_t *bmp, uint64_t bmp_bytes,
                                                uint8_t *jpeg_out, uint64_t jpeg_capacity, uint64_t *jpeg_bytes,
                                                int *width, int *height, void *cuda_stream)
{
    if (!enc || !bmp || !jpeg_out || !jpeg_bytes) {
        g_last_error = "bad argument";
        return JPEGB200_ERR_ARG;
    }
    if (bmp_bytes < 54 || bmp[0] != 'B' || bmp[1] != 'M') { g_last_error = "not a valid BMP file"; return JPEGB200_ERR_ARG; }
    if ((bmp[28] | (bmp[29] << 8)) != 24) { g_last_error = "only 24-bit BMP images are supported"; return JPEGB200_ERR_ARG; }
    if (le32(bmp + 30) != 0) { g_last_error = "compressed BMP images are not supported"; return JPEGB200_ERR_ARG; }
    const int32_t w = (int32_t)le32(bmp + 18);
    int32_t h = (int32_t)le32(bmp + 22);
    const bool bottom_up = h >= 0;                                   // bmp_handler.c:68-72
    if (h < 0) h = -h;
    if (w <= 0 || h <= 0) { g_last_error = "empty BMP image"; return JPEGB200_ERR_ARG; }
    const uint64_t pitch = ((uint64_t)w * 3u + 3u) & ~3ull;          // bmp_handler.c:75
    const uint64_t off = le32(bmp + 10), need = pitch * (uint64_t)h; // bmp_handler.c:88
    if (off > bmp_bytes || need > bmp_bytes - off) { g_last_error = "insufficient pixel data in BMP file"; return JPEGB200_ERR_ARG; }
    if (width) *width = w;
    if (height) *height = h;

    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    JB_CUDA(cudaSetDevice(enc->device));
    const uint64_t nb = (uint64_t)((w + 7) / 8) * (uint64_t)((h + 7) / 8);
    const uint64_t cap = 2 * (nb * (uint64_t)enc->bytes_per_block + 64);
    int rc = 0;
    if ((rc = enc->host_in.reserve(need + 64))) return rc;
    if ((rc = enc->host_scan.reserve(cap + 64))) return rc;
    uint8_t *d_pix = static_cast<uint8_t *>(enc->host_in.ptr);
    uint8_t *d_scan = static_cast<uint8_t *>(enc->host_scan.ptr);
    JB_CUDA(cudaMemcpyAsync(d_pix, bmp + off, need, cudaMemcpyHostToDevice, st));
    const uint8_t *top = bottom_up ? d_pix + (uint64_t)(h - 1) * pitch : d_pix;
    if ((rc = prepare(enc, top, w, h, 1, need, 0, bottom_up ? -(int64_t)pitch : (int64_t)pitch, /*bgr=*/true))) return rc;
    if ((rc = encode_launch(enc, d_scan, cap, misc_offsets(enc), st))) return rc;
    if (jpeg_capacity < 330) { g_last_error = "JPEG output buffer too small"; return JPEGB200_ERR_OUTPUT; }
    uint64_t scan_bytes = 0;
    rc = fetch_scan_to_host(enc, d_scan, cap, jpeg_out + 328, jpeg_capacity - 330, &scan_bytes, st, cuda_stream);
    *jpeg_bytes = 328 + scan_bytes + 2;
    if (rc == JPEGB200_ERR_OUTPUT) g_last_error = "JPEG output buffer too small";
    if (rc) return rc;
    jpegb200_jfif_header(w, h, jpeg_out);
    jpeg_out[328 + scan_bytes] = 0xFF;                               // EOI, jpeg_handler.c:113-117
    jpeg_out[328 + scan_bytes + 1] = 0xD9;
    return JPEGB200_OK;
}

namespace jb {
static std::mutex g_default_mutex;
static jpegb200_encoder *g_default_encoder = nullptr;

jpegb200_encoder *default_encoder()
{
    if (!g_default_encoder) {
        int dev = 0;
        if (const char *e = getenv("JPEGB200_DEVICE")) dev = atoi(e);
        g_default_encoder = jpegb200_encoder_create(dev);
        if (!g_default_encoder) fprintf(stderr, "jpegb200: no usable B200 device: %s\n", g_last_error.c_str());
    }
    return g_default_encoder;
}
}  // namespace jb

extern "C" JpegEncoderBuffer *jpegb200_encode_scan_dbg(const BMPImage *image, int16_t first_block[64])
{
    if (!image || !image->data || image->width <= 0 || image->height <= 0) return nullptr;
    std::lock_guard<std::mutex> lock(g_default_mutex);
    jpegb200_encoder *enc = default_encoder();
    if (!enc) return nullptr;
    const uint64_t nb = (uint64_t)((image->width + 7) / 8) * (uint64_t)((image->height + 7) / 8);
    for (int attempt = 0; attempt < 2; ++attempt) {
        const uint64_t cap = 2 * (nb * (uint64_t)enc->bytes_per_block + 64);
        uint8_t *host = (uint8_t *)malloc(cap);
        JpegEncoderBuffer *result = (JpegEncoderBuffer *)malloc(sizeof(JpegEncoderBuffer));
        uint64_t n = 0;
        int rc = JPEGB200_ERR_INTERNAL;
        if (host && result) rc = jpegb200_encode_host(enc, image->data, image->width, image->height, host, cap, &n, nullptr);
        if (rc == JPEGB200_OK) {
            if (first_block) {
                int16_t zz[64];
                if (read_block_coefficients(enc, 0, 1, zz) == JPEGB200_OK)
                    for (int k = 0; k < 64; ++k) first_block[kZigzag[k]] = zz[k];
            }
            result->data = host;
            result->size = n;
            result->capacity = cap;
            return result;
        }
        free(host);
        free(result);
        if ((rc == JPEGB200_ERR_WORKSPACE || rc == JPEGB200_ERR_OUTPUT) && attempt == 0) {
            enc->bytes_per_block = 184;              // worst case: 1463 bits per block
            continue;
        }
        fprintf(stderr, "jpegb200: encode failed: %s\n", g_last_error.c_str());
        break;
    }
    return nullptr;
}

extern "C" JpegEncoderBuffer *jpegb200_encode_scan(const BMPImage *image)
{
    return jpegb200_encode_scan_dbg(image, nullptr);
}

#include "stage_api.inl"
