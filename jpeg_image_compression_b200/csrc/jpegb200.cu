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
#include "entropy.cuh"
#include "fused_block.cuh"
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
    uint8_t dc_len[16];
    uint32_t dc_code[16], ac_code[256];
    std::vector<uint8_t> aclut;      // [63][256] bit cost of (run, int8 value)
};

static int bit_length(int v) { int a = v < 0 ? -v : v, n = 0; while (a) { ++n; a >>= 1; } return n; }

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
    t.aclut.assign(ACLUT_BYTES, 0);
    for (int run = 0; run < ACLUT_ROWS; ++run) {
        for (int b = 1; b < 256; ++b) {
            const int v = (int)(int8_t)b;
            const int sz = bit_length(v);
            if (sz > 10) continue;
            const int cost = (run >> 4) * ac[0xF0].len + ac[((run & 15) << 4) | sz].len + sz;
            t.aclut[run * 256 + b] = (uint8_t)cost;
        }
    }
}

// ---- encoder handle ---------------------------------------------------------------

struct DeviceBuffer {
    void *ptr = nullptr;
    size_t bytes = 0;
    int reserve(size_t need)
    {
        if (need <= bytes) return JPEGB200_OK;
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
        const size_t want = need + need / 8 + 256;
        if (!cuda_ok(cudaMalloc(&ptr, want), "cudaMalloc(workspace)")) return JPEGB200_ERR_CUDA;
        bytes = want;
        return JPEGB200_OK;
    }
    void release()
    {
        if (ptr) cudaFree(ptr);
        ptr = nullptr;
        bytes = 0;
    }
};

}  // namespace jb

struct jpegb200_encoder {
    int device = 0;
    int sm_count = 148;
    int dct_mode = 0;
    int bytes_per_block = 24;
    jb::DeviceBuffer coef, blockinfo, blockoff, tilebase, lookback, image_bits, image_base, packed, image_ff, aclut,
        misc, host_in, host_scan;
    // last launch
    jb::EntropyArgs args{};
    jb::Geom geom{};
    uint64_t lookback_words = 0;
    uint64_t total_blocks = 0;
    uint64_t launches = 0;
    bool stripe_ready = false;
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
    HostTables t;
    build_tables(t);
    {
        std::lock_guard<std::mutex> lock(g_tables_mutex);
        if (!g_tables_uploaded[enc->device & 63]) {
            JB_CUDA(cudaMemcpyToSymbol(c_ref_cos, kRefCos, sizeof(kRefCos)));
            JB_CUDA(cudaMemcpyToSymbol(c_ref_scale, t.ref_scale, sizeof(t.ref_scale)));
            JB_CUDA(cudaMemcpyToSymbol(c_quant_f, t.quant_f, sizeof(t.quant_f)));
            JB_CUDA(cudaMemcpyToSymbol(c_rk, t.rk, sizeof(t.rk)));
            JB_CUDA(cudaMemcpyToSymbol(c_zigzag, kZigzag, sizeof(kZigzag)));
            JB_CUDA(cudaMemcpyToSymbol(c_dc_len, t.dc_len, sizeof(t.dc_len)));
            JB_CUDA(cudaMemcpyToSymbol(c_dc_code, t.dc_code, sizeof(t.dc_code)));
            JB_CUDA(cudaMemcpyToSymbol(c_ac_code, t.ac_code, sizeof(t.ac_code)));
            JB_CUDA(cudaFuncSetAttribute(k_fused_blocks, cudaFuncAttributeMaxDynamicSharedMemorySize, K1_SMEM));
            JB_CUDA(cudaFuncSetAttribute(k_pack, cudaFuncAttributeMaxDynamicSharedMemorySize, K3_SMEM_WORDS * 4));
            g_tables_uploaded[enc->device & 63] = true;
        }
    }
    int rc = enc->aclut.reserve(ACLUT_BYTES);
    if (rc) return rc;
    JB_CUDA(cudaMemcpy(enc->aclut.ptr, t.aclut.data(), ACLUT_BYTES, cudaMemcpyHostToDevice));
    rc = enc->misc.reserve(256);
    if (rc) return rc;
    JB_CUDA(cudaMemset(enc->misc.ptr, 0, 256));
    return JPEGB200_OK;
}

// misc layout: [0] uint32 err, [8] uint64 flagged counter, [16..] stripe scratch
static uint32_t *misc_err(jpegb200_encoder *e) { return reinterpret_cast<uint32_t *>(e->misc.ptr); }
static unsigned long long *misc_flagged(jpegb200_encoder *e)
{
    return reinterpret_cast<unsigned long long *>(static_cast<uint8_t *>(e->misc.ptr) + 8);
}
static uint64_t *misc_offsets(jpegb200_encoder *e)
{
    return reinterpret_cast<uint64_t *>(static_cast<uint8_t *>(e->misc.ptr) + 16);
}

// Fill geometry + workspace for `count` images of w x h.
static int prepare(jpegb200_encoder *enc, const uint8_t *d_rgb, int w, int h, int count, uint64_t stride)
{
    if (!enc || !d_rgb || w <= 0 || h <= 0 || count <= 0) {
        g_last_error = "bad argument";
        return JPEGB200_ERR_ARG;
    }
    if (stride == 0) stride = (uint64_t)w * (uint64_t)h * 3u;
    JB_CUDA(cudaSetDevice(enc->device));
    Geom &g = enc->geom;
    g.rgb = d_rgb;
    g.image_stride = stride;
    g.w = w;
    g.h = h;
    g.bw = (w + 7) / 8;
    g.bh = (h + 7) / 8;
    g.spr = (g.bw + 31) / 32;
    g.count = count;
    g.blocks_per_image = (uint64_t)g.bw * (uint64_t)g.bh;
    g.total_strips = (uint64_t)g.spr * (uint64_t)g.bh * (uint64_t)count;
    const uint64_t nb = g.blocks_per_image, tb = nb * (uint64_t)count;
    enc->total_blocks = tb;
    const int tiles = (int)((nb + K2_TILE - 1) / K2_TILE);
    const uint64_t packed_per_image = ((nb * (uint64_t)enc->bytes_per_block + 15) & ~15ull) + 32;
    const int chunks_cap = (int)((packed_per_image + K4_CHUNK - 1) / K4_CHUNK);
    int rc = 0;
    if ((rc = enc->coef.reserve(tb * 64))) return rc;
    if ((rc = enc->blockinfo.reserve(tb * 4))) return rc;
    if ((rc = enc->blockoff.reserve(tb * 4))) return rc;
    if ((rc = enc->tilebase.reserve((uint64_t)tiles * count * 8))) return rc;
    // look-back state of K2 (one word per scan tile) and K4 (one word per 4 KiB chunk), contiguous so
    // that K1's prologue clears both
    enc->lookback_words = (uint64_t)tiles * count + (uint64_t)chunks_cap * count;
    if ((rc = enc->lookback.reserve(enc->lookback_words * 8))) return rc;
    if ((rc = enc->image_bits.reserve((uint64_t)count * 8))) return rc;
    if ((rc = enc->image_base.reserve((uint64_t)count * 8))) return rc;
    if ((rc = enc->image_ff.reserve((uint64_t)count * 8))) return rc;
    if ((rc = enc->packed.reserve(packed_per_image * (uint64_t)count + 64))) return rc;

    EntropyArgs &a = enc->args;
    a.coef = static_cast<const int8_t *>(enc->coef.ptr);
    a.blockinfo = static_cast<const uint32_t *>(enc->blockinfo.ptr);
    a.blockoff = static_cast<uint32_t *>(enc->blockoff.ptr);
    a.tilebase = static_cast<uint64_t *>(enc->tilebase.ptr);
    a.scan_state = static_cast<uint64_t *>(enc->lookback.ptr);
    a.image_bits = static_cast<uint64_t *>(enc->image_bits.ptr);
    a.image_base = static_cast<uint64_t *>(enc->image_base.ptr);
    a.packed = static_cast<uint32_t *>(enc->packed.ptr);
    a.packed_capacity = packed_per_image * (uint64_t)count;
    a.stuff_state = static_cast<uint64_t *>(enc->lookback.ptr) + (uint64_t)tiles * count;
    a.image_ff = static_cast<uint64_t *>(enc->image_ff.ptr);
    a.scan = nullptr;
    a.scan_capacity = 0;
    a.scan_offsets = nullptr;
    a.err = misc_err(enc);
    a.nb = nb;
    a.tiles = tiles;
    a.chunks_cap = chunks_cap;
    a.count = count;
    a.epoch = 1;
    a.dc_pred0 = 0;
    a.bit_phase = 0;
    return JPEGB200_OK;
}

// ---- kernel launches (optionally bracketed by cudaEvents for per-kernel timing) -----------

enum KernelId { KID_BLOCK = 0, KID_SCAN = 1, KID_PACK = 2, KID_STUFF = 3, KID_LAYOUT = 4, KID_ZERO = 5 };

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

static int launch_block_kernel(jpegb200_encoder *enc, cudaStream_t st)
{
    const Geom &g = enc->geom;
    const uint64_t want = (g.total_strips + K1_WARPS - 1) / K1_WARPS;
    const int grid = (int)std::min<uint64_t>(want, (uint64_t)enc->sm_count * 2);
    {
        TimedLaunch t(enc, st, KID_BLOCK);
        k_fused_blocks<<<grid, K1_THREADS, K1_SMEM, st>>>(g, static_cast<int8_t *>(enc->coef.ptr),
                                                           static_cast<uint32_t *>(enc->blockinfo.ptr),
                                                           static_cast<const uint8_t *>(enc->aclut.ptr), misc_flagged(enc),
                                                           enc->dct_mode, static_cast<uint64_t *>(enc->lookback.ptr),
                                                           enc->lookback_words);
    }
    JB_CUDA(cudaGetLastError());
    return JPEGB200_OK;
}

static int launch_bit_scan(jpegb200_encoder *enc, cudaStream_t st)
{
    const EntropyArgs &a = enc->args;
    {
        TimedLaunch t(enc, st, KID_SCAN);
        k_bit_scan<<<dim3(a.tiles, a.count), K2_THREADS, 0, st>>>(a);
    }
    JB_CUDA(cudaGetLastError());
    return JPEGB200_OK;
}

static int launch_pack(jpegb200_encoder *enc, cudaStream_t st)
{
    const EntropyArgs &a = enc->args;
    const unsigned ptiles = (unsigned)((a.nb + K3_THREADS - 1) / K3_THREADS);
    {
        TimedLaunch t(enc, st, KID_PACK);
        k_pack<<<dim3(ptiles, a.count), K3_THREADS, K3_SMEM_WORDS * 4, st>>>(a);
    }
    JB_CUDA(cudaGetLastError());
    return JPEGB200_OK;
}

static int launch_stuff(jpegb200_encoder *enc, const StuffArgs &sa, cudaStream_t st)
{
    const EntropyArgs &a = enc->args;
    {
        TimedLaunch t(enc, st, KID_STUFF);
        k_stuff<<<dim3(a.chunks_cap, a.count), K4_THREADS, 0, st>>>(a, sa);
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

static int encode_launch(jpegb200_encoder *enc, uint8_t *d_scan, uint64_t scan_capacity, uint64_t *d_scan_offsets,
                         cudaStream_t st)
{
    int rc = 0;
    enc->launches = 0;
    enc->stripe_ready = false;
    EntropyArgs &a = enc->args;
    a.scan = d_scan;
    a.scan_capacity = scan_capacity;
    a.scan_offsets = d_scan_offsets;
    if ((rc = launch_block_kernel(enc, st))) return rc;
    if ((rc = launch_bit_scan(enc, st))) return rc;
    if (a.count == 1) {
        if ((rc = launch_pack(enc, st))) return rc;
        if ((rc = launch_stuff(enc, StuffArgs{0, 0, 0}, st))) return rc;
    } else {
        {
            TimedLaunch t(enc, st, KID_LAYOUT);
            k_image_layout<<<1, 1024, 0, st>>>(a, 0);
        }
        const uint64_t ptiles = (a.nb + K3_THREADS - 1) / K3_THREADS * (uint64_t)a.count;
        {
            TimedLaunch t(enc, st, KID_ZERO);
            k_zero_shared_words<<<(unsigned)((ptiles + 255) / 256), 256, 0, st>>>(a);
        }
        JB_CUDA(cudaGetLastError());
        if ((rc = launch_pack(enc, st))) return rc;
        if ((rc = launch_stuff(enc, StuffArgs{1, 0, 0}, st))) return rc;
        {
            TimedLaunch t(enc, st, KID_LAYOUT);
            k_image_layout<<<1, 1024, 0, st>>>(a, 1);
        }
        JB_CUDA(cudaGetLastError());
        if ((rc = launch_stuff(enc, StuffArgs{2, 0, 0}, st))) return rc;
    }
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
    for (DeviceBuffer *b : {&enc->coef, &enc->blockinfo, &enc->blockoff, &enc->tilebase, &enc->lookback, &enc->image_bits,
                            &enc->image_base, &enc->packed, &enc->image_ff, &enc->aclut, &enc->misc, &enc->host_in,
                            &enc->host_scan})
        b->release();
    delete enc;
}

extern "C" int jpegb200_encoder_set_dct_mode(jpegb200_encoder *enc, int dct_mode)
{
    if (!enc || (dct_mode != 0 && dct_mode != 1)) return JPEGB200_ERR_ARG;
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
    int rc = prepare(enc, batch->d_rgb, batch->width, batch->height, batch->count, batch->image_stride);
    if (rc) return rc;
    return encode_launch(enc, d_scan, scan_capacity, d_scan_offsets, st);
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
    if (enc->args.count > 0 && enc->image_bits.ptr) {
        std::vector<uint64_t> bits((size_t)enc->args.count);
        JB_CUDA(cudaMemcpy(bits.data(), enc->image_bits.ptr, bits.size() * 8, cudaMemcpyDeviceToHost));
        for (uint64_t b : bits) out->packed_bytes += (b + 7) / 8;
    }
    return JPEGB200_OK;
}

extern "C" int jpegb200_encoder_read_coefficients(jpegb200_encoder *enc, int16_t *host_zz, uint64_t nblocks)
{
    if (!enc || !host_zz || nblocks > enc->total_blocks) return JPEGB200_ERR_ARG;
    JB_CUDA(cudaSetDevice(enc->device));
    JB_CUDA(cudaDeviceSynchronize());
    std::vector<int8_t> tmp((size_t)nblocks * 64);
    JB_CUDA(cudaMemcpy(tmp.data(), enc->coef.ptr, tmp.size(), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < tmp.size(); ++i) host_zz[i] = tmp[i];
    return JPEGB200_OK;
}

extern "C" int jpegb200_encoder_read_block_bits(jpegb200_encoder *enc, uint32_t *host_bits, uint64_t nblocks)
{
    if (!enc || !host_bits || nblocks > enc->total_blocks) return JPEGB200_ERR_ARG;
    JB_CUDA(cudaSetDevice(enc->device));
    JB_CUDA(cudaDeviceSynchronize());
    // bit cost of block b = offset(b+1) - offset(b) within its image
    const EntropyArgs &a = enc->args;
    std::vector<uint32_t> off((size_t)enc->total_blocks);
    std::vector<uint64_t> tb((size_t)a.tiles * a.count), ib((size_t)a.count);
    JB_CUDA(cudaMemcpy(off.data(), a.blockoff, off.size() * 4, cudaMemcpyDeviceToHost));
    JB_CUDA(cudaMemcpy(tb.data(), a.tilebase, tb.size() * 8, cudaMemcpyDeviceToHost));
    JB_CUDA(cudaMemcpy(ib.data(), a.image_bits, ib.size() * 8, cudaMemcpyDeviceToHost));
    for (uint64_t b = 0; b < nblocks; ++b) {
        const uint64_t img = b / a.nb, lb = b % a.nb;
        const uint64_t cur = tb[img * a.tiles + lb / K2_TILE] + off[b];
        const uint64_t nxt = lb + 1 < a.nb ? tb[img * a.tiles + (lb + 1) / K2_TILE] + off[b + 1] : ib[img];
        host_bits[b] = (uint32_t)(nxt - cur);
    }
    return JPEGB200_OK;
}

// ---- C ABI: MCU-row stripes -------------------------------------------------------------

extern "C" int jpegb200_stripe_analyze(jpegb200_encoder *enc, const uint8_t *d_rgb, int width, int stripe_height,
                                       jpegb200_stripe_summary *host_out, void *cuda_stream)
{
    if (!host_out) return JPEGB200_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    int rc = prepare(enc, d_rgb, width, stripe_height, 1, 0);
    if (rc) return rc;
    enc->launches = 0;
    if ((rc = launch_block_kernel(enc, st))) return rc;
    // pass 1 of the scan with predictor 0 only to learn the stripe's total; offsets are
    // recomputed in jpegb200_stripe_pack once the true predictor and bit phase are known.
    enc->args.packed_capacity = 0;                          // no word clearing in this pass
    if ((rc = launch_bit_scan(enc, st))) return rc;
    uint32_t first = 0, last = 0;
    uint64_t bits = 0;
    JB_CUDA(cudaMemcpyAsync(&first, enc->blockinfo.ptr, 4, cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaMemcpyAsync(&last, static_cast<uint32_t *>(enc->blockinfo.ptr) + (enc->total_blocks - 1), 4,
                            cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaMemcpyAsync(&bits, enc->image_bits.ptr, 8, cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaStreamSynchronize(st));
    host_out->first_dc = (int16_t)(first & 0xFFFFu);
    host_out->last_dc = (int16_t)(last & 0xFFFFu);
    host_out->reserved = 0;
    host_out->bits_pred0 = bits;
    enc->stripe_ready = true;
    return JPEGB200_OK;
}

extern "C" int jpegb200_stripe_pack(jpegb200_encoder *enc, int16_t dc_predictor, uint64_t bit_begin,
                                    jpegb200_stripe_packed *host_out, void *cuda_stream)
{
    if (!enc || !host_out || !enc->stripe_ready) {
        g_last_error = "jpegb200_stripe_pack without a preceding jpegb200_stripe_analyze";
        return JPEGB200_ERR_ARG;
    }
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    JB_CUDA(cudaSetDevice(enc->device));
    EntropyArgs &a = enc->args;
    a.dc_pred0 = dc_predictor;
    a.bit_phase = (uint32_t)(bit_begin & 7u);
    a.packed_capacity = ((a.nb * (uint64_t)enc->bytes_per_block + 15) & ~15ull) + 32;
    JB_CUDA(cudaMemsetAsync(enc->lookback.ptr, 0, enc->lookback_words * 8, st));   // second scan over the same tiles
    int rc = 0;
    if ((rc = launch_bit_scan(enc, st))) return rc;
    if ((rc = launch_pack(enc, st))) return rc;
    uint64_t bits = 0;
    JB_CUDA(cudaMemcpyAsync(&bits, enc->image_bits.ptr, 8, cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaStreamSynchronize(st));
    if ((rc = jpegb200_encoder_status(enc, cuda_stream))) return rc;
    const uint64_t nbytes = (bits + a.bit_phase + 7) >> 3;
    uint8_t head = 0, tail = 0;
    JB_CUDA(cudaMemcpyAsync(&head, enc->packed.ptr, 1, cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaMemcpyAsync(&tail, static_cast<uint8_t *>(enc->packed.ptr) + (nbytes - 1), 1, cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaStreamSynchronize(st));
    host_out->bit_begin = bit_begin;
    host_out->bit_end = bit_begin + bits;
    host_out->head_byte = head;
    host_out->tail_byte = tail;
    return JPEGB200_OK;
}

extern "C" int jpegb200_stripe_finish(jpegb200_encoder *enc, uint32_t or_into_last_byte, int owns_first_byte,
                                      int is_last_stripe, uint8_t *d_scan, uint64_t scan_capacity,
                                      uint64_t *host_scan_bytes, void *cuda_stream)
{
    (void)is_last_stripe;           // the zero padding of the final byte is already in the packed words
    if (!enc || !d_scan || !host_scan_bytes || !enc->stripe_ready) return JPEGB200_ERR_ARG;
    cudaStream_t st = static_cast<cudaStream_t>(cuda_stream);
    JB_CUDA(cudaSetDevice(enc->device));
    EntropyArgs &a = enc->args;
    a.scan = d_scan;
    a.scan_capacity = scan_capacity;
    a.scan_offsets = misc_offsets(enc);
    JB_CUDA(cudaMemsetAsync(misc_offsets(enc), 0, 16, st));
    int rc = 0;
    if ((rc = launch_stuff(enc, StuffArgs{0, owns_first_byte ? 0u : 1u, or_into_last_byte & 0xFFu}, st))) return rc;
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
    if ((rc = prepare(enc, d_rgb, width, height, 1, 0))) return rc;
    if ((rc = encode_launch(enc, d_scan, cap, d_off, st))) return rc;
    uint64_t offs[2] = {0, 0};
    uint32_t err = 0;
    JB_CUDA(cudaMemcpyAsync(offs, d_off, 16, cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaMemcpyAsync(&err, misc_err(enc), 4, cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaStreamSynchronize(st));
    if (err) return jpegb200_encoder_status(enc, cuda_stream);
    *host_scan_bytes = offs[1];
    if (offs[1] > host_capacity) {
        g_last_error = "host scan buffer too small";
        return JPEGB200_ERR_OUTPUT;
    }
    JB_CUDA(cudaMemcpyAsync(host_scan, d_scan, offs[1], cudaMemcpyDeviceToHost, st));
    JB_CUDA(cudaStreamSynchronize(st));
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
                int8_t zz[64];
                if (cuda_ok(cudaMemcpy(zz, enc->coef.ptr, 64, cudaMemcpyDeviceToHost), "cudaMemcpy(D2H block0)"))
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
