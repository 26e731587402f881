// stage_api.inl -- the reference's core-stage functions (same names, struct layouts,
// ownership and NULL-on-error behaviour), each executed by CUDA
// kernels from stages.cuh.  Included at the end of jpegb200.cu.
//
// Host pointers in, freshly malloc'd host structs out (callers index them and free them
// with the matching free*).  No CPU fallback: any CUDA failure yields NULL.

namespace jb {

struct Scratch {                 // RAII device allocations for one stage call
    std::vector<void *> ptrs;
    ~Scratch() { for (void *p : ptrs) cudaFree(p); }
    template <typename T> T *alloc(size_t n)
    {
        void *p = nullptr;
        if (!cuda_ok(cudaMalloc(&p, (n ? n : 1) * sizeof(T) + 64), "cudaMalloc(stage)")) return nullptr;
        ptrs.push_back(p);
        return static_cast<T *>(p);
    }
};

static bool stage_begin()
{
    std::lock_guard<std::mutex> lock(g_default_mutex);
    jpegb200_encoder *enc = default_encoder();
    return enc && cuda_ok(cudaSetDevice(enc->device), "cudaSetDevice");
}

static unsigned grid_for(uint64_t n, unsigned threads = 256)
{
    const uint64_t g = (n + threads - 1) / threads;
    return (unsigned)std::min<uint64_t>(std::max<uint64_t>(g, 1), 148u * 32u);
}

static bool finish(const char *what) { return cuda_ok(cudaGetLastError(), what) && cuda_ok(cudaDeviceSynchronize(), what); }

}  // namespace jb

extern "C" YImage *convertBMPToJPEGGrayscale(const BMPImage *image)
{
    if (!image || !image->data) return nullptr;                           // converter.c:6
    if (image->width <= 0 || image->height <= 0 || !stage_begin()) return nullptr;
    const int w = image->width, h = image->height, wp = (w + 7) & ~7, hp = (h + 7) & ~7;   // converter.c:15-16
    const size_t nin = (size_t)w * h * 3, nout = (size_t)wp * hp;
    Scratch s;
    uint8_t *d_in = s.alloc<uint8_t>(nin), *d_out = s.alloc<uint8_t>(nout);
    if (!d_in || !d_out || !cuda_ok(cudaMemcpy(d_in, image->data, nin, cudaMemcpyHostToDevice), "H2D")) return nullptr;
    k_stage_luma<<<grid_for(nout), 256>>>(d_in, w, h, wp, hp, d_out);
    if (!finish("k_stage_luma")) return nullptr;
    YImage *out = (YImage *)malloc(sizeof(YImage));
    if (!out) return nullptr;
    out->width = wp;
    out->height = hp;
    out->data = (uint8_t *)malloc(nout);
    if (!out->data || !cuda_ok(cudaMemcpy(out->data, d_out, nout, cudaMemcpyDeviceToHost), "D2H")) {
        free(out->data);
        free(out);
        return nullptr;
    }
    return out;
}

extern "C" CenteredYImage *centerYImage(const YImage *source)
{
    if (!source || !source->data) return nullptr;                         // converter.c:62
    if (source->width <= 0 || source->height <= 0 || !stage_begin()) return nullptr;
    const size_t n = (size_t)source->width * source->height;
    Scratch s;
    uint8_t *d_in = s.alloc<uint8_t>(n);
    int8_t *d_out = s.alloc<int8_t>(n);
    if (!d_in || !d_out || !cuda_ok(cudaMemcpy(d_in, source->data, n, cudaMemcpyHostToDevice), "H2D")) return nullptr;
    k_stage_center<<<grid_for(n), 256>>>(d_in, n, d_out);
    if (!finish("k_stage_center")) return nullptr;
    CenteredYImage *out = (CenteredYImage *)malloc(sizeof(CenteredYImage));
    if (!out) return nullptr;
    out->width = source->width;
    out->height = source->height;
    out->data = (int8_t *)malloc(n);
    if (!out->data || !cuda_ok(cudaMemcpy(out->data, d_out, n, cudaMemcpyDeviceToHost), "D2H")) {
        free(out->data);
        free(out);
        return nullptr;
    }
    return out;
}

extern "C" void freeCenteredYImage(CenteredYImage *img)
{
    if (!img) return;
    free(img->data);
    free(img);
}

static bool dct_on_device(const int8_t *host_in, int wp, int hp, float *host_out)
{
    const size_t n = (size_t)wp * hp;
    Scratch s;
    int8_t *d_in = s.alloc<int8_t>(n);
    float *d_out = s.alloc<float>(n);
    if (!d_in || !d_out || !cuda_ok(cudaMemcpy(d_in, host_in, n, cudaMemcpyHostToDevice), "H2D")) return false;
    // performDCT leaves coefficients of partial trailing blocks untouched (dct.c:119-121 loops
    // while a full block fits); inputs here are always multiples of 8.
    if (!cuda_ok(cudaMemset(d_out, 0, n * sizeof(float)), "memset")) return false;
    k_stage_dct<<<grid_for(n / 64, 4), 256>>>(d_in, wp, hp, d_out);
    return finish("k_stage_dct") && cuda_ok(cudaMemcpy(host_out, d_out, n * sizeof(float), cudaMemcpyDeviceToHost), "D2H");
}

extern "C" void computeDCTBlock(const int8_t inputBlock[8][8], float outputBlock[8][8])
{
    if (!inputBlock || !outputBlock || !stage_begin()) return;
    dct_on_device(&inputBlock[0][0], 8, 8, &outputBlock[0][0]);
}

extern "C" DCTImage *performDCT(const CenteredYImage *image)
{
    if (!image || !image->data) return nullptr;                           // dct.c:101
    if (image->width < 8 || image->height < 8 || (image->width & 7) || (image->height & 7) || !stage_begin()) return nullptr;
    const size_t n = (size_t)image->width * image->height;
    DCTImage *out = (DCTImage *)malloc(sizeof(DCTImage));
    if (!out) return nullptr;
    out->width = image->width;
    out->height = image->height;
    out->coefficients = (float *)malloc(n * sizeof(float));
    if (!out->coefficients || !dct_on_device(image->data, image->width, image->height, out->coefficients)) {
        free(out->coefficients);
        free(out);
        return nullptr;
    }
    return out;
}

extern "C" void freeDCTImage(DCTImage *img)
{
    if (!img) return;
    free(img->coefficients);
    free(img);
}

extern "C" QuantizedImage *quantizeImage(const DCTImage *dctImg)
{
    if (!dctImg || !dctImg->coefficients) return nullptr;                 // quantization.c:4
    if (dctImg->width <= 0 || dctImg->height <= 0 || !stage_begin()) return nullptr;
    const size_t n = (size_t)dctImg->width * dctImg->height;
    Scratch s;
    float *d_in = s.alloc<float>(n);
    int16_t *d_out = s.alloc<int16_t>(n);
    if (!d_in || !d_out || !cuda_ok(cudaMemcpy(d_in, dctImg->coefficients, n * sizeof(float), cudaMemcpyHostToDevice), "H2D"))
        return nullptr;
    k_stage_quant<<<grid_for(n), 256>>>(d_in, dctImg->width, n, d_out);
    if (!finish("k_stage_quant")) return nullptr;
    QuantizedImage *out = (QuantizedImage *)malloc(sizeof(QuantizedImage));
    if (!out) return nullptr;
    out->width = dctImg->width;
    out->height = dctImg->height;
    out->data = (int16_t *)malloc(n * sizeof(int16_t));
    if (!out->data || !cuda_ok(cudaMemcpy(out->data, d_out, n * sizeof(int16_t), cudaMemcpyDeviceToHost), "D2H")) {
        free(out->data);
        free(out);
        return nullptr;
    }
    return out;
}

extern "C" void freeQuantizedImage(QuantizedImage *img)
{
    if (!img) return;
    free(img->data);
    free(img);
}

extern "C" ZigZagData *performZigZag(const QuantizedImage *qImg)
{
    if (!qImg || !qImg->data) return nullptr;                             // zigzag.c:23
    if (qImg->width < 8 || qImg->height < 8 || !stage_begin()) return nullptr;
    const int wp = qImg->width, hp = qImg->height;
    const int bw = wp / 8, bh = hp / 8;                                   // zigzag.c:30-32
    const size_t nin = (size_t)wp * hp, nout = (size_t)bw * bh * 64;
    Scratch s;
    int16_t *d_in = s.alloc<int16_t>(nin), *d_out = s.alloc<int16_t>(nout);
    if (!d_in || !d_out || !cuda_ok(cudaMemcpy(d_in, qImg->data, nin * 2, cudaMemcpyHostToDevice), "H2D")) return nullptr;
    k_stage_zigzag<<<grid_for(nout), 256>>>(d_in, wp, bh * 8, d_out);
    if (!finish("k_stage_zigzag")) return nullptr;
    ZigZagData *out = (ZigZagData *)malloc(sizeof(ZigZagData));
    if (!out) return nullptr;
    out->numBlocksW = bw;
    out->numBlocksH = bh;
    out->totalBlocks = bw * bh;
    out->data = (int16_t *)malloc(nout * sizeof(int16_t));
    if (!out->data || !cuda_ok(cudaMemcpy(out->data, d_out, nout * 2, cudaMemcpyDeviceToHost), "D2H")) {
        free(out->data);
        free(out);
        return nullptr;
    }
    return out;
}

extern "C" void freeZigZagData(ZigZagData *zData)
{
    if (!zData) return;
    free(zData->data);
    free(zData);
}

extern "C" RLEData *performRLE(const ZigZagData *zz)
{
    if (!zz || !zz->data) return nullptr;                                 // rle.c:52
    if (zz->totalBlocks <= 0 || !stage_begin()) return nullptr;
    const uint64_t nb = (uint64_t)zz->totalBlocks;
    const unsigned tiles = (unsigned)((nb + 1023) / 1024);
    Scratch s;
    int16_t *d_zz = s.alloc<int16_t>(nb * 64);
    uint32_t *d_counts = s.alloc<uint32_t>(nb);
    uint64_t *d_off = s.alloc<uint64_t>(nb), *d_state = s.alloc<uint64_t>(tiles), *d_total = s.alloc<uint64_t>(1);
    uint32_t *d_err = s.alloc<uint32_t>(1);
    if (!d_zz || !d_counts || !d_off || !d_state || !d_total || !d_err) return nullptr;
    if (!cuda_ok(cudaMemcpy(d_zz, zz->data, nb * 128, cudaMemcpyHostToDevice), "H2D") ||
        !cuda_ok(cudaMemset(d_state, 0, tiles * 8), "memset") || !cuda_ok(cudaMemset(d_err, 0, 4), "memset"))
        return nullptr;
    k_rle_count<<<(unsigned)((nb + 255) / 256), 256>>>(d_zz, nb, d_counts);
    k_scan_u32<<<tiles, 256>>>(d_counts, d_off, nb, d_state, 1u, d_total, d_err);
    uint64_t total = 0;
    if (!finish("rle count/scan") || !cuda_ok(cudaMemcpy(&total, d_total, 8, cudaMemcpyDeviceToHost), "D2H")) return nullptr;
    uint8_t *d_sym = s.alloc<uint8_t>(total * 6);
    if (!d_sym) return nullptr;
    k_rle_emit<<<(unsigned)((nb + 255) / 256), 256>>>(d_zz, nb, d_off, d_sym);
    if (!finish("k_rle_emit")) return nullptr;
    RLEData *out = (RLEData *)malloc(sizeof(RLEData));
    if (!out) return nullptr;
    out->count = total;
    out->capacity = total > 4096 ? total : 4096;                          // rle.c:56 initial guess
    out->data = (RLESymbol *)malloc(out->capacity * sizeof(RLESymbol));
    if (!out->data || !cuda_ok(cudaMemcpy(out->data, d_sym, total * 6, cudaMemcpyDeviceToHost), "D2H")) {
        free(out->data);
        free(out);
        return nullptr;
    }
    return out;
}

extern "C" void freeRLEData(RLEData *rleData)
{
    if (!rleData) return;
    free(rleData->data);
    free(rleData);
}

extern "C" JpegEncoderBuffer *encodeHuffman(const RLEData *rle, int totalBlocks)
{
    if (!rle || !stage_begin()) return nullptr;
    static_assert(sizeof(RLESymbol) == 6, "RLESymbol layout");
    const uint64_t nsym = rle->data ? rle->count : 0;
    JpegEncoderBuffer *out = (JpegEncoderBuffer *)malloc(sizeof(JpegEncoderBuffer));
    if (!out) return nullptr;
    out->data = nullptr;
    out->size = 0;
    out->capacity = 0;
    if (nsym == 0 || totalBlocks <= 0) return out;                        // huffman.c:126-141: empty buffer
    const uint64_t nchunks = (nsym + HS_CHUNK - 1) / HS_CHUNK;
    const unsigned tiles = (unsigned)((nchunks + 1023) / 1024);
    Scratch s;
    uint8_t *d_sym = s.alloc<uint8_t>(nsym * 6), *d_maps = s.alloc<uint8_t>(nchunks * 64), *d_entry = s.alloc<uint8_t>(nchunks);
    uint32_t *d_cnt = s.alloc<uint32_t>(nchunks), *d_err = s.alloc<uint32_t>(1);
    uint64_t *d_cblk = s.alloc<uint64_t>(nchunks), *d_cbit = s.alloc<uint64_t>(nchunks);
    uint64_t *d_state = s.alloc<uint64_t>(2 * (size_t)tiles), *d_tot = s.alloc<uint64_t>(4);
    bool ok = d_sym && d_maps && d_entry && d_cnt && d_err && d_cblk && d_cbit && d_state && d_tot;
    ok = ok && cuda_ok(cudaMemcpy(d_sym, rle->data, nsym * 6, cudaMemcpyHostToDevice), "H2D") &&
         cuda_ok(cudaMemset(d_state, 0, 2 * (size_t)tiles * 8), "memset") && cuda_ok(cudaMemset(d_err, 0, 4), "memset") &&
         cuda_ok(cudaMemset(d_tot, 0, 32), "memset");
    uint64_t total_bits = 0;
    if (ok) {
        k_hs_maps<<<(unsigned)((nchunks + 3) / 4), 256>>>(d_sym, nsym, d_maps);
        k_hs_chain<<<1, 32>>>(d_maps, nchunks, d_entry);
        const unsigned wg = (unsigned)((nchunks + 127) / 128);
        k_hs_walk<<<wg, 128>>>(d_sym, nsym, d_entry, nullptr, nullptr, (uint64_t)totalBlocks, d_cnt, nullptr, 0);
        k_scan_u32<<<tiles, 256>>>(d_cnt, d_cblk, nchunks, d_state, 1u, d_tot, d_err);
        k_hs_walk<<<wg, 128>>>(d_sym, nsym, d_entry, d_cblk, nullptr, (uint64_t)totalBlocks, d_cnt, nullptr, 1);
        k_scan_u32<<<tiles, 256>>>(d_cnt, d_cbit, nchunks, d_state + tiles, 1u, d_tot + 1, d_err);
        ok = finish("huffman walk") && cuda_ok(cudaMemcpy(&total_bits, d_tot + 1, 8, cudaMemcpyDeviceToHost), "D2H");
    }
    uint8_t *host = nullptr;
    if (ok && total_bits > 0) {
        const uint64_t nbytes = (total_bits + 7) / 8;
        uint32_t *d_packed = s.alloc<uint32_t>(nbytes / 4 + 8);
        uint8_t *d_scan = s.alloc<uint8_t>(2 * nbytes + 16);
        uint64_t *d_misc = s.alloc<uint64_t>(8);            // [0] image_bits [1] image_ff [2..3] offsets
        const int chunks = (int)((nbytes + K4_CHUNK - 1) / K4_CHUNK);
        uint64_t *d_sstate = s.alloc<uint64_t>((size_t)chunks);
        ok = d_packed && d_scan && d_misc && d_sstate &&
             cuda_ok(cudaMemset(d_packed, 0, (nbytes / 4 + 8) * 4), "memset") &&
             cuda_ok(cudaMemset(d_sstate, 0, (size_t)chunks * 8), "memset") &&
             cuda_ok(cudaMemset(d_misc, 0, 64), "memset") &&
             cuda_ok(cudaMemcpy(d_misc, &total_bits, 8, cudaMemcpyHostToDevice), "H2D");
        if (ok) {
            const unsigned wg = (unsigned)((nchunks + 127) / 128);
            k_hs_walk<<<wg, 128>>>(d_sym, nsym, d_entry, d_cblk, d_cbit, (uint64_t)totalBlocks, nullptr, d_packed, 2);
            EntropyArgs a{};
            a.image_bits = d_misc;
            a.scan_offsets = d_misc + 2;
            a.packed = d_packed;
            a.packed_capacity = (nbytes / 4 + 8) * 4;
            a.stuff_state = d_sstate;
            a.scan = d_scan;
            a.scan_capacity = 2 * nbytes + 16;
            a.err = d_err;
            a.epoch = 1;
            k_stuff<<<(unsigned)chunks, K4_THREADS>>>(a);
            uint64_t offs[2] = {0, 0};
            uint32_t err = 0;
            ok = finish("huffman pack/stuff") && cuda_ok(cudaMemcpy(offs, d_misc + 2, 16, cudaMemcpyDeviceToHost), "D2H") &&
                 cuda_ok(cudaMemcpy(&err, d_err, 4, cudaMemcpyDeviceToHost), "D2H") && err == 0;
            if (ok) {
                host = (uint8_t *)malloc(offs[1] ? offs[1] : 1);
                ok = host && cuda_ok(cudaMemcpy(host, d_scan, offs[1], cudaMemcpyDeviceToHost), "D2H");
                if (ok) {
                    out->data = host;
                    out->size = offs[1];
                    out->capacity = offs[1] ? offs[1] : 1;
                }
            }
        }
    }
    if (!ok) {
        free(host);
        free(out);
        return nullptr;
    }
    return out;
}

extern "C" void freeJpegEncoderBuffer(JpegEncoderBuffer *buffer)
{
    if (!buffer) return;
    free(buffer->data);
    free(buffer);
}
