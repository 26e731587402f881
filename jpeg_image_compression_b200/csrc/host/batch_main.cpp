// jpeg_compression_batch <out_dir> <in1.bmp> [in2.bmp ...]
//
// Batch driver next to the drop-in 2-argument CLI: every input goes
// BMP file -> pinned host buffer -> jpegb200_encode_bmp_to_jpeg_host (pixel array read in place
// on the device, complete JPEG file back) -> <out_dir>/<name>.jpg.  Two worker threads per GPU, each with
// its own encoder handle and CUDA stream, so one image's PCIe copy overlaps the other's kernels
// and file I/O; files are dealt round-robin to the workers, i.e. the batch is sharded by image across
// all visible B200s (JPEGB200_DEVICE=<n> pins it to one, JPEGB200_BATCH_GPUS=<k> limits the count).
// Output files are byte-identical to jpeg_compression_app's.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "jpegb200.h"

struct Pinned {
    uint8_t *p = nullptr;
    size_t n = 0;
    bool reserve(size_t need)
    {
        if (need <= n) return true;
        if (p) cudaFreeHost(p);
        p = nullptr;
        n = 0;
        if (cudaHostAlloc((void **)&p, need + need / 4 + 4096, cudaHostAllocDefault) != cudaSuccess) return false;
        n = need + need / 4 + 4096;
        return true;
    }
    ~Pinned() { if (p) cudaFreeHost(p); }
};

static std::string out_name(const std::string &dir, const std::string &in)
{
    size_t a = in.find_last_of('/');
    std::string base = a == std::string::npos ? in : in.substr(a + 1);
    size_t dot = base.find_last_of('.');
    if (dot != std::string::npos) base = base.substr(0, dot);
    return dir + "/" + base + ".jpg";
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        fprintf(stderr, "Usage: %s <output_dir> <input1.bmp> [input2.bmp ...]\n", argv[0]);
        return 1;
    }
    const std::string dir = argv[1];
    std::vector<std::string> files(argv + 2, argv + argc);
    int first_device = 0, ngpus = jpegb200_device_count();
    if (const char *e = getenv("JPEGB200_DEVICE")) { first_device = atoi(e); ngpus = 1; }
    if (const char *e = getenv("JPEGB200_BATCH_GPUS")) ngpus = std::max(1, std::min(ngpus, atoi(e)));
    if (ngpus < 1) ngpus = 1;                                  // no device: the workers report the error
    const int nthreads = (int)std::max<size_t>(1, std::min<size_t>(files.size(), (size_t)2 * ngpus));
    std::vector<int> failed(nthreads, 0);
    std::vector<double> mpix(nthreads, 0.0);
    const auto t0 = std::chrono::steady_clock::now();
    auto worker = [&](int t) {
        const int device = first_device + t % ngpus;
        cudaSetDevice(device);
        jpegb200_encoder *enc = jpegb200_encoder_create(device);
        if (!enc) {
            fprintf(stderr, "Error: no usable B200 device: %s\n", jpegb200_last_error());
            failed[t] = (int)files.size();
            return;
        }
        cudaStream_t st;
        cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        Pinned in, out;
        for (size_t i = t; i < files.size(); i += nthreads) {
            FILE *f = fopen(files[i].c_str(), "rb");
            if (!f) {
                fprintf(stderr, "Error: Unable to open file: %s\n", files[i].c_str());
                ++failed[t];
                continue;
            }
            fseek(f, 0, SEEK_END);
            const long n = ftell(f);
            fseek(f, 0, SEEK_SET);
            bool ok = n > 0 && in.reserve((size_t)n) && out.reserve((size_t)n / 2 + 65536) && fread(in.p, 1, (size_t)n, f) == (size_t)n;
            fclose(f);
            uint64_t bytes = 0;
            int w = 0, h = 0, rc = JPEGB200_ERR_ARG;
            for (int attempt = 0; ok && attempt < 4; ++attempt) {       // WORKSPACE may be followed by OUTPUT: loop until neither
                rc = jpegb200_encode_bmp_to_jpeg_host(enc, in.p, (uint64_t)n, out.p, out.n, &bytes, &w, &h, st);
                if (rc == JPEGB200_ERR_WORKSPACE) jpegb200_encoder_set_bytes_per_block(enc, 184);   // very dense image
                else if (rc == JPEGB200_ERR_OUTPUT) ok = out.reserve(out.n * 4);
                else break;
            }
            if (!ok || rc != JPEGB200_OK) {
                fprintf(stderr, "Error: Failed to encode %s: %s\n", files[i].c_str(), ok ? jpegb200_last_error() : "I/O or memory");
                ++failed[t];
                continue;
            }
            const std::string o = out_name(dir, files[i]);
            FILE *g = fopen(o.c_str(), "wb");
            if (!g || fwrite(out.p, 1, bytes, g) != bytes) {
                perror("Error opening output file");
                ++failed[t];
            }
            if (g) fclose(g);
            mpix[t] += (double)w * h / 1e6;
        }
        cudaStreamDestroy(st);
        jpegb200_encoder_destroy(enc);
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < nthreads; ++t) pool.emplace_back(worker, t);
    for (auto &th : pool) th.join();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    int bad = 0;
    double mp = 0;
    for (int t = 0; t < nthreads; ++t) { bad += failed[t]; mp += mpix[t]; }
    printf("Encoded %zu of %zu files, %.1f Mpixel in %.3f s (%.1f Mpixel/s incl. file I/O and CUDA start-up; %d GPU%s, %d workers)\n",
           files.size() - bad, files.size(), mp, sec, mp / sec, ngpus, ngpus == 1 ? "" : "s", nthreads);
    return bad ? 1 : 0;
}
