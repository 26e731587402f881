// synth.cuh -- synthetic workload generator on the device.
// Integer-only and counter-based, so it is bit-identical to the numpy / C versions
// (jpeg_image_compression_b200/synth.py, oracle/jpeg_oracle.c:orc_synth_rgb).
#pragma once

#include "common.cuh"

namespace jb {

__device__ __forceinline__ uint32_t synth_tri(uint32_t t, uint32_t period)
{
    const uint32_t m = t % (2u * period);
    const uint32_t d = m > period ? m - period : period - m;
    return d * 255u / period;
}

__global__ void __launch_bounds__(256)
k_synth_rgb(uint8_t *__restrict__ rgb, const int w, const int h, const uint64_t image_stride,
            const uint32_t seed0, const int amp)
{
    const uint64_t npx = (uint64_t)w * (uint64_t)h;
    const uint32_t seed = seed0 + blockIdx.y;
    const uint32_t span = (uint32_t)(2 * amp + 1);
    uint8_t *img = rgb + (uint64_t)blockIdx.y * image_stride;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < npx; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t y = (uint32_t)(i / (uint32_t)w), x = (uint32_t)(i - (uint64_t)y * (uint32_t)w);
        const uint32_t base = (synth_tri(x + 2u * y, 419u) + synth_tri(3u * x + (1u << 20) - y, 1021u) +
                               synth_tri(y, 173u) + synth_tri(x, 67u)) / 4u;
#pragma unroll
        for (uint32_t c = 0; c < 3; ++c) {
            uint32_t hsh = (x * 0x9E3779B1u) ^ (y * 0x85EBCA77u) ^ ((c * 0xC2B2AE3Du) ^ (seed * 0x27D4EB2Fu));
            hsh ^= hsh >> 15; hsh *= 0x2C1B3C6Du;
            hsh ^= hsh >> 12; hsh *= 0x297A2D39u;
            hsh ^= hsh >> 15;
            const int noise = (int)((hsh >> 24) % span) - amp;
            int v = (int)base + (c == 0 ? 10 : (c == 1 ? 0 : -10)) + noise;
            v = v < 0 ? 0 : (v > 255 ? 255 : v);
            img[i * 3u + c] = (uint8_t)v;
        }
    }
}

}  // namespace jb
