// Micro-benchmark: issue rate of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a, alone and
// interleaved with integer ALU work (PRMT), at K1's occupancy (16 warps/SM) and at 64 warps/SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o ffma2 ffma2.cu ; run: ./ffma2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int ITERS = 4096, CHAINS = 8;

template <int MODE>   // 0 FFMA, 1 FFMA2, 2 FFMA+PRMT, 3 FFMA2+PRMT, 4 PRMT only
__global__ void k(float *out, float a, float b, uint32_t sel)
{
    float x[CHAINS * 2];
    uint32_t p[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS * 2; ++i) x[i] = threadIdx.x * 0.001f + i;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) p[i] = threadIdx.x + i;
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (MODE == 0 || MODE == 2) {
                x[2 * i] = fmaf(x[2 * i], a, b);
                x[2 * i + 1] = fmaf(x[2 * i + 1], a, b);
            }
            if (MODE == 1 || MODE == 3) {
                unsigned long long v, aa, bb;
                asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(x[2 * i]), "f"(x[2 * i + 1]));
                asm("mov.b64 %0, {%1, %1};" : "=l"(aa) : "f"(a));
                asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(v) : "l"(aa), "l"(bb));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(x[2 * i]), "=f"(x[2 * i + 1]) : "l"(v));
            }
            if (MODE >= 2) {
                p[i] = __byte_perm(p[i], sel, 0x3201);
                p[i] = __byte_perm(p[i], sel, 0x2103);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS * 2; ++i) s += x[i];
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += (float)p[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char *name, int ctas_per_sm, int threads, float *out)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int grid = 148 * ctas_per_sm;
    k<MODE><<<grid, threads>>>(out, 1.0001f, 0.5f, 0x01020304u);
    cudaEventRecord(e0);
    k<MODE><<<grid, threads>>>(out, 1.0001f, 0.5f, 0x01020304u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warps = (double)grid * threads / 32;
    const double fp_flop_pairs = warps * ITERS * CHAINS;      // one "pair" = 2 scalar FFMA or 1 FFMA2
    printf("%-14s warps/SM %2d: %.3f ms  -> %.2f fp32 FMA lanes/clk/SM @1.9GHz equiv (pairs %.3g)\n", name,
           ctas_per_sm * threads / 32, ms, fp_flop_pairs * 2 * 32 / (ms * 1e-3) / 148 / 1.9e9, fp_flop_pairs);
}

int main()
{
    float *out;
    cudaMalloc(&out, 148 * 8 * 1024 * sizeof(float));
    for (int occ = 0; occ < 2; ++occ) {
        const int ctas = occ ? 4 : 2, thr = occ ? 512 : 256;
        run<0>("FFMA", ctas, thr, out);
        run<1>("FFMA2", ctas, thr, out);
        run<2>("FFMA+PRMT", ctas, thr, out);
        run<3>("FFMA2+PRMT", ctas, thr, out);
        run<4>("PRMT", ctas, thr, out);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
