// Probe for the tensor-core DCT of K1 (tcgen05 / TMEM, sm_100a).
//
// Question 1 (semantics): does   D[128 x 128] (fp32, TMEM) = A[128 x 64] (fp16, TMEM) * B[128 x 64]^T (fp16, smem,
// K-major, no swizzle)   issued as 4 x tcgen05.mma.kind::f16 (M128 N128 K16) give the expected sums with
//   A row r = TMEM lane r, K elements packed two per 32-bit column,
//   B core matrices (8 rows x 16 bytes) at  (n/8)*SBO + (k/8)*LBO ?
// Question 2 (exactness): A holds integers in [-128,127], B integers in [-1024,1024] ("limbs" of a 22-bit fixed-point
// weight): all products and partial sums are integers below 2^24, so an fp32 accumulator that only ever rounds
// inexact results reproduces the integer sums exactly.  The probe checks that bit for bit, including the
// extreme patterns (all |a| = 128, all |b| = 1024, alternating signs).
// Question 3 (cost): cycles per [tcgen05.st A -> 4 MMA -> commit -> wait -> 4 x tcgen05.ld] round trip.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_dct_probe tc_dct_probe.cu ; run: ./tc_dct_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t n)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(n) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

#define TMEM_ST32(taddr, r)                                                                                            \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16," \
                 "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};\n" ::"r"(taddr),                   \
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),   \
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]),       \
                 "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]),      \
                 "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                   \
                 : "memory")

#define TMEM_LD32(taddr, r)                                                                                            \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"       \
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"                          \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),     \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),           \
                   "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),         \
                   "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),         \
                   "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                              \
                 : "r"(taddr)                                                                                          \
                 : "memory")

constexpr uint32_t IDESC = (1u << 4)            // D format: fp32
                           | (0u << 7)          // A format: fp16
                           | (0u << 10)         // B format: fp16
                           | (0u << 15)         // A: K-major
                           | (0u << 16)         // B: K-major
                           | ((128u >> 3) << 17)   // N = 128
                           | ((128u >> 4) << 24);  // M = 128

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo)
{
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo >> 4) & 0x3FFFu) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFFu) << 32) |
           (1ull << 46);                        // descriptor version 1 (Blackwell), no swizzle, base offset 0
}

// one CTA of 128 threads; thread t owns row t of A and of D
__global__ void __launch_bounds__(128) probe(const int8_t *__restrict__ a, const uint4 *__restrict__ bblob, float *__restrict__ d,
                                             uint32_t lbo, uint32_t sbo, uint32_t kstep_bytes, int iters, long long *cycles)
{
    __shared__ __align__(128) uint4 s_b[1024];      // 16 KB: B in the canonical K-major layout
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 1024; i += 128) s_b[i] = bblob[i];
    asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");     // generic-proxy writes -> visible to the MMA's async proxy
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(&s_tmem)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
    const uint32_t tbase = s_tmem;
    const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
    const uint32_t col_d = 0, col_a = 128;

    // A row -> 32 half2 words (k, k+1)
    uint32_t ar[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
        const __half2 h = __halves2half2(__int2half_rn((int)a[tid * 64 + 2 * j]), __int2half_rn((int)a[tid * 64 + 2 * j + 1]));
        ar[j] = *reinterpret_cast<const uint32_t *>(&h);
    }
    uint32_t parity = 0;
    uint32_t r[4][32];
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        TMEM_ST32(tbase + lane_base + col_a, ar);
        asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                const uint64_t bdesc = make_desc(smem_u32(s_b) + ks * kstep_bytes, lbo, sbo);
                const uint32_t accumulate = ks > 0 ? 1u : 0u;
                asm volatile(
                    "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
                    "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n}\n" ::"r"(tbase + col_d),
                    "r"(tbase + col_a + 8u * ks), "l"(bdesc), "r"(IDESC), "r"(accumulate)
                    : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(smem_u32(&s_bar)) : "memory");
        }
        mbar_wait(&s_bar, parity);
        parity ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory");
#pragma unroll
        for (int c = 0; c < 4; ++c) TMEM_LD32(tbase + lane_base + col_d + 32u * c, r[c]);
        asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory");
        ar[0] ^= (r[0][0] & 0u);                       // keep the loop body alive without changing A
    }
    const long long t1 = clock64();
    if (tid == 0 && cycles) *cycles = t1 - t0;
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int j = 0; j < 32; ++j) d[tid * 128 + 32 * c + j] = __uint_as_float(r[c][j]);
    asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tbase), "r"(256u) : "memory");
}

static int run_case(const char *name, const std::vector<int8_t> &A, const std::vector<int> &B, uint32_t lbo, uint32_t sbo, bool verbose)
{
    // B[n][k] -> blob: [kstep 4][n/8 16][kchunk 2][n%8 8][8 halves]  (kstep = 4096 B, SBO = 256, LBO = 128)
    std::vector<__half> blob(128 * 64);
    for (int n = 0; n < 128; ++n)
        for (int k = 0; k < 64; ++k) {
            const int ks = k / 16, kc = (k % 16) / 8, ke = k % 8;
            blob[(size_t)ks * 2048 + (n / 8) * 128 + kc * 64 + (n % 8) * 8 + ke] = __float2half((float)B[n * 64 + k]);
        }
    int8_t *da;
    uint4 *db;
    float *dd;
    long long *dc;
    cudaMalloc(&da, A.size());
    cudaMalloc(&db, 16384);
    cudaMalloc(&dd, 128 * 128 * 4);
    cudaMalloc(&dc, 8);
    cudaMemcpy(da, A.data(), A.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(db, blob.data(), 16384, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0xFF, 128 * 128 * 4);
    probe<<<1, 128>>>(da, db, dd, lbo, sbo, 4096, 1, dc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("%s: CUDA error %s\n", name, cudaGetErrorString(e));
        return -1;
    }
    std::vector<float> D(128 * 128);
    cudaMemcpy(D.data(), dd, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 128; ++n) {
            long long s = 0;
            for (int k = 0; k < 64; ++k) s += (long long)A[m * 64 + k] * B[n * 64 + k];
            if ((double)D[m * 128 + n] != (double)s) {
                if (verbose && bad < 6) printf("  mismatch m=%d n=%d got %.1f want %lld\n", m, n, D[m * 128 + n], s);
                ++bad;
            }
        }
    printf("%s (lbo=%u sbo=%u): %d / 16384 mismatches\n", name, lbo, sbo, bad);
    // timing
    probe<<<1, 128>>>(da, db, dd, lbo, sbo, 4096, 2000, dc);
    cudaDeviceSynchronize();
    long long cyc = 0;
    cudaMemcpy(&cyc, dc, 8, cudaMemcpyDeviceToHost);
    if (verbose) printf("  round trip (st A + 4 MMA + commit + wait + 4 ld): %.1f cycles\n", (double)cyc / 2000.0);
    cudaFree(da); cudaFree(db); cudaFree(dd); cudaFree(dc);
    return bad;
}

int main()
{
    std::vector<int8_t> A(128 * 64);
    std::vector<int> B(128 * 64);
    srand(12345);
    for (auto &v : A) v = (int8_t)(rand() % 256 - 128);
    for (auto &v : B) v = rand() % 2049 - 1024;
    int rc = run_case("random", A, B, 128, 256, true);
    if (rc != 0) {
        printf("trying swapped LBO/SBO interpretation\n");
        run_case("random-swapped", A, B, 256, 128, true);
    }
    // extremes: every product +-2^17, partial sums up to 2^23
    for (auto &v : A) v = -128;
    for (auto &v : B) v = -1024;
    rc |= run_case("all -128 x -1024 (sum 2^23)", A, B, 128, 256, true);
    for (size_t i = 0; i < A.size(); ++i) A[i] = (i & 1) ? 127 : -128;
    for (size_t i = 0; i < B.size(); ++i) B[i] = (i % 3) ? 1023 : -1024;
    rc |= run_case("alternating extremes", A, B, 128, 256, true);
    // many random trials for exactness
    int total_bad = 0;
    for (int trial = 0; trial < 20; ++trial) {
        for (auto &v : A) v = (int8_t)(rand() % 256 - 128);
        for (auto &v : B) v = rand() % 2049 - 1024;
        const int b = run_case("trial", A, B, 128, 256, false);
        total_bad += b < 0 ? 1000000 : b;
    }
    printf("20 random trials: %d mismatches in total\n", total_bad);
    return (rc || total_bad) ? 1 : 0;
}
