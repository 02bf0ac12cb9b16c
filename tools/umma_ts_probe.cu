// Probe: tcgen05.mma with the A operand in TENSOR MEMORY (written with tcgen05.st), B in shared memory (K-major, SWIZZLE_128B).
// Settles the A layout hypothesis "lane = row m, 32-bit column c of a K step holds elements k = 2c (low half), 2c + 1 (high half)".
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_ts_probe tools/umma_ts_probe.cu ; run on a B200.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1) probe(const uint32_t* a_words /* [128][32] */, const uint8_t* b_img /* 64 rows x 128 B */, int N, float* D) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) unsigned long long s_bar;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < N * 128; i += 128) sm[i] = b_img[i];
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&s_tmem)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&s_bar)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t tmem = s_tmem;
    // A row of this thread -> TMEM columns 256..287 of lane tid
    uint32_t r[32];
    for (int c = 0; c < 32; ++c) r[c] = a_words[tid * 32 + c];
    const uint32_t ta = tmem + ((uint32_t)(warp * 32) << 16) + 256u;
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,"
        "%26,%27,%28,%29,%30,%31,%32};" ::"r"(ta),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]),
        "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]),
        "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (tid == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;");
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
        const unsigned long long tmpl = ((unsigned long long)((1024 >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
        for (int ks = 0; ks < 4; ++ks) {
            const unsigned long long db = tmpl | (unsigned long long)(((base + ks * 32) >> 4) & 0x3FFF);
            const uint32_t a_t = tmem + 256u + (uint32_t)(ks * 8);
            const uint32_t acc = ks > 0;
            asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;}" ::"r"(tmem), "r"(a_t),
                         "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)));
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(done) : "r"(smem_u32(&s_bar)), "r"(0u));
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c0 = 0; c0 < N; c0 += 8) {
        uint32_t o[8];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(o[0]), "=r"(o[1]), "=r"(o[2]), "=r"(o[3]), "=r"(o[4]), "=r"(o[5]), "=r"(o[6]), "=r"(o[7])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int j = 0; j < 8; ++j) D[tid * N + c0 + j] = __uint_as_float(o[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)((u + 0x7fff + ((u >> 16) & 1)) >> 16); }
static float bf2f(uint16_t h) { uint32_t u = (uint32_t)h << 16; float f; memcpy(&f, &u, 4); return f; }

int main() {
    const int M = 128, N = 64, K = 64;
    static float A[128][64], B[64][64];
    srand(5);
    for (int m = 0; m < M; ++m) for (int k = 0; k < K; ++k) A[m][k] = bf2f(f2bf((rand() % 17 - 8) / 8.0f));
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) B[n][k] = bf2f(f2bf((rand() % 13 - 6) / 4.0f));
    static uint32_t aw[128 * 32];
    for (int m = 0; m < M; ++m) for (int c = 0; c < 32; ++c) aw[m * 32 + c] = (uint32_t)f2bf(A[m][2 * c]) | ((uint32_t)f2bf(A[m][2 * c + 1]) << 16);
    static uint8_t bimg[64 * 128];
    for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) {
        const int chunk = (k >> 3) ^ (n & 7);
        uint16_t v = f2bf(B[n][k]);
        memcpy(&bimg[n * 128 + chunk * 16 + (k & 7) * 2], &v, 2);
    }
    uint32_t* d_a; uint8_t* d_b; float* d_D;
    cudaMalloc(&d_a, sizeof(aw)); cudaMalloc(&d_b, sizeof(bimg)); cudaMalloc(&d_D, M * N * 4);
    cudaMemcpy(d_a, aw, sizeof(aw), cudaMemcpyHostToDevice); cudaMemcpy(d_b, bimg, sizeof(bimg), cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    probe<<<1, 128, 32 * 1024>>>(d_a, d_b, N, d_D);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("{\"probe\": \"umma_a_in_tmem\", \"error\": \"%s\"}\n", cudaGetErrorString(e)); return 1; }
    static float D[128 * 64];
    cudaMemcpy(D, d_D, sizeof(D), cudaMemcpyDeviceToHost);
    double worst = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0;
        for (int k = 0; k < K; ++k) ref += (double)A[m][k] * B[n][k];
        worst = fmax(worst, fabs(ref - D[m * N + n]));
    }
    printf("{\"probe\": \"umma_a_in_tmem\", \"hypothesis\": \"lane=row, column c of a K step = (k=2c | k=2c+1 << 16), K step = 8 columns\", \"max_abs_err\": %g, \"ok\": %s}\n",
           worst, worst < 1e-3 ? "true" : "false");
    return 0;
}
