// Hardware-semantics probe for tcgen05.mma shared-memory descriptors on sm_100a (development tool).
// The host hands in raw shared-memory images for A and B plus fully formed descriptors; the kernel
// issues the MMAs and dumps the TMEM accumulator.  tools/umma_probe.py builds the images for each
// layout hypothesis and checks the result, so one GPU run settles which encodings are right.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define A_REGION 0
#define B_REGION (96 * 1024)
#define SMEM_TOTAL (192 * 1024)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__global__ void __launch_bounds__(128, 1)
probe_kernel(const uint8_t* a_img, int a_bytes, const uint8_t* b_img, int b_bytes, uint32_t a_off, uint32_t b_off,
             unsigned long long a_tmpl, unsigned long long b_tmpl, uint32_t idesc, int ksteps, int a_step, int b_step,
             int kind_tf32, int N, float* D) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) unsigned long long s_bar;
    const uint32_t raw = smem_u32(smem_raw);
    const uint32_t base = (raw + 1023u) & ~1023u;
    uint8_t* sm = smem_raw + (base - raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    for (int i = tid; i < SMEM_TOTAL / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x7fc00000u;  // NaN fill: catches stray reads
    __syncthreads();
    for (int i = tid; i < a_bytes; i += 128) sm[A_REGION + i] = a_img[i];
    for (int i = tid; i < b_bytes; i += 128) sm[B_REGION + i] = b_img[i];
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

    if (tid == 0) {
        for (int ks = 0; ks < ksteps; ++ks) {
            unsigned long long da = a_tmpl | (unsigned long long)(((base + A_REGION + a_off + ks * a_step) >> 4) & 0x3FFF);
            unsigned long long db = b_tmpl | (unsigned long long)(((base + B_REGION + b_off + ks * b_step) >> 4) & 0x3FFF);
            uint32_t acc = ks > 0;
            if (kind_tf32)
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;}" ::"r"(tmem),
                             "l"(da), "l"(db), "r"(idesc), "r"(acc));
            else
                asm volatile("{.reg .pred p; setp.ne.b32 p, %4, 0; tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;}" ::"r"(tmem),
                             "l"(da), "l"(db), "r"(idesc), "r"(acc));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&s_bar)));
    }
    // wait phase 0
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}"
                         : "=r"(done)
                         : "r"(smem_u32(&s_bar)), "r"(0u));
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;");
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
            "%25,%26,%27,%28,%29,%30,%31}, [%32];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
              "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
              "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
              "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;");
        for (int j = 0; j < 32; ++j)
            if (c0 + j < N) D[(warp * 32 + lane) * N + c0 + j] = __uint_as_float(r[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

extern "C" __attribute__((visibility("default"))) int probe_run(const void* a_img, int a_bytes, const void* b_img, int b_bytes,
                                                                 unsigned a_off, unsigned b_off, unsigned long long a_tmpl,
                                                                 unsigned long long b_tmpl, unsigned idesc, int ksteps, int a_step,
                                                                 int b_step, int kind_tf32, int N, float* D, void* stream) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_TOTAL + 1024);
        if (e != cudaSuccess) return (int)e;
        attr = true;
    }
    if (a_bytes > B_REGION || b_bytes > SMEM_TOTAL - B_REGION) return -1;
    probe_kernel<<<1, 128, SMEM_TOTAL + 1024, (cudaStream_t)stream>>>((const uint8_t*)a_img, a_bytes, (const uint8_t*)b_img, b_bytes, a_off,
                                                                       b_off, a_tmpl, b_tmpl, idesc, ksteps, a_step, b_step, kind_tf32, N, D);
    return (int)cudaGetLastError();
}
