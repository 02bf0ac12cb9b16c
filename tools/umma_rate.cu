// Throughput probe (development tool): cycles per tcgen05.mma.kind::f16 (M = 128, K = 16, SS mode, K-major SWIZZLE_128B
// operands) as a function of N and of the number of independent accumulators, with a fully unrolled issue loop so the
// single issuing thread is not the limit.  Sizes the shared-memory operand bandwidth that bounds small-N implicit GEMMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/umma_rate tools/umma_rate.cu && tools/umma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../gridnext_b200/csrc/gn_ptx.cuh"
using namespace gnptx;

template <int N, int NACC>
__global__ void __launch_bounds__(128, 1) rate_kernel(int iters, long long* cycles_out) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ uint32_t s_tmem;
    __shared__ __align__(8) uint64_t bar;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    for (int i = threadIdx.x; i < 192 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;
    fence_proxy_async_smem();
    if (threadIdx.x < 32) tmem_alloc<512>(&s_tmem);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = idesc_bf16(128, N, 0, 0);
        constexpr uint64_t tmpl = smem_desc_template(0, 1024, LAYOUT_SW128);
        const uint64_t a0 = smem_desc(tmpl, smem_u32(sm)), b0 = smem_desc(tmpl, smem_u32(sm) + 128 * 1024);
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int u = 0; u < 16; ++u)      // 16 MMAs per iteration: 4 A tiles x 4 k-steps, accumulators round-robin
                umma_bf16(tmem + (uint32_t)((u % NACC) * (512 / NACC)), a0 + (uint64_t)((u / 4) * 1024 + (u % 4) * 2), b0 + (uint64_t)((u % 4) * 2), idesc, 1u);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        const long long t1 = clock64();
        if (blockIdx.x == 0) cycles_out[0] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

template <int N, int NACC>
static void run(long long* d) {
    const int smem = 193 * 1024 + 1024, iters = 512;
    cudaFuncSetAttribute(rate_kernel<N, NACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate_kernel<N, NACC><<<148, 128, smem>>>(iters, d);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
    long long c;
    cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    const double per = (double)c / (iters * 16);
    printf("{\"nacc\": %d, \"N\": %d, \"cyc_per_mma\": %.1f, \"floor\": %.1f, \"operand_B_per_clk\": %.1f, \"mac_per_clk\": %.0f}\n", NACC, N, per,
           128.0 * N / 256.0, (128 * 32 + N * 32) / per, 128.0 * N * 16 / per);
}

int main() {
    long long* d;
    cudaMalloc(&d, 8);
    run<16, 1>(d); run<32, 1>(d); run<64, 1>(d); run<128, 1>(d); run<256, 1>(d);
    run<16, 2>(d); run<32, 2>(d); run<64, 2>(d); run<128, 2>(d); run<256, 2>(d);
    run<32, 4>(d); run<64, 4>(d); run<128, 4>(d);
    return 0;
}
