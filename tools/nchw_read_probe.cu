// Read-bandwidth probe (development tool): how fast can 148 persistent CTAs bring an fp32 NCHW tensor (256 x 32 x 78 x 64, the g network's
// activation at the C4 corner) into shared memory with 1-D bulk copies, as a function of the SHAPE of a job -- CG channels x R grid rows,
// one copy of R * 256 contiguous bytes per channel -- and of the ring depth S?  The hex convolution kernels need all 32 channels of a cell
// together, i.e. 32 chunks 19,968 bytes apart per job; the question is whether the chunk length (R = 2: 512 B) bounds them.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tools/nchw_read_probe tools/nchw_read_probe.cu && tools/nchw_read_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../gridnext_b200/csrc/gn_ptx.cuh"
using namespace gnptx;

__global__ void __launch_bounds__(64, 1) probe_kernel(const float* x, int B, int C, int H, int W, int CG, int R, int S, int order) {
    extern __shared__ __align__(128) uint8_t sm[];
    __shared__ __align__(8) uint64_t full[16], empty[16];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        fence_barrier_init();
    }
    __syncthreads();
    const int rb = H / R, cg = C / CG;
    const long n_jobs = (long)B * rb * cg;
    const uint32_t chunk = (uint32_t)(R * W * 4), stage = chunk * CG;
    // order 0: a CTA owns a contiguous range of jobs (array-major, then row block, then channel group); order 1: jobs interleaved over CTAs
    const long j0 = order ? blockIdx.x : n_jobs * blockIdx.x / gridDim.x, j1 = order ? n_jobs : n_jobs * (blockIdx.x + 1) / gridDim.x;
    const long step = order ? gridDim.x : 1;
    if (warp == 0) {
        uint32_t k = 0;
        for (long j = j0; j < j1; j += step, ++k) {
            const uint32_t s = k % S;
            if (k >= (uint32_t)S) mbar_wait(&empty[s], ((k / S) - 1) & 1);
            const int g = (int)(j % cg), r = (int)((j / cg) % rb), b = (int)(j / ((long)cg * rb));
            if (lane == 0) mbar_arrive_expect_tx(&full[s], stage);
            __syncwarp();
            if (lane < CG) bulk_load_1d(sm + (size_t)s * stage + (size_t)lane * chunk, x + (((long)b * C + g * CG + lane) * H + r * R) * W, chunk, &full[s]);
        }
    } else if (lane == 0) {
        uint32_t k = 0;
        for (long j = j0; j < j1; j += step, ++k) {
            const uint32_t s = k % S;
            mbar_wait(&full[s], (k / S) & 1);
            mbar_arrive(&empty[s]);
        }
    }
}

int main() {
    const int B = 256, C = 32, H = 78, W = 64;
    const size_t bytes = (size_t)B * C * H * W * 4;
    float* x;
    cudaMalloc(&x, bytes);
    cudaMemset(x, 0, bytes);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int cfgs[][4] = {{32, 2, 4, 0}, {32, 2, 8, 0}, {32, 2, 12, 0}, {32, 2, 8, 1}, {32, 6, 4, 0}, {16, 6, 8, 0}, {8, 26, 3, 0}, {4, 26, 6, 0}, {4, 78, 2, 0},
                           {2, 78, 4, 0}, {1, 78, 8, 0}, {1, 78, 8, 1}, {32, 1, 8, 0}, {16, 2, 16, 0}, {8, 2, 16, 0}};
    for (auto& c : cfgs) {
        const int CG = c[0], R = c[1], S = c[2], order = c[3];
        const size_t smem = (size_t)S * CG * R * W * 4;
        if (smem > 200 * 1024) { printf("skip\n"); continue; }
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            probe_kernel<<<148, 64, smem>>>(x, B, C, H, W, CG, R, S, order);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf("{\"channels_per_job\": %d, \"rows\": %d, \"chunk_bytes\": %d, \"stages\": %d, \"in_flight_kb\": %d, \"order\": %d, \"ms\": %.4f, \"gbs\": %.0f}\n", CG, R,
               R * W * 4, S, (int)(smem / 1024), order, best, bytes / best / 1e6);
    }
    return 0;
}
