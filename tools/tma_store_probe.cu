// Probe: which TMA store forms are legal on sm_100a?  One CTA loads a box with cp.async.bulk.tensor.4d (global ->
// shared, SWIZZLE_128B) at a chosen shared-memory byte offset and stores it back with the 4-D store form at chosen
// coordinates.  Driven by tools/tma_store_probe.py, one case per process (a faulting form kills the context).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../gridnext_b200/csrc/gn_ptx.cuh"
using namespace gnptx;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmOut, int smem_off, int box_bytes, int lc0, int lc1,
             int lc2, int lc3, int sc0, int sc1, int sc2, int sc3, int nrep, int rep_stride_bytes, int rep_dy, int store_skip) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    if (threadIdx.x == 0) {
        mbar_init(&bar, 1);
        fence_barrier_init();
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&bar, (uint32_t)(box_bytes * nrep));
        for (int r = 0; r < nrep; ++r) tma_load_4d(&tmIn, &bar, sm + smem_off + r * rep_stride_bytes, lc0, lc1, lc2 + r * rep_dy, lc3);
        mbar_wait(&bar, 0);
        for (int r = 0; r < nrep; ++r) tma_store_4d(&tmOut, sm + smem_off + store_skip + r * rep_stride_bytes, sc0, sc1, sc2 + r * rep_dy, sc3);
        tma_store_commit();
        tma_store_wait_all<0>();
    }
    __syncthreads();
}

typedef CUresult (*encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                              const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int make_map(CUtensorMap* m, void* ptr, int C, int W, int H, int N, long ld, int boxc, int boxw) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess) return -1;
    cuuint64_t gd[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
    cuuint64_t gs[3] = {(cuuint64_t)ld * 2, (cuuint64_t)W * ld * 2, (cuuint64_t)H * W * ld * 2};
    cuuint32_t bx[4] = {(cuuint32_t)boxc, (cuuint32_t)boxw, 1, 1}, es[4] = {1, 1, 1, 1};
    CUresult r = ((encode_fn)p)(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, ptr, gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1000 + (int)r;
}

extern "C" __attribute__((visibility("default"))) int tma_store_probe(void* in, void* out, int Cin, int Cout, int W, int H, int N, long ld_in,
                                                                     long ld_out, int boxw, int smem_off, int lc0, int lc1, int lc2, int lc3, int sc0,
                                                                     int sc1, int sc2, int sc3, int nrep, int rep_dy, int boxw_store, int store_skip) {
    CUtensorMap tmIn, tmOut;
    int rc = make_map(&tmIn, in, Cin, W, H, N, ld_in, 64, boxw);
    if (rc) return rc;
    rc = make_map(&tmOut, out, Cout, W, H, N, ld_out, 64, boxw_store);
    if (rc) return rc;
    const int box_bytes = 128 * boxw;
    const int smem = 1024 + smem_off + nrep * box_bytes + 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<<<1, 128, smem>>>(tmIn, tmOut, smem_off, box_bytes, lc0, lc1, lc2, lc3, sc0, sc1, sc2, sc3, nrep, box_bytes, rep_dy, store_skip);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { fprintf(stderr, "probe: %s\n", cudaGetErrorString(e)); return (int)e; }
    return 0;
}
