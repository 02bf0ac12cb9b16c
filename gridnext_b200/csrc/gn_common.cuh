// Shared helpers for the gridnext_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#define GN_API extern "C" __attribute__((visibility("default")))

// Error convention of the C-ABI (include/gridnext_b200.h): 0 ok, <0 argument error, >0 cudaError_t.
enum { GN_OK = 0, GN_EINVAL = -1, GN_EUNSUPPORTED = -2, GN_EALIGN = -3, GN_EDRIVER = -4 };

void gn_set_error(const char* fmt, ...);

#define GN_REQUIRE(cond, code, ...)            \
    do {                                       \
        if (!(cond)) {                         \
            gn_set_error(__VA_ARGS__);         \
            return (code);                     \
        }                                      \
    } while (0)

#define GN_CUDA(call)                                                                  \
    do {                                                                               \
        cudaError_t e__ = (call);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            gn_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return (int)e__;                                                           \
        }                                                                              \
    } while (0)

#define GN_LAUNCH_CHECK() GN_CUDA(cudaGetLastError())

static inline int gn_ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

int gn_num_sms();

// development A/B switches: an environment variable set to a non-zero integer
static inline int gn_env_flag(const char* name) {
    const char* e = getenv(name);
    return e && atoi(e) != 0;
}

// ---- programmatic dependent launch ---------------------------------------------------------------
// A step is a fixed sequence of 300+ short launches; between two of them the GPU otherwise idles for the launch latency, the drain of
// the last CTAs and the next kernel's prologue (barrier init, tensor-memory allocation, descriptor fetch).  Kernels launched through
// gn_launch() carry cudaLaunchAttributeProgrammaticStreamSerialization: their CTAs may become resident as soon as every CTA of the
// preceding kernel has executed gn_pdl_trigger() (or exited), run their prologue, and block in gn_pdl_wait() until the preceding grid
// has completed and its memory is visible.  Rules that keep this equivalent to stream order:
//   * EVERY thread of a kernel launched this way executes gn_pdl_wait() before its first global-memory access and before any exit
//     (a grid that completed without waiting would let ITS successor overtake the predecessor's predecessor);
//   * gn_pdl_trigger() comes after the wait (so at most two grids overlap) and after the tensor-memory allocation (a resident successor
//     CTA that already holds tensor memory must never be waited on by a predecessor CTA still asking for it).
// GN_NO_PDL=1 launches without the attribute (the device instructions are then no-ops).
__device__ __forceinline__ void gn_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void gn_pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void gn_pdl_sync() {
    gn_pdl_wait();
    gn_pdl_trigger();
}

int gn_pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t gn_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = gn_pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

__device__ __forceinline__ float gn_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double gn_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
