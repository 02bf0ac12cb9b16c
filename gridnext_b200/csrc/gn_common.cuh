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
