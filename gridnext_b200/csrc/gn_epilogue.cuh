// Shared epilogue pieces of the tcgen05 kernels: one thread owns one output row and 32 consecutive columns.
#pragma once
#include "gn_common.cuh"

__device__ __forceinline__ uint32_t gn_pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// 32 floats -> bf16 row segment (vector path needs a 16-byte aligned destination)
__device__ __forceinline__ void gn_store_bf16_32(__nv_bfloat16* o, const float (&v)[32], int ncols) {
    if (ncols == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint4 t;
            t.x = gn_pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
            t.y = gn_pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
            t.z = gn_pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
            t.w = gn_pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
            *reinterpret_cast<uint4*>(o + 8 * q) = t;
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < ncols) o[j] = __float2bfloat16_rn(v[j]);
    }
}

__device__ __forceinline__ void gn_load_bf16_32(const __nv_bfloat16* p, float (&v)[32], int ncols) {
    if (ncols == 32 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint4 t = *reinterpret_cast<const uint4*>(p + 8 * q);
            const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
                v[8 * q + 2 * e] = f.x;
                v[8 * q + 2 * e + 1] = f.y;
            }
        }
    } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = j < ncols ? __bfloat162float(p[j]) : 0.f;
    }
}

// Column sums over the 32 rows held by a warp: on return lane l holds sum_rows v[l].
// Butterfly with halving: 16 + 8 + 4 + 2 + 1 = 31 shuffles for 32 columns.
__device__ __forceinline__ float gn_warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        const bool up = lane & 16;
        const float send = up ? v[j] : v[j + 16];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 16);
        v[j] = (up ? v[j + 16] : v[j]) + recv;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const bool up = lane & 8;
        const float send = up ? v[j] : v[j + 8];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 8);
        v[j] = (up ? v[j + 8] : v[j]) + recv;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const bool up = lane & 4;
        const float send = up ? v[j] : v[j + 4];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
        v[j] = (up ? v[j + 4] : v[j]) + recv;
    }
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const bool up = lane & 2;
        const float send = up ? v[j] : v[j + 2];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
        v[j] = (up ? v[j + 2] : v[j]) + recv;
    }
    {
        const bool up = lane & 1;
        const float send = up ? v[0] : v[1];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
        v[0] = (up ? v[1] : v[0]) + recv;
    }
    return v[0];
}

// Backward through  a = relu(bn(ref))  fused into a data-gradient epilogue (eval-mode BatchNorm, training.py:126).
//   ref_is_raw = 1: ref is the PRE-BN tensor:   a = ref * sc + sh,  xhat = (ref - p0) * p1   (p0 = mean, p1 = invstd)
//   ref_is_raw = 0: ref is the activated tensor: a = ref,            xhat = (ref - p0) * p1   (p0 = beta, p1 = 1/gamma)
//   g = acc * [a > 0];   out (=|+=) g * sc;   colsum[n] += sum_rows g,  colsum[ldsum + n] += sum_rows g * xhat
struct BnBwdEpi {
    const __nv_bfloat16* ref;
    long ldref;
    int ref_is_raw;
    const float* sc;
    const float* sh;
    const float* p0;
    const float* p1;
    float* colsum;     // [2][ldsum] fp32, atomically accumulated
    int ldsum;
    int rmw;           // out += instead of out =
};
