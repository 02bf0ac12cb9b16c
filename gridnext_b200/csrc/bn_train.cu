// Train-mode BatchNorm around the tensor-core kernels (f pre-training, /root/reference/gridnext/training.py:11-98, and the
// count f of GridNetHexMM, which training.py:126 leaves in train mode).
//
// The tensor-core kernels evaluate BatchNorm as a per-channel affine (scale, shift) on operand load / in the epilogue and
// their BN-backward epilogues return dx_eval = g * scale together with the column sums d_beta = sum g, d_gamma = sum g*xhat.
// With batch statistics the same kernels are reused; what changes is
//   forward   the statistics come from the data:   gn_colstats_bf16  (sum, sum of squares per channel, fp64 accumulators)
//                                                  gn_bn_train_coeffs (-> scale, shift, mean, invstd; running-stat update)
//             and a GEMM cannot normalise its own output (its statistics are only complete when it ends):
//                                                  gn_affine_relu_bf16 (a = relu(z * scale + shift), bf16 -> bf16)
//   backward  dx = dx_eval - (c0 + c1 * x)  with   c1 = scale * invstd * d_gamma / M,   c0 = scale * d_beta / M - c1 * mean
//             (the mean / variance terms of the BatchNorm gradient):
//                                                  gn_bn_train_fix_coeffs (c0, c1 from the column sums; accumulating form for
//                                                  DenseNet concat channels, which several layers normalise with the SAME
//                                                  batch statistics but their own gamma/beta)
//                                                  gn_bn_train_fix_bf16   (dx -= c0 + c1 * x in place)
// All are streaming kernels: 16-byte vectors along the channel dimension, HBM-bound.
#include "gn_common.cuh"
#include "gn_epilogue.cuh"

namespace {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(w[i] << 16);
        f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 o;
    o.x = gn_pack_bf16x2(f[0], f[1]);
    o.y = gn_pack_bf16x2(f[2], f[3]);
    o.z = gn_pack_bf16x2(f[4], f[5]);
    o.w = gn_pack_bf16x2(f[6], f[7]);
    return o;
}

constexpr int CS_THREADS = 256;
constexpr int CS_ROWS_PER_LANE = 32;      // rows one thread accumulates in fp32 before the fp64 reduction

// grid (row chunks, column tiles of TG groups); thread = (row lane, 8-channel group)
__global__ void __launch_bounds__(CS_THREADS) colstats_kernel(const __nv_bfloat16* __restrict__ x, long ld, long M, int C8, int TG,
                                                              double* __restrict__ sum, double* __restrict__ sumsq) {
    gn_pdl_sync();
    __shared__ float s_part[2][CS_THREADS][9];      // [sum | sumsq][thread][8 (+1 pad)]
    const int tg = threadIdx.x % TG, rl = threadIdx.x / TG, RL = CS_THREADS / TG;
    const int grp = blockIdx.y * TG + tg;
    const long row0 = (long)blockIdx.x * RL * CS_ROWS_PER_LANE;
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
    if (grp < C8 && rl < RL) {
        for (int i = 0; i < CS_ROWS_PER_LANE; ++i) {
            const long r = row0 + (long)i * RL + rl;
            if (r >= M) break;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(x + r * ld + grp * 8));
            float f[8];
            unpack8(v, f);
#pragma unroll
            for (int e = 0; e < 8; ++e) { s[e] += f[e]; q[e] = fmaf(f[e], f[e], q[e]); }
        }
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) { s_part[0][threadIdx.x][e] = s[e]; s_part[1][threadIdx.x][e] = q[e]; }
    __syncthreads();
    // TG*8 channels x 2 quantities, each reduced over RL row lanes in fp64 by one thread
    for (int o = threadIdx.x; o < TG * 16; o += CS_THREADS) {
        const int which = o / (TG * 8), c = o % (TG * 8), g = c / 8, e = c % 8;
        if (blockIdx.y * TG + g >= C8) continue;
        double acc = 0.0;
        for (int l = 0; l < RL; ++l) acc += (double)s_part[which][l * TG + g][e];
        atomicAdd((which ? sumsq : sum) + (blockIdx.y * TG + g) * 8 + e, acc);
    }
}

__global__ void bn_train_coeffs_kernel(const double* __restrict__ sum, const double* __restrict__ sumsq, double inv_m, double unbias,
                                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps, float momentum,
                                       float* __restrict__ rmean, float* __restrict__ rvar, float* __restrict__ sc, float* __restrict__ sh,
                                       float* __restrict__ mean, float* __restrict__ invstd, int C) {
    gn_pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const double m = sum[c] * inv_m;
    double var = sumsq[c] * inv_m - m * m;
    if (var < 0.0) var = 0.0;
    const double is = 1.0 / sqrt(var + (double)eps);
    const double g = gamma ? (double)gamma[c] : 1.0, b = beta ? (double)beta[c] : 0.0;
    sc[c] = (float)(g * is);
    sh[c] = (float)(b - m * g * is);
    mean[c] = (float)m;
    invstd[c] = (float)is;
    if (rmean) rmean[c] = (1.f - momentum) * rmean[c] + momentum * (float)m;
    if (rvar) rvar[c] = (1.f - momentum) * rvar[c] + momentum * (float)(var * unbias);
}

__global__ void __launch_bounds__(256) affine_relu_kernel(const __nv_bfloat16* __restrict__ x, long ldx, __nv_bfloat16* __restrict__ y, long ldy,
                                                          long M, int C8, const float* __restrict__ sc, const float* __restrict__ sh, int relu) {
    gn_pdl_sync();
    const long total = M * C8;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / C8;
        const int g = (int)(i - r * C8);
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + r * ldx + g * 8)), f);
        const float4 s0 = __ldg(reinterpret_cast<const float4*>(sc + g * 8)), s1 = __ldg(reinterpret_cast<const float4*>(sc + g * 8 + 4));
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(sh + g * 8)), t1 = __ldg(reinterpret_cast<const float4*>(sh + g * 8 + 4));
        const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w}, t[8] = {t0.x, t0.y, t0.z, t0.w, t1.x, t1.y, t1.z, t1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            f[e] = fmaf(f[e], s[e], t[e]);
            if (relu) f[e] = fmaxf(f[e], 0.f);
        }
        *reinterpret_cast<uint4*>(y + r * ldy + g * 8) = pack8(f);
    }
}

__global__ void bn_train_fix_coeffs_kernel(const float* __restrict__ dbeta, const float* __restrict__ dgamma, const float* __restrict__ sc,
                                           const float* __restrict__ invstd, const float* __restrict__ mean, float inv_m, int accumulate,
                                           float* __restrict__ c0, float* __restrict__ c1, int C) {
    gn_pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float k1 = sc[c] * invstd[c] * dgamma[c] * inv_m;
    const float k0 = sc[c] * dbeta[c] * inv_m - k1 * mean[c];
    if (accumulate) { c0[c] += k0; c1[c] += k1; }
    else { c0[c] = k0; c1[c] = k1; }
}

__global__ void __launch_bounds__(256) bn_train_fix_kernel(__nv_bfloat16* __restrict__ dx, long lddx, const __nv_bfloat16* __restrict__ x, long ldx,
                                                           long M, int C8, const float* __restrict__ c0, const float* __restrict__ c1) {
    gn_pdl_sync();
    const long total = M * C8;
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
        const long r = i / C8;
        const int g = (int)(i - r * C8);
        float d[8], v[8];
        uint4* dp = reinterpret_cast<uint4*>(dx + r * lddx + g * 8);
        unpack8(*dp, d);
        unpack8(__ldg(reinterpret_cast<const uint4*>(x + r * ldx + g * 8)), v);
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(c0 + g * 8)), a1 = __ldg(reinterpret_cast<const float4*>(c0 + g * 8 + 4));
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(c1 + g * 8)), b1 = __ldg(reinterpret_cast<const float4*>(c1 + g * 8 + 4));
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int e = 0; e < 8; ++e) d[e] -= fmaf(b[e], v[e], a[e]);
        *dp = pack8(d);
    }
}

int stream_blocks(long items) {
    int blocks = gn_ceil_div(items, 256);
    const int cap = gn_num_sms() * 16;
    return blocks > cap ? cap : (blocks < 1 ? 1 : blocks);
}

}  // namespace

// sum[c] += sum_r x[r, c], sumsq[c] += sum_r x[r, c]^2 for the C (multiple of 8) leading channels of bf16 rows of pitch ld.
// The caller zeroes sum / sumsq (fp64[C]).
GN_API int gn_colstats_bf16(const void* x, long ld, long M, int C, double* sum, double* sumsq, cudaStream_t stream) {
    GN_REQUIRE(x && sum && sumsq && M > 0 && C > 0 && ld >= C, GN_EINVAL, "colstats_bf16: bad arguments");
    GN_REQUIRE(C % 8 == 0 && ld % 8 == 0 && ((uintptr_t)x & 15) == 0, GN_EALIGN, "colstats_bf16: C, ld must be multiples of 8 and x 16-byte aligned");
    const int C8 = C / 8;
    int TG = 1;
    while (TG < C8 && TG < 32) TG *= 2;              // 8-channel groups per CTA (power of two <= 32)
    const int RL = CS_THREADS / TG;
    dim3 grid(gn_ceil_div(M, (long)RL * CS_ROWS_PER_LANE), gn_ceil_div(C8, TG));
    GN_CUDA(gn_launch(colstats_kernel, dim3(grid), dim3(CS_THREADS), 0, stream, (const __nv_bfloat16*)x, ld, M, C8, TG, sum, sumsq));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// Batch statistics -> the per-channel constants the tensor-core kernels consume; optional running-stat update
// (running_mean / running_var may be null; running_var receives the unbiased variance, as nn.BatchNorm does).
GN_API int gn_bn_train_coeffs(const double* sum, const double* sumsq, long M, const float* gamma, const float* beta, float eps, float momentum,
                              float* running_mean, float* running_var, float* sc, float* sh, float* mean, float* invstd, int C,
                              cudaStream_t stream) {
    GN_REQUIRE(sum && sumsq && sc && sh && mean && invstd && M > 0 && C > 0, GN_EINVAL, "bn_train_coeffs: bad arguments");
    const double unbias = M > 1 ? (double)M / (double)(M - 1) : 1.0;
    GN_CUDA(gn_launch(bn_train_coeffs_kernel, dim3(gn_ceil_div(C, 128)), dim3(128), 0, stream, sum, sumsq, 1.0 / (double)M, unbias, gamma, beta, eps, momentum, running_mean,
                                                                    running_var, sc, sh, mean, invstd, C));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// y[r, c] = [relu](x[r, c] * sc[c] + sh[c]) on the C (multiple of 8) leading channels; y may alias x.
GN_API int gn_affine_relu_bf16(const void* x, long ldx, void* y, long ldy, long M, int C, const float* sc, const float* sh, int relu,
                               cudaStream_t stream) {
    GN_REQUIRE(x && y && sc && sh && M > 0 && C > 0 && ldx >= C && ldy >= C, GN_EINVAL, "affine_relu_bf16: bad arguments");
    GN_REQUIRE(C % 8 == 0 && ldx % 8 == 0 && ldy % 8 == 0 && (((uintptr_t)x | (uintptr_t)y | (uintptr_t)sc | (uintptr_t)sh) & 15) == 0, GN_EALIGN,
               "affine_relu_bf16: C and pitches must be multiples of 8, pointers 16-byte aligned");
    GN_CUDA(gn_launch(affine_relu_kernel, dim3(stream_blocks(M * (C / 8))), dim3(256), 0, stream, (const __nv_bfloat16*)x, ldx, (__nv_bfloat16*)y, ldy, M, C / 8, sc, sh, relu));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// c1 = sc * invstd * dgamma / M, c0 = sc * dbeta / M - c1 * mean; accumulate != 0 adds to c0 / c1 instead of overwriting.
GN_API int gn_bn_train_fix_coeffs(const float* dbeta, const float* dgamma, const float* sc, const float* invstd, const float* mean, long M,
                                  int accumulate, float* c0, float* c1, int C, cudaStream_t stream) {
    GN_REQUIRE(dbeta && dgamma && sc && invstd && mean && c0 && c1 && M > 0 && C > 0, GN_EINVAL, "bn_train_fix_coeffs: bad arguments");
    GN_CUDA(gn_launch(bn_train_fix_coeffs_kernel, dim3(gn_ceil_div(C, 128)), dim3(128), 0, stream, dbeta, dgamma, sc, invstd, mean, 1.f / (float)M, accumulate, c0, c1, C));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// dx[r, c] -= c0[c] + c1[c] * x[r, c] on the C (multiple of 8) leading channels, in place.
GN_API int gn_bn_train_fix_bf16(void* dx, long lddx, const void* x, long ldx, long M, int C, const float* c0, const float* c1, cudaStream_t stream) {
    GN_REQUIRE(dx && x && c0 && c1 && M > 0 && C > 0 && lddx >= C && ldx >= C, GN_EINVAL, "bn_train_fix_bf16: bad arguments");
    GN_REQUIRE(C % 8 == 0 && lddx % 8 == 0 && ldx % 8 == 0 && (((uintptr_t)dx | (uintptr_t)x | (uintptr_t)c0 | (uintptr_t)c1) & 15) == 0, GN_EALIGN,
               "bn_train_fix_bf16: C and pitches must be multiples of 8, pointers 16-byte aligned");
    GN_CUDA(gn_launch(bn_train_fix_kernel, dim3(stream_blocks(M * (C / 8))), dim3(256), 0, stream, (__nv_bfloat16*)dx, lddx, (const __nv_bfloat16*)x, ldx, M, C / 8, c0, c1));
    GN_LAUNCH_CHECK();
    return GN_OK;
}
