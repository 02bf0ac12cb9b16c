// Small streaming kernels around the tensor-core GEMMs of the count MLP f (notebooks/Tutorial_visium_count.ipynb cell 12:
// Linear(G,500) Linear(500,100) BN1d ReLU Linear(100,100) Linear(100,50) BN1d ReLU Linear(50,n_cls)).
//   gn_cast_f32_bf16     the fp32 count slab (B, G, 78, 64) -> bf16, consumed in place by gn_gemm_tn_bf16 (no permute copy,
//                        /root/reference/gridnext/gridnet_models.py:168,83 force one)
//   gn_rows_affine_bf16  y = [relu](x * scale[c] + shift[c]) from the fp32 split-K accumulator of layer 1 to a bf16 row-major
//                        activation with a padded pitch (pad columns zero)
#include "gn_common.cuh"
#include "gn_epilogue.cuh"

__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float4* __restrict__ in, uint2* __restrict__ out, long n4, const float* in_tail,
                                                            __nv_bfloat16* out_tail, int tail) {
    gn_pdl_sync();
    for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n4; i += (long)gridDim.x * blockDim.x) {
        const float4 v = in[i];
        uint2 o;
        o.x = gn_pack_bf16x2(v.x, v.y);
        o.y = gn_pack_bf16x2(v.z, v.w);
        out[i] = o;
    }
    if (blockIdx.x == 0 && threadIdx.x < tail) out_tail[threadIdx.x] = __float2bfloat16_rn(in_tail[threadIdx.x]);
}

GN_API int gn_cast_f32_bf16(const float* in, void* out, long n, cudaStream_t stream) {
    GN_REQUIRE(in && out && n > 0, GN_EINVAL, "cast_f32_bf16: bad arguments");
    GN_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 7) == 0, GN_EALIGN, "cast_f32_bf16: buffers must be 16 / 8 byte aligned");
    const long n4 = n / 4;
    const int tail = (int)(n - 4 * n4);
    int blocks = gn_ceil_div(n4 > 0 ? n4 : 1, 256);
    if (blocks > gn_num_sms() * 16) blocks = gn_num_sms() * 16;
    GN_CUDA(gn_launch(cast_f32_bf16_kernel, dim3(blocks), dim3(256), 0, stream, (const float4*)in, (uint2*)out, n4, in + 4 * n4, (__nv_bfloat16*)out + 4 * n4, tail));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

__global__ void __launch_bounds__(256) rows_affine_bf16_kernel(const float* __restrict__ in, long ldi, const float* __restrict__ scale,
                                                               const float* __restrict__ shift, int relu, __nv_bfloat16* __restrict__ out, long ldo,
                                                               long N, int C, int Cpad) {
    gn_pdl_sync();
    const long total = N * Cpad;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const long r = e / Cpad;
        const int c = (int)(e - r * Cpad);
        float v = 0.f;
        if (c < C) {
            v = in[r * ldi + c];
            v = fmaf(v, scale ? __ldg(scale + c) : 1.f, shift ? __ldg(shift + c) : 0.f);
            if (relu) v = fmaxf(v, 0.f);
        }
        out[r * ldo + c] = __float2bfloat16_rn(v);
    }
}

// in: fp32 [N, ldi] (C columns used); out: bf16 [N, ldo], columns [C, Cpad) zeroed (Cpad <= ldo)
GN_API int gn_rows_affine_bf16(const float* in, long ldi, const float* scale, const float* shift, int relu, void* out, long ldo, long N, int C,
                               int Cpad, cudaStream_t stream) {
    GN_REQUIRE(in && out && N > 0 && C > 0 && Cpad >= C && ldo >= Cpad && ldi >= C, GN_EINVAL, "rows_affine_bf16: bad arguments");
    int blocks = gn_ceil_div(N * Cpad, 256);
    if (blocks > gn_num_sms() * 16) blocks = gn_num_sms() * 16;
    GN_CUDA(gn_launch(rows_affine_bf16_kernel, dim3(blocks), dim3(256), 0, stream, in, ldi, scale, shift, relu, (__nv_bfloat16*)out, ldo, N, C, Cpad));
    GN_LAUNCH_CHECK();
    return GN_OK;
}
