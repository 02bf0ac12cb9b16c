// Batched weight preparation and gradient un-packing for the DenseNet f network.
//
// The tensor-core kernels want bf16 / packed copies of the fp32 parameters (1x1 weights cast and transposed, 3x3 weights
// packed per tap for the forward and the data gradient, the stem weights in the 8-pixel x 4-channel window order).  Doing
// that per layer cost ~400 tiny launches per step; here ONE launch walks a device-resident job table (built once per
// model, parameter addresses are stable) and fills one flat bf16 buffer, and ONE launch turns the packed fp32 gradient
// accumulators back into the parameters' (CO, CI, kh, kw) layout.
#include "gn_common.cuh"
#include <cuda_bf16.h>

struct GnPrepJob {
    const float* src;      // fp32 source (a parameter, or the packed gradient accumulator for the un-pack kinds)
    long dst_off;          // element offset in the destination buffer
    long start;            // first work item of this job (prefix sum); item = one destination element
    int kind;              // see below
    int a, b, ld;
};

enum { PREP_CAST = 0, PREP_TRANSPOSE = 1, PREP_C3PACK0 = 2, PREP_C3PACK1 = 3, PREP_STEM = 4, UNPACK_C3 = 5, UNPACK_STEM = 6 };

__device__ __forceinline__ int prep_find(const GnPrepJob* jobs, int n, long item) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (jobs[mid].start <= item) lo = mid; else hi = mid - 1;
    }
    return lo;
}

template <typename OutT>
__global__ void __launch_bounds__(256) prep_kernel(const GnPrepJob* __restrict__ jobs, int n_jobs, long total, OutT* __restrict__ dst) {
    gn_pdl_sync();
    for (long item = blockIdx.x * (long)blockDim.x + threadIdx.x; item < total; item += (long)gridDim.x * blockDim.x) {
        const GnPrepJob j = jobs[prep_find(jobs, n_jobs, item)];
        const int e = (int)(item - j.start);
        float v = 0.f;
        switch (j.kind) {
            case PREP_CAST: {            // src [a, b] -> dst [a, ld], pad columns zero
                const int r = e / j.ld, c = e - r * j.ld;
                if (c < j.b) v = j.src[(long)r * j.b + c];
                break;
            }
            case PREP_TRANSPOSE: {       // src [a, b] -> dst [b, ld >= a]
                const int c = e / j.ld, r = e - c * j.ld;
                if (r < j.a) v = j.src[(long)r * j.b + c];
                break;
            }
            case PREP_C3PACK0: {         // w [CO = a, CI = b, 3, 3] -> dst [(t*CO + co), ld >= CI]
                const int row = e / j.ld, c = e - row * j.ld;
                const int t = row / j.a, co = row - t * j.a;
                if (c < j.b) v = j.src[((long)co * j.b + c) * 9 + t];
                break;
            }
            case PREP_C3PACK1: {         // data gradient: dst [(t*CI + c), ld >= CO] = w[co, c, 2-ky, 2-kx]
                const int row = e / j.ld, co = e - row * j.ld;
                const int t = row / j.b, c = row - t * j.b;
                if (co < j.a) v = j.src[((long)co * j.b + c) * 9 + (8 - t)];
                break;
            }
            case PREP_STEM: {            // w [CO = a, 3, 7, 7] -> dst [ky][co][kxp*4 + c]
                const int k = e & 31, co = (e >> 5) % j.a, ky = (e >> 5) / j.a;
                const int kxp = k >> 2, c = k & 3;
                if (kxp >= 1 && c < 3) v = j.src[((co * 3 + c) * 7 + ky) * 7 + (kxp - 1)];
                break;
            }
            case UNPACK_C3: {            // dwp [9][CI = b][CO = a] -> dw [co][c][ky][kx]
                const int t = e % 9, c = (e / 9) % j.b, co = e / (9 * j.b);
                v = j.src[((long)t * j.b + c) * j.a + co];
                break;
            }
            case UNPACK_STEM: {          // dwq [CO = a][7*32] -> dw [co][c][ky][kx]
                const int kx = e % 7, ky = (e / 7) % 7, c = (e / 49) % 3, co = e / 147;
                v = j.src[(long)co * 224 + ky * 32 + (kx + 1) * 4 + c];
                break;
            }
        }
        dst[j.dst_off + e] = (OutT)v;
    }
}

static int prep_launch(const void* jobs, int n_jobs, long total, void* dst, int dst_bf16, cudaStream_t stream) {
    GN_REQUIRE(jobs && dst && n_jobs > 0 && total > 0, GN_EINVAL, "prepare_weights: bad arguments");
    int blocks = gn_ceil_div(total, 256);
    if (blocks > gn_num_sms() * 16) blocks = gn_num_sms() * 16;
    if (dst_bf16) GN_CUDA(gn_launch(prep_kernel<__nv_bfloat16>, dim3(blocks), dim3(256), 0, stream, (const GnPrepJob*)jobs, n_jobs, total, (__nv_bfloat16*)dst));
    else GN_CUDA(gn_launch(prep_kernel<float>, dim3(blocks), dim3(256), 0, stream, (const GnPrepJob*)jobs, n_jobs, total, (float*)dst));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// jobs: device array of n_jobs records {src fp32*, dst_off, start, kind, a, b, ld} (40 bytes each, see GnPrepJob); dst: bf16 buffer
GN_API int gn_prepare_weights(const void* jobs, int n_jobs, long total_items, void* dst_bf16, cudaStream_t stream) {
    return prep_launch(jobs, n_jobs, total_items, dst_bf16, 1, stream);
}
// same job walker with an fp32 destination: packed gradient accumulators -> parameter layout
GN_API int gn_unpack_gradients(const void* jobs, int n_jobs, long total_items, float* dst, cudaStream_t stream) {
    return prep_launch(jobs, n_jobs, total_items, dst, 0, stream);
}
GN_API int gn_prep_job_bytes(void) { return (int)sizeof(GnPrepJob); }
