// Spot-patch gather WITH resize: window_size != patch_size in grid_from_wsi_visium
// (/root/reference/gridnext/imgprocess.py:188-195,220-221): the 2*(w//2)-pixel window around every spot is resized to
// (P, P) by ``Image.fromarray(patch).resize((P, P))`` -- Pillow's default BICUBIC resampler for RGB images.
//
// Pillow (src/libImaging/Resample.c, 8 bits per channel) works in FIXED POINT, so the result is reproducible bit for bit:
//   * per output coordinate: support = 2 * max(in/out, 1), taps [xmin, xmin + n) around centre (xx + 0.5) * in/out, bicubic
//     weights (a = -0.5) normalised to sum 1 in double, then rounded to int32 with 22 fractional bits
//     (the host computes this table exactly as precompute_coeffs / normalize_coeffs_8bpc do and passes it in);
//   * horizontal pass over all input rows into a uint8 intermediate: out = clip8((2^21 + sum pix * k) >> 22);
//   * vertical pass over the intermediate, same arithmetic.
// Edge padding by w//2 (np.pad mode='edge', imgprocess.py:198) is coordinate clamping.  Output: raw values as float, or
// ToTensor (/255) + Normalize(mean, std) in IEEE fp32 like the crop kernel; out-of-tissue cells stay 0.
//
// One CTA = one cell x one tile of output rows: the input rows that tile needs are resampled horizontally into shared
// memory (uint8 [rows][P][3]), then resampled vertically and written as three NCHW planes.
#include "gn_common.cuh"

#define RS_PRECISION_BITS 22

__device__ __forceinline__ int rs_clip8(int v) {
    v >>= RS_PRECISION_BITS;
    return v < 0 ? 0 : (v > 255 ? 255 : v);
}

template <typename OutT>
__global__ void __launch_bounds__(256) patch_resize_kernel(const unsigned char* __restrict__ img, long pitch, int H, int W,
                                                           const int* __restrict__ cells, int ws, int P, const int* __restrict__ bounds,
                                                           const int* __restrict__ kk, int ksize, int rows_per_tile, int max_in_rows,
                                                           const float* __restrict__ mean, const float* __restrict__ stdv,
                                                           OutT* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char sm[];
    float* lut = reinterpret_cast<float*>(sm);                  // [3][256]
    unsigned char* tmp = sm + 3 * 256 * sizeof(float);          // [max_in_rows][P][3]
    const int cell = blockIdx.x;
    const int y0 = blockIdx.y * rows_per_tile;
    const int ny = min(rows_per_tile, P - y0);
    const int tid = threadIdx.x;
    const int cx = cells[3 * cell + 0], cy = cells[3 * cell + 1], valid = cells[3 * cell + 2];
    OutT* ocell = out + (long)cell * 3 * P * P;
    if (!valid) {
        for (int e = tid; e < 3 * ny * P; e += 256) {
            const int c = e / (ny * P), r = (e / P) % ny, x = e % P;
            ocell[((long)c * P + y0 + r) * P + x] = OutT(0.f);
        }
        return;
    }
    for (int e = tid; e < 3 * 256; e += 256) {
        const int c = e >> 8, v = e & 255;
        float f = (float)v;
        if (mean != nullptr) f = __fdiv_rn(__fsub_rn(__fdiv_rn(f, 255.0f), mean[c]), stdv[c]);
        lut[e] = f;
    }
    const int hw = ws / 2;
    const int in_lo = bounds[2 * y0];                                        // first input row of the window this tile needs
    const int in_hi = bounds[2 * (y0 + ny - 1)] + bounds[2 * (y0 + ny - 1) + 1];
    const int n_in = in_hi - in_lo;                                          // <= max_in_rows (host-checked)
    // ---- horizontal pass: tmp[r][xx][c] for window rows in_lo .. in_hi-1
    for (int e = tid; e < n_in * P; e += 256) {
        const int r = e / P, xx = e % P;
        const int gy = min(max(cy - hw + in_lo + r, 0), H - 1);
        const unsigned char* row = img + (long)gy * pitch;
        const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
        const int* k = kk + (long)xx * ksize;
        int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0;
        for (int t = 0; t < n; ++t) {
            const int gx = min(max(cx - hw + xmin + t, 0), W - 1);
            const int kv = __ldg(k + t);
            s0 += (int)row[3 * gx + 0] * kv;
            s1 += (int)row[3 * gx + 1] * kv;
            s2 += (int)row[3 * gx + 2] * kv;
        }
        unsigned char* d = tmp + ((long)r * P + xx) * 3;
        d[0] = (unsigned char)rs_clip8(s0);
        d[1] = (unsigned char)rs_clip8(s1);
        d[2] = (unsigned char)rs_clip8(s2);
    }
    __syncthreads();
    // ---- vertical pass + value map + NCHW store (consecutive threads -> consecutive x of one channel plane)
    for (int e = tid; e < 3 * ny * P; e += 256) {
        const int c = e / (ny * P), r = (e / P) % ny, xx = e % P;
        const int yy = y0 + r;
        const int ymin = bounds[2 * yy], n = bounds[2 * yy + 1];
        const int* k = kk + (long)yy * ksize;
        int s = 1 << (RS_PRECISION_BITS - 1);
        for (int t = 0; t < n; ++t) s += (int)tmp[((long)(ymin - in_lo + t) * P + xx) * 3 + c] * __ldg(k + t);
        ocell[((long)c * P + yy) * P + xx] = OutT(lut[c * 256 + rs_clip8(s)]);
    }
}

GN_API int gn_patch_gather_resize(const unsigned char* img, long pitch, int H, int W, const int* cells, int n_cells, int ws, int P,
                                  const int* bounds, const int* kk, int ksize, int max_span, const float* mean, const float* stdv,
                                  void* out, int out_bf16, cudaStream_t stream) {
    GN_REQUIRE(img && cells && bounds && kk && out && n_cells > 0 && H > 0 && W > 0 && pitch >= 3L * W, GN_EINVAL, "patch_gather_resize: bad arguments");
    GN_REQUIRE(ws >= 2 && ws % 2 == 0 && P >= 1 && ksize >= 1 && max_span >= 1, GN_EINVAL, "patch_gather_resize: window %d / patch %d", ws, P);
    GN_REQUIRE((mean == nullptr) == (stdv == nullptr), GN_EINVAL, "patch_gather_resize: mean and std go together");
    // rows of the output per CTA: as many as keep the uint8 intermediate within ~96 KB of shared memory
    // (input rows needed by t output rows <= max_span + (t - 1) * ceil(ws / P) + 1, max_span = the widest single-row support)
    const int step = (ws + P - 1) / P + 1;
    int rows_per_tile = 16;
    while (rows_per_tile > 1 && (long)(max_span + (rows_per_tile - 1) * step) * P * 3 > 96 * 1024) rows_per_tile >>= 1;
    const int max_in_rows = max_span + (rows_per_tile - 1) * step;
    const size_t smem = 3 * 256 * sizeof(float) + (size_t)max_in_rows * P * 3;
    GN_REQUIRE(smem <= 200 * 1024, GN_EUNSUPPORTED, "patch_gather_resize: window %d -> patch %d needs %zu B of shared memory per row tile", ws, P, smem);
    dim3 grid(n_cells, (P + rows_per_tile - 1) / rows_per_tile);
    if (out_bf16) {
        GN_CUDA(cudaFuncSetAttribute(patch_resize_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        patch_resize_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(img, pitch, H, W, cells, ws, P, bounds, kk, ksize, rows_per_tile, max_in_rows,
                                                                        mean, stdv, (__nv_bfloat16*)out);
    } else {
        GN_CUDA(cudaFuncSetAttribute(patch_resize_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        patch_resize_kernel<float><<<grid, 256, smem, stream>>>(img, pitch, H, W, cells, ws, P, bounds, kk, ksize, rows_per_tile, max_in_rows, mean,
                                                                stdv, (float*)out);
    }
    GN_LAUNCH_CHECK();
    return GN_OK;
}
