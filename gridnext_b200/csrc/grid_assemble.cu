// Dataset tensor assembly on the device: the per-array python loops that place spot data on the
// (h_st, w_st) grid before GridNet sees it.
//   /root/reference/gridnext/utils.py:144-166          read_annotated_starray: counts_grid[y, x] = cmat[spot], annots_grid[y, x] = label + 1
//   /root/reference/gridnext/image_datasets.py:205-232 PatchGridDataset.__getitem__: patch_grid[y, x] = patch, annots_grid[y, x] = label + 1
//   /root/reference/gridnext/multimodal_datasets.py:237-244  78x64 python double loop of .max() calls ("spots must have image data and
//                                                      annotations to be marked as foreground")
// The grid is written ONCE, cell by cell (gather through the inverse spot->cell map), instead of zero-filled and then scattered:
// every output byte is touched exactly once with coalesced stores, and duplicates resolve like the reference's sequential loop
// (the last spot wins).  All kernels are HBM-bound byte movers.
#include "gn_common.cuh"

namespace {

__global__ void fill_i32_kernel(int* p, int n, int v) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = v;
}

__global__ void cell_inverse_kernel(const int* __restrict__ cell, int n_rows, int* __restrict__ inv, int n_cells) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n_rows; r += gridDim.x * blockDim.x) {
        const int c = cell[r];
        if (c >= 0 && c < n_cells) atomicMax(inv + c, r);
    }
}

// dst[c, :] = inv[c] >= 0 ? src[inv[c], :] : 0;  one CTA per cell, VEC-byte words
template <typename V>
__global__ void __launch_bounds__(256) gather_rows_kernel(const unsigned char* __restrict__ src, long src_pitch, const int* __restrict__ inv,
                                                          unsigned char* __restrict__ dst, long dst_pitch, long n_cells, long row_words) {
    for (long c = blockIdx.x; c < n_cells; c += gridDim.x) {
        const int r = inv[c];
        V* d = reinterpret_cast<V*>(dst + c * dst_pitch);
        if (r >= 0) {
            const V* s = reinterpret_cast<const V*>(src + (long)r * src_pitch);
            for (long i = threadIdx.x; i < row_words; i += blockDim.x) d[i] = s[i];
        } else {
            V z;
            memset(&z, 0, sizeof(V));
            for (long i = threadIdx.x; i < row_words; i += blockDim.x) d[i] = z;
        }
    }
}

// dst[g, c] = inv[c] >= 0 ? src[g, inv[c]] : 0   (fp32; genes x spots -> genes x cells: the channels-first count slab)
__global__ void __launch_bounds__(256) gather_cols_kernel(const float* __restrict__ src, long src_pitch, const int* __restrict__ inv,
                                                          float* __restrict__ dst, long dst_pitch, int G, int n_cells) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const int r = inv[c];
    for (int g = blockIdx.y; g < G; g += gridDim.y) dst[(long)g * dst_pitch + c] = r >= 0 ? __ldg(src + (long)g * src_pitch + r) : 0.f;
}

__global__ void grid_labels_kernel(const long long* __restrict__ labels, const int* __restrict__ inv, long long* __restrict__ annots, int n_cells) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells) return;
    const int r = inv[c];
    long long a = 0;
    if (r >= 0 && labels) {
        const long long l = labels[r];
        a = l >= 0 ? l + 1 : 0;            // 0 is reserved for background / un-annotated
    }
    annots[c] = a;
}

// One CTA per cell: pmax = max(patch[c, :]) with torch.max's NaN propagation; empty = (pmax == 0);
//   annots[c] = empty ? 0 : annots[c];  flags[c] bit0 = empty (counts column must be zeroed);  patch[c, :] = 0 if annots[c] == 0.
__global__ void __launch_bounds__(256) mm_fg_patch_kernel(float* __restrict__ patch, long F, long long* __restrict__ annots,
                                                          unsigned char* __restrict__ flags, int n_cells) {
    __shared__ float s_m[8];
    __shared__ int s_zero;
    for (int c = blockIdx.x; c < n_cells; c += gridDim.x) {
        float* row = patch + (long)c * F;
        float m = -INFINITY;
        for (long i = threadIdx.x; i < F; i += blockDim.x) {
            const float v = row[i];
            m = (v > m || v != v) ? v : m;         // NaN sticks, like torch.max
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float v = __shfl_xor_sync(0xffffffffu, m, o);
            m = (v > m || v != v) ? v : m;
        }
        if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) {
            float mm = s_m[0];
            for (int w = 1; w < (int)(blockDim.x >> 5); ++w) {
                const float v = s_m[w];
                mm = (v > mm || v != v) ? v : mm;
            }
            const bool empty = mm == 0.f;
            long long a = annots[c];
            if (empty) a = 0;
            annots[c] = a;
            flags[c] = empty ? 1 : 0;
            s_zero = a == 0;
        }
        __syncthreads();
        if (s_zero)
            for (long i = threadIdx.x; i < F; i += blockDim.x) row[i] = 0.f;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) mm_fg_counts_kernel(float* __restrict__ counts, long pitch, int G, const unsigned char* __restrict__ flags,
                                                           int n_cells) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_cells || !flags[c]) return;
    for (int g = blockIdx.y; g < G; g += gridDim.y) counts[(long)g * pitch + c] = 0.f;
}

}  // namespace

// inv[c] = the LAST row r with cell[r] == c, or -1.  cell[r] < 0 (or >= n_cells) drops the spot.
GN_API int gn_cell_inverse(const int* cell, int n_rows, int* inv, int n_cells, cudaStream_t stream) {
    GN_REQUIRE(inv && n_cells > 0 && n_rows >= 0 && (cell || n_rows == 0), GN_EINVAL, "cell_inverse: bad arguments");
    fill_i32_kernel<<<gn_ceil_div(n_cells, 256), 256, 0, stream>>>(inv, n_cells, -1);
    GN_LAUNCH_CHECK();
    if (n_rows > 0) {
        cell_inverse_kernel<<<gn_ceil_div(n_rows, 256), 256, 0, stream>>>(cell, n_rows, inv, n_cells);
        GN_LAUNCH_CHECK();
    }
    return GN_OK;
}

// dst[c, 0:row_bytes] = inv[c] >= 0 ? src[inv[c], 0:row_bytes] : 0 for every cell (patch grids, spot-major feature grids).
GN_API int gn_grid_gather_rows(const void* src, long src_pitch_bytes, const int* inv, void* dst, long dst_pitch_bytes, long n_cells,
                               long row_bytes, cudaStream_t stream) {
    GN_REQUIRE(dst && inv && n_cells > 0 && row_bytes > 0 && dst_pitch_bytes >= row_bytes, GN_EINVAL, "grid_gather_rows: bad arguments");
    const int grid = (int)(n_cells < 148L * 32 ? n_cells : 148L * 32);
    const uintptr_t a = (uintptr_t)src | (uintptr_t)dst | (uintptr_t)src_pitch_bytes | (uintptr_t)dst_pitch_bytes | (uintptr_t)row_bytes;
    if ((a & 15) == 0)
        gather_rows_kernel<uint4><<<grid, 256, 0, stream>>>((const unsigned char*)src, src_pitch_bytes, inv, (unsigned char*)dst, dst_pitch_bytes,
                                                            n_cells, row_bytes / 16);
    else if ((a & 3) == 0)
        gather_rows_kernel<uint32_t><<<grid, 256, 0, stream>>>((const unsigned char*)src, src_pitch_bytes, inv, (unsigned char*)dst,
                                                               dst_pitch_bytes, n_cells, row_bytes / 4);
    else
        gather_rows_kernel<unsigned char><<<grid, 256, 0, stream>>>((const unsigned char*)src, src_pitch_bytes, inv, (unsigned char*)dst,
                                                                    dst_pitch_bytes, n_cells, row_bytes);
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// dst[g, c] = inv[c] >= 0 ? src[g, inv[c]] : 0: (genes x spots) count matrix -> channels-first (G, h_st*w_st) slab.
GN_API int gn_grid_gather_cols(const float* src, long src_pitch, const int* inv, float* dst, long dst_pitch, int G, int n_cells,
                               cudaStream_t stream) {
    GN_REQUIRE(dst && inv && G > 0 && n_cells > 0 && dst_pitch >= n_cells, GN_EINVAL, "grid_gather_cols: bad arguments");
    dim3 grid(gn_ceil_div(n_cells, 256), G < 1024 ? G : 1024);
    gather_cols_kernel<<<grid, 256, 0, stream>>>(src, src_pitch, inv, dst, dst_pitch, G, n_cells);
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// annots[c] = inv[c] >= 0 && labels[inv[c]] >= 0 ? labels[inv[c]] + 1 : 0  (labels nullable: all background)
GN_API int gn_grid_labels(const long long* labels, const int* inv, long long* annots, int n_cells, cudaStream_t stream) {
    GN_REQUIRE(inv && annots && n_cells > 0, GN_EINVAL, "grid_labels: bad arguments");
    grid_labels_kernel<<<gn_ceil_div(n_cells, 256), 256, 0, stream>>>(labels, inv, annots, n_cells);
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// multimodal_datasets.py:237-244 in place: patch (n_cells x F) fp32, counts (G x n_cells, pitch in elements) fp32, annots int64,
// flags: n_cells bytes of workspace.
GN_API int gn_mm_fg_consistency(float* patch, long F, float* counts, long counts_pitch, int G, long long* annots, unsigned char* flags,
                                int n_cells, cudaStream_t stream) {
    GN_REQUIRE(patch && annots && flags && F > 0 && n_cells > 0 && (counts || G == 0) && counts_pitch >= (G ? n_cells : 0), GN_EINVAL,
               "mm_fg_consistency: bad arguments");
    mm_fg_patch_kernel<<<n_cells < 148 * 16 ? n_cells : 148 * 16, 256, 0, stream>>>(patch, F, annots, flags, n_cells);
    GN_LAUNCH_CHECK();
    if (G > 0) {
        dim3 grid(gn_ceil_div(n_cells, 256), G < 1024 ? G : 1024);
        mm_fg_counts_kernel<<<grid, 256, 0, stream>>>(counts, counts_pitch, G, flags, n_cells);
        GN_LAUNCH_CHECK();
    }
    return GN_OK;
}
