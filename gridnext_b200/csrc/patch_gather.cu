// Spot-patch gather: crop a P x P window around every Visium spot of a full-resolution RGB image,
// optionally apply ToTensor (/255) + Normalize(mean, std), and write the (78, 64, 3, P, P) grid tensor.
//
// Replaces the python loop of /root/reference/gridnext/imgprocess.py:185-238 (grid_from_wsi_visium):
//   * pseudo_hex_to_oddr (imgprocess.py:26-32): x = col//2 (even row) | (col-1)//2 (odd row), y = row
//   * centre = int(np.rint(pxl))  -> rint() in fp64 here (round-half-even)        (imgprocess.py:213-214)
//   * np.pad(mode='edge') by w//2  -> coordinates are clamped to the image instead (imgprocess.py:198)
//   * patch rows [cy - w//2, cy + w//2), cols [cx - w//2, cx + w//2)                (imgprocess.py:220)
//   * out-of-tissue cells stay exactly 0                                           (imgprocess.py:206)
// Pure-crop case only (2*(w//2) == P): a same-size PIL resize is the identity.
//
// One CTA stages RPC patch rows (3*P contiguous bytes each, arbitrary byte alignment) into shared
// memory with 16-byte loads, then every thread turns 4 pixels x 3 channels into three vector
// stores, one per NCHW channel plane.  HBM-bound: 3*P*P bytes in, 3*P*P*sizeof(out) bytes out per spot.
#include "gn_common.cuh"
#include "gn_ptx.cuh"
#include "gn_tma.cuh"

// cells[n] = {cx, cy, valid} for grid cell n = y_ind * w_st + x_ind
__global__ void spot_table_kernel(const unsigned char* __restrict__ in_tissue, const int* __restrict__ array_row,
                                  const int* __restrict__ array_col, const double* __restrict__ pxl_row,
                                  const double* __restrict__ pxl_col, int n_spots, int h_st, int w_st, int* __restrict__ cells,
                                  int* __restrict__ n_dropped) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_spots) return;
    if (in_tissue[i] != 1) return;
    const int row = array_row[i], col = array_col[i];
    // python: int(col/2) for even rows, int((col-1)/2) for odd rows (true division, truncation toward zero)
    const int num = (row % 2 == 0) ? col : col - 1;
    const int x_ind = num / 2;     // C division truncates toward zero like int(float)
    const int y_ind = row;
    if (y_ind < 0 || y_ind >= h_st || x_ind < 0 || x_ind >= w_st) {
        atomicAdd(n_dropped, 1);   // reference prints a warning and skips (imgprocess.py:232-234)
        return;
    }
    int* c = cells + 3 * (y_ind * w_st + x_ind);
    c[0] = (int)rint(pxl_col[i]);
    c[1] = (int)rint(pxl_row[i]);
    c[2] = 1;
}

template <typename OutT> struct Pack4;
template <> struct Pack4<float> {
    static __device__ __forceinline__ void store(float* p, float a, float b, float c, float d) {
        *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
    }
};
template <> struct Pack4<__nv_bfloat16> {
    static __device__ __forceinline__ void store(__nv_bfloat16* p, float a, float b, float c, float d) {
        __nv_bfloat162 lo = __floats2bfloat162_rn(a, b), hi = __floats2bfloat162_rn(c, d);
        uint2 u;
        u.x = *reinterpret_cast<unsigned*>(&lo);
        u.y = *reinterpret_cast<unsigned*>(&hi);
        *reinterpret_cast<uint2*>(p) = u;
    }
};

template <typename OutT>
__global__ void __launch_bounds__(256) patch_gather_kernel(const unsigned char* __restrict__ img, long pitch, long img_bytes, int H, int W,
                                                           const int* __restrict__ cells, int P, int rpc, int row_buf,
                                                           const float* __restrict__ mean, const float* __restrict__ stdv,
                                                           OutT* __restrict__ out, int only_rest) {
    extern __shared__ __align__(128) unsigned char sm[];
    float* lut = reinterpret_cast<float*>(sm);                 // [3][256]
    unsigned char* rows = sm + 3 * 256 * sizeof(float);        // [rpc][row_buf]
    __shared__ int s_off[64];                                  // byte offset of the staged segment inside its row buffer

    const int cell = blockIdx.x;
    const int r0 = blockIdx.y * rpc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int cx = cells[3 * cell + 0], cy = cells[3 * cell + 1], valid = cells[3 * cell + 2];
    const int hw = P / 2;
    if (only_rest && valid && cx - hw >= 0 && cx - hw + P <= W && cy - hw >= 0 && cy - hw + P <= H) return;   // the TMA kernel owns this cell
    OutT* ocell = out + (long)cell * 3 * P * P;
    const int nrows = min(rpc, P - r0);
    const int groups = P / 4;

    if (!valid) {   // out-of-tissue cell: exact zeros
        for (int e = tid; e < 3 * nrows * groups; e += 256) {
            int c = e / (nrows * groups), r = (e / groups) % nrows, g = e % groups;
            Pack4<OutT>::store(ocell + ((long)c * P + r0 + r) * P + 4 * g, 0.f, 0.f, 0.f, 0.f);
        }
        return;
    }
    // value table: raw u8 -> float, or ((v / 255) - mean) / std in IEEE fp32 like ToTensor + Normalize
    for (int e = tid; e < 3 * 256; e += 256) {
        const int c = e >> 8, v = e & 255;
        float f = (float)v;
        if (mean != nullptr) f = __fdiv_rn(__fsub_rn(__fdiv_rn(f, 255.0f), mean[c]), stdv[c]);
        lut[e] = f;
    }
    const int x_lo = cx - hw;
    const int xs = max(x_lo, 0), xe = min(x_lo + P, W);       // staged source columns [xs, xe)
    for (int r = warp; r < nrows; r += 8) {
        const int gy = min(max(cy - hw + r0 + r, 0), H - 1);
        unsigned char* dst = rows + (long)r * row_buf;
        if (xe > xs) {
            const long a_begin = (long)gy * pitch + 3L * xs;
            const long a_end = (long)gy * pitch + 3L * xe;
            const long a0 = a_begin & ~15L;
            if (lane == 0) s_off[r] = (int)(a_begin - a0);
            const int nchunks = (int)((a_end - a0 + 15) >> 4);
            for (int ch = lane; ch < nchunks; ch += 32) {
                const long a = a0 + 16L * ch;
                uint4 v;
                if (a + 16 <= img_bytes) {
                    v = __ldg(reinterpret_cast<const uint4*>(img + a));
                } else {   // last bytes of the allocation: guarded byte loads
                    unsigned char tmp[16];
                    for (int q = 0; q < 16; ++q) tmp[q] = (a + q < img_bytes) ? img[a + q] : 0;
                    v = *reinterpret_cast<uint4*>(tmp);
                }
                *reinterpret_cast<uint4*>(dst + 16 * ch) = v;
            }
        } else if (lane == 0) {
            s_off[r] = 0;
        }
    }
    __syncthreads();
    // all pixels of a row outside the image horizontally (xe <= xs) cannot happen for centres inside the image,
    // but clamp anyway: source column = clamp(x_lo + px, 0, W-1); if nothing was staged read the edge pixel directly.
    for (int e = tid; e < nrows * groups; e += 256) {
        const int r = e / groups, g = e % groups;
        const unsigned char* src = rows + (long)r * row_buf + s_off[r];
        float v[3][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int sxg = min(max(x_lo + 4 * g + j, 0), W - 1);
            if (xe > xs) {
                const unsigned char* p = src + 3 * (sxg - xs);
                v[0][j] = lut[p[0]]; v[1][j] = lut[256 + p[1]]; v[2][j] = lut[512 + p[2]];
            } else {
                const int gy = min(max(cy - hw + r0 + r, 0), H - 1);
                const unsigned char* p = img + (long)gy * pitch + 3L * sxg;
                v[0][j] = lut[p[0]]; v[1][j] = lut[256 + p[1]]; v[2][j] = lut[512 + p[2]];
            }
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
            Pack4<OutT>::store(ocell + ((long)c * P + r0 + r) * P + 4 * g, v[c][0], v[c][1], v[c][2], v[c][3]);
    }
}

// ------------------------------------------------------------------------------------------------
// Persistent TMA-pipelined gather (the default whenever the image can be described by a tensor map: 16-byte row pitch, W % 4 == 0,
// P % 32 == 0).  Work item = (cell, tile of 32 patch rows); 2 CTAs per SM walk the item list with a ring of PG_STAGES shared-memory
// tiles: one thread issues the TMA load of item i + PG_STAGES while all 256 threads convert item i, so the loads of several tiles
// are always in flight and the value table (3 x 256 floats, two IEEE divisions each) is built once per CTA instead of once per
// tile.  (The first version launched one short-lived CTA per tile: 163 us for 4,992 spots at P = 128 -> bf16, 69 % of the copy
// bandwidth, with the table construction and the un-overlapped load latency of every CTA on the critical path.)
// TMA boxes must start on a 16-byte boundary of the innermost dimension (a misaligned start faults -- measured), so the image is
// viewed as rows of uint32 and the box starts at the 16-byte boundary below the window's first byte; the byte offset d (0..15, the
// same for every row of a cell) is removed with funnel shifts when a thread unpacks its 4 pixels (12 bytes).
// Cells whose window hangs over an image border (edge clamp == np.pad(mode='edge')) take clamped per-pixel loads in the same
// kernel; out-of-tissue cells are zero-filled: one launch writes the whole grid.
#define PG_RPC 32            // patch rows per tile
#define PG_STAGES 4

__device__ __forceinline__ bool pg_interior(int cx, int cy, int hw, int P, int H, int W) {
    return cx - hw >= 0 && cx - hw + P <= W && cy - hw >= 0 && cy - hw + P <= H;
}

// REP: copies of the value table, copy (lane % REP) is the one a lane reads.  The table look-ups are data dependent, so with one copy the
// 32 lanes of a look-up hit random banks: ncu counted 62 % of all shared-memory wavefronts as conflict replays and the shared-memory pipe,
// not HBM, set the pace (12 look-ups x ~3.5 wavefronts per 4 pixels).  Entry (c, v) of copy k sits at word (c*256 + v) * REP + k, i.e. in bank
// (v * REP + k) % 32: with REP = 16 only the two lanes that share a copy can collide (and only when v differs by a multiple of 2).
template <typename OutT, int REP>
__global__ void __launch_bounds__(256, 2) patch_gather_tma_kernel(const __grid_constant__ CUtensorMap tmImg, const unsigned char* __restrict__ img,
                                                                  long pitch, int H, int W, const int* __restrict__ cells, int n_cells, int P,
                                                                  int row_bytes, int stages, const float* __restrict__ mean,
                                                                  const float* __restrict__ stdv, OutT* __restrict__ out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t bar[PG_STAGES];
    const int tid = threadIdx.x;
    float* lut = reinterpret_cast<float*>(sm);                        // [3][256][REP]
    unsigned char* ring = sm + 3 * 256 * REP * sizeof(float);         // [stages][PG_RPC][row_bytes]
    const float* lutp = lut + (tid & (REP - 1));                      // this lane's copy
    const int tile_bytes = PG_RPC * row_bytes;
    const int tiles = P / PG_RPC, groups = P / 4, hw = P / 2;
    const int n_items = n_cells * tiles;

    auto issue = [&](int item, int s) {                               // thread 0 only
        const int cell = item / tiles, r0 = (item - cell * tiles) * PG_RPC;
        const int cx = __ldg(cells + 3 * cell), cy = __ldg(cells + 3 * cell + 1), valid = __ldg(cells + 3 * cell + 2);
        if (valid && pg_interior(cx, cy, hw, P, H, W)) {
            const int a0 = (3 * (cx - hw)) & ~15;
            gnptx::mbar_arrive_expect_tx(&bar[s], (uint32_t)tile_bytes);
            gnptx::tma_load_2d(&tmImg, &bar[s], ring + (size_t)s * tile_bytes, a0 >> 2, cy - hw + r0);
        } else {
            gnptx::mbar_arrive(&bar[s]);                              // nothing to load: complete the phase
        }
    };

    if (tid == 0) {
        for (int s = 0; s < stages; ++s) gnptx::mbar_init(&bar[s], 1);
        gnptx::fence_barrier_init();
    }
    gn_pdl_sync();
    // value table: raw u8 -> float, or ((v / 255) - mean) / std in IEEE fp32 like ToTensor + Normalize
    for (int e = tid; e < 3 * 256; e += 256) {
        const int c = e >> 8, v = e & 255;
        float f = (float)v;
        if (mean != nullptr) f = __fdiv_rn(__fsub_rn(__fdiv_rn(f, 255.0f), mean[c]), stdv[c]);
#pragma unroll
        for (int k = 0; k < REP; ++k) lut[e * REP + k] = f;
    }
    __syncthreads();
    if (tid == 0)
        for (int s = 0; s < stages; ++s) {
            const int item = blockIdx.x + s * gridDim.x;
            if (item < n_items) issue(item, s);
        }
    int s = 0;
    uint32_t phase = 0;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int cell = item / tiles, r0 = (item - cell * tiles) * PG_RPC;
        const int cx = __ldg(cells + 3 * cell), cy = __ldg(cells + 3 * cell + 1), valid = __ldg(cells + 3 * cell + 2);
        OutT* ocell = out + (long)cell * 3 * P * P;
        gnptx::mbar_wait(&bar[s], phase);
        if (!valid) {
            for (int e = tid; e < 3 * PG_RPC * groups; e += 256) {
                const int c = e / (PG_RPC * groups), r = (e / groups) % PG_RPC, g = e % groups;
                Pack4<OutT>::store(ocell + ((long)c * P + r0 + r) * P + 4 * g, 0.f, 0.f, 0.f, 0.f);
            }
        } else if (pg_interior(cx, cy, hw, P, H, W)) {
            // thread = 4 pixels (12 bytes): four 4-byte shared loads from the word holding the first byte, one funnel shift per word
            // (the byte phase (d & 3) is uniform), 12 table look-ups, and one 8/16-byte store per channel -- consecutive lanes write
            // consecutive addresses of the output row.
            const int b0 = 3 * (cx - hw), d = b0 - (b0 & ~15);
            const int sh = (d & 3) * 8;
            const unsigned char* rows = ring + (size_t)s * tile_bytes;
            // warp = patch rows (r = warp, warp + 8, ...), lane = 4-pixel groups: no integer division in the inner loop (the kernel is
            // bound by instruction issue, ~145 instructions per 4 pixels in the first version: ncu, profiles/r02d_ncu_full_gather_corrector.txt)
            const int warp = tid >> 5, lane = tid & 31;
            for (int r = warp; r < PG_RPC; r += 8) {
                const unsigned char* rowp = rows + (size_t)r * row_bytes;
                OutT* orow = ocell + (long)(r0 + r) * P;
                for (int g = lane; g < groups; g += 32) {
                    const unsigned* src = reinterpret_cast<const unsigned*>(rowp + ((12 * g + d) & ~3));
                    const unsigned i0 = src[0], i1 = src[1], i2 = src[2], i3 = src[3];
                    const unsigned w[3] = {__funnelshift_r(i0, i1, sh), __funnelshift_r(i1, i2, sh), __funnelshift_r(i2, i3, sh)};
                    float v[3][4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int c = 0; c < 3; ++c) {
                            const int byte = 3 * j + c;
                            v[c][j] = lutp[(c * 256 + (int)__byte_perm(w[byte >> 2], 0, 0x4440 | (byte & 3))) * REP];
                        }
#pragma unroll
                    for (int c = 0; c < 3; ++c) Pack4<OutT>::store(orow + (long)c * P * P + 4 * g, v[c][0], v[c][1], v[c][2], v[c][3]);
                }
            }
        } else {
            // window over an image border: clamped coordinates (a few dozen cells per array)
            for (int e = tid; e < PG_RPC * groups; e += 256) {
                const int r = e / groups, g = e - r * groups;
                const int gy = min(max(cy - hw + r0 + r, 0), H - 1);
                float v[3][4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int gx = min(max(cx - hw + 4 * g + j, 0), W - 1);
                    const unsigned char* px = img + (long)gy * pitch + 3L * gx;
                    v[0][j] = lutp[px[0] * REP]; v[1][j] = lutp[(256 + px[1]) * REP]; v[2][j] = lutp[(512 + px[2]) * REP];
                }
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    Pack4<OutT>::store(ocell + ((long)c * P + r0 + r) * P + 4 * g, v[c][0], v[c][1], v[c][2], v[c][3]);
            }
        }
        __syncthreads();                                              // every thread is done with stage s
        if (tid == 0) {
            const int nxt = item + stages * gridDim.x;
            if (nxt < n_items) issue(nxt, s);
        }
        if (++s == stages) { s = 0; phase ^= 1; }
    }
}

// ------------------------------------------------------------------------------------------------
GN_API int gn_spot_table(const unsigned char* in_tissue, const int* array_row, const int* array_col, const double* pxl_row,
                         const double* pxl_col, int n_spots, int h_st, int w_st, int* cells /* [h_st*w_st][3] */,
                         int* n_dropped /* [1] */, cudaStream_t stream) {
    GN_REQUIRE(in_tissue && array_row && array_col && pxl_row && pxl_col && cells && n_dropped && n_spots >= 0 && h_st > 0 && w_st > 0,
               GN_EINVAL, "spot_table: bad arguments");
    GN_CUDA(cudaMemsetAsync(cells, 0, (size_t)h_st * w_st * 3 * sizeof(int), stream));
    GN_CUDA(cudaMemsetAsync(n_dropped, 0, sizeof(int), stream));
    if (n_spots > 0) {
        spot_table_kernel<<<gn_ceil_div(n_spots, 256), 256, 0, stream>>>(in_tissue, array_row, array_col, pxl_row, pxl_col, n_spots, h_st,
                                                                          w_st, cells, n_dropped);
        GN_LAUNCH_CHECK();
    }
    return GN_OK;
}

GN_API int gn_patch_gather(const unsigned char* img, long pitch, int H, int W, const int* cells, int n_cells, int P, const float* mean,
                           const float* stdv, void* out, int out_bf16, cudaStream_t stream) {
    GN_REQUIRE(img && cells && out && H > 0 && W > 0 && n_cells > 0, GN_EINVAL, "patch_gather: bad arguments");
    GN_REQUIRE(pitch >= 3L * W, GN_EINVAL, "patch_gather: pitch %ld < 3*W", pitch);
    GN_REQUIRE(P >= 4 && P % 4 == 0 && P <= 1024, GN_EUNSUPPORTED, "patch_gather: patch size %d must be a multiple of 4 in [4, 1024]", P);
    GN_REQUIRE((mean == nullptr) == (stdv == nullptr), GN_EINVAL, "patch_gather: mean/std must come together");
    GN_REQUIRE(((uintptr_t)img & 15) == 0 && ((uintptr_t)out & 15) == 0, GN_EALIGN, "patch_gather: img/out must be 16-byte aligned");
    const int row_buf = ((3 * P + 15 + 15) / 16 + 1) * 16;
    int rpc = (40 * 1024) / row_buf;
    if (rpc > 64) rpc = 64;
    if (rpc > P) rpc = P;
    if (rpc >= 8) rpc &= ~7;
    GN_REQUIRE(rpc >= 1, GN_EUNSUPPORTED, "patch_gather: patch too wide");
    const size_t smem = 3 * 256 * sizeof(float) + (size_t)rpc * row_buf;
    const long img_bytes = (long)(H - 1) * pitch + 3L * W;
    // interior cells go through TMA when the image can be described by a tensor map of uint32 rows (16-byte pitch)
    const int row_bytes = ((3 * P + 16 + 15) / 16) * 16;          // window + up to 15 bytes of lead-in, a whole number of 16-byte units
    const bool use_tma = (P % 32 == 0) && (row_bytes / 4 <= 256) && (pitch % 16 == 0) && (W % 4 == 0) && H >= P && W >= P;
    if (use_tma) {
        CUtensorMap tm;
        uint64_t dims[2] = {(uint64_t)(3 * W / 4), (uint64_t)H};       // rows as uint32; the lead-out past 3*W is out of bounds = zero-filled, never read
        uint64_t strides[1] = {(uint64_t)pitch};
        uint32_t box[2] = {(uint32_t)(row_bytes / 4), PG_RPC};
        int rc = gn_tmap_encode(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, img, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
        // 16 table copies when three or more ring stages still fit beside them (two CTAs per SM), else 8
        const size_t ring1 = (size_t)PG_RPC * row_bytes;
        const int rep = (3 * 256 * 16 * sizeof(float) + 3 * ring1 + 16 <= 100 * 1024) ? 16 : 8;
        const size_t lut_bytes = 3 * 256 * (size_t)rep * sizeof(float);
        int stages = PG_STAGES;
        while (stages > 2 && lut_bytes + (size_t)stages * ring1 + 16 > 100 * 1024) --stages;   // two CTAs per SM
        const size_t smem_t = lut_bytes + (size_t)stages * ring1 + 16;
        const int n_items = n_cells * (P / PG_RPC);
        int grid_t = 2 * gn_num_sms();
        if (grid_t > n_items) grid_t = n_items;
#define PG_LAUNCH(T, R)                                                                                                                        \
    do {                                                                                                                                       \
        GN_CUDA(cudaFuncSetAttribute(patch_gather_tma_kernel<T, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));                \
        GN_CUDA(gn_launch(patch_gather_tma_kernel<T, R>, dim3(grid_t), dim3(256), smem_t, stream, tm, img, pitch, H, W, cells, n_cells, P,     \
                          row_bytes, stages, mean, stdv, (T*)out));                                                                            \
    } while (0)
        if (out_bf16) { if (rep == 16) PG_LAUNCH(__nv_bfloat16, 16); else PG_LAUNCH(__nv_bfloat16, 8); }
        else { if (rep == 16) PG_LAUNCH(float, 16); else PG_LAUNCH(float, 8); }
#undef PG_LAUNCH
        GN_LAUNCH_CHECK();
        return GN_OK;                                                   // the persistent kernel wrote every cell (interior, border, off-tissue)
    }
    dim3 grid(n_cells, gn_ceil_div(P, rpc));
    if (out_bf16)
        patch_gather_kernel<__nv_bfloat16><<<grid, 256, smem, stream>>>(img, pitch, img_bytes, H, W, cells, P, rpc, row_buf, mean, stdv,
                                                                        (__nv_bfloat16*)out, 0);
    else
        patch_gather_kernel<float><<<grid, 256, smem, stream>>>(img, pitch, img_bytes, H, W, cells, P, rpc, row_buf, mean, stdv, (float*)out, 0);
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// ------------------------------------------------------------------------------------------------
// ToTensor (/255) + Normalize(mean, std) of an already-cropped uint8 patch grid (cells, 3, P, P) -> bf16 | fp32,
// the device half of PatchGridDataset.__getitem__ (/root/reference/gridnext/image_datasets.py:205-232 applies the
// torchvision transform per present patch; absent cells stay exactly 0).  valid: nullable u8[cells].
template <typename OutT>
__global__ void __launch_bounds__(256) normalize_u8_kernel(const unsigned char* __restrict__ in, const unsigned char* __restrict__ valid,
                                                           long n_cells, long plane, const float* __restrict__ mean,
                                                           const float* __restrict__ stdv, OutT* __restrict__ out) {
    __shared__ float lut[3 * 256];
    for (int e = threadIdx.x; e < 3 * 256; e += 256) {
        const int c = e >> 8, v = e & 255;
        float f = (float)v;
        if (mean != nullptr) f = __fdiv_rn(__fsub_rn(__fdiv_rn(f, 255.0f), mean[c]), stdv[c]);
        lut[e] = f;
    }
    __syncthreads();
    const long groups = n_cells * 3 * plane / 16;
    for (long g = blockIdx.x * (long)blockDim.x + threadIdx.x; g < groups; g += (long)gridDim.x * blockDim.x) {
        const long e0 = g * 16;
        const long cell = e0 / (3 * plane);
        const int c = (int)((e0 / plane) % 3);
        const bool ok = valid == nullptr || valid[cell] != 0;
        const uint4 raw = ok ? __ldg(reinterpret_cast<const uint4*>(in + e0)) : make_uint4(0, 0, 0, 0);
        const unsigned w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            float f[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) f[j] = ok ? lut[c * 256 + ((w[q] >> (8 * j)) & 0xff)] : 0.f;
            Pack4<OutT>::store(out + e0 + 4 * q, f[0], f[1], f[2], f[3]);
        }
    }
}

GN_API int gn_normalize_u8(const unsigned char* in, const unsigned char* valid, long n_cells, int P, const float* mean, const float* stdv,
                           void* out, int out_bf16, cudaStream_t stream) {
    GN_REQUIRE(in && out && n_cells > 0 && P > 0 && (P * P) % 16 == 0, GN_EINVAL, "normalize_u8: bad arguments (P*P must be a multiple of 16)");
    GN_REQUIRE((mean == nullptr) == (stdv == nullptr), GN_EINVAL, "normalize_u8: mean/std must come together");
    GN_REQUIRE(((uintptr_t)in & 15) == 0 && ((uintptr_t)out & 15) == 0, GN_EALIGN, "normalize_u8: in/out must be 16-byte aligned");
    const long plane = (long)P * P;
    long blocks = (n_cells * 3 * plane / 16 + 255) / 256;
    if (blocks > 16L * gn_num_sms()) blocks = 16L * gn_num_sms();
    if (out_bf16)
        normalize_u8_kernel<__nv_bfloat16><<<(unsigned)blocks, 256, 0, stream>>>(in, valid, n_cells, plane, mean, stdv, (__nv_bfloat16*)out);
    else
        normalize_u8_kernel<float><<<(unsigned)blocks, 256, 0, stream>>>(in, valid, n_cells, plane, mean, stdv, (float*)out);
    GN_LAUNCH_CHECK();
    return GN_OK;
}
