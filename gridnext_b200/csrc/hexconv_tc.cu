// Hexagonal convolution (kernel_size 1, <= 32 channels) on tcgen05 with fp32-grade accuracy.
//
// Replaces hexagdly.Conv2d.forward and its data gradient as used by the g network
// (/root/reference/gridnext/gridnet_models.py:128-148) for large batches, where the 7-tap, 32-channel convolution is
// far above the FP32-FMA ridge (56 FLOP/B vs ~11) and only tensor cores can make it bandwidth-bound.  fp32 accuracy
// (north_star: 1e-5) comes from a bf16 x 3 split: x = x_hi + x_lo, w = w_hi + w_lo (bf16 each, 16 mantissa bits
// together), y = sum x_hi w_hi + x_lo w_hi + x_hi w_lo with fp32 accumulation in TMEM.
//
// Layout trick: the hex neighbourhood depends on the parity of the Visium row, which breaks the "every tap is a constant
// row shift of one operand buffer" property an implicit GEMM needs.  So the input is first rewritten (one streaming
// kernel, which also applies the previous layer's BatchNorm+ReLU, gridnet_models.py:134-136) into two PLANES per array,
// even rows and odd rows, each row padded to W+2 slots whose last two are zero, every slot = 32 hi | 32 lo bf16 (128 B):
//     planes[b][parity][q][x'][64]          position index inside a plane  s = q*(W+2) + x'
// In plane coordinates every tap of an even-row output is a constant offset into the even plane (same row: -1, 0, +1)
// or the odd plane (-(W+2)-1, -(W+2), -1, 0), and symmetrically for odd-row outputs; row ends read the zero slots,
// rows outside the grid are TMA out-of-bounds zero-fill.  A tile is 128 consecutive positions of BOTH planes: two TMA
// boxes (one window per plane), 2 parities x 7 taps x (4 + 2) MMAs of 128 x 32 x 16, outputs written straight back to
// NCHW fp32 (consecutive lanes = consecutive x: coalesced per channel) with the bias and the BatchNorm batch
// statistics (fp64) in the epilogue.
//
//   warp 0: TMA producer   warp 1: MMA issuer   warps 2-9: epilogue (TMEM lane group x output parity)
#include "gn_common.cuh"
#include "gn_ptx.cuh"
#include "gn_tma.cuh"
#include "gn_epilogue.cuh"

using namespace gnptx;

#define HTC_C 32                 // channels per operand half (hi | lo)
#define HTC_TAPS 7
#define HTC_W_TILE 4096          // one weight tile: 32 rows (cout) x 128 B
#define HTC_W_BYTES (HTC_TAPS * 2 * HTC_W_TILE)

struct HexTcParams {
    int B, H, W, Cout, Q, WP, PS, tiles_per_img, n_tiles, win_rows, stages;
    const float* bias;
    float* y;
    double* stats;
    int row_off[2][HTC_TAPS];    // [output parity][tap]: first buffer row of the tap's operand
    int src[2][HTC_TAPS];        // [output parity][tap]: 0 = even-plane window, 1 = odd-plane window
};

// ---------------------------------------------------------------------------------------------- NCHW fp32 -> planes
// One thread per slot x' of one grid row: the 32 channel values are read coalesced across the warp (consecutive x), split
// into hi | lo in registers and written as the slot's 128 contiguous bytes; slots W, W+1 and rows >= H are zeros.
__global__ void __launch_bounds__(128) hex_to_planes_kernel(const float* __restrict__ x, const float* __restrict__ in_scale,
                                                            const float* __restrict__ in_shift, int Cin, int H, int W, int Q, int WP,
                                                            __nv_bfloat16* __restrict__ planes) {
    gn_pdl_sync();
    const int yy = blockIdx.x, b = blockIdx.y;                               // yy in [0, 2Q): rows >= H are written as zeros
    const int par = yy & 1, q = yy >> 1;
    uint4* dst_row = reinterpret_cast<uint4*>(planes + (((long)b * 2 + par) * Q + q) * WP * 64);
    const long plane = (long)H * W;
    const float* src = x + ((long)b * Cin * H + yy) * W;
    for (int xx = threadIdx.x; xx < WP; xx += blockDim.x) {
        uint32_t hi[HTC_C / 2], lo[HTC_C / 2];
#pragma unroll
        for (int j = 0; j < HTC_C / 2; ++j) { hi[j] = 0u; lo[j] = 0u; }
        if (xx < W && yy < H) {
#pragma unroll
            for (int j = 0; j < HTC_C / 2; ++j) {
                float v0 = 0.f, v1 = 0.f;
                if (2 * j < Cin) v0 = src[(long)(2 * j) * plane + xx];
                if (2 * j + 1 < Cin) v1 = src[(long)(2 * j + 1) * plane + xx];
                if (in_scale != nullptr) {
                    if (2 * j < Cin) v0 = fmaxf(fmaf(v0, __ldg(in_scale + 2 * j), __ldg(in_shift + 2 * j)), 0.f);
                    if (2 * j + 1 < Cin) v1 = fmaxf(fmaf(v1, __ldg(in_scale + 2 * j + 1), __ldg(in_shift + 2 * j + 1)), 0.f);
                }
                const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
                const float2 hf = __bfloat1622float2(h);
                const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - hf.x, v1 - hf.y);
                hi[j] = *reinterpret_cast<const uint32_t*>(&h);
                lo[j] = *reinterpret_cast<const uint32_t*>(&l);
            }
        }
        uint4* d = dst_row + (long)xx * 8;
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] = make_uint4(hi[4 * j], hi[4 * j + 1], hi[4 * j + 2], hi[4 * j + 3]);
#pragma unroll
        for (int j = 0; j < 4; ++j) d[4 + j] = make_uint4(lo[4 * j], lo[4 * j + 1], lo[4 * j + 2], lo[4 * j + 3]);
    }
}

// wp fp32 [7][cin][cout] -> wtc bf16 [7][2][32 (cout)][64]: tile 0 = [w_hi | w_hi], tile 1 = [w_lo | 0]
__global__ void hex_pack_tc_kernel(const float* __restrict__ wp, int cin, int cout, __nv_bfloat16* __restrict__ wtc) {
    gn_pdl_sync();
    const int total = HTC_TAPS * 2 * HTC_C * 64;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int kk = e & 63, co = (e >> 6) & 31, j = (e >> 11) & 1, t = e >> 12;
        const int ci = kk & 31;
        float w = 0.f;
        if (ci < cin && co < cout) w = wp[((long)t * cin + ci) * cout + co];
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        float v;
        if (j == 0) v = __bfloat162float(hi);
        else v = kk < HTC_C ? w - __bfloat162float(hi) : 0.f;
        wtc[e] = __float2bfloat16_rn(v);
    }
}

// ---------------------------------------------------------------------------------------------- the convolution
__global__ void __launch_bounds__(320, 1)
hexconv_tc_kernel(const __grid_constant__ CUtensorMap tmP, const __grid_constant__ CUtensorMap tmW, const HexTcParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_w, bar_full[2], bar_empty[2], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_slot;
    __shared__ double s_stat[2 * HTC_C];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int win_bytes = ((p.win_rows * 128 + 1023) / 1024) * 1024;
    uint8_t* s_w = sm;
    uint8_t* s_a = sm + HTC_W_BYTES;                       // [stages][2 windows][win_bytes]

    if (threadIdx.x < 2 * HTC_C) s_stat[threadIdx.x] = 0.0;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmP);
        tma_prefetch_desc(&tmW);
        mbar_init(&bar_w, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
            mbar_init(&bar_tfull[s], 1);
            mbar_init(&bar_tempty[s], 8);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<128>(&tmem_slot);
    gn_pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    gn_pdl_trigger();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(&bar_w, HTC_W_BYTES);
            for (int i = 0; i < HTC_TAPS * 2; ++i) tma_load_2d(&tmW, &bar_w, s_w + i * HTC_W_TILE, 0, i * HTC_C);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const int b = tile / p.tiles_per_img, s0 = (tile - b * p.tiles_per_img) * 128;
                mbar_wait(&bar_empty[stage], phase ^ 1);
                uint8_t* st = s_a + (size_t)stage * 2 * win_bytes;
                mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)(2 * p.win_rows * 128));
                tma_load_3d(&tmP, &bar_full[stage], st, 0, s0 - 1, 2 * b);                       // even-plane window
                tma_load_3d(&tmP, &bar_full[stage], st + win_bytes, 0, s0 - p.WP - 1, 2 * b + 1);  // odd-plane window
                if (p.stages == 2) { stage ^= 1; if (stage == 0) phase ^= 1; } else { phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = idesc_bf16(128, HTC_C, 0, 0);
            const uint64_t tmpl = smem_desc_template(0, 1024, LAYOUT_SW128);
            const uint64_t descW = smem_desc(tmpl, smem_u32(s_w));
            mbar_wait(&bar_w, 0);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
                mbar_wait(&bar_full[stage], phase);
                tc_fence_after();
                const uint32_t st = smem_u32(s_a + (size_t)stage * 2 * win_bytes);
                const uint64_t descA[2] = {smem_desc(tmpl, st), smem_desc(tmpl, st + win_bytes)};
#pragma unroll
                for (int par = 0; par < 2; ++par) {
                    const uint32_t d = tmem_base + (uint32_t)((acc * 2 + par) * HTC_C);
                    uint32_t first = 1;
#pragma unroll
                    for (int t = 0; t < HTC_TAPS; ++t) {
                        const uint64_t da = descA[p.src[par][t]] + (uint64_t)(p.row_off[par][t] * 8);
                        const uint64_t dw = descW + (uint64_t)(t * 2 * (HTC_W_TILE >> 4));
#pragma unroll
                        for (int k = 0; k < 4; ++k) {       // [x_hi | x_lo] . [w_hi | w_hi]
                            umma_bf16(d, da + (uint64_t)(k * 2), dw + (uint64_t)(k * 2), idesc, first ^ 1u);
                            first = 0;
                        }
#pragma unroll
                        for (int k = 0; k < 2; ++k)         // x_hi . w_lo
                            umma_bf16(d, da + (uint64_t)(k * 2), dw + (uint64_t)((HTC_W_TILE >> 4) + k * 2), idesc, 1u);
                    }
                }
                umma_commit(&bar_empty[stage]);
                umma_commit(&bar_tfull[acc]);
                if (p.stages == 2) { stage ^= 1; if (stage == 0) phase ^= 1; } else { phase ^= 1; }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else {
        // 8 epilogue warps: warp = (TMEM lane group, output parity); thread = one position of the tile
        const int g = warp & 3;
        const int par = (warp - 2) >> 2;
        const int i = g * 32 + lane;
        int acc = 0;
        uint32_t acc_phase = 0;
        float bias[32];
#pragma unroll
        for (int co = 0; co < 32; ++co) bias[co] = (p.bias != nullptr && co < p.Cout) ? __ldg(p.bias + co) : 0.f;
        const long chan = (long)p.H * p.W;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            const int b = tile / p.tiles_per_img, s = (tile - b * p.tiles_per_img) * 128 + i;
            const int q = s / p.WP, xx = s - q * p.WP;
            const int yy = 2 * q + par;
            const bool valid = xx < p.W && yy < p.H;
            mbar_wait(&bar_tfull[acc], acc_phase);
            tc_fence_after();
            uint32_t r[32];
            __syncwarp();
            tmem_ld32(tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)((acc * 2 + par) * HTC_C), r);
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[acc]);         // accumulator is in registers: the MMA warp may reuse it
            float v[32], v2[32];
            float* out = p.y + ((long)b * p.Cout * p.H + yy) * p.W + xx;
#pragma unroll
            for (int co = 0; co < 32; ++co) {
                float val = 0.f;
                if (co < p.Cout && valid) {
                    val = __uint_as_float(r[co]) + bias[co];
                    out[co * chan] = val;
                }
                v[co] = val;
                v2[co] = val * val;
            }
            if (p.stats != nullptr) {
                const float sg = gn_warp_colsum32(v, lane), sq = gn_warp_colsum32(v2, lane);
                if (lane < p.Cout) {
                    atomicAdd(&s_stat[lane], (double)sg);
                    atomicAdd(&s_stat[HTC_C + lane], (double)sq);
                }
            }
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.stats != nullptr && threadIdx.x < p.Cout) {
        atomicAdd(p.stats + threadIdx.x, s_stat[threadIdx.x]);
        atomicAdd(p.stats + p.Cout + threadIdx.x, s_stat[HTC_C + threadIdx.x]);
    }
    if (warp == 1) tmem_dealloc<128>(tmem_base);
}

// ---------------------------------------------------------------------------------------------- C-ABI
static inline long htc_planes_bytes(int B, int H, int W) { return (long)B * 2 * ((H + 1) / 2) * (W + 2) * 128; }

// bytes of caller-owned scratch for gn_hexconv_fwd_tc (plane image of the input + packed weights)
GN_API long gn_hexconv_tc_workspace_bytes(int B, int H, int W) { return htc_planes_bytes(B, H, W) + HTC_W_BYTES + 1024; }

// 1 if the tensor-core path supports this shape
GN_API int gn_hexconv_tc_supported(int cin, int cout, int H, int W, int ksize) {
    return ksize == 1 && cin >= 1 && cin <= HTC_C && cout >= 1 && cout <= HTC_C && H >= 2 && W >= 1 && W + 2 <= 124;
}

// y = hexconv(x') + bias with x' = in_scale ? relu(x*in_scale+in_shift) : x  (same contract as gn_hexconv_fwd; wp from
// gn_hexconv_pack, mode 0 for the forward, mode 1 for the data gradient), kernel_size 1, channels <= 32.
GN_API int gn_hexconv_fwd_tc(const float* x, const float* wp, const float* bias, const float* in_scale, const float* in_shift, float* y,
                             double* stats, int B, int cin, int cout, int H, int W, void* workspace, cudaStream_t stream) {
    GN_REQUIRE(x && wp && y && workspace && B > 0, GN_EINVAL, "hexconv_fwd_tc: bad arguments");
    GN_REQUIRE(gn_hexconv_tc_supported(cin, cout, H, W, 1), GN_EUNSUPPORTED, "hexconv_fwd_tc: needs kernel_size 1, <= 32 channels, W <= 122");
    GN_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), GN_EINVAL, "hexconv_fwd_tc: in_scale/in_shift must come together");
    GN_REQUIRE(B <= 32767 && ((uintptr_t)workspace & 1023) == 0, GN_EALIGN, "hexconv_fwd_tc: workspace must be 1024-byte aligned, batch <= 32767");
    HexTcParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = H; p.W = W; p.Cout = cout;
    p.Q = (H + 1) / 2; p.WP = W + 2; p.PS = p.Q * p.WP;
    p.tiles_per_img = gn_ceil_div(p.PS, 128);
    p.n_tiles = B * p.tiles_per_img;
    p.win_rows = 128 + p.WP + 2;
    p.bias = bias; p.y = y; p.stats = stats;
    // tap tables from the same HexTaps convention as hexconv.cu: tap t has (dy, dx_even, dx_odd)
    //   k = 1: t0..t2 same row dx = -1, 0, +1; t3,t4 row y-1; t5,t6 row y+1 with dx_even in {-1, 0}, dx_odd in {0, +1}
    static const int dy[HTC_TAPS] = {0, 0, 0, -1, -1, 1, 1};
    static const int dxe[HTC_TAPS] = {-1, 0, 1, -1, 0, -1, 0};
    static const int dxo[HTC_TAPS] = {-1, 0, 1, 0, 1, 0, 1};
    for (int t = 0; t < HTC_TAPS; ++t) {
        // even-row output at plane position s: same row -> even plane s+dx; row y-1 -> odd plane (q-1): s-WP+dx; row y+1 -> odd plane q: s+dx
        if (dy[t] == 0) { p.src[0][t] = 0; p.row_off[0][t] = dxe[t] + 1; }
        else { p.src[0][t] = 1; p.row_off[0][t] = (dy[t] < 0 ? -p.WP : 0) + dxe[t] + p.WP + 1; }
        // odd-row output: same row -> odd plane s+dx; row y-1 -> even plane q: s+dx; row y+1 -> even plane q+1: s+WP+dx
        if (dy[t] == 0) { p.src[1][t] = 1; p.row_off[1][t] = dxo[t] + p.WP + 1; }
        else { p.src[1][t] = 0; p.row_off[1][t] = (dy[t] > 0 ? p.WP : 0) + dxo[t] + 1; }
    }
    __nv_bfloat16* planes = (__nv_bfloat16*)workspace;
    __nv_bfloat16* wtc = (__nv_bfloat16*)((uint8_t*)workspace + ((htc_planes_bytes(B, H, W) + 1023) / 1024) * 1024);
    GN_CUDA(gn_launch(hex_pack_tc_kernel, dim3(gn_ceil_div(HTC_TAPS * 2 * HTC_C * 64, 256)), dim3(256), 0, stream, wp, cin, cout, wtc));
    GN_LAUNCH_CHECK();
    dim3 cgrid(2 * p.Q, B);
    GN_CUDA(gn_launch(hex_to_planes_kernel, cgrid, dim3(p.WP <= 96 ? 96 : 128), 0, stream, x, in_scale, in_shift, cin, H, W, p.Q, p.WP, planes));
    GN_LAUNCH_CHECK();

    CUtensorMap tmP, tmW;
    {
        uint64_t dims[3] = {64, (uint64_t)p.PS, (uint64_t)2 * B};
        uint64_t strides[2] = {128, (uint64_t)p.PS * 128};
        uint32_t box[3] = {64, (uint32_t)p.win_rows, 1};
        int rc = gn_tmap_encode(&tmP, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, planes, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    int rc = gn_tmap_bf16_2d(&tmW, wtc, (uint64_t)HTC_TAPS * 2 * HTC_C, 64, 64, 64, HTC_C);
    if (rc) return rc;
    const int win_bytes = ((p.win_rows * 128 + 1023) / 1024) * 1024;
    p.stages = 2;
    const size_t smem = (size_t)HTC_W_BYTES + (size_t)p.stages * 2 * win_bytes + 1024;
    static size_t attr_set = 0;
    if (smem > attr_set) {
        GN_CUDA(cudaFuncSetAttribute(hexconv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = smem;
    }
    const int grid = p.n_tiles < gn_num_sms() ? p.n_tiles : gn_num_sms();
    GN_CUDA(gn_launch(hexconv_tc_kernel, dim3(grid), dim3(320), smem, stream, tmP, tmW, p));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// ================================================================================================
// Weight gradient on tensor cores:  dWp[t][ci][co] += sum_s X'[s + off_t, ci] * dY[s, co]  over both parity planes.
// Both operands are "MN-major" (rows = positions = the reduction index): A = the X' window started at the tap's row
// (64 columns = hi | lo of ci, the second 64-column group aliased: LBO = 0), B = the dY tile (64 columns = hi | lo of co).
// All four cross terms hi/lo x hi/lo land in the 64 x 64 accumulator of the tap; the seven accumulators (448 TMEM
// columns) live for the whole persistent CTA and are folded and added to global memory once at the end.
struct HexTcWgParams {
    int B, H, W, Cin, Cout, Q, WP, PS, tiles_per_img, n_tiles, win_rows;
    float* dwp;                  // [7][Cin][Cout] fp32, +=
    int row_off[2][HTC_TAPS];
    int src[2][HTC_TAPS];
};

__global__ void __launch_bounds__(192, 1)
hexconv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const HexTcWgParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_done;
    __shared__ uint32_t tmem_slot;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int win_bytes = ((p.win_rows * 128 + 1023) / 1024) * 1024;
    const int stage_bytes = 2 * win_bytes + 2 * 16384;            // X' windows (even, odd) + dY tiles (even, odd)

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmDY);
        for (int s = 0; s < 2; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(&tmem_slot);
    gn_pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    gn_pdl_trigger();
    const uint32_t tmem_base = tmem_slot;
    const bool has_work = (int)blockIdx.x < p.n_tiles;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const int b = tile / p.tiles_per_img, s0 = (tile - b * p.tiles_per_img) * 128;
                mbar_wait(&bar_empty[stage], phase ^ 1);
                uint8_t* st = sm + (size_t)stage * stage_bytes;
                mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)(2 * p.win_rows * 128 + 2 * 16384));
                tma_load_3d(&tmX, &bar_full[stage], st, 0, s0 - 1, 2 * b);
                tma_load_3d(&tmX, &bar_full[stage], st + win_bytes, 0, s0 - p.WP - 1, 2 * b + 1);
                tma_load_3d(&tmDY, &bar_full[stage], st + 2 * win_bytes, 0, s0, 2 * b);
                tma_load_3d(&tmDY, &bar_full[stage], st + 2 * win_bytes + 16384, 0, s0, 2 * b + 1);
                stage ^= 1;
                if (stage == 0) phase ^= 1;
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_bf16(128, 64, 1, 1);
            const uint64_t tmplA = smem_desc_template(0, 1024, LAYOUT_SW128);      // MN-major, the two 64-wide groups aliased
            const uint64_t tmplB = smem_desc_template(0, 1024, LAYOUT_SW128);
            int stage = 0;
            uint32_t phase = 0;
            uint32_t started = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                mbar_wait(&bar_full[stage], phase);
                tc_fence_after();
                const uint32_t st = smem_u32(sm + (size_t)stage * stage_bytes);
                const uint64_t descX[2] = {smem_desc(tmplA, st), smem_desc(tmplA, st + win_bytes)};
#pragma unroll
                for (int par = 0; par < 2; ++par) {
                    const uint64_t descDY = smem_desc(tmplB, st + 2 * win_bytes + par * 16384);
#pragma unroll
                    for (int t = 0; t < HTC_TAPS; ++t) {
                        const uint64_t da = descX[p.src[par][t]] + (uint64_t)(p.row_off[par][t] * 8);
#pragma unroll
                        for (int k = 0; k < 8; ++k) {        // 128 positions = 8 x 16 reduction rows (2048 B each)
                            umma_bf16(tmem_base + (uint32_t)(t * 64), da + (uint64_t)(k * 128), descDY + (uint64_t)(k * 128), idesc, (started >> t) & 1u);
                            started |= 1u << t;
                        }
                    }
                }
                umma_commit(&bar_empty[stage]);
                stage ^= 1;
                if (stage == 0) phase ^= 1;
            }
            umma_commit(&bar_done);
        }
    } else if (has_work) {
        const int g = warp & 3;
        if (g < 2) {                                   // accumulator rows 0..31 = x_hi[ci], 32..63 = x_lo[ci]
            mbar_wait(&bar_done, 0);
            tc_fence_after();
            const int ci = lane;
            for (int t = 0; t < HTC_TAPS; ++t) {
                uint32_t r0[32], r1[32];
                __syncwarp();
                tmem_ld32(tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(t * 64), r0);          // x dY_hi[co]
                tmem_ld32(tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(t * 64 + 32), r1);     // x dY_lo[co]
                tmem_ld_wait();
                if (ci < p.Cin) {
                    float* o = p.dwp + ((long)t * p.Cin + ci) * p.Cout;
                    for (int co = 0; co < p.Cout; ++co) atomicAdd(o + co, __uint_as_float(r0[co]) + __uint_as_float(r1[co]));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

GN_API long gn_hexconv_tc_wgrad_workspace_bytes(int B, int H, int W) { return 2 * (((htc_planes_bytes(B, H, W) + 1023) / 1024) * 1024) + 1024; }

// dwp[7][cin][cout] += sum dY * x'   (x' = in_scale ? relu(x*in_scale+in_shift) : x); same contract as gn_hexconv_wgrad except
// that the bias gradient is not produced here (it is a plain channel sum of dY: gn_bn_stats).
GN_API int gn_hexconv_wgrad_tc(const float* x, const float* in_scale, const float* in_shift, const float* dy, float* dwp, int B, int cin,
                               int cout, int H, int W, void* workspace, cudaStream_t stream) {
    GN_REQUIRE(x && dy && dwp && workspace && B > 0, GN_EINVAL, "hexconv_wgrad_tc: bad arguments");
    GN_REQUIRE(gn_hexconv_tc_supported(cin, cout, H, W, 1), GN_EUNSUPPORTED, "hexconv_wgrad_tc: needs kernel_size 1, <= 32 channels, W <= 122");
    GN_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), GN_EINVAL, "hexconv_wgrad_tc: in_scale/in_shift must come together");
    GN_REQUIRE(B <= 32767 && ((uintptr_t)workspace & 1023) == 0, GN_EALIGN, "hexconv_wgrad_tc: workspace must be 1024-byte aligned, batch <= 32767");
    HexTcWgParams p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = H; p.W = W; p.Cin = cin; p.Cout = cout;
    p.Q = (H + 1) / 2; p.WP = W + 2; p.PS = p.Q * p.WP;
    p.tiles_per_img = gn_ceil_div(p.PS, 128);
    p.n_tiles = B * p.tiles_per_img;
    p.win_rows = 128 + p.WP + 2;
    p.dwp = dwp;
    static const int dy_[HTC_TAPS] = {0, 0, 0, -1, -1, 1, 1};
    static const int dxe[HTC_TAPS] = {-1, 0, 1, -1, 0, -1, 0};
    static const int dxo[HTC_TAPS] = {-1, 0, 1, 0, 1, 0, 1};
    for (int t = 0; t < HTC_TAPS; ++t) {
        if (dy_[t] == 0) { p.src[0][t] = 0; p.row_off[0][t] = dxe[t] + 1; }
        else { p.src[0][t] = 1; p.row_off[0][t] = (dy_[t] < 0 ? -p.WP : 0) + dxe[t] + p.WP + 1; }
        if (dy_[t] == 0) { p.src[1][t] = 1; p.row_off[1][t] = dxo[t] + p.WP + 1; }
        else { p.src[1][t] = 0; p.row_off[1][t] = (dy_[t] > 0 ? p.WP : 0) + dxo[t] + 1; }
    }
    const long pb = ((htc_planes_bytes(B, H, W) + 1023) / 1024) * 1024;
    __nv_bfloat16* px = (__nv_bfloat16*)workspace;
    __nv_bfloat16* pdy = (__nv_bfloat16*)((uint8_t*)workspace + pb);
    dim3 cgrid(2 * p.Q, B);
    GN_CUDA(gn_launch(hex_to_planes_kernel, cgrid, dim3(p.WP <= 96 ? 96 : 128), 0, stream, x, in_scale, in_shift, cin, H, W, p.Q, p.WP, px));
    GN_LAUNCH_CHECK();
    GN_CUDA(gn_launch(hex_to_planes_kernel, cgrid, dim3(p.WP <= 96 ? 96 : 128), 0, stream, dy, nullptr, nullptr, cout, H, W, p.Q, p.WP, pdy));
    GN_LAUNCH_CHECK();
    CUtensorMap tmX, tmDY;
    {
        uint64_t dims[3] = {64, (uint64_t)p.PS, (uint64_t)2 * B};
        uint64_t strides[2] = {128, (uint64_t)p.PS * 128};
        uint32_t box[3] = {64, (uint32_t)p.win_rows, 1};
        int rc = gn_tmap_encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, px, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        uint32_t box2[3] = {64, 128, 1};
        rc = gn_tmap_encode(&tmDY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, pdy, dims, strides, box2, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    const int win_bytes = ((p.win_rows * 128 + 1023) / 1024) * 1024;
    const size_t smem = (size_t)2 * (2 * win_bytes + 2 * 16384) + 1024;
    static size_t attr_set = 0;
    if (smem > attr_set) {
        GN_CUDA(cudaFuncSetAttribute(hexconv_tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = smem;
    }
    const int grid = p.n_tiles < gn_num_sms() ? p.n_tiles : gn_num_sms();
    GN_CUDA(gn_launch(hexconv_tc_wgrad_kernel, dim3(grid), dim3(192), smem, stream, tmX, tmDY, p));
    GN_LAUNCH_CHECK();
    return GN_OK;
}
