// 3x3 / pad 1 / stride 1 convolution in NHWC bf16 as a "padded-position" implicit GEMM on tcgen05.
//
// DenseNet's conv2 of every dense layer (/root/reference/gridnext/densenet.py:30-31, 128 -> 32 channels) and its
// data gradient (32 -> 128 with flipped taps).  Index space: every image is addressed with a one-pixel zero border,
// PP = (H+2)(W+2) positions per image, global position P = n*PP + y'*(W+2) + x'.  A tile is 128 CONSECUTIVE
// positions; the input position of tap (ky, kx) is P + (ky-1)(W+2) + (kx-1): the SAME constant row shift for every
// row of the tile.  So the activations of a tile (plus W+3 positions of halo on both sides) are loaded ONCE by TMA as
// whole padded image rows (box = 64 channels x (W+2) x 1 x 1, out-of-bounds coordinates zero-filled = the padding),
// and the nine taps are nine UMMA descriptor start addresses into that one SWIZZLE_128B buffer (the 128B swizzle is
// a function of the absolute shared-memory address, so any 128-byte row offset is legal -- checked on hardware with
// tools/umma_probe).  Weights for all nine taps stay resident in shared memory for the life of the persistent CTA.
// Outputs at border positions are computed and dropped (efficiency HW / PP).
//
//   warp 0: TMA producer   warp 1: MMA issuer (9 taps x kblocks x k16 MMAs of 128 x CO x 16)   warps 2-5: epilogue
#include "gn_common.cuh"
#include "gn_ptx.cuh"
#include "gn_tma.cuh"
#include "gn_epilogue.cuh"

using namespace gnptx;

struct Conv3Params {
    int Nimg, H, W, CI, CO;
    int NP;               // CO rounded up to 16 (UMMA N)
    int kblocks;          // ceil(CI / 64)
    int w_row_bytes;      // bytes per weight row per k-block: 128 (SW128) or 64 (SW64)
    int buf_rows;         // rows (positions) per activation stage buffer
    int stages;
    int n_tiles;
    __nv_bfloat16* out;   // [Nimg*H*W, ldo]
    long ldo;
    int epi_mode;         // 0 plain store, 1 BnBwdEpi
    BnBwdEpi bn;
};

__global__ void __launch_bounds__(192, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const Conv3Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_w, bar_full[2], bar_empty[2], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_slot;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int W2 = p.W + 2, H2 = p.H + 2, PP = W2 * H2, HALO = p.W + 3;
    const int w_tile_bytes = p.NP * p.w_row_bytes;                       // one (tap, k-block) weight tile (rows beyond CO: next tap / zero fill, never stored)
    const int w_bytes = ((9 * p.kblocks * w_tile_bytes + 1023) / 1024) * 1024;
    const int kb_buf_bytes = p.buf_rows * 128;                           // one k-block of one stage
    const int stage_bytes = p.kblocks * kb_buf_bytes;
    uint8_t* s_w = sm;
    uint8_t* s_a = sm + w_bytes;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        mbar_init(&bar_w, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
            mbar_init(&bar_tfull[s], 1);
            mbar_init(&bar_tempty[s], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            // resident weights: 9 * kblocks boxes of {row bytes / 2 channels, CO rows}
            mbar_arrive_expect_tx(&bar_w, 9 * p.kblocks * w_tile_bytes);
            for (int t = 0; t < 9; ++t)
                for (int kb = 0; kb < p.kblocks; ++kb)
                    tma_load_2d(&tmW, &bar_w, s_w + (t * p.kblocks + kb) * w_tile_bytes, kb * 64, t * p.CO);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                mbar_wait(&bar_empty[stage], phase ^ 1);
                const long P0 = (long)tile * 128;
                const long lo = P0 - HALO;
                const long R0 = lo >= 0 ? lo / W2 : -((-lo + W2 - 1) / W2);     // floor division
                const long R1 = (P0 + 127 + HALO) / W2;
                const int nr = (int)(R1 - R0 + 1);
                mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)(nr * W2 * 128 * p.kblocks));
                for (int r = 0; r < nr; ++r) {
                    const long R = R0 + r;
                    int n, yp;
                    if (R >= 0) { n = (int)(R / H2); yp = (int)(R % H2); } else { n = -1; yp = 0; }
                    for (int kb = 0; kb < p.kblocks; ++kb)
                        tma_load_4d(&tmX, &bar_full[stage], s_a + (size_t)stage * stage_bytes + (size_t)kb * kb_buf_bytes + (size_t)r * W2 * 128,
                                    kb * 64, -1, yp - 1, n);
                }
                if (p.stages == 2) { stage ^= 1; if (stage == 0) phase ^= 1; } else { phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = idesc_bf16(128, p.NP, 0, 0);
            const uint64_t tmplA = smem_desc_template(0, 1024, LAYOUT_SW128);
            const uint64_t tmplW = p.w_row_bytes == 128 ? smem_desc_template(0, 1024, LAYOUT_SW128) : smem_desc_template(0, 512, LAYOUT_SW64);
            mbar_wait(&bar_w, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
                mbar_wait(&bar_full[stage], phase);
                tc_fence_after();
                const long P0 = (long)tile * 128;
                const long lo = P0 - HALO;
                const long R0 = lo >= 0 ? lo / W2 : -((-lo + W2 - 1) / W2);
                const int base_row = (int)(P0 - R0 * W2);                 // buffer row of position P0
                const uint32_t d = tmem_base + (uint32_t)(acc * 256);
                const uint32_t a_stage = smem_u32(s_a + (size_t)stage * stage_bytes);
                const uint32_t w_base = smem_u32(s_w);
                uint32_t first = 1;
                for (int t = 0; t < 9; ++t) {
                    const int row = base_row + (t / 3 - 1) * W2 + (t % 3 - 1);
                    for (int kb = 0; kb < p.kblocks; ++kb) {
                        const int kc = min(64, p.CI - kb * 64);
                        const int k16n = (kc + 15) >> 4;
                        const uint32_t a_addr = a_stage + kb * kb_buf_bytes + row * 128;
                        const uint32_t w_addr = w_base + (t * p.kblocks + kb) * w_tile_bytes;
                        for (int k = 0; k < k16n; ++k) {
                            umma_bf16(d, smem_desc(tmplA, a_addr + k * 32), smem_desc(tmplW, w_addr + k * 32), idesc, first ^ 1u);
                            first = 0;
                        }
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
        const int g = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        const int nchunks = (p.CO + 31) / 32;
        float cs_g[8], cs_x[8];       // per-lane column partial sums, columns chunk*32 + lane
#pragma unroll
        for (int i = 0; i < 8; ++i) { cs_g[i] = 0.f; cs_x[i] = 0.f; }
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(&bar_tfull[acc], acc_phase);
            tc_fence_after();
            const long P = (long)tile * 128 + g * 32 + lane;
            const int n = (int)(P / PP);
            const int q = (int)(P % PP);
            const int yp = q / W2, xp = q % W2;
            const bool valid = n < p.Nimg && yp >= 1 && yp <= p.H && xp >= 1 && xp <= p.W;
            const long m = ((long)n * p.H + (yp - 1)) * p.W + (xp - 1);
            const uint32_t taddr = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(acc * 256);
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                if (ch >= nchunks) break;
                uint32_t r[32];
                tmem_ld32(taddr + ch * 32, r);
                tmem_ld_wait();
                const int col = ch * 32;
                const int ncols = min(32, p.CO - col);
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (p.epi_mode == 0) {
                    if (valid) gn_store_bf16_32(p.out + m * p.ldo + col, v, ncols);
                } else {
                    float gx[32];
                    if (valid) {
                        float ref[32];
                        gn_load_bf16_32(p.bn.ref + m * p.bn.ldref + col, ref, ncols);
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (j < ncols) {
                                const float sc = __ldg(p.bn.sc + col + j);
                                const float a = p.bn.ref_is_raw ? fmaf(ref[j], sc, __ldg(p.bn.sh + col + j)) : ref[j];
                                const float gg = a > 0.f ? v[j] : 0.f;
                                gx[j] = gg * (ref[j] - __ldg(p.bn.p0 + col + j)) * __ldg(p.bn.p1 + col + j);
                                v[j] = gg;
                                ref[j] = gg * sc;
                            } else {
                                gx[j] = 0.f; v[j] = 0.f; ref[j] = 0.f;
                            }
                        }
                        gn_store_bf16_32(p.out + m * p.ldo + col, ref, ncols);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { v[j] = 0.f; gx[j] = 0.f; }
                    }
                    if (p.bn.colsum != nullptr) {
                        cs_g[ch] += gn_warp_colsum32(v, lane);
                        cs_x[ch] += gn_warp_colsum32(gx, lane);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (p.epi_mode == 1 && p.bn.colsum != nullptr) {
#pragma unroll
            for (int ch = 0; ch < 8; ++ch) {
                const int col = ch * 32 + lane;
                if (ch < nchunks && col < p.CO) {
                    atomicAdd(p.bn.colsum + col, cs_g[ch]);
                    atomicAdd(p.bn.colsum + p.bn.ldsum + col, cs_x[ch]);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// weight repack:  mode 0 (forward)   wp[(t*CO + co), c] = w[co, c, ky, kx]          rows: 9*CO, cols: CI (pitch ldw)
//                 mode 1 (data grad) wp[(t*CI + c), co] = w[co, c, 2-ky, 2-kx]      rows: 9*CI, cols: CO
__global__ void conv3_pack_kernel(const float* __restrict__ w, int CO, int CI, int mode, __nv_bfloat16* __restrict__ wp, int ldw) {
    const int rows = 9 * (mode == 0 ? CO : CI), cols = mode == 0 ? CI : CO;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < (long)rows * ldw; e += (long)gridDim.x * blockDim.x) {
        const int r = (int)(e / ldw), c = (int)(e % ldw);
        float v = 0.f;
        if (c < cols) {
            const int t = r / (mode == 0 ? CO : CI), o = r % (mode == 0 ? CO : CI);
            const int ky = t / 3, kx = t % 3;
            if (mode == 0) v = w[(((long)o * CI + c) * 3 + ky) * 3 + kx];
            else v = w[(((long)c * CI + o) * 3 + (2 - ky)) * 3 + (2 - kx)];
        }
        wp[e] = __float2bfloat16_rn(v);
    }
}

GN_API int gn_conv3x3_pack(const float* w, int CO, int CI, int mode, void* wp, int ldw, cudaStream_t stream) {
    GN_REQUIRE(w && wp && CO > 0 && CI > 0 && (mode == 0 || mode == 1), GN_EINVAL, "conv3x3_pack: bad arguments");
    const int cols = mode == 0 ? CI : CO;
    GN_REQUIRE(ldw >= cols && ldw % 8 == 0, GN_EALIGN, "conv3x3_pack: ldw must be >= %d and a multiple of 8", cols);
    const long total = 9L * (mode == 0 ? CO : CI) * ldw;
    int blocks = gn_ceil_div(total, 256);
    if (blocks > 1024) blocks = 1024;
    conv3_pack_kernel<<<blocks, 256, 0, stream>>>(w, CO, CI, mode, (__nv_bfloat16*)wp, ldw);
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// x: NHWC bf16 [Nimg, H, W, >=CI] with channel pitch ldx; wp: packed weights [9*CO, ldw] (gn_conv3x3_pack, for the
// data gradient pass CI := conv Cout, CO := conv Cin and the mode-1 pack); out: [Nimg*H*W, ldo] bf16.
// epilogue 1 (bn_* given): out = acc * [a > 0] * bn_sc, column sums into colsum (see gn_epilogue.cuh).
GN_API int gn_conv3x3_bf16(const void* x, long ldx, int Nimg, int H, int W, int CI, const void* wp, int ldw, int CO, void* out, long ldo,
                           const void* bn_ref, long bn_ldref, int bn_ref_is_raw, const float* bn_sc, const float* bn_sh,
                           const float* bn_p0, const float* bn_p1, float* bn_colsum, int bn_ldsum, cudaStream_t stream) {
    GN_REQUIRE(x && wp && out && Nimg > 0 && H > 0 && W > 0 && CI > 0 && CO > 0, GN_EINVAL, "conv3x3: bad arguments");
    GN_REQUIRE(CO % 8 == 0 && CO <= 256, GN_EUNSUPPORTED, "conv3x3: output channels %d must be a multiple of 8, <= 256", CO);
    GN_REQUIRE(CI % 8 == 0 && CI <= 256, GN_EUNSUPPORTED, "conv3x3: input channels %d must be a multiple of 8, <= 256", CI);
    GN_REQUIRE(ldx % 8 == 0 && ldw % 8 == 0 && ldx >= CI && ldw >= CI && ldo >= CO, GN_EALIGN, "conv3x3: bad pitches");
    GN_REQUIRE(W + 2 <= 256, GN_EUNSUPPORTED, "conv3x3: width %d too large", W);
    Conv3Params p;
    memset(&p, 0, sizeof(p));
    p.Nimg = Nimg; p.H = H; p.W = W; p.CI = CI; p.CO = CO; p.NP = ((CO + 15) / 16) * 16;
    p.kblocks = gn_ceil_div(CI, 64);
    p.w_row_bytes = (CI <= 32) ? 64 : 128;
    const int W2 = W + 2, HALO = W + 3;
    const int nr_max = (127 + 2 * HALO) / W2 + 2;
    p.buf_rows = nr_max * W2;
    const long PPt = (long)(H + 2) * W2 * Nimg;
    p.n_tiles = (int)((PPt + 127) / 128);
    p.out = (__nv_bfloat16*)out; p.ldo = ldo;
    p.epi_mode = bn_ref != nullptr ? 1 : 0;
    if (p.epi_mode == 1) {
        GN_REQUIRE(bn_sc && bn_p0 && bn_p1 && (!bn_ref_is_raw || bn_sh), GN_EINVAL, "conv3x3: incomplete BN-backward epilogue arguments");
        p.bn.ref = (const __nv_bfloat16*)bn_ref; p.bn.ldref = bn_ldref; p.bn.ref_is_raw = bn_ref_is_raw;
        p.bn.sc = bn_sc; p.bn.sh = bn_sh; p.bn.p0 = bn_p0; p.bn.p1 = bn_p1; p.bn.colsum = bn_colsum; p.bn.ldsum = bn_ldsum; p.bn.rmw = 0;
    }
    const int w_bytes = ((9 * p.kblocks * p.NP * p.w_row_bytes + 1023) / 1024) * 1024;
    const int stage_bytes = p.kblocks * p.buf_rows * 128;
    const int budget = 227 * 1024 - 1024 - 256;
    GN_REQUIRE(w_bytes + stage_bytes <= budget, GN_EUNSUPPORTED, "conv3x3: tile does not fit shared memory (weights %d B + stage %d B)", w_bytes,
               stage_bytes);
    p.stages = (w_bytes + 2 * stage_bytes <= budget) ? 2 : 1;
    const size_t smem = (size_t)w_bytes + (size_t)p.stages * stage_bytes + 1024;

    CUtensorMap tmX, tmW;
    {
        uint64_t dims[4] = {(uint64_t)CI, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        uint64_t strides[3] = {(uint64_t)ldx * 2, (uint64_t)W * ldx * 2, (uint64_t)H * W * ldx * 2};
        uint32_t box[4] = {64, (uint32_t)W2, 1, 1};
        int rc = gn_tmap_encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)CI, (uint64_t)9 * CO};
        uint64_t strides[1] = {(uint64_t)ldw * 2};
        uint32_t box[2] = {(uint32_t)(p.w_row_bytes / 2), (uint32_t)p.NP};
        int rc = gn_tmap_encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wp, dims, strides, box,
                                p.w_row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
    }
    static int max_set = 0;
    if ((int)smem > max_set) {
        GN_CUDA(cudaFuncSetAttribute(conv3x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        max_set = (int)smem;
    }
    const int grid = p.n_tiles < gn_num_sms() ? p.n_tiles : gn_num_sms();
    conv3x3_kernel<<<grid, 192, smem, stream>>>(tmX, tmW, p);
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// ================================================================================================
// Weight gradient of the 3x3 convolution:  dW[t][c][co] += sum_P  X[P + off_t, c] * dY[P, co]
//
// Same padded-position index space.  The reduction index is the position P, so both operands are "MN-major":
// A = the activation halo buffer (rows = positions, 64 channels per 128-byte row, two channel groups -> UMMA M = 128),
// started at the row of tap t; B = the dY tile in padded layout (its border positions are TMA zero-fill, so halo
// products vanish).  Nine accumulators (one per tap) of 128 x NP fp32 live in TMEM for the whole persistent CTA and are
// added to global memory with vector atomics once at the end.
struct Conv3WgParams {
    int Nimg, H, W, CI, CO, NP;     // NP = CO rounded up to 16 (UMMA N)
    int a_rows, b_rows;             // buffer rows (positions) per stage for X halo and dY
    int n_tiles;
    int stages;
    float* dwp;                     // [9][CI][CO] fp32
};

__global__ void __launch_bounds__(192, 1)
conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDY, const Conv3WgParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[2], bar_empty[2], bar_done;
    __shared__ uint32_t tmem_slot;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W2 = p.W + 2, H2 = p.H + 2, HALO = p.W + 3;
    const int a_group_bytes = p.a_rows * 128;
    const int a_bytes = 2 * a_group_bytes;
    const int b_bytes = p.b_rows * 128;
    const int stage_bytes = a_bytes + b_bytes;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmDY);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
        }
        mbar_init(&bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const bool has_work = (int)blockIdx.x < p.n_tiles;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                mbar_wait(&bar_empty[stage], phase ^ 1);
                uint8_t* sa = sm + (size_t)stage * stage_bytes;
                const long P0 = (long)tile * 128;
                const long lo = P0 - HALO;
                const long R0 = lo >= 0 ? lo / W2 : -((-lo + W2 - 1) / W2);
                const long R1 = (P0 + 127 + HALO) / W2;
                const int nr = (int)(R1 - R0 + 1);
                const long Q0 = P0 / W2, Q1 = (P0 + 127) / W2;
                const int nq = (int)(Q1 - Q0 + 1);
                mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)((2 * nr + nq) * W2 * 128));
                for (int r = 0; r < nr; ++r) {
                    const long R = R0 + r;
                    int n, yp;
                    if (R >= 0) { n = (int)(R / H2); yp = (int)(R % H2); } else { n = -1; yp = 0; }
                    for (int g = 0; g < 2; ++g)
                        tma_load_4d(&tmX, &bar_full[stage], sa + (size_t)g * a_group_bytes + (size_t)r * W2 * 128, g * 64, -1, yp - 1, n);
                }
                for (int r = 0; r < nq; ++r) {
                    const long R = Q0 + r;
                    tma_load_4d(&tmDY, &bar_full[stage], sa + a_bytes + (size_t)r * W2 * 128, 0, -1, (int)(R % H2) - 1, (int)(R / H2));
                }
                if (p.stages == 2) { stage ^= 1; if (stage == 0) phase ^= 1; } else { phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = idesc_bf16(128, p.NP, 1, 1);
            const uint64_t tmplA = smem_desc_template((uint32_t)a_group_bytes, 1024, LAYOUT_SW128);
            const uint64_t tmplB = smem_desc_template(0, 1024, LAYOUT_SW128);
            int stage = 0;
            uint32_t phase = 0;
            bool first_tile = true;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                mbar_wait(&bar_full[stage], phase);
                tc_fence_after();
                const long P0 = (long)tile * 128;
                const long lo = P0 - HALO;
                const long R0 = lo >= 0 ? lo / W2 : -((-lo + W2 - 1) / W2);
                const int a_row0 = (int)(P0 - R0 * W2);
                const int b_row0 = (int)(P0 - (P0 / W2) * W2);
                const uint32_t a_base = smem_u32(sm + (size_t)stage * stage_bytes);
                const uint32_t b_base = a_base + a_bytes + b_row0 * 128;
                for (int t = 0; t < 9; ++t) {
                    const uint32_t a_addr = a_base + (a_row0 + (t / 3 - 1) * W2 + (t % 3 - 1)) * 128;
                    const uint32_t d = tmem_base + (uint32_t)(t * p.NP);
#pragma unroll
                    for (int k = 0; k < 8; ++k)    // 128 positions = 8 x 16 reduction rows
                        umma_bf16(d, smem_desc(tmplA, a_addr + k * 2048), smem_desc(tmplB, b_base + k * 2048), idesc,
                                  (uint32_t)(!first_tile || k != 0));
                }
                umma_commit(&bar_empty[stage]);
                first_tile = false;
                if (p.stages == 2) { stage ^= 1; if (stage == 0) phase ^= 1; } else { phase ^= 1; }
            }
            umma_commit(&bar_done);
        }
    } else if (has_work) {
        const int g = warp & 3;
        mbar_wait(&bar_done, 0);
        tc_fence_after();
        const int c = g * 32 + lane;                 // activation channel (accumulator row)
        for (int t = 0; t < 9; ++t) {
            for (int c0 = 0; c0 < p.NP; c0 += 32) {
                uint32_t r[32];
                tmem_ld32(tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(t * p.NP + c0), r);
                tmem_ld_wait();
                if (c < p.CI) {
                    float* o = p.dwp + ((long)t * p.CI + c) * p.CO + c0;
                    const int ncols = min(32, p.CO - c0);   // columns beyond NP-c0 hold the next tap and are never used
                    if (ncols == 32 && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            atomicAdd(reinterpret_cast<float4*>(o + 4 * q),
                                      make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                                  __uint_as_float(r[4 * q + 3])));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < ncols) atomicAdd(o + j, __uint_as_float(r[j]));
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// dw[co][c][ky][kx] = dwp[t][c][co]
__global__ void conv3_unpack_grad_kernel(const float* __restrict__ dwp, int CO, int CI, float* __restrict__ dw) {
    const long total = 9L * CO * CI;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int t = (int)(e % 9);
        const int c = (int)((e / 9) % CI);
        const int co = (int)(e / (9L * CI));
        dw[e] = dwp[((long)t * CI + c) * CO + co];
    }
}

// x: activations [Nimg*H*W, >=CI] (pitch ldx), dy: output gradient [Nimg*H*W, >=CO] (pitch ldy), dwp: [9][CI][CO] fp32 (+=)
GN_API int gn_conv3x3_wgrad_bf16(const void* x, long ldx, const void* dy, long ldy, int Nimg, int H, int W, int CI, int CO, float* dwp,
                                 cudaStream_t stream) {
    GN_REQUIRE(x && dy && dwp && Nimg > 0 && H > 0 && W > 0 && CI > 0 && CO > 0, GN_EINVAL, "conv3x3_wgrad: bad arguments");
    GN_REQUIRE(CI <= 128 && CI % 8 == 0, GN_EUNSUPPORTED, "conv3x3_wgrad: activation channels %d must be a multiple of 8, <= 128", CI);
    GN_REQUIRE(CO <= 48 && CO % 8 == 0, GN_EUNSUPPORTED, "conv3x3_wgrad: gradient channels %d must be a multiple of 8, <= 48", CO);
    GN_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= CI && ldy >= CO, GN_EALIGN, "conv3x3_wgrad: bad pitches");
    GN_REQUIRE(W + 2 <= 256, GN_EUNSUPPORTED, "conv3x3_wgrad: width %d too large", W);
    Conv3WgParams p;
    p.Nimg = Nimg; p.H = H; p.W = W; p.CI = CI; p.CO = CO; p.NP = ((CO + 15) / 16) * 16;
    const int W2 = W + 2, HALO = W + 3;
    p.a_rows = ((127 + 2 * HALO) / W2 + 2) * W2;
    p.b_rows = (127 / W2 + 2) * W2;
    p.n_tiles = (int)(((long)(H + 2) * W2 * Nimg + 127) / 128);
    p.dwp = dwp;
    const size_t stage_b = (size_t)(2 * p.a_rows + p.b_rows) * 128;
    p.stages = (2 * stage_b + 1024 <= 227 * 1024 - 256) ? 2 : 1;
    const size_t smem = p.stages * stage_b + 1024;
    GN_REQUIRE(smem <= 227 * 1024 - 256, GN_EUNSUPPORTED, "conv3x3_wgrad: tile does not fit shared memory (%zu B)", smem);
    CUtensorMap tmX, tmDY;
    {
        uint64_t dims[4] = {(uint64_t)CI, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        uint64_t strides[3] = {(uint64_t)ldx * 2, (uint64_t)W * ldx * 2, (uint64_t)H * W * ldx * 2};
        uint32_t box[4] = {64, (uint32_t)W2, 1, 1};
        int rc = gn_tmap_encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)CO, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        uint64_t strides[3] = {(uint64_t)ldy * 2, (uint64_t)W * ldy * 2, (uint64_t)H * W * ldy * 2};
        uint32_t box[4] = {64, (uint32_t)W2, 1, 1};
        int rc = gn_tmap_encode(&tmDY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dy, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    static int max_set = 0;
    if ((int)smem > max_set) {
        GN_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        max_set = (int)smem;
    }
    const int grid = p.n_tiles < gn_num_sms() ? p.n_tiles : gn_num_sms();
    conv3x3_wgrad_kernel<<<grid, 192, smem, stream>>>(tmX, tmDY, p);
    GN_LAUNCH_CHECK();
    return GN_OK;
}

GN_API int gn_conv3x3_unpack_grad(const float* dwp, int CO, int CI, float* dw, cudaStream_t stream) {
    GN_REQUIRE(dwp && dw && CO > 0 && CI > 0, GN_EINVAL, "conv3x3_unpack_grad: bad arguments");
    conv3_unpack_grad_kernel<<<gn_ceil_div(9L * CO * CI, 256), 256, 0, stream>>>(dwp, CO, CI, dw);
    GN_LAUNCH_CHECK();
    return GN_OK;
}
