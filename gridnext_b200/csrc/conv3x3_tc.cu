// 3x3 / pad 1 / stride 1 convolution in NHWC bf16 as a "padded-position" implicit GEMM on tcgen05.
//
// DenseNet's conv2 of every dense layer (/root/reference/gridnext/densenet.py:30-31, 128 -> 32 channels) and its
// data gradient (32 -> 128 with flipped taps).  Index space: every image is addressed with a one-pixel zero border,
// (H+2) padded rows of (W+2) positions per image.  A tile is R = floor(128 / (W+2)) CONSECUTIVE PADDED ROWS (they may
// span images); accumulator row i is position i of the tile and the input position of tap (ky, kx) is
// i + (ky-1)(W+2) + (kx-1): the SAME constant row shift for every row of the tile.  So the activations of a tile
// (R + 2 padded rows) are loaded ONCE by TMA as whole padded image rows (box = 64 channels x (W+2) x 1 x 1,
// out-of-bounds coordinates zero-filled = the padding), and the nine taps are nine UMMA descriptor start addresses
// into that one SWIZZLE_128B buffer (the 128B swizzle is a function of the absolute shared-memory address, so any
// 128-byte row offset is legal -- checked on hardware with tools/umma_probe).  Weights for all nine taps stay
// resident in shared memory for the life of the persistent CTA.  The epilogue is staged through shared memory:
// the BN reference tile arrives by TMA (one box per padded row), results are written in place and leave as TMA row
// stores of the W interior positions of every padded row (TMA stores fault on a negative start coordinate --
// measured with tools/tma_store_probe -- so the x' = 0 border is skipped by the source address instead).
//
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
    int R;                // padded image rows per tile (R * (W+2) <= 128 positions)
    int a_rows;           // 128-byte rows per k-block of one activation stage (slack included), multiple of 8
    int stages;
    int n_tiles;
    int nsub;             // 64-channel output sub-tiles per tile
    int e_stages;         // epilogue sub-tile ring depth
    int epi_mode;         // 0 plain store, 1 BnBwdEpi (reference tile fetched by TMA)
    int img;              // > 0: small feature maps -- R = img * (H+2): a tile is img whole padded images, loaded with one TMA box per image
    int stack;            // 1: the three kx taps of a kernel row are stacked along UMMA N (3 MMAs chains of N = 3*CO instead of 9 of N = CO)
    int eager_drain;      // 1: a staging slot is released as soon as ITS store has read it (not one sub-tile later): the slot ring, not any pipe, paces the data gradient
    __nv_bfloat16* out;   // stacked forward: the epilogue stores straight from registers (32 B = one sector per thread and 16-channel half)
    long ldo;
    BnBwdEpi bn;
};

#define C3_SUB_BYTES 16384
#define C3_MAX_ESTAGES 8
#define C3_EPI_WARPS 8
#define C3_MAX_CO 256

__device__ __forceinline__ uint32_t c3_pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 c3_unpack_bf16x2(uint32_t w) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w)); }

// padded row index -> (image, padded y); rows outside [0, Nimg*H2) map to an out-of-bounds image index.
// 32-bit arithmetic only: a 64-bit division per TMA row made the single producer thread the bottleneck.
__device__ __forceinline__ void c3_row_coords(int Rg, int total_rows, int H2, int Nimg, int& n, int& yp) {
    if (Rg < 0) { n = -1; yp = H2 + Rg; }          // only Rg = -1 occurs: continue into image 0 at yp = 0 after one step
    else if (Rg >= total_rows) { n = Nimg; yp = 0; }
    else { n = Rg / H2; yp = Rg - n * H2; }
}
__device__ __forceinline__ void c3_row_next(int H2, int& n, int& yp) {
    if (++yp == H2) { yp = 0; ++n; }
}

// All MMAs of one tile, straight-line when the k-block structure is known at compile time (KB_T k-blocks of K16_T
// 16-channel steps; KB_T = 0: run-time loops).  The single issuing thread must not spend more than ~45 cycles per
// MMA on index arithmetic or it, not the tensor pipe, sets the pace (measured with tools/umma_rate).
template <int KB_T, int K16_T>
__device__ __forceinline__ void c3_issue_tile(uint32_t d, uint64_t descA, uint64_t descW, uint32_t idesc, int W2, int kb_buf16, int w_tile16,
                                              int kblocks, int CI) {
    uint32_t acc = 0;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
        const int row = (t / 3) * W2 + (t % 3) - 1;          // buffer row feeding accumulator row 0 for this tap
        const uint64_t da = descA + (uint64_t)(int64_t)(row * 8);
        if (KB_T > 0) {
            const uint64_t dw = descW + (uint64_t)(t * KB_T * w_tile16);
#pragma unroll
            for (int kb = 0; kb < KB_T; ++kb)
#pragma unroll
                for (int k = 0; k < K16_T; ++k) {
                    umma_bf16(d, da + (uint64_t)(kb * kb_buf16 + k * 2), dw + (uint64_t)(kb * w_tile16 + k * 2), idesc, acc);
                    acc = 1;
                }
        } else {
            const uint64_t dw = descW + (uint64_t)(t * kblocks * w_tile16);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int k16n = (min(64, CI - kb * 64) + 15) >> 4;
                for (int k = 0; k < k16n; ++k) {
                    umma_bf16(d, da + (uint64_t)(kb * kb_buf16 + k * 2), dw + (uint64_t)(kb * w_tile16 + k * 2), idesc, acc);
                    acc = 1;
                }
            }
        }
    }
}

// N-stacked variant for narrow outputs (CO <= 32: the forward conv2, 128 -> 32).  A UMMA with N = 32 costs the same ~45 cycles
// as one with N = 96 (the A-operand fetch from shared memory, 128 B/clk, sets the floor: tools/umma_rate), so the three kx
// taps of one kernel row share ONE MMA: B = the 3*CO packed weight rows (kx, co) of that ky -- contiguous in the packed
// layout -- and A = the tile shifted by (ky-1) padded rows only.  Accumulator column block kx of row q then holds
//   E_kx[q] = sum_ky X[q + (ky-1)(W+2)] . W[ky, kx]     and     out[p] = E_0[p-1] + E_1[p] + E_2[p+1],
// a neighbour-lane sum the epilogue does with two warp shuffles per channel (rows p-1 / p+1 live in adjacent TMEM lanes;
// a tile is whole padded rows, so both neighbours of every interior position are inside the tile).  24 MMAs per tile
// instead of 72.
template <int KB_T, int K16_T>
__device__ __forceinline__ void c3_issue_tile_stacked(uint32_t d, uint64_t descA, uint64_t descW, uint32_t idesc, int W2, int kb_buf16,
                                                      int w_tile16, int kblocks, int CI) {
    uint32_t acc = 0;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const uint64_t da = descA + (uint64_t)(int64_t)(ky * W2 * 8);
        if (KB_T > 0) {
            const uint64_t dw = descW + (uint64_t)(ky * KB_T * w_tile16);
#pragma unroll
            for (int kb = 0; kb < KB_T; ++kb)
#pragma unroll
                for (int k = 0; k < K16_T; ++k) {
                    umma_bf16(d, da + (uint64_t)(kb * kb_buf16 + k * 2), dw + (uint64_t)(kb * w_tile16 + k * 2), idesc, acc);
                    acc = 1;
                }
        } else {
            const uint64_t dw = descW + (uint64_t)(ky * kblocks * w_tile16);
            for (int kb = 0; kb < kblocks; ++kb) {
                const int k16n = (min(64, CI - kb * 64) + 15) >> 4;
                for (int k = 0; k < k16n; ++k) {
                    umma_bf16(d, da + (uint64_t)(kb * kb_buf16 + k * 2), dw + (uint64_t)(kb * w_tile16 + k * 2), idesc, acc);
                    acc = 1;
                }
            }
        }
    }
}

// Tile = R consecutive PADDED image rows (R * (W+2) <= 128 positions, accumulator row i = position i of the tile).
//   warp 0: TMA producer (R+2 padded rows per k-block)      warp 1: MMA issuer (9 taps x kblocks x k16)
//   warp 2: epilogue feeder (TMA loads of the BN reference tile, one box per padded row)
//   warp 3: TMEM allocator + epilogue drain (TMA row stores; border positions fall outside the tensor map and are dropped)
//   warps 4-11: epilogue (tcgen05.ld -> math -> swizzled st.shared in place)
//   EPI_WARPS = 16 (BN-backward epilogue with <= 128 output channels): warp (g, cb) owns accumulator rows 32g..32g+31 and the
//   FIXED 32-column block cb for the whole kernel, so the BN column sums (sum g, sum g*ref) stay in 64 registers per thread
//   across all tiles and are reduced over the 32 rows ONCE at the end.  (With 8 warps each sub-tile paid two 31-step shuffle
//   butterflies -- 57 % of the epilogue's instructions -- and two resident warps per scheduler could not hide their latency:
//   the data-gradient convolution was bound by its epilogue at 2.9 TB/s.)  The two sub-tiles of a tile are processed
//   concurrently by the two halves of the epilogue.
template <int EPI_MODE, int EPI_WARPS>
__global__ void __launch_bounds__((4 + EPI_WARPS) * 32, 1)
conv3x3_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmXb, const __grid_constant__ CUtensorMap tmW,
               const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmRef, const Conv3Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_w, bar_full[2], bar_empty[2], bar_tfull[2], bar_tempty[2];
    __shared__ __align__(8) uint64_t bar_efull[C3_MAX_ESTAGES], bar_eready[C3_MAX_ESTAGES], bar_eempty[C3_MAX_ESTAGES];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_xchg[2][4][2][32];                 // stacked forward: boundary lanes' E_0 / E_2 rows between epilogue warps
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const int W2 = p.W + 2, H2 = p.H + 2, R = p.R;
    const int total_rows = p.Nimg * H2;
    const bool stack = EPI_MODE == 0 && p.stack;
    const int w_groups = stack ? 3 : 9;                                  // weight tiles per k-block: kernel rows (stacked) or taps
    const int w_tile_bytes = (stack ? 3 * p.CO : p.NP) * p.w_row_bytes;  // one (group, k-block) weight tile (rows beyond CO: next tap / zero fill, never stored)
    const int w_bytes = ((w_groups * p.kblocks * w_tile_bytes + 1023) / 1024) * 1024;
    const int kb_buf_bytes = p.a_rows * 128;                             // one k-block of one stage
    const int stage_bytes = p.kblocks * kb_buf_bytes;
    uint8_t* s_w = sm;
    uint8_t* s_a = sm + w_bytes;
    uint8_t* s_slots = s_a + (size_t)p.stages * stage_bytes;
    const int slot_bytes = stack ? C3_SUB_BYTES / 2 : C3_SUB_BYTES;      // stacked forward: 128 positions x 64 B (<= 32 channels, SWIZZLE_64B)
    const int slot_row = stack ? 64 : 128;
    float* s_epi = reinterpret_cast<float*>(s_slots + (size_t)p.e_stages * slot_bytes);         // [4][C3_MAX_CO]

    float* s_cs = s_epi + 4 * C3_MAX_CO;                                                        // [2][C3_MAX_CO] column sums of this CTA
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmXb);
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmOut);
        if (EPI_MODE == 1) tma_prefetch_desc(&tmRef);
        mbar_init(&bar_w, 1);
        for (int s = 0; s < 2; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 1);
            mbar_init(&bar_tfull[s], 1);
            mbar_init(&bar_tempty[s], (EPI_MODE == 0 && EPI_WARPS == 16) ? 8 : EPI_WARPS);
        }
        for (int s = 0; s < C3_MAX_ESTAGES; ++s) {
            mbar_init(&bar_efull[s], 1);
            mbar_init(&bar_eready[s], C3_EPI_WARPS);
            mbar_init(&bar_eempty[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc<512>(&tmem_slot);
    gn_pdl_wait();
    if (EPI_MODE == 1) {
        // the BN-backward epilogue multiplies by what the slot holds even for rows outside the image (their gradient is zero): no NaN bit patterns
        for (int i = threadIdx.x; i < p.e_stages * (slot_bytes / 16); i += blockDim.x) reinterpret_cast<uint4*>(s_slots)[i] = make_uint4(0, 0, 0, 0);
        fence_proxy_async_smem();
        for (int i = threadIdx.x; i < C3_MAX_CO; i += blockDim.x) {
            s_cs[i] = 0.f;
            s_cs[C3_MAX_CO + i] = 0.f;
            const bool in = i < p.CO;
            s_epi[i] = in ? p.bn.sc[i] : 0.f;
            s_epi[C3_MAX_CO + i] = (in && p.bn.sh) ? p.bn.sh[i] : 0.f;
            s_epi[2 * C3_MAX_CO + i] = in ? p.bn.p0[i] : 0.f;
            s_epi[3 * C3_MAX_CO + i] = in ? p.bn.p1[i] : 0.f;
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    gn_pdl_trigger();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            // resident weights: 9 * kblocks boxes of {row bytes / 2 channels, CO rows}
            mbar_arrive_expect_tx(&bar_w, w_groups * p.kblocks * w_tile_bytes);
            for (int t = 0; t < w_groups; ++t)
                for (int kb = 0; kb < p.kblocks; ++kb)
                    tma_load_2d(&tmW, &bar_w, s_w + (t * p.kblocks + kb) * w_tile_bytes, kb * 64, t * (stack ? 3 : 1) * p.CO);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                mbar_wait(&bar_empty[stage], phase ^ 1);
                const int Rg0 = tile * R;
                int n, yp;
                c3_row_coords(Rg0 - 1, total_rows, H2, p.Nimg, n, yp);
                uint8_t* dst = s_a + (size_t)stage * stage_bytes + 1024;
                if (p.img > 0) {
                    // whole images: the halo rows above / below a tile belong to other images and only feed dropped border outputs
                    mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)(R * W2 * 128 * p.kblocks));
                    for (int i = 0; i < p.img; ++i)
                        for (int kb = 0; kb < p.kblocks; ++kb)
                            tma_load_4d(&tmXb, &bar_full[stage], dst + (size_t)kb * kb_buf_bytes + (size_t)(1 + i * H2) * W2 * 128, kb * 64, -1, -1,
                                        tile * p.img + i);
                    if (p.stages == 2) { stage ^= 1; if (stage == 0) phase ^= 1; } else { phase ^= 1; }
                    continue;
                }
                mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)((R + 2) * W2 * 128 * p.kblocks));
                if (n >= 0 && yp + R + 1 < H2) {
                    // all R + 2 padded rows lie in one image: ONE box per k-block (a TMA instruction costs the issuing thread
                    // ~80 cycles; ten per tile kept the two-stage ring from covering the DRAM latency)
                    for (int kb = 0; kb < p.kblocks; ++kb)
                        tma_load_4d(&tmXb, &bar_full[stage], dst + (size_t)kb * kb_buf_bytes, kb * 64, -1, yp - 1, n);
                } else {
                    for (int lr = 0; lr < R + 2; ++lr) {
                        const int nn = n < p.Nimg ? n : p.Nimg;     // past the last image: any out-of-bounds index (zero fill)
                        for (int kb = 0; kb < p.kblocks; ++kb)
                            tma_load_4d(&tmX, &bar_full[stage], dst + (size_t)kb * kb_buf_bytes, kb * 64, -1, yp - 1, nn);
                        dst += W2 * 128;
                        c3_row_next(H2, n, yp);
                    }
                }
                if (p.stages == 2) { stage ^= 1; if (stage == 0) phase ^= 1; } else { phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = idesc_bf16(128, stack ? 3 * p.CO : p.NP, 0, 0);
            const uint64_t tmplA = smem_desc_template(0, 1024, LAYOUT_SW128);
            const uint64_t tmplW = p.w_row_bytes == 128 ? smem_desc_template(0, 1024, LAYOUT_SW128) : smem_desc_template(0, 512, LAYOUT_SW64);
            mbar_wait(&bar_w, 0);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            const uint32_t w_base = smem_u32(s_w);
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
                mbar_wait(&bar_full[stage], phase);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * 256);
                const uint64_t descA = smem_desc(tmplA, smem_u32(s_a + (size_t)stage * stage_bytes) + 1024);   // buffer row 0 = local padded row -1, x' = 0
                const uint64_t descW = smem_desc(tmplW, w_base);
                const int kb_buf16 = kb_buf_bytes >> 4, w_tile16 = w_tile_bytes >> 4;
                if (stack) {
                    if (p.CI == 128) c3_issue_tile_stacked<2, 4>(d, descA, descW, idesc, W2, kb_buf16, w_tile16, p.kblocks, p.CI);
                    else c3_issue_tile_stacked<0, 0>(d, descA, descW, idesc, W2, kb_buf16, w_tile16, p.kblocks, p.CI);
                } else if (p.CI == 128) c3_issue_tile<2, 4>(d, descA, descW, idesc, W2, kb_buf16, w_tile16, p.kblocks, p.CI);
                else if (p.CI == 32) c3_issue_tile<1, 2>(d, descA, descW, idesc, W2, kb_buf16, w_tile16, p.kblocks, p.CI);
                else c3_issue_tile<0, 0>(d, descA, descW, idesc, W2, kb_buf16, w_tile16, p.kblocks, p.CI);
                umma_commit(&bar_empty[stage]);
                umma_commit(&bar_tfull[acc]);
                if (p.stages == 2) { stage ^= 1; if (stage == 0) phase ^= 1; } else { phase ^= 1; }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp == 2) {
        // ---- epilogue feeder (BN-backward epilogue only; the plain epilogue waits for the drain's release itself: one hop less)
        if (EPI_MODE == 1 && elect_one()) {
            int es = 0;
            uint32_t eph = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const int Rg0 = tile * R;
                for (int j = 0; j < p.nsub; ++j) {
                    mbar_wait(&bar_eempty[es], eph ^ 1);
                    uint8_t* slot = s_slots + (size_t)es * slot_bytes;
                    mbar_arrive_expect_tx(&bar_efull[es], (uint32_t)(R * W2 * 128));
                    if (p.img > 0) {
                        for (int i = 0; i < p.img; ++i)
                            tma_load_4d(&tmRef, &bar_efull[es], slot + (size_t)(i * H2) * W2 * 128, j * 64, -1, -1, tile * p.img + i);
                    } else {
                        int n, yp;
                        c3_row_coords(Rg0, total_rows, H2, p.Nimg, n, yp);
                        for (int r = 0; r < R; ++r) {
                            tma_load_4d(&tmRef, &bar_efull[es], slot + (size_t)r * W2 * 128, j * 64, -1, yp - 1, n < p.Nimg ? n : p.Nimg);
                            c3_row_next(H2, n, yp);
                        }
                    }
                    if (++es == p.e_stages) { es = 0; eph ^= 1; }
                }
            }
        }
    } else if (warp == 3) {
        // ---- epilogue drain (the two-group stacked forward stores from registers: nothing to drain)
        if (!(EPI_MODE == 0 && EPI_WARPS == 16) && elect_one()) {
            int es = 0, prev_es = -1;
            uint32_t eph = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                const int Rg0 = tile * R;
                int n0, yp0;
                c3_row_coords(Rg0, total_rows, H2, p.Nimg, n0, yp0);
                for (int j = 0; j < p.nsub; ++j) {
                    mbar_wait(&bar_eready[es], eph);
                    const uint8_t* slot = s_slots + (size_t)es * slot_bytes;
                    int n = n0, yp = yp0;
                    for (int r = 0; r < R; ++r) {
                        if (n < p.Nimg && yp >= 1 && yp <= p.H)
                            tma_store_4d(&tmOut, slot + ((size_t)r * W2 + 1 + (stack ? 1 : 0)) * slot_row, j * 64, 0, yp - 1, n);   // skip the x' = 0 border position
                        c3_row_next(H2, n, yp);
                    }
                    tma_store_commit();
                    if (p.eager_drain) {
                        tma_store_wait_read<0>();
                        mbar_arrive(&bar_eempty[es]);
                    } else {
                        if (prev_es >= 0) {
                            tma_store_wait_read<1>();
                            mbar_arrive(&bar_eempty[prev_es]);
                        }
                        prev_es = es;
                    }
                    if (++es == p.e_stages) { es = 0; eph ^= 1; }
                }
            }
            tma_store_wait_all<0>();
        }
        __syncwarp();
    } else if (EPI_MODE == 0 && EPI_WARPS == 16) {
        // ---- stacked forward, two epilogue groups of 8 warps: group gi owns TMEM accumulator buffer gi, i.e. every other tile of this
        // CTA, so two tiles are in the epilogue at once (one group of 8 in-order warps needed ~2,200 cycles per tile for ~200 dependent
        // instructions and set the pace of the whole kernel).  Results leave as 32-byte register stores: no staging slot, no drain hop.
        // warp (g, h) of a group: accumulator rows 32g..32g+31, channels 16h..16h+15
        const int gi = (warp - 4) >> 3;
        const int g = warp & 3;
        const int h = ((warp - 4) >> 2) & 1;
        const int trow = g * 32 + lane;
        const int r_loc = trow / W2, xp = trow % W2;
        uint32_t acc_phase = 0;
        int it = 0;
        const uint32_t taddr = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(gi * 256);
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            if ((it & 1) != gi) continue;
            int n, yp;
            c3_row_coords(tile * R + r_loc, total_rows, H2, p.Nimg, n, yp);
            const bool valid = r_loc < R && n >= 0 && n < p.Nimg && yp >= 1 && yp <= p.H && xp >= 1 && xp <= p.W;
            mbar_wait(&bar_tfull[gi], acc_phase);
            acc_phase ^= 1;
            tc_fence_after();
            uint32_t r0[16], r1[16], r2[16];
            tmem_ld16(taddr + 16 * h, r0);
            tmem_ld16(taddr + p.CO + 16 * h, r1);
            tmem_ld16(taddr + 2 * p.CO + 16 * h, r2);
            tmem_ld_wait();
            tc_fence_before();
            float4* xw = reinterpret_cast<float4*>(&s_xchg[gi][g][0][16 * h]);
            if (lane == 31) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    xw[q] = make_float4(__uint_as_float(r0[4 * q]), __uint_as_float(r0[4 * q + 1]), __uint_as_float(r0[4 * q + 2]),
                                        __uint_as_float(r0[4 * q + 3]));
            }
            if (lane == 0) {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    xw[8 + q] = make_float4(__uint_as_float(r2[4 * q]), __uint_as_float(r2[4 * q + 1]), __uint_as_float(r2[4 * q + 2]),
                                            __uint_as_float(r2[4 * q + 3]));
            }
            named_bar_sync(2 + gi, 8 * 32);           // the accumulator is in registers: the MMA warp may refill the buffer
            if (lane == 0) mbar_arrive(&bar_tempty[gi]);
            float v[16];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(r0[e]), 1);
                const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[e]), 1);
                v[e] = __uint_as_float(r1[e]) + (lane > 0 ? up : 0.f) + (lane < 31 ? dn : 0.f);
            }
            if (lane == 0 && g > 0) {
                const float4* xr = reinterpret_cast<const float4*>(&s_xchg[gi][g - 1][0][16 * h]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = xr[q];
                    v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
                }
            }
            if (lane == 31 && g < 3) {
                const float4* xr = reinterpret_cast<const float4*>(&s_xchg[gi][g + 1][1][16 * h]);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t = xr[q];
                    v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
                }
            }
            if (valid) {
                __nv_bfloat16* o = p.out + (((long)n * p.H + (yp - 1)) * p.W + (xp - 1)) * p.ldo + 16 * h;
#pragma unroll
                for (int q = 0; q < 2; ++q)
                    if (16 * h + 8 * q < p.CO)
                        *reinterpret_cast<uint4*>(o + 8 * q) = make_uint4(c3_pack_bf16x2(v[8 * q], v[8 * q + 1]), c3_pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                                                          c3_pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), c3_pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
            }
            named_bar_sync(2 + gi, 8 * 32);           // the exchange rows are read: the group's next tile may overwrite them
        }
    } else if (EPI_WARPS == 16) {
        // ---- epilogue, register-accumulating form (see the kernel comment): warp (g, cb), sub-tile j = cb / 2, half h = cb % 2
        const int g = warp & 3;
        const int cb = (warp - 4) >> 2;
        const int j = cb >> 1, h = cb & 1;
        const int trow = g * 32 + lane;
        const uint32_t sw = (uint32_t)(trow & 7);
        const int r_loc = trow / W2, xp = trow % W2;
        const bool active = j < p.nsub;
        const bool is_raw = p.bn.ref_is_raw != 0;
        // column sums as packed fp32 pairs: FADD2 / FFMA2 / FMUL2 do two columns per instruction (this epilogue is bound by instruction issue)
        uint64_t sg2[16], sx2[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) sg2[e] = sx2[e] = 0ull;
        int acc = 0;
        uint32_t acc_phase = 0;
        int es = j % p.e_stages;
        uint32_t eph = (uint32_t)((j / p.e_stages) & 1);
        const float* cst = s_epi + cb * 32;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(&bar_tfull[acc], acc_phase);
            tc_fence_after();
            if (active) {
                int n, yp;
                c3_row_coords(tile * R + r_loc, total_rows, H2, p.Nimg, n, yp);
                const bool valid = r_loc < R && n >= 0 && n < p.Nimg && yp >= 1 && yp <= p.H && xp >= 1 && xp <= p.W;
                const uint32_t taddr = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(acc * 256 + cb * 32);
                mbar_wait(&bar_efull[es], eph);
                uint8_t* rowp = s_slots + (size_t)es * slot_bytes + trow * 128;
#pragma unroll
                for (int ch = 0; ch < 8; ++ch) {                               // eight 4-column chunks (register budget: 96 per thread, 64 of them sums)
                    uint32_t r[4];
                    tmem_ld4(taddr + 4 * ch, r);
                    const uint32_t off = ((((uint32_t)(h * 4 + (ch >> 1))) ^ sw) << 4) + (uint32_t)((ch & 1) * 8);
                    const uint2 rv = *reinterpret_cast<const uint2*>(rowp + off);
                    const uint32_t rw[2] = {rv.x, rv.y};
                    tmem_ld_wait();
                    uint32_t res[2];
#pragma unroll
                    for (int e2 = 0; e2 < 2; ++e2) {
                        // rows outside the image keep whatever the slot held (finite: the slots are cleared once at kernel start and only
                        // ever hold TMA-loaded activations or results), their gradient is forced to zero by `valid`
                        const uint64_t ref2 = bf16x2_to_f32x2(rw[e2]);
                        const float2 sc2f = *reinterpret_cast<const float2*>(cst + 4 * ch + 2 * e2);
                        const uint64_t sc2 = f32x2(sc2f.x, sc2f.y);
                        float a_lo, a_hi;
                        if (is_raw) {
                            const float2 sh2f = *reinterpret_cast<const float2*>(cst + C3_MAX_CO + 4 * ch + 2 * e2);
                            f32x2_unpack(ffma2(ref2, sc2, f32x2(sh2f.x, sh2f.y)), a_lo, a_hi);
                        } else {
                            f32x2_unpack(ref2, a_lo, a_hi);
                        }
                        const float g_lo = (valid && a_lo > 0.f) ? __uint_as_float(r[2 * e2]) : 0.f;
                        const float g_hi = (valid && a_hi > 0.f) ? __uint_as_float(r[2 * e2 + 1]) : 0.f;
                        const uint64_t g2 = f32x2(g_lo, g_hi);
                        const int e = 2 * ch + e2;
                        asm("add.rn.f32x2 %0, %0, %1;" : "+l"(sg2[e]) : "l"(g2));
                        sx2[e] = ffma2(g2, ref2, sx2[e]);                      // sum g*xhat = p1 * (sum g*ref - p0 * sum g), applied at the flush
                        res[e2] = f32x2_to_bf16x2(fmul2(g2, sc2));
                    }
                    *reinterpret_cast<uint2*>(rowp + off) = make_uint2(res[0], res[1]);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_eready[es]);
                es += p.nsub;
                while (es >= p.e_stages) { es -= p.e_stages; eph ^= 1; }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (active && p.bn.colsum != nullptr) {
            float sg[32], sx[32];
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                f32x2_unpack(sg2[e], sg[2 * e], sg[2 * e + 1]);
                f32x2_unpack(sx2[e], sx[2 * e], sx[2 * e + 1]);
            }
            const float tg = gn_warp_colsum32(sg, lane), tx = gn_warp_colsum32(sx, lane);
            const int col = cb * 32 + lane;
            if (col < p.CO) {
                atomicAdd(p.bn.colsum + col, tg);
                atomicAdd(p.bn.colsum + p.bn.ldsum + col, __ldg(p.bn.p1 + col) * (tx - __ldg(p.bn.p0 + col) * tg));
            }
        }
    } else {
        // ---- epilogue: thread = (accumulator row, 32-channel half of each 64-channel sub-tile)
        const int g = warp & 3;
        const int h = (warp - 4) >> 2;
        const int trow = g * 32 + lane;
        const uint32_t sw = (uint32_t)(trow & 7);
        const int r_loc = trow / W2, xp = trow % W2;
        int acc = 0;
        uint32_t acc_phase = 0;
        int es = 0;
        uint32_t eph = 0;
        const bool want_sums = EPI_MODE == 1 && p.bn.colsum != nullptr;
        const bool is_raw = p.bn.ref_is_raw != 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            mbar_wait(&bar_tfull[acc], acc_phase);
            tc_fence_after();
            int n, yp;
            c3_row_coords(tile * R + r_loc, total_rows, H2, p.Nimg, n, yp);
            const bool valid = r_loc < R && n >= 0 && n < p.Nimg && yp >= 1 && yp <= p.H && xp >= 1 && xp <= p.W;
            const uint32_t taddr = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(acc * 256);
#pragma unroll 1
            for (int j = 0; j < p.nsub; ++j) {
                if (EPI_MODE == 1) mbar_wait(&bar_efull[es], eph);
                else mbar_wait(&bar_eempty[es], eph ^ 1);
                // stacked forward: 64-byte rows, stored one row down so that the TMA store source (position r*W2 + 1, W even) stays 128-byte aligned
                uint8_t* rowp = s_slots + (size_t)es * slot_bytes + (trow + (stack ? 1 : 0)) * slot_row;
                const int c0 = j * 64 + h * 32;
                __syncwarp();
                if (EPI_MODE == 0 && stack) {
                    // out[p] = E_0[p-1] + E_1[p] + E_2[p+1]: the neighbours' rows come from the adjacent lanes (and, at the warp
                    // boundaries, from the adjacent epilogue warp through s_xchg; one named barrier per tile, double-buffered)
                    // warp (g, h): rows 32g..32g+31, channels 16h..16h+15
                    uint32_t r0[16], r1[16], r2[16];
                    tmem_ld16(taddr + 16 * h, r0);
                    tmem_ld16(taddr + p.CO + 16 * h, r1);
                    tmem_ld16(taddr + 2 * p.CO + 16 * h, r2);
                    tmem_ld_wait();
                    float4* xw = reinterpret_cast<float4*>(&s_xchg[acc][g][0][16 * h]);
                    if (lane == 31) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            xw[q] = make_float4(__uint_as_float(r0[4 * q]), __uint_as_float(r0[4 * q + 1]), __uint_as_float(r0[4 * q + 2]),
                                                __uint_as_float(r0[4 * q + 3]));
                    }
                    if (lane == 0) {
#pragma unroll
                        for (int q = 0; q < 4; ++q)
                            xw[8 + q] = make_float4(__uint_as_float(r2[4 * q]), __uint_as_float(r2[4 * q + 1]), __uint_as_float(r2[4 * q + 2]),
                                                    __uint_as_float(r2[4 * q + 3]));
                    }
                    named_bar_sync(2, C3_EPI_WARPS * 32);
                    float v[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float up = __shfl_up_sync(0xffffffffu, __uint_as_float(r0[e]), 1);
                        const float dn = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[e]), 1);
                        v[e] = __uint_as_float(r1[e]) + (lane > 0 ? up : 0.f) + (lane < 31 ? dn : 0.f);
                    }
                    if (lane == 0 && g > 0) {
                        const float4* xr = reinterpret_cast<const float4*>(&s_xchg[acc][g - 1][0][16 * h]);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 t = xr[q];
                            v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
                        }
                    }
                    if (lane == 31 && g < 3) {
                        const float4* xr = reinterpret_cast<const float4*>(&s_xchg[acc][g + 1][1][16 * h]);
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 t = xr[q];
                            v[4 * q] += t.x; v[4 * q + 1] += t.y; v[4 * q + 2] += t.z; v[4 * q + 3] += t.w;
                        }
                    }
                    const uint32_t sw64 = (uint32_t)(((trow + 1) >> 1) & 3);   // SWIZZLE_64B: 16-byte chunk index ^= address bits [8:7]
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        if (trow == 127) break;                                // never an interior position; its shifted row would leave the slot
                        const uint32_t off = (((uint32_t)(2 * h + q)) ^ sw64) << 4;
                        *reinterpret_cast<uint4*>(rowp + off) = make_uint4(c3_pack_bf16x2(v[8 * q], v[8 * q + 1]), c3_pack_bf16x2(v[8 * q + 2], v[8 * q + 3]),
                                                                           c3_pack_bf16x2(v[8 * q + 4], v[8 * q + 5]), c3_pack_bf16x2(v[8 * q + 6], v[8 * q + 7]));
                    }
                } else if (c0 < p.NP) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c0, r);
                    tmem_ld_wait();
                    const float* cst = s_epi + c0;
                    float v[32], gx[32];
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const uint32_t off = (((uint32_t)(h * 4 + q)) ^ sw) << 4;
                        uint32_t res[4];
                        if (EPI_MODE == 0) {
#pragma unroll
                            for (int e2 = 0; e2 < 4; ++e2)
                                res[e2] = c3_pack_bf16x2(__uint_as_float(r[8 * q + 2 * e2]), __uint_as_float(r[8 * q + 2 * e2 + 1]));
                        } else {
                            const uint4 rv = *reinterpret_cast<const uint4*>(rowp + off);
                            const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
                            const float4 sc0 = *reinterpret_cast<const float4*>(cst + 8 * q), sc1 = *reinterpret_cast<const float4*>(cst + 8 * q + 4);
                            const float scv[8] = {sc0.x, sc0.y, sc0.z, sc0.w, sc1.x, sc1.y, sc1.z, sc1.w};
                            float shv[8];
                            if (is_raw) {
                                const float4 sh0 = *reinterpret_cast<const float4*>(cst + C3_MAX_CO + 8 * q),
                                             sh1 = *reinterpret_cast<const float4*>(cst + C3_MAX_CO + 8 * q + 4);
                                shv[0] = sh0.x; shv[1] = sh0.y; shv[2] = sh0.z; shv[3] = sh0.w; shv[4] = sh1.x; shv[5] = sh1.y; shv[6] = sh1.z; shv[7] = sh1.w;
                            }
#pragma unroll
                            for (int e2 = 0; e2 < 4; ++e2) {
                                const float2 rf = c3_unpack_bf16x2(rw[e2]);
                                float o2[2];
#pragma unroll
                                for (int u = 0; u < 2; ++u) {
                                    const int e = 8 * q + 2 * e2 + u;
                                    const float ref = u ? rf.y : rf.x;
                                    const float sc = scv[2 * e2 + u];
                                    const float a = is_raw ? fmaf(ref, sc, shv[2 * e2 + u]) : ref;
                                    const bool on = valid && a > 0.f;       // dropped rows hold stale shared memory: select, never multiply
                                    const float gg = on ? __uint_as_float(r[e]) : 0.f;
                                    gx[e] = on ? gg * ref : 0.f;            // sum g*xhat = p1 * (sum g*ref - p0 * sum g), applied at the flush
                                    v[e] = gg;
                                    o2[u] = gg * sc;
                                }
                                res[e2] = c3_pack_bf16x2(o2[0], o2[1]);
                            }
                        }
                        *reinterpret_cast<uint4*>(rowp + off) = make_uint4(res[0], res[1], res[2], res[3]);
                    }
                    if (want_sums) {
                        const float sg = gn_warp_colsum32(v, lane), sx = gn_warp_colsum32(gx, lane);
                        atomicAdd(&s_cs[c0 + lane], sg);
                        atomicAdd(&s_cs[C3_MAX_CO + c0 + lane], sx);
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_eready[es]);
                if (++es == p.e_stages) { es = 0; eph ^= 1; }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        if (want_sums) {
            named_bar_sync(1, C3_EPI_WARPS * 32);          // every epilogue warp has added its last partial sums
            for (int col = threadIdx.x - 4 * 32; col < p.CO; col += C3_EPI_WARPS * 32) {
                const float sg = s_cs[col], sx = s_cs[C3_MAX_CO + col];
                atomicAdd(p.bn.colsum + col, sg);
                atomicAdd(p.bn.colsum + p.bn.ldsum + col, __ldg(p.bn.p1 + col) * (sx - __ldg(p.bn.p0 + col) * sg));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) tmem_dealloc<512>(tmem_base);
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
    Conv3Params p;
    memset(&p, 0, sizeof(p));
    p.Nimg = Nimg; p.H = H; p.W = W; p.CI = CI; p.CO = CO; p.NP = ((CO + 15) / 16) * 16;
    p.kblocks = gn_ceil_div(CI, 64);
    p.w_row_bytes = (CI <= 32) ? 64 : 128;
    const int W2 = W + 2;
    GN_REQUIRE(W2 <= 128, GN_EUNSUPPORTED, "conv3x3: width %d too large (W + 2 must fit one 128-position tile)", W);
    p.R = 128 / W2;
    p.img = ((H + 2) * W2 <= 128) ? 128 / ((H + 2) * W2) : 0;
    if (p.img > 0) p.R = p.img * (H + 2);
    p.a_rows = ((2 * W2 + 137 + 7) / 8) * 8;
    const long total_rows = (long)(H + 2) * Nimg;
    p.n_tiles = (int)((total_rows + p.R - 1) / p.R);
    p.nsub = (CO + 63) / 64;
    p.epi_mode = bn_ref != nullptr ? 1 : 0;
    p.stack = (p.epi_mode == 0 && CO <= 32 && (3 * CO) % 16 == 0 && W % 2 == 0) ? 1 : 0;
    p.out = (__nv_bfloat16*)out; p.ldo = ldo;
    p.eager_drain = gn_env_flag("GN_C3_LAZY_DRAIN") ? 0 : 1;
    GN_REQUIRE(((uintptr_t)out & 15) == 0 && ldo % 8 == 0, GN_EALIGN, "conv3x3: output view must be 16-byte aligned with a pitch that is a multiple of 8");
    if (p.epi_mode == 1) {
        GN_REQUIRE(bn_sc && bn_p0 && bn_p1 && (!bn_ref_is_raw || bn_sh), GN_EINVAL, "conv3x3: incomplete BN-backward epilogue arguments");
        GN_REQUIRE(((uintptr_t)bn_ref & 15) == 0 && bn_ldref % 8 == 0 && bn_ldref >= CO, GN_EALIGN, "conv3x3: reference view must be 16-byte aligned with a pitch that is a multiple of 8");
        p.bn.ref = (const __nv_bfloat16*)bn_ref; p.bn.ldref = bn_ldref; p.bn.ref_is_raw = bn_ref_is_raw;
        p.bn.sc = bn_sc; p.bn.sh = bn_sh; p.bn.p0 = bn_p0; p.bn.p1 = bn_p1; p.bn.colsum = bn_colsum; p.bn.ldsum = bn_ldsum; p.bn.rmw = 0;
    }
    const int w_bytes = (((p.stack ? 3 * 3 * CO : 9 * p.NP) * p.kblocks * p.w_row_bytes + 1023) / 1024) * 1024;
    const int stage_bytes = p.kblocks * p.a_rows * 128;
    const int epi_fixed = 6 * C3_MAX_CO * 4;
    const int budget = 227 * 1024 - 1024 - 512 - 2048;      // static shared memory: barriers + the stacked epilogue's exchange rows
    const int slot_bytes = p.stack ? C3_SUB_BYTES / 2 : C3_SUB_BYTES;
    GN_REQUIRE(w_bytes + stage_bytes + 2 * slot_bytes + epi_fixed <= budget, GN_EUNSUPPORTED,
               "conv3x3: tile does not fit shared memory (weights %d B + stage %d B)", w_bytes, stage_bytes);
    p.stages = (w_bytes + 2 * stage_bytes + 2 * slot_bytes + epi_fixed <= budget) ? 2 : 1;
    p.e_stages = (budget - w_bytes - p.stages * stage_bytes - epi_fixed) / slot_bytes;
    if (p.e_stages > C3_MAX_ESTAGES) p.e_stages = C3_MAX_ESTAGES;
    const size_t smem = (size_t)w_bytes + (size_t)p.stages * stage_bytes + (size_t)p.e_stages * slot_bytes + epi_fixed + 1024;

    CUtensorMap tmX, tmXb, tmW, tmOut, tmRef;
    memset(&tmRef, 0, sizeof(tmRef));
    {
        uint64_t dims[4] = {(uint64_t)CI, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        uint64_t strides[3] = {(uint64_t)ldx * 2, (uint64_t)W * ldx * 2, (uint64_t)H * W * ldx * 2};
        uint32_t box[4] = {64, (uint32_t)W2, 1, 1};
        int rc = gn_tmap_encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
        box[2] = p.img > 0 ? (uint32_t)(H + 2) : (uint32_t)(p.R + 2);   // a whole padded image | the whole halo tile of one k-block when it lies inside one image
        rc = gn_tmap_encode(&tmXb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {(uint64_t)CI, (uint64_t)9 * CO};
        uint64_t strides[1] = {(uint64_t)ldw * 2};
        uint32_t box[2] = {(uint32_t)(p.w_row_bytes / 2), (uint32_t)(p.stack ? 3 * CO : p.NP)};
        int rc = gn_tmap_encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wp, dims, strides, box,
                                p.w_row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)CO, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        uint64_t strides[3] = {(uint64_t)ldo * 2, (uint64_t)W * ldo * 2, (uint64_t)H * W * ldo * 2};
        uint32_t box[4] = {p.stack ? 32u : 64u, (uint32_t)W, 1, 1};        // stores may not start at a negative coordinate (tools/tma_store_probe)
        int rc = gn_tmap_encode(&tmOut, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, out, dims, strides, box,
                                p.stack ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    if (p.epi_mode == 1) {
        uint64_t dims[4] = {(uint64_t)CO, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        uint64_t strides[3] = {(uint64_t)bn_ldref * 2, (uint64_t)W * bn_ldref * 2, (uint64_t)H * W * bn_ldref * 2};
        uint32_t box[4] = {64, (uint32_t)W2, p.img > 0 ? (uint32_t)(H + 2) : 1u, 1};
        int rc = gn_tmap_encode(&tmRef, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, bn_ref, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    static int max_set[2] = {0, 0};
    if ((int)smem > max_set[p.epi_mode]) {
        if (p.epi_mode) {
            GN_CUDA(cudaFuncSetAttribute(conv3x3_kernel<1, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            GN_CUDA(cudaFuncSetAttribute(conv3x3_kernel<1, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        } else {
            GN_CUDA(cudaFuncSetAttribute(conv3x3_kernel<0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            GN_CUDA(cudaFuncSetAttribute(conv3x3_kernel<0, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        }
        max_set[p.epi_mode] = (int)smem;
    }
    const int grid = p.n_tiles < gn_num_sms() ? p.n_tiles : gn_num_sms();
    if (p.epi_mode && p.nsub <= 2 && p.e_stages >= 2) GN_CUDA(gn_launch(conv3x3_kernel<1, 16>, dim3(grid), dim3(640), smem, stream, tmX, tmXb, tmW, tmOut, tmRef, p));
    else if (p.epi_mode) GN_CUDA(gn_launch(conv3x3_kernel<1, 8>, dim3(grid), dim3(384), smem, stream, tmX, tmXb, tmW, tmOut, tmRef, p));
    else if (p.stack && !gn_env_flag("GN_C3_FWD_ONE_GROUP")) GN_CUDA(gn_launch(conv3x3_kernel<0, 16>, dim3(grid), dim3(640), smem, stream, tmX, tmXb, tmW, tmOut, tmRef, p));
    else GN_CUDA(gn_launch(conv3x3_kernel<0, 8>, dim3(grid), dim3(384), smem, stream, tmX, tmXb, tmW, tmOut, tmRef, p));
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
    int stack;                      // 1 (CO <= 32): the three kx taps share one MMA, B = dY shifted by +1 / 0 / -1 positions stacked along N
    int img;                        // > 0: small images -- a tile is img WHOLE padded images (img * (H+2) * (W+2) <= 128 positions), one TMA box per
                                    // image and operand: no halo rows from neighbouring images and 3 * img TMA instructions per tile instead of ~47
    int k16;                        // reduction steps of 16 positions per tile
    // chunk mode (CO <= 32): a tile is ONE contiguous run of padded positions fetched with three TMA boxes (two channel groups of X, one dY):
    //   rt > 0: rt interior rows of one image (X box = rt + 2 rows: the halo rows come with it, 1 + 2/rt re-read instead of 1.86x)
    //   rt = 0: ni whole padded images (box spans the image dimension)
    // the position count is the REDUCTION length, so it is not tied to 128: ~300 positions per tile, three kx taps stacked along N
    int chunk, rt, ni, tiles_per_img;
    int x_front;                    // slack positions in front of the X box (rt = 0: tap row ky = 0 reaches W + 2 positions back)
    int xa_rows;                    // 128-byte rows per channel group of one X stage
    int y_bytes;                    // bytes of one dY stage (1024 B of zeros in front: position -1)
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
    const int a_group_bytes = (p.chunk ? p.xa_rows : p.a_rows) * 128;
    const int a_bytes = 2 * a_group_bytes;
    const int b_row_bytes = p.stack ? 64 : 128;      // stacked: dY as [position][32 channels] SWIZZLE_64B, so that a 32-channel group is one MN atom
    const int b_ext = p.stack ? 1 : 0;               // ... and positions P0-1 .. P0+128 are needed
    const int b_bytes = p.chunk ? p.y_bytes : p.b_rows * b_row_bytes;
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
    if (p.img > 0 || p.chunk) {
        // positions between the images' data and the next multiple of 16, and the slack rows the shifted taps reach, are never
        // written by TMA: they must hold zeros (dY) / finite values (X) for the whole kernel
        uint4* z = reinterpret_cast<uint4*>(sm);
        const int n16 = p.stages * stage_bytes / 16;
        for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async_smem();
    }
    if (warp == 1) tmem_alloc<512>(&tmem_slot);
    gn_pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    gn_pdl_trigger();
    const uint32_t tmem_base = tmem_slot;
    const bool has_work = (int)blockIdx.x < p.n_tiles;
    const int img_pos = H2 * W2;                 // positions of one padded image
    const int img_front = W2 + 1;                // slack rows in front of the first image of a tile (tap (0, 0) reaches back W2 + 1 positions)

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                mbar_wait(&bar_empty[stage], phase ^ 1);
                uint8_t* sa = sm + (size_t)stage * stage_bytes;
                if (p.chunk) {
                    const int nimg = p.rt > 0 ? 1 : p.ni;
                    const int xr = p.rt > 0 ? p.rt + 2 : H2, yr = p.rt > 0 ? p.rt : H2;
                    int n, y0;
                    if (p.rt > 0) { n = tile / p.tiles_per_img; y0 = (tile - n * p.tiles_per_img) * p.rt; } else { n = tile * p.ni; y0 = -1; }
                    mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)(nimg * W2 * (2 * xr * 128 + yr * 64)));
                    for (int g = 0; g < 2; ++g)
                        tma_load_4d(&tmX, &bar_full[stage], sa + (size_t)g * a_group_bytes + (size_t)p.x_front * 128, g * 64, -1, p.rt > 0 ? y0 - 1 : -1, n);
                    tma_load_4d(&tmDY, &bar_full[stage], sa + a_bytes + 1024, 0, -1, y0, n);
                    if (p.stages == 2) { stage ^= 1; if (stage == 0) phase ^= 1; } else { phase ^= 1; }
                    continue;
                }
                if (p.img > 0) {
                    mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)(3 * p.img * img_pos * 128));
                    for (int i = 0; i < p.img; ++i) {
                        const int n = tile * p.img + i;           // past the last image: out of bounds = zero fill
                        for (int g = 0; g < 2; ++g)
                            tma_load_4d(&tmX, &bar_full[stage], sa + (size_t)g * a_group_bytes + (size_t)(img_front + i * img_pos) * 128, g * 64, -1, -1, n);
                        tma_load_4d(&tmDY, &bar_full[stage], sa + a_bytes + (size_t)(i * img_pos) * 128, 0, -1, -1, n);
                    }
                    if (p.stages == 2) { stage ^= 1; if (stage == 0) phase ^= 1; } else { phase ^= 1; }
                    continue;
                }
                const int P0 = tile * 128;
                const int lo = P0 - HALO;
                const int R0 = lo >= 0 ? lo / W2 : -((-lo + W2 - 1) / W2);
                const int R1 = (P0 + 127 + HALO) / W2;
                const int nr = (int)(R1 - R0 + 1);
                const int Q0 = P0 - b_ext >= 0 ? (P0 - b_ext) / W2 : -1, Q1 = (P0 + 127 + b_ext) / W2;
                const int nq = (int)(Q1 - Q0 + 1);
                mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)(2 * nr * W2 * 128 + nq * W2 * b_row_bytes));
                for (int r = 0; r < nr; ++r) {
                    const int R = R0 + r;
                    int n, yp;
                    if (R >= 0) { n = R / H2; yp = R - n * H2; } else { n = -1; yp = 0; }
                    for (int g = 0; g < 2; ++g)
                        tma_load_4d(&tmX, &bar_full[stage], sa + (size_t)g * a_group_bytes + (size_t)r * W2 * 128, g * 64, -1, yp - 1, n);
                }
                for (int r = 0; r < nq; ++r) {
                    const int R = Q0 + r;
                    int n, yp;
                    if (R >= 0) { n = R / H2; yp = R - n * H2; } else { n = -1; yp = 0; }
                    tma_load_4d(&tmDY, &bar_full[stage], sa + a_bytes + (size_t)r * W2 * b_row_bytes, 0, -1, yp - 1, n);
                }
                if (p.stages == 2) { stage ^= 1; if (stage == 0) phase ^= 1; } else { phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = idesc_bf16(128, p.NP, 1, 1);
            const uint64_t tmplA = smem_desc_template((uint32_t)a_group_bytes, 1024, LAYOUT_SW128);
            const uint64_t tmplB = smem_desc_template(0, 1024, LAYOUT_SW128);
            const uint32_t idesc_s = idesc_bf16(128, 96, 1, 1);
            const uint64_t tmplBs = smem_desc_template(64, 512, LAYOUT_SW64);
            int stage = 0;
            uint32_t phase = 0;
            bool first_tile = true;
            for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
                mbar_wait(&bar_full[stage], phase);
                tc_fence_after();
                const int P0 = tile * 128;
                const int lo = P0 - HALO;
                const int R0 = lo >= 0 ? lo / W2 : -((-lo + W2 - 1) / W2);
                const int a_row0 = p.img > 0 ? img_front : (int)(P0 - R0 * W2);
                const uint32_t a_base = smem_u32(sm + (size_t)stage * stage_bytes);
                if (p.chunk) {
                    // same stacked contraction as below over the tile's k16 * 16 positions; position -1 is the 64 zero bytes in front of the dY box
                    const uint32_t b_base = a_base + a_bytes + 1024 - 64;
                    const int q0 = p.x_front + (p.rt > 0 ? W2 : 0);          // buffer row of the tile's first position
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const uint32_t a_addr = a_base + (q0 + (ky - 1) * W2) * 128;
                        const uint32_t d = tmem_base + (uint32_t)(ky * 96);
                        uint64_t da = smem_desc(tmplA, a_addr), db = smem_desc(tmplBs, b_base);
                        uint32_t acc = first_tile ? 0u : 1u;
                        for (int k = 0; k < p.k16; ++k) {
                            umma_bf16(d, da, db, idesc_s, acc);
                            da += 128; db += 64; acc = 1u;                  // 2048 B / 1024 B per step in descriptor units of 16 B
                        }
                    }
                } else if (p.img > 0) {
                    const uint32_t b_base = a_base + a_bytes;
                    for (int t = 0; t < 9; ++t) {
                        const uint32_t a_addr = a_base + (a_row0 + (t / 3 - 1) * W2 + (t % 3 - 1)) * 128;
                        const uint32_t d = tmem_base + (uint32_t)(t * p.NP);
                        for (int k = 0; k < p.k16; ++k)
                            umma_bf16(d, smem_desc(tmplA, a_addr + k * 2048), smem_desc(tmplB, b_base + k * 2048), idesc,
                                      (uint32_t)(!first_tile || k != 0));
                    }
                } else if (p.stack) {
                    // D_ky[ci, (j, co)] += sum_q X[q + (ky-1)W2][ci] * dY[q - 1 + j][co],  j = 2 - kx: the three N groups are the SAME dY
                    // buffer starting one position (64 B = LBO) later each -- overlapping MN atoms, legal because the swizzle is a
                    // function of the absolute shared-memory address
                    const int Q0 = P0 - 1 >= 0 ? (P0 - 1) / W2 : -1;
                    const uint32_t b_base = a_base + a_bytes + (uint32_t)(P0 - 1 - Q0 * W2) * 64;
#pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const uint32_t a_addr = a_base + (a_row0 + (ky - 1) * W2) * 128;
                        const uint32_t d = tmem_base + (uint32_t)(ky * 96);
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            umma_bf16(d, smem_desc(tmplA, a_addr + k * 2048), smem_desc(tmplBs, b_base + k * 1024), idesc_s,
                                      (uint32_t)(!first_tile || k != 0));
                    }
                } else {
                    const int b_row0 = (int)(P0 - (P0 / W2) * W2);
                    const uint32_t b_base = a_base + a_bytes + b_row0 * 128;
                    for (int t = 0; t < 9; ++t) {
                        const uint32_t a_addr = a_base + (a_row0 + (t / 3 - 1) * W2 + (t % 3 - 1)) * 128;
                        const uint32_t d = tmem_base + (uint32_t)(t * p.NP);
#pragma unroll
                        for (int k = 0; k < 8; ++k)    // 128 positions = 8 x 16 reduction rows
                            umma_bf16(d, smem_desc(tmplA, a_addr + k * 2048), smem_desc(tmplB, b_base + k * 2048), idesc,
                                      (uint32_t)(!first_tile || k != 0));
                    }
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
                const uint32_t col = p.stack ? (uint32_t)((t / 3) * 96 + (2 - t % 3) * 32) : (uint32_t)(t * p.NP + c0);
                tmem_ld32(tmem_base + ((uint32_t)(g * 32) << 16) + col, r);
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
    memset(&p, 0, sizeof(p));
    p.Nimg = Nimg; p.H = H; p.W = W; p.CI = CI; p.CO = CO; p.NP = ((CO + 15) / 16) * 16;
    const int W2 = W + 2, H2 = H + 2, HALO = W + 3;
    p.dwp = dwp;
    size_t smem = 0;
    static int chunk_npos = -1;         // target positions per tile (GN_WG_NPOS = 0 switches the chunk mode off: development A/B)
    if (chunk_npos < 0) {
        const char* e = getenv("GN_WG_NPOS");
        chunk_npos = e ? atoi(e) : 304;
    }
    if (CO <= 32 && chunk_npos > 0 && W2 <= 256 && W2 <= chunk_npos) {
        const int img_pos = H2 * W2;
        int npos;
        if (2 * img_pos <= chunk_npos) {                       // several whole padded images per tile
            p.ni = chunk_npos / img_pos;
            if (p.ni > Nimg) p.ni = Nimg;
            if (p.ni > 256) p.ni = 256;
            p.rt = 0;
            npos = p.ni * img_pos;
            p.x_front = ((W2 + 7) / 8) * 8;
            p.n_tiles = (Nimg + p.ni - 1) / p.ni;
            p.tiles_per_img = 0;
        } else {                                               // rt interior rows of one image, split evenly
            int rmax = chunk_npos / W2;
            if (rmax > H) rmax = H;
            if (rmax > 254) rmax = 254;
            p.tiles_per_img = (H + rmax - 1) / rmax;
            p.rt = (H + p.tiles_per_img - 1) / p.tiles_per_img;
            p.ni = 1;
            npos = p.rt * W2;
            p.x_front = 0;
            p.n_tiles = Nimg * p.tiles_per_img;
        }
        p.k16 = (npos + 15) / 16;
        p.xa_rows = ((p.x_front + (p.rt > 0 ? W2 : 0) + W2 + 16 * p.k16 + 7) / 8) * 8;
        const int x_box_rows = p.rt > 0 ? (p.rt + 2) * W2 : npos;
        if (p.xa_rows < ((p.x_front + x_box_rows + 7) / 8) * 8) p.xa_rows = ((p.x_front + x_box_rows + 7) / 8) * 8;
        p.y_bytes = 1024 + (((16 * p.k16 + 2) * 64 + 1023) / 1024) * 1024;
        const size_t stage_b = (size_t)2 * p.xa_rows * 128 + (size_t)p.y_bytes;
        if (stage_b + 1024 <= 227 * 1024 - 256) {
            p.chunk = 1;
            p.stack = 1;
            p.stages = (2 * stage_b + 1024 <= 227 * 1024 - 256) ? 2 : 1;
            smem = p.stages * stage_b + 1024;
        }
    }
    if (!p.chunk) {
        p.rt = p.ni = p.tiles_per_img = p.x_front = p.xa_rows = p.y_bytes = 0;
        p.a_rows = ((127 + 2 * HALO) / W2 + 2) * W2;
        p.img = (H + 2) * (W + 2) <= 128 ? 128 / ((H + 2) * (W + 2)) : 0;
        p.k16 = p.img > 0 ? (p.img * (H + 2) * (W + 2) + 15) / 16 : 8;
        p.stack = (p.img == 0 && CO <= 32 && W % 2 == 0 && W >= 24) ? 1 : 0;   // 64-byte dY rows: TMA destinations stay 128-byte aligned only for even W + 2
        p.b_rows = ((127 + 2 * p.stack) / W2 + 2) * W2;
        p.n_tiles = (int)(((long)(H + 2) * W2 * Nimg + 127) / 128);
        if (p.img > 0) {
            p.a_rows = ((W2 + 1) + 128 + (W2 + 1) + 7) / 8 * 8;
            p.b_rows = 128;
            p.n_tiles = (Nimg + p.img - 1) / p.img;
        }
        const size_t stage_b = (size_t)(2 * p.a_rows) * 128 + (size_t)p.b_rows * (p.stack ? 64 : 128);
        p.stages = (2 * stage_b + 1024 <= 227 * 1024 - 256) ? 2 : 1;
        smem = p.stages * stage_b + 1024;
    }
    GN_REQUIRE(smem <= 227 * 1024 - 256, GN_EUNSUPPORTED, "conv3x3_wgrad: tile does not fit shared memory (%zu B)", smem);
    CUtensorMap tmX, tmDY;
    {
        uint64_t dims[4] = {(uint64_t)CI, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        uint64_t strides[3] = {(uint64_t)ldx * 2, (uint64_t)W * ldx * 2, (uint64_t)H * W * ldx * 2};
        uint32_t box[4] = {64, (uint32_t)W2, p.img > 0 ? (uint32_t)(H + 2) : 1u, 1};
        if (p.chunk) { box[2] = p.rt > 0 ? (uint32_t)(p.rt + 2) : (uint32_t)H2; box[3] = (uint32_t)p.ni; }
        int rc = gn_tmap_encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    {
        uint64_t dims[4] = {(uint64_t)CO, (uint64_t)W, (uint64_t)H, (uint64_t)Nimg};
        uint64_t strides[3] = {(uint64_t)ldy * 2, (uint64_t)W * ldy * 2, (uint64_t)H * W * ldy * 2};
        uint32_t box[4] = {p.stack ? 32u : 64u, (uint32_t)W2, p.img > 0 ? (uint32_t)(H + 2) : 1u, 1};
        if (p.chunk) { box[2] = p.rt > 0 ? (uint32_t)p.rt : (uint32_t)H2; box[3] = (uint32_t)p.ni; }
        int rc = gn_tmap_encode(&tmDY, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dy, dims, strides, box,
                                p.stack ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    static int max_set = 0;
    if ((int)smem > max_set) {
        GN_CUDA(cudaFuncSetAttribute(conv3x3_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        max_set = (int)smem;
    }
    const int grid = p.n_tiles < gn_num_sms() ? p.n_tiles : gn_num_sms();
    GN_CUDA(gn_launch(conv3x3_wgrad_kernel, dim3(grid), dim3(192), smem, stream, tmX, tmDY, p));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

GN_API int gn_conv3x3_unpack_grad(const float* dwp, int CO, int CI, float* dw, cudaStream_t stream) {
    GN_REQUIRE(dwp && dw && CO > 0 && CI > 0, GN_EINVAL, "conv3x3_unpack_grad: bad arguments");
    conv3_unpack_grad_kernel<<<gn_ceil_div(9L * CO * CI, 256), 256, 0, stream>>>(dwp, CO, CI, dw);
    GN_LAUNCH_CHECK();
    return GN_OK;
}
