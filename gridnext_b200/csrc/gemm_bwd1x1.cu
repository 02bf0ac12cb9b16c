// Backward of DenseNet's bottleneck 1x1 convolution as ONE tcgen05 kernel: data gradient + BatchNorm/ReLU backward + weight gradient.
//
//   forward (/root/reference/gridnext/densenet.py:12-18,26-27):  z = conv1(relu(norm1(cat)))   cat = C[:, :c_in],  z: [M, 128]
//   backward, per dense layer (what autograd derives for those lines):
//     (1) dC[:, :c_in] += (dz @ W1) * [a > 0] * sc        a = C * sc + sh        -- gn_gemm_bf16 with the BnBwdEpi epilogue
//     (2) dW1[128, c_in] += dz^T @ relu(a)                                        -- gn_gemm_tn_bf16 with the operand transform
// Both are HBM-bound streams over the SAME two operands: (1) reads dz [M,128] and C [M,c_in] and read-modify-writes dC, (2) reads
// dz and C again.  Layer by layer (2) cost 8.9 ms of the 74 ms DenseNet-121 step for nothing but re-reading.  Here the tile that (1)
// already holds on chip feeds (2):
//   * the dz tile [128 positions x 128 channels] sits in shared memory as the K-major A operand of (1); read through an MN-major
//     descriptor the same bytes are dz^T, the A operand of (2);
//   * the epilogue of (1) has the raw C tile in its staging slot and computes a = C*sc + sh for the ReLU mask anyway; it now also
//     writes relu(a) as bf16 into a second slot with the same 128-byte-swizzled layout, which IS the MN-major B operand of (2);
//   * the weight-gradient accumulator [128 x 256 fp32] lives in tensor memory for the whole life of the persistent CTA (every CTA
//     owns ONE 256-column block of c_in for all its row tiles) and is added to dW1 with vector atomics once, at the end.
// Tensor memory: columns 0..255 = two 128-column halves of the data-gradient accumulator (the epilogue works on one half while the
// tensor core fills the other), columns 256..511 = the weight-gradient accumulator.
//
//   warp 0  TMA producer: resident W1 block once, then one dz tile (2 boxes) per row tile
//   warp 1  MMA issuer:   dgrad(unit k+2) is issued before wgrad(unit k) is waited for, so the epilogue never waits for the tensor core
//   warp 2  epilogue feeder: raw C sub-tiles (128 x 64) by TMA, L2 prefetch several sub-tiles ahead
//   warp 3  TMEM allocator + drain: one TMA reduction (dC +=) per finished sub-tile
//   warps 4-11 epilogue
#include "gn_common.cuh"
#include "gn_ptx.cuh"
#include "gn_tma.cuh"
#include "gn_epilogue.cuh"

using namespace gnptx;

#define BW_BM 128
#define BW_K 128                       // bottleneck width (bn_size * growth_rate of DenseNet-121)
#define BW_BN 256                      // c_in columns owned by one CTA
#define BW_KB_BYTES 16384              // one 64-channel k-block of a dz tile: 128 rows x 128 B
#define BW_A_BYTES (2 * BW_KB_BYTES)
#define BW_SUB_BYTES 16384             // one 128 x 64 bf16 sub-tile (C reference / result / relu(a))
#define BW_MAX_ATILES 3
#define BW_ESTAGES 3
#define BW_XSTAGES 2
#define BW_EPI_WARPS 8
#define BW_PREFETCH 6                  // sub-tiles of C the feeder asks L2 for ahead of the shared-memory ring

#define BW_MAX_NB 8                    // column blocks (N <= 2048)

struct BwdParams {
    int M, N;
    int num_m_blocks, nnb;
    // Column blocks are balanced (N = 288 is 3 + 2 sub-tiles of 64 columns, not 4 + 1) and every block gets CTAs in proportion to its
    // width, so all CTAs stream the same number of bytes and walk the row tiles at the same pace (the dz tile of a row tile is then
    // fetched from HBM once and found in L2 by the other blocks' CTAs).  With fixed 256-column blocks and an equal CTA split the
    // c_in = 288..480 layers of dense block 2 ran at 3.2-4.8 TB/s: half the CTAs owned 32..224 columns, the other half 256.
    int sub_begin[BW_MAX_NB + 1];      // first 64-column sub-tile of block nb
    int cta_begin[BW_MAX_NB + 1];      // first CTA of block nb
    int a_tiles;                       // dz tile ring depth
    int b_rows;                        // rows of the resident weight block: 128 (N <= 128) or 256
    float* dw;                         // [128, lddw] fp32, atomically accumulated
    long lddw;
    BnBwdEpi bn;
};

__device__ __forceinline__ uint32_t bw_pack(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 bw_unpack(uint32_t w) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w)); }

__global__ void __launch_bounds__(384, 1)
gemm_bwd1x1_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                   const __grid_constant__ CUtensorMap tmRef, const BwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_afull[BW_MAX_ATILES], bar_aempty[BW_MAX_ATILES], bar_tfull[2], bar_tempty[2];
    __shared__ __align__(8) uint64_t bar_efull[BW_ESTAGES], bar_eready[BW_ESTAGES], bar_eempty[BW_ESTAGES];
    __shared__ __align__(8) uint64_t bar_xfull[BW_XSTAGES], bar_xempty[BW_XSTAGES], bar_bres, bar_d2full;
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(16) float s_epi[4][BW_BN];       // sc, sh, sum g, sum g*ref of this CTA's columns

    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // this CTA's column block and row tiles
    int nb = 0;
    while (nb + 1 < p.nnb && (int)blockIdx.x >= p.cta_begin[nb + 1]) ++nb;
    const int cta_m = blockIdx.x - p.cta_begin[nb], m_stride = p.cta_begin[nb + 1] - p.cta_begin[nb];
    const int n_my = cta_m < p.num_m_blocks ? (p.num_m_blocks - cta_m + m_stride - 1) / m_stride : 0;
    const int col_base = p.sub_begin[nb] * 64;
    const int ncols = min((p.sub_begin[nb + 1] - p.sub_begin[nb]) * 64, p.N - col_base);
    const int halves = (ncols + 127) >> 7;
    const int nsub_all = (ncols + 63) >> 6;
    const int n_units = n_my * halves;

    uint8_t* s_a = sm;                                              // [a_tiles][2][128 x 128 B]
    uint8_t* s_b = s_a + (size_t)p.a_tiles * BW_A_BYTES;            // [2][b_rows x 128 B]
    const int b_kb_bytes = p.b_rows * 128;
    uint8_t* s_slots = s_b + 2 * (size_t)b_kb_bytes;                // [BW_ESTAGES][16 KB]
    uint8_t* s_xf = s_slots + (size_t)BW_ESTAGES * BW_SUB_BYTES;    // [BW_XSTAGES][16 KB]

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmOut);
        tma_prefetch_desc(&tmRef);
        for (int s = 0; s < BW_MAX_ATILES; ++s) { mbar_init(&bar_afull[s], 1); mbar_init(&bar_aempty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&bar_tfull[s], 1); mbar_init(&bar_tempty[s], BW_EPI_WARPS); }
        for (int s = 0; s < BW_ESTAGES; ++s) { mbar_init(&bar_efull[s], 1); mbar_init(&bar_eready[s], BW_EPI_WARPS); mbar_init(&bar_eempty[s], 1); }
        for (int s = 0; s < BW_XSTAGES; ++s) { mbar_init(&bar_xfull[s], BW_EPI_WARPS); mbar_init(&bar_xempty[s], 1); }
        mbar_init(&bar_bres, 1);
        mbar_init(&bar_d2full, 1);
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc<512>(&tmem_slot);
    gn_pdl_wait();
    for (int i = threadIdx.x; i < BW_BN; i += blockDim.x) {
        const bool in = i < ncols;
        s_epi[0][i] = in ? p.bn.sc[col_base + i] : 0.f;
        s_epi[1][i] = in ? p.bn.sh[col_base + i] : 0.f;
        s_epi[2][i] = 0.f;
        s_epi[3][i] = 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    gn_pdl_trigger();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one() && n_my > 0) {
            mbar_arrive_expect_tx(&bar_bres, (uint32_t)(2 * b_kb_bytes));
            for (int kb = 0; kb < 2; ++kb) tma_load_2d(&tmB, &bar_bres, s_b + (size_t)kb * b_kb_bytes, kb * 64, col_base);
            int slot = 0;
            uint32_t phase = 0;
            for (int t = 0; t < n_my; ++t) {
                const int mb = cta_m + t * m_stride;
                if (t + p.a_tiles < n_my) {                       // the tile after the ones the ring can hold: bring it to L2 now
                    const int mbp = cta_m + (t + p.a_tiles) * m_stride;
                    tma_prefetch_l2_2d(&tmA, 0, mbp * BW_BM);
                    tma_prefetch_l2_2d(&tmA, 64, mbp * BW_BM);
                }
                mbar_wait(&bar_aempty[slot], phase ^ 1);
                uint8_t* sa = s_a + (size_t)slot * BW_A_BYTES;
                mbar_arrive_expect_tx(&bar_afull[slot], BW_A_BYTES);
                tma_load_2d(&tmA, &bar_afull[slot], sa, 0, mb * BW_BM);
                tma_load_2d(&tmA, &bar_afull[slot], sa + BW_KB_BYTES, 64, mb * BW_BM);
                if (++slot == p.a_tiles) { slot = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one() && n_my > 0) {
            constexpr uint64_t tmplK = smem_desc_template(0, 1024, LAYOUT_SW128);                 // K-major: dz tile and W1 block
            constexpr uint64_t tmplMN = smem_desc_template(BW_KB_BYTES, 1024, LAYOUT_SW128);      // MN-major: dz^T (two 64-channel groups) and relu(a)
            constexpr uint32_t idesc_w = idesc_bf16(128, 64, 1, 1);
            const uint32_t a_base = smem_u32(s_a), b_base = smem_u32(s_b), x_base = smem_u32(s_xf);
            mbar_wait(&bar_bres, 0);
            int xs = 0;
            uint32_t xph = 0;
            auto dgrad = [&](int k) {
                const int t = k / halves, hf = k - t * halves, buf = k & 1;
                const int slot = t % p.a_tiles;
                if (hf == 0) mbar_wait(&bar_afull[slot], (uint32_t)((t / p.a_tiles) & 1));
                mbar_wait(&bar_tempty[buf], (uint32_t)(((k >> 1) & 1) ^ 1));
                tc_fence_after();
                const int nh = ((min(128, ncols - hf * 128) + 63) >> 6) << 6;
                const uint32_t idesc_d = idesc_bf16(128, nh, 0, 0);
                const uint32_t d = tmem_base + (uint32_t)(buf * 128);
                const uint32_t a_addr = a_base + (uint32_t)(slot * BW_A_BYTES);
                const uint32_t b_addr = b_base + (uint32_t)(hf * 128 * 128);
#pragma unroll
                for (int kb = 0; kb < 2; ++kb)
#pragma unroll
                    for (int k16 = 0; k16 < 4; ++k16)
                        umma_bf16(d, smem_desc(tmplK, a_addr + kb * BW_KB_BYTES + k16 * 32), smem_desc(tmplK, b_addr + kb * b_kb_bytes + k16 * 32), idesc_d,
                                  (uint32_t)((kb | k16) != 0));
                umma_commit(&bar_tfull[buf]);
            };
            auto wgrad = [&](int k) {
                const int t = k / halves, hf = k - t * halves;
                const int slot = t % p.a_tiles;
                const uint32_t a_addr = a_base + (uint32_t)(slot * BW_A_BYTES);
                const int j1 = min(nsub_all, 2 * hf + 2);
                for (int j = 2 * hf; j < j1; ++j) {
                    mbar_wait(&bar_xfull[xs], xph);
                    tc_fence_after();
                    const uint32_t d = tmem_base + 256u + (uint32_t)(j * 64);
                    const uint32_t x_addr = x_base + (uint32_t)(xs * BW_SUB_BYTES);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk)
                        umma_bf16(d, smem_desc(tmplMN, a_addr + kk * 2048), smem_desc(tmplMN, x_addr + kk * 2048), idesc_w, (uint32_t)((t | kk) != 0));
                    umma_commit(&bar_xempty[xs]);
                    if (++xs == BW_XSTAGES) { xs = 0; xph ^= 1; }
                }
                if (hf == halves - 1) umma_commit(&bar_aempty[slot]);     // every MMA that reads this dz tile has been issued
            };
            dgrad(0);
            if (n_units > 1) dgrad(1);
            for (int k = 0; k < n_units; ++k) {
                wgrad(k);
                if (k + 2 < n_units) dgrad(k + 2);
            }
            umma_commit(&bar_d2full);
        }
    } else if (warp == 2) {
        // ===================== epilogue feeder =====================
        if (elect_one()) {
            int es = 0;
            uint32_t eph = 0;
            int tp = 0, jp = 0;                                   // prefetch cursor, BW_PREFETCH sub-tiles ahead
            for (int q = 0; q < BW_PREFETCH && tp < n_my; ++q) {
                tma_prefetch_l2_2d(&tmRef, col_base + jp * 64, (cta_m + tp * m_stride) * BW_BM);
                if (++jp == nsub_all) { jp = 0; ++tp; }
            }
            for (int t = 0; t < n_my; ++t) {
                const int mb = cta_m + t * m_stride;
                for (int j = 0; j < nsub_all; ++j) {
                    if (tp < n_my) {
                        tma_prefetch_l2_2d(&tmRef, col_base + jp * 64, (cta_m + tp * m_stride) * BW_BM);
                        if (++jp == nsub_all) { jp = 0; ++tp; }
                    }
                    mbar_wait(&bar_eempty[es], eph ^ 1);
                    mbar_arrive_expect_tx(&bar_efull[es], BW_SUB_BYTES);
                    tma_load_2d(&tmRef, &bar_efull[es], s_slots + (size_t)es * BW_SUB_BYTES, col_base + j * 64, mb * BW_BM);
                    if (++es == BW_ESTAGES) { es = 0; eph ^= 1; }
                }
            }
        }
    } else if (warp == 3) {
        // ===================== epilogue drain =====================
        if (elect_one()) {
            int es = 0;
            uint32_t eph = 0;
            for (int t = 0; t < n_my; ++t) {
                const int mb = cta_m + t * m_stride;
                for (int j = 0; j < nsub_all; ++j) {
                    mbar_wait(&bar_eready[es], eph);
                    if (p.bn.rmw) tma_reduce_add_2d(&tmOut, s_slots + (size_t)es * BW_SUB_BYTES, col_base + j * 64, mb * BW_BM);
                    else tma_store_2d(&tmOut, s_slots + (size_t)es * BW_SUB_BYTES, col_base + j * 64, mb * BW_BM);
                    tma_store_commit();
                    tma_store_wait_read<0>();
                    mbar_arrive(&bar_eempty[es]);
                    if (++es == BW_ESTAGES) { es = 0; eph ^= 1; }
                }
            }
            tma_store_wait_all<0>();
        }
        __syncwarp();
    } else {
        // ===================== epilogue: thread = (tile row, 32-column half of each 64-column sub-tile) =====================
        const int g = warp & 3;
        const int h = (warp - 4) >> 2;
        const int trow = g * 32 + lane;
        const uint32_t sw = (uint32_t)(trow & 7);
        const bool want_sums = p.bn.colsum != nullptr;
        int es = 0, xs = 0;
        uint32_t eph = 0, xph = 0;
        for (int k = 0; k < n_units; ++k) {
            const int t = k / halves, hf = k - t * halves, buf = k & 1;
            mbar_wait(&bar_tfull[buf], (uint32_t)((k >> 1) & 1));
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(buf * 128 + h * 32);
            const int j1 = min(nsub_all, 2 * hf + 2);
#pragma unroll 1
            for (int j = 2 * hf; j < j1; ++j) {
                mbar_wait(&bar_efull[es], eph);
                mbar_wait(&bar_xempty[xs], xph ^ 1);
                uint8_t* row_p = s_slots + (size_t)es * BW_SUB_BYTES + trow * 128;       // C reference in, dC contribution out (in place)
                uint8_t* row_x = s_xf + (size_t)xs * BW_SUB_BYTES + trow * 128;          // relu(a): the weight gradient's B operand
                __syncwarp();
                uint32_t r[32];
                tmem_ld32(taddr + (uint32_t)((j - 2 * hf) * 64), r);
                tmem_ld_wait();
                const int c0 = j * 64 + h * 32;
                float v[32], gx[32];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const uint32_t off = (((uint32_t)(h * 4 + q)) ^ sw) << 4;
                    const float4 c0a = *reinterpret_cast<const float4*>(&s_epi[0][c0 + 8 * q]), c0b = *reinterpret_cast<const float4*>(&s_epi[0][c0 + 8 * q + 4]);
                    const float4 c1a = *reinterpret_cast<const float4*>(&s_epi[1][c0 + 8 * q]), c1b = *reinterpret_cast<const float4*>(&s_epi[1][c0 + 8 * q + 4]);
                    const float k0[8] = {c0a.x, c0a.y, c0a.z, c0a.w, c0b.x, c0b.y, c0b.z, c0b.w};
                    const float k1[8] = {c1a.x, c1a.y, c1a.z, c1a.w, c1b.x, c1b.y, c1b.z, c1b.w};
                    const uint4 rv = *reinterpret_cast<const uint4*>(row_p + off);
                    const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
                    uint32_t res[4], xf[4];
#pragma unroll
                    for (int e2 = 0; e2 < 4; ++e2) {
                        // packed fp32 pairs: a = ref*sc + sh (FFMA2), relu(a) -> bf16 in the pack instruction, g*ref and g*sc as FMUL2
                        const uint64_t ref2 = bf16x2_to_f32x2(rw[e2]);
                        const uint64_t sc2 = f32x2(k0[2 * e2], k0[2 * e2 + 1]);
                        const uint64_t a2 = ffma2(ref2, sc2, f32x2(k1[2 * e2], k1[2 * e2 + 1]));
                        float a_lo, a_hi;
                        f32x2_unpack(a2, a_lo, a_hi);
                        const float g_lo = a_lo > 0.f ? __uint_as_float(r[8 * q + 2 * e2]) : 0.f;
                        const float g_hi = a_hi > 0.f ? __uint_as_float(r[8 * q + 2 * e2 + 1]) : 0.f;
                        const uint64_t g2 = f32x2(g_lo, g_hi);
                        v[8 * q + 2 * e2] = g_lo;
                        v[8 * q + 2 * e2 + 1] = g_hi;
                        f32x2_unpack(fmul2(g2, ref2), gx[8 * q + 2 * e2], gx[8 * q + 2 * e2 + 1]);   // sum g*xhat = p1 * (sum g*ref - p0 * sum g), applied at the flush
                        res[e2] = f32x2_to_bf16x2(fmul2(g2, sc2));
                        xf[e2] = f32x2_to_bf16x2_relu(a2);
                    }
                    *reinterpret_cast<uint4*>(row_p + off) = make_uint4(res[0], res[1], res[2], res[3]);
                    *reinterpret_cast<uint4*>(row_x + off) = make_uint4(xf[0], xf[1], xf[2], xf[3]);
                }
                if (want_sums) {
                    const float sg = gn_warp_colsum32(v, lane), sx = gn_warp_colsum32(gx, lane);
                    atomicAdd(&s_epi[2][c0 + lane], sg);
                    atomicAdd(&s_epi[3][c0 + lane], sx);
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&bar_eready[es]);
                    mbar_arrive(&bar_xfull[xs]);
                }
                if (++es == BW_ESTAGES) { es = 0; eph ^= 1; }
                if (++xs == BW_XSTAGES) { xs = 0; xph ^= 1; }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[buf]);
        }
        if (n_my > 0) {
            // ---- weight gradient: TMEM columns 256.. -> dW1[row = bottleneck channel, col_base + c] with vector atomics
            mbar_wait(&bar_d2full, 0);
            tc_fence_after();
            const uint32_t taddr2 = tmem_base + ((uint32_t)(g * 32) << 16) + 256u;
            float* orow = p.dw + (long)trow * p.lddw + col_base;
#pragma unroll 1
            for (int c0 = h * 32; c0 < nsub_all * 64; c0 += 64) {
                uint32_t r[32];
                tmem_ld32(taddr2 + (uint32_t)c0, r);
                tmem_ld_wait();
                if (c0 < ncols) {
                    float* o = orow + c0;
                    if (c0 + 32 <= ncols && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                        for (int q = 0; q < 8; ++q)
                            atomicAdd(reinterpret_cast<float4*>(o + 4 * q), make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]),
                                                                                        __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])));
                    } else {
#pragma unroll
                        for (int e = 0; e < 32; ++e)
                            if (c0 + e < ncols) atomicAdd(o + e, __uint_as_float(r[e]));
                    }
                }
            }
            tc_fence_before();
            if (want_sums) {
                named_bar_sync(1, BW_EPI_WARPS * 32);
                for (int c = threadIdx.x - 4 * 32; c < ncols; c += BW_EPI_WARPS * 32) {
                    const float sg = s_epi[2][c], sx = s_epi[3][c];
                    const int col = col_base + c;
                    atomicAdd(p.bn.colsum + col, sg);
                    atomicAdd(p.bn.colsum + p.bn.ldsum + col, __ldg(p.bn.p1 + col) * (sx - __ldg(p.bn.p0 + col) * sg));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) tmem_dealloc<512>(tmem_base);
}

// dz: [M, 128] bf16 (pitch lddz); wt: W1 transposed, [N = c_in, 128] bf16 (pitch ldw); ref: raw C[:, :c_in] (pitch ldref);
// dx: dC[:, :c_in] bf16 (pitch lddx), dx (+)= (dz @ wt^T) * [ref*sc + sh > 0] * sc; colsum as in BnBwdEpi; dw: [128, lddw] fp32, dw += dz^T relu(ref*sc + sh).
GN_API int gn_conv1x1_bwd_bf16(const void* dz, long lddz, const void* wt, long ldw, int M, int N, void* dx, long lddx, const void* ref, long ldref,
                               const float* sc, const float* sh, const float* p0, const float* p1, float* colsum, int ldsum, int rmw, float* dw,
                               long lddw, cudaStream_t stream) {
    GN_REQUIRE(dz && wt && dx && ref && sc && sh && p0 && p1 && dw && M > 0 && N > 0, GN_EINVAL, "conv1x1_bwd: bad arguments");
    GN_REQUIRE(lddz >= BW_K && ldw >= BW_K && lddx >= N && ldref >= N && lddw >= N, GN_EINVAL, "conv1x1_bwd: pitch smaller than extent");
    auto ok = [](const void* ptr, long ld) { return ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0) && (ld % 8 == 0); };
    GN_REQUIRE(ok(dz, lddz) && ok(wt, ldw) && ok(dx, lddx) && ok(ref, ldref), GN_EALIGN,
               "conv1x1_bwd: operands must be 16-byte aligned with pitches that are multiples of 8 elements");
    BwdParams p;
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N;
    p.num_m_blocks = gn_ceil_div(M, BW_BM);
    const int nsub = gn_ceil_div(N, 64);
    p.nnb = gn_ceil_div(nsub, BW_BN / 64);
    GN_REQUIRE(p.nnb <= BW_MAX_NB, GN_EUNSUPPORTED, "conv1x1_bwd: at most %d input channels", BW_MAX_NB * BW_BN);
    int grid = 0;
    {
        const int sms = gn_num_sms();
        int given = 0;
        p.sub_begin[0] = 0;
        for (int b = 0; b < p.nnb; ++b) p.sub_begin[b + 1] = p.sub_begin[b] + nsub / p.nnb + (b < nsub % p.nnb ? 1 : 0);
        int cnt[BW_MAX_NB];
        for (int b = 0; b < p.nnb; ++b) {
            cnt[b] = sms * (p.sub_begin[b + 1] - p.sub_begin[b]) / nsub;
            if (cnt[b] < 1) cnt[b] = 1;
            given += cnt[b];
        }
        for (int b = 0; given < sms; b = (b + 1) % p.nnb) { ++cnt[b]; ++given; }      // leftovers to the (wider) first blocks
        p.cta_begin[0] = 0;
        for (int b = 0; b < p.nnb; ++b) {
            if (cnt[b] > p.num_m_blocks) cnt[b] = p.num_m_blocks;
            p.cta_begin[b + 1] = p.cta_begin[b] + cnt[b];
        }
        grid = p.cta_begin[p.nnb];
    }
    p.b_rows = N <= 128 ? 128 : 256;
    p.a_tiles = p.b_rows == 128 ? 3 : 2;
    p.dw = dw; p.lddw = lddw;
    p.bn.ref = (const __nv_bfloat16*)ref; p.bn.ldref = ldref; p.bn.ref_is_raw = 1;
    p.bn.sc = sc; p.bn.sh = sh; p.bn.p0 = p0; p.bn.p1 = p1; p.bn.colsum = colsum; p.bn.ldsum = ldsum; p.bn.rmw = rmw;
    CUtensorMap tmA, tmB, tmOut, tmRef;
    int rc = gn_tmap_bf16_2d(&tmA, dz, (uint64_t)M, (uint64_t)BW_K, (uint64_t)lddz, 64, BW_BM);
    if (rc) return rc;
    rc = gn_tmap_bf16_2d(&tmB, wt, (uint64_t)N, (uint64_t)BW_K, (uint64_t)ldw, 64, (uint32_t)p.b_rows);
    if (rc) return rc;
    rc = gn_tmap_bf16_2d(&tmOut, dx, (uint64_t)M, (uint64_t)N, (uint64_t)lddx, 64, BW_BM);
    if (rc) return rc;
    rc = gn_tmap_bf16_2d(&tmRef, ref, (uint64_t)M, (uint64_t)N, (uint64_t)ldref, 64, BW_BM);
    if (rc) return rc;
    const size_t smem = (size_t)p.a_tiles * BW_A_BYTES + 2 * (size_t)p.b_rows * 128 + (size_t)(BW_ESTAGES + BW_XSTAGES) * BW_SUB_BYTES + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        GN_CUDA(cudaFuncSetAttribute(gemm_bwd1x1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     2 * BW_A_BYTES + 2 * 256 * 128 + (BW_ESTAGES + BW_XSTAGES) * BW_SUB_BYTES + 1024));   // = 3 tiles + the 128-row block
        attr_set = true;
    }
    GN_CUDA(gn_launch(gemm_bwd1x1_kernel, dim3(grid), dim3(384), smem, stream, tmA, tmB, tmOut, tmRef, p));
    GN_LAUNCH_CHECK();
    return GN_OK;
}
