// tcgen05 / TMEM / TMA GEMM for the f networks (bf16 operands, fp32 accumulation in tensor memory).
//
//   D[M, N] = op(A)[M, K] * B[N, K]^T            A, B row-major ("K-major"), D row-major with pitch ldc
//
// used for every 1x1 convolution of DenseNet-121 in NHWC (/root/reference/gridnext/densenet.py:26-27,52-53:
// conv1 of each dense layer, the transitions), their data gradients, and the Linear layers of the count MLP
// (notebooks/Tutorial_visium_count.ipynb cell 12).
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0      TMA producer     (cp.async.bulk.tensor 2-D boxes of 128 x 64 / BN x 64 bf16, SWIZZLE_128B)
//   warp 1      MMA issuer       (one elected thread, tcgen05.mma.cta_group::1.kind::f16, UMMA 128 x BN x 16)
//   warps 2-5   epilogue         (tcgen05.ld 32x32b -> affine / ReLU -> bf16|fp32 global stores), double-buffered
//                                 TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 6-9   operand transform (XFORM only): DenseNet's pre-activation BatchNorm+ReLU
//                                 (densenet.py:12-18: conv(relu(norm(cat)))) is applied IN PLACE to the A tile in
//                                 shared memory between TMA arrival and the MMA, so the activated tensor is never
//                                 written to HBM.  BN is eval-mode (training.py:126): x * scale[k] + shift[k].
#include "gn_common.cuh"
#include "gn_ptx.cuh"
#include "gn_tma.cuh"
#include "gn_epilogue.cuh"

using namespace gnptx;

#define GEMM_BM 128
#define GEMM_BK 64
#define GEMM_A_BYTES (GEMM_BM * GEMM_BK * 2)
#define GEMM_MAX_XF_K 1024

struct GemmParams {
    int M, N, K;
    int num_m_blocks, num_n_blocks, num_k_blocks;
    void* out;
    long ldc;
    int out_fp32;
    int accumulate;            // fp32 only: out += result
    const float* scale;        // per output column (nullable)
    const float* shift;        // per output column (nullable)
    int relu;
    const float* xf_scale;     // per K index (XFORM kernels)
    const float* xf_shift;
    int epi_mode;              // 0: affine/ReLU store, 1: BN+ReLU backward (BnBwdEpi), bf16 output
    BnBwdEpi bn;
};

template <int BN> struct GemmCfg {
    static constexpr int STAGES = (BN <= 128) ? 6 : 4;
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = GEMM_A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr size_t smem_bytes(bool xform) { return (size_t)STAGES * STAGE_BYTES + 1024 + (xform ? 2 * GEMM_MAX_XF_K * 4 : 0); }
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}

// store 32 consecutive columns of one row (mode-0 epilogue)
__device__ __forceinline__ void epi_store_row(const GemmParams& p, int row, int col, const uint32_t (&r)[32]) {
    if (row >= p.M || col >= p.N) return;
    float v[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
    const int ncols = min(32, p.N - col);
    if (p.scale != nullptr || p.shift != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            if (j < ncols) {
                const float sc = p.scale ? __ldg(p.scale + col + j) : 1.f;
                const float sh = p.shift ? __ldg(p.shift + col + j) : 0.f;
                v[j] = fmaf(v[j], sc, sh);
            }
        }
    }
    if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (p.out_fp32) {
        float* o = reinterpret_cast<float*>(p.out) + (long)row * p.ldc + col;
        const bool vec = (ncols == 32) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
        if (vec) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                float4 t = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
                if (p.accumulate) {
                    const float4 old = *reinterpret_cast<const float4*>(o + 4 * q);
                    t.x += old.x; t.y += old.y; t.z += old.z; t.w += old.w;
                }
                *reinterpret_cast<float4*>(o + 4 * q) = t;
            }
        } else {
            for (int j = 0; j < ncols; ++j) o[j] = p.accumulate ? o[j] + v[j] : v[j];
        }
    } else {
        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (long)row * p.ldc + col;
        const bool vec = (ncols == 32) && ((reinterpret_cast<uintptr_t>(o) & 15) == 0);
        if (vec) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                uint4 t;
                t.x = pack_bf16x2(v[8 * q + 0], v[8 * q + 1]);
                t.y = pack_bf16x2(v[8 * q + 2], v[8 * q + 3]);
                t.z = pack_bf16x2(v[8 * q + 4], v[8 * q + 5]);
                t.w = pack_bf16x2(v[8 * q + 6], v[8 * q + 7]);
                *reinterpret_cast<uint4*>(o + 8 * q) = t;
            }
        } else {
            for (int j = 0; j < ncols; ++j) o[j] = __float2bfloat16_rn(v[j]);
        }
    }
}

template <int BN, bool XFORM>
__global__ void __launch_bounds__(XFORM ? 320 : 192, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[STAGES], bar_xf[STAGES], bar_empty[STAGES], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_slot;

    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    float* s_xf = reinterpret_cast<float*>(sm + (size_t)STAGES * Cfg::STAGE_BYTES);   // [2][GEMM_MAX_XF_K]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (XFORM) {
        const int kpad = p.num_k_blocks * GEMM_BK;
        for (int i = threadIdx.x; i < kpad; i += blockDim.x) {
            s_xf[i] = i < p.K ? p.xf_scale[i] : 0.f;
            s_xf[GEMM_MAX_XF_K + i] = i < p.K ? p.xf_shift[i] : 0.f;
        }
    }
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_xf[s], 4);
            mbar_init(&bar_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bar_tfull[a], 1);
            mbar_init(&bar_tempty[a], 4);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<Cfg::TMEM_COLS>(&tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const int n_tiles = p.num_m_blocks * p.num_n_blocks;
    const int nkb = p.num_k_blocks;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int nb = tile / p.num_m_blocks, mb = tile % p.num_m_blocks;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    uint8_t* sa = sm + (size_t)stage * Cfg::STAGE_BYTES;
                    mbar_arrive_expect_tx(&bar_full[stage], Cfg::STAGE_BYTES);
                    tma_load_2d(&tmA, &bar_full[stage], sa, kb * GEMM_BK, mb * GEMM_BM);
                    tma_load_2d(&tmB, &bar_full[stage], sa + GEMM_A_BYTES, kb * GEMM_BK, nb * BN);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_bf16(GEMM_BM, BN, 0, 0);
            constexpr uint64_t tmpl = smem_desc_template(0, 1024, LAYOUT_SW128);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(XFORM ? &bar_xf[stage] : &bar_full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sm + (size_t)stage * Cfg::STAGE_BYTES);
                    const uint32_t b_addr = a_addr + GEMM_A_BYTES;
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k)
                        umma_bf16(d, smem_desc(tmpl, a_addr + k * 32), smem_desc(tmpl, b_addr + k * 32), idesc, (uint32_t)((kb | k) != 0));
                    umma_commit(&bar_empty[stage]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bar_tfull[acc]);
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp < 6) {
        // ===================== epilogue =====================
        const int g = warp & 3;   // TMEM lane group this warp may access
        int acc = 0;
        uint32_t acc_phase = 0;
        constexpr int NCH = BN / 32;
        float cs_g[NCH], cs_x[NCH];   // BN-backward column partial sums (column = nb*BN + ch*32 + lane)
#pragma unroll
        for (int i = 0; i < NCH; ++i) { cs_g[i] = 0.f; cs_x[i] = 0.f; }
        int cur_nb = -1;
        auto flush_colsums = [&]() {
            if (p.epi_mode == 1 && p.bn.colsum != nullptr && cur_nb >= 0) {
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    const int col = cur_nb * BN + ch * 32 + lane;
                    if (col < p.N) {
                        atomicAdd(p.bn.colsum + col, cs_g[ch]);
                        atomicAdd(p.bn.colsum + p.bn.ldsum + col, cs_x[ch]);
                    }
                    cs_g[ch] = 0.f; cs_x[ch] = 0.f;
                }
            }
        };
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int nb = tile / p.num_m_blocks, mb = tile % p.num_m_blocks;
            if (nb != cur_nb) { flush_colsums(); cur_nb = nb; }
            mbar_wait(&bar_tfull[acc], acc_phase);
            tc_fence_after();
            const int row = mb * GEMM_BM + g * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(acc * BN);
            if (p.epi_mode == 0) {
#pragma unroll 1
                for (int c0 = 0; c0 < BN; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c0, r);
                    tmem_ld_wait();
                    epi_store_row(p, row, nb * BN + c0, r);
                }
            } else {
#pragma unroll
                for (int ch = 0; ch < NCH; ++ch) {
                    uint32_t r[32];
                    tmem_ld32(taddr + ch * 32, r);
                    tmem_ld_wait();
                    const int col = nb * BN + ch * 32;
                    const int ncols = min(32, p.N - col);
                    float v[32], gx[32];
                    if (row < p.M && ncols > 0) {
                        float ref[32];
                        gn_load_bf16_32(p.bn.ref + (long)row * p.bn.ldref + col, ref, ncols);
                        __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (long)row * p.ldc + col;
                        float old[32];
                        if (p.bn.rmw) gn_load_bf16_32(o, old, ncols);
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (j < ncols) {
                                const float sc = __ldg(p.bn.sc + col + j);
                                const float a = p.bn.ref_is_raw ? fmaf(ref[j], sc, __ldg(p.bn.sh + col + j)) : ref[j];
                                const float gg = a > 0.f ? __uint_as_float(r[j]) : 0.f;
                                gx[j] = gg * (ref[j] - __ldg(p.bn.p0 + col + j)) * __ldg(p.bn.p1 + col + j);
                                v[j] = gg;
                                ref[j] = p.bn.rmw ? fmaf(gg, sc, old[j]) : gg * sc;
                            } else {
                                gx[j] = 0.f; v[j] = 0.f; ref[j] = 0.f;
                            }
                        }
                        gn_store_bf16_32(o, ref, ncols);
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
        flush_colsums();
    } else if (XFORM) {
        // ===================== operand transform: A <- relu(A * scale[k] + shift[k]) in place =====================
        const int w = warp - 6;
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            for (int kb = 0; kb < nkb; ++kb) {
                mbar_wait(&bar_full[stage], phase);
                uint8_t* sa = sm + (size_t)stage * Cfg::STAGE_BYTES;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int row = w * 32 + i * 4 + (lane >> 3);
                    const int pc = lane & 7;                        // physical 16-byte chunk in the 128-byte row
                    const int k0 = kb * GEMM_BK + ((pc ^ (row & 7)) << 3);   // logical K index of its first element
                    uint4* ptr = reinterpret_cast<uint4*>(sa + row * 128 + pc * 16);
                    uint4 v = *ptr;
                    const float4 s0 = *reinterpret_cast<const float4*>(s_xf + k0);
                    const float4 s1 = *reinterpret_cast<const float4*>(s_xf + k0 + 4);
                    const float4 t0 = *reinterpret_cast<const float4*>(s_xf + GEMM_MAX_XF_K + k0);
                    const float4 t1 = *reinterpret_cast<const float4*>(s_xf + GEMM_MAX_XF_K + k0 + 4);
                    float2 f;
                    f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v.x));
                    v.x = pack_bf16x2(fmaxf(fmaf(f.x, s0.x, t0.x), 0.f), fmaxf(fmaf(f.y, s0.y, t0.y), 0.f));
                    f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v.y));
                    v.y = pack_bf16x2(fmaxf(fmaf(f.x, s0.z, t0.z), 0.f), fmaxf(fmaf(f.y, s0.w, t0.w), 0.f));
                    f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v.z));
                    v.z = pack_bf16x2(fmaxf(fmaf(f.x, s1.x, t1.x), 0.f), fmaxf(fmaf(f.y, s1.y, t1.y), 0.f));
                    f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v.w));
                    v.w = pack_bf16x2(fmaxf(fmaf(f.x, s1.z, t1.z), 0.f), fmaxf(fmaf(f.y, s1.w, t1.w), 0.f));
                    *ptr = v;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_xf[stage]);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
}

template <int BN, bool XFORM>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    static bool attr_set = false;
    const size_t smem = Cfg::smem_bytes(XFORM);
    if (!attr_set) {
        GN_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN, XFORM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int tiles = p.num_m_blocks * p.num_n_blocks;
    const int grid = tiles < gn_num_sms() ? tiles : gn_num_sms();
    gemm_bf16_kernel<BN, XFORM><<<grid, XFORM ? 320 : 192, smem, stream>>>(tmA, tmB, p);
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// a: [M, K] bf16 with row pitch lda; b: [N, K] bf16 with row pitch ldb; out: [M, ldc] bf16 or fp32
GN_API int gn_gemm_bf16(const void* a, long lda, const void* b, long ldb, int M, int N, int K, void* out, long ldc, int out_fp32,
                        int accumulate, const float* scale, const float* shift, int relu, const float* xf_scale, const float* xf_shift,
                        const void* bn_ref, long bn_ldref, int bn_ref_is_raw, const float* bn_sc, const float* bn_sh, const float* bn_p0,
                        const float* bn_p1, float* bn_colsum, int bn_ldsum, int bn_rmw, cudaStream_t stream) {
    GN_REQUIRE(a && b && out && M > 0 && N > 0 && K > 0, GN_EINVAL, "gemm_bf16: bad arguments");
    GN_REQUIRE(lda >= K && ldb >= K && ldc >= N, GN_EINVAL, "gemm_bf16: pitch smaller than extent");
    GN_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, GN_EALIGN, "gemm_bf16: lda/ldb must be multiples of 8 elements (16 bytes)");
    GN_REQUIRE((xf_scale == nullptr) == (xf_shift == nullptr), GN_EINVAL, "gemm_bf16: xf_scale/xf_shift must come together");
    GN_REQUIRE(!accumulate || out_fp32, GN_EUNSUPPORTED, "gemm_bf16: accumulate needs an fp32 output");
    const bool xform = xf_scale != nullptr;
    GN_REQUIRE(!xform || K <= GEMM_MAX_XF_K, GN_EUNSUPPORTED, "gemm_bf16: operand transform supports K <= %d", GEMM_MAX_XF_K);
    const int BN = N <= 32 ? 32 : N <= 64 ? 64 : N <= 128 ? 128 : 256;
    GemmParams p;
    p.M = M; p.N = N; p.K = K;
    p.num_m_blocks = gn_ceil_div(M, GEMM_BM);
    p.num_n_blocks = gn_ceil_div(N, BN);
    p.num_k_blocks = gn_ceil_div(K, GEMM_BK);
    p.out = out; p.ldc = ldc; p.out_fp32 = out_fp32; p.accumulate = accumulate;
    p.scale = scale; p.shift = shift; p.relu = relu; p.xf_scale = xf_scale; p.xf_shift = xf_shift;
    p.epi_mode = bn_ref != nullptr ? 1 : 0;
    memset(&p.bn, 0, sizeof(p.bn));
    if (p.epi_mode == 1) {
        GN_REQUIRE(!out_fp32 && !accumulate && !scale && !shift && !relu, GN_EINVAL, "gemm_bf16: BN-backward epilogue needs a plain bf16 output");
        GN_REQUIRE(bn_sc && bn_p0 && bn_p1 && (!bn_ref_is_raw || bn_sh), GN_EINVAL, "gemm_bf16: incomplete BN-backward epilogue arguments");
        p.bn.ref = (const __nv_bfloat16*)bn_ref; p.bn.ldref = bn_ldref; p.bn.ref_is_raw = bn_ref_is_raw;
        p.bn.sc = bn_sc; p.bn.sh = bn_sh; p.bn.p0 = bn_p0; p.bn.p1 = bn_p1; p.bn.colsum = bn_colsum; p.bn.ldsum = bn_ldsum; p.bn.rmw = bn_rmw;
    }
    CUtensorMap tmA, tmB;
    int rc = gn_tmap_bf16_2d(&tmA, a, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BK, GEMM_BM);
    if (rc) return rc;
    rc = gn_tmap_bf16_2d(&tmB, b, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, GEMM_BK, (uint32_t)BN);
    if (rc) return rc;
#define GN_DISPATCH(BNv)                                                                  \
    case BNv:                                                                             \
        return xform ? launch_gemm<BNv, true>(tmA, tmB, p, stream) : launch_gemm<BNv, false>(tmA, tmB, p, stream);
    switch (BN) {
        GN_DISPATCH(32)
        GN_DISPATCH(64)
        GN_DISPATCH(128)
        GN_DISPATCH(256)
    }
#undef GN_DISPATCH
    return GN_EINVAL;
}
