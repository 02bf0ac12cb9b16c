// tcgen05 weight-gradient GEMM:   D[Mo, No] += sum_k A[k, Mo] * op(B)[k, No]     (fp32 output, atomically accumulated)
//
// Both operands are row-major matrices whose ROWS are the reduction index (pixels / spots), i.e. "MN-major" for
// the tensor core: 64-row x 64-channel TMA boxes land in shared memory as SWIZZLE_128B tiles that the UMMA reads
// through MN-major descriptors (validated with tools/umma_probe).  Used for
//   dW(conv 1x1)[Cout, Cin] = dZ[pix, Cout]^T * relu(bn(C[pix, Cin]))     (/root/reference/gridnext/densenet.py:26-27,52-53 backward)
//   dW(Linear)[out, in]     = dY[spot, out]^T * X[spot, in]                (count MLP backward)
// op(B) = relu(B * xf_scale[n] + xf_shift[n]) is DenseNet's pre-activation BatchNorm+ReLU recomputed in shared
// memory (the activated tensor is never stored).  The reduction dimension is split over CTAs; each CTA keeps its
// 128 x 256 partial in TMEM and adds it to D with vector atomics.
#include "gn_common.cuh"
#include "gn_ptx.cuh"
#include "gn_tma.cuh"

using namespace gnptx;

#define TN_BM 128          // D rows per tile  (A columns)
#define TN_BN 256          // D cols per tile  (B columns)
#define TN_BK 64           // reduction rows per stage
#define TN_STAGES 4
#define TN_A_BYTES (2 * 64 * 128)     // 2 groups of 64 channels x 64 rows x 128 B
#define TN_B_BYTES (4 * 64 * 128)
#define TN_STAGE_BYTES (TN_A_BYTES + TN_B_BYTES)
#define TN_GROUP_BYTES (64 * 128)

struct TnParams {
    int Mo, No, Kp;
    int mo_blocks, no_blocks, ksplit, kb_total;
    float* out;
    long ldo;
    const float* xf_scale;   // per B column (nullable)
    const float* xf_shift;
};

template <bool XFORM>
__global__ void __launch_bounds__(XFORM ? 320 : 192, 1)
gemm_tn_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TnParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[TN_STAGES], bar_xf[TN_STAGES], bar_empty[TN_STAGES], bar_tfull[2], bar_tempty[2];
    __shared__ uint32_t tmem_slot;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        for (int s = 0; s < TN_STAGES; ++s) {
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
    if (warp == 1) tmem_alloc<512>(&tmem_slot);
    gn_pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    gn_pdl_trigger();
    const uint32_t tmem_base = tmem_slot;
    const int n_units = p.mo_blocks * p.no_blocks * p.ksplit;
    const int kb_per = (p.kb_total + p.ksplit - 1) / p.ksplit;

    // unit -> (mb, nb, [kb0, kb1))
    auto decode = [&](int u, int& mb, int& nb, int& kb0, int& kb1) {
        mb = u % p.mo_blocks;
        nb = (u / p.mo_blocks) % p.no_blocks;
        const int ks = u / (p.mo_blocks * p.no_blocks);
        kb0 = ks * kb_per;
        kb1 = min(kb0 + kb_per, p.kb_total);
    };

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                int mb, nb, kb0, kb1;
                decode(u, mb, nb, kb0, kb1);
                const int ngroups = (min(TN_BN, p.No - nb * TN_BN) + 63) >> 6;      // 64-column groups of B that hold data
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    uint8_t* sa = sm + (size_t)stage * TN_STAGE_BYTES;
                    mbar_arrive_expect_tx(&bar_full[stage], TN_A_BYTES + ngroups * TN_GROUP_BYTES);
#pragma unroll
                    for (int g = 0; g < 2; ++g) tma_load_2d(&tmA, &bar_full[stage], sa + g * TN_GROUP_BYTES, mb * TN_BM + g * 64, kb * TN_BK);
                    for (int g = 0; g < ngroups; ++g)
                        tma_load_2d(&tmB, &bar_full[stage], sa + TN_A_BYTES + g * TN_GROUP_BYTES, nb * TN_BN + g * 64, kb * TN_BK);
                    if (++stage == TN_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint64_t tmpl = smem_desc_template(TN_GROUP_BYTES, 1024, LAYOUT_SW128);
            int stage = 0;
            uint32_t phase = 0;
            int acc = 0;
            uint32_t acc_phase = 0;
            for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
                int mb, nb, kb0, kb1;
                decode(u, mb, nb, kb0, kb1);
                mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * TN_BN);
                const uint32_t idesc = idesc_bf16(TN_BM, ((min(TN_BN, p.No - nb * TN_BN) + 63) >> 6) * 64, 1, 1);
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(XFORM ? &bar_xf[stage] : &bar_full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sm + (size_t)stage * TN_STAGE_BYTES);
                    const uint32_t b_addr = a_addr + TN_A_BYTES;
#pragma unroll
                    for (int k = 0; k < TN_BK / 16; ++k)   // 16 reduction rows = 2048 bytes per step
                        umma_bf16(d, smem_desc(tmpl, a_addr + k * 2048), smem_desc(tmpl, b_addr + k * 2048), idesc,
                                  (uint32_t)((kb != kb0) || (k != 0)));
                    umma_commit(&bar_empty[stage]);
                    if (++stage == TN_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&bar_tfull[acc]);      // (an empty K range still signals: accumulator is then skipped below)
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp < 6) {
        const int g = warp & 3;
        int acc = 0;
        uint32_t acc_phase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            int mb, nb, kb0, kb1;
            decode(u, mb, nb, kb0, kb1);
            mbar_wait(&bar_tfull[acc], acc_phase);
            tc_fence_after();
            const int row = mb * TN_BM + g * 32 + lane;
            const uint32_t taddr = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(acc * TN_BN);
            if (kb1 > kb0) {
                const int ncols_tile = ((min(TN_BN, p.No - nb * TN_BN) + 63) >> 6) * 64;
#pragma unroll 1
                for (int c0 = 0; c0 < ncols_tile; c0 += 32) {
                    uint32_t r[32];
                    tmem_ld32(taddr + c0, r);
                    tmem_ld_wait();
                    const int col = nb * TN_BN + c0;
                    if (row < p.Mo && col < p.No) {
                        float* o = p.out + (long)row * p.ldo + col;
                        if (col + 32 <= p.No && ((reinterpret_cast<uintptr_t>(o) & 15) == 0)) {
#pragma unroll
                            for (int q = 0; q < 8; ++q)
                                atomicAdd(reinterpret_cast<float4*>(o + 4 * q),
                                          make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]),
                                                      __uint_as_float(r[4 * q + 3])));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (col + j < p.No) atomicAdd(o + j, __uint_as_float(r[j]));
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
    } else if (XFORM) {
        // B tile [4 groups][64 rows][128 B]: channel n = nb*256 + g*64 + (pc ^ (row & 7))*8 + j
        const int w = warp - 6;
        int stage = 0;
        uint32_t phase = 0;
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            int mb, nb, kb0, kb1;
            decode(u, mb, nb, kb0, kb1);
            const int ngroups = (min(TN_BN, p.No - nb * TN_BN) + 63) >> 6;
            const bool active = w < ngroups;
            const int pc = lane & 7;
            // row & 7 of the rows this lane touches is ((i & 1) * 4 + (lane >> 3)): two sets of 8 constants per unit
            float cs[2][8], ct[2][8];
#pragma unroll
            for (int par = 0; par < 2; ++par) {
                const int n0 = nb * TN_BN + w * 64 + ((pc ^ (par * 4 + (lane >> 3))) << 3);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const bool ok = active && n0 + j < p.No;
                    cs[par][j] = ok ? __ldg(p.xf_scale + n0 + j) : 0.f;
                    ct[par][j] = ok ? __ldg(p.xf_shift + n0 + j) : 0.f;
                }
            }
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(&bar_full[stage], phase);
                if (active) {
                    uint8_t* sb = sm + (size_t)stage * TN_STAGE_BYTES + TN_A_BYTES + w * TN_GROUP_BYTES + (lane >> 3) * 128 + pc * 16;
                    // 64 rows-of-128B of group w; 2 batches of 8 iterations x (4 rows x 8 chunks)
#pragma unroll
                    for (int b = 0; b < 2; ++b) {
                        uint4 v[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const uint4*>(sb + (b * 8 + i) * 512);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            const int par = i & 1;
                            uint32_t* vv = reinterpret_cast<uint32_t*>(&v[i]);
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float2 f = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&vv[q]));
                                __nv_bfloat162 o = __floats2bfloat162_rn(fmaxf(fmaf(f.x, cs[par][2 * q], ct[par][2 * q]), 0.f),
                                                                          fmaxf(fmaf(f.y, cs[par][2 * q + 1], ct[par][2 * q + 1]), 0.f));
                                vv[q] = *reinterpret_cast<uint32_t*>(&o);
                            }
                            *reinterpret_cast<uint4*>(sb + (b * 8 + i) * 512) = v[i];
                        }
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_xf[stage]);
                if (++stage == TN_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// a: [Kp, lda] bf16 (Mo columns used), b: [Kp, ldb] bf16 (No columns used), out: [Mo, ldo] fp32, out += a^T op(b)
GN_API int gn_gemm_tn_bf16(const void* a, long lda, const void* b, long ldb, int Mo, int No, int Kp, float* out, long ldo,
                           const float* xf_scale, const float* xf_shift, cudaStream_t stream) {
    GN_REQUIRE(a && b && out && Mo > 0 && No > 0 && Kp > 0, GN_EINVAL, "gemm_tn_bf16: bad arguments");
    GN_REQUIRE(lda >= Mo && ldb >= No && ldo >= No, GN_EINVAL, "gemm_tn_bf16: pitch smaller than extent");
    GN_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, GN_EALIGN, "gemm_tn_bf16: lda/ldb must be multiples of 8 elements (16 bytes)");
    GN_REQUIRE((xf_scale == nullptr) == (xf_shift == nullptr), GN_EINVAL, "gemm_tn_bf16: xf_scale/xf_shift must come together");
    TnParams p;
    p.Mo = Mo; p.No = No; p.Kp = Kp;
    p.mo_blocks = gn_ceil_div(Mo, TN_BM);
    p.no_blocks = gn_ceil_div(No, TN_BN);
    p.kb_total = gn_ceil_div(Kp, TN_BK);
    const int tiles = p.mo_blocks * p.no_blocks;
    int ksplit = (2 * gn_num_sms() + tiles - 1) / tiles;
    if (ksplit > p.kb_total) ksplit = p.kb_total;
    if (ksplit < 1) ksplit = 1;
    // make every split non-empty
    const int kb_per = (p.kb_total + ksplit - 1) / ksplit;
    ksplit = (p.kb_total + kb_per - 1) / kb_per;
    p.ksplit = ksplit;
    p.out = out; p.ldo = ldo; p.xf_scale = xf_scale; p.xf_shift = xf_shift;
    CUtensorMap tmA, tmB;
    int rc = gn_tmap_bf16_2d(&tmA, a, (uint64_t)Kp, (uint64_t)Mo, (uint64_t)lda, 64, TN_BK);
    if (rc) return rc;
    rc = gn_tmap_bf16_2d(&tmB, b, (uint64_t)Kp, (uint64_t)No, (uint64_t)ldb, 64, TN_BK);
    if (rc) return rc;
    const size_t smem = (size_t)TN_STAGES * TN_STAGE_BYTES + 1024;
    const bool xform = xf_scale != nullptr;
    static bool attr_set = false;
    if (!attr_set) {
        GN_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        GN_CUDA(cudaFuncSetAttribute(gemm_tn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int units = tiles * ksplit;
    const int grid = units < gn_num_sms() ? units : gn_num_sms();
    if (xform)
        GN_CUDA(gn_launch(gemm_tn_kernel<true>, dim3(grid), dim3(320), smem, stream, tmA, tmB, p));
    else
        GN_CUDA(gn_launch(gemm_tn_kernel<false>, dim3(grid), dim3(192), smem, stream, tmA, tmB, p));
    GN_LAUNCH_CHECK();
    return GN_OK;
}
