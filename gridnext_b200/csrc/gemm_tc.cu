// tcgen05 / TMEM / TMA GEMM for the f networks (bf16 operands, fp32 accumulation in tensor memory).
//
//   D[M, N] = op(A)[M, K] * B[N, K]^T            A, B row-major ("K-major"), D row-major with pitch ldc
//
// used for every 1x1 convolution of DenseNet-121 in NHWC (/root/reference/gridnext/densenet.py:26-27,52-53:
// conv1 of each dense layer, the transitions), their data gradients, and the Linear layers of the count MLP
// (notebooks/Tutorial_visium_count.ipynb cell 12).
//
// These GEMMs are HBM-bound (K = 64..1024 against N = 128, or K = 128 against N = 64..992 with a read-modify-write
// output), so the kernel is built to STREAM: every global access on the data path is a TMA bulk copy.
//
// Persistent, warp-specialised, one CTA per SM:
//   warp 0       TMA producer      (A/B k-blocks: boxes of 128 x 64 / BN x 64 bf16, SWIZZLE_128B, `stages` deep ring)
//   warp 1       MMA issuer        (one elected thread, tcgen05.mma.cta_group::1.kind::f16, UMMA 128 x BN x 16)
//   warp 2       epilogue feeder   (TMA loads of the tiles the epilogue reads: the BN reference tile and, for
//                                   read-modify-write, the old output tile; `e_stages` deep ring of 128 x 64 sub-tiles)
//   warp 3       TMEM allocator + epilogue drain (one TMA store per finished sub-tile, frees the slot once read)
//   warps 4-11   epilogue          (tcgen05.ld 32x32b -> math -> swizzled st.shared, in place over the staged tile);
//                                   two warps per TMEM lane group, double-buffered TMEM accumulators so the epilogue
//                                   of tile i overlaps the MMAs of tile i+1
//   warps 12-15  operand transform (XFORM only): DenseNet's pre-activation BatchNorm+ReLU
//                                   (densenet.py:12-18: conv(relu(norm(cat)))) is applied IN PLACE to the A tile in
//                                   shared memory between TMA arrival and the MMA, so the activated tensor is never
//                                   written to HBM.  BN is eval-mode (training.py:126): x * scale[k] + shift[k].
// EPI_DIRECT keeps a register -> global epilogue for what TMA cannot express (fp32 output / accumulate, unaligned
// views, N > 1024).
#include "gn_common.cuh"
#include "gn_ptx.cuh"
#include "gn_tma.cuh"
#include "gn_epilogue.cuh"

using namespace gnptx;

#define GEMM_BM 128
#define GEMM_BK 64
#define GEMM_A_BYTES (GEMM_BM * GEMM_BK * 2)
#define GEMM_MAX_XF_K 1024
#define GEMM_EPI_MAX_N 1024
#define GEMM_SUB_BYTES 16384      // one epilogue sub-tile: 128 rows x 64 bf16 (128-byte rows, SWIZZLE_128B)
#define GEMM_MAX_STAGES 8
#define GEMM_MAX_ESTAGES 8
#define GEMM_EPI_WARPS 8
#define GEMM_EPI_THREADS (GEMM_EPI_WARPS * 32)

enum { EPI_STORE = 0, EPI_BNBWD = 1, EPI_DIRECT = 2 };

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
    int eager_drain;           // release a staging slot as soon as its own store has read it
    int res_b;                 // 1: all k-blocks of B (one n-block) stay resident in shared memory for the life of the CTA.  Re-streaming the
                               // weights with every M tile doubled the L2 -> SM traffic of the forward conv1 (K = c_in against N = 128), and
                               // L2 -> SM bandwidth is no higher than HBM bandwidth: 4.7 TB/s algorithmic at 6.0 TB/s into the SMs
    int stage_bytes;           // ring slot: A (+ B unless resident)
    int epi_ld, xf_ld;         // pitches of the per-column / per-k constant tables in shared memory: N and K rounded up to 64
    int epi_mode;              // 0: affine/ReLU store, 1: BN+ReLU backward (BnBwdEpi), bf16 output
    BnBwdEpi bn;
    int stages;                // main-loop ring depth
    int e_stages;              // epilogue sub-tile ring depth (TMA epilogues)
    int slot_bytes;            // GEMM_SUB_BYTES
};

template <int BN> struct GemmCfg {
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = GEMM_A_BYTES + B_BYTES;
    static constexpr int TMEM_COLS = (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
    static constexpr int NSUB = BN >= 64 ? BN / 64 : 1;      // 64-column epilogue sub-tiles per tile
};

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
    __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t w) { return __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w)); }

// store 32 consecutive columns of one row (EPI_DIRECT, mode 0)
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
        gn_store_bf16_32(o, v, ncols);
    }
}

// The it-th tile of this CTA -> (n-block, m-block); false when the CTA is done.  The TMA epilogues keep their column sums in
// shared memory for ALL columns, so tiles are ordered m-major (the n-blocks of one m-block are consecutive tile numbers and
// run on neighbouring CTAs at the same time): the A tile (dZ, shared by every n-block) comes from L2 instead of HBM.  The
// host picks a grid size coprime with the number of n-blocks so that every CTA sees all n-blocks (the last one is narrow).
// The register-accumulating direct epilogue flushes on every n-block change and keeps the n-major order.
template <int EPI>
__device__ __forceinline__ bool gemm_next_tile(int it, const GemmParams& p, int& nb, int& mb) {
    const int tile = blockIdx.x + it * gridDim.x;
    if (tile >= p.num_m_blocks * p.num_n_blocks) return false;
    if (EPI == EPI_DIRECT) {
        nb = tile / p.num_m_blocks;
        mb = tile - nb * p.num_m_blocks;
    } else {
        mb = tile / p.num_n_blocks;
        nb = tile - mb * p.num_n_blocks;
    }
    return true;
}

// Operand-transform warps: 4, or 8 in two groups that take alternate k-blocks (EPI_STORE: the forward conv1).  One warp per
// scheduler cannot hide the LDS -> FMA -> STS latency of a 16 KB tile inside the ~760 cycles HBM needs to deliver it; with two
// groups every group has two k-block periods per tile.
template <bool XFORM, int EPI, int XFW = 8>
struct GemmThreads { static constexpr int XF_WARPS = XFORM ? (EPI == EPI_STORE ? XFW : 4) : 0; static constexpr int N = (12 + XF_WARPS) * 32; };

// XT (operand transform into TENSOR MEMORY): the transform warps read the raw A tile from shared memory once
// (thread = tile row), apply BatchNorm+ReLU in registers and write the bf16 operand with tcgen05.st into a ring of four 32-column TMEM
// stages; the MMA takes A from there.  Against the in-place transform this drops the transform's store and the tensor core's A fetch from
// the shared-memory port (per 64-channel k-block 704 -> 512 wavefronts), which -- not HBM, not issue slots -- bounded the forward conv1.
#define GEMM_XT_STAGES 4
#define GEMM_XT_COL0 256
template <int BN, bool XFORM, int EPI, int XFW = 8, bool XT = false>
__global__ void __launch_bounds__(GemmThreads<XFORM, EPI, XFW>::N, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
                 const __grid_constant__ CUtensorMap tmRef, const GemmParams p) {
    using Cfg = GemmCfg<BN>;
    constexpr int NSUB = Cfg::NSUB;
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[GEMM_MAX_STAGES], bar_xf[GEMM_MAX_STAGES], bar_empty[GEMM_MAX_STAGES], bar_tfull[2], bar_tempty[2];
    __shared__ __align__(8) uint64_t bar_efull[GEMM_MAX_ESTAGES], bar_eready[GEMM_MAX_ESTAGES], bar_eempty[GEMM_MAX_ESTAGES];
    __shared__ __align__(8) uint64_t bar_bres, bar_tfree[GEMM_XT_STAGES];
    __shared__ uint32_t tmem_slot;

    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int STAGES = p.stages;
    const int stage_bytes = p.stage_bytes;
    uint8_t* s_bres = sm + (size_t)STAGES * stage_bytes;                                         // resident B: num_k_blocks x B_BYTES
    uint8_t* s_slots = s_bres + (p.res_b ? (size_t)p.num_k_blocks * Cfg::B_BYTES : 0);           // e_stages x slot_bytes
    const int epi_ld = p.epi_ld, xf_ld = p.xf_ld;                                                 // row pitches of the constant tables (floats)
    float* s_epi = reinterpret_cast<float*>(s_slots + (EPI == EPI_DIRECT ? 0 : (size_t)p.e_stages * p.slot_bytes));   // [4][epi_ld]
    float* s_cs = s_epi + 2 * epi_ld;                                                     // [2][epi_ld] column sums (EPI_BNBWD; reuses the p0/p1 rows)
    float* s_xf = s_epi + (EPI == EPI_DIRECT ? 0 : 4 * epi_ld);                           // [2][xf_ld]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
        if (EPI != EPI_DIRECT) tma_prefetch_desc(&tmOut);
        if (EPI == EPI_BNBWD) tma_prefetch_desc(&tmRef);
        for (int s = 0; s < GEMM_MAX_STAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_xf[s], 4);
            mbar_init(&bar_empty[s], XT ? (p.res_b ? 4 : 5) : 1);   // XT: the four transform warps release the stage (+ the MMA when it holds streamed weights)
        }
        for (int s = 0; s < GEMM_XT_STAGES; ++s) mbar_init(&bar_tfree[s], 1);
        for (int s = 0; s < GEMM_MAX_ESTAGES; ++s) {
            mbar_init(&bar_efull[s], 1);
            mbar_init(&bar_eready[s], GEMM_EPI_WARPS);
            mbar_init(&bar_eempty[s], 1);
        }
        mbar_init(&bar_bres, 1);
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bar_tfull[a], 1);
            mbar_init(&bar_tempty[a], GEMM_EPI_WARPS);
        }
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc<(XT ? 512 : Cfg::TMEM_COLS)>(&tmem_slot);
    gn_pdl_wait();      // everything above overlaps the tail of the preceding kernel; the constant tables below are global reads
    if (XFORM) {
        const int kpad = p.num_k_blocks * GEMM_BK;
        for (int i = threadIdx.x; i < kpad; i += blockDim.x) {
            const float sc = i < p.K ? p.xf_scale[i] : 0.f, sh = i < p.K ? p.xf_shift[i] : 0.f;
            if (XT) {                                   // {scale k, scale k+1, shift k, shift k+1}: one 16-byte broadcast load per channel pair
                s_xf[(i >> 1) * 4 + (i & 1)] = sc;
                s_xf[(i >> 1) * 4 + 2 + (i & 1)] = sh;
            } else {
                s_xf[i] = sc;
                s_xf[xf_ld + i] = sh;
            }
        }
    }
    if (EPI != EPI_DIRECT) {
        // per-column epilogue constants, zero beyond N (those accumulator columns are zero as well)
        const int npad = epi_ld;
        for (int i = threadIdx.x; i < npad; i += blockDim.x) {
            const bool in = i < p.N;
            if (EPI == EPI_STORE) {
                s_epi[i] = in ? (p.scale ? p.scale[i] : 1.f) : 0.f;
                s_epi[epi_ld + i] = in ? (p.shift ? p.shift[i] : 0.f) : 0.f;
            } else {
                s_epi[i] = in ? p.bn.sc[i] : 0.f;
                s_epi[epi_ld + i] = (in && p.bn.sh) ? p.bn.sh[i] : 0.f;
                s_epi[2 * epi_ld + i] = 0.f;          // s_cs: sum g
                s_epi[3 * epi_ld + i] = 0.f;          // s_cs: sum g * ref
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    gn_pdl_trigger();
    const uint32_t tmem_base = tmem_slot;
    const int n_tiles = p.num_m_blocks * p.num_n_blocks;
    const int nkb = p.num_k_blocks;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            if (p.res_b) {
                mbar_arrive_expect_tx(&bar_bres, (uint32_t)(nkb * Cfg::B_BYTES));
                for (int kb = 0; kb < nkb; ++kb) tma_load_2d(&tmB, &bar_bres, s_bres + (size_t)kb * Cfg::B_BYTES, kb * GEMM_BK, 0);
            }
            for (int it = 0;; ++it) {
                int nb, mb;
                if (!gemm_next_tile<EPI>(it, p, nb, mb)) break;
                for (int kb = 0; kb < nkb; ++kb) {
                    mbar_wait(&bar_empty[stage], phase ^ 1);
                    uint8_t* sa = sm + (size_t)stage * stage_bytes;
                    mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)stage_bytes);
                    tma_load_2d(&tmA, &bar_full[stage], sa, kb * GEMM_BK, mb * GEMM_BM);
                    if (!p.res_b) tma_load_2d(&tmB, &bar_full[stage], sa + GEMM_A_BYTES, kb * GEMM_BK, nb * BN);
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
            uint32_t kc = 0;                                   // XT: running k-block count of this CTA -> TMEM operand stage
            if (p.res_b) mbar_wait(&bar_bres, 0);
            for (int it = 0;; ++it) {
                int nb_, mb_;
                if (!gemm_next_tile<EPI>(it, p, nb_, mb_)) break;
                mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d = tmem_base + (uint32_t)(acc * BN);
                for (int kb = 0; kb < nkb; ++kb) {
                    if (XT) {
                        const int ts = (int)(kc & (GEMM_XT_STAGES - 1));
                        mbar_wait(&bar_xf[ts], (kc / GEMM_XT_STAGES) & 1);
                        if (!p.res_b) mbar_wait(&bar_full[stage], phase);      // the B half of the stage (already complete: the transform waited for it)
                        tc_fence_after();
                        const uint32_t a_t = tmem_base + (uint32_t)(GEMM_XT_COL0 + ts * 32);
                        const uint32_t b_x = p.res_b ? smem_u32(s_bres + (size_t)kb * Cfg::B_BYTES)
                                                     : smem_u32(sm + (size_t)stage * stage_bytes) + GEMM_A_BYTES;
#pragma unroll
                        for (int k = 0; k < GEMM_BK / 16; ++k)
                            umma_bf16_ts(d, a_t + (uint32_t)(k * 8), smem_desc(tmpl, b_x + k * 32), idesc, (uint32_t)((kb | k) != 0));
                        umma_commit(&bar_tfree[ts]);
                        if (!p.res_b) umma_commit(&bar_empty[stage]);          // streamed weights: the stage is free once these MMAs have read B
                        ++kc;
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                        continue;
                    }
                    mbar_wait(XFORM ? &bar_xf[stage] : &bar_full[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(sm + (size_t)stage * stage_bytes);
                    const uint32_t b_addr = p.res_b ? smem_u32(s_bres + (size_t)kb * Cfg::B_BYTES) : a_addr + GEMM_A_BYTES;
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
    } else if (warp == 2) {
        // ===================== epilogue feeder =====================
        if (EPI != EPI_DIRECT && elect_one()) {
            int es = 0;
            uint32_t eph = 0;
            for (int it = 0;; ++it) {
                int nb, mb;
                if (!gemm_next_tile<EPI>(it, p, nb, mb)) break;
                const int nsub = (min(BN, p.N - nb * BN) + 63) >> 6;
                for (int j = 0; j < nsub; ++j) {
                    mbar_wait(&bar_eempty[es], eph ^ 1);
                    if (EPI == EPI_BNBWD) {
                        uint8_t* slot = s_slots + (size_t)es * p.slot_bytes;
                        mbar_arrive_expect_tx(&bar_efull[es], (uint32_t)p.slot_bytes);
                        tma_load_2d(&tmRef, &bar_efull[es], slot, nb * BN + j * 64, mb * GEMM_BM);
                    } else {
                        mbar_arrive(&bar_efull[es]);
                    }
                    if (++es == p.e_stages) { es = 0; eph ^= 1; }
                }
            }
        }
    } else if (warp == 3) {
        // ===================== epilogue drain: one TMA store per finished sub-tile =====================
        if (EPI != EPI_DIRECT && elect_one()) {
            int es = 0, prev_es = -1;
            uint32_t eph = 0;
            for (int it = 0;; ++it) {
                int nb, mb;
                if (!gemm_next_tile<EPI>(it, p, nb, mb)) break;
                const int nsub = (min(BN, p.N - nb * BN) + 63) >> 6;
                for (int j = 0; j < nsub; ++j) {
                    mbar_wait(&bar_eready[es], eph);
                    if (EPI == EPI_BNBWD && p.bn.rmw) tma_reduce_add_2d(&tmOut, s_slots + (size_t)es * p.slot_bytes, nb * BN + j * 64, mb * GEMM_BM);
                    else tma_store_2d(&tmOut, s_slots + (size_t)es * p.slot_bytes, nb * BN + j * 64, mb * GEMM_BM);
                    tma_store_commit();
                    if (p.eager_drain) {
                        tma_store_wait_read<0>();
                        mbar_arrive(&bar_eempty[es]);
                    } else {
                        if (prev_es >= 0) {
                            tma_store_wait_read<1>();          // the previous sub-tile's store has drained its slot
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
    } else if (warp >= 4 && warp < 4 + GEMM_EPI_WARPS) {
        // ===================== epilogue =====================
        const int g = warp & 3;            // TMEM lane group this warp may access
        const int h = (warp - 4) >> 2;     // which 32-column half of every 64-column sub-tile
        int acc = 0;
        uint32_t acc_phase = 0;
        float cs_g[NSUB], cs_x[NSUB];      // BN-backward column partial sums (column = nb*BN + j*64 + h*32 + lane)
#pragma unroll
        for (int i = 0; i < NSUB; ++i) { cs_g[i] = 0.f; cs_x[i] = 0.f; }
        int cur_nb = -1;
        const bool want_sums = (EPI == EPI_BNBWD || (EPI == EPI_DIRECT && p.epi_mode == 1)) && p.bn.colsum != nullptr;
        auto flush_colsums = [&]() {
            if (EPI == EPI_DIRECT && want_sums && cur_nb >= 0) {
#pragma unroll
                for (int j = 0; j < NSUB; ++j) {
                    const int col = cur_nb * BN + j * 64 + h * 32 + lane;
                    if (col < p.N) {
                        atomicAdd(p.bn.colsum + col, cs_g[j]);
                        atomicAdd(p.bn.colsum + p.bn.ldsum + col, cs_x[j]);
                    }
                    cs_g[j] = 0.f; cs_x[j] = 0.f;
                }
            }
        };
        int es = 0;
        uint32_t eph = 0;
        const bool active = (BN >= 64) || h == 0;       // BN = 32: the upper half-warps have no accumulator columns
        const int trow = g * 32 + lane;                 // row of the tile owned by this thread
        const uint32_t sw = (uint32_t)(trow & 7);
        for (int it = 0;; ++it) {
            int nb, mb;
            if (!gemm_next_tile<EPI>(it, p, nb, mb)) break;
            if (nb != cur_nb) { flush_colsums(); cur_nb = nb; }
            mbar_wait(&bar_tfull[acc], acc_phase);
            tc_fence_after();
            const int row = mb * GEMM_BM + trow;
            const uint32_t taddr = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(acc * BN);
            if (EPI == EPI_DIRECT) {
#pragma unroll
                for (int j = 0; j < NSUB; ++j) {
                    const int c0 = j * 64 + h * 32;
                    if (c0 < BN) {
                        uint32_t r[32];
                        __syncwarp();
                        tmem_ld32(taddr + c0, r);
                        tmem_ld_wait();
                        const int col = nb * BN + c0;
                        if (p.epi_mode == 0) {
                            epi_store_row(p, row, col, r);
                        } else {
                            const int ncols = min(32, p.N - col);
                            float v[32], gx[32];
                            if (row < p.M && ncols > 0) {
                                float ref[32];
                                gn_load_bf16_32(p.bn.ref + (long)row * p.bn.ldref + col, ref, ncols);
                                __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (long)row * p.ldc + col;
                                float old[32];
                                if (p.bn.rmw) gn_load_bf16_32(o, old, ncols);
#pragma unroll
                                for (int e = 0; e < 32; ++e) {
                                    if (e < ncols) {
                                        const float sc = __ldg(p.bn.sc + col + e);
                                        const float a = p.bn.ref_is_raw ? fmaf(ref[e], sc, __ldg(p.bn.sh + col + e)) : ref[e];
                                        const float gg = a > 0.f ? __uint_as_float(r[e]) : 0.f;
                                        gx[e] = gg * (ref[e] - __ldg(p.bn.p0 + col + e)) * __ldg(p.bn.p1 + col + e);
                                        v[e] = gg;
                                        ref[e] = p.bn.rmw ? fmaf(gg, sc, old[e]) : gg * sc;
                                    } else {
                                        gx[e] = 0.f; v[e] = 0.f; ref[e] = 0.f;
                                    }
                                }
                                gn_store_bf16_32(o, ref, ncols);
                            } else {
#pragma unroll
                                for (int e = 0; e < 32; ++e) { v[e] = 0.f; gx[e] = 0.f; }
                            }
                            if (want_sums) {
                                cs_g[j] += gn_warp_colsum32(v, lane);
                                cs_x[j] += gn_warp_colsum32(gx, lane);
                            }
                        }
                    }
                }
            } else {
                const int nsub = (min(BN, p.N - nb * BN) + 63) >> 6;
#pragma unroll 1
                for (int j = 0; j < nsub; ++j) {
                    mbar_wait(&bar_efull[es], eph);
                    uint8_t* row_p = s_slots + (size_t)es * p.slot_bytes + trow * 128;      // results replace the staged tile in place
                    __syncwarp();
                    if (active) {
                        uint32_t r[32];
                        tmem_ld32(taddr + j * 64 + h * 32, r);
                        tmem_ld_wait();
                        const int col0 = nb * BN + j * 64 + h * 32;
                        const float* cst = s_epi + col0;
                        float v[32], gx[32];
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const uint32_t off = (((uint32_t)(h * 4 + q)) ^ sw) << 4;
                            const float4 c0a = *reinterpret_cast<const float4*>(cst + 8 * q), c0b = *reinterpret_cast<const float4*>(cst + 8 * q + 4);
                            const float4 c1a = *reinterpret_cast<const float4*>(cst + epi_ld + 8 * q),
                                         c1b = *reinterpret_cast<const float4*>(cst + epi_ld + 8 * q + 4);
                            const float k0[8] = {c0a.x, c0a.y, c0a.z, c0a.w, c0b.x, c0b.y, c0b.z, c0b.w};     // scale
                            const float k1[8] = {c1a.x, c1a.y, c1a.z, c1a.w, c1b.x, c1b.y, c1b.z, c1b.w};     // shift
                            uint32_t res[4];
                            if (EPI == EPI_STORE) {
#pragma unroll
                                for (int e2 = 0; e2 < 4; ++e2) {
                                    const uint64_t x = ffma2(f32x2(__uint_as_float(r[8 * q + 2 * e2]), __uint_as_float(r[8 * q + 2 * e2 + 1])),
                                                             f32x2(k0[2 * e2], k0[2 * e2 + 1]), f32x2(k1[2 * e2], k1[2 * e2 + 1]));
                                    res[e2] = p.relu ? f32x2_to_bf16x2_relu(x) : f32x2_to_bf16x2(x);
                                }
                            } else {
                                const bool is_raw = p.bn.ref_is_raw != 0;
                                const uint4 rv = *reinterpret_cast<const uint4*>(row_p + off);
                                const uint32_t rw[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
                                for (int e2 = 0; e2 < 4; ++e2) {
                                    const float2 rf = unpack_bf16x2(rw[e2]);
                                    float o2[2];
#pragma unroll
                                    for (int u = 0; u < 2; ++u) {
                                        const int e = 8 * q + 2 * e2 + u;
                                        const float ref = u ? rf.y : rf.x;
                                        const float sc = k0[2 * e2 + u];
                                        const float a = is_raw ? fmaf(ref, sc, k1[2 * e2 + u]) : ref;
                                        const float gg = a > 0.f ? __uint_as_float(r[e]) : 0.f;
                                        gx[e] = gg * ref;             // sum g*xhat = p1 * (sum g*ref - p0 * sum g), applied at the flush
                                        v[e] = gg;
                                        o2[u] = gg * sc;              // += (read-modify-write) is done by the TMA reduction in L2
                                    }
                                    res[e2] = pack_bf16x2(o2[0], o2[1]);
                                }
                            }
                            *reinterpret_cast<uint4*>(row_p + off) = make_uint4(res[0], res[1], res[2], res[3]);
                        }
                        if (EPI == EPI_BNBWD && want_sums) {
                            const float sg = gn_warp_colsum32(v, lane), sx = gn_warp_colsum32(gx, lane);
                            atomicAdd(&s_cs[col0 + lane], sg);
                            atomicAdd(&s_cs[epi_ld + col0 + lane], sx);
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bar_eready[es]);
                    if (++es == p.e_stages) { es = 0; eph ^= 1; }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_tempty[acc]);
            acc ^= 1;
            if (acc == 0) acc_phase ^= 1;
        }
        flush_colsums();
        if (EPI == EPI_BNBWD && want_sums) {
            named_bar_sync(1, GEMM_EPI_THREADS);            // every epilogue warp has added its last partial sums
            for (int col = threadIdx.x - 4 * 32; col < p.N; col += GEMM_EPI_THREADS) {
                const float sg = s_cs[col], sx = s_cs[epi_ld + col];
                atomicAdd(p.bn.colsum + col, sg);
                atomicAdd(p.bn.colsum + p.bn.ldsum + col, __ldg(p.bn.p1 + col) * (sx - __ldg(p.bn.p0 + col) * sg));
            }
        }
    } else if (XFORM && warp >= 12) {
        // ===================== operand transform: A <- relu(A * scale[k] + shift[k]) in place =====================
        const int w = (warp - 12) & 3;
        const int grp = (warp - 12) >> 2;                           // which of the alternating k-block groups this warp belongs to
        constexpr int NGRP = GemmThreads<XFORM, EPI, XFW>::XF_WARPS / 4;
        int stage = 0, cnt = 0;
        uint32_t phase = 0;
        uint32_t kc = 0;
        for (int it = 0;; ++it) {
            int nb_, mb_;
            if (!gemm_next_tile<EPI>(it, p, nb_, mb_)) break;
            for (int kb = 0; kb < nkb; ++kb, cnt = (cnt + 1 == NGRP ? 0 : cnt + 1), ++kc) {
                if (NGRP > 1 && cnt != grp) {
                    // the other group's k-block: still wait for its data.  A parity wait tells phases apart only modulo 2, so a group must
                    // never reach use u of a stage before use u-1 (the other group's, when the ring depth is odd) has landed -- with an odd
                    // ring the XT variant failed intermittently (launch failure after a stage was read and released one phase early)
                    mbar_wait(&bar_full[stage], phase);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    continue;
                }
                if (XT) {
                    // thread = tile row (TMEM lane 32w + lane): 8 swizzled 16-byte chunks of the raw row -> BatchNorm + ReLU -> 32 bf16 pairs -> TMEM
                    const int ts = (int)(kc & (GEMM_XT_STAGES - 1));
                    const int row = w * 32 + lane;
                    const uint32_t swz = (uint32_t)(row & 7);
                    mbar_wait(&bar_full[stage], phase);
                    const uint8_t* rowp = sm + (size_t)stage * stage_bytes + row * 128;
                    uint4 v[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) v[c] = *reinterpret_cast<const uint4*>(rowp + (((uint32_t)c ^ swz) << 4));
                    uint32_t r[32];
                    const float4* cst4 = reinterpret_cast<const float4*>(s_xf) + kb * (GEMM_BK / 2);
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        const uint32_t wv[4] = {v[c].x, v[c].y, v[c].z, v[c].w};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float4 k4 = cst4[c * 4 + q];
                            r[c * 4 + q] = f32x2_to_bf16x2_relu(ffma2(bf16x2_to_f32x2(wv[q]), f32x2(k4.x, k4.y), f32x2(k4.z, k4.w)));
                        }
                    }
                    mbar_wait(&bar_tfree[ts], ((kc / GEMM_XT_STAGES) & 1) ^ 1);      // the MMAs that read this TMEM stage last time are done
                    tc_fence_after();
                    tmem_st32(tmem_base + ((uint32_t)(w * 32) << 16) + (uint32_t)(GEMM_XT_COL0 + ts * 32), r);
                    tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&bar_xf[ts]);
                        mbar_arrive(&bar_empty[stage]);                              // the raw tile is in registers / TMEM: the TMA may refill the stage
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    continue;
                }
                // row & 7 of the rows this lane touches is ((i & 1) * 4 + (lane >> 3)): two sets of 8 constants per k-block
                const int pc = lane & 7;                            // physical 16-byte chunk in the 128-byte row
                uint64_t cs[2][4], ct[2][4];                        // packed fp32 pairs: one FFMA2 + one F2FP.RELU per bf16 pair
#pragma unroll
                for (int par = 0; par < 2; ++par) {
                    const int k0 = kb * GEMM_BK + ((pc ^ (par * 4 + (lane >> 3))) << 3);   // logical K index of the chunk's first element
                    const float4 s0 = *reinterpret_cast<const float4*>(s_xf + k0);
                    const float4 s1 = *reinterpret_cast<const float4*>(s_xf + k0 + 4);
                    const float4 t0 = *reinterpret_cast<const float4*>(s_xf + xf_ld + k0);
                    const float4 t1 = *reinterpret_cast<const float4*>(s_xf + xf_ld + k0 + 4);
                    cs[par][0] = f32x2(s0.x, s0.y); cs[par][1] = f32x2(s0.z, s0.w); cs[par][2] = f32x2(s1.x, s1.y); cs[par][3] = f32x2(s1.z, s1.w);
                    ct[par][0] = f32x2(t0.x, t0.y); ct[par][1] = f32x2(t0.z, t0.w); ct[par][2] = f32x2(t1.x, t1.y); ct[par][3] = f32x2(t1.z, t1.w);
                }
                mbar_wait(&bar_full[stage], phase);
                uint8_t* sa = sm + (size_t)stage * stage_bytes + (w * 32 + (lane >> 3)) * 128 + pc * 16;
                uint4 v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = *reinterpret_cast<const uint4*>(sa + i * 512);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const int par = i & 1;
                    v[i].x = f32x2_to_bf16x2_relu(ffma2(bf16x2_to_f32x2(v[i].x), cs[par][0], ct[par][0]));
                    v[i].y = f32x2_to_bf16x2_relu(ffma2(bf16x2_to_f32x2(v[i].y), cs[par][1], ct[par][1]));
                    v[i].z = f32x2_to_bf16x2_relu(ffma2(bf16x2_to_f32x2(v[i].z), cs[par][2], ct[par][2]));
                    v[i].w = f32x2_to_bf16x2_relu(ffma2(bf16x2_to_f32x2(v[i].w), cs[par][3], ct[par][3]));
                    *reinterpret_cast<uint4*>(sa + i * 512) = v[i];
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
    if (warp == 3) tmem_dealloc<(XT ? 512 : Cfg::TMEM_COLS)>(tmem_base);
}

template <int BN, bool XFORM, int EPI, int XFW, bool XT>
static int launch_gemm_kernel(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmRef, const GemmParams& p,
                              int grid, size_t smem, cudaStream_t stream) {
    static size_t attr_set = 0;
    if (smem > attr_set) {
        GN_CUDA(cudaFuncSetAttribute(gemm_bf16_kernel<BN, XFORM, EPI, XFW, XT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = smem;
    }
    GN_CUDA(gn_launch(gemm_bf16_kernel<BN, XFORM, EPI, XFW, XT>, dim3(grid), dim3(GemmThreads<XFORM, EPI, XFW>::N), smem, stream, tmA, tmB, tmOut, tmRef, p));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

template <int BN, bool XFORM, int EPI, int XFW = 8>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmRef, GemmParams& p,
                       cudaStream_t stream) {
    using Cfg = GemmCfg<BN>;
    const int budget = 227 * 1024 - 1024 - 512;        // dynamic shared memory minus alignment slack and static barriers
    p.xf_ld = p.num_k_blocks * GEMM_BK;
    p.epi_ld = p.num_n_blocks * BN < GEMM_EPI_MAX_N ? p.num_n_blocks * BN : GEMM_EPI_MAX_N;      // every column a tile can touch
    int fixed = XFORM ? 2 * p.xf_ld * 4 : 0;
    if (EPI != EPI_DIRECT) {
        p.slot_bytes = GEMM_SUB_BYTES;       // read-modify-write goes through the TMA reduction: no old-output tile in shared memory
        p.e_stages = EPI == EPI_BNBWD ? 6 : 4;
        { const char* e = getenv("GN_GEMM_ESTAGES"); if (e && atoi(e) >= 2 && atoi(e) <= GEMM_MAX_ESTAGES) p.e_stages = atoi(e); }
        fixed += 4 * p.epi_ld * 4 + p.e_stages * p.slot_bytes;
    } else {
        p.slot_bytes = 0;
        p.e_stages = 1;
    }
    p.res_b = 0;
    p.stage_bytes = Cfg::STAGE_BYTES;
    const int res_bytes = p.num_k_blocks * Cfg::B_BYTES;
    if (p.num_n_blocks == 1 && p.num_m_blocks > 2 * gn_num_sms() && budget - fixed - res_bytes >= 4 * GEMM_A_BYTES && !gn_env_flag("GN_GEMM_NO_RESB")) {
        p.res_b = 1;
        p.stage_bytes = GEMM_A_BYTES;
        fixed += res_bytes;
    }
    int stages = (budget - fixed) / p.stage_bytes;
    if (stages > GEMM_MAX_STAGES) stages = GEMM_MAX_STAGES;
    if (!p.res_b && stages > 2 * p.num_k_blocks && stages > 2) stages = 2 * p.num_k_blocks > 2 ? 2 * p.num_k_blocks : 2;
    GN_REQUIRE(stages >= 2, GN_EUNSUPPORTED, "gemm_bf16: shared memory budget exhausted (BN %d)", BN);
    p.stages = stages;
    const size_t smem = (size_t)stages * p.stage_bytes + fixed + 1024;
    const int tiles = p.num_m_blocks * p.num_n_blocks;
    int grid = tiles < gn_num_sms() ? tiles : gn_num_sms();
    if (EPI != EPI_DIRECT && p.num_n_blocks > 1 && grid > 1)
        while (grid % 2 == 0 && p.num_n_blocks % 2 == 0 || grid % 3 == 0 && p.num_n_blocks % 3 == 0) --grid;      // coprime with the n-block count
    if constexpr (XFORM && EPI == EPI_STORE && BN == 128) {
        // the forward conv1 of a dense layer (N = 128 output channels, resident weights): BatchNorm+ReLU operand transform into tensor memory
        static const bool xt_on = !gn_env_flag("GN_GEMM_NO_XT");
        if (xt_on) return launch_gemm_kernel<BN, XFORM, EPI, XFW, true>(tmA, tmB, tmOut, tmRef, p, grid, smem, stream);
    }
    return launch_gemm_kernel<BN, XFORM, EPI, XFW, false>(tmA, tmB, tmOut, tmRef, p, grid, smem, stream);
}

template <int BN>
static int dispatch_gemm(bool xform, int epi, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmOut, const CUtensorMap& tmRef,
                         GemmParams& p, cudaStream_t stream) {
    if (epi == EPI_BNBWD) return launch_gemm<BN, false, EPI_BNBWD>(tmA, tmB, tmOut, tmRef, p, stream);
    if (epi == EPI_STORE) {
        if (!xform) return launch_gemm<BN, false, EPI_STORE>(tmA, tmB, tmOut, tmRef, p, stream);
        return launch_gemm<BN, true, EPI_STORE, 8>(tmA, tmB, tmOut, tmRef, p, stream);
    }
    return xform ? launch_gemm<BN, true, EPI_DIRECT>(tmA, tmB, tmOut, tmRef, p, stream)
                 : launch_gemm<BN, false, EPI_DIRECT>(tmA, tmB, tmOut, tmRef, p, stream);
}

static inline bool tma_ok(const void* ptr, long ld_elems) { return ((reinterpret_cast<uintptr_t>(ptr) & 15) == 0) && (ld_elems % 8 == 0); }

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
    memset(&p, 0, sizeof(p));
    p.M = M; p.N = N; p.K = K;
    p.num_m_blocks = gn_ceil_div(M, GEMM_BM);
    p.num_n_blocks = gn_ceil_div(N, BN);
    p.num_k_blocks = gn_ceil_div(K, GEMM_BK);
    p.out = out; p.ldc = ldc; p.out_fp32 = out_fp32; p.accumulate = accumulate;
    p.scale = scale; p.shift = shift; p.relu = relu; p.xf_scale = xf_scale; p.xf_shift = xf_shift;
    p.eager_drain = gn_env_flag("GN_GEMM_LAZY_DRAIN") ? 0 : 1;
    p.epi_mode = bn_ref != nullptr ? 1 : 0;
    if (p.epi_mode == 1) {
        GN_REQUIRE(!out_fp32 && !accumulate && !scale && !shift && !relu, GN_EINVAL, "gemm_bf16: BN-backward epilogue needs a plain bf16 output");
        GN_REQUIRE(!xform, GN_EUNSUPPORTED, "gemm_bf16: BN-backward epilogue cannot be combined with the operand transform");
        GN_REQUIRE(bn_sc && bn_p0 && bn_p1 && (!bn_ref_is_raw || bn_sh), GN_EINVAL, "gemm_bf16: incomplete BN-backward epilogue arguments");
        p.bn.ref = (const __nv_bfloat16*)bn_ref; p.bn.ldref = bn_ldref; p.bn.ref_is_raw = bn_ref_is_raw;
        p.bn.sc = bn_sc; p.bn.sh = bn_sh; p.bn.p0 = bn_p0; p.bn.p1 = bn_p1; p.bn.colsum = bn_colsum; p.bn.ldsum = bn_ldsum; p.bn.rmw = bn_rmw;
    }
    // TMA-staged epilogue whenever the output (and reference) views can be described by a tensor map
    int epi = EPI_DIRECT;
    if (!out_fp32 && N <= GEMM_EPI_MAX_N && tma_ok(out, ldc) && (p.epi_mode == 0 || tma_ok(bn_ref, bn_ldref)))
        epi = p.epi_mode == 1 ? EPI_BNBWD : EPI_STORE;
    if (epi == EPI_STORE && gn_env_flag("GN_GEMM_FORCE_DIRECT")) epi = EPI_DIRECT;
    CUtensorMap tmA, tmB, tmOut, tmRef;
    memset(&tmOut, 0, sizeof(tmOut));
    memset(&tmRef, 0, sizeof(tmRef));
    int rc = gn_tmap_bf16_2d(&tmA, a, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BK, GEMM_BM);
    if (rc) return rc;
    rc = gn_tmap_bf16_2d(&tmB, b, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, GEMM_BK, (uint32_t)BN);
    if (rc) return rc;
    if (epi != EPI_DIRECT) {
        rc = gn_tmap_bf16_2d(&tmOut, out, (uint64_t)M, (uint64_t)N, (uint64_t)ldc, 64, GEMM_BM);
        if (rc) return rc;
        if (epi == EPI_BNBWD) {
            rc = gn_tmap_bf16_2d(&tmRef, bn_ref, (uint64_t)M, (uint64_t)N, (uint64_t)bn_ldref, 64, GEMM_BM);
            if (rc) return rc;
        }
    }
    switch (BN) {
        case 32: return dispatch_gemm<32>(xform, epi, tmA, tmB, tmOut, tmRef, p, stream);
        case 64: return dispatch_gemm<64>(xform, epi, tmA, tmB, tmOut, tmRef, p, stream);
        case 128: return dispatch_gemm<128>(xform, epi, tmA, tmB, tmOut, tmRef, p, stream);
        case 256: return dispatch_gemm<256>(xform, epi, tmA, tmB, tmOut, tmRef, p, stream);
    }
    return GN_EINVAL;
}
