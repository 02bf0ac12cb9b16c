// DenseNet stem: 7x7 / stride 2 / pad 3 convolution of the 3-channel patch (/root/reference/gridnext/densenet.py:107,
// conv0) with norm0 + relu0 (:108-109) in the epilogue, and its weight gradient -- WITHOUT an im2col buffer.
//
// The patch is kept as NHWC4 bf16 (RGB + a zero channel = 8 bytes per pixel).  For one kernel row ky, the operand row
// of output pixel ox is the 8-pixel window starting at input x = 2*ox - 4 (one zero-weight pixel in front of the 7
// taps, so the window starts on a 16-byte boundary): 32 contiguous bf16.  Consecutive output pixels start 16 bytes
// apart, i.e. the rows of the implicit-GEMM operand OVERLAP in memory.  A no-swizzle UMMA shared-memory descriptor
// only encodes strides (16 B between the two 8-element chunks of a K step, 128 B between 8-row groups), so it reads
// that Toeplitz operand straight out of the raw image rows (validated on hardware: tools/umma_probe.py, cases
// toeplitz_*).  One TMA box brings the 2*RT+5 image rows a tile needs (out-of-bounds rows/pixels zero-filled = the
// padding); 7 ky x 2 K-steps of tcgen05.mma produce 128 "virtual" output positions v = rt*(P+8) + ox of which the
// ox < P/2 ones are real (the others are computed and dropped: the stem is 4 % of the network's MACs).
//
//   forward : D[v, co]      = sum_ky A_ky[v, 32] * Wq[ky][co, 32]^T     -> relu(D*scale+shift) -> bf16 act0 rows
//   wgrad   : dWq[ky][co,32] += dZ0[v, co]^T * A_ky[v, 32]              (both operands MN-major; 7 accumulators stay
//                                                                        in TMEM for the whole persistent CTA)
#include "gn_common.cuh"
#include "gn_ptx.cuh"
#include "gn_tma.cuh"
#include "gn_epilogue.cuh"

using namespace gnptx;

#define STEM_KQ 32                 // packed K per kernel row: 8 pixels x 4 channels
#define STEM_SUB_BYTES 16384
#define STEM_MAX_ESTAGES 4
#define STEM_EPI_WARPS 8
#define STEM_MAX_CO 128
#define STEM_SLACK 2560            // bytes the 128-row operand window may read past the last image row of a strip

struct StemParams {
    int N, P, Ho, CO, NP;
    int RT;                // output rows per tile
    int G, S;              // forward: G row-tiles share one pipeline stage (one strip box, one accumulator hand-off: the ~1 us of barrier traffic per
                           // stage dwarfed the 14 MMAs of a 64-position tile), S of them share one staging slot / TMA store
    int supers_per_img, n_super;
    int vs;                // virtual positions per output row = P + 8
    int tiles_per_img, n_tiles;
    int rows_in;           // image rows per strip = 2*RT + 5
    int pitchB;            // bytes per strip row = (P + 8) * 8
    int strip_alloc;       // bytes per strip stage (multiple of 1024, slack included)
    int stages, e_stages, nsub;
    unsigned kstep_mask;   // wgrad: which of the eight 16-position K steps contain real outputs
    const float* scale;
    const float* shift;
    int relu;
    float* dwq;            // wgrad: [CO][7*32] fp32, +=
};

// ------------------------------------------------------------------------------------------------ packing
template <typename InT>
__global__ void __launch_bounds__(256) stem_pack_input_kernel(const InT* __restrict__ x, int P, uint2* __restrict__ xq) {
    gn_pdl_sync();
    const int plane = P * P;
    const long n = blockIdx.y;
    const InT* src = x + n * 3 * plane;
    uint2* dst = xq + n * plane;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < plane; r += gridDim.x * blockDim.x) {
        uint2 o;
        o.x = gn_pack_bf16x2((float)src[r], (float)src[plane + r]);
        o.y = gn_pack_bf16x2((float)src[2 * plane + r], 0.f);
        dst[r] = o;
    }
}

// wq[ky][co][kxp*4 + c] = w[co][c][ky][kxp - 1]   (kxp = 0 and c = 3 are zero)
__global__ void stem_pack_weight_kernel(const float* __restrict__ w, int CO, __nv_bfloat16* __restrict__ wq) {
    const int total = 7 * CO * STEM_KQ;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int k = e % STEM_KQ, co = (e / STEM_KQ) % CO, ky = e / (STEM_KQ * CO);
        const int kxp = k >> 2, c = k & 3;
        float v = 0.f;
        if (kxp >= 1 && c < 3) v = w[((co * 3 + c) * 7 + ky) * 7 + (kxp - 1)];
        wq[e] = __float2bfloat16_rn(v);
    }
}
// dw[co][c][ky][kx] = dwq[co][ky*32 + (kx+1)*4 + c]
__global__ void stem_unpack_wgrad_kernel(const float* __restrict__ dwq, int CO, float* __restrict__ dw) {
    const int total = CO * 147;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int kx = e % 7, ky = (e / 7) % 7, c = (e / 49) % 3, co = e / 147;
        dw[e] = dwq[co * (7 * STEM_KQ) + ky * STEM_KQ + (kx + 1) * 4 + c];
    }
}

GN_API int gn_stem_pack_input(const void* x, int x_is_bf16, int N, int P, void* xq, cudaStream_t stream) {
    GN_REQUIRE(x && xq && N > 0 && P > 0, GN_EINVAL, "stem_pack_input: bad arguments");
    GN_REQUIRE(N <= 65535, GN_EUNSUPPORTED, "stem_pack_input: at most 65535 patches per call");
    dim3 grid((unsigned)gn_ceil_div((long)P * P, 1024), (unsigned)N);
    if (x_is_bf16) GN_CUDA(gn_launch(stem_pack_input_kernel<__nv_bfloat16>, grid, dim3(256), 0, stream, (const __nv_bfloat16*)x, P, (uint2*)xq));
    else GN_CUDA(gn_launch(stem_pack_input_kernel<float>, grid, dim3(256), 0, stream, (const float*)x, P, (uint2*)xq));
    GN_LAUNCH_CHECK();
    return GN_OK;
}
GN_API int gn_stem_pack_weight(const float* w, int CO, void* wq, cudaStream_t stream) {
    GN_REQUIRE(w && wq && CO > 0, GN_EINVAL, "stem_pack_weight: bad arguments");
    stem_pack_weight_kernel<<<gn_ceil_div(7L * CO * STEM_KQ, 256), 256, 0, stream>>>(w, CO, (__nv_bfloat16*)wq);
    GN_LAUNCH_CHECK();
    return GN_OK;
}
GN_API int gn_stem_unpack_wgrad(const float* dwq, int CO, float* dw, cudaStream_t stream) {
    GN_REQUIRE(dwq && dw && CO > 0, GN_EINVAL, "stem_unpack_wgrad: bad arguments");
    stem_unpack_wgrad_kernel<<<gn_ceil_div(147L * CO, 256), 256, 0, stream>>>(dwq, CO, dw);
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// ------------------------------------------------------------------------------------------------ forward
//   warp 0: TMA producer (one strip box per tile)   warp 1: MMA issuer (14 MMAs per tile)   warp 2: epilogue slot feeder
//   warp 3: TMEM allocator + drain (TMA row stores)  warps 4-11: epilogue (BN + ReLU -> bf16, swizzled staging tile)
__global__ void __launch_bounds__(384, 1)
stem_fwd_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmOut,
                const StemParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_w, bar_full[4], bar_empty[4], bar_tfull[2], bar_tempty[2];
    __shared__ __align__(8) uint64_t bar_efull[STEM_MAX_ESTAGES], bar_eready[STEM_MAX_ESTAGES], bar_eempty[STEM_MAX_ESTAGES];
    __shared__ uint32_t tmem_slot;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w_bytes = ((7 * p.NP * 64 + 1023) / 1024) * 1024;
    uint8_t* s_w = sm;
    uint8_t* s_x = sm + w_bytes;
    uint8_t* s_slots = s_x + (size_t)p.stages * p.strip_alloc;
    float* s_epi = reinterpret_cast<float*>(s_slots + (size_t)p.e_stages * STEM_SUB_BYTES);      // [2][STEM_MAX_CO]

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW);
        tma_prefetch_desc(&tmOut);
        mbar_init(&bar_w, 1);
        for (int s = 0; s < 4; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&bar_tfull[s], 1); mbar_init(&bar_tempty[s], STEM_EPI_WARPS); }
        for (int s = 0; s < STEM_MAX_ESTAGES; ++s) {
            mbar_init(&bar_efull[s], 1);
            mbar_init(&bar_eready[s], STEM_EPI_WARPS);
            mbar_init(&bar_eempty[s], 1);
        }
        fence_barrier_init();
    }
    if (warp == 3) tmem_alloc<512>(&tmem_slot);
    gn_pdl_wait();
    for (int i = threadIdx.x; i < STEM_MAX_CO; i += blockDim.x) {
        const bool in = i < p.CO;
        s_epi[i] = in ? (p.scale ? p.scale[i] : 1.f) : 0.f;
        s_epi[STEM_MAX_CO + i] = (in && p.shift) ? p.shift[i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    gn_pdl_trigger();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            mbar_arrive_expect_tx(&bar_w, 7 * p.NP * 64);
            for (int ky = 0; ky < 7; ++ky) tma_load_2d(&tmW, &bar_w, s_w + ky * p.NP * 64, 0, ky * p.CO);
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.n_super; tile += gridDim.x) {
                const int n = tile / p.supers_per_img, oy0 = (tile - n * p.supers_per_img) * p.G * p.RT;
                mbar_wait(&bar_empty[stage], phase ^ 1);
                mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)(p.rows_in * p.pitchB));
                tma_load_3d(&tmX, &bar_full[stage], s_x + (size_t)stage * p.strip_alloc, -4, 2 * oy0 - 3, n);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = idesc_bf16(128, p.NP, 0, 0);
            const uint64_t tmplA = smem_desc_template(16, 128, LAYOUT_NONE);      // overlapping rows: K chunks 16 B apart, 8-row groups 128 B apart
            const uint64_t tmplW = smem_desc_template(0, 512, LAYOUT_SW64);
            const uint64_t descW = smem_desc(tmplW, smem_u32(s_w));
            const int pitch16 = p.pitchB >> 4, wk16 = (p.NP * 64) >> 4;
            mbar_wait(&bar_w, 0);
            int stage = 0, acc = 0;
            uint32_t phase = 0, acc_phase = 0;
            for (int tile = blockIdx.x; tile < p.n_super; tile += gridDim.x) {
                mbar_wait(&bar_tempty[acc], acc_phase ^ 1);
                mbar_wait(&bar_full[stage], phase);
                tc_fence_after();
                for (int g = 0; g < p.G; ++g) {
                    const uint32_t d = tmem_base + (uint32_t)(acc * 256 + g * p.NP);
                    const uint64_t descA = smem_desc(tmplA, smem_u32(s_x + (size_t)stage * p.strip_alloc) + (uint32_t)(g * 2 * p.RT * p.pitchB));
#pragma unroll
                    for (int ky = 0; ky < 7; ++ky)
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks)
                            umma_bf16(d, descA + (uint64_t)(ky * pitch16 + ks * 2), descW + (uint64_t)(ky * wk16 + ks * 2), idesc, (uint32_t)((ky | ks) != 0));
                }
                umma_commit(&bar_empty[stage]);
                umma_commit(&bar_tfull[acc]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
                acc ^= 1;
                if (acc == 0) acc_phase ^= 1;
            }
        }
    } else if (warp == 2) {
        if (elect_one()) {
            int es = 0;
            uint32_t eph = 0;
            for (int tile = blockIdx.x; tile < p.n_super; tile += gridDim.x)
                for (int j = 0; j < (p.G / p.S) * p.nsub; ++j) {
                    mbar_wait(&bar_eempty[es], eph ^ 1);
                    mbar_arrive(&bar_efull[es]);
                    if (++es == p.e_stages) { es = 0; eph ^= 1; }
                }
        }
    } else if (warp == 3) {
        if (elect_one()) {
            int es = 0, prev_es = -1;
            uint32_t eph = 0;
            for (int tile = blockIdx.x; tile < p.n_super; tile += gridDim.x) {
                const int n = tile / p.supers_per_img, oy0 = (tile - n * p.supers_per_img) * p.G * p.RT;
                const int row0 = (n * p.Ho + oy0) * p.Ho;
                for (int sp = 0; sp < p.G / p.S; ++sp)
                    for (int j = 0; j < p.nsub; ++j) {
                        mbar_wait(&bar_eready[es], eph);
                        // the S * RT output rows of a slot are consecutive rows of the [N*Ho*Ho, CO] output: one box
                        tma_store_2d(&tmOut, s_slots + (size_t)es * STEM_SUB_BYTES, j * 64, row0 + sp * p.S * p.RT * p.Ho);
                        tma_store_commit();
                        tma_store_wait_read<0>();
                        mbar_arrive(&bar_eempty[es]);
                        if (++es == p.e_stages) { es = 0; eph ^= 1; }
                    }
            }
            (void)prev_es;
            tma_store_wait_all<0>();
        }
        __syncwarp();
    } else {
        const int g = warp & 3, h = (warp - 4) >> 2;
        const int trow = g * 32 + lane;
        const int rt = trow / p.vs, ox = trow - rt * p.vs;
        const bool valid = rt < p.RT && ox < p.Ho;
        const int mrow = rt * p.Ho + ox;                       // dense row of the staging tile
        const uint32_t sw = (uint32_t)(mrow & 7);
        int acc = 0, es = 0;
        uint32_t acc_phase = 0, eph = 0;
        for (int tile = blockIdx.x; tile < p.n_super; tile += gridDim.x) {
            mbar_wait(&bar_tfull[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(acc * 256);
            for (int sp = 0; sp < p.G / p.S; ++sp)
                for (int j = 0; j < p.nsub; ++j) {
                    mbar_wait(&bar_efull[es], eph);
                    const int c0 = j * 64 + h * 32;
                    __syncwarp();
                    if (c0 < p.NP) {
                        for (int si = 0; si < p.S; ++si) {
                            uint32_t r[32];
                            tmem_ld32(taddr + (uint32_t)((sp * p.S + si) * p.NP + c0), r);
                            tmem_ld_wait();
                            if (valid) {
                                uint8_t* rowp = s_slots + (size_t)es * STEM_SUB_BYTES + (si * p.RT * p.Ho + mrow) * 128;      // S > 1: RT * Ho is a multiple of 8
                                const float* cst = s_epi + c0;
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    float o[8];
#pragma unroll
                                    for (int e = 0; e < 8; ++e) {
                                        const float x = fmaf(__uint_as_float(r[8 * q + e]), cst[8 * q + e], cst[STEM_MAX_CO + 8 * q + e]);
                                        o[e] = p.relu ? fmaxf(x, 0.f) : x;
                                    }
                                    uint4 t;
                                    t.x = gn_pack_bf16x2(o[0], o[1]); t.y = gn_pack_bf16x2(o[2], o[3]);
                                    t.z = gn_pack_bf16x2(o[4], o[5]); t.w = gn_pack_bf16x2(o[6], o[7]);
                                    *reinterpret_cast<uint4*>(rowp + ((((uint32_t)(h * 4 + q)) ^ sw) << 4)) = t;
                                }
                            }
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
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 3) tmem_dealloc<512>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ weight gradient
//   warp 0: TMA producer (strip + the dZ0 rows of the tile)   warp 1: MMA issuer   warps 2-5: final accumulator drain
__global__ void __launch_bounds__(192, 1)
stem_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDz, const StemParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[4], bar_empty[4], bar_done;
    __shared__ uint32_t tmem_slot;
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int a_bytes = 2 * 128 * 128;                                   // dZ0 tile: 2 channel groups x 128 positions x 128 B
    const int stage_bytes = p.strip_alloc + p.G * a_bytes;               // G row-tiles per stage: one strip, G dZ0 tiles
    const int ngroups = (p.CO + 63) >> 6;

    // zero everything once: the dropped positions of the dZ0 tile and the slack behind each strip must read as 0 forever
    for (int i = threadIdx.x; i < p.stages * stage_bytes / 16; i += blockDim.x) reinterpret_cast<uint4*>(sm)[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async_smem();
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmDz);
        for (int s = 0; s < 4; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<256>(&tmem_slot);
    gn_pdl_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    gn_pdl_trigger();
    const uint32_t tmem_base = tmem_slot;
    const bool has_work = (int)blockIdx.x < p.n_super;

    if (warp == 0) {
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < p.n_super; tile += gridDim.x) {
                const int n = tile / p.supers_per_img, oy0 = (tile - n * p.supers_per_img) * p.G * p.RT;
                const int row0 = (n * p.Ho + oy0) * p.Ho;
                mbar_wait(&bar_empty[stage], phase ^ 1);
                uint8_t* st = sm + (size_t)stage * stage_bytes;
                mbar_arrive_expect_tx(&bar_full[stage], (uint32_t)(p.rows_in * p.pitchB + p.G * p.RT * ngroups * p.Ho * 128));
                tma_load_3d(&tmX, &bar_full[stage], st, -4, 2 * oy0 - 3, n);
                for (int g = 0; g < p.G; ++g)
                    for (int rt = 0; rt < p.RT; ++rt)
                        for (int gi = 0; gi < ngroups; ++gi)
                            tma_load_2d(&tmDz, &bar_full[stage], st + p.strip_alloc + g * a_bytes + gi * 16384 + (size_t)rt * p.vs * 128, gi * 64,
                                        row0 + (g * p.RT + rt) * p.Ho);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = idesc_bf16(128, STEM_KQ, 1, 1);
            const uint64_t tmplA = smem_desc_template(16384, 1024, LAYOUT_SW128);      // dZ0^T: MN-major, 64-channel groups 16 KB apart
            const uint64_t tmplB = smem_desc_template(128, 16, LAYOUT_NONE);           // overlapping windows: 8-position groups 128 B, 8-element groups 16 B
            const int pitch16 = p.pitchB >> 4;
            int stage = 0;
            uint32_t phase = 0;
            uint32_t started = 0;                                                      // bit ky: accumulator ky holds data
            for (int tile = blockIdx.x; tile < p.n_super; tile += gridDim.x) {
                mbar_wait(&bar_full[stage], phase);
                tc_fence_after();
                const uint32_t st = smem_u32(sm + (size_t)stage * stage_bytes);
                for (int g = 0; g < p.G; ++g) {
                    const uint64_t descB = smem_desc(tmplB, st + (uint32_t)(g * 2 * p.RT * p.pitchB)), descA = smem_desc(tmplA, st + p.strip_alloc + g * a_bytes);
#pragma unroll
                    for (int ky = 0; ky < 7; ++ky) {
#pragma unroll
                        for (int s = 0; s < 8; ++s) {
                            if (p.kstep_mask & (1u << s)) {
                                umma_bf16(tmem_base + (uint32_t)(ky * STEM_KQ), descA + (uint64_t)(s * 128), descB + (uint64_t)(ky * pitch16 + s * 16), idesc,
                                          (started >> ky) & 1u);
                                started |= 1u << ky;
                            }
                        }
                    }
                }
                umma_commit(&bar_empty[stage]);
                if (++stage == p.stages) { stage = 0; phase ^= 1; }
            }
            umma_commit(&bar_done);
        }
    } else if (has_work) {
        const int g = warp & 3;
        mbar_wait(&bar_done, 0);
        tc_fence_after();
        const int co = g * 32 + lane;
        for (int ky = 0; ky < 7; ++ky) {
            uint32_t r[32];
            __syncwarp();
            tmem_ld32(tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(ky * STEM_KQ), r);
            tmem_ld_wait();
            if (co < p.CO) {
                float* o = p.dwq + (long)co * (7 * STEM_KQ) + ky * STEM_KQ;
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    atomicAdd(reinterpret_cast<float4*>(o + 4 * q),
                              make_float4(__uint_as_float(r[4 * q]), __uint_as_float(r[4 * q + 1]), __uint_as_float(r[4 * q + 2]), __uint_as_float(r[4 * q + 3])));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc<256>(tmem_base);
}

// ------------------------------------------------------------------------------------------------ host
static int stem_geometry(StemParams& p, int N, int P, int CO) {
    GN_REQUIRE(N > 0 && P >= 8 && P % 4 == 0 && P <= 248, GN_EUNSUPPORTED, "stem: patch size %d must be a multiple of 4 in [8, 248]", P);
    GN_REQUIRE(CO > 0 && CO % 8 == 0 && CO <= STEM_MAX_CO, GN_EUNSUPPORTED, "stem: %d output channels (must be a multiple of 8, <= %d)", CO, STEM_MAX_CO);
    memset(&p, 0, sizeof(p));
    p.N = N; p.P = P; p.Ho = P / 2; p.CO = CO; p.NP = ((CO + 15) / 16) * 16;
    p.vs = P + 8;
    const int rtmax = (128 - p.Ho) / p.vs + 1;
    p.RT = 1;
    for (int r = rtmax; r >= 1; --r)
        if (p.Ho % r == 0) { p.RT = r; break; }
    p.tiles_per_img = p.Ho / p.RT;
    GN_REQUIRE((long)N * p.tiles_per_img < (1L << 31) && (long)N * p.Ho * p.Ho < (1L << 31), GN_EUNSUPPORTED, "stem: too many positions");
    p.n_tiles = N * p.tiles_per_img;
    p.rows_in = 2 * p.RT + 5;
    p.pitchB = p.vs * 8;
    p.strip_alloc = ((p.rows_in * p.pitchB + STEM_SLACK + 1023) / 1024) * 1024;
    p.nsub = (CO + 63) / 64;
    p.kstep_mask = 0;
    for (int rt = 0; rt < p.RT; ++rt)
        for (int ox = 0; ox < p.Ho; ++ox) p.kstep_mask |= 1u << ((rt * p.vs + ox) >> 4);
    return GN_OK;
}

static int stem_xmap(CUtensorMap* tm, const void* xq, const StemParams& p) {
    uint64_t dims[3] = {(uint64_t)p.P, (uint64_t)p.P, (uint64_t)p.N};
    uint64_t strides[2] = {(uint64_t)p.P * 8, (uint64_t)p.P * p.P * 8};
    uint32_t box[3] = {(uint32_t)p.vs, (uint32_t)p.rows_in, 1};
    return gn_tmap_encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, xq, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

// xq: NHWC4 bf16 patches (gn_stem_pack_input); wq: packed weights (gn_stem_pack_weight); out: [N*Ho*Ho, ldo] bf16
GN_API int gn_stem_conv_fwd(const void* xq, int N, int P, const void* wq, int CO, const float* scale, const float* shift, int relu, void* out,
                            long ldo, cudaStream_t stream) {
    GN_REQUIRE(xq && wq && out, GN_EINVAL, "stem_conv_fwd: bad arguments");
    StemParams p;
    int rc = stem_geometry(p, N, P, CO);
    if (rc) return rc;
    GN_REQUIRE(ldo >= CO && ldo % 8 == 0 && ((uintptr_t)out & 15) == 0, GN_EALIGN, "stem_conv_fwd: output pitch must be a multiple of 8 and 16-byte aligned");
    p.scale = scale; p.shift = shift; p.relu = relu;
    // row-tiles per pipeline stage: two accumulator buffers of G * NP columns in 512 TMEM columns (the second buffer starts at column 256)
    p.G = 1;
    if (!gn_env_flag("GN_STEM_G1"))
        for (int g = 4; g > 1; g >>= 1)
            if (p.tiles_per_img % g == 0 && g * p.NP <= 256) { p.G = g; break; }
    p.S = 1;
    if ((p.RT * p.Ho) % 8 == 0)
        for (int sv = p.G; sv > 1; sv >>= 1)
            if (p.G % sv == 0 && sv * p.RT * p.Ho <= 128) { p.S = sv; break; }
    p.supers_per_img = p.tiles_per_img / p.G;
    p.n_super = N * p.supers_per_img;
    p.rows_in = 2 * p.G * p.RT + 5;
    p.strip_alloc = ((p.rows_in * p.pitchB + STEM_SLACK + 1023) / 1024) * 1024;
    const int w_bytes = ((7 * p.NP * 64 + 1023) / 1024) * 1024;
    const int budget = 227 * 1024 - 1024 - 512;
    p.e_stages = 3;
    p.stages = (budget - w_bytes - p.e_stages * STEM_SUB_BYTES - 2 * STEM_MAX_CO * 4) / p.strip_alloc;
    if (p.stages > 4) p.stages = 4;
    GN_REQUIRE(p.stages >= 1, GN_EUNSUPPORTED, "stem_conv_fwd: strip does not fit shared memory");
    const size_t smem = (size_t)w_bytes + (size_t)p.stages * p.strip_alloc + (size_t)p.e_stages * STEM_SUB_BYTES + 2 * STEM_MAX_CO * 4 + 1024;
    CUtensorMap tmX, tmW, tmOut;
    rc = stem_xmap(&tmX, xq, p);
    if (rc) return rc;
    rc = gn_tmap_bf16_2d(&tmW, wq, (uint64_t)7 * CO, STEM_KQ, STEM_KQ, STEM_KQ, (uint32_t)p.NP, CU_TENSOR_MAP_SWIZZLE_64B);
    if (rc) return rc;
    rc = gn_tmap_bf16_2d(&tmOut, out, (uint64_t)N * p.Ho * p.Ho, (uint64_t)CO, (uint64_t)ldo, 64, (uint32_t)(p.S * p.RT * p.Ho));
    if (rc) return rc;
    static size_t attr_set = 0;
    if (smem > attr_set) {
        GN_CUDA(cudaFuncSetAttribute(stem_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = smem;
    }
    const int grid = p.n_super < gn_num_sms() ? p.n_super : gn_num_sms();
    GN_CUDA(gn_launch(stem_fwd_kernel, dim3(grid), dim3(384), smem, stream, tmX, tmW, tmOut, p));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// dwq[CO][7*32] (fp32) += sum over positions dz[m, co] * window(m)[ky*32 + k];  dz: [N*Ho*Ho, ldz] bf16
GN_API int gn_stem_conv_wgrad(const void* xq, int N, int P, const void* dz, long ldz, int CO, float* dwq, cudaStream_t stream) {
    GN_REQUIRE(xq && dz && dwq, GN_EINVAL, "stem_conv_wgrad: bad arguments");
    StemParams p;
    int rc = stem_geometry(p, N, P, CO);
    if (rc) return rc;
    GN_REQUIRE(ldz >= CO && ldz % 8 == 0 && ((uintptr_t)dz & 15) == 0, GN_EALIGN, "stem_conv_wgrad: gradient pitch must be a multiple of 8 and 16-byte aligned");
    p.dwq = dwq;
    p.G = 1;
    if (!gn_env_flag("GN_STEM_G1"))
        for (int g = 4; g > 1; g >>= 1)       // largest group that still leaves two pipeline stages
            if (p.tiles_per_img % g == 0 &&
                2 * ((((2 * g * p.RT + 5) * p.pitchB + STEM_SLACK + 1023) / 1024) * 1024 + g * 2 * 128 * 128) <= 227 * 1024 - 1024 - 512) { p.G = g; break; }
    p.S = 1;
    p.supers_per_img = p.tiles_per_img / p.G;
    p.n_super = N * p.supers_per_img;
    p.rows_in = 2 * p.G * p.RT + 5;
    p.strip_alloc = ((p.rows_in * p.pitchB + STEM_SLACK + 1023) / 1024) * 1024;
    const int stage_bytes = p.strip_alloc + p.G * 2 * 128 * 128;
    const int budget = 227 * 1024 - 1024 - 512;
    p.stages = budget / stage_bytes;
    if (p.stages > 4) p.stages = 4;
    GN_REQUIRE(p.stages >= 1, GN_EUNSUPPORTED, "stem_conv_wgrad: strip does not fit shared memory");
    const size_t smem = (size_t)p.stages * stage_bytes + 1024;
    CUtensorMap tmX, tmDz;
    rc = stem_xmap(&tmX, xq, p);
    if (rc) return rc;
    rc = gn_tmap_bf16_2d(&tmDz, dz, (uint64_t)N * p.Ho * p.Ho, (uint64_t)CO, (uint64_t)ldz, 64, (uint32_t)p.Ho);
    if (rc) return rc;
    static size_t attr_set = 0;
    if (smem > attr_set) {
        GN_CUDA(cudaFuncSetAttribute(stem_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = smem;
    }
    const int grid = p.n_super < gn_num_sms() ? p.n_super : gn_num_sms();
    GN_CUDA(gn_launch(stem_wgrad_kernel, dim3(grid), dim3(192), smem, stream, tmX, tmDz, p));
    GN_LAUNCH_CHECK();
    return GN_OK;
}
