// Weight (and bias) gradient of the hexagonal convolution (kernel_size 1, <= 32 channels, grid width <= 64) on tcgen05 -- second
// generation: reads the fp32 NCHW activations and output gradients ONCE, straight from global memory, no intermediate layout.
//
// Replaces autograd's weight gradient of hexagdly.Conv2d as used by the g network (/root/reference/gridnext/gridnet_models.py:128-148)
// at large batches.  hexconv_tc.cu (first generation) rewrote x' and dY into bf16 parity planes (two launches, 128 B read + 128 B
// written per cell each) and then streamed both planes through the weight-gradient kernel: ~770 B of traffic per cell against the
// 256 algorithmic bytes, three launches, 0.34 ms per layer at 256 arrays -- the dominant kernel of the g-only step.  Here:
//
//   * dWp[t][ci][co] = sum over cells s of dY[s, co] * x'[s + off_t(parity of s), ci] is a GEMM whose reduction index is the CELL.  In the
//     cell-major operand rows of the forward kernel (one 128-byte row per cell: [32 channels hi | 32 channels lo] bf16) the cell is the
//     row index, i.e. both operands are "MN-major" (gn_ptx.cuh) with the 64 (hi | lo) channel slots as M / N;
//   * sixteen converter warps read the fp32 rows with coalesced ld.global (lane = grid column: 128 B per warp and channel; the TMA boxes
//     of the forward kernel walk such 256-byte rows one at a time and top out at 2 TB/s), apply the previous BatchNorm+ReLU to x
//     (gridnet_models.py:134-136), split into bf16 hi + lo and write the operand rows into two rings of row PAIRS (x: 5 pairs, dY: 3);
//     loads run two jobs ahead of the conversion (24 registers per thread, ~48 KB in flight per SM).  An x row slot holds 72 cells: four
//     zero cells on either side of the 64 columns, so that the column shifts of the neighbourhood are operand START ADDRESSES;
//   * one thread issues, per grid row y (its 64 cells = 4 K steps of 16 cells) and K step, A = dY row (M = 64 channel slots, aliased
//     to 128 with LBO = 0) against
//         same row  : B = x row y   from column -1,                     N = 192 = taps (x-1, x, x+1)
//         row above : B = x row y-1 from column -1 (y even) / 0 (odd),  N = 128 = taps (a = 0, 1)
//         row below : B = x row y+1 likewise,                           N = 128
//     where the taps are stacked along N with LBO = 128 bytes: the next 64-slot group of the MN-major operand is the SAME rows one cell
//     further on.  12 MMAs per grid row instead of 28 (GRIDNEXT_B200_HEXWG2_STACK=0 issues the 28);
//   * the 128 x 448 fp32 accumulator (7 taps x 64 slots) stays in tensor memory for the life of the persistent CTA; at the end
//     hi x hi + hi x lo + lo x hi are folded and added to dWp with atomics.  The bias gradient (channel sums of dY) is accumulated by the
//     converters on the way (the first generation needed another pass over dY for it).
//
//   warp 0: MMA issuer   warps 1-16: converters (row of the pair x column half x channel quarter), then the final fold
#include "gn_common.cuh"
#include "gn_ptx.cuh"
#include <stdlib.h>

using namespace gnptx;

#define W2_XR 5                        // x ring depth in row pairs: a dY pair reads three, the converters may be two ahead
#define W2_DR 3                        // dY ring depth in row pairs
#define W2_XROW (72 * 128)             // one x grid row: 4 zero cells | 64 cells | 4 zero cells, 128 B each
#define W2_XSLOT (2 * W2_XROW)
#define W2_DROW (64 * 128)
#define W2_DSLOT (2 * W2_DROW)
#define W2_CONV_WARPS 16
#define W2_THREADS (32 + 32 * W2_CONV_WARPS)
#define W2_LA 3                        // converter look-ahead: loads of job j + 2 are in flight while job j is converted
#define W2_SMEM (W2_XR * W2_XSLOT + W2_DR * W2_DSLOT + 1024)

struct HexWg2Params {
    int B, H, W, Cin, Cout, NPA;       // NPA: row pairs per array
    int stack;
    const float* x;
    const float* dy;
    const float* in_scale;
    const float* in_shift;
    float* dwp;                        // [7][Cin][Cout] fp32, +=
    float* dbias;                      // [Cout] fp32, += (nullable)
};

// The job sequence of one CTA, generated identically by the converters and the MMA issuer: for every dY row pair g of the CTA's range
// (x pairs p-1, p at the start of the range or of an array), x pair p+1, dY pair p.
struct W2Gen {
    int g, g1, g0, b, p, NPA, sub;
    uint32_t xi, di;
    __device__ void init(int g0_, int g1_, int NPA_) {
        g = g0 = g0_; g1 = g1_; NPA = NPA_; b = g0_ / NPA_; p = g0_ - b * NPA_; sub = 0; xi = di = 0;
    }
    // type 0: x pair, 1: dY pair; returns false when the range is exhausted
    __device__ bool next(int& type, int& bb, int& pair, uint32_t& idx) {
        if (g >= g1) return false;
        const bool start = g == g0 || p == 0;
        if (sub < 2 && !start) sub = 2;
        bb = b;
        if (sub == 0) { type = 0; pair = p - 1; idx = xi++; sub = 1; }
        else if (sub == 1) { type = 0; pair = p; idx = xi++; sub = 2; }
        else if (sub == 2) { type = 0; pair = p + 1; idx = xi++; sub = 3; }
        else {
            type = 1; pair = p; idx = di++; sub = 0;
            ++g;
            if (++p == NPA) { p = 0; ++b; }
        }
        return true;
    }
};

__global__ void __launch_bounds__(W2_THREADS, 1) hexconv_wgrad_tc2_kernel(const HexWg2Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t xfull[W2_XR], xfree[W2_XR], dfull[W2_DR], dfree[W2_DR], bar_done;
    __shared__ uint32_t tmem_slot;
    __shared__ float s_pro[2][32];
    __shared__ float s_db[32];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* s_x = sm;                                   // [W2_XR][2 rows][72 cells][128 B]
    uint8_t* s_d = sm + W2_XR * W2_XSLOT;                // [W2_DR][2 rows][64 cells][128 B]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // the zero cells on either side of every x row (never written again)
    for (int i = threadIdx.x; i < W2_XR * 2 * 8 * 8; i += blockDim.x) {
        const int row = i >> 6, cell = (i >> 3) & 7, ch = i & 7;
        *reinterpret_cast<uint4*>(s_x + (size_t)row * W2_XROW + (size_t)(cell < 4 ? cell : 64 + cell) * 128 + ch * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
    if (threadIdx.x < 32) s_db[threadIdx.x] = 0.f;
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < W2_XR; ++s) { mbar_init(&xfull[s], W2_CONV_WARPS); mbar_init(&xfree[s], 1); }
        for (int s = 0; s < W2_DR; ++s) { mbar_init(&dfull[s], W2_CONV_WARPS); mbar_init(&dfree[s], 1); }
        mbar_init(&bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    gn_pdl_wait();
    if (threadIdx.x < 32) {
        s_pro[0][threadIdx.x] = (p.in_scale != nullptr && threadIdx.x < p.Cin) ? p.in_scale[threadIdx.x] : 1.f;
        s_pro[1][threadIdx.x] = (p.in_shift != nullptr && threadIdx.x < p.Cin) ? p.in_shift[threadIdx.x] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    gn_pdl_trigger();
    const uint32_t tmem_base = tmem_slot;

    // this CTA's contiguous range of dY row pairs (global pair index = array * NPA + pair)
    const long total = (long)p.B * p.NPA;
    const int g0 = (int)(total * blockIdx.x / gridDim.x), g1 = (int)(total * (blockIdx.x + 1) / gridDim.x);

    if (warp == 0) {
        // ------------------------------------------------------------------------------------------ MMA issuer
        if (elect_one() && g1 > g0) {
            const uint32_t idS = idesc_bf16(128, 192, 1, 1), idU = idesc_bf16(128, 128, 1, 1), id1 = idesc_bf16(128, 64, 1, 1);
            const uint64_t tA = smem_desc_template(0, 1024, LAYOUT_SW128);          // M = 128: the second 64-slot group aliases the first
            const uint64_t tB = smem_desc_template(128, 1024, LAYOUT_SW128);        // next 64-slot group = one cell further on
            const uint32_t x0 = smem_u32(s_x), d0 = smem_u32(s_d);
            W2Gen gen;
            gen.init(g0, g1, p.NPA);
            uint32_t xw = 0, dw = 0, started = 0;
            int type, bb, pair;
            uint32_t idx;
            while (gen.next(type, bb, pair, idx)) {
                if (type == 0) continue;
                // dY pair idx with x pairs xi-3, xi-2, xi-1
                const uint32_t xi = gen.xi;
                while (xw < xi) { mbar_wait(&xfull[xw % W2_XR], (xw / W2_XR) & 1); ++xw; }
                while (dw <= idx) { mbar_wait(&dfull[dw % W2_DR], (dw / W2_DR) & 1); ++dw; }
                tc_fence_after();
                const uint32_t xs[3] = {x0 + ((xi - 3) % W2_XR) * W2_XSLOT, x0 + ((xi - 2) % W2_XR) * W2_XSLOT, x0 + ((xi - 1) % W2_XR) * W2_XSLOT};
                const uint32_t ds = d0 + (idx % W2_DR) * W2_DSLOT;
#pragma unroll
                for (int r = 0; r < 2; ++r) {
                    // row y = 2 pair + r (parity r): above = (r ? pair p row 0 : pair p-1 row 1), below = (r ? pair p+1 row 0 : pair p row 1)
                    const uint32_t same = xs[1] + r * W2_XROW;
                    const uint32_t up = r ? xs[1] : xs[0] + W2_XROW;
                    const uint32_t dn = r ? xs[2] : xs[1] + W2_XROW;
                    const uint32_t c_ud = (r ? 4 : 3) * 128;             // first tap of the rows above / below: column x-1 (even row) / x (odd row)
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t a = smem_desc(tA, ds + r * W2_DROW + ks * 2048);
                        if (p.stack) {
                            umma_bf16(tmem_base, a, smem_desc(tB, same + 3 * 128 + ks * 2048), idS, started & 1u);
                            umma_bf16(tmem_base + 192, a, smem_desc(tB, up + c_ud + ks * 2048), idU, started & 1u);
                            umma_bf16(tmem_base + 320, a, smem_desc(tB, dn + c_ud + ks * 2048), idU, started & 1u);
                        } else {
#pragma unroll
                            for (int t = 0; t < 3; ++t) umma_bf16(tmem_base + 64 * t, a, smem_desc(tA, same + (3 + t) * 128 + ks * 2048), id1, started & 1u);
#pragma unroll
                            for (int t = 0; t < 2; ++t) {
                                umma_bf16(tmem_base + 192 + 64 * t, a, smem_desc(tA, up + c_ud + t * 128 + ks * 2048), id1, started & 1u);
                                umma_bf16(tmem_base + 320 + 64 * t, a, smem_desc(tA, dn + c_ud + t * 128 + ks * 2048), id1, started & 1u);
                            }
                        }
                        started = 1u;
                    }
                }
                umma_commit(&dfree[idx % W2_DR]);
                umma_commit(&xfree[(xi - 3) % W2_XR]);
                if (gen.g >= gen.g1 || gen.p == 0) {                     // the next dY pair starts over with its own x pairs
                    umma_commit(&xfree[(xi - 2) % W2_XR]);
                    umma_commit(&xfree[(xi - 1) % W2_XR]);
                }
            }
            umma_commit(&bar_done);
        }
        __syncwarp();
    } else {
        // ------------------------------------------------------------------------------------------ converters
        const int cw = warp - 1;                       // 0..15
        const int q = cw & 3;                          // channels 8q .. 8q+7
        const int r = (cw >> 2) & 1;                   // row of the pair
        const int x = ((cw >> 3) << 5) | lane;         // column 0..63
        const long chan = (long)p.H * p.W;
        const bool has_pro = p.in_scale != nullptr;
        float sc[8], sh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = s_pro[0][8 * q + j]; sh[j] = s_pro[1][8 * q + j]; }
        float db[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) db[j] = 0.f;

        W2Gen gen;
        gen.init(g0, g1, p.NPA);
        float v[W2_LA][8];
        int jt[W2_LA];                                  // -1: none, 0: x (cell in the grid), 1: dY, 2: x (outside the grid: zeros)
        uint32_t jidx[W2_LA];
        auto load = [&](int s) {
            int type, bb, pair;
            uint32_t idx;
            if (!gen.next(type, bb, pair, idx)) { jt[s] = -1; return; }
            const int y = 2 * pair + r;
            const int C = type ? p.Cout : p.Cin;
            const bool in = y >= 0 && y < p.H && x < p.W;
            const float* src = (type ? p.dy : p.x) + ((long)bb * C * p.H + (in ? y : 0)) * p.W + (in ? x : 0) + (long)(8 * q) * chan;
#pragma unroll
            for (int j = 0; j < 8; ++j) v[s][j] = (in && 8 * q + j < C) ? __ldg(src + j * chan) : 0.f;
            jt[s] = type ? 1 : (in ? 0 : 2);
            jidx[s] = idx;
        };
#pragma unroll
        for (int s = 0; s < W2_LA - 1; ++s) load(s);
        bool more = true;
        while (more) {
#pragma unroll
            for (int s = 0; s < W2_LA; ++s) {
                load((s + W2_LA - 1) % W2_LA);          // job j + 2 into the register set job j - 1 has left
                if (jt[s] < 0) { more = false; break; }
                const bool isd = jt[s] == 1;
                const uint32_t idx = jidx[s];
                float w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] = v[s][j];
                if (isd) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) db[j] += w[j];
                } else if (has_pro && jt[s] == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[j] = 8 * q + j < p.Cin ? fmaxf(fmaf(w[j], sc[j], sh[j]), 0.f) : 0.f;
                }
                uint32_t hi[4], lo[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const __nv_bfloat162 h = __floats2bfloat162_rn(w[2 * j], w[2 * j + 1]);
                    const float2 hf = __bfloat1622float2(h);
                    const __nv_bfloat162 l = __floats2bfloat162_rn(w[2 * j] - hf.x, w[2 * j + 1] - hf.y);
                    hi[j] = *reinterpret_cast<const uint32_t*>(&h);
                    lo[j] = *reinterpret_cast<const uint32_t*>(&l);
                }
                uint8_t* row;
                uint32_t sw;
                if (isd) {
                    const uint32_t slot = idx % W2_DR;
                    if (idx >= W2_DR) mbar_wait_backoff(&dfree[slot], ((idx / W2_DR) - 1) & 1);
                    row = s_d + (size_t)slot * W2_DSLOT + (size_t)r * W2_DROW + (size_t)x * 128;
                    sw = (uint32_t)(x & 7);
                } else {
                    const uint32_t slot = idx % W2_XR;
                    if (idx >= W2_XR) mbar_wait_backoff(&xfree[slot], ((idx / W2_XR) - 1) & 1);
                    row = s_x + (size_t)slot * W2_XSLOT + (size_t)r * W2_XROW + (size_t)(x + 4) * 128;
                    sw = (uint32_t)((x + 4) & 7);
                }
                // chunk c of the 128-byte operand row is stored at (c ^ row-in-atom) (SWIZZLE_128B; slots and rows are 1024-byte aligned)
                *reinterpret_cast<uint4*>(row + (((uint32_t)q ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(row + (((uint32_t)(4 + q) ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) mbar_arrive(isd ? &dfull[idx % W2_DR] : &xfull[idx % W2_XR]);
            }
        }
        // bias gradient: channel sums of dY
        if (p.dbias != nullptr) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float a = gn_warp_sum(db[j]);
                if (lane == 0 && 8 * q + j < p.Cout) atomicAdd(&s_db[8 * q + j], a);
            }
        }
        // ------------------------------------------------------------------------------------------ final fold (8 of the converter warps)
        // accumulator lanes 0..31 = dY_hi[co], 32..63 = dY_lo[co]; columns 64 t + (0..31) = x_hi[ci], + (32..63) = x_lo[ci]
        const int lg = warp & 3;                        // a warp may only touch the TMEM lane quarter warp % 4
        if (lg < 2 && g1 > g0) {
            const int part = (warp - 1) >> 2;           // 0..3: the warps 1, 5, 9, 13 (lane quarter 1) and 4, 8, 12, 16 (quarter 0) share the 7 taps
            mbar_wait_backoff(&bar_done, 0);
            tc_fence_after();
            const int co = lane;
            for (int t = part; t < 7; t += 4) {
                uint32_t r0[32], r1[32];
                tmem_ld32(tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(t * 64), r0);               // x x_hi[ci]
                if (lg == 0) tmem_ld32(tmem_base + (uint32_t)(t * 64 + 32), r1);                            // dY_hi x x_lo[ci]
                tmem_ld_wait();
                if (co < p.Cout) {
                    float* o = p.dwp + (long)t * p.Cin * p.Cout + co;
#pragma unroll
                    for (int ci = 0; ci < 32; ++ci) {
                        float a = __uint_as_float(r0[ci]);
                        if (lg == 0) a += __uint_as_float(r1[ci]);
                        if (ci < p.Cin) atomicAdd(o + (long)ci * p.Cout, a);
                    }
                }
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.dbias != nullptr && threadIdx.x < p.Cout) atomicAdd(p.dbias + threadIdx.x, s_db[threadIdx.x]);
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// dwp[7][cin][cout] += sum dY * x'   (x' = in_scale ? relu(x * in_scale + in_shift) : x);  dbias[cout] += sum dY (nullable).
// Same contract as gn_hexconv_wgrad (caller zeroes both); shapes as gn_hexconv_tc2_supported.
GN_API int gn_hexconv_wgrad_tc2(const float* x, const float* in_scale, const float* in_shift, const float* dy, float* dwp, float* dbias, int B,
                                int cin, int cout, int H, int W, cudaStream_t stream) {
    GN_REQUIRE(x && dy && dwp && B > 0, GN_EINVAL, "hexconv_wgrad_tc2: bad arguments");
    GN_REQUIRE(cin >= 1 && cin <= 32 && cout >= 1 && cout <= 32 && H >= 2 && W >= 1 && W <= 64, GN_EUNSUPPORTED,
               "hexconv_wgrad_tc2: needs kernel_size 1, <= 32 channels, W <= 64");
    GN_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), GN_EINVAL, "hexconv_wgrad_tc2: in_scale/in_shift must come together");
    GN_REQUIRE((long)B * ((H + 1) / 2) < (1L << 30), GN_EUNSUPPORTED, "hexconv_wgrad_tc2: batch too large");
    HexWg2Params p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = H; p.W = W; p.Cin = cin; p.Cout = cout;
    p.NPA = (H + 1) / 2;
    p.x = x; p.dy = dy; p.in_scale = in_scale; p.in_shift = in_shift; p.dwp = dwp; p.dbias = dbias;
    {
        const char* e = getenv("GRIDNEXT_B200_HEXWG2_STACK");
        p.stack = e ? atoi(e) != 0 : 1;
    }
    static bool attr_set = false;
    if (!attr_set) {
        GN_CUDA(cudaFuncSetAttribute(hexconv_wgrad_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, W2_SMEM));
        attr_set = true;
    }
    const long total = (long)B * p.NPA;
    const int grid = total < gn_num_sms() ? (int)total : gn_num_sms();
    GN_CUDA(gn_launch(hexconv_wgrad_tc2_kernel, dim3(grid), dim3(W2_THREADS), (size_t)W2_SMEM, stream, p));
    GN_LAUNCH_CHECK();
    return GN_OK;
}
