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
//   * two producer warps bring the fp32 rows into a staging ring with 16-byte asynchronous copies (a job = one row pair of one tensor = per
//     channel 2 W contiguous floats = one warp instruction; four jobs = 64 KB in flight per SM); sixteen converter warps apply the previous
//     BatchNorm+ReLU to x (gridnet_models.py:134-136), split into bf16 hi + lo and write the operand rows into two rings of row PAIRS
//     (x: 5 pairs, dY: 3).  An x row slot holds 72 cells: four zero cells on either side of the 64 columns, so that the column shifts
//     of the neighbourhood are operand START ADDRESSES;
//   * one thread issues, per grid row y (its 64 cells = 4 K steps of 16 cells) and K step, A = dY row (M = 64 channel slots; the first
//     version aliased them to M = 128 with LBO = 0 and kept the tensor pipe 62 % busy with half of it wasted) against
//         same row  : B = x row y   from column -1,                     N = 192 = taps (x-1, x, x+1)
//         row above : B = x row y-1 from column -1 (y even) / 0 (odd),  N = 128 = taps (a = 0, 1)
//         row below : B = x row y+1 likewise,                           N = 128
//     where the taps are stacked along N with LBO = 128 bytes: the next 64-slot group of the MN-major operand is the SAME rows one cell
//     further on.  12 MMAs per grid row instead of 28 (GRIDNEXT_B200_HEXWG2_STACK=0 issues the 28);
//   * the 64 x 448 fp32 accumulator (7 taps x 64 slots) stays in tensor memory for the life of the persistent CTA; at the end
//     hi x hi + hi x lo + lo x hi are folded and added to dWp with atomics.  The bias gradient (channel sums of dY) is one more MMA per
//     K step against an all-ones operand (accumulator column 448; the first generation needed another pass over dY for it).
//
//   warp 0: MMA issuer   warps 1-2: copy producers   warps 3-18: converters (channel quarter x row of the pair x column half), then the final fold
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
#define W2_PROD_WARPS 2
#define W2_THREADS (32 + 32 * W2_PROD_WARPS + 32 * W2_CONV_WARPS)
#define W2_NSTG 4                      // fp32 staging ring: jobs (row pairs of one tensor) in flight
#define W2_STAGE 16384                 // 32 channels x 2 rows x 64 columns fp32
#define W2_ONES 2048                    // the all-ones operand of the bias gradient: 16 cells x 128 B
#define W2_SMEM (W2_XR * W2_XSLOT + W2_DR * W2_DSLOT + W2_ONES + W2_NSTG * W2_STAGE + 1024)

struct HexWg2Params {
    int B, H, W, Cin, Cout, NPA;       // NPA: row pairs per array
    int stack;
    int dbg;                           // development (GRIDNEXT_B200_HEXWG2_DBG): 1 no MMAs, 2 no conversion, 4 no copies
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
    __shared__ __align__(8) uint64_t xfull[W2_XR], xfree[W2_XR], dfull[W2_DR], dfree[W2_DR], stg_full[W2_NSTG], stg_free[W2_NSTG], bar_done;
    __shared__ uint32_t tmem_slot;
    __shared__ float2 s_pro[32];                        // {scale, shift} of the BatchNorm+ReLU in front of x
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* s_x = sm;                                   // [W2_XR][2 rows][72 cells][128 B]
    uint8_t* s_d = sm + W2_XR * W2_XSLOT;                // [W2_DR][2 rows][64 cells][128 B]
    uint8_t* s_one = s_d + W2_DR * W2_DSLOT;             // [16 cells][128 B]: slot 0 = 1.0, the rest 0 -- dY x ones = the bias gradient
    uint8_t* s_stg = s_one + W2_ONES;                    // [W2_NSTG][32 channels][2 rows][W] fp32
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // the zero cells on either side of every x row (never written again)
    for (int i = threadIdx.x; i < W2_XR * 2 * 8 * 8; i += blockDim.x) {
        const int row = i >> 6, cell = (i >> 3) & 7, ch = i & 7;
        *reinterpret_cast<uint4*>(s_x + (size_t)row * W2_XROW + (size_t)(cell < 4 ? cell : 64 + cell) * 128 + ch * 16) = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = threadIdx.x; i < W2_ONES / 16; i += blockDim.x) {
        const int k = i >> 3, ch = i & 7;                // cell k, 16-byte chunk ch: element 0 of the row sits in chunk (0 ^ (k & 7))
        *reinterpret_cast<uint4*>(s_one + i * 16) = make_uint4(ch == (k & 7) ? 0x3F80u : 0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < W2_XR; ++s) { mbar_init(&xfull[s], W2_CONV_WARPS); mbar_init(&xfree[s], 1); }
        for (int s = 0; s < W2_DR; ++s) { mbar_init(&dfull[s], W2_CONV_WARPS); mbar_init(&dfree[s], 1); }
        for (int s = 0; s < W2_NSTG; ++s) { mbar_init(&stg_full[s], 32 * W2_PROD_WARPS); mbar_init(&stg_free[s], W2_CONV_WARPS); }
        mbar_init(&bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc<512>(&tmem_slot);
    gn_pdl_wait();
    if (threadIdx.x < 32)
        s_pro[threadIdx.x] = make_float2((p.in_scale != nullptr && threadIdx.x < p.Cin) ? p.in_scale[threadIdx.x] : 1.f,
                                         (p.in_shift != nullptr && threadIdx.x < p.Cin) ? p.in_shift[threadIdx.x] : 0.f);
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
            const uint32_t idS = idesc_bf16(64, 192, 1, 1), idU = idesc_bf16(64, 128, 1, 1), id1 = idesc_bf16(64, 64, 1, 1);
            const uint64_t tA = smem_desc_template(0, 1024, LAYOUT_SW128);          // M = 64: one group of 64 channel slots
            const uint64_t tB = smem_desc_template(128, 1024, LAYOUT_SW128);        // next 64-slot group = one cell further on
            const uint32_t x0 = smem_u32(s_x), d0 = smem_u32(s_d);
            const uint64_t dOne = smem_desc(tA, smem_u32(s_one));
            const bool want_db = p.dbias != nullptr;
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
                for (int r = 0; r < ((p.dbg & 1) ? 0 : 2); ++r) {
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
                        if (want_db) umma_bf16(tmem_base + 448, a, dOne, id1, started & 1u);        // column 448: sum over cells of dY (hi rows, lo rows)
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
    } else if (warp <= W2_PROD_WARPS) {
        // ------------------------------------------------------------------------------------------ producers: 16-byte asynchronous copies
        // A job's bytes are, per channel, the row pair's 2 W contiguous floats (512 B): one warp instruction, lane l copying bytes 16 l ..
        // 16 l + 15, into the stage as [channel][row][W].  Measured alternatives for these 32 chunks per job, 19,968 bytes apart
        // (tools/nchw_read_probe.cu, profiles/r02wg_*): 1-D bulk copies cost ~60-85 cycles EACH whatever their size (512 B chunks: 1.65 TB/s
        // with 64, 128 or 192 KB in flight; 1.5 KB: 4.3; 6.6 KB: 5.7) -- the same rate the forward kernel's TMA boxes reach on these rows;
        // register loads by the converters ran one memory latency per job whatever the look-ahead in the source (six scoreboards per warp).
        const int pw = warp - 1;
        W2Gen gen;
        gen.init(g0, g1, p.NPA);
        int type, bb, pair;
        uint32_t idx, jobno = 0;
        const long chan = (long)p.H * p.W;
        while (gen.next(type, bb, pair, idx)) {
            const uint32_t st = jobno % W2_NSTG;
            if (jobno >= W2_NSTG) mbar_wait(&stg_free[st], ((jobno / W2_NSTG) - 1) & 1);
            const int C = type ? p.Cout : p.Cin;
            const int y0 = 2 * pair;
            const int rows = (y0 < 0 || y0 >= p.H || (p.dbg & 4)) ? 0 : (y0 + 1 < p.H ? 2 : 1);
            const int bytes = rows * p.W * 4;
            if (16 * lane < bytes) {
                const uint8_t* src = reinterpret_cast<const uint8_t*>((type ? p.dy : p.x) + (long)bb * C * chan + (long)y0 * p.W) + 16 * lane;
                uint8_t* dst = s_stg + (size_t)st * W2_STAGE + 16 * lane;
                for (int c = pw; c < C; c += W2_PROD_WARPS) cp_async_16(dst + (size_t)c * 2 * p.W * 4, src + (size_t)c * chan * 4);
            }
            cp_async_arrive(&stg_full[st]);                              // every lane, also without copies: the barrier expects the producers' 64 arrivals
            ++jobno;
        }
    } else {
        // ------------------------------------------------------------------------------------------ converters
        // All sixteen warps work on every job; thread = (channel quarter, row of the pair, column).  The fp32 values come from the staging
        // ring (lanes = consecutive columns: conflict-free), so a job's serial path holds no global-memory latency.
        const int cw = warp - 1 - W2_PROD_WARPS;       // 0..15
        const int q = cw & 3;                          // channels 8q .. 8q+7
        const int r = (cw >> 2) & 1;                   // row of the pair
        const int x = ((cw >> 3) << 5) | lane;         // column 0..63
        const bool has_pro = p.in_scale != nullptr;
        float sc[8], sh[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { sc[j] = s_pro[8 * q + j].x; sh[j] = s_pro[8 * q + j].y; }

        W2Gen gen;
        gen.init(g0, g1, p.NPA);
        int type, bb, pair;
        uint32_t idx, jobno = 0;
        while (gen.next(type, bb, pair, idx)) {
            const uint32_t st = jobno % W2_NSTG;
            const int y = 2 * pair + r;
            const int C = type ? p.Cout : p.Cin;
            const bool in = y >= 0 && y < p.H && x < p.W && !(p.dbg & 6);
            float w[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) w[j] = 0.f;
            mbar_wait_backoff(&stg_full[st], (jobno / W2_NSTG) & 1);
            if (in) {
                const float* src = reinterpret_cast<const float*>(s_stg + (size_t)st * W2_STAGE) + (8 * q) * 2 * p.W + r * p.W + x;      // [channel][row][W]
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (8 * q + j < C) w[j] = src[j * 2 * p.W];
                if (has_pro && type == 0) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) w[j] = 8 * q + j < C ? fmaxf(fmaf(w[j], sc[j], sh[j]), 0.f) : 0.f;
                }
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
            __syncwarp();
            if (lane == 0) mbar_arrive(&stg_free[st]);                   // the values are in registers (the conversion consumed them)
            uint8_t* row;
            uint32_t sw;
            if (type) {
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
            if (!(p.dbg & 2)) {
                *reinterpret_cast<uint4*>(row + (((uint32_t)q ^ sw) << 4)) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                *reinterpret_cast<uint4*>(row + (((uint32_t)(4 + q) ^ sw) << 4)) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                fence_proxy_async_smem();
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(type ? &dfull[idx % W2_DR] : &xfull[idx % W2_XR]);
            ++jobno;
        }
        // ------------------------------------------------------------------------------------------ final fold (all sixteen converter warps)
        // M = 64: accumulator row i sits in TMEM lane 32 (i / 16) + i % 16 (tools/umma_m64_probe.py), i.e. lane quarter 0 / 1 hold dY_hi of
        // channels 0..15 / 16..31 in their first 16 lanes, quarters 2 / 3 dY_lo; columns 64 t + (0..31) = x_hi[ci], + (32..63) = x_lo[ci]
        if (g1 > g0) {
            const int qd = warp & 3;                    // a warp may only touch the TMEM lane quarter warp % 4
            const int part = (warp - 1 - W2_PROD_WARPS) >> 2;       // 0..3: four warps (one per quarter) share a subset of the 7 taps
            const bool hi_rows = qd < 2;
            const int co = (qd & 1) * 16 + lane;
            const bool live = lane < 16 && co < p.Cout;
            mbar_wait_backoff(&bar_done, 0);
            tc_fence_after();
            for (int t = part; t < 7; t += 4) {
                uint32_t r0[32], r1[32];
                tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(t * 64), r0);                     // x x_hi[ci]
                if (hi_rows) tmem_ld32(tmem_base + ((uint32_t)(qd * 32) << 16) + (uint32_t)(t * 64 + 32), r1);   // dY_hi x x_lo[ci]
                tmem_ld_wait();
                if (live) {
                    float* o = p.dwp + (long)t * p.Cin * p.Cout + co;
#pragma unroll
                    for (int ci = 0; ci < 32; ++ci) {
                        float a = __uint_as_float(r0[ci]);
                        if (hi_rows) a += __uint_as_float(r1[ci]);
                        if (ci < p.Cin) atomicAdd(o + (long)ci * p.Cout, a);
                    }
                }
            }
            if (part == 3 && p.dbias != nullptr) {      // the warps with one tap only
                uint32_t rb[4];
                tmem_ld4(tmem_base + ((uint32_t)(qd * 32) << 16) + 448u, rb);
                tmem_ld_wait();
                if (live) atomicAdd(p.dbias + co, __uint_as_float(rb[0]));
            }
            tc_fence_before();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<512>(tmem_base);
}

// dwp[7][cin][cout] += sum dY * x'   (x' = in_scale ? relu(x * in_scale + in_shift) : x);  dbias[cout] += sum dY (nullable).
// Same contract as gn_hexconv_wgrad (caller zeroes both); shapes as gn_hexconv_tc2_supported.
GN_API int gn_hexconv_wgrad_tc2(const float* x, const float* in_scale, const float* in_shift, const float* dy, float* dwp, float* dbias, int B,
                                int cin, int cout, int H, int W, cudaStream_t stream) {
    GN_REQUIRE(x && dy && dwp && B > 0, GN_EINVAL, "hexconv_wgrad_tc2: bad arguments");
    GN_REQUIRE(cin >= 1 && cin <= 32 && cout >= 1 && cout <= 32 && H >= 2 && W >= 4 && W <= 64 && W % 4 == 0, GN_EUNSUPPORTED,
               "hexconv_wgrad_tc2: needs kernel_size 1, <= 32 channels, W <= 64 and W % 4 == 0");
    GN_REQUIRE((((uintptr_t)x | (uintptr_t)dy) & 15) == 0, GN_EALIGN, "hexconv_wgrad_tc2: x and dy must be 16-byte aligned (bulk copies)");
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
        e = getenv("GRIDNEXT_B200_HEXWG2_DBG");
        p.dbg = e ? atoi(e) : 0;
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
