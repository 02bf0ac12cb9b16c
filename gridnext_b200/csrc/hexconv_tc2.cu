// Hexagonal convolution (kernel_size 1, <= 32 channels, grid width <= 64) on tcgen05 -- second generation: fp32 NCHW in, fp32 NCHW
// out, NO intermediate layout in global memory.
//
// Replaces hexagdly.Conv2d.forward and its data gradient as used by the g network
// (/root/reference/gridnext/gridnet_models.py:128-148) at large batches.  hexconv_tc.cu (first generation) first rewrote the input
// into two bf16 parity planes (read 128 B + write 128 B per cell), then loaded two overlapping plane windows per tile (196 B per
// cell): 580 B of traffic per cell against the 256 algorithmic bytes -- at most 44 % of the HBM roofline, measured 25-29 %.
// Here the traffic is the algorithmic one (+ one halo row pair per 26 rows):
//
//   * a persistent CTA walks segments of consecutive grid rows of one array (H2Seg); the producer warp TMA-loads the fp32 rows (box =
//     64 columns x 1 row x 32 channels, out-of-grid rows / columns / channels zero-filled) into a staging ring;
//   * four converter warps apply the previous layer's BatchNorm+ReLU (gridnet_models.py:134-136), split every value into
//     bf16 hi + lo (x = hi + lo to 16 mantissa bits; fp32-grade accuracy with x_hi w_hi + x_lo w_hi + x_hi w_lo, north_star
//     1e-5) and write the cell's 128-byte K-major operand row [32 hi | 32 lo] (SWIZZLE_128B) into a ring of 12 grid-row slots
//     (+ a mirror of slot 0 behind slot 11, so that any two consecutive rows are contiguous); every row is loaded and converted once;
//   * a tile is two grid rows (y even, y + 1) x 64 columns = the 128 accumulator rows of one UMMA.  The hexagonal neighbourhood
//     mixes row shifts (y - 1, y + 1) with parity-dependent column shifts.  Row shifts are operand START ADDRESSES (a whole
//     8 KB slot); column shifts are NOT applied to the operand: the taps are stacked along UMMA N instead --
//         same row : N = 96 = taps (x-1, x, x+1) x 32 channels      A = rows (y,   y+1)
//         row above: N = 64 = taps (a = 0, 1)  x 32 channels        A = rows (y-1, y  )
//         row below: N = 64                                         A = rows (y+1, y+2)
//     so accumulator column block j of lane q holds E_j[q] = X[row-shifted q] . W_j, and the output is a neighbour-LANE sum
//         out[x] = L[x-1] + C[x] + R[x+1],  L = E_l + (even row: E_u0 + E_d0), R = E_r + (odd row: E_u1 + E_d1), C = the rest,
//     two warp shuffles per channel in the epilogue (lanes are columns; the x = 31|32 warp boundary goes through shared memory).
//     A UMMA with N = 32 costs the same ~45 cycles as one with N = 96 (A-operand fetch bound), so stacking turns 42 MMAs per
//     tile into 18 (912 cycles per 128 cells against the 1,400 cycles HBM needs for their 32 KB): the kernel can be HBM-bound;
//   * eight epilogue warps read TMEM, do the lane sums, add the bias, write NCHW fp32 (one coalesced 128-byte store per channel
//     and warp) and keep the BatchNorm batch statistics of their 16 channels in registers across all tiles (one reduction at the
//     end, fp64 atomics per CTA only).
//
//   warp 0: TMA producer   warp 1: MMA issuer   warps 2-5: converters   warps 6-13: epilogue (TMEM lane group x channel half)
#include "gn_common.cuh"
#include "gn_ptx.cuh"
#include "gn_tma.cuh"
#include <stdlib.h>

using namespace gnptx;

#define H2_TRACE(ev, idx)                                                                 \
    do {                                                                                  \
        if (p.trace != nullptr && blockIdx.x == 0 && (idx) < 64) p.trace[(ev) * 64 + (idx)] = clock64(); \
    } while (0)

// a segment of nt tiles (H2Seg below) loads nt + 2 row pairs: pair k = rows y0 - 2 + 2k, y0 - 1 + 2k (first / last: one halo row)
#define H2_STAGES 3                  // fp32 staging ring (one row pair = 16 KB per stage)
#define H2_RP 6                      // operand ring depth in row pairs: a tile reads 3, so the converters may run 3 pairs ahead of the MMAs
                                     // (with 4 the converter <-> MMA barrier hand-offs were on the critical path: 2.4 us per tile with all work switched off)
#define H2_ROW_F32 8192              // one staged grid row: 32 channels x 64 columns fp32
#define H2_SLOT 8192                 // one operand row slot: 64 cells x 128 B
#define H2_W_ROWS 448                // packed weight rows: S_hi 96, S_lo 96, U_hi 64, U_lo 64, D_hi 64, D_lo 64
#define H2_W_BYTES (H2_W_ROWS * 128)
#define H2_THREADS 448

struct Hex2Params {
    int B, H, W, Cin, Cout;
    int TPA;                 // tiles (row pairs) per array
    long n_tiles;            // B * TPA
    long long* trace;        // development: per-role clock64 timestamps of CTA 0 ([13 events][64]), see tools/hextc_trace.py
    int dbg;                 // development switches (GRIDNEXT_B200_H2_DBG): 1 no output stores, 2 no MMAs, 4 no conversion, 8 no TMEM reads
    int tma_in;              // 0 (GRIDNEXT_B200_H2_CPASYNC=1): the input rows by 16-byte asynchronous copies instead of TMA boxes
    const float* x;
    const float* bias;
    const float* in_scale;
    const float* in_shift;
    float* y;
    double* stats;
};

// A CTA owns a contiguous range of the tiles (array-major); it walks it in SEGMENTS of consecutive row pairs of one array -- the
// range start and every array boundary start a new segment with its own halo pairs.  (First version: fixed strips of 26 rows dealt
// round-robin, 768 strips on 148 CTAs = 6 strips for some and 5 for others, 15 pairs loaded per 13 tiles.)
struct H2Seg {
    long cur, t1;
    int TPA, b, y0, nt;
    __device__ void init(const long n_tiles, int TPA_) {
        cur = n_tiles * blockIdx.x / gridDim.x;
        t1 = n_tiles * (blockIdx.x + 1) / gridDim.x;
        TPA = TPA_;
    }
    __device__ bool next() {
        if (cur >= t1) return false;
        b = (int)(cur / TPA);
        const int p0 = (int)(cur - (long)b * TPA);
        const long left = t1 - cur;
        nt = left < (long)(TPA - p0) ? (int)left : TPA - p0;
        y0 = 2 * p0;
        cur += nt;
        return true;
    }
};

// wp fp32 [7][cin][cout] (gn_hexconv_pack, mode 0 forward / mode 1 data gradient) -> bf16 B-operand tiles, 128-byte rows:
//   hi tile row = [w_hi(ci 0..31) | w_hi(ci 0..31)]  (against A = [x_hi | x_lo]),  lo tile row = [w_lo(ci 0..31) | 0] (against x_hi)
//   rows: S_hi [tap 0..2][co], S_lo, U_hi [tap 3,4][co], U_lo, D_hi [tap 5,6][co], D_lo
__global__ void hex2_pack_kernel(const float* __restrict__ wp, int cin, int cout, __nv_bfloat16* __restrict__ wt) {
    gn_pdl_sync();
    const int total = H2_W_ROWS * 64;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int kk = e & 63, row = e >> 6;
        int grp_row = row, tap0, lo;
        if (row < 192) { lo = row >= 96; grp_row = row - (lo ? 96 : 0); tap0 = 0; }
        else if (row < 320) { lo = row >= 256; grp_row = row - (lo ? 256 : 192); tap0 = 3; }
        else { lo = row >= 384; grp_row = row - (lo ? 384 : 320); tap0 = 5; }
        const int t = tap0 + (grp_row >> 5), co = grp_row & 31, ci = kk & 31;
        float w = 0.f;
        if (ci < cin && co < cout) w = wp[((long)t * cin + ci) * cout + co];
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        float v;
        if (!lo) v = __bfloat162float(hi);
        else v = kk < 32 ? w - __bfloat162float(hi) : 0.f;
        wt[e] = __float2bfloat16_rn(v);
    }
}

__global__ void __launch_bounds__(H2_THREADS, 1)
hexconv_tc2_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmX2, const __grid_constant__ CUtensorMap tmW,
                   const Hex2Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_w, stg_full[H2_STAGES], stg_free[H2_STAGES], ring_full[H2_RP], ring_free[H2_RP], tm_full[2], tm_free[2];
    __shared__ uint32_t tmem_slot;
    __shared__ float s_pro[2][32];
    __shared__ __align__(16) float s_xchg[2][4][2][2][16];       // [tile parity][lane group][channel half][L of lane 31 | R of lane 0][16]
    __shared__ double s_stat[2 * 32];
    __shared__ __align__(16) float s_bias[32];
    const uint32_t raw = smem_u32(smem_raw);
    uint8_t* sm = smem_raw + (((raw + 1023u) & ~1023u) - raw);
    uint8_t* s_w = sm;                                           // packed weights, 56 KB
    uint8_t* s_ring = s_w + H2_W_BYTES;                          // 2 * H2_RP operand row slots + a mirror of slot 0 behind the last one, 104 KB
    uint8_t* s_stg = s_ring + (2 * H2_RP + 1) * H2_SLOT;         // staging: [stage][32 channels][row of the pair][64] fp32, 48 KB
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x < 64) s_stat[threadIdx.x] = 0.0;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmX2);
        tma_prefetch_desc(&tmW);
        mbar_init(&bar_w, 1);
        for (int s = 0; s < H2_STAGES; ++s) { mbar_init(&stg_full[s], p.tma_in ? 1 : 32); mbar_init(&stg_free[s], 4); }
        for (int s = 0; s < H2_RP; ++s) { mbar_init(&ring_full[s], 4); mbar_init(&ring_free[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tm_full[s], 1); mbar_init(&tm_free[s], 8); }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc<512>(&tmem_slot);
    gn_pdl_wait();
    if (threadIdx.x < 32) {
        s_bias[threadIdx.x] = (p.bias != nullptr && threadIdx.x < p.Cout) ? p.bias[threadIdx.x] : 0.f;
        s_pro[0][threadIdx.x] = (p.in_scale != nullptr && threadIdx.x < p.Cin) ? p.in_scale[threadIdx.x] : 1.f;
        s_pro[1][threadIdx.x] = (p.in_shift != nullptr && threadIdx.x < p.Cin) ? p.in_shift[threadIdx.x] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    gn_pdl_trigger();
    const uint32_t tmem_base = tmem_slot;
    const bool has_pro = p.in_scale != nullptr;

    if (warp == 0 && !p.tma_in) {
        // ------------------------------------------------------------------------------------------ producer: 16-byte asynchronous copies
        // (experiment, GRIDNEXT_B200_H2_CPASYNC=1) One warp instruction per channel: lane = (row of the pair, 4 columns), 512 contiguous
        // bytes of a regular pair; rows / columns / channels outside the tensor are written as zeros (src-size 0), which is what the TMA
        // boxes' out-of-bounds fill does.  This input path took the weight gradient (hexconv_wgrad_tc2.cu) from 1.6 to 2.4 TB/s; HERE it
        // changes nothing (0.161 vs 0.151 ms at 256 arrays, profiles/r02wg_hextc_time.txt): with eight boxes per pair the forward is bound
        // by its epilogue (~3,000 cycles per tile), not by its loads.
        if (lane == 0) {
            mbar_arrive_expect_tx(&bar_w, H2_W_BYTES);
            for (int i = 0; i < H2_W_ROWS / 64; ++i) tma_load_2d(&tmW, &bar_w, s_w + i * 64 * 128, 0, i * 64);      // 7 boxes of 64 rows
        }
        const int rr = lane >> 4, x4 = (lane & 15) * 4;
        const long chan = (long)p.H * p.W;
        uint32_t gk = 0;
        H2Seg seg;
        seg.init(p.n_tiles, p.TPA);
        while (seg.next()) {
            const int b = seg.b, y0 = seg.y0, nt = seg.nt, nchunks = nt + 2;
            (void)nt; (void)nchunks;
            for (int k = 0; k < nchunks; ++k, ++gk) {
                const int st = gk % H2_STAGES;
                if (gk >= H2_STAGES) mbar_wait(&stg_free[st], ((gk / H2_STAGES) - 1) & 1);
                uint8_t* dst = s_stg + (size_t)st * 2 * H2_ROW_F32;
                const int r_lo = y0 - 2 + 2 * k;
                const bool first = k == 0, last = k == nchunks - 1;
                // staged layout [channel][row of the pair][x] for a regular pair; a halo pair's one row as [channel][x] in its half of the stage
                const bool mine = first ? rr == 1 : (last ? rr == 0 : true);
                if (mine) {
                    const int row = r_lo + rr;
                    const bool ok = row >= 0 && row < p.H && x4 < p.W;
                    const float* src = ok ? p.x + ((long)b * p.Cin * p.H + row) * p.W + x4 : p.x;
                    uint8_t* d = dst + ((first || last) ? (first ? H2_ROW_F32 : 0) + x4 * 4 : rr * 256 + x4 * 4);
                    const int cstride = (first || last) ? 256 : 512;
#pragma unroll 8
                    for (int c = 0; c < 32; ++c) {
                        const bool okc = ok && c < p.Cin;
                        cp_async_16_zfill(d + c * cstride, okc ? src + c * chan : p.x, okc ? 16u : 0u);
                    }
                }
                cp_async_arrive(&stg_full[st]);
            }
        }
    } else if (warp == 0) {
        // ------------------------------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            mbar_arrive_expect_tx(&bar_w, H2_W_BYTES);
            for (int i = 0; i < H2_W_ROWS / 64; ++i) tma_load_2d(&tmW, &bar_w, s_w + i * 64 * 128, 0, i * 64);      // 7 boxes of 64 rows
            uint32_t gk = 0;                                            // running row-pair index of this CTA
            H2Seg seg;
            seg.init(p.n_tiles, p.TPA);
            while (seg.next()) {
                const int b = seg.b, y0 = seg.y0, nt = seg.nt, nchunks = nt + 2;
                (void)nt; (void)nchunks;
                for (int k = 0; k < nchunks; ++k, ++gk) {
                    const int st = gk % H2_STAGES;
                    if (gk >= H2_STAGES) mbar_wait(&stg_free[st], ((gk / H2_STAGES) - 1) & 1);
                    uint8_t* dst = s_stg + (size_t)st * 2 * H2_ROW_F32;
                    const int r_lo = y0 - 2 + 2 * k;
                    const bool first = k == 0, last = k == nchunks - 1;
                    // staged layout [channel][row of the pair][x]: a regular pair is ONE box (64 x 2 rows x 32 channels: 512 contiguous bytes per
                    // channel); the halo pairs at the segment ends load their single row into the same layout with the one-row map
                    mbar_arrive_expect_tx(&stg_full[st], (first || last) ? H2_ROW_F32 : 2 * H2_ROW_F32);
                    // A TMA operation walks the rows of its box (here 256-byte rows 20 KB apart) nearly one at a time: with one 64-row box
                    // per pair the loads alone took 0.19 ms for 256 arrays (1.7 TB/s; measured with every other role switched off).  Eight
                    // boxes of 4 channels per pair keep 32 operations in flight per SM instead of 4.
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (first) tma_load_4d(&tmX, &stg_full[st], dst + H2_ROW_F32 + i * 1024, 0, r_lo + 1, 4 * i, b);
                        else if (last) tma_load_4d(&tmX, &stg_full[st], dst + i * 1024, 0, r_lo, 4 * i, b);
                        else tma_load_4d(&tmX2, &stg_full[st], dst + i * 2048, 0, r_lo, 4 * i, b);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------------------------------ MMA issuer
        if (elect_one()) {
            const uint32_t idS = idesc_bf16(128, 96, 0, 0), idU = idesc_bf16(128, 64, 0, 0);
            const uint64_t tmpl = smem_desc_template(0, 1024, LAYOUT_SW128);
            const uint32_t w0 = smem_u32(s_w), ring0 = smem_u32(s_ring);
            const uint64_t dS_hi = smem_desc(tmpl, w0), dS_lo = smem_desc(tmpl, w0 + 96 * 128);
            const uint64_t dU_hi = smem_desc(tmpl, w0 + 192 * 128), dU_lo = smem_desc(tmpl, w0 + 256 * 128);
            const uint64_t dD_hi = smem_desc(tmpl, w0 + 320 * 128), dD_lo = smem_desc(tmpl, w0 + 384 * 128);
            mbar_wait(&bar_w, 0);
            uint32_t gk = 0, waited = 0, tile_seq = 0;
            H2Seg seg;
            seg.init(p.n_tiles, p.TPA);
            while (seg.next()) {
                const int nt = seg.nt, nchunks = nt + 2;
                for (int t = 0; t < nt; ++t, ++tile_seq) {
                    // row pairs gk + t, + t + 1, + t + 2 must be converted
                    while (waited < gk + t + 3) { mbar_wait(&ring_full[waited % H2_RP], (waited / H2_RP) & 1); ++waited; }
                    const int acc = tile_seq & 1;
                    if (tile_seq >= 2) mbar_wait(&tm_free[acc], ((tile_seq >> 1) - 1) & 1);
                    tc_fence_after();
                    H2_TRACE(4, tile_seq);
                    // slot of row index i (running over all segments: 2 * (gk + pair) + row): i mod 12; a pair starting at the last slot continues
                    // in the mirror slot behind it
                    const uint32_t i_own = 2 * (gk + t + 1);                     // first own row (even slot)
                    const uint32_t a_own = ring0 + ((i_own % (2 * H2_RP)) * H2_SLOT);
                    const uint32_t a_up = ring0 + (((i_own - 1) % (2 * H2_RP)) * H2_SLOT);
                    const uint32_t a_dn = ring0 + (((i_own + 1) % (2 * H2_RP)) * H2_SLOT);
                    const uint32_t d = tmem_base + (uint32_t)(acc * 256);
                    const uint64_t dA_own = smem_desc(tmpl, a_own), dA_up = smem_desc(tmpl, a_up), dA_dn = smem_desc(tmpl, a_dn);
                    if (!(p.dbg & 2)) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(d, dA_own + (uint64_t)(2 * k), dS_hi + (uint64_t)(2 * k), idS, k > 0);
#pragma unroll
                    for (int k = 0; k < 2; ++k) umma_bf16(d, dA_own + (uint64_t)(2 * k), dS_lo + (uint64_t)(2 * k), idS, 1u);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(d + 96, dA_up + (uint64_t)(2 * k), dU_hi + (uint64_t)(2 * k), idU, k > 0);
#pragma unroll
                    for (int k = 0; k < 2; ++k) umma_bf16(d + 96, dA_up + (uint64_t)(2 * k), dU_lo + (uint64_t)(2 * k), idU, 1u);
                    // the rows below go into the SAME two column blocks as the rows above: both taps a = 0 / 1 of either row take the same lane
                    // shift, so the tensor core adds them (64 accumulator columns and 32 TMEM loads + adds per epilogue thread fewer)
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_bf16(d + 96, dA_dn + (uint64_t)(2 * k), dD_hi + (uint64_t)(2 * k), idU, 1u);
#pragma unroll
                    for (int k = 0; k < 2; ++k) umma_bf16(d + 96, dA_dn + (uint64_t)(2 * k), dD_lo + (uint64_t)(2 * k), idU, 1u);
                    }
                    umma_commit(&tm_full[acc]);
                    umma_commit(&ring_free[(gk + t) % H2_RP]);                   // row pair gk + t is not read again
                    if (t == nt - 1) {
                        umma_commit(&ring_free[(gk + t + 1) % H2_RP]);
                        umma_commit(&ring_free[(gk + t + 2) % H2_RP]);
                    }
                    H2_TRACE(5, tile_seq);
                }
                gk += nchunks;
            }
        }
    } else if (warp < 6) {
        // ------------------------------------------------------------------------------------------ converters: thread = (row of the pair, column)
        const int cw = warp - 2;                    // 0..3
        const int r = cw >> 1;                      // row of the pair
        const int x = ((cw & 1) << 5) | lane;       // column 0..63
        uint32_t gk = 0;
        H2Seg seg;
        seg.init(p.n_tiles, p.TPA);
        while (seg.next()) {
            const int b = seg.b, y0 = seg.y0, nt = seg.nt, nchunks = nt + 2;
            (void)nt; (void)nchunks;
            (void)b;
            for (int k = 0; k < nchunks; ++k, ++gk) {
                const int st = gk % H2_STAGES, rs = gk % H2_RP;
                mbar_wait_backoff(&stg_full[st], (gk / H2_STAGES) & 1);
                if (threadIdx.x == 64) H2_TRACE(1, gk);
                if (gk >= H2_RP) mbar_wait_backoff(&ring_free[rs], ((gk / H2_RP) - 1) & 1);
                if (threadIdx.x == 64) H2_TRACE(2, gk);
                const int gy = y0 - 2 + 2 * k + r;
                const bool loaded = !((k == 0 && r == 0) || (k == nchunks - 1 && r == 1));
                if (loaded && !(p.dbg & 4)) {
                    // regular pair: [c][r][x] (128 floats per channel); halo pair: its one row as [c][x] in the half of the stage it was loaded to
                    const bool halo = k == 0 || k == nchunks - 1;
                    const int cs = halo ? 64 : 128;
                    const float* src = reinterpret_cast<const float*>(s_stg + (size_t)st * 2 * H2_ROW_F32) + (halo ? r * (H2_ROW_F32 / 4) : r * 64) + x;
                    const bool live = has_pro && gy >= 0 && gy < p.H && x < p.W;     // zero padding stays zero
                    uint32_t hi[16], lo[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        float v0 = src[(2 * j) * cs], v1 = src[(2 * j + 1) * cs];
                        if (live) {
                            v0 = 2 * j < p.Cin ? fmaxf(fmaf(v0, s_pro[0][2 * j], s_pro[1][2 * j]), 0.f) : 0.f;
                            v1 = 2 * j + 1 < p.Cin ? fmaxf(fmaf(v1, s_pro[0][2 * j + 1], s_pro[1][2 * j + 1]), 0.f) : 0.f;
                        }
                        const __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
                        const float2 hf = __bfloat1622float2(h);
                        const __nv_bfloat162 l = __floats2bfloat162_rn(v0 - hf.x, v1 - hf.y);
                        hi[j] = *reinterpret_cast<const uint32_t*>(&h);
                        lo[j] = *reinterpret_cast<const uint32_t*>(&l);
                    }
                    // operand row of this cell: 8 chunks of 16 B, chunk c stored at (c ^ (x & 7)) (SWIZZLE_128B: slots are 1024-byte aligned)
                    const uint32_t slot = (2 * gk + r) % (2 * H2_RP);
                    uint8_t* row = s_ring + (size_t)slot * H2_SLOT + (size_t)x * 128;
                    const uint32_t sw = (uint32_t)(x & 7);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        *reinterpret_cast<uint4*>(row + ((c ^ sw) << 4)) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                        *reinterpret_cast<uint4*>(row + (((4 + c) ^ sw) << 4)) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
                    }
                    if (slot == 0) {
                        uint8_t* mrow = row + 2 * H2_RP * H2_SLOT;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            *reinterpret_cast<uint4*>(mrow + ((c ^ sw) << 4)) = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                            *reinterpret_cast<uint4*>(mrow + (((4 + c) ^ sw) << 4)) = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
                        }
                    }
                }
                fence_proxy_async_smem();            // generic-proxy writes -> visible to the tensor core's (async proxy) operand reads
                __syncwarp();
                if (lane == 0) {
                    mbar_arrive(&stg_free[st]);
                    mbar_arrive(&ring_full[rs]);
                }
                if (threadIdx.x == 64) H2_TRACE(3, gk);
            }
        }
    } else {
        // ------------------------------------------------------------------------------------------ epilogue: warp = (TMEM lane group g, channel half h)
        const int ew = warp - 6;                     // 0..7
        const int g = warp & 3;                      // TMEM lanes 32g .. 32g+31 (a warp may only touch the lane quarter warp % 4)
        const int h = ew >> 2;                       // channels 16h .. 16h+15
        const int par = g >> 1;                      // row of the tile = parity of the grid row (tiles start on even rows)
        const int x = ((g & 1) << 5) | lane;
        const bool right_half = g & 1;
        float sg[16], sq[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) sg[e] = sq[e] = 0.f;
        const long chan = (long)p.H * p.W;
        uint32_t tile_seq = 0;
        H2Seg seg;
        seg.init(p.n_tiles, p.TPA);
        while (seg.next()) {
            const int b = seg.b, y0 = seg.y0, nt = seg.nt, nchunks = nt + 2;
            (void)nt; (void)nchunks;
            for (int t = 0; t < nt; ++t, ++tile_seq) {
                const int acc = tile_seq & 1;
                const int yy = y0 + 2 * t + par;
                const bool valid = yy < p.H && x < p.W;
                mbar_wait_backoff(&tm_full[acc], (tile_seq >> 1) & 1);
                tc_fence_after();
                if (threadIdx.x == 192) H2_TRACE(6, tile_seq);
                if (threadIdx.x == 416) H2_TRACE(11, tile_seq);
                const uint32_t ta = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(acc * 256 + 16 * h);
                float L[16], C[16], R[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) L[e] = C[e] = R[e] = 0.f;
                if (!(p.dbg & 8)) {
                    uint32_t el[16], ec[16], er[16];
                    tmem_ld16(ta, el);
                    tmem_ld16(ta + 32, ec);
                    tmem_ld16(ta + 64, er);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) { L[e] = __uint_as_float(el[e]); C[e] = __uint_as_float(ec[e]); R[e] = __uint_as_float(er[e]); }
                }
#pragma unroll
                for (int q8 = 0; q8 < ((p.dbg & 8) ? 0 : 2); ++q8) {                           // 8 channels at a time (register budget: 146 per thread)
                    uint32_t u0[8], u1[8];
                    tmem_ld8(ta + 96 + 8 * q8, u0);
                    tmem_ld8(ta + 128 + 8 * q8, u1);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const float a0 = __uint_as_float(u0[e]);      // tap a = 0 of the rows above + below (summed in the accumulator)
                        const float a1 = __uint_as_float(u1[e]);      // tap a = 1
                        if (par == 0) { L[8 * q8 + e] += a0; C[8 * q8 + e] += a1; }            // even row: columns x-1, x
                        else { C[8 * q8 + e] += a0; R[8 * q8 + e] += a1; }                     // odd row:  columns x, x+1
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tm_free[acc]);                // the accumulator is in registers
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 b4 = *reinterpret_cast<const float4*>(&s_bias[16 * h + 4 * q]);
                    C[4 * q] += b4.x; C[4 * q + 1] += b4.y; C[4 * q + 2] += b4.z; C[4 * q + 3] += b4.w;
                }
                if (threadIdx.x == 192) H2_TRACE(7, tile_seq);
                if (threadIdx.x == 416) H2_TRACE(12, tile_seq);
                // boundary lanes: column 31 | 32 sits between the two warps of a grid row
                float* xw = &s_xchg[acc][g][h][0][0];
                if (lane == 31) {
#pragma unroll
                    for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(xw + e) = make_float4(L[e], L[e + 1], L[e + 2], L[e + 3]);
                }
                if (lane == 0) {
#pragma unroll
                    for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(xw + 16 + e) = make_float4(R[e], R[e + 1], R[e + 2], R[e + 3]);
                }
                if (threadIdx.x == 192) H2_TRACE(9, tile_seq);
                named_bar_sync(3, 256);
                if (threadIdx.x == 192) H2_TRACE(10, tile_seq);
                // every lane reads the neighbour warp's boundary values (a broadcast load) and selects: lane-conditional loads compiled into one
                // divergent branch per channel and made this tail 5,000 cycles long (the whole kernel ran at 6,300 cycles per tile; clock64 trace)
                const float4* nb4 = reinterpret_cast<const float4*>(right_half ? &s_xchg[acc][g - 1][h][0][0] : &s_xchg[acc][g + 1][h][1][0]);
                float nbv[16];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4 t4 = nb4[q];
                    nbv[4 * q] = t4.x; nbv[4 * q + 1] = t4.y; nbv[4 * q + 2] = t4.z; nbv[4 * q + 3] = t4.w;
                }
                float* outp = p.y + ((long)b * p.Cout * p.H + yy) * p.W + x + (long)(16 * h) * chan;      // advanced by one channel plane per element
                const int n_ch = min(16, p.Cout - 16 * h);                                                // channels of this half that exist
                const bool store = valid && !(p.dbg & 1);
                // all 32 shuffles first, back to back (independent), then the arithmetic: issued element by element the in-order warp waited out
                // a shuffle + a shared-memory latency per channel (2,650 cycles for this loop; clock64 trace)
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    L[e] = __shfl_up_sync(0xffffffffu, L[e], 1);
                    R[e] = __shfl_down_sync(0xffffffffu, R[e], 1);
                }
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    // x = 0 / x = 63: zero padding; x = 32 takes L of the left warp's lane 31, x = 31 takes R of the right warp's lane 0
                    const float lf = lane == 0 ? (right_half ? nbv[e] : 0.f) : L[e];
                    const float rt = lane == 31 ? (right_half ? 0.f : nbv[e]) : R[e];
                    const bool ok = valid && e < n_ch;
                    const float val = ok ? C[e] + lf + rt : 0.f;                                           // the bias is already in C
                    if (store && e < n_ch) *outp = val;
                    outp += chan;
                    C[e] = val;
                }
                if (p.stats != nullptr) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        sg[e] += C[e];
                        sq[e] = fmaf(C[e], C[e], sq[e]);
                    }
                }
                if (threadIdx.x == 192) H2_TRACE(8, tile_seq);
            }
        }
        if (p.stats != nullptr) {
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                const float a = gn_warp_sum(sg[e]), q = gn_warp_sum(sq[e]);
                if (lane == 0 && 16 * h + e < p.Cout) {
                    atomicAdd(&s_stat[16 * h + e], (double)a);
                    atomicAdd(&s_stat[32 + 16 * h + e], (double)q);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.stats != nullptr && threadIdx.x < p.Cout) {
        atomicAdd(p.stats + threadIdx.x, s_stat[threadIdx.x]);
        atomicAdd(p.stats + p.Cout + threadIdx.x, s_stat[32 + threadIdx.x]);
    }
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ---------------------------------------------------------------------------------------------- C-ABI
static long long* g_h2_trace = nullptr;
// development: device buffer of 9 x 64 int64 that CTA 0 of the next launches fills with clock64() timestamps per pipeline event
GN_API int gn_hexconv_tc2_set_trace(long long* dev_buf) {
    g_h2_trace = dev_buf;
    return GN_OK;
}

GN_API long gn_hexconv_tc2_workspace_bytes(void) { return H2_W_BYTES + 1024; }

GN_API int gn_hexconv_tc2_supported(int cin, int cout, int H, int W, int ksize) {
    return ksize == 1 && cin >= 1 && cin <= 32 && cout >= 1 && cout <= 32 && H >= 2 && W >= 4 && W <= 64 && W % 4 == 0;
}

// y = hexconv(x') + bias with x' = in_scale ? relu(x * in_scale + in_shift) : x.  Same contract as gn_hexconv_fwd / gn_hexconv_fwd_tc
// (wp from gn_hexconv_pack: mode 0 forward, mode 1 data gradient); stats (nullable): fp64 [2 * cout] sum / sum of squares, accumulated.
// workspace: gn_hexconv_tc2_workspace_bytes() of caller-owned device memory, 1024-byte aligned (packed weights).
GN_API int gn_hexconv_fwd_tc2(const float* x, const float* wp, const float* bias, const float* in_scale, const float* in_shift, float* y,
                              double* stats, int B, int cin, int cout, int H, int W, void* workspace, cudaStream_t stream) {
    GN_REQUIRE(x && wp && y && workspace && B > 0, GN_EINVAL, "hexconv_fwd_tc2: bad arguments");
    GN_REQUIRE(gn_hexconv_tc2_supported(cin, cout, H, W, 1), GN_EUNSUPPORTED, "hexconv_fwd_tc2: needs kernel_size 1, <= 32 channels, W <= 64 and W % 4 == 0");
    GN_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), GN_EINVAL, "hexconv_fwd_tc2: in_scale/in_shift must come together");
    GN_REQUIRE(((uintptr_t)workspace & 1023) == 0 && ((uintptr_t)x & 15) == 0, GN_EALIGN, "hexconv_fwd_tc2: workspace must be 1024-byte aligned, x 16-byte aligned");
    Hex2Params p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.H = H; p.W = W; p.Cin = cin; p.Cout = cout;
    p.TPA = (H + 1) / 2;
    p.n_tiles = (long)B * p.TPA;
    p.bias = bias; p.in_scale = in_scale; p.in_shift = in_shift; p.y = y; p.stats = stats;
    {
        const char* e = getenv("GRIDNEXT_B200_H2_DBG");
        p.dbg = e ? atoi(e) : 0;
        p.tma_in = !gn_env_flag("GRIDNEXT_B200_H2_CPASYNC");
        p.x = x;
    }
    p.trace = g_h2_trace;
    __nv_bfloat16* wt = (__nv_bfloat16*)workspace;
    GN_CUDA(gn_launch(hex2_pack_kernel, dim3(gn_ceil_div(H2_W_ROWS * 64, 256)), dim3(256), 0, stream, wp, cin, cout, wt));
    GN_LAUNCH_CHECK();
    CUtensorMap tmX, tmX2, tmW;
    {
        uint64_t dims[4] = {(uint64_t)W, (uint64_t)H, (uint64_t)cin, (uint64_t)B};
        uint64_t strides[3] = {(uint64_t)W * 4, (uint64_t)H * W * 4, (uint64_t)cin * H * W * 4};
        uint32_t box[4] = {64, 1, 4, 1};
        int rc = gn_tmap_encode(&tmX, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
        box[1] = 2;
        rc = gn_tmap_encode(&tmX2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc) return rc;
    }
    {
        uint64_t dims[2] = {64, (uint64_t)H2_W_ROWS};
        uint64_t strides[1] = {128};
        uint32_t box[2] = {64, 64};
        int rc = gn_tmap_encode(&tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, wt, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc) return rc;
    }
    const size_t smem = (size_t)H2_W_BYTES + (2 * H2_RP + 1) * H2_SLOT + (size_t)H2_STAGES * 2 * H2_ROW_F32 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        GN_CUDA(cudaFuncSetAttribute(hexconv_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_set = true;
    }
    const int grid = p.n_tiles < gn_num_sms() ? (int)p.n_tiles : gn_num_sms();
    GN_CUDA(gn_launch(hexconv_tc2_kernel, dim3(grid), dim3(H2_THREADS), smem, stream, tmX, tmX2, tmW, p));
    GN_LAUNCH_CHECK();
    return GN_OK;
}
