// Memory-bound pieces of the DenseNet-BC f network in NHWC bf16 (everything that is not a GEMM / 3x3 conv):
// stem max-pool, transition BN-ReLU-avgpool, head BN-ReLU-global-average-pool + classifier, and their
// backward passes with the BatchNorm parameter-gradient column sums fused in.
//
// Reference: /root/reference/gridnext/densenet.py  conv0/norm0/relu0/pool0 (:103-112), _Transition (:47-54),
// norm_final + F.relu + adaptive_avg_pool2d + classifier (:134,152-158).  BatchNorm is eval-mode (training.py:126):
// y = x * scale[c] + shift[c], scale = gamma / sqrt(var + eps), shift = beta - mean * scale.
// Each thread handles 8 consecutive channels (one 16-byte vector) of one pixel.
#include "gn_common.cuh"
#include "gn_epilogue.cuh"
#include <algorithm>

struct V8 { float v[8]; };

__device__ __forceinline__ V8 ld_bf16x8(const __nv_bfloat16* p) {
    const uint4 t = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
    V8 r;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&w[e]));
        r.v[2 * e] = f.x; r.v[2 * e + 1] = f.y;
    }
    return r;
}
__device__ __forceinline__ void st_bf16x8(__nv_bfloat16* p, const V8& a) {
    uint4 t;
    t.x = gn_pack_bf16x2(a.v[0], a.v[1]); t.y = gn_pack_bf16x2(a.v[2], a.v[3]);
    t.z = gn_pack_bf16x2(a.v[4], a.v[5]); t.w = gn_pack_bf16x2(a.v[6], a.v[7]);
    *reinterpret_cast<uint4*>(p) = t;
}
__device__ __forceinline__ V8 ld_f32x8(const float* p) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    V8 r; r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w; r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}

// per-thread column partial sums -> shared -> global (colsum[0][c] += g, colsum[1][c] += gx)
struct ColAcc {
    float g[8], x[8];
    int cg;   // channel group the registers belong to (-1: none yet)
    __device__ __forceinline__ void init() {
#pragma unroll
        for (int j = 0; j < 8; ++j) { g[j] = 0.f; x[j] = 0.f; }
        cg = -1;
    }
    __device__ __forceinline__ void flush(float* s_sum, int C) {
        if (cg >= 0) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                atomicAdd(&s_sum[cg * 8 + j], g[j]);
                atomicAdd(&s_sum[C + cg * 8 + j], x[j]);
                g[j] = 0.f; x[j] = 0.f;
            }
        }
    }
    __device__ __forceinline__ void select(int new_cg, float* s_sum, int C) {
        if (new_cg != cg) { flush(s_sum, C); cg = new_cg; }
    }
};

__device__ __forceinline__ void colsum_block_begin(float* s_sum, int C) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) s_sum[i] = 0.f;
    __syncthreads();
}
__device__ __forceinline__ void colsum_block_end(float* s_sum, int C, float* colsum, int ldsum) {
    __syncthreads();
    for (int i = threadIdx.x; i < C; i += blockDim.x) {
        const float a = s_sum[i], b = s_sum[C + i];
        if (a != 0.f) atomicAdd(colsum + i, a);
        if (b != 0.f) atomicAdd(colsum + ldsum + i, b);
    }
}

// ------------------------------------------------------------------------------------------------ stem (pool0; conv0 lives in stem_conv.cu)
// 3x3 / stride 2 / pad 1 max pool, NHWC; idx = ky*3+kx of the first maximum (PyTorch tie rule).
// The kernel was bound by instruction issue (ncu: 835 M warp instructions, 650 per 8-channel item, ALU pipe 65 %, DRAM 36 %): a
// compare + two selects + conversions per (tap, channel).  Now every (tap, channel) is ONE signed-integer max on a 32-bit key:
//   key = order(bf16 bits) << 16 | (15 - tap)        order(b) = b ^ (b < 0 ? 0x7fff : 0): bf16 bit patterns in signed-integer order
// so the maximum carries its own arg-max, and among equal values the lowest tap wins (the tie rule) -- 4 instructions per pair.
__device__ __forceinline__ int mp_key(uint32_t w_shifted_or_masked_with_low) {
    const int k = (int)w_shifted_or_masked_with_low;
    return k ^ ((k >> 31) & 0x7fff0000);
}
__global__ void __launch_bounds__(256) maxpool3s2_fwd_kernel(const __nv_bfloat16* __restrict__ in, long ldi, int N, int Hi, int Wi, int C,
                                                              __nv_bfloat16* __restrict__ out, long ldo, unsigned char* __restrict__ idx) {
    gn_pdl_sync();
    const int Ho = Hi / 2, Wo = Wi / 2, G = C / 8;
    const int total = N * Ho * Wo * G;                     // < 2^31 (checked by the launcher): 32-bit index arithmetic
    const uint32_t ld16 = (uint32_t)(ldi >> 3);            // pixel pitch in 16-byte units (checked: ldi % 8 == 0, whole tensor < 2^35 bytes)
    const uint4* in16 = reinterpret_cast<const uint4*>(in);
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int op = e / G, cg = e - op * G;
        const int q = op / Wo, ox = op - q * Wo;
        const int n = q / Ho, oy = q - n * Ho;
        int best[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) best[j] = (int)0x80000000;
        // Only the first row / column of windows reaches outside the image (Hi, Wi even).  Such a tap is redirected to the in-image tap
        // of the same window that reads the same pixel row / column (ky 0 -> 1, kx 0 -> 1) WITH that tap's index: a duplicate key
        // changes nothing, and the loop stays free of divergent branches.
        const int ky_min = oy == 0 ? 1 : 0, kx_min = ox == 0 ? 1 : 0;
        const uint32_t pix0 = (uint32_t)((n * Hi + 2 * oy - 1) * Wi + 2 * ox - 1);      // pixel of tap (0, 0); may wrap, only used with in-image offsets
        // all nine loads first (one round trip to memory per item instead of nine dependent ones), then one sign test for the item
        uint4 t[9];
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int eky = ky == 0 ? ky_min : ky;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ekx = kx == 0 ? kx_min : kx;
                t[ky * 3 + kx] = __ldg(in16 + (size_t)(pix0 + (uint32_t)(eky * Wi + ekx)) * ld16 + cg);
            }
        }
        uint32_t any = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) any |= t[k].x | t[k].y | t[k].z | t[k].w;
        const bool nonneg = (any & 0x80008000u) == 0u;    // always, behind a ReLU: bf16 bit patterns already are in integer order
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int eky = ky == 0 ? ky_min : ky;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ekx = kx == 0 ? kx_min : kx;
                const uint32_t w[4] = {t[ky * 3 + kx].x, t[ky * 3 + kx].y, t[ky * 3 + kx].z, t[ky * 3 + kx].w};
                const uint32_t low = (uint32_t)(15 - (eky * 3 + ekx));
                if (nonneg) {
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        best[2 * h] = max(best[2 * h], (int)__byte_perm(w[h], low, 0x1054));          // (lower bf16 << 16) | low
                        best[2 * h + 1] = max(best[2 * h + 1], (int)__byte_perm(w[h], low, 0x3254));  // (upper bf16 << 16) | low
                    }
                } else {
#pragma unroll
                    for (int h = 0; h < 4; ++h) {
                        best[2 * h] = max(best[2 * h], mp_key(__byte_perm(w[h], low, 0x1054)));
                        best[2 * h + 1] = max(best[2 * h + 1], mp_key(__byte_perm(w[h], low, 0x3254)));
                    }
                }
            }
        }
        uint32_t vb[8], bi[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            vb[j] = (uint32_t)mp_key((uint32_t)best[j]);          // the transform is its own inverse on the upper half
            bi[j] = 15u - ((uint32_t)best[j] & 15u);
        }
        uint4 o;
        o.x = __byte_perm(vb[0], vb[1], 0x7632); o.y = __byte_perm(vb[2], vb[3], 0x7632);
        o.z = __byte_perm(vb[4], vb[5], 0x7632); o.w = __byte_perm(vb[6], vb[7], 0x7632);
        *reinterpret_cast<uint4*>(out + (long)op * ldo + cg * 8) = o;
        uint2 pk;
        pk.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | (bi[3] << 24);
        pk.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | (bi[7] << 24);
        *reinterpret_cast<uint2*>(idx + (long)op * C + cg * 8) = pk;
    }
}

// dZ0[pixel, c] = (sum of dPool over the windows whose arg-max is this pixel) * [act0 > 0] * s0[c]; BN0 column sums.
// act0 is the ACTIVATED stem output: xhat = (act0 - beta) / gamma where act0 > 0.
// Thread layout of the backward pooling kernels: a thread owns ONE group of 8 channels for its whole life (its BatchNorm
// constants and column partial sums stay in registers) and walks pixels with 32-bit index arithmetic; blockDim = G * ppb.
// Column sums: sum g and sum g*ref per channel; xhat is applied once at the end: p1 * (sum g*ref - p0 * sum g).
__device__ __forceinline__ void colsum_thread_flush(float* s_sum, int C, int cg, const float (&sg)[8], const float (&sx)[8],
                                                    const float* __restrict__ p0, const float* __restrict__ p1) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = cg * 8 + j;
        atomicAdd(&s_sum[c], sg[j]);
        atomicAdd(&s_sum[C + c], __ldg(p1 + c) * (sx[j] - __ldg(p0 + c) * sg[j]));
    }
}

// Work unit = the 2x2 input quad {2a, 2a+1} x {2b, 2b+1}: it touches exactly the four windows (a, b), (a, b+1), (a+1, b),
// (a+1, b+1), so every pooled gradient / arg-max byte is loaded once per quad instead of once per input pixel.
// A thread owns FOUR channels (8-byte vectors): with eight the kernel needed 125 registers, ran at 24 % occupancy and was bound by
// load latency (ncu: issue-active 46 %, DRAM 53 %); with four it fits 64 registers and twice as many quads are in flight per SM.
__global__ void __launch_bounds__(256, 4) maxpool3s2_bnrelu_bwd_kernel(const __nv_bfloat16* __restrict__ dpool, long ldp,
                                                                       const unsigned char* __restrict__ idx,
                                                                       const __nv_bfloat16* __restrict__ act, long lda, int N, int Hi, int Wi,
                                                                       int C, const float* __restrict__ sc, const float* __restrict__ p0,
                                                                       const float* __restrict__ p1, __nv_bfloat16* __restrict__ dz, long ldz,
                                                                       float* __restrict__ colsum, int ldsum, int G4, int ppb) {
    gn_pdl_sync();
    extern __shared__ float s_sum[];
    colsum_block_begin(s_sum, C);
    const int Ho = Hi / 2, Wo = Wi / 2;
    const int cg = threadIdx.x % G4, pl = threadIdx.x / G4;      // cg: group of 4 channels
    const float4 s = *reinterpret_cast<const float4*>(sc + cg * 4);
    const float sv[4] = {s.x, s.y, s.z, s.w};
    float sg[4], sx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) { sg[j] = 0.f; sx[j] = 0.f; }
    const int nquads = N * Ho * Wo;
    for (int qd = blockIdx.x * ppb + pl; qd < nquads; qd += gridDim.x * ppb) {
        const int b = qd % Wo, t = qd / Wo, a = t % Ho, n = t / Ho;
        // the four windows; w = wy*2 + wx with (wy, wx) = window offset from (a, b)
        uint2 d[4];          // 4 bf16 gradients per window
        uint32_t pk[4];      // 4 arg-max bytes per window
#pragma unroll
        for (int w = 0; w < 4; ++w) {
            const int oy = a + (w >> 1), ox = b + (w & 1);
            if (oy < Ho && ox < Wo) {
                const long op = ((long)n * Ho + oy) * Wo + ox;
                pk[w] = *reinterpret_cast<const uint32_t*>(idx + op * C + cg * 4);
                d[w] = *reinterpret_cast<const uint2*>(dpool + op * ldp + cg * 4);
            } else {
                pk[w] = 0xffffffffu;                                   // tap 255 never matches
                d[w] = make_uint2(0u, 0u);
            }
        }
        const long ip00 = ((long)n * Hi + 2 * a) * Wi + 2 * b;
        uint2 av4[4];
#pragma unroll
        for (int pi = 0; pi < 4; ++pi) av4[pi] = *reinterpret_cast<const uint2*>(act + (ip00 + (pi >> 1) * Wi + (pi & 1)) * lda + cg * 4);
#pragma unroll
        for (int py = 0; py < 2; ++py)
#pragma unroll
            for (int px = 0; px < 2; ++px) {
                float g[4] = {0.f, 0.f, 0.f, 0.f};
                // pixel (2a+py, 2b+px) lies in window (a+wy, b+wx) at tap ky = py + 1 - 2 wy, kx = px + 1 - 2 wx (wy <= py, wx <= px)
#pragma unroll
                for (int wy = 0; wy <= py; ++wy)
#pragma unroll
                    for (int wx = 0; wx <= px; ++wx) {
                        const int w = wy * 2 + wx;
                        const unsigned k = (unsigned)((py + 1 - 2 * wy) * 3 + (px + 1 - 2 * wx));
                        const float dv[4] = {__uint_as_float(d[w].x << 16), __uint_as_float(d[w].x & 0xffff0000u), __uint_as_float(d[w].y << 16),
                                             __uint_as_float(d[w].y & 0xffff0000u)};
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (((pk[w] >> (8 * j)) & 0xffu) == k) g[j] += dv[j];
                    }
                const uint2 aw = av4[py * 2 + px];
                const float av[4] = {__uint_as_float(aw.x << 16), __uint_as_float(aw.x & 0xffff0000u), __uint_as_float(aw.y << 16),
                                     __uint_as_float(aw.y & 0xffff0000u)};
                float o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float gg = av[j] > 0.f ? g[j] : 0.f;
                    sg[j] += gg;
                    sx[j] = fmaf(gg, av[j], sx[j]);
                    o[j] = gg * sv[j];
                }
                *reinterpret_cast<uint2*>(dz + (ip00 + py * Wi + px) * ldz + cg * 4) = make_uint2(gn_pack_bf16x2(o[0], o[1]), gn_pack_bf16x2(o[2], o[3]));
            }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int c = cg * 4 + j;
        atomicAdd(&s_sum[c], sg[j]);
        atomicAdd(&s_sum[C + c], __ldg(p1 + c) * (sx[j] - __ldg(p0 + c) * sg[j]));
    }
    colsum_block_end(s_sum, C, colsum, ldsum);
}

// ------------------------------------------------------------------------------------------------ transition
// out[n, y, x, c] = 1/4 sum_{2x2} relu(in * sc + sh)
__global__ void __launch_bounds__(256) bnrelu_avgpool2_fwd_kernel(const __nv_bfloat16* __restrict__ in, long ldi, int N, int H, int W, int C,
                                                                   const float* __restrict__ sc, const float* __restrict__ sh,
                                                                   __nv_bfloat16* __restrict__ out, long ldo) {
    gn_pdl_sync();
    const int Ho = H / 2, Wo = W / 2, G = C / 8;
    const int total = N * Ho * Wo * G;                     // < 2^31 (checked by the launcher)
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x) {
        const int op = e / G, cg = e - op * G;
        const int q = op / Wo, ox = op - q * Wo;
        const int n = q / Ho, oy = q - n * Ho;
        const V8 s = ld_f32x8(sc + cg * 8), t = ld_f32x8(sh + cg * 8);
        V8 acc;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
                const V8 v = ld_bf16x8(in + (((long)n * H + 2 * oy + dy) * W + 2 * ox + dx) * ldi + cg * 8);
#pragma unroll
                for (int j = 0; j < 8; ++j) acc.v[j] += fmaxf(fmaf(v.v[j], s.v[j], t.v[j]), 0.f);
            }
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] *= 0.25f;
        st_bf16x8(out + (long)op * ldo + cg * 8, acc);
    }
}

// dC[pixel, c] = dP[pooled pixel, c] / 4 * [raw*sc+sh > 0] * sc   (WRITE: first contribution to the block's gradient buffer)
// pool_div = 4 with (H, W) -> (H/2, W/2) for the transitions; for the head pass Hp = Wp = 1 semantics via `gap` = 1:
//   gap: dP is fp32 [N, C] (gradient of the pooled features), divisor H*W.
template <bool GAP>
__global__ void __launch_bounds__(256, 4) pool_bnrelu_bwd_kernel(const void* __restrict__ dpool, long ldp, const __nv_bfloat16* __restrict__ raw,
                                                               long ldr, int N, int H, int W, int C, const float* __restrict__ sc,
                                                               const float* __restrict__ sh, const float* __restrict__ p0,
                                                               const float* __restrict__ p1, __nv_bfloat16* __restrict__ dC, long ldc,
                                                               float* __restrict__ colsum, int ldsum, int G, int ppb) {
    gn_pdl_sync();
    extern __shared__ float s_sum[];
    colsum_block_begin(s_sum, C);
    const int cg = threadIdx.x % G, pl = threadIdx.x / G;
    const V8 s = ld_f32x8(sc + cg * 8), t = ld_f32x8(sh + cg * 8);
    float sg[8], sx[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { sg[j] = 0.f; sx[j] = 0.f; }
    const int Ho = H / 2, Wo = W / 2;
    const int n_units = GAP ? N : N * Ho * Wo;           // one unit = one pooled value per channel
    const int pix_per_unit = GAP ? H * W : 4;
    const float inv = GAP ? 1.f / (float)(H * W) : 0.25f;
    for (int u = blockIdx.x * ppb + pl; u < n_units; u += gridDim.x * ppb) {
        V8 d;
        long ip0;
        if (GAP) {
            d = ld_f32x8(reinterpret_cast<const float*>(dpool) + (long)u * ldp + cg * 8);
            ip0 = (long)u * H * W;
        } else {
            d = ld_bf16x8(reinterpret_cast<const __nv_bfloat16*>(dpool) + (long)u * ldp + cg * 8);
            const int ox = u % Wo, q = u / Wo, oy = q % Ho, n = q / Ho;
            ip0 = ((long)n * H + 2 * oy) * W + 2 * ox;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) d.v[j] *= inv;
        for (int pi = 0; pi < pix_per_unit; ++pi) {
            const long ip = GAP ? ip0 + pi : ip0 + (pi >> 1) * W + (pi & 1);
            const V8 r = ld_bf16x8(raw + ip * ldr + cg * 8);
            V8 o;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float a = fmaf(r.v[j], s.v[j], t.v[j]);
                const float gg = a > 0.f ? d.v[j] : 0.f;
                sg[j] += gg;
                sx[j] += gg * r.v[j];
                o.v[j] = gg * s.v[j];
            }
            st_bf16x8(dC + ip * ldc + cg * 8, o);
        }
    }
    colsum_thread_flush(s_sum, C, cg, sg, sx, p0, p1);
    colsum_block_end(s_sum, C, colsum, ldsum);
}

// ------------------------------------------------------------------------------------------------ head
// feat[n, c] = mean_{hw} relu(in * sc + sh)     (fp32)
__global__ void __launch_bounds__(256) bnrelu_gap_fwd_kernel(const __nv_bfloat16* __restrict__ in, long ldi, int N, int HW, int C,
                                                              const float* __restrict__ sc, const float* __restrict__ sh,
                                                              float* __restrict__ feat, long ldf) {
    gn_pdl_sync();
    const int G = C / 8;
    const long total = (long)N * G;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int cg = (int)(e % G);
        const long n = e / G;
        const V8 s = ld_f32x8(sc + cg * 8), t = ld_f32x8(sh + cg * 8);
        V8 acc;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc.v[j] = 0.f;
        for (int p = 0; p < HW; ++p) {
            const V8 v = ld_bf16x8(in + (n * HW + p) * ldi + cg * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc.v[j] += fmaxf(fmaf(v.v[j], s.v[j], t.v[j]), 0.f);
        }
        const float inv = 1.f / (float)HW;
#pragma unroll
        for (int j = 0; j < 8; ++j) feat[n * ldf + cg * 8 + j] = acc.v[j] * inv;
    }
}

// logits[n, j] = sum_c feat[n, c] * w[j, c] + b[j]        one warp per spot, J <= 64
__global__ void __launch_bounds__(256) linear_small_fwd_kernel(const float* __restrict__ feat, long ldf, const float* __restrict__ w,
                                                               const float* __restrict__ b, int N, int C, int J, float* __restrict__ out) {
    gn_pdl_sync();
    const int lane = threadIdx.x & 31;
    for (long n = blockIdx.x * (long)(blockDim.x >> 5) + (threadIdx.x >> 5); n < N; n += (long)gridDim.x * (blockDim.x >> 5)) {
        for (int j = 0; j < J; ++j) {
            float acc = 0.f;
            for (int c = lane; c < C; c += 32) acc = fmaf(feat[n * ldf + c], __ldg(w + (long)j * C + c), acc);
            acc = gn_warp_sum(acc);
            if (lane == 0) out[n * J + j] = acc + (b ? b[j] : 0.f);
        }
    }
}
// dfeat[n, c] = sum_j dlog[n, j] * w[j, c]
__global__ void __launch_bounds__(256) linear_small_bwd_data_kernel(const float* __restrict__ dlog, const float* __restrict__ w, int N, int C,
                                                                    int J, float* __restrict__ dfeat, long ldf) {
    gn_pdl_sync();
    const long total = (long)N * C;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        const int c = (int)(e % C);
        const long n = e / C;
        float acc = 0.f;
        for (int j = 0; j < J; ++j) acc = fmaf(__ldg(dlog + n * J + j), __ldg(w + (long)j * C + c), acc);
        dfeat[n * ldf + c] = acc;
    }
}
// dw[j, c] += sum_n dlog[n, j] * feat[n, c];  db[j] += sum_n dlog[n, j]     (grid.y splits the spots)
__global__ void __launch_bounds__(256) linear_small_bwd_weight_kernel(const float* __restrict__ dlog, const float* __restrict__ feat, long ldf,
                                                                      int N, int C, int J, float* __restrict__ dw, float* __restrict__ db) {
    gn_pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const long per = (N + gridDim.y - 1) / gridDim.y;
    const long n0 = blockIdx.y * per, n1 = min((long)N, n0 + per);
    for (int j0 = 0; j0 < J; j0 += 8) {
        float acc[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) acc[q] = 0.f;
        if (c < C) {
            for (long n = n0; n < n1; ++n) {
                const float f = feat[n * ldf + c];
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    if (j0 + q < J) acc[q] = fmaf(__ldg(dlog + n * J + j0 + q), f, acc[q]);
            }
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (j0 + q < J) atomicAdd(dw + (long)(j0 + q) * C + c, acc[q]);
        }
    }
    if (db != nullptr && blockIdx.x == 0 && threadIdx.x < J) {
        float s = 0.f;
        for (long n = n0; n < n1; ++n) s += dlog[n * J + threadIdx.x];
        atomicAdd(db + threadIdx.x, s);
    }
}

// eval-mode BatchNorm constants for a whole list of layers at once: scale, shift, invstd   (C total entries)
__global__ void bn_eval_consts_kernel(const float* __restrict__ gamma, const float* __restrict__ beta, const float* __restrict__ mean,
                                      const float* __restrict__ var, float eps, int C, float* __restrict__ scale, float* __restrict__ shift,
                                      float* __restrict__ invstd, float* __restrict__ inv_gamma) {
    gn_pdl_sync();
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    const float is = 1.f / sqrtf(var[c] + eps);
    const float s = gamma[c] * is;
    scale[c] = s;
    shift[c] = beta[c] - mean[c] * s;
    invstd[c] = is;
    inv_gamma[c] = 1.f / gamma[c];
}

// ------------------------------------------------------------------------------------------------ C-ABI
static inline unsigned grid_for(long items, int mult) {
    long b = (items + 255) / 256;
    const long cap = (long)gn_num_sms() * mult;
    if (b > cap) b = cap;
    return (unsigned)(b < 1 ? 1 : b);
}
// grid whose total thread count is a multiple of G (keeps a thread's channel group fixed in grid-stride loops)
static inline unsigned grid_for_groups(long items, int mult, int G) {
    unsigned g = grid_for(items, mult);
    if ((256 % G) == 0) return g;
    // make 256 * g a multiple of G
    long step = G / std::__gcd((long)256, (long)G);
    g = (unsigned)(((g + step - 1) / step) * step);
    return g;
}

GN_API int gn_maxpool3s2_fwd(const void* in, long ldi, int N, int Hi, int Wi, int C, void* out, long ldo, unsigned char* idx,
                             cudaStream_t stream) {
    GN_REQUIRE(in && out && idx && N > 0 && Hi % 2 == 0 && Wi % 2 == 0 && C % 8 == 0 && ldi % 8 == 0 && ldo % 8 == 0, GN_EINVAL,
               "maxpool3s2_fwd: bad arguments");
    const long total = (long)N * (Hi / 2) * (Wi / 2) * (C / 8);
    GN_REQUIRE(total < (1L << 31) && (long)N * Hi * Wi < (1L << 31), GN_EUNSUPPORTED, "maxpool3s2_fwd: too many elements for 32-bit indexing");
    GN_REQUIRE(((uintptr_t)in & 15) == 0, GN_EALIGN, "maxpool3s2_fwd: input must be 16-byte aligned");
    GN_CUDA(gn_launch(maxpool3s2_fwd_kernel, dim3(grid_for(total, 16)), dim3(256), 0, stream, (const __nv_bfloat16*)in, ldi, N, Hi, Wi, C, (__nv_bfloat16*)out, ldo, idx));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

GN_API int gn_maxpool3s2_bnrelu_bwd(const void* dpool, long ldp, const unsigned char* idx, const void* act, long lda, int N, int Hi, int Wi,
                                    int C, const float* sc, const float* p0, const float* p1, void* dz, long ldz, float* colsum, int ldsum,
                                    cudaStream_t stream) {
    GN_REQUIRE(dpool && idx && act && sc && p0 && p1 && dz && colsum && N > 0 && C % 8 == 0 && C <= 2048, GN_EINVAL,
               "maxpool3s2_bnrelu_bwd: bad arguments");
    const int G = C / 4;           // threads own four channels
    GN_REQUIRE(G <= 256 && (long)N * Hi * Wi < (1L << 31), GN_EUNSUPPORTED, "maxpool3s2_bnrelu_bwd: at most 1024 channels and 2^31 pixels");
    GN_REQUIRE(ldp % 4 == 0 && lda % 4 == 0 && ldz % 4 == 0, GN_EALIGN, "maxpool3s2_bnrelu_bwd: pitches must be multiples of 4 elements");
    const int ppb = 256 / G;
    GN_REQUIRE(Hi % 2 == 0 && Wi % 2 == 0, GN_EINVAL, "maxpool3s2_bnrelu_bwd: odd spatial size");
    unsigned grid = (unsigned)gn_ceil_div((long)N * (Hi / 2) * (Wi / 2), ppb);
    if (grid > (unsigned)gn_num_sms() * 8) grid = gn_num_sms() * 8;
    GN_CUDA(gn_launch(maxpool3s2_bnrelu_bwd_kernel, dim3(grid), dim3(G * ppb), 2 * C * sizeof(float), stream, 
        (const __nv_bfloat16*)dpool, ldp, idx, (const __nv_bfloat16*)act, lda, N, Hi, Wi, C, sc, p0, p1, (__nv_bfloat16*)dz, ldz, colsum, ldsum, G, ppb));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

GN_API int gn_bnrelu_avgpool2_fwd(const void* in, long ldi, int N, int H, int W, int C, const float* sc, const float* sh, void* out, long ldo,
                                  cudaStream_t stream) {
    GN_REQUIRE(in && out && sc && sh && N > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0 && ldi % 8 == 0 && ldo % 8 == 0, GN_EINVAL,
               "bnrelu_avgpool2_fwd: bad arguments");
    const long total = (long)N * (H / 2) * (W / 2) * (C / 8);
    GN_REQUIRE(total < (1L << 31), GN_EUNSUPPORTED, "bnrelu_avgpool2_fwd: too many elements for 32-bit indexing");
    GN_CUDA(gn_launch(bnrelu_avgpool2_fwd_kernel, dim3(grid_for(total, 16)), dim3(256), 0, stream, (const __nv_bfloat16*)in, ldi, N, H, W, C, sc, sh, (__nv_bfloat16*)out, ldo));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// gap = 0: dpool bf16 [N*(H/2)*(W/2), ldp] (transition);  gap = 1: dpool fp32 [N, ldp] (head)
GN_API int gn_pool_bnrelu_bwd(const void* dpool, long ldp, int gap, const void* raw, long ldr, int N, int H, int W, int C, const float* sc,
                              const float* sh, const float* p0, const float* p1, void* dC, long ldc, float* colsum, int ldsum,
                              cudaStream_t stream) {
    GN_REQUIRE(dpool && raw && sc && sh && p0 && p1 && dC && colsum && N > 0 && C % 8 == 0 && C <= 2048, GN_EINVAL, "pool_bnrelu_bwd: bad arguments");
    GN_REQUIRE(gap || (H % 2 == 0 && W % 2 == 0), GN_EINVAL, "pool_bnrelu_bwd: odd spatial size");
    const int G = C / 8;
    GN_REQUIRE(G <= 256 && (long)N * H * W < (1L << 31), GN_EUNSUPPORTED, "pool_bnrelu_bwd: at most 2048 channels and 2^31 pixels");
    const int ppb = 256 / G;
    const long units = gap ? N : (long)N * (H / 2) * (W / 2);
    unsigned grid = (unsigned)gn_ceil_div(units, ppb);
    if (grid > (unsigned)gn_num_sms() * 8) grid = gn_num_sms() * 8;
    const size_t smem = 2 * C * sizeof(float);
    if (gap)
        GN_CUDA(gn_launch(pool_bnrelu_bwd_kernel<true>, dim3(grid), dim3(G * ppb), smem, stream, dpool, ldp, (const __nv_bfloat16*)raw, ldr, N, H, W, C, sc, sh, p0, p1,
                                                                     (__nv_bfloat16*)dC, ldc, colsum, ldsum, G, ppb));
    else
        GN_CUDA(gn_launch(pool_bnrelu_bwd_kernel<false>, dim3(grid), dim3(G * ppb), smem, stream, dpool, ldp, (const __nv_bfloat16*)raw, ldr, N, H, W, C, sc, sh, p0, p1,
                                                                      (__nv_bfloat16*)dC, ldc, colsum, ldsum, G, ppb));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

GN_API int gn_bnrelu_gap_fwd(const void* in, long ldi, int N, int HW, int C, const float* sc, const float* sh, float* feat, long ldf,
                             cudaStream_t stream) {
    GN_REQUIRE(in && sc && sh && feat && N > 0 && HW > 0 && C % 8 == 0 && ldi % 8 == 0, GN_EINVAL, "bnrelu_gap_fwd: bad arguments");
    GN_CUDA(gn_launch(bnrelu_gap_fwd_kernel, dim3(grid_for((long)N * (C / 8), 16)), dim3(256), 0, stream, (const __nv_bfloat16*)in, ldi, N, HW, C, sc, sh, feat, ldf));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

GN_API int gn_linear_small_fwd(const float* feat, long ldf, const float* w, const float* b, int N, int C, int J, float* out, cudaStream_t stream) {
    GN_REQUIRE(feat && w && out && N > 0 && C > 0 && J > 0, GN_EINVAL, "linear_small_fwd: bad arguments");
    GN_CUDA(gn_launch(linear_small_fwd_kernel, dim3(grid_for((long)N * 32, 16)), dim3(256), 0, stream, feat, ldf, w, b, N, C, J, out));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

GN_API int gn_linear_small_bwd(const float* dlog, const float* feat, long ldf, const float* w, int N, int C, int J, float* dfeat, long lddf,
                               float* dw, float* db, cudaStream_t stream) {
    GN_REQUIRE(dlog && feat && w && N > 0 && C > 0 && J > 0 && J <= 256, GN_EINVAL, "linear_small_bwd: bad arguments");
    if (dfeat) {
        GN_CUDA(gn_launch(linear_small_bwd_data_kernel, dim3(grid_for((long)N * C, 16)), dim3(256), 0, stream, dlog, w, N, C, J, dfeat, lddf));
        GN_LAUNCH_CHECK();
    }
    if (dw) {
        int ys = (2 * gn_num_sms()) / gn_ceil_div(C, 256);
        if (ys < 1) ys = 1;
        if (ys > N) ys = N;
        dim3 grid(gn_ceil_div(C, 256), ys);
        GN_CUDA(gn_launch(linear_small_bwd_weight_kernel, dim3(grid), dim3(256), 0, stream, dlog, feat, ldf, N, C, J, dw, db));
        GN_LAUNCH_CHECK();
    }
    return GN_OK;
}

GN_API int gn_bn_eval_consts(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int C, float* scale,
                             float* shift, float* invstd, float* inv_gamma, cudaStream_t stream) {
    GN_REQUIRE(gamma && beta && mean && var && scale && shift && invstd && inv_gamma && C > 0, GN_EINVAL, "bn_eval_consts: bad arguments");
    GN_CUDA(gn_launch(bn_eval_consts_kernel, dim3(gn_ceil_div(C, 256)), dim3(256), 0, stream, gamma, beta, mean, var, eps, C, scale, shift, invstd, inv_gamma));
    GN_LAUNCH_CHECK();
    return GN_OK;
}
