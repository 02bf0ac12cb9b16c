// Hexagonal convolution over Visium odd-r grids (fp32, NCHW), forward / data-grad / weight-grad.
//
// Replaces hexagdly.Conv2d as used by the reference's g network
// (/root/reference/gridnext/gridnet_models.py:128-148) AND the rot90/flip re-indexing around it
// (gridnet_models.py:177-185): the kernels work directly in the Visium (H=78, W=64) layout, where
// the hexagonal neighbourhood depends on the parity of the ROW index y:
//
//   out[b,o,y,x] = bias[o] + sum_{c} sum_{tap t=(i,side,a)} kernel_i[o,c,a,side] *
//                  in[b,c, y + dy_t, x + dx_t(y&1)]
//   dy_t = 0 (i=0) | -i (side 0) | +i (side 1);  dx_t(p) = a - (k - i/2) + (i&1)*p
//
// Weights are repacked once per call into Wp[t][cin][cout] (cout contiguous) so that the inner
// loop reads them as broadcast float4s.  The data gradient is the same convolution applied to dY
// with point-reflected, channel-transposed kernels (pack mode 1), so one kernel serves both.
//
// Optional fusions (both used by the corrector):
//   * prologue: in' = relu(in * in_scale[c] + in_shift[c])  (BatchNorm2d-apply + ReLU of the
//     previous layer, gridnet_models.py:134-136,142-144), padding stays exactly zero;
//   * epilogue: per-output-channel sum / sum-of-squares accumulated in fp64 for the next
//     BatchNorm2d's batch statistics.
#include "gn_common.cuh"

#define HEX_MAX_K 3
#define HEX_MAX_TAPS 37

struct HexTaps {
    int n;
    int k;
    int square;                      // 1: Cartesian K x K window (K = 2k+1), one weight tensor (Cout, Cin, K, K)
    signed char dy[HEX_MAX_TAPS];
    signed char dxe[HEX_MAX_TAPS];
    signed char dxo[HEX_MAX_TAPS];
    signed char ki[HEX_MAX_TAPS];    // which kernel_i
    signed char ka[HEX_MAX_TAPS];    // row tap a
    signed char kside[HEX_MAX_TAPS]; // side
};

static HexTaps make_taps(int k) {
    HexTaps t;
    memset(&t, 0, sizeof(t));
    t.k = k;
    int n = 0;
    for (int i = 0; i <= k; ++i)
        for (int side = 0; side < (i == 0 ? 1 : 2); ++side)
            for (int a = 0; a < 2 * k + 1 - i; ++a) {
                t.dy[n] = (signed char)(i == 0 ? 0 : (side ? i : -i));
                int base = a - (k - i / 2);
                t.dxe[n] = (signed char)base;
                t.dxo[n] = (signed char)(base + (i & 1));
                t.ki[n] = (signed char)i;
                t.ka[n] = (signed char)a;
                t.kside[n] = (signed char)side;
                ++n;
            }
    t.n = n;
    return t;
}

// Cartesian K x K, stride 1, padding K/2 (the base GridNet corrector's nn.Conv2d layers, gridnet_models.py:51-66): the same
// tile kernels with a parity-independent tap table; the weight tensor plays the role of kernel_0 with (a, side) = (row, col).
static HexTaps make_square_taps(int K) {
    HexTaps t;
    memset(&t, 0, sizeof(t));
    t.k = K / 2;
    t.square = 1;
    int n = 0;
    for (int r = 0; r < K; ++r)
        for (int c = 0; c < K; ++c) {
            t.dy[n] = (signed char)(r - K / 2);
            t.dxe[n] = t.dxo[n] = (signed char)(c - K / 2);
            t.ki[n] = 0;
            t.ka[n] = (signed char)r;
            t.kside[n] = (signed char)c;
            ++n;
        }
    t.n = n;
    return t;
}

struct HexKernelPtrs {
    const float* k[HEX_MAX_K + 1];
};
struct HexKernelPtrsMut {
    float* k[HEX_MAX_K + 1];
};

// ------------------------------------------------------------------------------------------------
// weight (un)packing
// mode 0: Wp[t][ci][co] = K_i[co][ci][a][side]                      (conv cin=Cin,  cout=Cout)
// mode 1: Wp[t][o ][c ] = K_i[o ][c ][na-1-a][ns-1-side]            (conv cin=Cout, cout=Cin)
__global__ void hex_pack_kernel(HexKernelPtrs kp, HexTaps taps, int Cin, int Cout, int mode, float* __restrict__ wp) {
    gn_pdl_sync();
    long total = (long)taps.n * Cin * Cout;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        int t = (int)(e / ((long)Cin * Cout));
        int r = (int)(e % ((long)Cin * Cout));
        int i = taps.ki[t], a = taps.ka[t], side = taps.kside[t];
        int na = 2 * taps.k + 1 - i, ns = taps.square ? na : ((i == 0) ? 1 : 2);
        int co, ci;
        if (mode == 0) {
            ci = r / Cout; co = r % Cout;
        } else {
            co = r / Cin; ci = r % Cin;      // packed "input" channel is the layer's output channel
            a = na - 1 - a;
            side = ns - 1 - side;
        }
        wp[e] = kp.k[i][(((long)co * Cin + ci) * na + a) * ns + side];
    }
}

// dK_i[co][ci][a][side] = dWp[t][ci][co]
__global__ void hex_unpack_grad_kernel(const float* __restrict__ dwp, HexTaps taps, int Cin, int Cout, HexKernelPtrsMut kp) {
    gn_pdl_sync();
    long total = (long)taps.n * Cin * Cout;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        int t = (int)(e / ((long)Cin * Cout));
        int r = (int)(e % ((long)Cin * Cout));
        int ci = r / Cout, co = r % Cout;
        int i = taps.ki[t], a = taps.ka[t], side = taps.kside[t];
        int na = 2 * taps.k + 1 - i, ns = taps.square ? na : ((i == 0) ? 1 : 2);
        kp.k[i][(((long)co * Cin + ci) * na + a) * ns + side] = dwp[e];
    }
}

// ------------------------------------------------------------------------------------------------
// forward / data-grad
#define HEX_TR 8       // rows per tile (each thread owns 2 adjacent rows: one even, one odd)
#define HEX_TW 64      // columns per tile
#define HEX_CC 8       // input channels staged per chunk
#define HEX_THREADS 256

template <int CO>
__global__ void __launch_bounds__(HEX_THREADS)
hexconv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wp, const float* __restrict__ bias,
                   const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                   float* __restrict__ y, double* __restrict__ stats,
                   int Cin, int Cout, int H, int W, HexTaps taps) {
    gn_pdl_sync();
    extern __shared__ __align__(16) float smem[];
    const int k = taps.k;
    const int SR = HEX_TR + 2 * k;          // staged rows
    const int SW = HEX_TW + 2 * k;          // staged cols
    float* s_in = smem;                                       // [CC][SR][SW]
    float* s_w = smem + HEX_CC * SR * SW;                     // [T][CC][CO]
    __shared__ double s_stat[2 * 32];

    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int n_xt = (W + HEX_TW - 1) / HEX_TW;
    const int x0 = (blockIdx.x % n_xt) * HEX_TW;
    const int co0 = (blockIdx.x / n_xt) * CO;
    const int y0 = blockIdx.y * HEX_TR;
    const int lx = tid % HEX_TW;
    const int lr = (tid / HEX_TW) * 2;      // local even row; lr+1 is the odd one (y0 is even)

    float acc0[CO], acc1[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) { acc0[o] = 0.f; acc1[o] = 0.f; }

    const float* xb = x + (long)b * Cin * H * W;
    for (int c0 = 0; c0 < Cin; c0 += HEX_CC) {
        __syncthreads();
        // stage the input chunk (zero outside the grid; optional BN+ReLU prologue)
        for (int e = tid; e < HEX_CC * SR * SW; e += HEX_THREADS) {
            int c = e / (SR * SW);
            int r = (e / SW) % SR;
            int cx = e % SW;
            int gy = y0 + r - k, gx = x0 + cx - k, gc = c0 + c;
            float v = 0.f;
            if (gc < Cin && gy >= 0 && gy < H && gx >= 0 && gx < W) {
                v = __ldg(xb + ((long)gc * H + gy) * W + gx);
                if (in_scale != nullptr) v = fmaxf(fmaf(v, __ldg(in_scale + gc), __ldg(in_shift + gc)), 0.f);
            }
            s_in[e] = v;
        }
        // stage the weight chunk: s_w[t][c][o] = wp[t][c0+c][co0+o]
        for (int e = tid; e < taps.n * HEX_CC * CO; e += HEX_THREADS) {
            int t = e / (HEX_CC * CO);
            int c = (e / CO) % HEX_CC;
            int o = e % CO;
            float v = 0.f;
            if (c0 + c < Cin && co0 + o < Cout) v = __ldg(wp + ((long)t * Cin + (c0 + c)) * Cout + co0 + o);
            s_w[e] = v;
        }
        __syncthreads();
        const int cmax = min(HEX_CC, Cin - c0);
        for (int c = 0; c < cmax; ++c) {
            const float* sc = s_in + c * SR * SW;
            for (int t = 0; t < taps.n; ++t) {
                const int dy = taps.dy[t];
                const float v0 = sc[(lr + k + dy) * SW + lx + k + taps.dxe[t]];
                const float v1 = sc[(lr + 1 + k + dy) * SW + lx + k + taps.dxo[t]];
                const float4* w4 = reinterpret_cast<const float4*>(s_w + (t * HEX_CC + c) * CO);
#pragma unroll
                for (int o4 = 0; o4 < CO / 4; ++o4) {
                    const float4 w = w4[o4];
                    acc0[4 * o4 + 0] = fmaf(v0, w.x, acc0[4 * o4 + 0]);
                    acc0[4 * o4 + 1] = fmaf(v0, w.y, acc0[4 * o4 + 1]);
                    acc0[4 * o4 + 2] = fmaf(v0, w.z, acc0[4 * o4 + 2]);
                    acc0[4 * o4 + 3] = fmaf(v0, w.w, acc0[4 * o4 + 3]);
                    acc1[4 * o4 + 0] = fmaf(v1, w.x, acc1[4 * o4 + 0]);
                    acc1[4 * o4 + 1] = fmaf(v1, w.y, acc1[4 * o4 + 1]);
                    acc1[4 * o4 + 2] = fmaf(v1, w.z, acc1[4 * o4 + 2]);
                    acc1[4 * o4 + 3] = fmaf(v1, w.w, acc1[4 * o4 + 3]);
                }
            }
        }
    }

    // epilogue: bias, store, optional BN statistics
    const int gx = x0 + lx;
    const int gy0 = y0 + lr, gy1 = gy0 + 1;
    const bool ok0 = gx < W && gy0 < H, ok1 = gx < W && gy1 < H;
    if (stats != nullptr) {
        if (tid < 2 * 32) s_stat[tid] = 0.0;
        __syncthreads();
    }
#pragma unroll
    for (int o = 0; o < CO; ++o) {
        const int go = co0 + o;
        if (go < Cout) {   // warp-uniform
            const float bv = bias != nullptr ? __ldg(bias + go) : 0.f;
            const float r0 = acc0[o] + bv, r1 = acc1[o] + bv;
            float* yo = y + (((long)b * Cout + go) * H) * W;
            if (ok0) yo[(long)gy0 * W + gx] = r0;
            if (ok1) yo[(long)gy1 * W + gx] = r1;
            if (stats != nullptr) {
                float s = (ok0 ? r0 : 0.f) + (ok1 ? r1 : 0.f);
                float q = (ok0 ? r0 * r0 : 0.f) + (ok1 ? r1 * r1 : 0.f);
                s = gn_warp_sum(s);
                q = gn_warp_sum(q);
                if ((tid & 31) == 0) {
                    atomicAdd(&s_stat[o], (double)s);
                    atomicAdd(&s_stat[32 + o], (double)q);
                }
            }
        }
    }
    if (stats != nullptr) {
        __syncthreads();
        if (tid < CO && co0 + tid < Cout) {
            atomicAdd(stats + co0 + tid, s_stat[tid]);
            atomicAdd(stats + Cout + co0 + tid, s_stat[32 + tid]);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// weight grad: dWp[t][c][o] += sum_{b,y,x} dY[b,o,y,x] * X'[b,c,y+dy_t,x+dx_t(y&1)],  dbias[o] += sum dY
#define HEXW_CO 32
#define HEXW_MAXI 10   // work items per thread: ceil(37 taps * 8 ch * 8 cout-groups / 256)

__global__ void __launch_bounds__(HEX_THREADS)
hexconv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ in_scale, const float* __restrict__ in_shift,
                     const float* __restrict__ dy, float* __restrict__ dwp, float* __restrict__ dbias,
                     int B, int Cin, int Cout, int H, int W, HexTaps taps) {
    gn_pdl_sync();
    extern __shared__ __align__(16) float smem[];
    const int k = taps.k;
    const int SR = HEX_TR + 2 * k;
    const int SW = HEX_TW + 2 * k;
    float* s_in = smem;                                     // [CC][SR][SW]
    float* s_dy = smem + HEX_CC * SR * SW;                  // [TR*TW][HEXW_CO]
    __shared__ float s_db[HEXW_CO];

    const int tid = threadIdx.x;
    const int c0 = blockIdx.y * HEX_CC;
    const int co0 = blockIdx.z * HEXW_CO;
    const int n_xt = (W + HEX_TW - 1) / HEX_TW;
    const int n_yt = (H + HEX_TR - 1) / HEX_TR;
    const long n_tiles = (long)B * n_yt * n_xt;
    const int OG = HEXW_CO / 4;
    const int n_items = taps.n * HEX_CC * OG;

    float acc[HEXW_MAXI][4];
#pragma unroll
    for (int j = 0; j < HEXW_MAXI; ++j) { acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f; }
    float db_acc = 0.f;
    const bool do_bias = (dbias != nullptr) && (blockIdx.y == 0);

    for (long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int b = (int)(tile / (n_yt * n_xt));
        const int y0 = (int)((tile / n_xt) % n_yt) * HEX_TR;
        const int x0 = (int)(tile % n_xt) * HEX_TW;
        __syncthreads();
        const float* xb = x + (long)b * Cin * H * W;
        for (int e = tid; e < HEX_CC * SR * SW; e += HEX_THREADS) {
            int c = e / (SR * SW);
            int r = (e / SW) % SR;
            int cx = e % SW;
            int gy = y0 + r - k, gx = x0 + cx - k, gc = c0 + c;
            float v = 0.f;
            if (gc < Cin && gy >= 0 && gy < H && gx >= 0 && gx < W) {
                v = __ldg(xb + ((long)gc * H + gy) * W + gx);
                if (in_scale != nullptr) v = fmaxf(fmaf(v, __ldg(in_scale + gc), __ldg(in_shift + gc)), 0.f);
            }
            s_in[e] = v;
        }
        const float* dyb = dy + (long)b * Cout * H * W;
        for (int e = tid; e < HEXW_CO * HEX_TR * HEX_TW; e += HEX_THREADS) {
            int o = e / (HEX_TR * HEX_TW);
            int pix = e % (HEX_TR * HEX_TW);
            int gy = y0 + pix / HEX_TW, gx = x0 + pix % HEX_TW, go = co0 + o;
            float v = 0.f;
            if (go < Cout && gy < H && gx < W) v = __ldg(dyb + ((long)go * H + gy) * W + gx);
            s_dy[pix * HEXW_CO + o] = v;
        }
        __syncthreads();
        if (do_bias) {
            const int o = tid % HEXW_CO;
            for (int pix = tid / HEXW_CO; pix < HEX_TR * HEX_TW; pix += HEX_THREADS / HEXW_CO) db_acc += s_dy[pix * HEXW_CO + o];
        }
#pragma unroll
        for (int j = 0; j < HEXW_MAXI; ++j) {
            const int id = tid + j * HEX_THREADS;
            if (id < n_items) {
                const int og = id % OG;
                const int c = (id / OG) % HEX_CC;
                const int t = id / (OG * HEX_CC);
                const int tdy = taps.dy[t], dxe = taps.dxe[t], dxo = taps.dxo[t];
                const float* sc = s_in + c * SR * SW;
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
                for (int r = 0; r < HEX_TR; ++r) {
                    const float* row = sc + (r + k + tdy) * SW + k + ((r & 1) ? dxo : dxe);
                    const float4* d4 = reinterpret_cast<const float4*>(s_dy + (r * HEX_TW) * HEXW_CO) + og;
#pragma unroll 8
                    for (int xx = 0; xx < HEX_TW; ++xx) {
                        const float v = row[xx];
                        const float4 d = d4[xx * (HEXW_CO / 4)];
                        a0 = fmaf(v, d.x, a0); a1 = fmaf(v, d.y, a1); a2 = fmaf(v, d.z, a2); a3 = fmaf(v, d.w, a3);
                    }
                }
                acc[j][0] += a0; acc[j][1] += a1; acc[j][2] += a2; acc[j][3] += a3;
            }
        }
    }
    // flush
#pragma unroll
    for (int j = 0; j < HEXW_MAXI; ++j) {
        const int id = tid + j * HEX_THREADS;
        if (id < n_items) {
            const int og = id % OG;
            const int c = (id / OG) % HEX_CC;
            const int t = id / (OG * HEX_CC);
            if (c0 + c < Cin) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int go = co0 + og * 4 + q;
                    if (go < Cout) atomicAdd(dwp + ((long)t * Cin + c0 + c) * Cout + go, acc[j][q]);
                }
            }
        }
    }
    if (do_bias) {
        if (tid < HEXW_CO) s_db[tid] = 0.f;
        __syncthreads();
        atomicAdd(&s_db[tid % HEXW_CO], db_acc);
        __syncthreads();
        if (tid < HEXW_CO && co0 + tid < Cout) atomicAdd(dbias + co0 + tid, s_db[tid]);
    }
}

// ------------------------------------------------------------------------------------------------
// C-ABI
GN_API int gn_hexconv_n_taps(int ksize) { return 1 + 3 * ksize * (ksize + 1); }

GN_API int gn_hexconv_pack(const float* k0, const float* k1, const float* k2, const float* k3, int ksize, int cin, int cout,
                           int mode, float* wp, cudaStream_t stream) {
    GN_REQUIRE(ksize >= 1 && ksize <= HEX_MAX_K, GN_EUNSUPPORTED, "hexconv: kernel_size %d not in 1..3", ksize);
    GN_REQUIRE(cin > 0 && cout > 0 && wp && k0 && k1, GN_EINVAL, "hexconv_pack: bad arguments");
    GN_REQUIRE((ksize < 2 || k2) && (ksize < 3 || k3), GN_EINVAL, "hexconv_pack: missing kernel pointer");
    HexKernelPtrs kp = {{k0, k1, k2, k3}};
    HexTaps taps = make_taps(ksize);
    long total = (long)taps.n * cin * cout;
    GN_CUDA(gn_launch(hex_pack_kernel, dim3(gn_ceil_div(total, 256) > 1024 ? 1024 : gn_ceil_div(total, 256)), dim3(256), 0, stream, kp, taps, cin, cout, mode, wp));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

GN_API int gn_hexconv_unpack_grad(const float* dwp, float* dk0, float* dk1, float* dk2, float* dk3, int ksize, int cin, int cout,
                                  cudaStream_t stream) {
    GN_REQUIRE(ksize >= 1 && ksize <= HEX_MAX_K, GN_EUNSUPPORTED, "hexconv: kernel_size %d not in 1..3", ksize);
    GN_REQUIRE(cin > 0 && cout > 0 && dwp && dk0 && dk1, GN_EINVAL, "hexconv_unpack_grad: bad arguments");
    HexKernelPtrsMut kp = {{dk0, dk1, dk2, dk3}};
    HexTaps taps = make_taps(ksize);
    long total = (long)taps.n * cin * cout;
    GN_CUDA(gn_launch(hex_unpack_grad_kernel, dim3(gn_ceil_div(total, 256) > 1024 ? 1024 : gn_ceil_div(total, 256)), dim3(256), 0, stream, dwp, taps, cin, cout, kp));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

template <int CO>
static int launch_fwd(const float* x, const float* wp, const float* bias, const float* in_scale, const float* in_shift, float* y,
                      double* stats, int B, int cin, int cout, int H, int W, const HexTaps& taps, cudaStream_t stream) {
    const int k = taps.k;
    size_t smem = ((size_t)HEX_CC * (HEX_TR + 2 * k) * (HEX_TW + 2 * k) + (size_t)taps.n * HEX_CC * CO) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        GN_CUDA(cudaFuncSetAttribute(hexconv_fwd_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
        attr_set = true;
    }
    const int n_xt = gn_ceil_div(W, HEX_TW), n_ct = gn_ceil_div(cout, CO);
    dim3 grid(n_xt * n_ct, gn_ceil_div(H, HEX_TR), B);
    GN_CUDA(gn_launch(hexconv_fwd_kernel<CO>, grid, dim3(HEX_THREADS), smem, stream, x, wp, bias, in_scale, in_shift, y, stats, cin, cout, H, W, taps));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

static int conv_fwd_impl(const float* x, const float* wp, const float* bias, const float* in_scale, const float* in_shift, float* y,
                         double* stats, int B, int cin, int cout, int H, int W, const HexTaps& taps, cudaStream_t stream) {
    GN_REQUIRE(x && wp && y && B > 0 && cin > 0 && cout > 0 && H > 0 && W > 0, GN_EINVAL, "hexconv_fwd: bad arguments");
    GN_REQUIRE(B <= 65535, GN_EUNSUPPORTED, "hexconv_fwd: batch %d > 65535", B);
    GN_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), GN_EINVAL, "hexconv_fwd: in_scale/in_shift must come together");
    if (cout <= 8) return launch_fwd<8>(x, wp, bias, in_scale, in_shift, y, stats, B, cin, cout, H, W, taps, stream);
    if (cout <= 16) return launch_fwd<16>(x, wp, bias, in_scale, in_shift, y, stats, B, cin, cout, H, W, taps, stream);
    return launch_fwd<32>(x, wp, bias, in_scale, in_shift, y, stats, B, cin, cout, H, W, taps, stream);
}

GN_API int gn_hexconv_fwd(const float* x, const float* wp, const float* bias, const float* in_scale, const float* in_shift, float* y,
                          double* stats, int B, int cin, int cout, int H, int W, int ksize, cudaStream_t stream) {
    GN_REQUIRE(ksize >= 1 && ksize <= HEX_MAX_K, GN_EUNSUPPORTED, "hexconv: kernel_size %d not in 1..3", ksize);
    return conv_fwd_impl(x, wp, bias, in_scale, in_shift, y, stats, B, cin, cout, H, W, make_taps(ksize), stream);
}

static int conv_wgrad_impl(const float* x, const float* in_scale, const float* in_shift, const float* dy, float* dwp, float* dbias,
                           int B, int cin, int cout, int H, int W, const HexTaps& taps, cudaStream_t stream) {
    GN_REQUIRE(x && dy && dwp && B > 0 && cin > 0 && cout > 0 && H > 0 && W > 0, GN_EINVAL, "hexconv_wgrad: bad arguments");
    GN_REQUIRE((in_scale == nullptr) == (in_shift == nullptr), GN_EINVAL, "hexconv_wgrad: in_scale/in_shift must come together");
    const int k = taps.k;
    size_t smem = ((size_t)HEX_CC * (HEX_TR + 2 * k) * (HEX_TW + 2 * k) + (size_t)HEX_TR * HEX_TW * HEXW_CO) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
        GN_CUDA(cudaFuncSetAttribute(hexconv_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024));
        attr_set = true;
    }
    const int gy = gn_ceil_div(cin, HEX_CC), gz = gn_ceil_div(cout, HEXW_CO);
    const long n_tiles = (long)B * gn_ceil_div(H, HEX_TR) * gn_ceil_div(W, HEX_TW);
    long gx = (2L * gn_num_sms() + gy * gz - 1) / (gy * gz);
    if (gx > n_tiles) gx = n_tiles;
    if (gx < 1) gx = 1;
    dim3 grid((unsigned)gx, gy, gz);
    GN_CUDA(gn_launch(hexconv_wgrad_kernel, grid, dim3(HEX_THREADS), smem, stream, x, in_scale, in_shift, dy, dwp, dbias, B, cin, cout, H, W, taps));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

GN_API int gn_hexconv_wgrad(const float* x, const float* in_scale, const float* in_shift, const float* dy, float* dwp, float* dbias,
                            int B, int cin, int cout, int H, int W, int ksize, cudaStream_t stream) {
    GN_REQUIRE(ksize >= 1 && ksize <= HEX_MAX_K, GN_EUNSUPPORTED, "hexconv: kernel_size %d not in 1..3", ksize);
    return conv_wgrad_impl(x, in_scale, in_shift, dy, dwp, dbias, B, cin, cout, H, W, make_taps(ksize), stream);
}

// ---- Cartesian K x K (K in {1, 3, 5}), stride 1, zero padding K/2: nn.Conv2d of the base GridNet corrector
// (/root/reference/gridnext/gridnet_models.py:51-66).  Same packed layout Wp[t = r*K + c][cin][cout], same fusions.
#define SQ_OK(K) ((K) == 1 || (K) == 3 || (K) == 5)

GN_API int gn_sqconv_pack(const float* w, int K, int cin, int cout, int mode, float* wp, cudaStream_t stream) {
    GN_REQUIRE(SQ_OK(K), GN_EUNSUPPORTED, "sqconv: kernel size %d not in {1, 3, 5}", K);
    GN_REQUIRE(cin > 0 && cout > 0 && wp && w, GN_EINVAL, "sqconv_pack: bad arguments");
    HexKernelPtrs kp = {{w, nullptr, nullptr, nullptr}};
    HexTaps taps = make_square_taps(K);
    long total = (long)taps.n * cin * cout;
    GN_CUDA(gn_launch(hex_pack_kernel, dim3(gn_ceil_div(total, 256) > 1024 ? 1024 : gn_ceil_div(total, 256)), dim3(256), 0, stream, kp, taps, cin, cout, mode, wp));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

GN_API int gn_sqconv_unpack_grad(const float* dwp, float* dw, int K, int cin, int cout, cudaStream_t stream) {
    GN_REQUIRE(SQ_OK(K), GN_EUNSUPPORTED, "sqconv: kernel size %d not in {1, 3, 5}", K);
    GN_REQUIRE(cin > 0 && cout > 0 && dwp && dw, GN_EINVAL, "sqconv_unpack_grad: bad arguments");
    HexKernelPtrsMut kp = {{dw, nullptr, nullptr, nullptr}};
    HexTaps taps = make_square_taps(K);
    long total = (long)taps.n * cin * cout;
    GN_CUDA(gn_launch(hex_unpack_grad_kernel, dim3(gn_ceil_div(total, 256) > 1024 ? 1024 : gn_ceil_div(total, 256)), dim3(256), 0, stream, dwp, taps, cin, cout, kp));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

GN_API int gn_sqconv_fwd(const float* x, const float* wp, const float* bias, const float* in_scale, const float* in_shift, float* y,
                         double* stats, int B, int cin, int cout, int H, int W, int K, cudaStream_t stream) {
    GN_REQUIRE(SQ_OK(K), GN_EUNSUPPORTED, "sqconv: kernel size %d not in {1, 3, 5}", K);
    return conv_fwd_impl(x, wp, bias, in_scale, in_shift, y, stats, B, cin, cout, H, W, make_square_taps(K), stream);
}

GN_API int gn_sqconv_wgrad(const float* x, const float* in_scale, const float* in_shift, const float* dy, float* dwp, float* dbias,
                           int B, int cin, int cout, int H, int W, int K, cudaStream_t stream) {
    GN_REQUIRE(SQ_OK(K), GN_EUNSUPPORTED, "sqconv: kernel size %d not in {1, 3, 5}", K);
    return conv_wgrad_impl(x, in_scale, in_shift, dy, dwp, dbias, B, cin, cout, H, W, make_square_taps(K), stream);
}
