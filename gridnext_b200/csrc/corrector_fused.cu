// The whole g corrector in ONE launch per direction (small and medium batches).
//
// Reference: the nn.Sequential of GridNetHex._init_corrector (/root/reference/gridnext/gridnet_models.py:128-148):
//   hexagdly.Conv2d hexagdly.Conv2d [BatchNorm2d] ReLU hexagdly.Conv2d hexagdly.Conv2d [BatchNorm2d] ReLU hexagdly.Conv2d,
// kernel_size 1, 32 channels wide, evaluated per grid cell with train-mode batch statistics over B*H*W cells.
//
// Why one kernel: one Visium array through g is 4,992 cells x 32 channels -- 0.64 MB per activation, 36 MMAC per layer.  Launched
// layer by layer (pack, conv, BN finalize, ... ~40 launches forward+backward) the step is launch- and tail-bound: 1.15 ms of the
// 1.42 ms count-model step (BASELINE configs[0]) while the arithmetic needs a few microseconds.  Here every stage is a phase of
// a persistent grid (one CTA per SM, all co-resident: cooperative launch) separated by grid-wide barriers; activations travel
// through L2; BatchNorm statistics are reduced between phases; weights are read straight from the parameter tensors
// (kernel0 (Cout, Cin, 3, 1), kernel1 (Cout, Cin, 2, 2): no pack / unpack launches) and weight gradients are written back in
// that layout.  FP32 FMA throughout (exact fp32 semantics, <= 1e-5 against the oracle); large batches, where the convolution is
// throughput-bound, run on the tensor-core kernel instead (hexconv_tc.cu).
//
// Work unit = one grid row (b, y) x 64 columns.  Forward / data gradient: thread = (column x, group of 8 output channels);
// input rows y-1..y+1 staged in shared memory [c][row][x] (x contiguous: conflict-free), weights [tap][c][o] read as broadcast
// float4s.  Weight gradient: thread = (tap, 4 input channels, 8 output channels) accumulating over the 64 columns from
// channel-contiguous copies ([row][x][c], 16-byte chunks XOR-swizzled by x; [x][o]) -- 3 shared loads per 32 FMAs -- and kept in
// registers across all units of the phase; flushed with float4 atomics into a packed [tap][c][o] buffer, un-packed by the last
// phase.
//
// Hex geometry in the Visium layout (parity p = y & 1): taps 0..2 = row y, x-1 / x / x+1 (kernel0[a]);
// taps 3,4 = row y-1, x-1+p / x+p (kernel1[a][0]); taps 5,6 = row y+1, x-1+p / x+p (kernel1[a][1]).  The data gradient is the
// same convolution with point-reflected, channel-transposed weights.
#include "gn_common.cuh"

#define CF_MAX_STAGES 8
#define CF_C 32                 // widest channel count (padded in shared memory)
#define CF_T 7
#define CF_THREADS 256
#define CF_TW 64
#define CF_SW (CF_TW + 2)

struct CorrFusedArgs {
    int B, H, W, L, n_xt;
    int bn_training;                         // BatchNorm layers use batch statistics (and update the running ones)
    int cin[CF_MAX_STAGES], cout[CF_MAX_STAGES];
    int pro[CF_MAX_STAGES];                  // transform of stage j's INPUT: 0 none, 1 ReLU, 2 BatchNorm2d + ReLU
    const float* k0[CF_MAX_STAGES];
    const float* k1[CF_MAX_STAGES];
    const float* bias[CF_MAX_STAGES];
    const float* gamma[CF_MAX_STAGES];       // BatchNorm in front of stage j (pro[j] == 2)
    const float* beta[CF_MAX_STAGES];
    float* rmean[CF_MAX_STAGES];
    float* rvar[CF_MAX_STAGES];
    float momentum[CF_MAX_STAGES], eps[CF_MAX_STAGES];
    const float* x;                          // (B, cin[0], H, W)
    float* act[CF_MAX_STAGES];               // act[j] = output of stage j, (B, cout[j], H, W)
    double* stats;                           // [L][64]: sum / sum of squares of act[j-1] per channel (slot j), zeroed by the host
    float* consts;                           // [L][128]: scale, shift, mean, invstd of the BatchNorm in front of stage j
    unsigned* sync;                          // [2] grid barrier counter + exit counter (self-resetting)
    // backward only
    const float* dout;                       // (B, cout[L-1], H, W)
    float* gb[CF_MAX_STAGES];                // gb[j] = gradient w.r.t. stage j's convolution input (after its transform); gb[0] may be null
    float* dwp;                              // [L][7][32][32] packed weight gradients, zeroed by the host
    float* dbias_acc;                        // [L][32], zeroed by the host
    double* sums;                            // [L][64]: sum g, sum g*xhat of the BatchNorm in front of stage j, zeroed by the host
    float* dk0[CF_MAX_STAGES];
    float* dk1[CF_MAX_STAGES];
    float* dbias[CF_MAX_STAGES];
    float* dgamma[CF_MAX_STAGES];
    float* dbeta[CF_MAX_STAGES];
};

__device__ __forceinline__ unsigned cf_ld_acquire(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// All CTAs are co-resident (cooperative launch, grid <= SMs x occupancy).  Barrier k waits for (k+1) * gridDim.x arrivals on a
// monotonic counter; the counters are reset by the last CTA to leave the kernel.
__device__ __forceinline__ void cf_grid_barrier(unsigned* sync, unsigned& epoch) {
    __syncthreads();
    if (threadIdx.x == 0) {
        ++epoch;
        __threadfence();
        atomicAdd(sync, 1u);
        const unsigned target = epoch * gridDim.x;
        while (cf_ld_acquire(sync) < target) {}
        __threadfence();
    }
    __syncthreads();
}

__device__ __forceinline__ void cf_exit(unsigned* sync) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(sync + 1, 1u) == gridDim.x - 1) {
            sync[0] = 0u;
            sync[1] = 0u;
            __threadfence();
        }
    }
}

// s_w[t][ci][co] (co contiguous, CF_C wide, zero padded) for the forward (mode 0: ci = layer input channel) or the data
// gradient (mode 1: ci = layer OUTPUT channel, taps point-reflected).
// (The 7168 / 256 = 28 elements per thread are fetched seven at a time -- independent loads in flight -- because a phase is only
// a few microseconds long and a dependent chain of 28 L2 round trips would cost more than its arithmetic.)
__device__ void cf_stage_weights(float* s_w, const float* __restrict__ k0, const float* __restrict__ k1, int Cin, int Cout, int mode) {
    for (int e0 = threadIdx.x; e0 < CF_T * CF_C * CF_C; e0 += 7 * CF_THREADS) {
        float v[7];
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            const int e = e0 + q * CF_THREADS;
            const int t = e / (CF_C * CF_C), ci = (e / CF_C) % CF_C, co = e % CF_C;
            int o, c;                      // layer output / input channel of this element
            if (mode == 0) { c = ci; o = co; } else { o = ci; c = co; }
            v[q] = 0.f;
            if (o < Cout && c < Cin) {
                if (t < 3) {
                    const int a = mode == 0 ? t : 2 - t;
                    v[q] = __ldg(k0 + ((long)o * Cin + c) * 3 + a);
                } else {
                    int a = (t - 3) & 1, side = (t - 3) >> 1;
                    if (mode == 1) { a = 1 - a; side = 1 - side; }
                    v[q] = __ldg(k1 + (((long)o * Cin + c) * 2 + a) * 2 + side);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 7; ++q) s_w[e0 + q * CF_THREADS] = v[q];
    }
}

// What a stage's input transform needs, per channel, in shared memory: sc, sh (forward) and mean, invstd, mg, mgx (backward).
struct CfBn {
    float sc[CF_C], sh[CF_C], mean[CF_C], invstd[CF_C], mg[CF_C], mgx[CF_C];
};

// 7-tap hex convolution of one unit from staged rows: acc[8] for output channels 8*og .. 8*og+7 at column x.
__device__ __forceinline__ void cf_conv_unit(const float* __restrict__ s_in, const float* __restrict__ s_w, int Cin, int x, int og, int p, float acc[8]) {
    const int dxs[CF_T] = {-1, 0, 1, -1 + p, p, -1 + p, p};
    const int rws[CF_T] = {1, 1, 1, 0, 0, 2, 2};
#pragma unroll
    for (int t = 0; t < CF_T; ++t) {
        const float* src = s_in + rws[t] * CF_SW + x + 1 + dxs[t];
        const float* w = s_w + t * CF_C * CF_C + 8 * og;
#pragma unroll 4
        for (int c = 0; c < Cin; ++c) {
            const float v = src[c * 3 * CF_SW];
            const float4 w0 = *reinterpret_cast<const float4*>(w + c * CF_C);
            const float4 w1 = *reinterpret_cast<const float4*>(w + c * CF_C + 4);
            acc[0] = fmaf(v, w0.x, acc[0]); acc[1] = fmaf(v, w0.y, acc[1]); acc[2] = fmaf(v, w0.z, acc[2]); acc[3] = fmaf(v, w0.w, acc[3]);
            acc[4] = fmaf(v, w1.x, acc[4]); acc[5] = fmaf(v, w1.y, acc[5]); acc[6] = fmaf(v, w1.z, acc[6]); acc[7] = fmaf(v, w1.w, acc[7]);
        }
    }
}

// ================================================================================================ forward
__global__ void __launch_bounds__(CF_THREADS, 1) corrector_fused_fwd_kernel(const CorrFusedArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;                                  // [7][32][32]
    float* s_in = s_w + CF_T * CF_C * CF_C;             // [32][3][66]
    __shared__ CfBn bn;
    __shared__ double s_stat[2 * CF_C];
    const int tid = threadIdx.x, x = tid & 63, og = tid >> 6;
    const long HW = (long)a.H * a.W;
    const int n_units = a.B * a.H * a.n_xt;
    unsigned epoch = 0;

    for (int j = 0; j < a.L; ++j) {
        const int Cin = a.cin[j], Cout = a.cout[j];
        const float* in = j == 0 ? a.x : a.act[j - 1];
        float* out = a.act[j];
        // ---- input transform of this stage
        if (tid < CF_C) {
            float sc = 1.f, sh = 0.f, mean = 0.f, invstd = 1.f;
            if (a.pro[j] == 2 && tid < Cin) {
                const float g = a.gamma[j][tid], be = a.beta[j][tid];
                if (a.bn_training) {
                    const double cnt = (double)a.B * (double)HW;
                    const double m = __ldcg(a.stats + j * 64 + tid) / cnt;
                    double var = __ldcg(a.stats + j * 64 + 32 + tid) / cnt - m * m;
                    if (var < 0.0) var = 0.0;
                    mean = (float)m;
                    invstd = (float)(1.0 / sqrt(var + (double)a.eps[j]));
                    if (blockIdx.x == 0) {
                        const double unbiased = cnt > 1.0 ? var * cnt / (cnt - 1.0) : var;
                        a.rmean[j][tid] = (1.f - a.momentum[j]) * a.rmean[j][tid] + a.momentum[j] * (float)m;
                        a.rvar[j][tid] = (1.f - a.momentum[j]) * a.rvar[j][tid] + a.momentum[j] * (float)unbiased;
                    }
                } else {
                    mean = a.rmean[j][tid];
                    invstd = 1.f / sqrtf(a.rvar[j][tid] + a.eps[j]);
                }
                sc = g * invstd;
                sh = be - mean * sc;
            }
            bn.sc[tid] = sc; bn.sh[tid] = sh;
            if (blockIdx.x == 0 && a.pro[j] == 2 && tid < Cin) {
                float* k = a.consts + j * 128;
                k[tid] = sc; k[32 + tid] = sh; k[64 + tid] = mean; k[96 + tid] = invstd;
            }
        }
        const bool want_stats = j + 1 < a.L && a.pro[j + 1] == 2 && a.bn_training;
        if (tid < 2 * CF_C) s_stat[tid] = 0.0;
        if (j == 0) cf_stage_weights(s_w, a.k0[0], a.k1[0], Cin, Cout, 0);      // later stages: staged before the preceding barrier
        __syncthreads();
        const int pro = a.pro[j];
        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int xt = u % a.n_xt, y = (u / a.n_xt) % a.H, b = u / (a.n_xt * a.H);
            const int x0 = xt * CF_TW;
            const float* inb = in + (long)b * Cin * HW;
            const int n_in = Cin * 3 * CF_SW;
            for (int e0 = tid; e0 < n_in; e0 += 8 * CF_THREADS) {
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int e = e0 + q * CF_THREADS;
                    const int c = e / (3 * CF_SW), r = (e / CF_SW) % 3, xx = e % CF_SW;
                    const int gy = y + r - 1, gx = x0 + xx - 1;
                    v[q] = 0.f;
                    if (e < n_in && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) v[q] = __ldcg(inb + (long)c * HW + (long)gy * a.W + gx);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int e = e0 + q * CF_THREADS;
                    if (e < n_in) {
                        const int c = e / (3 * CF_SW), r = (e / CF_SW) % 3, xx = e % CF_SW;
                        const int gy = y + r - 1, gx = x0 + xx - 1;
                        const bool in_grid = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
                        s_in[e] = (pro && in_grid) ? fmaxf(fmaf(v[q], bn.sc[c], bn.sh[c]), 0.f) : v[q];
                    }
                }
            }
            __syncthreads();
            float acc[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] = 0.f;
            if (8 * og < Cout) cf_conv_unit(s_in, s_w, Cin, x, og, y & 1, acc);
            const bool ok = x0 + x < a.W;
            float* ob = out + (long)b * Cout * HW + (long)y * a.W + x0 + x;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int o = 8 * og + q;
                if (o < Cout) {               // warp-uniform
                    const float r = acc[q] + (a.bias[j] ? __ldg(a.bias[j] + o) : 0.f);
                    if (ok) ob[(long)o * HW] = r;
                    if (want_stats) {
                        const float s = gn_warp_sum(ok ? r : 0.f), qq = gn_warp_sum(ok ? r * r : 0.f);
                        if ((tid & 31) == 0) {
                            atomicAdd(&s_stat[o], (double)s);
                            atomicAdd(&s_stat[CF_C + o], (double)qq);
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (want_stats) {
            __syncthreads();
            if (tid < 2 * CF_C && (tid & 31) < Cout && s_stat[tid] != 0.0) atomicAdd(a.stats + (j + 1) * 64 + tid, s_stat[tid]);
        }
        if (j + 1 < a.L) {
            // the next stage's weights do not depend on the barrier: fetch them while the slower CTAs finish this phase
            __syncthreads();
            cf_stage_weights(s_w, a.k0[j + 1], a.k1[j + 1], a.cin[j + 1], a.cout[j + 1], 0);
            cf_grid_barrier(a.sync, epoch);
        }
    }
    cf_exit(a.sync);
}

// ================================================================================================ backward
// Gradient of the loss w.r.t. act[j] (= dY of stage j) from gb[j+1] (gradient w.r.t. stage j+1's transformed input) through that
// transform's backward: none | ReLU | BatchNorm(+ReLU) with the batch-statistics terms.
__device__ __forceinline__ float cf_dy(float dA, float h, int pro, int training, const CfBn& bn, int c) {
    if (pro == 0) return dA;
    if (pro == 1) return h > 0.f ? dA : 0.f;
    const float g = fmaf(h, bn.sc[c], bn.sh[c]) > 0.f ? dA : 0.f;
    if (!training) return bn.sc[c] * g;
    const float xhat = (h - bn.mean[c]) * bn.invstd[c];
    return bn.sc[c] * (g - bn.mg[c] - xhat * bn.mgx[c]);
}

__global__ void __launch_bounds__(CF_THREADS, 1) corrector_fused_bwd_kernel(const CorrFusedArgs a) {
    extern __shared__ __align__(16) float smem[];
    float* s_w = smem;                                  // [7][32][32]   data-gradient weights
    float* s_dy = s_w + CF_T * CF_C * CF_C;             // [32][3][66]   dY of this stage, rows y-1..y+1
    float* s_x = s_dy + CF_C * 3 * CF_SW;               // [3][66][32]   transformed input of this stage, channel-contiguous (swizzled)
    float* s_dyT = s_x + 3 * CF_SW * CF_C;              // [64][32]      dY of row y, channel-contiguous
    __shared__ CfBn bn_out;                              // BatchNorm between this stage and the next one (its backward makes dY)
    __shared__ CfBn bn_in;                               // BatchNorm in front of this stage (its forward makes X'; its sums are accumulated)
    __shared__ double s_sum[2 * CF_C];
    __shared__ float s_db[CF_C];
    const int tid = threadIdx.x, x = tid & 63, og = tid >> 6;
    const long HW = (long)a.H * a.W;
    const int n_units = a.B * a.H * a.n_xt;
    const double cnt = (double)a.B * (double)HW;
    unsigned epoch = 0;
    // weight-gradient tile of this thread
    const int wt = tid >> 5, wcg = (tid >> 2) & 7, wog = tid & 3;

    for (int j = a.L - 1; j >= 0; --j) {
        const int Cin = a.cin[j], Cout = a.cout[j];
        const int pro_out = j + 1 < a.L ? a.pro[j + 1] : 0;
        const int pro_in = a.pro[j];
        const float* dsrc = j + 1 < a.L ? a.gb[j + 1] : a.dout;        // (B, Cout, H, W)
        const float* hout = a.act[j];                                   // pre-transform values the next stage saw
        const float* xin = j == 0 ? a.x : a.act[j - 1];                 // (B, Cin, H, W), pre-transform
        float* gout = a.gb[j];                                          // (B, Cin, H, W) or null (stage 0 when dx is not needed)
        if (tid < CF_C) {
            // transform between this stage and the next: constants from the forward pass + the finished backward sums
            float sc = 1.f, sh = 0.f, mean = 0.f, invstd = 1.f, mg = 0.f, mgx = 0.f;
            if (pro_out == 2 && tid < Cout) {
                const float* k = a.consts + (j + 1) * 128;
                sc = k[tid]; sh = k[32 + tid]; mean = k[64 + tid]; invstd = k[96 + tid];
                if (a.bn_training) {
                    const double sg = __ldcg(a.sums + (j + 1) * 64 + tid), sgx = __ldcg(a.sums + (j + 1) * 64 + 32 + tid);
                    mg = (float)(sg / cnt);
                    mgx = (float)(sgx / cnt);
                    if (blockIdx.x == 0) {
                        a.dbeta[j + 1][tid] = (float)sg;
                        a.dgamma[j + 1][tid] = (float)sgx;
                    }
                } else if (blockIdx.x == 0) {
                    a.dbeta[j + 1][tid] = (float)__ldcg(a.sums + (j + 1) * 64 + tid);
                    a.dgamma[j + 1][tid] = (float)__ldcg(a.sums + (j + 1) * 64 + 32 + tid);
                }
            }
            bn_out.sc[tid] = sc; bn_out.sh[tid] = sh; bn_out.mean[tid] = mean; bn_out.invstd[tid] = invstd; bn_out.mg[tid] = mg; bn_out.mgx[tid] = mgx;
            sc = 1.f; sh = 0.f; mean = 0.f; invstd = 1.f;
            if (pro_in == 2 && tid < Cin) {
                const float* k = a.consts + j * 128;
                sc = k[tid]; sh = k[32 + tid]; mean = k[64 + tid]; invstd = k[96 + tid];
            }
            bn_in.sc[tid] = sc; bn_in.sh[tid] = sh; bn_in.mean[tid] = mean; bn_in.invstd[tid] = invstd;
            s_db[tid] = 0.f;
        }
        if (tid < 2 * CF_C) s_sum[tid] = 0.0;
        if (j == a.L - 1) cf_stage_weights(s_w, a.k0[j], a.k1[j], Cin, Cout, 1);      // earlier stages: staged before the preceding barrier
        float wacc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int q = 0; q < 8; ++q) wacc[i][q] = 0.f;
        float dbacc = 0.f;
        __syncthreads();
        const bool want_sums = pro_in == 2;                 // sum g, sum g*xhat of the BatchNorm in front of this stage (also the eval-mode parameter gradients)
        const bool need_dgrad = gout != nullptr;

        for (int u = blockIdx.x; u < n_units; u += gridDim.x) {
            const int xt = u % a.n_xt, y = (u / a.n_xt) % a.H, b = u / (a.n_xt * a.H);
            const int x0 = xt * CF_TW;
            // ---- dY rows y-1..y+1 (zero outside the grid), [o][row][x]
            const float* db_ = dsrc + (long)b * Cout * HW;
            const float* hb_ = hout + (long)b * Cout * HW;
            const int n_dy = Cout * 3 * CF_SW;
            for (int e0 = tid; e0 < n_dy; e0 += 4 * CF_THREADS) {
                float v[4], h[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int e = e0 + q * CF_THREADS;
                    const int o = e / (3 * CF_SW), r = (e / CF_SW) % 3, xx = e % CF_SW;
                    const int gy = y + r - 1, gx = x0 + xx - 1;
                    v[q] = 0.f; h[q] = 0.f;
                    if (e < n_dy && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) {
                        const long idx = (long)o * HW + (long)gy * a.W + gx;
                        v[q] = __ldcg(db_ + idx);
                        if (pro_out) h[q] = __ldcg(hb_ + idx);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int e = e0 + q * CF_THREADS;
                    if (e < n_dy) {
                        const int o = e / (3 * CF_SW), r = (e / CF_SW) % 3, xx = e % CF_SW;
                        const int gy = y + r - 1, gx = x0 + xx - 1;
                        const bool in_grid = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
                        s_dy[e] = (pro_out && in_grid) ? cf_dy(v[q], h[q], pro_out, a.bn_training, bn_out, o) : v[q];
                    }
                }
            }
            // ---- transformed input rows y-1..y+1, channel-contiguous: chunk (c >> 2) ^ (xx & 7), element c & 3
            const float* xb_ = xin + (long)b * Cin * HW;
            const int n_x = Cin * 3 * CF_SW;
            for (int e0 = tid; e0 < n_x; e0 += 8 * CF_THREADS) {
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int e = e0 + q * CF_THREADS;
                    const int c = e / (3 * CF_SW), r = (e / CF_SW) % 3, xx = e % CF_SW;
                    const int gy = y + r - 1, gx = x0 + xx - 1;
                    v[q] = 0.f;
                    if (e < n_x && gy >= 0 && gy < a.H && gx >= 0 && gx < a.W) v[q] = __ldcg(xb_ + (long)c * HW + (long)gy * a.W + gx);
                }
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int e = e0 + q * CF_THREADS;
                    if (e < n_x) {
                        const int c = e / (3 * CF_SW), r = (e / CF_SW) % 3, xx = e % CF_SW;
                        const int gy = y + r - 1, gx = x0 + xx - 1;
                        const bool in_grid = gy >= 0 && gy < a.H && gx >= 0 && gx < a.W;
                        s_x[(r * CF_SW + xx) * CF_C + ((((c >> 2) ^ (xx & 7)) << 2) | (c & 3))] =
                            (pro_in && in_grid) ? fmaxf(fmaf(v[q], bn_in.sc[c], bn_in.sh[c]), 0.f) : v[q];
                    }
                }
            }
            __syncthreads();
            // ---- dY of row y, channel-contiguous (zero padded to 32 channels)
            for (int e = tid; e < CF_TW * CF_C; e += CF_THREADS) {
                const int xx = e >> 5, o = e & 31;
                s_dyT[e] = o < Cout ? s_dy[(o * 3 + 1) * CF_SW + xx + 1] : 0.f;
            }
            // ---- data gradient: the same 7-tap convolution with reflected, transposed weights ("input" channels = Cout)
            if (need_dgrad) {
                float acc[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) acc[q] = 0.f;
                if (8 * og < Cin) cf_conv_unit(s_dy, s_w, Cout, x, og, y & 1, acc);
                const bool ok = x0 + x < a.W;
                const long cell = (long)y * a.W + x0 + x;
                float* gb_ = gout + (long)b * Cin * HW + cell;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int c = 8 * og + q;
                    if (c < Cin) {            // warp-uniform
                        if (ok) gb_[(long)c * HW] = acc[q];
                        if (want_sums) {
                            float g = 0.f, gx_ = 0.f;
                            if (ok) {
                                const float h = __ldcg(xb_ + (long)c * HW + cell);
                                g = fmaf(h, bn_in.sc[c], bn_in.sh[c]) > 0.f ? acc[q] : 0.f;
                                gx_ = g * ((h - bn_in.mean[c]) * bn_in.invstd[c]);
                            }
                            g = gn_warp_sum(g);
                            gx_ = gn_warp_sum(gx_);
                            if ((tid & 31) == 0) {
                                atomicAdd(&s_sum[c], (double)g);
                                atomicAdd(&s_sum[CF_C + c], (double)gx_);
                            }
                        }
                    }
                }
            }
            __syncthreads();
            // ---- weight gradient: thread (tap wt, input channels 4*wcg.., output channels 8*wog..) over the 64 columns of row y
            if (wt < CF_T && 4 * wcg < Cin && 8 * wog < Cout) {
                const int p = y & 1;
                const int dx = wt < 3 ? wt - 1 : ((wt - 3) & 1) - 1 + p;
                const int rw = wt < 3 ? 1 : (wt < 5 ? 0 : 2);
                const float* xrow = s_x + (rw * CF_SW + 1 + dx) * CF_C;
#pragma unroll 4
                for (int xx = 0; xx < CF_TW; ++xx) {
                    const int col = xx + 1 + dx;               // staged column index of the input cell
                    const float4 xv = *reinterpret_cast<const float4*>(xrow + xx * CF_C + ((wcg ^ (col & 7)) << 2));
                    const float4 d0 = *reinterpret_cast<const float4*>(s_dyT + xx * CF_C + 8 * wog);
                    const float4 d1 = *reinterpret_cast<const float4*>(s_dyT + xx * CF_C + 8 * wog + 4);
                    const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        wacc[i][0] = fmaf(xs[i], d0.x, wacc[i][0]); wacc[i][1] = fmaf(xs[i], d0.y, wacc[i][1]);
                        wacc[i][2] = fmaf(xs[i], d0.z, wacc[i][2]); wacc[i][3] = fmaf(xs[i], d0.w, wacc[i][3]);
                        wacc[i][4] = fmaf(xs[i], d1.x, wacc[i][4]); wacc[i][5] = fmaf(xs[i], d1.y, wacc[i][5]);
                        wacc[i][6] = fmaf(xs[i], d1.z, wacc[i][6]); wacc[i][7] = fmaf(xs[i], d1.w, wacc[i][7]);
                    }
                }
            }
            // ---- bias gradient: column sums of dY over row y
            if (a.dbias[j] != nullptr && tid < Cout) {
                float s = 0.f;
                for (int xx = 0; xx < CF_TW; ++xx) s += s_dyT[xx * CF_C + tid];
                dbacc += s;
            }
            __syncthreads();
        }
        // ---- flush this CTA's partial sums
        if (wt < CF_T && 4 * wcg < Cin && 8 * wog < Cout) {
            float* dst = a.dwp + (long)j * CF_T * CF_C * CF_C + ((long)wt * CF_C + 4 * wcg) * CF_C + 8 * wog;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                atomicAdd(reinterpret_cast<float4*>(dst + i * CF_C), make_float4(wacc[i][0], wacc[i][1], wacc[i][2], wacc[i][3]));
                atomicAdd(reinterpret_cast<float4*>(dst + i * CF_C + 4), make_float4(wacc[i][4], wacc[i][5], wacc[i][6], wacc[i][7]));
            }
        }
        if (a.dbias[j] != nullptr && tid < Cout && dbacc != 0.f) atomicAdd(a.dbias_acc + j * CF_C + tid, dbacc);
        if (want_sums && tid < 2 * CF_C && (tid & 31) < Cin && s_sum[tid] != 0.0) atomicAdd(a.sums + j * 64 + tid, s_sum[tid]);
        if (j > 0) {
            __syncthreads();
            cf_stage_weights(s_w, a.k0[j - 1], a.k1[j - 1], a.cin[j - 1], a.cout[j - 1], 1);
        }
        cf_grid_barrier(a.sync, epoch);
    }
    // ---- last phase: packed weight gradients -> parameter layout; bias gradients; a BatchNorm in front of stage 0 (not a
    // pattern the models build, but complete): its parameter gradients
    const long gtid = (long)blockIdx.x * CF_THREADS + tid, gstride = (long)gridDim.x * CF_THREADS;
    for (int j = 0; j < a.L; ++j) {
        const int Cin = a.cin[j], Cout = a.cout[j];
        const float* src = a.dwp + (long)j * CF_T * CF_C * CF_C;
        for (long e = gtid; e < (long)Cout * Cin * CF_T; e += gstride) {
            const int t = (int)(e % CF_T), c = (int)((e / CF_T) % Cin), o = (int)(e / ((long)CF_T * Cin));
            const float v = __ldcg(src + ((long)t * CF_C + c) * CF_C + o);
            if (t < 3) a.dk0[j][((long)o * Cin + c) * 3 + t] = v;
            else a.dk1[j][(((long)o * Cin + c) * 2 + ((t - 3) & 1)) * 2 + ((t - 3) >> 1)] = v;
        }
        if (a.dbias[j] != nullptr)
            for (long e = gtid; e < Cout; e += gstride) a.dbias[j][e] = __ldcg(a.dbias_acc + j * CF_C + e);
    }
    if (a.pro[0] == 2 && blockIdx.x == 0 && tid < a.cin[0]) {
        a.dbeta[0][tid] = (float)__ldcg(a.sums + tid);
        a.dgamma[0][tid] = (float)__ldcg(a.sums + 32 + tid);
    }
    cf_exit(a.sync);
}

// ================================================================================================ host
#define CF_FWD_SMEM ((CF_T * CF_C * CF_C + CF_C * 3 * CF_SW) * 4)
#define CF_BWD_SMEM ((CF_T * CF_C * CF_C + CF_C * 3 * CF_SW + 3 * CF_SW * CF_C + CF_TW * CF_C) * 4)

static int cf_fill(CorrFusedArgs& a, int L, int B, int H, int W, const int* cin, const int* cout, const int* pro, const void* const* params,
                   const float* momentum, const float* eps, int bn_training, const float* x, float* const* act, double* stats, float* consts,
                   unsigned* sync) {
    GN_REQUIRE(L >= 1 && L <= CF_MAX_STAGES && B > 0 && H > 0 && W > 0, GN_EINVAL, "corrector_fused: bad dimensions");
    GN_REQUIRE(x && act && params && cin && cout && pro && stats && consts && sync, GN_EINVAL, "corrector_fused: null argument");
    memset(&a, 0, sizeof(a));
    a.B = B; a.H = H; a.W = W; a.L = L; a.n_xt = (W + CF_TW - 1) / CF_TW;
    a.bn_training = bn_training;
    for (int j = 0; j < L; ++j) {
        GN_REQUIRE(cin[j] >= 1 && cin[j] <= CF_C && cout[j] >= 1 && cout[j] <= CF_C, GN_EUNSUPPORTED, "corrector_fused: stage %d has %d -> %d channels (max %d)", j,
                   cin[j], cout[j], CF_C);
        GN_REQUIRE(j == 0 || cin[j] == cout[j - 1], GN_EINVAL, "corrector_fused: stage %d input channels do not match stage %d output", j, j - 1);
        a.cin[j] = cin[j]; a.cout[j] = cout[j]; a.pro[j] = pro[j];
        const void* const* p = params + 7 * j;
        a.k0[j] = (const float*)p[0]; a.k1[j] = (const float*)p[1]; a.bias[j] = (const float*)p[2];
        a.gamma[j] = (const float*)p[3]; a.beta[j] = (const float*)p[4]; a.rmean[j] = (float*)p[5]; a.rvar[j] = (float*)p[6];
        GN_REQUIRE(a.k0[j] && a.k1[j] && act[j], GN_EINVAL, "corrector_fused: stage %d kernels / output missing", j);
        GN_REQUIRE(pro[j] != 2 || (a.gamma[j] && a.beta[j] && a.rmean[j] && a.rvar[j]), GN_EINVAL, "corrector_fused: stage %d BatchNorm parameters missing", j);
        a.momentum[j] = momentum ? momentum[j] : 0.1f;
        a.eps[j] = eps ? eps[j] : 1e-5f;
        a.act[j] = act[j];
    }
    a.x = x; a.stats = stats; a.consts = consts; a.sync = sync;
    return GN_OK;
}

template <typename K>
static int cf_launch(K kernel, CorrFusedArgs& a, size_t smem, cudaStream_t stream) {
    static int occ_cache[2] = {0, 0};
    const int which = smem == (size_t)CF_FWD_SMEM ? 0 : 1;
    if (occ_cache[which] == 0) {
        GN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int occ = 0;
        GN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, CF_THREADS, smem));
        GN_REQUIRE(occ >= 1, GN_EUNSUPPORTED, "corrector_fused: kernel does not fit an SM");
        occ_cache[which] = 1;               // one CTA per SM: the phases are short, more CTAs only lengthen the barrier
    }
    const int units = a.B * a.H * a.n_xt;
    int grid = gn_num_sms();
    if (grid > units) grid = units;
    void* args[] = {&a};
    GN_CUDA(cudaLaunchCooperativeKernel((const void*)kernel, dim3(grid), dim3(CF_THREADS), args, smem, stream));
    return GN_OK;
}

GN_API int gn_corrector_fused_supported(int L, const int* cin, const int* cout) {
    if (L < 1 || L > CF_MAX_STAGES) return 0;
    for (int j = 0; j < L; ++j)
        if (cin[j] < 1 || cin[j] > CF_C || cout[j] < 1 || cout[j] > CF_C) return 0;
    return 1;
}

// stats [L][64] fp64 and sync [2] u32: stats zeroed by the caller before every forward; sync zeroed once (the kernel resets it).
GN_API int gn_corrector_fused_fwd(const float* x, float* const* act, const void* const* params, const int* cin, const int* cout, const int* pro,
                                  const float* momentum, const float* eps, int L, int B, int H, int W, int bn_training, double* stats,
                                  float* consts, unsigned* sync, cudaStream_t stream) {
    CorrFusedArgs a;
    int rc = cf_fill(a, L, B, H, W, cin, cout, pro, params, momentum, eps, bn_training, x, act, stats, consts, sync);
    if (rc) return rc;
    return cf_launch(corrector_fused_fwd_kernel, a, CF_FWD_SMEM, stream);
}

// gb[j] (B, cin[j], H, W): gradient w.r.t. stage j's transformed input (gb[0] = dx; null when not needed).  grads: per stage
// dk0, dk1, dbias, dgamma, dbeta (dgamma/dbeta of the BatchNorm IN FRONT of the stage).  dwp [L][7][32][32], dbias_acc [L][32],
// sums [L][64] fp64: zeroed by the caller.
GN_API int gn_corrector_fused_bwd(const float* x, float* const* act, const float* dout, float* const* gb, const void* const* params,
                                  void* const* grads, const int* cin, const int* cout, const int* pro, const float* eps, int L, int B, int H,
                                  int W, int bn_training, float* dwp, float* dbias_acc, double* sums, double* stats, float* consts, unsigned* sync,
                                  cudaStream_t stream) {
    CorrFusedArgs a;
    int rc = cf_fill(a, L, B, H, W, cin, cout, pro, params, nullptr, eps, bn_training, x, act, stats, consts, sync);
    if (rc) return rc;
    GN_REQUIRE(dout && gb && grads && dwp && dbias_acc && sums, GN_EINVAL, "corrector_fused_bwd: null argument");
    a.dout = dout; a.dwp = dwp; a.dbias_acc = dbias_acc; a.sums = sums;
    for (int j = 0; j < L; ++j) {
        a.gb[j] = gb[j];
        GN_REQUIRE(j == 0 || gb[j], GN_EINVAL, "corrector_fused_bwd: gradient buffer of stage %d missing", j);
        void* const* g = grads + 5 * j;
        a.dk0[j] = (float*)g[0]; a.dk1[j] = (float*)g[1]; a.dbias[j] = (float*)g[2]; a.dgamma[j] = (float*)g[3]; a.dbeta[j] = (float*)g[4];
        GN_REQUIRE(a.dk0[j] && a.dk1[j], GN_EINVAL, "corrector_fused_bwd: weight gradient outputs of stage %d missing", j);
        GN_REQUIRE(pro[j] != 2 || (a.dgamma[j] && a.dbeta[j]), GN_EINVAL, "corrector_fused_bwd: BatchNorm gradient outputs of stage %d missing", j);
        GN_REQUIRE((a.bias[j] != nullptr) == (a.dbias[j] != nullptr), GN_EINVAL, "corrector_fused_bwd: bias / bias gradient of stage %d disagree", j);
    }
    return cf_launch(corrector_fused_bwd_kernel, a, CF_BWD_SMEM, stream);
}
