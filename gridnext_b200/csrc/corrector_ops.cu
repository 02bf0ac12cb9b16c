// BatchNorm2d(+ReLU) pieces and the foreground-masked cross-entropy of the g network (fp32, NCHW).
//
// Reference semantics:
//   nn.BatchNorm2d(32) + nn.ReLU inside the corrector   /root/reference/gridnext/gridnet_models.py:134-136,142-144
//     train: batch mean / biased variance over B*H*W cells, running stats with momentum 0.1 and the
//            unbiased variance, eps 1e-5;  eval: running-stat affine.
//   masked CE                                            /root/reference/gridnext/training.py:152-160
//     outputs[labels > 0], labels - 1, nn.CrossEntropyLoss (mean), torch.max(outputs, 1) for accuracy.
//
// The BN *apply* + ReLU is not a kernel of its own on the hot path: it is the prologue of the next
// hexconv (hexconv.cu); its statistics are the epilogue of the previous one.  What lives here is the
// tiny per-channel finalize, the backward reductions, and stand-alone apply kernels for eval/generic use.
#include "gn_common.cuh"

// ------------------------------------------------------------------------------------------------
__global__ void bn_finalize_kernel(const double* __restrict__ stats, const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum, float eps,
                                   double count, float* __restrict__ scale, float* __restrict__ shift,
                                   float* __restrict__ mean_invstd, int C, int update_running) {
    gn_pdl_sync();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    double mean = stats[c] / count;
    double var = stats[C + c] / count - mean * mean;
    if (var < 0.0) var = 0.0;
    float invstd = (float)(1.0 / sqrt(var + (double)eps));
    float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    float sc = g * invstd;
    scale[c] = sc;
    shift[c] = b - (float)mean * sc;
    mean_invstd[c] = (float)mean;
    mean_invstd[C + c] = invstd;
    if (update_running) {
        double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)mean;
        running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unbiased;
    }
}

__global__ void bn_eval_affine_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                                      const float* __restrict__ running_mean, const float* __restrict__ running_var, float eps,
                                      float* __restrict__ scale, float* __restrict__ shift, float* __restrict__ mean_invstd, int C) {
    gn_pdl_sync();
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float invstd = 1.f / sqrtf(running_var[c] + eps);
    float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
    float sc = g * invstd;
    scale[c] = sc;
    shift[c] = b - running_mean[c] * sc;
    if (mean_invstd) {
        mean_invstd[c] = running_mean[c];
        mean_invstd[C + c] = invstd;
    }
}

// per-channel sum / sumsq of an NCHW tensor (used when the producer is not one of our hexconvs)
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, double* __restrict__ stats, int B, int C, long HW) {
    gn_pdl_sync();
    const int c = blockIdx.x;
    double s = 0.0, q = 0.0;
    const long per = (long)B * HW;
    for (long e = blockIdx.y * (long)blockDim.x + threadIdx.x; e < per; e += (long)gridDim.y * blockDim.x) {
        long b = e / HW, p = e % HW;
        float v = __ldg(x + (b * C + c) * HW + p);
        s += v;
        q += (double)v * v;
    }
    __shared__ double sh[2][8];
    s = gn_warp_sum(s);
    q = gn_warp_sum(q);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = s; sh[1][threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0, tq = 0;
        for (int i = 0; i < 8; ++i) { ts += sh[0][i]; tq += sh[1][i]; }
        atomicAdd(stats + c, ts);
        atomicAdd(stats + C + c, tq);
    }
}

// y = [relu](x * scale[c] + shift[c])
__global__ void __launch_bounds__(256) bn_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ scale,
                                                         const float* __restrict__ shift, float* __restrict__ y, int C, long HW,
                                                         long total, int relu) {
    gn_pdl_sync();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
        int c = (int)((e / HW) % C);
        float v = fmaf(__ldg(x + e), __ldg(scale + c), __ldg(shift + c));
        y[e] = relu ? fmaxf(v, 0.f) : v;
    }
}

// sums[c] += sum g ; sums[C+c] += sum g * xhat,  g = dA * [relu: (h*scale+shift) > 0]
// Block (c, j) walks the planes b = j, j + gridDim.y, ... of channel c with 16-byte loads (no per-element index division; fp32 partial sums of
// the <= HW / 1024 vectors a thread sees per plane, fp64 across planes): at 256 arrays x 32 channels the two kernels of the BatchNorm
// backward took 0.44 ms against the 0.13 ms their 0.8 GB need (64-bit divisions and fp64 multiply-adds per element).
__global__ void __launch_bounds__(256) bn_act_bwd_reduce_kernel(const float* __restrict__ dA, const float* __restrict__ h,
                                                                const float* __restrict__ scale, const float* __restrict__ shift,
                                                                const float* __restrict__ mean_invstd, double* __restrict__ sums,
                                                                int B, int C, long HW, int relu, int vec) {
    gn_pdl_sync();
    const int c = blockIdx.x;
    const float sc = scale[c], sh_ = shift[c], mean = mean_invstd[c], invstd = mean_invstd[C + c];
    double s = 0.0, q = 0.0;
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const float* hp = h + ((long)b * C + c) * HW;
        const float* gp = dA + ((long)b * C + c) * HW;
        float fs = 0.f, fq = 0.f;
        if (vec) {
            const int n4 = (int)(HW >> 2);
            for (int i = threadIdx.x; i < n4; i += blockDim.x) {
                const float4 hv = __ldg(reinterpret_cast<const float4*>(hp) + i);
                float4 g = __ldg(reinterpret_cast<const float4*>(gp) + i);
                if (relu) {
                    if (!(fmaf(hv.x, sc, sh_) > 0.f)) g.x = 0.f;
                    if (!(fmaf(hv.y, sc, sh_) > 0.f)) g.y = 0.f;
                    if (!(fmaf(hv.z, sc, sh_) > 0.f)) g.z = 0.f;
                    if (!(fmaf(hv.w, sc, sh_) > 0.f)) g.w = 0.f;
                }
                fs += (g.x + g.y) + (g.z + g.w);
                fq = fmaf(g.x, (hv.x - mean) * invstd, fq);
                fq = fmaf(g.y, (hv.y - mean) * invstd, fq);
                fq = fmaf(g.z, (hv.z - mean) * invstd, fq);
                fq = fmaf(g.w, (hv.w - mean) * invstd, fq);
            }
        } else {
            for (long i = threadIdx.x; i < HW; i += blockDim.x) {
                const float hv = __ldg(hp + i);
                float g = __ldg(gp + i);
                if (relu && !(fmaf(hv, sc, sh_) > 0.f)) g = 0.f;
                fs += g;
                fq = fmaf(g, (hv - mean) * invstd, fq);
            }
        }
        s += (double)fs;
        q += (double)fq;
    }
    __shared__ double shm[2][8];
    s = gn_warp_sum(s);
    q = gn_warp_sum(q);
    if ((threadIdx.x & 31) == 0) { shm[0][threadIdx.x >> 5] = s; shm[1][threadIdx.x >> 5] = q; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ts = 0, tq = 0;
        for (int i = 0; i < 8; ++i) { ts += shm[0][i]; tq += shm[1][i]; }
        atomicAdd(sums + c, ts);
        atomicAdd(sums + C + c, tq);
    }
}

// training: dH = scale * (g - sum_g/n - xhat * sum_gx/n);  eval: dH = scale * g
// dgamma[c] = sum_gx, dbeta[c] = sum_g (written by block y == 0)
__global__ void __launch_bounds__(256) bn_act_bwd_apply_kernel(const float* __restrict__ dA, const float* __restrict__ h,
                                                               const float* __restrict__ scale, const float* __restrict__ shift,
                                                               const float* __restrict__ mean_invstd, const double* __restrict__ sums,
                                                               double count, int training, float* __restrict__ dH,
                                                               float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                               int B, int C, long HW, int relu, int vec) {
    gn_pdl_sync();
    const int c = blockIdx.x;
    const float sc = scale[c], sh_ = shift[c], mean = mean_invstd[c], invstd = mean_invstd[C + c];
    const float mg = training ? (float)(sums[c] / count) : 0.f;
    const float mgx = training ? (float)(sums[C + c] / count) : 0.f;
    if (blockIdx.y == 0 && threadIdx.x == 0) {
        if (dgamma) dgamma[c] = (float)sums[C + c];
        if (dbeta) dbeta[c] = (float)sums[c];
    }
    for (int b = blockIdx.y; b < B; b += gridDim.y) {
        const long off = ((long)b * C + c) * HW;
        const float* hp = h + off;
        const float* gp = dA + off;
        float* op = dH + off;
        if (vec) {
            const int n4 = (int)(HW >> 2);
            for (int i = threadIdx.x; i < n4; i += blockDim.x) {
                const float4 hv = __ldg(reinterpret_cast<const float4*>(hp) + i);
                float4 g = __ldg(reinterpret_cast<const float4*>(gp) + i);
                if (relu) {
                    if (!(fmaf(hv.x, sc, sh_) > 0.f)) g.x = 0.f;
                    if (!(fmaf(hv.y, sc, sh_) > 0.f)) g.y = 0.f;
                    if (!(fmaf(hv.z, sc, sh_) > 0.f)) g.z = 0.f;
                    if (!(fmaf(hv.w, sc, sh_) > 0.f)) g.w = 0.f;
                }
                float4 o;
                o.x = sc * (g.x - mg - (hv.x - mean) * invstd * mgx);
                o.y = sc * (g.y - mg - (hv.y - mean) * invstd * mgx);
                o.z = sc * (g.z - mg - (hv.z - mean) * invstd * mgx);
                o.w = sc * (g.w - mg - (hv.w - mean) * invstd * mgx);
                reinterpret_cast<float4*>(op)[i] = o;
            }
        } else {
            for (long i = threadIdx.x; i < HW; i += blockDim.x) {
                const float hv = __ldg(hp + i);
                float g = __ldg(gp + i);
                if (relu && !(fmaf(hv, sc, sh_) > 0.f)) g = 0.f;
                op[i] = sc * (g - mg - (hv - mean) * invstd * mgx);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// masked cross-entropy over (B, C, H, W) logits and (B, H, W) int64 labels (0 = background)
// acc[0] = sum of per-spot losses, acc[1] = n_foreground, acc[2] = n_correct, acc[3] = labels > C (out of range)   (fp64)
__global__ void __launch_bounds__(256) ce_count_kernel(const long long* __restrict__ labels, long n, double* __restrict__ acc) {
    gn_pdl_sync();
    int cnt = 0;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) cnt += labels[e] > 0;
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(acc + 1, (double)cnt);
}

#define CE_MAX_C 64
__global__ void __launch_bounds__(256) ce_main_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                      float* __restrict__ dlogits, double* __restrict__ acc,
                                                      const double* __restrict__ n_fg_ptr, float grad_scale, int C, long HW, long n) {
    gn_pdl_sync();
    const double nfg = n_fg_ptr[0];
    const float gs = nfg > 0 ? (float)(grad_scale / nfg) : 0.f;
    double loss = 0.0;
    int correct = 0, bad = 0;
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const long b = e / HW, p = e % HW;
        const float* lp = logits + b * C * HW + p;
        const long long lab = labels[e];
        if (lab > 0) {
            float v[CE_MAX_C];
            float m = -INFINITY;
            int am = 0;
#pragma unroll 8
            for (int c = 0; c < C; ++c) {
                v[c % CE_MAX_C] = __ldg(lp + (long)c * HW);
                if (v[c % CE_MAX_C] > m) { m = v[c % CE_MAX_C]; am = c; }
            }
            float se = 0.f;
            for (int c = 0; c < C; ++c) se += expf(v[c % CE_MAX_C] - m);
            const int cls = (int)lab - 1;
            const float lse = m + logf(se);
            if (cls < C) loss += (double)(lse - v[cls % CE_MAX_C]);
            else ++bad;                      // nn.CrossEntropyLoss raises "Target out of bounds"; the host checks acc[3]
            correct += (am == cls);
            if (dlogits) {
                const float inv = 1.f / se;
                float* dp = dlogits + b * C * HW + p;
                for (int c = 0; c < C; ++c) dp[(long)c * HW] = (expf(v[c % CE_MAX_C] - m) * inv - (c == cls ? 1.f : 0.f)) * gs;
            }
        } else if (dlogits) {
            float* dp = dlogits + b * C * HW + p;
            for (int c = 0; c < C; ++c) dp[(long)c * HW] = 0.f;
        }
    }
    loss = gn_warp_sum(loss);
    correct = __reduce_add_sync(0xffffffffu, correct);
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        if (loss != 0.0) atomicAdd(acc + 0, loss);
        if (correct) atomicAdd(acc + 2, (double)correct);
        if (bad) atomicAdd(acc + 3, (double)bad);
    }
}

__global__ void ce_finalize_kernel(const double* __restrict__ acc, const double* __restrict__ n_fg_ptr, float loss_scale,
                                   float* __restrict__ loss_out) {
    gn_pdl_sync();
    const double nfg = n_fg_ptr[0];
    loss_out[0] = nfg > 0 ? (float)(acc[0] / nfg) * loss_scale : NAN;   // CrossEntropyLoss over an empty set is NaN
}

// ------------------------------------------------------------------------------------------------
// Evaluation loop (utils.all_fgd_predictions, /root/reference/gridnext/utils.py:44-52): foreground mask, labels - 1, soft-max and
// arg-max over the classes, written compacted in grid order.  offsets[cell] = number of foreground cells before `cell`.
__global__ void __launch_bounds__(256) fg_predictions_kernel(const float* __restrict__ logits, const long long* __restrict__ labels,
                                                             const int* __restrict__ offsets, long long* __restrict__ true_out,
                                                             long long* __restrict__ pred_out, float* __restrict__ smax_out, int C, long HW,
                                                             long n) {
    gn_pdl_sync();
    for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < n; e += (long)gridDim.x * blockDim.x) {
        const long long lab = labels[e];
        if (lab <= 0) continue;
        const long b = e / HW, p = e % HW;
        const float* lp = logits + b * C * HW + p;
        float m = -INFINITY;
        int am = 0;
        for (int c = 0; c < C; ++c) {
            const float v = __ldg(lp + (long)c * HW);
            if (v > m) { m = v; am = c; }                  // first maximum, like torch.argmax
        }
        float se = 0.f;
        for (int c = 0; c < C; ++c) se += expf(__ldg(lp + (long)c * HW) - m);
        const float inv = 1.f / se;
        const long o = offsets[e];
        true_out[o] = lab - 1;
        pred_out[o] = am;
        float* sp = smax_out + o * C;
        for (int c = 0; c < C; ++c) sp[c] = expf(__ldg(lp + (long)c * HW) - m) * inv;
    }
}

GN_API int gn_fg_predictions(const float* logits, const long long* labels, const int* offsets, long long* true_out, long long* pred_out,
                             float* smax_out, int B, int C, long HW, cudaStream_t stream) {
    GN_REQUIRE(logits && labels && offsets && true_out && pred_out && smax_out && B > 0 && C > 0 && HW > 0, GN_EINVAL, "fg_predictions: bad arguments");
    const long n = (long)B * HW;
    long blocks = (n + 255) / 256;
    if (blocks > 8L * gn_num_sms()) blocks = 8L * gn_num_sms();
    GN_CUDA(gn_launch(fg_predictions_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, logits, labels, offsets, true_out, pred_out, smax_out, C, HW, n));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// ------------------------------------------------------------------------------------------------
// C-ABI
GN_API int gn_bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                          float momentum, float eps, double count, float* scale, float* shift, float* mean_invstd, int C,
                          int update_running, cudaStream_t stream) {
    GN_REQUIRE(stats && scale && shift && mean_invstd && C > 0 && count > 0, GN_EINVAL, "bn_finalize: bad arguments");
    GN_REQUIRE(!update_running || (running_mean && running_var), GN_EINVAL, "bn_finalize: running stats missing");
    GN_CUDA(gn_launch(bn_finalize_kernel, dim3(gn_ceil_div(C, 128)), dim3(128), 0, stream, stats, gamma, beta, running_mean, running_var, momentum, eps, count,
                                                                 scale, shift, mean_invstd, C, update_running));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

GN_API int gn_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps,
                             float* scale, float* shift, float* mean_invstd, int C, cudaStream_t stream) {
    GN_REQUIRE(running_mean && running_var && scale && shift && C > 0, GN_EINVAL, "bn_eval_affine: bad arguments");
    GN_CUDA(gn_launch(bn_eval_affine_kernel, dim3(gn_ceil_div(C, 128)), dim3(128), 0, stream, gamma, beta, running_mean, running_var, eps, scale, shift, mean_invstd, C));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// blocks per channel of the plane-wise BatchNorm-backward kernels: ~8 blocks per SM over all channels, at most one per array
static inline int plane_chunks_for(int B, int C) {
    long want = (8L * gn_num_sms() + C - 1) / C;
    if (want > B) want = B;
    return (int)(want < 1 ? 1 : want);
}
static inline int bn_vec_ok(const void* a, const void* b, const void* c, long HW) {
    return (HW & 3) == 0 && ((((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) & 15) == 0);
}

static inline int chunks_for(long per, int C) {
    long want = (4L * gn_num_sms() + C - 1) / C;
    long maxc = (per + 255) / 256;
    if (want > maxc) want = maxc;
    return (int)(want < 1 ? 1 : want);
}

GN_API int gn_bn_stats(const float* x, double* stats, int B, int C, long HW, cudaStream_t stream) {
    GN_REQUIRE(x && stats && B > 0 && C > 0 && HW > 0, GN_EINVAL, "bn_stats: bad arguments");
    dim3 grid(C, chunks_for((long)B * HW, C));
    GN_CUDA(gn_launch(bn_stats_kernel, dim3(grid), dim3(256), 0, stream, x, stats, B, C, HW));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

GN_API int gn_bn_act_fwd(const float* x, const float* scale, const float* shift, float* y, int B, int C, long HW, int relu,
                         cudaStream_t stream) {
    GN_REQUIRE(x && scale && shift && y && B > 0 && C > 0 && HW > 0, GN_EINVAL, "bn_act_fwd: bad arguments");
    long total = (long)B * C * HW;
    long blocks = (total + 255) / 256;
    if (blocks > 8L * gn_num_sms()) blocks = 8L * gn_num_sms();
    GN_CUDA(gn_launch(bn_act_fwd_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, x, scale, shift, y, C, HW, total, relu));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

GN_API int gn_bn_act_bwd(const float* dA, const float* h, const float* scale, const float* shift, const float* mean_invstd,
                         double* sums /* [2C] workspace, zeroed here */, double count, int training, float* dH, float* dgamma,
                         float* dbeta, int B, int C, long HW, int relu, cudaStream_t stream) {
    GN_REQUIRE(dA && h && scale && shift && mean_invstd && sums && dH && B > 0 && C > 0 && HW > 0, GN_EINVAL, "bn_act_bwd: bad arguments");
    GN_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)C * sizeof(double), stream));
    dim3 grid(C, plane_chunks_for(B, C));
    const int vec = bn_vec_ok(dA, h, dH, HW);
    GN_CUDA(gn_launch(bn_act_bwd_reduce_kernel, dim3(grid), dim3(256), 0, stream, dA, h, scale, shift, mean_invstd, sums, B, C, HW, relu, vec));
    GN_LAUNCH_CHECK();
    GN_CUDA(gn_launch(bn_act_bwd_apply_kernel, dim3(grid), dim3(256), 0, stream, dA, h, scale, shift, mean_invstd, sums, count, training, dH, dgamma, dbeta, B, C,
                                                      HW, relu, vec));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

// The two halves of gn_bn_act_bwd as separate calls, so that a data-parallel SyncBN can all-reduce `sums` in between.
GN_API int gn_bn_act_bwd_reduce(const float* dA, const float* h, const float* scale, const float* shift, const float* mean_invstd,
                                double* sums /* [2C], zeroed here */, int B, int C, long HW, int relu, cudaStream_t stream) {
    GN_REQUIRE(dA && h && scale && shift && mean_invstd && sums && B > 0 && C > 0 && HW > 0, GN_EINVAL, "bn_act_bwd_reduce: bad arguments");
    GN_CUDA(cudaMemsetAsync(sums, 0, 2 * (size_t)C * sizeof(double), stream));
    dim3 grid(C, plane_chunks_for(B, C));
    GN_CUDA(gn_launch(bn_act_bwd_reduce_kernel, dim3(grid), dim3(256), 0, stream, dA, h, scale, shift, mean_invstd, sums, B, C, HW, relu, bn_vec_ok(dA, h, nullptr, HW)));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

GN_API int gn_bn_act_bwd_apply(const float* dA, const float* h, const float* scale, const float* shift, const float* mean_invstd,
                               const double* sums, double count, int training, float* dH, float* dgamma, float* dbeta, int B, int C,
                               long HW, int relu, cudaStream_t stream) {
    GN_REQUIRE(dA && h && scale && shift && mean_invstd && sums && dH && B > 0 && C > 0 && HW > 0, GN_EINVAL, "bn_act_bwd_apply: bad arguments");
    dim3 grid(C, plane_chunks_for(B, C));
    GN_CUDA(gn_launch(bn_act_bwd_apply_kernel, dim3(grid), dim3(256), 0, stream, dA, h, scale, shift, mean_invstd, sums, count, training, dH, dgamma, dbeta, B, C,
                                                      HW, relu, bn_vec_ok(dA, h, dH, HW)));
    GN_LAUNCH_CHECK();
    return GN_OK;
}

GN_API int gn_masked_ce(const float* logits, const long long* labels, float* dlogits, double* acc /* [4] */, float* loss_out,
                        const double* n_fg_override, float loss_scale, int B, int C, long HW, cudaStream_t stream) {
    GN_REQUIRE(logits && labels && acc && loss_out && B > 0 && C > 0 && HW > 0, GN_EINVAL, "masked_ce: bad arguments");
    GN_REQUIRE(C <= CE_MAX_C, GN_EUNSUPPORTED, "masked_ce: n_classes %d > %d", C, CE_MAX_C);
    const long n = (long)B * HW;
    GN_CUDA(cudaMemsetAsync(acc, 0, 4 * sizeof(double), stream));
    long blocks = (n + 255) / 256;
    if (blocks > 4L * gn_num_sms()) blocks = 4L * gn_num_sms();
    GN_CUDA(gn_launch(ce_count_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, labels, n, acc));
    GN_LAUNCH_CHECK();
    const double* nfg = n_fg_override ? n_fg_override : acc + 1;
    GN_CUDA(gn_launch(ce_main_kernel, dim3((unsigned)blocks), dim3(256), 0, stream, logits, labels, dlogits, acc, nfg, loss_scale, C, HW, n));
    GN_LAUNCH_CHECK();
    GN_CUDA(gn_launch(ce_finalize_kernel, dim3(1), dim3(1), 0, stream, acc, nfg, loss_scale, loss_out));
    GN_LAUNCH_CHECK();
    return GN_OK;
}
