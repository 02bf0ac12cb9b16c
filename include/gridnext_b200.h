/* gridnext_b200 -- C-ABI of the B200 (sm_100a) GridNet hot path.
 *
 * The reference (adaly/gridnext) is pure Python/PyTorch and has no FFI; the functions below are the
 * entry points a maintainer binds (ctypes stub in INTEGRATION.md) to replace the ATen/cuDNN call
 * sequences named beside each one.  Conventions:
 *   - plain C, raw DEVICE pointers, sizes, and the cudaStream_t to enqueue on (void* here so the
 *     header needs no CUDA include); no torch types;
 *   - the CALLER owns every buffer, workspaces included; the library never allocates or frees
 *     device memory and keeps no pointer after returning;
 *   - all work is stream-ordered and asynchronous; functions are re-entrant;
 *   - return value: 0 ok, <0 argument/shape/alignment error, >0 a cudaError_t.
 *     gn_last_error() returns a thread-local description of the last failure.
 *   - tensors are contiguous; activations of the g network are fp32 NCHW in the Visium odd-r
 *     layout (B, C, H=78, W=64), labels int64 (B, H, W).
 */
#ifndef GRIDNEXT_B200_H
#define GRIDNEXT_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef void* gn_stream_t; /* cudaStream_t */

int gn_version(void);
const char* gn_last_error(void);
int gn_device_sm_count(void);
/* Runtime switch without a counterpart in the reference (which has no native code): programmatic dependent launch of the hot-path
 * kernels (csrc/gn_common.cuh: each kernel's prologue overlaps the tail of its predecessor)
 * on (default; GN_NO_PDL=1 in the environment forces off) or off for every later launch of this process; returns the previous setting.
 * The host side switches it off in data-parallel runs (NCCL kernels between kernels that trigger their dependents early: 2-GPU hang). */
int gn_set_pdl(int on);

/* ---- hexagonal convolution: replaces hexagdly.Conv2d.forward/backward (stride 1) as built by
 * gridnext/gridnet_models.py:128-148 plus the rot90/flip pair of gridnet_models.py:177-185.
 * kernel_i has shape (Cout, Cin, 2k+1-i, 1 if i==0 else 2); ksize k in 1..3. */
int gn_hexconv_n_taps(int ksize);
/* mode 0: wp[T][cin][cout] for the forward; mode 1: reflected+transposed weights so that the SAME
 * forward kernel computes dX from dY (call gn_hexconv_fwd with cin:=Cout, cout:=Cin, bias NULL). */
int gn_hexconv_pack(const float* k0, const float* k1, const float* k2, const float* k3, int ksize, int cin, int cout,
                    int mode, float* wp, gn_stream_t stream);
int gn_hexconv_unpack_grad(const float* dwp, float* dk0, float* dk1, float* dk2, float* dk3, int ksize, int cin, int cout,
                           gn_stream_t stream);
/* y = hexconv(x') + bias, x' = in_scale ? relu(x*in_scale[c]+in_shift[c]) : x  (BatchNorm2d-apply +
 * ReLU of gridnet_models.py:134-136 fused as prologue).  stats (nullable, fp64 [2*cout], caller
 * zeroes) accumulates per-channel sum and sum of squares of y for the next BatchNorm2d. */
int gn_hexconv_fwd(const float* x, const float* wp, const float* bias, const float* in_scale, const float* in_shift,
                   float* y, double* stats, int B, int cin, int cout, int H, int W, int ksize, gn_stream_t stream);
/* Tensor-core variant of gn_hexconv_fwd for kernel_size 1 and <= 32 channels (tcgen05, bf16 x 3 split, fp32 accumulate): same
 * contract; `workspace` is caller-owned scratch of gn_hexconv_tc_workspace_bytes(B, H, W) bytes, 1024-byte aligned. */
int gn_hexconv_tc_supported(int cin, int cout, int H, int W, int ksize);
long gn_hexconv_tc_workspace_bytes(int B, int H, int W);
int gn_hexconv_fwd_tc(const float* x, const float* wp, const float* bias, const float* in_scale, const float* in_shift, float* y,
                      double* stats, int B, int cin, int cout, int H, int W, void* workspace, gn_stream_t stream);
/* Second generation of the same operation (csrc/hexconv_tc2.cu): fp32 NCHW in and out with NO intermediate layout in global
 * memory -- rows are TMA-loaded once, converted to the bf16 hi | lo operand in shared memory (previous BatchNorm+ReLU fused there),
 * the taps of a grid row are stacked along UMMA N and the column shifts become neighbour-lane sums in the epilogue.  Grid width
 * <= 64 and a multiple of 4; same contract as gn_hexconv_fwd_tc; workspace = gn_hexconv_tc2_workspace_bytes(), 1024-byte aligned. */
int gn_hexconv_tc2_supported(int cin, int cout, int H, int W, int ksize);
long gn_hexconv_tc2_workspace_bytes(void);
int gn_hexconv_tc2_set_trace(long long* dev_buf);   /* development: 9 x 64 clock64() timestamps of CTA 0's pipeline events, or NULL */
int gn_hexconv_fwd_tc2(const float* x, const float* wp, const float* bias, const float* in_scale, const float* in_shift, float* y,
                       double* stats, int B, int cin, int cout, int H, int W, void* workspace, gn_stream_t stream);
/* Tensor-core weight gradient (same shapes as gn_hexconv_fwd_tc); dbias is not produced: it is gn_bn_stats' channel sum of dY. */
long gn_hexconv_tc_wgrad_workspace_bytes(int B, int H, int W);
int gn_hexconv_wgrad_tc(const float* x, const float* in_scale, const float* in_shift, const float* dy, float* dwp, int B, int cin,
                        int cout, int H, int W, void* workspace, gn_stream_t stream);
/* Second generation of the tensor-core weight gradient -- autograd's dW / db of the hexagdly.Conv2d layers built at
 * /root/reference/gridnext/gridnet_models.py:128-148, reached from training.py:141-171 (csrc/hexconv_wgrad_tc2.cu): x and dY are read once, straight from the fp32 NCHW
 * tensors, converted to bf16 hi | lo operand rows in shared memory; the cell is the reduction index, the neighbourhood's column shifts
 * are operand start addresses, taps stacked along UMMA N; no workspace.  Shapes as gn_hexconv_tc2_supported.  Also
 * produces dbias (nullable) on the way.  Same contract as gn_hexconv_wgrad: dwp / dbias are accumulated into, caller zeroes. */
int gn_hexconv_wgrad_tc2(const float* x, const float* in_scale, const float* in_shift, const float* dy, float* dwp, float* dbias, int B,
                         int cin, int cout, int H, int W, gn_stream_t stream);
/* dwp[T][cin][cout] += sum dY * x' ; dbias[cout] += sum dY  (caller zeroes both). */
int gn_hexconv_wgrad(const float* x, const float* in_scale, const float* in_shift, const float* dy, float* dwp,
                     float* dbias, int B, int cin, int cout, int H, int W, int ksize, gn_stream_t stream);

/* ---- BatchNorm2d (+ReLU): replaces nn.BatchNorm2d(32)/nn.ReLU of gridnet_models.py:134-136,142-144 */
int gn_bn_finalize(const double* stats, const float* gamma, const float* beta, float* running_mean, float* running_var,
                   float momentum, float eps, double count, float* scale, float* shift, float* mean_invstd, int C,
                   int update_running, gn_stream_t stream);
int gn_bn_eval_affine(const float* gamma, const float* beta, const float* running_mean, const float* running_var, float eps,
                      float* scale, float* shift, float* mean_invstd, int C, gn_stream_t stream);
int gn_bn_stats(const float* x, double* stats, int B, int C, long HW, gn_stream_t stream);
int gn_bn_act_fwd(const float* x, const float* scale, const float* shift, float* y, int B, int C, long HW, int relu,
                  gn_stream_t stream);
/* dH from dA through [ReLU](BN(h)); sums is a [2*C] fp64 workspace; dgamma/dbeta nullable. */
int gn_bn_act_bwd(const float* dA, const float* h, const float* scale, const float* shift, const float* mean_invstd,
                  double* sums, double count, int training, float* dH, float* dgamma, float* dbeta, int B, int C, long HW,
                  int relu, gn_stream_t stream);
/* The same in two calls (sum reduction, then apply) so that a data-parallel SyncBN can all-reduce `sums` in between
 * (the reference is single-device, gridnext/training.py:21,111). */
int gn_bn_act_bwd_reduce(const float* dA, const float* h, const float* scale, const float* shift, const float* mean_invstd,
                         double* sums, int B, int C, long HW, int relu, gn_stream_t stream);
int gn_bn_act_bwd_apply(const float* dA, const float* h, const float* scale, const float* shift, const float* mean_invstd,
                        const double* sums, double count, int training, float* dH, float* dgamma, float* dbeta, int B, int C,
                        long HW, int relu, gn_stream_t stream);

/* ---- the whole corrector nn.Sequential of gridnet_models.py:128-148 (hexagdly.Conv2d kernel_size 1, <= 32 channels, [BatchNorm2d]
 * + ReLU in front of any stage) as ONE persistent launch per direction (csrc/corrector_fused.cu): activations stay in L2, BatchNorm
 * statistics are reduced between phases by grid-wide barriers, weights are read from / gradients written to the parameter layout
 * kernel0 (Cout, Cin, 3, 1), kernel1 (Cout, Cin, 2, 2).  All pointer arrays are HOST arrays of device pointers:
 *   act[L]     outputs of every stage, (B, cout[j], H, W) fp32 (act[L-1] is the result; all are kept for the backward)
 *   params[7L] per stage: kernel0, kernel1, bias (nullable), then the BatchNorm IN FRONT of the stage: gamma, beta, running_mean,
 *              running_var (null when pro[j] != 2);  pro[j]: 0 none, 1 ReLU, 2 BatchNorm + ReLU applied to the stage's input
 *   stats [L][64] fp64 zeroed by the caller; consts [L][128] fp32 (written by fwd, read by bwd); sync: 2 x u32, zeroed ONCE.
 * bwd: gb[L] gradient w.r.t. every stage's (transformed) input, gb[0] = dx (nullable); grads[5L] per stage: d kernel0, d kernel1,
 * d bias, d gamma, d beta; dwp [L][7][32][32], dbias_acc [L][32], sums [L][64] fp64 zeroed by the caller. */
int gn_corrector_fused_supported(int L, const int* cin, const int* cout);
int gn_corrector_fused_fwd(const float* x, float* const* act, const void* const* params, const int* cin, const int* cout, const int* pro,
                           const float* momentum, const float* eps, int L, int B, int H, int W, int bn_training, double* stats,
                           float* consts, unsigned* sync, gn_stream_t stream);
int gn_corrector_fused_bwd(const float* x, float* const* act, const float* dout, float* const* gb, const void* const* params,
                           void* const* grads, const int* cin, const int* cout, const int* pro, const float* eps, int L, int B, int H,
                           int W, int bn_training, float* dwp, float* dbias_acc, double* sums, double* stats, float* consts, unsigned* sync,
                           gn_stream_t stream);

/* ---- foreground-masked cross-entropy: replaces gridnext/training.py:152-160 (+ its backward).
 * acc fp64[4]: {sum of spot losses, n_foreground, n_correct, n labels > C}; loss_out[0] = mean * loss_scale;
 * dlogits (nullable) = d(loss_out)/d(logits).  n_fg_override (nullable, device fp64[1]) replaces the
 * local foreground count as the normaliser (global count under data parallelism). */
int gn_masked_ce(const float* logits, const long long* labels, float* dlogits, double* acc, float* loss_out,
                 const double* n_fg_override, float loss_scale, int B, int C, long HW, gn_stream_t stream);

/* ---- evaluation loop: replaces the mask / labels - 1 / softmax / argmax of gridnext/utils.py:44-52 (all_fgd_predictions).
 * offsets[cell] = number of foreground cells (label > 0) before `cell` in (b, y, x) order; outputs are compacted in that order:
 * true_out[n_fg] = label - 1, pred_out[n_fg] = argmax over classes, smax_out[n_fg][C] = softmax. */
int gn_fg_predictions(const float* logits, const long long* labels, const int* offsets, long long* true_out, long long* pred_out,
                      float* smax_out, int B, int C, long HW, gn_stream_t stream);

/* ---- spot-patch gather: replaces gridnext/imgprocess.py:185-238 (grid_from_wsi_visium) */
int gn_spot_table(const unsigned char* in_tissue, const int* array_row, const int* array_col, const double* pxl_row,
                  const double* pxl_col, int n_spots, int h_st, int w_st, int* cells, int* n_dropped, gn_stream_t stream);
/* window_size != patch_size (gridnext/imgprocess.py:188-195,221: Image.fromarray(patch).resize((P, P)), Pillow BICUBIC):
 * the ws x ws window (ws even, edge-clamped) around every spot is resampled to P x P in Pillow's 8-bit fixed-point arithmetic
 * (22 fractional bits, horizontal then vertical pass, uint8 intermediate) -- bit-exact.  bounds [P][2] = (first tap, count),
 * kk [P][ksize] int32 coefficients, max_span = the largest tap count: the host builds them like Resample.c precompute_coeffs. */
int gn_patch_gather_resize(const unsigned char* img, long pitch, int H, int W, const int* cells, int n_cells, int ws, int P,
                           const int* bounds, const int* kk, int ksize, int max_span, const float* mean, const float* stdv,
                           void* out, int out_bf16, gn_stream_t stream);
/* ToTensor + Normalize of a pre-cropped uint8 patch grid (cells, 3, P, P); valid (nullable) marks present cells */
int gn_normalize_u8(const unsigned char* in, const unsigned char* valid, long n_cells, int P, const float* mean, const float* stdv,
                    void* out, int out_bf16, gn_stream_t stream);
int gn_patch_gather(const unsigned char* img, long pitch, int H, int W, const int* cells, int n_cells, int P,
                    const float* mean, const float* stdv, void* out, int out_bf16, gn_stream_t stream);

/* ---- tensor-core GEMM (tcgen05/TMEM/TMA): D[M,N] = op(A)[M,K] * B[N,K]^T, bf16 operands, fp32 accumulate.
 * Replaces F.conv2d 1x1 (gridnext/densenet.py:26-27,52-53) in NHWC and nn.Linear (count MLP, notebooks/
 * Tutorial_visium_count.ipynb cell 12).  a: [M, lda] bf16, b: [N, ldb] bf16, out: [M, ldc] bf16 | fp32.
 * Epilogue: out = [relu](acc * scale[n] + shift[n]) (scale/shift nullable), accumulate: out += (fp32 only).
 * xf_scale/xf_shift (nullable, [K]): op(A) = relu(A * xf_scale[k] + xf_shift[k]) applied in shared memory
 * (DenseNet's pre-activation eval-mode BatchNorm + ReLU, densenet.py:12-18). */
int gn_gemm_bf16(const void* a, long lda, const void* b, long ldb, int M, int N, int K, void* out, long ldc, int out_fp32,
                 int accumulate, const float* scale, const float* shift, int relu, const float* xf_scale, const float* xf_shift,
                 const void* bn_ref, long bn_ldref, int bn_ref_is_raw, const float* bn_sc, const float* bn_sh, const float* bn_p0,
                 const float* bn_p1, float* bn_colsum, int bn_ldsum, int bn_rmw, gn_stream_t stream);
/* bn_ref != NULL selects the BN+ReLU-backward epilogue of gn_conv3x3_bf16 (bf16 out; bn_rmw: out += g * bn_sc). */

/* count MLP glue (notebooks/Tutorial_visium_count.ipynb cell 12): fp32 count slab -> bf16 (consumed in place by gn_gemm_tn_bf16, so
 * the permute+reshape copy of gridnet_models.py:168,83 never happens); fp32 rows -> [relu](x*scale+shift) -> bf16 rows, pad columns zero */
int gn_cast_f32_bf16(const float* in, void* out, long n, gn_stream_t stream);
int gn_rows_affine_bf16(const float* in, long ldi, const float* scale, const float* shift, int relu, void* out, long ldo, long N, int C,
                        int Cpad, gn_stream_t stream);

/* Weight-gradient GEMM: out[Mo, No] (fp32, pitch ldo) += a[Kp, Mo]^T * op(b)[Kp, No]; rows of a and b are the
 * reduction index (pixels / spots).  op(b) = relu(b * xf_scale[n] + xf_shift[n]) when given.  Split over the
 * reduction dimension with vector atomics: the caller zeroes (or pre-loads) out.  Replaces the weight gradient
 * of F.conv2d 1x1 / nn.Linear that autograd computes for densenet.py:26-27,52-53 and the count MLP. */
int gn_gemm_tn_bf16(const void* a, long lda, const void* b, long ldb, int Mo, int No, int Kp, float* out, long ldo,
                    const float* xf_scale, const float* xf_shift, gn_stream_t stream);

/* Backward of the bottleneck 1x1 convolution of a dense layer in ONE kernel (gridnext/densenet.py:12-18,26-27: conv1(relu(norm1(cat)));
 * what autograd computes for those lines): dx[M, N] (+)= (dz[M,128] @ wt[N,128]^T) * [ref*sc + sh > 0] * sc  with the BatchNorm
 * column sums of gn_gemm_bf16's bn_* epilogue (ref = raw concat columns, p0 = mean, p1 = invstd), AND
 * dw[128, N] (fp32, pitch lddw) += dz^T @ relu(ref*sc + sh), the result of gn_gemm_tn_bf16 on the same operands, from the tiles
 * the data gradient already holds on chip (no second pass over dz and ref).  Bottleneck width is 128; all bf16 views must be
 * 16-byte aligned with pitches that are multiples of 8 elements. */
int gn_conv1x1_bwd_bf16(const void* dz, long lddz, const void* wt, long ldw, int M, int N, void* dx, long lddx, const void* ref, long ldref,
                        const float* sc, const float* sh, const float* p0, const float* p1, float* colsum, int ldsum, int rmw, float* dw,
                        long lddw, gn_stream_t stream);

/* ---- 3x3 / pad 1 convolution, NHWC bf16, tcgen05 implicit GEMM: replaces F.conv2d of densenet.py:30-31 (conv2 of
 * each dense layer) and its data gradient.  gn_conv3x3_pack: fp32 (CO, CI, 3, 3) weights -> bf16 [9*CO, ldw]
 * (mode 0, forward) or flipped/transposed [9*CI, ldw] (mode 1, data gradient; then call gn_conv3x3_bf16 with the
 * roles of CI and CO swapped).  Optional BN+ReLU-backward epilogue (bn_ref != NULL):
 *   g = acc * [a > 0], out = g * bn_sc[n], colsum[n] += sum g, colsum[bn_ldsum + n] += sum g * (ref - p0[n]) * p1[n]
 *   with a = ref * sc + sh (bn_ref_is_raw) or a = ref. */
int gn_conv3x3_pack(const float* w, int CO, int CI, int mode, void* wp, int ldw, gn_stream_t stream);
int gn_conv3x3_bf16(const void* x, long ldx, int Nimg, int H, int W, int CI, const void* wp, int ldw, int CO, void* out, long ldo,
                    const void* bn_ref, long bn_ldref, int bn_ref_is_raw, const float* bn_sc, const float* bn_sh,
                    const float* bn_p0, const float* bn_p1, float* bn_colsum, int bn_ldsum, gn_stream_t stream);

/* Weight gradient of the 3x3 convolution: dwp[9][CI][CO] (fp32) += sum over pixels x[p + tap] * dy[p];
 * gn_conv3x3_unpack_grad reorders to the (CO, CI, 3, 3) parameter layout. */
int gn_conv3x3_wgrad_bf16(const void* x, long ldx, const void* dy, long ldy, int Nimg, int H, int W, int CI, int CO, float* dwp,
                          gn_stream_t stream);
int gn_conv3x3_unpack_grad(const float* dwp, int CO, int CI, float* dw, gn_stream_t stream);

/* ---- memory-bound DenseNet pieces (NHWC bf16, eval-mode BatchNorm folded to scale/shift); each replaces the ATen
 * ops of the cited densenet.py lines.  colsum is fp32 [2][ldsum]: row 0 += sum g (d beta), row 1 += sum g*xhat (d gamma). */
/* Batched derivation of the bf16 / packed weight copies the tensor-core kernels consume (one launch for the whole network) and the
 * inverse for the packed gradient accumulators.  jobs: device array of n_jobs records {const float* src; long dst_off; long start;
 * int kind, a, b, ld;} (gn_prep_job_bytes() each), kinds: 0 cast [a,b]->[a,ld]; 1 transpose [a,b]->[b,ld]; 2/3 3x3 pack for the
 * forward / data gradient (what gn_conv3x3_pack does); 4 stem pack (gn_stem_pack_weight); 5 3x3 gradient un-pack; 6 stem un-pack. */
int gn_prep_job_bytes(void);
int gn_prepare_weights(const void* jobs, int n_jobs, long total_items, void* dst_bf16, gn_stream_t stream);
int gn_unpack_gradients(const void* jobs, int n_jobs, long total_items, float* dst, gn_stream_t stream);

/* Stem without an im2col buffer (densenet.py:107-109): the patch is repacked once to NHWC4 bf16 (RGB + zero channel), the
 * 7x7 / stride 2 convolution reads its overlapping operand rows straight from the image rows through a no-swizzle UMMA
 * descriptor; norm0 + relu0 run in the epilogue.  wq: [7][CO][32] bf16 (gn_stem_pack_weight), dwq: [CO][224] fp32. */
int gn_stem_pack_input(const void* x, int x_is_bf16, int N, int P, void* xq, gn_stream_t stream);
int gn_stem_pack_weight(const float* w, int CO, void* wq, gn_stream_t stream);
int gn_stem_conv_fwd(const void* xq, int N, int P, const void* wq, int CO, const float* scale, const float* shift, int relu, void* out,
                     long ldo, gn_stream_t stream);
int gn_stem_conv_wgrad(const void* xq, int N, int P, const void* dz, long ldz, int CO, float* dwq, gn_stream_t stream);
int gn_stem_unpack_wgrad(const float* dwq, int CO, float* dw, gn_stream_t stream);
/* pool0: MaxPool2d(3, 2, 1) (densenet.py:111-112); idx keeps the arg-max tap for the backward */
int gn_maxpool3s2_fwd(const void* in, long ldi, int N, int Hi, int Wi, int C, void* out, long ldo, unsigned char* idx, gn_stream_t stream);
int gn_maxpool3s2_bnrelu_bwd(const void* dpool, long ldp, const unsigned char* idx, const void* act, long lda, int N, int Hi, int Wi, int C,
                             const float* sc, const float* p0, const float* p1, void* dz, long ldz, float* colsum, int ldsum,
                             gn_stream_t stream);
/* _Transition (densenet.py:47-54): avg-pool commutes with the 1x1 conv, so BN+ReLU+AvgPool2d(2) runs first */
int gn_bnrelu_avgpool2_fwd(const void* in, long ldi, int N, int H, int W, int C, const float* sc, const float* sh, void* out, long ldo,
                           gn_stream_t stream);
int gn_pool_bnrelu_bwd(const void* dpool, long ldp, int gap, const void* raw, long ldr, int N, int H, int W, int C, const float* sc,
                       const float* sh, const float* p0, const float* p1, void* dC, long ldc, float* colsum, int ldsum, gn_stream_t stream);
/* head (densenet.py:134,153-158): norm_final + relu + adaptive_avg_pool2d(1) -> fp32 features; Linear classifier */
int gn_bnrelu_gap_fwd(const void* in, long ldi, int N, int HW, int C, const float* sc, const float* sh, float* feat, long ldf,
                      gn_stream_t stream);
int gn_linear_small_fwd(const float* feat, long ldf, const float* w, const float* b, int N, int C, int J, float* out, gn_stream_t stream);
int gn_linear_small_bwd(const float* dlog, const float* feat, long ldf, const float* w, int N, int C, int J, float* dfeat, long lddf,
                        float* dw, float* db, gn_stream_t stream);
int gn_bn_eval_consts(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int C, float* scale,
                      float* shift, float* invstd, float* inv_gamma, gn_stream_t stream);

/* Dataset tensor assembly (utils.py:144-166 read_annotated_starray, image_datasets.py:205-232 PatchGridDataset item,
 * multimodal_datasets.py:237-244): inverse spot->cell map (last spot wins, cell < 0 drops the spot), then one pass over the
 * OUTPUT grid: rows (patches), columns (genes x spots -> channels-first count slab), labels (+1, 0 = background), and the
 * multimodal foreground-consistency rule in place (flags: n_cells bytes of workspace). */
int gn_cell_inverse(const int* cell, int n_rows, int* inv, int n_cells, gn_stream_t stream);
int gn_grid_gather_rows(const void* src, long src_pitch_bytes, const int* inv, void* dst, long dst_pitch_bytes, long n_cells,
                        long row_bytes, gn_stream_t stream);
int gn_grid_gather_cols(const float* src, long src_pitch, const int* inv, float* dst, long dst_pitch, int G, int n_cells,
                        gn_stream_t stream);
int gn_grid_labels(const long long* labels, const int* inv, long long* annots, int n_cells, gn_stream_t stream);
int gn_mm_fg_consistency(float* patch, long F, float* counts, long counts_pitch, int G, long long* annots, unsigned char* flags,
                         int n_cells, gn_stream_t stream);

/* Cartesian K x K convolution (K in {1, 3, 5}), stride 1, zero padding K/2: nn.Conv2d of the base GridNet corrector
 * (gridnet_models.py:51-66: 3x3, 5x5, 5x5, 3x3).  Same tile kernels, packed layout Wp[r*K + c][cin][cout], BN/ReLU prologue and
 * BN-statistics epilogue as gn_hexconv_*; w / dw: (Cout, Cin, K, K) fp32.  pack mode 1 = reflected + transposed (data gradient). */
int gn_sqconv_pack(const float* w, int K, int cin, int cout, int mode, float* wp, gn_stream_t stream);
int gn_sqconv_unpack_grad(const float* dwp, float* dw, int K, int cin, int cout, gn_stream_t stream);
int gn_sqconv_fwd(const float* x, const float* wp, const float* bias, const float* in_scale, const float* in_shift, float* y,
                  double* stats, int B, int cin, int cout, int H, int W, int K, gn_stream_t stream);
int gn_sqconv_wgrad(const float* x, const float* in_scale, const float* in_shift, const float* dy, float* dwp, float* dbias,
                    int B, int cin, int cout, int H, int W, int K, gn_stream_t stream);

/* Train-mode BatchNorm around the same tensor-core kernels (f pre-training, training.py:11-98; the count f that training.py:126
 * leaves in train mode inside GridNetHexMM).  Forward: batch statistics of bf16 rows (fp64 sums; caller zeroes them) -> the
 * per-channel constants (scale, shift, mean, invstd) the kernels above consume, with nn.BatchNorm's running-stat update
 * (running_* nullable; running_var gets the unbiased variance); y = [relu](x*sc+sh) as a pass of its own, because a GEMM cannot
 * normalise its own output.  Backward: the BN-backward epilogues return dx_eval = g*scale and the column sums (d_beta, d_gamma);
 * dx = dx_eval - (c0 + c1*x) with c1 = scale*invstd*d_gamma/M, c0 = scale*d_beta/M - c1*mean (accumulate: several DenseNet
 * layers normalise the same concat channel).  C and pitches are multiples of 8, pointers 16-byte aligned. */
int gn_colstats_bf16(const void* x, long ld, long M, int C, double* sum, double* sumsq, gn_stream_t stream);
int gn_bn_train_coeffs(const double* sum, const double* sumsq, long M, const float* gamma, const float* beta, float eps, float momentum,
                       float* running_mean, float* running_var, float* sc, float* sh, float* mean, float* invstd, int C,
                       gn_stream_t stream);
int gn_affine_relu_bf16(const void* x, long ldx, void* y, long ldy, long M, int C, const float* sc, const float* sh, int relu,
                        gn_stream_t stream);
int gn_bn_train_fix_coeffs(const float* dbeta, const float* dgamma, const float* sc, const float* invstd, const float* mean, long M,
                           int accumulate, float* c0, float* c1, int C, gn_stream_t stream);
int gn_bn_train_fix_bf16(void* dx, long lddx, const void* x, long ldx, long M, int C, const float* c0, const float* c1, gn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif
