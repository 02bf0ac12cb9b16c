"""Python faces of the tensor-core (tcgen05) entry points.  Thin: argument checks + the ctypes call."""
import os
import torch
from . import _lib
from ._lib import ptr, stream, call


def _rows_pitch(t):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError('expected a 2-D tensor with unit column stride, got strides %s' % (t.stride(),))
    return t.shape[0], t.shape[1], t.stride(0)


def _bn_args(bn):
    if bn is None:
        return (None, 0, 0, None, None, None, None, None, 0)
    _, _, ldref = _rows_pitch(bn['ref'])
    cs = bn.get('colsum')
    return (ptr(bn['ref']), ldref, 1 if bn['ref_is_raw'] else 0, ptr(bn['sc']), ptr(bn.get('sh')), ptr(bn['p0']), ptr(bn['p1']),
            ptr(cs), cs.stride(0) if cs is not None else 0)


def gemm_bf16(a, b, out=None, out_dtype=torch.bfloat16, accumulate=False, scale=None, shift=None, relu=False,
              xf_scale=None, xf_shift=None, bn=None):
    """out[M, N] = [relu]((op(a) @ b.T) * scale + shift);  a [M, K], b [N, K] bf16 (row pitch may exceed K).

    ``out`` may be a column slice of a wider row-major buffer (e.g. a DenseNet concat buffer)."""
    _lib.require_cuda(a, b)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16:
        raise ValueError('gemm_bf16: operands must be bfloat16')
    M, K, lda = _rows_pitch(a)
    N, Kb, ldb = _rows_pitch(b)
    if K != Kb:
        raise ValueError('gemm_bf16: K mismatch %d vs %d' % (K, Kb))
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    Mo, No, ldc = _rows_pitch(out)
    if (Mo, No) != (M, N):
        raise ValueError('gemm_bf16: out shape %s != (%d, %d)' % (tuple(out.shape), M, N))
    if out.dtype not in (torch.bfloat16, torch.float32):
        raise ValueError('gemm_bf16: out must be bfloat16 or float32')
    call('gn_gemm_bf16', ptr(a), lda, ptr(b), ldb, M, N, K, ptr(out), ldc, 1 if out.dtype == torch.float32 else 0,
         1 if accumulate else 0, ptr(scale), ptr(shift), 1 if relu else 0, ptr(xf_scale), ptr(xf_shift),
         *_bn_args(bn), 1 if (bn is not None and bn.get('rmw')) else 0, stream())
    return out


def conv1x1_bwd_fusable(dz, wt, dx, ref):
    """True when gn_conv1x1_bwd_bf16 can take these views (bottleneck width 128, TMA-addressable operands)."""
    if os.environ.get('GRIDNEXT_B200_FUSED_BWD1X1', '1') == '0':
        return False
    if dz.shape[1] != 128 or wt.shape[1] != 128:
        return False
    for t in (dz, wt, dx, ref):
        if t.dtype != torch.bfloat16 or t.stride(1) != 1 or t.stride(0) % 8 != 0 or t.data_ptr() % 16 != 0:
            return False
    return True


def conv1x1_bwd_bf16(dz, wt, dx, bn, dw):
    """dx (+)= BN/ReLU-backward(dz @ wt.T) and dw += dz.T @ relu(bn(ref)) in one pass over dz and ref (see gn_conv1x1_bwd_bf16)."""
    _lib.require_cuda(dz, wt, dx, dw)
    M, K, lddz = _rows_pitch(dz)
    N, Kb, ldw = _rows_pitch(wt)
    Mo, No, lddx = _rows_pitch(dx)
    if K != 128 or Kb != 128 or (Mo, No) != (M, N) or tuple(dw.shape) != (128, N) or dw.dtype != torch.float32:
        raise ValueError('conv1x1_bwd_bf16: shape mismatch')
    if not bn['ref_is_raw']:
        raise ValueError('conv1x1_bwd_bf16: the reference must be the raw (pre-BatchNorm) tensor')
    _, _, ldref = _rows_pitch(bn['ref'])
    cs = bn.get('colsum')
    call('gn_conv1x1_bwd_bf16', ptr(dz), lddz, ptr(wt), ldw, M, N, ptr(dx), lddx, ptr(bn['ref']), ldref, ptr(bn['sc']), ptr(bn['sh']),
         ptr(bn['p0']), ptr(bn['p1']), ptr(cs), cs.stride(0) if cs is not None else 0, 1 if bn.get('rmw') else 0, ptr(dw), dw.stride(0), stream())
    return dx


def gemm_tn_bf16(a, b, out, xf_scale=None, xf_shift=None):
    """out[Mo, No] (fp32) += a[Kp, Mo].T @ op(b)[Kp, No]  (weight gradients; reduction over rows)."""
    _lib.require_cuda(a, b, out)
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16 or out.dtype != torch.float32:
        raise ValueError('gemm_tn_bf16: a, b must be bfloat16 and out float32')
    Kp, Mo, lda = _rows_pitch(a)
    Kb, No, ldb = _rows_pitch(b)
    if Kp != Kb:
        raise ValueError('gemm_tn_bf16: reduction length mismatch %d vs %d' % (Kp, Kb))
    Mo2, No2, ldo = _rows_pitch(out)
    if (Mo2, No2) != (Mo, No):
        raise ValueError('gemm_tn_bf16: out shape %s != (%d, %d)' % (tuple(out.shape), Mo, No))
    call('gn_gemm_tn_bf16', ptr(a), lda, ptr(b), ldb, Mo, No, Kp, ptr(out), ldo, ptr(xf_scale), ptr(xf_shift), stream())
    return out


def conv3x3_pack(w, mode, ldw=None):
    """fp32 (CO, CI, 3, 3) -> bf16 [9*CO, ldw] (mode 0) | flipped+transposed [9*CI, ldw] (mode 1, data gradient)."""
    _lib.require_cuda(w)
    CO, CI = int(w.shape[0]), int(w.shape[1])
    cols = CI if mode == 0 else CO
    ldw = ldw or ((cols + 7) // 8) * 8
    wp = torch.empty((9 * (CO if mode == 0 else CI), ldw), device=w.device, dtype=torch.bfloat16)
    call('gn_conv3x3_pack', ptr(w.contiguous().float()), CO, CI, mode, ptr(wp), ldw, stream())
    return wp


def conv3x3_bf16(x2d, Nimg, H, W, CI, wp, CO, out2d, bn=None):
    """x2d: [Nimg*H*W, >=CI] bf16 NHWC rows (pitch = stride(0)); out2d: [Nimg*H*W, CO] view (may be a column slice).
    bn: dict(ref, ref_is_raw, sc, sh, p0, p1, colsum) for the fused BN+ReLU-backward epilogue."""
    _lib.require_cuda(x2d, wp, out2d)
    M, _, ldx = _rows_pitch(x2d)
    Mo, COo, ldo = _rows_pitch(out2d)
    if M != Nimg * H * W or Mo != M or COo != CO:
        raise ValueError('conv3x3_bf16: shape mismatch')
    args = _bn_args(bn)
    call('gn_conv3x3_bf16', ptr(x2d), ldx, Nimg, H, W, CI, ptr(wp), wp.stride(0), CO, ptr(out2d), ldo, *args, stream())
    return out2d


def conv3x3_wgrad_into(x2d, dy2d, Nimg, H, W, CI, CO, dwp):
    """dwp [9, CI, CO] fp32 (+=): packed weight gradient of the 3x3 convolution (un-packed later in one batched launch)."""
    _lib.require_cuda(x2d, dy2d, dwp)
    M, _, ldx = _rows_pitch(x2d)
    M2, _, ldy = _rows_pitch(dy2d)
    if M != Nimg * H * W or M2 != M or tuple(dwp.shape) != (9, CI, CO) or not dwp.is_contiguous():
        raise ValueError('conv3x3_wgrad_into: shape mismatch')
    call('gn_conv3x3_wgrad_bf16', ptr(x2d), ldx, ptr(dy2d), ldy, Nimg, H, W, CI, CO, ptr(dwp), stream())
    return dwp


def conv3x3_wgrad_bf16(x2d, dy2d, Nimg, H, W, CI, CO):
    """-> fp32 (CO, CI, 3, 3) weight gradient of the 3x3 convolution from activations x2d [M, >=CI] and dy2d [M, >=CO]."""
    _lib.require_cuda(x2d, dy2d)
    M, _, ldx = _rows_pitch(x2d)
    M2, _, ldy = _rows_pitch(dy2d)
    if M != Nimg * H * W or M2 != M:
        raise ValueError('conv3x3_wgrad_bf16: shape mismatch')
    dwp = torch.zeros((9, CI, CO), device=x2d.device, dtype=torch.float32)
    call('gn_conv3x3_wgrad_bf16', ptr(x2d), ldx, ptr(dy2d), ldy, Nimg, H, W, CI, CO, ptr(dwp), stream())
    dw = torch.empty((CO, CI, 3, 3), device=x2d.device, dtype=torch.float32)
    call('gn_conv3x3_unpack_grad', ptr(dwp), CO, CI, ptr(dw), stream())
    return dw


# ---- stem (conv0 7x7 / stride 2 / pad 3 + norm0 + relu0) without an im2col buffer -----------------------------------
def stem_pack_input(x):
    """(N, 3, P, P) fp32 | bf16 -> NHWC4 bf16 (N, P, P, 4): RGB + a zero channel, 8 bytes per pixel."""
    _lib.require_cuda(x)
    N, _, P, _ = x.shape
    xq = torch.empty((N, P, P, 4), device=x.device, dtype=torch.bfloat16)
    call('gn_stem_pack_input', ptr(x), 1 if x.dtype == torch.bfloat16 else 0, N, P, ptr(xq), stream())
    return xq


def stem_pack_weight(w):
    """fp32 (CO, 3, 7, 7) -> bf16 [7, CO, 32] (per kernel row: 8 pixel slots x 4 channels, slot 0 and channel 3 zero)."""
    _lib.require_cuda(w)
    CO = int(w.shape[0])
    wq = torch.empty((7, CO, 32), device=w.device, dtype=torch.bfloat16)
    call('gn_stem_pack_weight', ptr(w.contiguous().float()), CO, ptr(wq), stream())
    return wq


def stem_conv_fwd(xq, wq, scale=None, shift=None, relu=False, out=None):
    """out[(n, oy, ox), co] = [relu](conv7x7s2(x)[n, co, oy, ox] * scale[co] + shift[co]), bf16 rows of CO channels."""
    _lib.require_cuda(xq, wq)
    N, P = int(xq.shape[0]), int(xq.shape[1])
    CO = int(wq.shape[1])
    Ho = P // 2
    if out is None:
        out = torch.empty((N * Ho * Ho, CO), device=xq.device, dtype=torch.bfloat16)
    _, _, ldo = _rows_pitch(out)
    call('gn_stem_conv_fwd', ptr(xq), N, P, ptr(wq), CO, ptr(scale), ptr(shift), 1 if relu else 0, ptr(out), ldo, stream())
    return out


def stem_conv_wgrad_into(xq, dz, CO, dwq):
    """dwq [CO, 224] fp32 (+=): packed conv0 weight gradient."""
    _lib.require_cuda(xq, dz, dwq)
    N, P = int(xq.shape[0]), int(xq.shape[1])
    _, _, ldz = _rows_pitch(dz)
    call('gn_stem_conv_wgrad', ptr(xq), N, P, ptr(dz), ldz, CO, ptr(dwq), stream())
    return dwq


def stem_conv_wgrad(xq, dz, CO):
    """-> fp32 (CO, 3, 7, 7) weight gradient from NHWC4 patches and dz [(n, oy, ox), >=CO] bf16."""
    _lib.require_cuda(xq, dz)
    N, P = int(xq.shape[0]), int(xq.shape[1])
    _, _, ldz = _rows_pitch(dz)
    dwq = torch.zeros((CO, 7 * 32), device=xq.device, dtype=torch.float32)
    call('gn_stem_conv_wgrad', ptr(xq), N, P, ptr(dz), ldz, CO, ptr(dwq), stream())
    dw = torch.empty((CO, 3, 7, 7), device=xq.device, dtype=torch.float32)
    call('gn_stem_unpack_wgrad', ptr(dwq), CO, ptr(dw), stream())
    return dw
