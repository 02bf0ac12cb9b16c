"""Python faces of the tensor-core (tcgen05) entry points.  Thin: argument checks + the ctypes call."""
import torch
from . import _lib
from ._lib import ptr, stream, call


def _rows_pitch(t):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError('expected a 2-D tensor with unit column stride, got strides %s' % (t.stride(),))
    return t.shape[0], t.shape[1], t.stride(0)


def gemm_bf16(a, b, out=None, out_dtype=torch.bfloat16, accumulate=False, scale=None, shift=None, relu=False,
              xf_scale=None, xf_shift=None):
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
         1 if accumulate else 0, ptr(scale), ptr(shift), 1 if relu else 0, ptr(xf_scale), ptr(xf_shift), stream())
    return out
