"""Foreground-masked cross-entropy of the grid-wise training step as one fused CUDA op.

Reference: /root/reference/gridnext/training.py:152-160 -- permute/reshape, boolean-mask gathers,
``labels -= 1``, ``nn.CrossEntropyLoss()`` (mean over foreground spots), ``torch.max(outputs, 1)``.
"""
import torch
from . import _lib
from ._lib import ptr, stream, call


class _MaskedCEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, loss_scale, n_fg_override):
        _lib.require_cuda(logits, labels)
        ctx.in_dtype = logits.dtype
        logits = logits.contiguous().float()
        labels = labels.contiguous()
        if labels.dtype != torch.int64:
            labels = labels.long()
        B, C, H, W = logits.shape
        if tuple(labels.shape) != (B, H, W):
            raise ValueError('masked_cross_entropy: labels %s do not match logits %s' % (tuple(labels.shape), tuple(logits.shape)))
        acc = torch.empty(4, device=logits.device, dtype=torch.float64)
        loss = torch.empty(1, device=logits.device, dtype=torch.float32)
        need_grad = ctx.needs_input_grad[0]
        dlogits = torch.empty_like(logits) if need_grad else None
        call('gn_masked_ce', ptr(logits), ptr(labels), ptr(dlogits), ptr(acc), ptr(loss), ptr(n_fg_override), float(loss_scale),
             B, C, H * W, stream())
        ctx.dlogits = dlogits
        ctx.mark_non_differentiable(acc)
        return loss.reshape(()), acc

    @staticmethod
    def backward(ctx, dloss, _dacc):
        d = ctx.dlogits
        ctx.dlogits = None
        if d is None:
            return None, None, None, None
        return (d * dloss).to(ctx.in_dtype), None, None, None


def masked_cross_entropy(logits, labels, accum_iters=1, n_fg_override=None):
    """logits (B, C, H, W) fp32, labels (B, H, W) int64 with 0 = background, classes 1..C.

    Returns (loss, acc): loss = mean CE over foreground spots / accum_iters (0-dim tensor with grad),
    acc = fp64[4] device tensor {sum of spot losses, n_foreground, n_correct, n labels > C}.  A label above C makes
    nn.CrossEntropyLoss raise on the host; here it is counted in acc[3] (the training loop checks it at its one host read
    per phase) so the step stays free of host synchronisation."""
    return _MaskedCEFn.apply(logits, labels, 1.0 / accum_iters, n_fg_override)


MAX_FUSED_CLASSES = 64          # CE_MAX_C of csrc/corrector_ops.cu; wider outputs take the generic (reference) path


def is_plain_cross_entropy(criterion):
    """True when ``criterion`` is an nn.CrossEntropyLoss the fused kernel reproduces exactly."""
    import torch.nn as nn
    return (type(criterion) is nn.CrossEntropyLoss and criterion.weight is None and criterion.reduction == 'mean'
            and getattr(criterion, 'label_smoothing', 0.0) == 0.0 and getattr(criterion, 'ignore_index', -100) < 0)
