"""Count-MLP spot classifier f on the tcgen05 GEMM kernels.

The reference's count f is a user ``nn.Sequential`` of Linear / BatchNorm1d / ReLU modules
(/root/reference/notebooks/Tutorial_visium_count.ipynb cell 12: Linear(G,500) Linear(500,100) BN ReLU Linear(100,100)
Linear(100,50) BN ReLU Linear(50,n_cls)) applied to every one of the B*78*64 grid cells after a permute+reshape COPY of the
(B, G, 78, 64) slab (/root/reference/gridnext/gridnet_models.py:168,83).  Here the module tree and its parameters stay
what the user built (state-dict keys untouched); ``compile_count_mlp`` recognises the pattern and runs it as

  layer 1   out[m, d] = sum_g slab[g, m] * W1[d, g]   gn_gemm_tn_bf16 straight on the slab (rows = genes are the
                                                        reduction index, so the slab IS the MN-major operand: no copy)
  layer i   gn_gemm_bf16 on spot-major bf16 activations, Linear bias + eval-mode BatchNorm1d + ReLU in the epilogue
  backward  weight gradients by gn_gemm_tn_bf16 (bias gradients = the same GEMM against a column of ones), data gradients by
            gn_gemm_bf16 with the BatchNorm+ReLU backward (and the BN parameter column sums) fused in the epilogue.

BatchNorm1d must be in eval mode (training.py:126 puts ``patch_classifier`` in eval).  GridNetHexMM's count f keeps
train-mode BatchNorm during the train phase: that case, like any unrecognised module, is not
compiled (``compile_count_mlp`` returns None) and runs through the generic module call exactly as in the reference.
"""
import torch
import torch.nn as nn

from . import _lib, tc
from ._lib import ptr, stream, call

BF = torch.bfloat16


def _pad8(n):
    return (n + 7) // 8 * 8


class _Stage:
    __slots__ = ('lin', 'bn', 'relu')

    def __init__(self, lin):
        self.lin, self.bn, self.relu = lin, None, False


def _parse(module):
    if not isinstance(module, nn.Sequential) or len(module) == 0:
        return None
    stages = []
    for m in module:
        if isinstance(m, nn.Linear):
            stages.append(_Stage(m))
        elif isinstance(m, nn.BatchNorm1d):
            if not stages or stages[-1].bn is not None or stages[-1].relu or not m.affine or not m.track_running_stats:
                return None
            stages[-1].bn = m
        elif isinstance(m, nn.ReLU):
            if not stages or stages[-1].relu:
                return None
            stages[-1].relu = True
        else:
            return None
    if not stages or len(stages) < 2:
        return None
    for a, b in zip(stages[:-1], stages[1:]):
        if a.lin.out_features != b.lin.in_features:
            return None
    for s in stages:
        if s.bn is not None and (not s.relu or s.bn.num_features != s.lin.out_features):
            return None        # BN without ReLU is not a pattern of the hot path
        if s.lin.bias is None:
            return None
    if stages[-1].bn is not None or stages[-1].relu:
        return None
    return stages


def compile_count_mlp(module):
    """-> object with ``forward_grid(x, f_dim)`` or None when the module is not the Linear/BN1d(eval)/ReLU pattern."""
    stages = _parse(module)
    if stages is None:
        return None
    if any(s.bn is not None and s.bn.training for s in stages):
        return None
    return _Compiled(module, stages)


def _cast_bf16(x):
    out = torch.empty(x.shape, device=x.device, dtype=BF)
    call('gn_cast_f32_bf16', ptr(x), ptr(out), x.numel(), stream())
    return out


def _rows_to_bf16(x2d, C, scale=None, shift=None, relu=False):
    """fp32 [N, >=C] -> bf16 [N, pad8(C)] = [relu](x*scale+shift), pad columns zero."""
    N = x2d.shape[0]
    Cp = _pad8(C)
    out = torch.empty((N, Cp), device=x2d.device, dtype=BF)
    call('gn_rows_affine_bf16', ptr(x2d), x2d.stride(0), ptr(scale), ptr(shift), 1 if relu else 0, ptr(out), Cp, N, C, Cp, stream())
    return out


def _w_bf16(w):
    """fp32 [R, C] -> bf16 [R, pad8(C)] view [:, :C] (row pitch a multiple of 8 elements for TMA)."""
    R, C = w.shape
    buf = torch.zeros((R, _pad8(C)), device=w.device, dtype=BF)
    buf[:, :C] = w
    return buf[:, :C]


class _Compiled:
    def __init__(self, module, stages):
        self.module, self.stages = module, stages

    def forward_grid(self, x, f_dim):
        """x: (B, G, H, W) fp32 CUDA -> (B, f_dim, H, W) fp32 (the reference's permuted view of (N, f_dim))."""
        if x.dim() != 4 or x.dtype != torch.float32 or (x.shape[2] * x.shape[3]) % 8 != 0 or x.requires_grad:
            return None
        if self.stages[0].lin.in_features != x.shape[1] or self.stages[-1].lin.out_features != f_dim:
            return None
        B, G, H, W = x.shape
        out = _CountMLPFn.apply(x.contiguous(), self, *self.module.parameters())
        return out.reshape(B, H, W, f_dim).permute(0, 3, 1, 2)


def _fold(stage):
    """(scale, shift) of the stage's epilogue: y = (W x) * scale + shift  [then ReLU]."""
    b = stage.lin.bias.detach().float()
    if stage.bn is None:
        return None, b.contiguous(), None
    bn = stage.bn
    s = (bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)).contiguous()
    t = (bn.bias.detach().float() - bn.running_mean.float() * s + b * s).contiguous()
    return s, t, dict(beta=bn.bias.detach().float().contiguous(), inv_gamma=torch.where(bn.weight.detach() != 0, 1.0 / bn.weight.detach().float(), torch.zeros_like(bn.weight.detach().float())).contiguous())


class _CountMLPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, comp, *params):
        _lib.require_cuda(x)
        stages = comp.stages
        B, G, H, W = x.shape
        HW, N = H * W, B * H * W
        xb = _cast_bf16(x).view(B, G, HW)
        # ---- layer 1 straight on the slab
        st = stages[0]
        D1 = st.lin.out_features
        w1t = _w_bf16(st.lin.weight.detach().t().contiguous())                  # [G, D1]
        acc = torch.zeros((N, _pad8(D1)), device=x.device, dtype=torch.float32)
        for b in range(B):
            tc.gemm_tn_bf16(xb[b], w1t, acc[b * HW:(b + 1) * HW, :D1])
        s, t, _ = _fold(st)
        acts = [_rows_to_bf16(acc, D1, s, t, st.relu)]
        del acc
        # ---- layers 2..n on spot-major activations
        for i, st in enumerate(stages[1:], start=1):
            cin, cout = st.lin.in_features, st.lin.out_features
            s, t, _ = _fold(st)
            last = i == len(stages) - 1
            if last:
                o = torch.empty((N, cout), device=x.device, dtype=torch.float32)
                tc.gemm_bf16(acts[-1][:, :cin], _w_bf16(st.lin.weight.detach()), out=o, scale=s, shift=t, relu=st.relu)
                out = o
            else:
                o = torch.empty((N, _pad8(cout)), device=x.device, dtype=BF)
                tc.gemm_bf16(acts[-1][:, :cin], _w_bf16(st.lin.weight.detach()), out=o[:, :cout], scale=s, shift=t, relu=st.relu)
                acts.append(o)
        ctx.comp, ctx.xb, ctx.acts, ctx.dims = comp, xb, acts, (B, G, HW)
        ctx.plist = list(comp.module.parameters())
        return out

    @staticmethod
    def backward(ctx, dout):
        comp, xb, acts = ctx.comp, ctx.xb, ctx.acts
        stages = comp.stages
        B, G, HW = ctx.dims
        N = B * HW
        dev = dout.device
        grads = {}
        ones = torch.ones((N, 8), device=dev, dtype=BF)
        f = stages[-1].lin.out_features
        dY = _rows_to_bf16(dout.contiguous().float(), f)                         # gradient w.r.t. the last Linear's output
        for i in range(len(stages) - 1, -1, -1):
            st = stages[i]
            cin, cout = st.lin.in_features, st.lin.out_features
            dYv = dY[:, :cout]
            db = torch.zeros((cout, 8), device=dev, dtype=torch.float32)
            tc.gemm_tn_bf16(dYv, ones, db)
            grads[id(st.lin.bias)] = db[:, 0].contiguous()
            if i == 0:
                # dW1^T [G, D1] += slab_b [G, HW] @ dY1_b [HW, D1]
                D = torch.zeros((G, cout), device=dev, dtype=torch.float32)
                for b in range(B):
                    dyt = dYv[b * HW:(b + 1) * HW].t().contiguous()              # [D1, HW]
                    tc.gemm_bf16(xb[b], dyt, out=D, accumulate=True)
                grads[id(st.lin.weight)] = D.t().contiguous()
                break
            prev = stages[i - 1]
            a_prev = acts[i - 1][:, :cin]
            dW = torch.zeros((cout, cin), device=dev, dtype=torch.float32)
            tc.gemm_tn_bf16(dYv, a_prev, dW)
            grads[id(st.lin.weight)] = dW
            wt = _w_bf16(st.lin.weight.detach().t().contiguous())                # [cin, cout]
            dprev = torch.empty((N, _pad8(cin)), device=dev, dtype=BF)
            if prev.relu:
                s, _, extra = _fold(prev)
                if prev.bn is not None:
                    colsum = torch.zeros((2, cin), device=dev, dtype=torch.float32)
                    tc.gemm_bf16(dYv, wt, out=dprev[:, :cin],
                                 bn=dict(ref=a_prev, ref_is_raw=False, sc=s, sh=None, p0=extra['beta'], p1=extra['inv_gamma'], colsum=colsum))
                    grads[id(prev.bn.bias)] = colsum[0].clone()
                    grads[id(prev.bn.weight)] = colsum[1].clone()
                else:
                    one = torch.ones(cin, device=dev, dtype=torch.float32)
                    zero = torch.zeros(cin, device=dev, dtype=torch.float32)
                    tc.gemm_bf16(dYv, wt, out=dprev[:, :cin], bn=dict(ref=a_prev, ref_is_raw=False, sc=one, sh=None, p0=zero, p1=one, colsum=None))
            else:
                tc.gemm_bf16(dYv, wt, out=dprev[:, :cin])
            dY = dprev
        ctx.xb = ctx.acts = None
        return (None, None) + tuple(grads.get(id(p)) for p in ctx.plist)
