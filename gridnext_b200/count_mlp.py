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

BatchNorm1d in eval mode (training.py:126 puts ``patch_classifier`` in eval) folds into the GEMM epilogue.  GridNetHexMM's
count f keeps TRAIN-mode BatchNorm during the train phase (training.py:126 only reaches
``patch_classifier``), and so does f pre-training (training.py:11-98): there the Linear output is stored raw, its batch
statistics give the per-channel constants (csrc/bn_train.cu: gn_colstats_bf16 / gn_bn_train_coeffs, running statistics updated
as nn.BatchNorm1d does), BN+ReLU is a pass of its own, and the backward adds the mean/variance terms of the BatchNorm gradient
as dx -= c0 + c1*x after the fused eval-style epilogue.  Any unrecognised module pattern is not compiled
(``compile_count_mlp`` returns None) and runs through the generic module call exactly as in the reference.
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


def tensor_core_mlp_enabled():
    """GRIDNEXT_B200_MLP_TC=0 keeps a recognised count MLP on the generic fp32 module path (bit-for-bit what the reference's
    modules compute on the GPU); the default runs it on the bf16 tensor-core GEMMs (2e-2 of the fp32 result, north_star)."""
    import os
    return os.environ.get('GRIDNEXT_B200_MLP_TC', '1') != '0'


def compile_count_mlp(module):
    """-> object with ``forward_grid(x, f_dim)`` or None when the module is not the Linear/BN1d(eval)/ReLU pattern (or the
    tensor-core MLP path is switched off).  The parse is cached on the module and redone only when its children change."""
    if not tensor_core_mlp_enabled() or not isinstance(module, nn.Sequential):
        return None
    key = tuple(id(m) for m in module)
    cached = module.__dict__.get('_b200_compiled')
    if cached is not None and cached[0] == key:
        return cached[1]
    stages = _parse(module)
    compiled = None if stages is None else _Compiled(module, stages)
    module.__dict__['_b200_compiled'] = (key, compiled)
    return compiled


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

    def forward_spots(self, x):
        """x: (N, G) fp32 CUDA spot-major batch (f pre-training, training.py:11-98) -> (N, n_out) fp32, or None if unsupported."""
        if x.dim() != 2 or x.dtype != torch.float32 or x.requires_grad or self.stages[0].lin.in_features != x.shape[1]:
            return None
        return _CountMLPFn.apply(x.contiguous(), self, *self.module.parameters())


def _fold(stage):
    """(scale, shift, extra) of the stage's epilogue: y = (W x) * scale + shift  [then ReLU] (eval-mode BN folded in)."""
    b = stage.lin.bias.detach().float()
    if stage.bn is None or stage.bn.training:
        return None, b.contiguous(), None
    bn = stage.bn
    s = (bn.weight.detach().float() / torch.sqrt(bn.running_var.float() + bn.eps)).contiguous()
    t = (bn.bias.detach().float() - bn.running_mean.float() * s + b * s).contiguous()
    return s, t, dict(beta=bn.bias.detach().float().contiguous(), inv_gamma=torch.where(bn.weight.detach() != 0, 1.0 / bn.weight.detach().float(), torch.zeros_like(bn.weight.detach().float())).contiguous())


def _bn_train_consts(bn, raw, cout):
    """Batch statistics of the raw Linear output [N, pad8(cout)] -> padded (sc, sh, mean, invstd) rows; running stats updated."""
    N, Cp = raw.shape
    if N < 2:
        raise ValueError('Expected more than 1 value per channel when training, got %d' % N)
    st = torch.zeros((2, Cp), device=raw.device, dtype=torch.float64)
    call('gn_colstats_bf16', ptr(raw), raw.stride(0), N, Cp, ptr(st[0]), ptr(st[1]), stream())
    k = torch.zeros((4, Cp), device=raw.device, dtype=torch.float32)
    mom = bn.momentum if bn.momentum is not None else 1.0 / (int(bn.num_batches_tracked.item()) + 1)
    track = bn.track_running_stats and bn.running_mean is not None
    call('gn_bn_train_coeffs', ptr(st[0]), ptr(st[1]), N, ptr(bn.weight.detach()), ptr(bn.bias.detach()), float(bn.eps), float(mom),
         ptr(bn.running_mean) if track else None, ptr(bn.running_var) if track else None, ptr(k[0]), ptr(k[1]), ptr(k[2]), ptr(k[3]), cout, stream())
    if track:
        bn.num_batches_tracked += 1
    return dict(sc=k[0], sh=k[1], mean=k[2], invstd=k[3])


def _bn_relu_pass(raw, k, relu):
    act = torch.empty_like(raw)
    call('gn_affine_relu_bf16', ptr(raw), raw.stride(0), ptr(act), act.stride(0), raw.shape[0], raw.shape[1], ptr(k['sc']), ptr(k['sh']),
         1 if relu else 0, stream())
    return act


class _CountMLPFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, comp, *params):
        _lib.require_cuda(x)
        stages = comp.stages
        rows = x.dim() == 2                     # spot-major (N, G) batch instead of the (B, G, H, W) slab
        if rows:
            N, G = x.shape
            B, HW = 1, N
            xb = _rows_to_bf16(x, G)            # [N, pad8(G)]
        else:
            B, G, H, W = x.shape
            HW, N = H * W, B * H * W
            xb = _cast_bf16(x).view(B, G, HW)
        # per stage: the operand the next stage consumes (acts), and for train-mode BN the raw Linear output + batch constants
        acts, raws, tks = [], [None] * len(stages), [None] * len(stages)
        # ---- layer 1 straight on the slab
        st = stages[0]
        D1 = st.lin.out_features
        acc = torch.zeros((N, _pad8(D1)), device=x.device, dtype=torch.float32)
        if rows:
            tc.gemm_bf16(xb[:, :G], _w_bf16(st.lin.weight.detach()), out=acc[:, :D1])
        else:
            w1t = _w_bf16(st.lin.weight.detach().t().contiguous())              # [G, D1]
            for b in range(B):
                tc.gemm_tn_bf16(xb[b], w1t, acc[b * HW:(b + 1) * HW, :D1])
        s, t, _ = _fold(st)
        if st.bn is not None and st.bn.training:
            raws[0] = _rows_to_bf16(acc, D1, None, t, False)
            tks[0] = _bn_train_consts(st.bn, raws[0], D1)
            acts.append(_bn_relu_pass(raws[0], tks[0], st.relu))
        else:
            acts.append(_rows_to_bf16(acc, D1, s, t, st.relu))
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
            elif st.bn is not None and st.bn.training:
                raw = torch.zeros((N, _pad8(cout)), device=x.device, dtype=BF)
                tc.gemm_bf16(acts[-1][:, :cin], _w_bf16(st.lin.weight.detach()), out=raw[:, :cout], shift=t)
                raws[i] = raw
                tks[i] = _bn_train_consts(st.bn, raw, cout)
                acts.append(_bn_relu_pass(raw, tks[i], st.relu))
            else:
                o = torch.empty((N, _pad8(cout)), device=x.device, dtype=BF)
                tc.gemm_bf16(acts[-1][:, :cin], _w_bf16(st.lin.weight.detach()), out=o[:, :cout], scale=s, shift=t, relu=st.relu)
                acts.append(o)
        ctx.comp, ctx.xb, ctx.acts, ctx.raws, ctx.tks, ctx.dims, ctx.rows = comp, xb, acts, raws, tks, (B, G, HW), rows
        ctx.plist = list(comp.module.parameters())
        return out

    @staticmethod
    def backward(ctx, dout):
        comp, xb, acts, raws, tks = ctx.comp, ctx.xb, ctx.acts, ctx.raws, ctx.tks
        if acts is None:
            raise RuntimeError('count MLP (B200): backward a second time (saved activations were released; use a fresh forward)')
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
            if i == 0 and ctx.rows:
                dW = torch.zeros((cout, G), device=dev, dtype=torch.float32)
                tc.gemm_tn_bf16(dYv, xb[:, :G], dW)
                grads[id(st.lin.weight)] = dW
                break
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
            cinp = _pad8(cin)
            dprev = (torch.zeros if cinp != cin else torch.empty)((N, cinp), device=dev, dtype=BF)
            if prev.relu and tks[i - 1] is not None:
                # train-mode BN of the previous stage: eval-style fused epilogue on the RAW Linear output, then the mean/variance terms
                k, raw = tks[i - 1], raws[i - 1]
                colsum = torch.zeros((2, cin), device=dev, dtype=torch.float32)
                tc.gemm_bf16(dYv, wt, out=dprev[:, :cin],
                             bn=dict(ref=raw[:, :cin], ref_is_raw=True, sc=k['sc'], sh=k['sh'], p0=k['mean'], p1=k['invstd'], colsum=colsum))
                grads[id(prev.bn.bias)] = colsum[0].clone()
                grads[id(prev.bn.weight)] = colsum[1].clone()
                Fz = torch.zeros((2, cinp), device=dev, dtype=torch.float32)
                call('gn_bn_train_fix_coeffs', ptr(colsum[0]), ptr(colsum[1]), ptr(k['sc']), ptr(k['invstd']), ptr(k['mean']), N, 0,
                     ptr(Fz[0]), ptr(Fz[1]), cin, stream())
                call('gn_bn_train_fix_bf16', ptr(dprev), cinp, ptr(raw), raw.stride(0), N, cinp, ptr(Fz[0]), ptr(Fz[1]), stream())
            elif prev.relu:
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
        ctx.xb = ctx.acts = ctx.raws = ctx.tks = None
        return (None, None) + tuple(grads.get(id(p)) for p in ctx.plist)
