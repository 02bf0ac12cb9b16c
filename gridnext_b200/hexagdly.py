"""Drop-in ``hexagdly.Conv2d`` backed by the sm_100a hexconv kernels.

Mirrors the third-party class the reference imports at /root/reference/gridnext/gridnet_models.py:10
and instantiates at :130-147: ``Conv2d(in_channels, out_channels, kernel_size=1, stride=1, bias=True,
debug=False)``, parameters ``kernel0..kernel{k}`` of shape (Cout, Cin, 2k+1-i, 1 | 2) and
``bias_tensor``, all U(-1/sqrt(n), 1/sqrt(n)), n = in_channels * (1 + 3k(k+1)).

A stand-alone module call takes HexagDLy's own layout (B, C, rows, cols) with odd 0-indexed columns
shifted down; inside ``GridNetHex*`` the corrector runs through ``corrector.run_corrector`` directly
in the Visium layout instead, without any re-indexing copies.
"""
import math
import ctypes
import os

import torch
import torch.nn as nn

from . import _lib
from ._lib import ptr, stream, call


def n_taps(k):
    return 1 + 3 * k * (k + 1)


def _kernels(mod):
    return [getattr(mod, 'kernel%d' % i) for i in range(mod.hexbase_size + 1)]


def _kptrs(ks):
    p = [ptr(t) for t in ks]
    return p + [None] * (4 - len(p))


def pack_weights(ks, ksize, cin, cout, mode, kind='hex'):
    """[T][cin][cout] (mode 0) or reflected/transposed [T][cout][cin] (mode 1) packed copy.
    kind 'sq': Cartesian ksize x ksize window (nn.Conv2d of the base GridNet corrector), ks = [weight (Cout, Cin, K, K)]."""
    T = ksize * ksize if kind == 'sq' else n_taps(ksize)
    wp = torch.empty((T, cin, cout) if mode == 0 else (T, cout, cin), device=ks[0].device, dtype=torch.float32)
    if kind == 'sq':
        call('gn_sqconv_pack', ptr(ks[0]), ksize, cin, cout, mode, ptr(wp), stream())
    else:
        call('gn_hexconv_pack', *_kptrs(ks), ksize, cin, cout, mode, ptr(wp), stream())
    return wp


# 'auto': tensor cores once the convolution is throughput-bound (>= 16 Visium arrays' worth of cells); below that the g network
# is launch-latency-bound and the exact-fp32 FMA kernel costs the same.  '1' / '0' force one path (tests, benchmarks).
TENSOR_CORE_MODE = os.environ.get('GRIDNEXT_B200_HEX_TC', 'auto')
TENSOR_CORE_MIN_CELLS = 16 * 78 * 64
# '2': the second-generation kernel (csrc/hexconv_tc2.cu: fp32 NCHW in and out, operands converted in shared memory) wherever it
# supports the shape (grid width <= 64 and a multiple of 4), else the first generation; '1' forces the first generation.
TENSOR_CORE_GEN = os.environ.get('GRIDNEXT_B200_HEX_TC_GEN', '2')


def _use_tc(B, cin, cout, H, W, ksize):
    if TENSOR_CORE_MODE == '0' or not _lib.load().gn_hexconv_tc_supported(cin, cout, H, W, ksize):
        return False
    return TENSOR_CORE_MODE == '1' or B * H * W >= TENSOR_CORE_MIN_CELLS


def hexconv_fwd(x, wp, bias, cout, ksize, in_scale=None, in_shift=None, stats=None, kind='hex'):
    """y = hexconv(x') + bias.  kernel_size 1 with <= 32 channels runs on tcgen05 (bf16 x 3 split, fp32 accumulate: the
    7-tap, 32-channel convolution is far above the FP32-FMA ridge); everything else on the FP32-FMA kernel."""
    B, cin, H, W = x.shape
    y = torch.empty((B, cout, H, W), device=x.device, dtype=torch.float32)
    if kind == 'sq':
        call('gn_sqconv_fwd', ptr(x), ptr(wp), ptr(bias), ptr(in_scale), ptr(in_shift), ptr(y), ptr(stats), B, cin, cout, H, W, ksize, stream())
        return y
    if _use_tc(B, cin, cout, H, W, ksize):
        lib = _lib.load()
        if TENSOR_CORE_GEN != '1' and lib.gn_hexconv_tc2_supported(cin, cout, H, W, ksize):
            ws, wptr = _aligned_workspace(lib.gn_hexconv_tc2_workspace_bytes(), x.device)
            call('gn_hexconv_fwd_tc2', ptr(x), ptr(wp), ptr(bias), ptr(in_scale), ptr(in_shift), ptr(y), ptr(stats), B, cin, cout, H, W, wptr, stream())
            return y
        ws, wptr = _aligned_workspace(_lib.load().gn_hexconv_tc_workspace_bytes(B, H, W), x.device)
        call('gn_hexconv_fwd_tc', ptr(x), ptr(wp), ptr(bias), ptr(in_scale), ptr(in_shift), ptr(y), ptr(stats), B, cin, cout, H, W, wptr, stream())
        return y
    call('gn_hexconv_fwd', ptr(x), ptr(wp), ptr(bias), ptr(in_scale), ptr(in_shift), ptr(y), ptr(stats),
         B, cin, cout, H, W, ksize, stream())
    return y


def _aligned_workspace(nbytes, device):
    ws = torch.empty((int(nbytes) + 1024,), device=device, dtype=torch.uint8)
    return ws, ctypes.c_void_p(ws.data_ptr() + (-ws.data_ptr()) % 1024)


def hexconv_wgrad(x, dy, ksize, in_scale=None, in_shift=None, want_bias=True, kind='hex'):
    B, cin, H, W = x.shape
    cout = dy.shape[1]
    if kind == 'sq':
        dwp = torch.zeros((ksize * ksize, cin, cout), device=x.device, dtype=torch.float32)
        db = torch.zeros((cout,), device=x.device, dtype=torch.float32) if want_bias else None
        call('gn_sqconv_wgrad', ptr(x), ptr(in_scale), ptr(in_shift), ptr(dy), ptr(dwp), ptr(db), B, cin, cout, H, W, ksize, stream())
        return dwp, db
    dwp = torch.zeros((n_taps(ksize), cin, cout), device=x.device, dtype=torch.float32)
    if _use_tc(B, cin, cout, H, W, ksize):
        if (TENSOR_CORE_GEN != '1' and _lib.load().gn_hexconv_tc2_supported(cin, cout, H, W, ksize)
                and x.data_ptr() % 16 == 0 and dy.data_ptr() % 16 == 0):       # 16-byte asynchronous copies; an odd view takes generation 1
            db = torch.zeros((cout,), device=x.device, dtype=torch.float32) if want_bias else None
            call('gn_hexconv_wgrad_tc2', ptr(x), ptr(in_scale), ptr(in_shift), ptr(dy), ptr(dwp), ptr(db), B, cin, cout, H, W, stream())
            return dwp, db
        ws, wptr = _aligned_workspace(_lib.load().gn_hexconv_tc_wgrad_workspace_bytes(B, H, W), x.device)
        call('gn_hexconv_wgrad_tc', ptr(x), ptr(in_scale), ptr(in_shift), ptr(dy), ptr(dwp), B, cin, cout, H, W, wptr, stream())
        db = None
        if want_bias:
            st = torch.zeros(2 * cout, device=x.device, dtype=torch.float64)
            call('gn_bn_stats', ptr(dy), ptr(st), B, cout, H * W, stream())
            db = st[:cout].float()
        return dwp, db
    db = torch.zeros((cout,), device=x.device, dtype=torch.float32) if want_bias else None
    call('gn_hexconv_wgrad', ptr(x), ptr(in_scale), ptr(in_shift), ptr(dy), ptr(dwp), ptr(db), B, cin, cout, H, W, ksize, stream())
    return dwp, db


def unpack_grad(dwp, shapes, ksize, cin, cout, kind='hex'):
    gs = [torch.empty(s, device=dwp.device, dtype=torch.float32) for s in shapes]
    if kind == 'sq':
        call('gn_sqconv_unpack_grad', ptr(dwp), ptr(gs[0]), ksize, cin, cout, stream())
    else:
        call('gn_hexconv_unpack_grad', ptr(dwp), *_kptrs(gs), ksize, cin, cout, stream())
    return gs


class _HexConvFn(torch.autograd.Function):
    """Single hexagonal convolution in the Visium layout (parity on dim 2)."""

    @staticmethod
    def forward(ctx, x, bias, ksize, *ks):
        _lib.require_cuda(x, *ks)
        x = x.contiguous().float()
        ks = [k.contiguous() for k in ks]
        cout, cin = ks[0].shape[0], ks[0].shape[1]
        if x.shape[1] != cin:
            raise ValueError('hexagdly.Conv2d: expected %d input channels, got %d' % (cin, x.shape[1]))
        wp = pack_weights(ks, ksize, cin, cout, 0)
        y = hexconv_fwd(x, wp, bias, cout, ksize)
        ctx.save_for_backward(x, *ks)
        ctx.ksize, ctx.has_bias = ksize, bias is not None
        return y

    @staticmethod
    def backward(ctx, dy):
        x, *ks = ctx.saved_tensors
        ksize = ctx.ksize
        cout, cin = ks[0].shape[0], ks[0].shape[1]
        dy = dy.contiguous().float()
        dx = None
        if ctx.needs_input_grad[0]:
            wpt = pack_weights(ks, ksize, cin, cout, 1)
            dx = hexconv_fwd(dy, wpt, None, cin, ksize)
        dwp, db = hexconv_wgrad(x, dy, ksize, want_bias=ctx.has_bias)
        gks = unpack_grad(dwp, [k.shape for k in ks], ksize, cin, cout)
        return (dx, db, None) + tuple(gks)


def hexconv_visium(x, ks, bias=None):
    """Functional hexagonal convolution on (B, C, H, W) with the hex parity taken on the row index.
    kernel_size 1, <= 32 channels, below the tensor-core batch threshold: the persistent one-launch kernels of the corrector
    (csrc/corrector_fused.cu) with a single stage -- all SMs busy at any batch size, no pack / unpack launches."""
    ksize = len(ks) - 1
    if ksize == 1 and x.is_cuda:
        from . import corrector as _corr
        cout, cin = int(ks[0].shape[0]), int(ks[0].shape[1])
        meta = [dict(kind='hex', ksize=1, cin=cin, cout=cout, nk=2, has_bias=bias is not None, bn=None, relu=False)]
        if x.dim() == 4 and x.shape[1] == cin and _corr._fused_eligible(meta, x):
            params = list(ks) + ([bias] if bias is not None else [])
            return _corr._CorrectorFusedFn.apply(x, meta, *params)
    return _HexConvFn.apply(x, bias, ksize, *ks)


class Conv2d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, bias=True, debug=False):
        super().__init__()
        if stride != 1:
            raise NotImplementedError('gridnext_b200.hexagdly.Conv2d: only stride 1 is on the GridNet hot path')
        if not 1 <= kernel_size <= 3:
            raise NotImplementedError('gridnext_b200.hexagdly.Conv2d: kernel_size must be 1, 2 or 3')
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.hexbase_size = kernel_size
        self.hexbase_stride = stride
        self.debug = debug
        self.bias = bias
        for i in range(kernel_size + 1):
            shape = (out_channels, in_channels, 2 * kernel_size + 1 - i, 1 if i == 0 else 2)
            setattr(self, 'kernel' + str(i), nn.Parameter(torch.empty(shape)))
        if bias:
            self.bias_tensor = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter('bias_tensor', None)
        self.init_parameters(debug)

    def init_parameters(self, debug):
        if debug:
            for i in range(self.hexbase_size + 1):
                getattr(self, 'kernel' + str(i)).data.fill_(1.0)
            if self.bias_tensor is not None:
                self.bias_tensor.data.fill_(0.0)
        else:
            stdv = 1.0 / math.sqrt(self.in_channels * n_taps(self.hexbase_size))
            for p in self.parameters():
                p.data.uniform_(-stdv, stdv)

    def forward_visium(self, x):
        return hexconv_visium(x, _kernels(self), self.bias_tensor)

    def forward(self, x):
        # HexagDLy layout (B, C, rows, cols): parity lives on the LAST index -> transpose in and out
        return self.forward_visium(x.transpose(2, 3).contiguous()).transpose(2, 3).contiguous()

    def __repr__(self):
        return 'Conv2d({}, {}, kernel_size={}, stride={})'.format(
            self.in_channels, self.out_channels, self.hexbase_size, self.hexbase_stride)
