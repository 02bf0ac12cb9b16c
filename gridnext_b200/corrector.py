"""Fused execution of the g corrector (forward + backward): hexagonal in the Visium layout, or the base GridNet's Cartesian one.

Runs the ``nn.Sequential`` built by ``GridNetHex._init_corrector``
(/root/reference/gridnext/gridnet_models.py:128-148: hex hex [BN] ReLU hex hex [BN] ReLU hex) as ONE
autograd node: BatchNorm statistics come out of the preceding hexconv's epilogue, BatchNorm-apply +
ReLU is the next hexconv's prologue (the activated tensor is never materialised), and the rot90/flip
re-indexing of gridnet_models.py:177-185 is gone because the kernels take the row parity directly.

The base ``GridNet``'s Cartesian corrector (gridnet_models.py:51-66: Conv2d 3x3, 5x5, 5x5, 3x3 with [BN] ReLU in between) runs
through the same node: a K x K window is a parity-independent tap table of the same tile kernels (csrc/hexconv.cu, gn_sqconv_*).
"""
import torch
import torch.nn as nn

from . import _lib
from ._lib import ptr, stream, call
from . import hexagdly as hx
from . import parallel


def _is_square_conv(m):
    if not isinstance(m, nn.Conv2d):
        return False
    K = m.kernel_size[0]
    return (m.kernel_size == (K, K) and K in (1, 3, 5) and m.stride == (1, 1) and m.padding == (K // 2, K // 2) and m.dilation == (1, 1)
            and m.groups == 1 and m.padding_mode == 'zeros')


def parse_corrector(seq):
    """-> list of stages [(conv_module, bn_module | None, relu_before: bool)] or None if not fusable.

    Stage j = optional (BatchNorm2d, ReLU | ReLU) applied to the running tensor, then a hex conv (hexagdly.Conv2d) or a
    Cartesian one (nn.Conv2d K x K, stride 1, 'same' zero padding)."""
    if not isinstance(seq, nn.Sequential):
        return None
    stages, bn, relu = [], None, False
    for m in seq:
        if isinstance(m, hx.Conv2d) or _is_square_conv(m):
            if bn is not None and not relu:
                return None           # BN without ReLU before a conv: not a pattern we fuse
            stages.append((m, bn, relu))
            bn, relu = None, False
        elif isinstance(m, nn.BatchNorm2d):
            if bn is not None or relu or not stages or not m.track_running_stats or not m.affine:
                return None
            bn = m
        elif isinstance(m, nn.ReLU):
            if relu or not stages:
                return None
            relu = True
        else:
            return None
    if bn is not None or relu or not stages:
        return None
    return stages


class _CorrectorFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, meta, *params):
        """meta: list per stage of dict(ksize, cin, cout, nk, has_bias, bn: None | dict(training, momentum, eps), relu)
        params: flattened per stage: kernels..., [bias], [bn.weight, bn.bias]; BN running buffers travel in meta."""
        _lib.require_cuda(x)
        x = x.contiguous().float()
        B, _, H, W = x.shape
        dev = x.device
        saved = []          # per stage: input tensor, scale, shift, mean_invstd
        it = iter(params)
        cur = x
        cur_stats = None    # fp64 [2*C] stats of `cur` if a BN follows
        per_stage_params = []
        for j, st in enumerate(meta):
            ks = [next(it).contiguous() for _ in range(st['nk'])]
            bias = next(it) if st['has_bias'] else None
            gamma = beta = None
            scale = shift = mi = None
            if st['bn'] is not None:
                gamma, beta = next(it), next(it)
                C = st['cin']
                scale = torch.empty(C, device=dev); shift = torch.empty(C, device=dev); mi = torch.empty(2 * C, device=dev)
                bn = st['bn']
                if bn['training']:
                    count = float(B * H * W)
                    if bn['sync']:          # SyncBN: statistics over every rank's cells (equal per-rank batch sizes)
                        parallel.allreduce_sum_(cur_stats)
                        count *= parallel.world_size()
                    call('gn_bn_finalize', ptr(cur_stats), ptr(gamma), ptr(beta), ptr(bn['running_mean']), ptr(bn['running_var']),
                         bn['momentum'], bn['eps'], count, ptr(scale), ptr(shift), ptr(mi), C, 1, stream())
                else:
                    call('gn_bn_eval_affine', ptr(gamma), ptr(beta), ptr(bn['running_mean']), ptr(bn['running_var']), bn['eps'],
                         ptr(scale), ptr(shift), ptr(mi), C, stream())
            elif st['relu']:
                C = st['cin']
                scale = torch.ones(C, device=dev); shift = torch.zeros(C, device=dev)
            if st.get('tr'):          # Cartesian conv inside a hexagonal model: the reference applies it in HexagDLy layout = transposed
                ks = [ks[0].transpose(2, 3).contiguous()]
            wp = hx.pack_weights(ks, st['ksize'], st['cin'], st['cout'], 0, st['kind'])
            nxt = meta[j + 1] if j + 1 < len(meta) else None
            want_stats = nxt is not None and nxt['bn'] is not None and nxt['bn']['training']
            stats = torch.zeros(2 * st['cout'], device=dev, dtype=torch.float64) if want_stats else None
            out = hx.hexconv_fwd(cur, wp, bias, st['cout'], st['ksize'], scale, shift, stats, st['kind'])
            saved.append((cur, scale, shift, mi))
            per_stage_params.append((ks, bias, gamma, beta))
            cur, cur_stats = out, stats
        ctx.meta = meta
        ctx.saved = saved
        ctx.stage_params = per_stage_params
        ctx.dims = (B, H, W)
        return cur

    @staticmethod
    def backward(ctx, dout):
        if ctx.saved is None:
            raise RuntimeError('corrector: backward a second time (saved activations were released; use a fresh forward)')
        meta, saved, sp = ctx.meta, ctx.saved, ctx.stage_params
        B, H, W = ctx.dims
        grad = dout.contiguous().float()
        dev = grad.device
        grads = [None] * len(meta)
        for j in range(len(meta) - 1, -1, -1):
            st = meta[j]
            inp, scale, shift, mi = saved[j]
            ks, bias, gamma, beta = sp[j]
            dwp, db = hx.hexconv_wgrad(inp, grad, st['ksize'], scale, shift, want_bias=bias is not None, kind=st['kind'])
            gks = hx.unpack_grad(dwp, [k.shape for k in ks], st['ksize'], st['cin'], st['cout'], st['kind'])
            if st.get('tr'):
                gks = [gks[0].transpose(2, 3).contiguous()]
            dgamma = dbeta = None
            need_dx = j > 0 or ctx.needs_input_grad[0]
            if need_dx:
                wpt = hx.pack_weights(ks, st['ksize'], st['cin'], st['cout'], 1, st['kind'])
                dA = hx.hexconv_fwd(grad, wpt, None, st['cin'], st['ksize'], kind=st['kind'])
                if scale is not None:
                    C = st['cin']
                    dH = torch.empty_like(dA)
                    sums = torch.empty(2 * C, device=dev, dtype=torch.float64)
                    if st['bn'] is not None:
                        dgamma = torch.empty(C, device=dev); dbeta = torch.empty(C, device=dev)
                        training = 1 if st['bn']['training'] else 0
                    else:
                        mi = torch.zeros(2 * C, device=dev); mi[C:] = 1.0
                        training = 0
                    if st['bn'] is not None and st['bn']['sync'] and training:
                        # SyncBN backward: this rank's parameter gradients from ITS sums, the data gradient from the global ones
                        call('gn_bn_act_bwd_reduce', ptr(dA), ptr(inp), ptr(scale), ptr(shift), ptr(mi), ptr(sums), B, C, H * W, 1, stream())
                        dbeta.copy_(sums[:C])
                        dgamma.copy_(sums[C:])
                        parallel.allreduce_sum_(sums)
                        call('gn_bn_act_bwd_apply', ptr(dA), ptr(inp), ptr(scale), ptr(shift), ptr(mi), ptr(sums),
                             float(B * H * W) * parallel.world_size(), training, ptr(dH), None, None, B, C, H * W, 1, stream())
                    else:
                        call('gn_bn_act_bwd', ptr(dA), ptr(inp), ptr(scale), ptr(shift), ptr(mi), ptr(sums), float(B * H * W), training,
                             ptr(dH), ptr(dgamma), ptr(dbeta), B, C, H * W, 1, stream())
                    grad = dH
                else:
                    grad = dA
            elif st['bn'] is not None:
                raise RuntimeError('corrector: first stage cannot have a BatchNorm prologue')
            grads[j] = (gks, db, dgamma, dbeta)
        flat = []
        for j, st in enumerate(meta):
            gks, db, dgamma, dbeta = grads[j]
            flat.extend(gks)
            if st['has_bias']:
                flat.append(db)
            if st['bn'] is not None:
                flat.extend([dgamma, dbeta])
        dx = grad if ctx.needs_input_grad[0] else None
        ctx.saved = None
        return (dx, None) + tuple(flat)


# ---- the whole corrector in one launch per direction (csrc/corrector_fused.cu) ---------------------------------------------------
# 'auto': below the batch size where the tensor-core hexconv takes over, all-hexagonal kernel_size-1 correctors up to 32 channels wide
# run as ONE persistent kernel forward and ONE backward; '0' disables, '1' forces it for every eligible corrector.
import os as _os
import ctypes as _ct
FUSED_MODE = _os.environ.get('GRIDNEXT_B200_G_FUSED', 'auto')
_SYNC = {}


def _fused_eligible(meta, x):
    if FUSED_MODE == '0' or len(meta) > 8:
        return False
    modes = set()
    for m in meta:
        if m['kind'] != 'hex' or m['ksize'] != 1 or m['cin'] > 32 or m['cout'] > 32:
            return False
        if m['bn'] is not None:
            if m['bn']['sync']:
                return False
            modes.add(bool(m['bn']['training']))
    if len(modes) > 1 or meta[0]['bn'] is not None or meta[0]['relu']:
        return False
    if FUSED_MODE == '1':
        return True
    return x.shape[0] * x.shape[2] * x.shape[3] < hx.TENSOR_CORE_MIN_CELLS


def _ptr_array(ptrs):
    return (_ct.c_void_p * len(ptrs))(*[p.value if isinstance(p, _ct.c_void_p) else p for p in ptrs])


class _CorrectorFusedFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, meta, *params):
        _lib.require_cuda(x)
        x = x.contiguous().float()
        B, _, H, W = x.shape
        dev = x.device
        L = len(meta)
        it = iter(params)
        per_stage, pptr = [], []
        for st in meta:
            k0, k1 = next(it).contiguous(), next(it).contiguous()
            bias = next(it) if st['has_bias'] else None
            gamma = beta = None
            rm = rv = None
            if st['bn'] is not None:
                gamma, beta = next(it), next(it)
                rm, rv = st['bn']['running_mean'], st['bn']['running_var']
            per_stage.append((k0, k1, bias, gamma, beta))
            pptr += [ptr(k0), ptr(k1), ptr(bias), ptr(gamma), ptr(beta), ptr(rm), ptr(rv)]
        cin = (_ct.c_int * L)(*[st['cin'] for st in meta])
        cout = (_ct.c_int * L)(*[st['cout'] for st in meta])
        pro = (_ct.c_int * L)(*[2 if st['bn'] is not None else (1 if st['relu'] else 0) for st in meta])
        mom = (_ct.c_float * L)(*[st['bn']['momentum'] if st['bn'] is not None else 0.1 for st in meta])
        eps = (_ct.c_float * L)(*[st['bn']['eps'] if st['bn'] is not None else 1e-5 for st in meta])
        bn_training = 1 if any(st['bn'] is not None and st['bn']['training'] for st in meta) else 0
        acts = [torch.empty((B, st['cout'], H, W), device=dev, dtype=torch.float32) for st in meta]
        stats = torch.zeros((L, 64), device=dev, dtype=torch.float64)
        consts = torch.empty((L, 128), device=dev, dtype=torch.float32)
        sync = _SYNC.get(dev)
        if sync is None:
            sync = _SYNC[dev] = torch.zeros(2, device=dev, dtype=torch.int32)
        call('gn_corrector_fused_fwd', ptr(x), _ptr_array([ptr(a) for a in acts]), _ptr_array(pptr), cin, cout, pro, mom, eps, L, B, H, W, bn_training,
             ptr(stats), ptr(consts), ptr(sync), stream())
        ctx.meta, ctx.per_stage, ctx.x, ctx.acts, ctx.consts, ctx.stats = meta, per_stage, x, acts, consts, stats
        ctx.arrs = (pptr, cin, cout, pro, eps, bn_training, sync)
        return acts[-1]

    @staticmethod
    def backward(ctx, dout):
        if ctx.acts is None:
            raise RuntimeError('corrector: backward through the fused node a second time (its saved activations were released)')
        meta, per_stage, x, acts, consts = ctx.meta, ctx.per_stage, ctx.x, ctx.acts, ctx.consts
        pptr, cin, cout, pro, eps, bn_training, sync = ctx.arrs
        L = len(meta)
        B, _, H, W = x.shape
        dev = x.device
        dout = dout.contiguous().float()
        need_dx = ctx.needs_input_grad[0]
        gb = [torch.empty((B, st['cin'], H, W), device=dev, dtype=torch.float32) if (j > 0 or need_dx) else None for j, st in enumerate(meta)]
        # one zero-filled workspace: [sums fp64 L*64][dwp L*7*32*32][dbias_acc L*32]
        zb = torch.zeros(L * 128 + L * 7168 + L * 32, device=dev, dtype=torch.float32)
        sums = zb[:L * 128].view(torch.float64)
        dwp = zb[L * 128:L * 128 + L * 7168]
        dba = zb[L * 128 + L * 7168:]
        grads, gptr = [], []
        for st, (k0, k1, bias, gamma, beta) in zip(meta, per_stage):
            g = [torch.empty_like(k0), torch.empty_like(k1), torch.empty_like(bias) if bias is not None else None,
                 torch.empty_like(gamma) if gamma is not None else None, torch.empty_like(beta) if beta is not None else None]
            grads.append(g)
            gptr += [ptr(t) for t in g]
        call('gn_corrector_fused_bwd', ptr(x), _ptr_array([ptr(a) for a in acts]), ptr(dout), _ptr_array([ptr(t) for t in gb]), _ptr_array(pptr),
             _ptr_array(gptr), cin, cout, pro, eps, L, B, H, W, bn_training, ptr(dwp), ptr(dba), ptr(sums), ptr(ctx.stats), ptr(consts), ptr(sync), stream())
        flat = []
        for st, g in zip(meta, grads):
            flat += [g[0], g[1]]
            if st['has_bias']:
                flat.append(g[2])
            if st['bn'] is not None:
                flat += [g[3], g[4]]
        ctx.acts = None
        return (gb[0] if need_dx else None, None) + tuple(flat)


def run_corrector(stages, x, training, sq_transposed=False):
    """x: (B, f_dim, H, W) Visium layout -> (B, n_out, H, W).  ``sq_transposed``: Cartesian nn.Conv2d stages act on the
    TRANSPOSED grid (a hexagonal model hands its corrector the HexagDLy layout, gridnet_models.py:177-185), i.e. with
    their kernels' two spatial axes swapped."""
    meta, params = [], []
    for (hexm, bn, relu) in stages:
        if isinstance(hexm, nn.Conv2d):
            m = dict(kind='sq', ksize=hexm.kernel_size[0], cin=hexm.in_channels, cout=hexm.out_channels, nk=1, has_bias=hexm.bias is not None,
                     bn=None, relu=relu, tr=bool(sq_transposed))
            params.append(hexm.weight)
            if hexm.bias is not None:
                params.append(hexm.bias)
        else:
            k = hexm.hexbase_size
            m = dict(kind='hex', ksize=k, cin=hexm.in_channels, cout=hexm.out_channels, nk=k + 1, has_bias=hexm.bias_tensor is not None, bn=None,
                     relu=relu)
            params.extend(getattr(hexm, 'kernel%d' % i) for i in range(k + 1))
            if hexm.bias_tensor is not None:
                params.append(hexm.bias_tensor)
        if bn is not None:
            bn_train = bool(training and bn.training)
            if bn_train and bn.momentum is None:
                raise NotImplementedError('BatchNorm2d(momentum=None) (cumulative average) is not supported')
            m['bn'] = dict(training=bn_train, momentum=float(bn.momentum if bn.momentum is not None else 0.1), eps=float(bn.eps),
                           running_mean=bn.running_mean, running_var=bn.running_var, sync=bn_train and parallel.sync_bn_active())
            params.extend([bn.weight, bn.bias])
            if bn_train and bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
        meta.append(m)
    if x.shape[1] != meta[0]['cin']:
        raise ValueError('corrector: expected %d input channels, got %d' % (meta[0]['cin'], x.shape[1]))
    if _fused_eligible(meta, x):
        return _CorrectorFusedFn.apply(x, meta, *params)
    return _CorrectorFn.apply(x, meta, *params)
