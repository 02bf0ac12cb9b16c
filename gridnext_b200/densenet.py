"""DenseNet-BC spot classifier (f) with the reference's constructor, module tree and state-dict keys,
executed on hand-written sm_100a kernels.

Mirrors /root/reference/gridnext/densenet.py: ``DenseNet(growth_rate, block_config, compression,
num_init_features, bn_size, drop_rate, num_classes, small_inputs, efficient, classify)`` (:93-95), parameters
``features.conv0 / norm0 / denseblock{i}.denselayer{j}.{norm1,conv1,norm2,conv2} / transition{i}.{norm,conv} /
norm_final`` and ``classifier`` (:102-138), He-normal conv init (:141-150), ``forward`` (:152-159).

Underneath (B200-first, not a translation of the module graph):
  * activations live in NHWC bf16; every dense block owns ONE pre-allocated concat buffer and each layer's conv2
    writes its growth channels in place (no ``torch.cat``, densenet.py:14,75);
  * conv1 (1x1) = tcgen05 GEMM whose A operand gets norm1+ReLU applied in shared memory between TMA and MMA and whose
    epilogue applies norm2+ReLU, so neither activated tensor round-trips HBM un-fused;
  * conv2 (3x3) = padded-position implicit GEMM with the nine taps as descriptor offsets into one TMA-loaded tile;
  * backward = the mirrored kernels with BatchNorm/ReLU backward and BN parameter gradients fused into the
    data-gradient epilogues; weight gradients are split-K tcgen05 GEMMs.
On the grid-wise hot path f is always in eval mode (/root/reference/gridnext/training.py:126): BatchNorm is a per-channel
affine from the running statistics.  In train mode (f pre-training, training.py:11-98) the SAME kernels run with constants
derived from batch statistics (csrc/bn_train.cu): the statistics of a concat channel are computed once, when it is produced,
and shared by every layer that normalises it; the mean/variance terms of the BatchNorm gradient are a per-channel correction
dx -= c0 + c1*x applied once per channel after all of its consumers have accumulated into the gradient buffer.
"""
import math
from collections import OrderedDict

import os
import torch
import torch.nn as nn

from . import _lib
from ._lib import ptr, stream, call
from . import tc

BN_EPS = 1e-5
MAX_SPOTS_RESIDENT = 8192     # spots whose activations are kept for the backward; beyond that f is re-run chunk by chunk


class _DenseLayer(nn.Module):
    def __init__(self, num_input_features, growth_rate, bn_size, drop_rate, efficient=False):
        super().__init__()
        self.add_module('norm1', nn.BatchNorm2d(num_input_features))
        self.add_module('relu1', nn.ReLU(inplace=True))
        self.add_module('conv1', nn.Conv2d(num_input_features, bn_size * growth_rate, kernel_size=1, stride=1, bias=False))
        self.add_module('norm2', nn.BatchNorm2d(bn_size * growth_rate))
        self.add_module('relu2', nn.ReLU(inplace=True))
        self.add_module('conv2', nn.Conv2d(bn_size * growth_rate, growth_rate, kernel_size=3, stride=1, padding=1, bias=False))
        self.drop_rate = drop_rate
        self.efficient = efficient


class _Transition(nn.Sequential):
    def __init__(self, num_input_features, num_output_features):
        super().__init__()
        self.add_module('norm', nn.BatchNorm2d(num_input_features))
        self.add_module('relu', nn.ReLU(inplace=True))
        self.add_module('conv', nn.Conv2d(num_input_features, num_output_features, kernel_size=1, stride=1, bias=False))
        self.add_module('pool', nn.AvgPool2d(kernel_size=2, stride=2))


class _DenseBlock(nn.Module):
    def __init__(self, num_layers, num_input_features, bn_size, growth_rate, drop_rate, efficient=False):
        super().__init__()
        for i in range(num_layers):
            self.add_module('denselayer%d' % (i + 1),
                            _DenseLayer(num_input_features + i * growth_rate, growth_rate, bn_size, drop_rate, efficient))


def _bn_list(net):
    """BatchNorm modules in execution order."""
    out = [net.features.norm0]
    for name, m in net.features.named_children():
        if name.startswith('denseblock'):
            for layer in m.children():
                out += [layer.norm1, layer.norm2]
        elif name.startswith('transition'):
            out.append(m.norm)
    out.append(net.features.norm_final)
    return out


class _Consts:
    """Per-call derived constants: eval-BN scale/shift/invstd/1/gamma for every BN, bf16 weight copies."""

    def __init__(self, net):
        bns = _bn_list(net)
        dev = net.classifier.weight.device
        sizes = [b.num_features for b in bns]
        tot = sum(sizes)
        g = torch.cat([b.weight.detach() for b in bns]).float().contiguous()
        be = torch.cat([b.bias.detach() for b in bns]).float().contiguous()
        mu = torch.cat([b.running_mean for b in bns]).float().contiguous()
        var = torch.cat([b.running_var for b in bns]).float().contiguous()
        buf = torch.empty((4, tot), device=dev, dtype=torch.float32)
        call('gn_bn_eval_consts', ptr(g), ptr(be), ptr(mu), ptr(var), float(bns[0].eps), tot, ptr(buf[0]), ptr(buf[1]), ptr(buf[2]), ptr(buf[3]), stream())
        self.bn = {}
        off = 0
        for b, n in zip(bns, sizes):
            self.bn[id(b)] = dict(sc=buf[0, off:off + n], sh=buf[1, off:off + n], invstd=buf[2, off:off + n], inv_gamma=buf[3, off:off + n],
                                  mean=mu[off:off + n], beta=be[off:off + n], off=off, n=n)
            off += n
        self.bn_total = tot
        self.bns = bns

    def of(self, bn):
        return self.bn[id(bn)]


_JOB_DTYPE = [('src', '<u8'), ('dst', '<i8'), ('start', '<i8'), ('kind', '<i4'), ('a', '<i4'), ('b', '<i4'), ('ld', '<i4')]
_CAST, _TRANSPOSE, _C3PACK0, _C3PACK1, _STEM, _UNPACK_C3, _UNPACK_STEM = range(7)


def _pad8(n):
    return (n + 7) // 8 * 8


class _Plan:
    """Derived-weight and gradient workspaces of one DenseNet on one device, laid out once.

    ``prepare()`` fills ONE flat bf16 buffer with every copy the tensor-core kernels consume (1x1 weights cast and
    transposed, 3x3 weights packed for the forward and the data gradient, the stem window order) in a single launch driven
    by a device-resident job table; ``Grads`` hands out views of ONE flat fp32 accumulator (a single fill per step) and a
    second table turns the packed 3x3 / stem accumulators back into parameter layout in a single launch.  Parameter
    addresses are stable across optimizer steps, so the tables are built once (and rebuilt if a parameter moves)."""

    def __init__(self, net):
        import numpy as np
        assert _lib.load().gn_prep_job_bytes() == np.dtype(_JOB_DTYPE).itemsize
        self.key = _Plan.key_of(net)
        dev = net.classifier.weight.device
        f = net.features
        jobs, self.wviews, self.wshapes = [], {}, {}
        off = [0, 0]          # [bf16 weight elements, job items]

        def add_w(name, param, kind, a, b, ld, rows):
            n = rows * ld
            jobs.append((param.data_ptr(), off[0], off[1], kind, a, b, ld))
            self.wshapes[name] = (off[0], rows, ld)
            off[0] += _pad8(n) + 56      # keep every view 128-byte aligned (TMA needs 16)
            off[0] = (off[0] + 63) // 64 * 64
            off[1] += n

        c0 = f.conv0.out_channels
        add_w(('stem',), f.conv0.weight, _STEM, c0, 0, 32, 7 * c0)
        for name, m in f.named_children():
            if name.startswith('denseblock'):
                for layer in m.children():
                    bott, cin = layer.conv1.out_channels, layer.conv1.in_channels
                    g = layer.conv2.out_channels
                    add_w(('w1', id(layer)), layer.conv1.weight, _CAST, bott, cin, _pad8(cin), bott)              # [bott, cin]
                    add_w(('w1t', id(layer)), layer.conv1.weight, _TRANSPOSE, bott, cin, _pad8(bott), cin)        # [cin, bott]
                    add_w(('wp2', id(layer)), layer.conv2.weight, _C3PACK0, g, bott, _pad8(bott), 9 * g)          # [9*g, bott]
                    add_w(('wp2t', id(layer)), layer.conv2.weight, _C3PACK1, g, bott, _pad8(g), 9 * bott)         # [9*bott, g]
            elif name.startswith('transition'):
                co, ci = m.conv.out_channels, m.conv.in_channels
                add_w(('wt', id(m)), m.conv.weight, _CAST, co, ci, _pad8(ci), co)
                add_w(('wtt', id(m)), m.conv.weight, _TRANSPOSE, co, ci, _pad8(co), ci)
        self.n_w_items = off[1]
        self.wbuf = torch.zeros(off[0], device=dev, dtype=torch.bfloat16)
        self.wjobs = torch.from_numpy(np.array(jobs, dtype=_JOB_DTYPE).view(np.uint8).copy()).to(dev)
        self.n_wjobs = len(jobs)

        # ---- gradient accumulators (fp32): [conv0 dwq][per layer: dw1, dwp2][per transition: dwt][classifier w, b]
        goff = [0]
        self.gslots = {}

        def add_g(name, numel):
            self.gslots[name] = (goff[0], numel)
            goff[0] += (numel + 3) // 4 * 4

        ujobs, uoff = [], [0, 0]
        self.uslots = {}

        def add_u(param, kind, a, b, gname):
            n = param.numel()
            ujobs.append((gname, uoff[0], uoff[1], kind, a, b, 0))
            self.uslots[id(param)] = (uoff[0], tuple(param.shape))
            uoff[0] += (n + 3) // 4 * 4
            uoff[1] += n

        add_g(('stem',), c0 * 224)
        add_u(f.conv0.weight, _UNPACK_STEM, c0, 0, ('stem',))
        for name, m in f.named_children():
            if name.startswith('denseblock'):
                for layer in m.children():
                    bott, cin, g = layer.conv1.out_channels, layer.conv1.in_channels, layer.conv2.out_channels
                    add_g(('dw1', id(layer)), bott * cin)
                    add_g(('dwp2', id(layer)), 9 * bott * g)
                    add_u(layer.conv2.weight, _UNPACK_C3, g, bott, ('dwp2', id(layer)))
            elif name.startswith('transition'):
                add_g(('dwt', id(m)), m.conv.out_channels * m.conv.in_channels)
        J, Cf = net.classifier.out_features, net.classifier.in_features
        add_g(('cls_w',), J * Cf)
        add_g(('cls_b',), J)
        self.g_total = goff[0]
        self.u_total_items = uoff[1]
        self.u_numel = uoff[0]
        self.n_ujobs = len(ujobs)
        self.dev = dev
        # persistent accumulator (stable address: the un-pack table can live on the device, and the step stays graph-capturable)
        self.gflat = torch.zeros(self.g_total, device=dev, dtype=torch.float32)
        rows = [(self.gflat.data_ptr() + 4 * self.gslots[gname][0], dst, start, kind, a, b, ld) for gname, dst, start, kind, a, b, ld in ujobs]
        self.ujobs = torch.from_numpy(np.array(rows, dtype=_JOB_DTYPE).view(np.uint8).copy()).to(dev)

    @staticmethod
    def key_of(net):
        return tuple(p.data_ptr() for p in net.parameters()) + (str(net.classifier.weight.device),)

    @staticmethod
    def of(net):
        plan = getattr(net, '_b200_plan', None)
        if plan is None or plan.key != _Plan.key_of(net):
            plan = _Plan(net)
            object.__setattr__(net, '_b200_plan', plan)
        return plan

    def prepare(self):
        call('gn_prepare_weights', ptr(self.wjobs), self.n_wjobs, self.n_w_items, ptr(self.wbuf), stream())

    def w(self, *name):
        o, rows, ld = self.wshapes[name]
        return self.wbuf[o:o + rows * ld].view(rows, ld)

    def g(self, name, shape):
        o, n = self.gslots[name]
        return self.gflat[o:o + n].view(shape)

    def finish(self):
        """-> (copy of the flat accumulator, flat buffer with the conv2 / conv0 gradients in parameter layout): fresh tensors, so
        the gradients handed to autograd do not alias the persistent workspace."""
        out = torch.empty(self.u_numel, device=self.dev, dtype=torch.float32)
        call('gn_unpack_gradients', ptr(self.ujobs), self.n_ujobs, self.u_total_items, ptr(out), stream())
        return self.gflat.clone(), out


class _Geometry:
    def __init__(self, net, P):
        if P % 4 != 0:
            raise ValueError('DenseNet (B200): patch size %d must be a multiple of 4' % P)
        self.P = P
        self.H0 = P // 2            # conv0 output
        self.H1 = self.H0 // 2      # pool0 output = block 1 resolution
        self.blocks = []            # (block module, transition | None, H, c_in, c_total, layers)
        H = self.H1
        names = [n for n, _ in net.features.named_children()]
        c = net.features.conv0.out_channels
        nb = sum(1 for n in names if n.startswith('denseblock'))
        for i in range(1, nb + 1):
            blk = getattr(net.features, 'denseblock%d' % i)
            layers = list(blk.children())
            g = layers[0].conv2.out_channels
            ct = c + len(layers) * g
            tr = getattr(net.features, 'transition%d' % i) if i != nb else None
            self.blocks.append(dict(block=blk, trans=tr, H=H, c_in=c, c_tot=ct, layers=layers, growth=g, bott=layers[0].conv1.out_channels))
            if tr is not None:
                if H % 2 != 0:
                    raise ValueError('DenseNet (B200): odd feature-map size %d before a transition (patch size %d)' % (H, P))
                c = tr.conv.out_channels
                H //= 2
            else:
                c = ct
        self.c_final = c


def _outside_kernel_envelope(net):
    """None when the tcgen05 kernels cover this configuration, else the reason (string).  The envelope is what the GridNet hot
    path uses (every notebook: growth_rate 32, bn_size 4, 64 stem channels, small_inputs=False, drop_rate 0): TMA needs 16-byte
    aligned channel slices (multiples of 8 bf16), UMMA needs N a multiple of 16."""
    if net.small_inputs:
        return 'small_inputs=True (3x3 stem without norm0 / pool0)'
    f = net.features
    if f.conv0.out_channels % 8 or f.conv0.out_channels > 128:
        return 'num_init_features must be a multiple of 8, at most 128'
    for name, m in f.named_children():
        if name.startswith('denseblock'):
            for l in m.children():
                g, b = l.conv2.out_channels, l.conv1.out_channels
                if g % 8 or g > 48 or b % 16 or b > 128:
                    return 'growth_rate must be a multiple of 8 (<= 48) and bn_size*growth_rate a multiple of 16 (<= 128)'
                if l.drop_rate > 0 and l.training:
                    return 'dropout in training mode'
        if name.startswith('transition') and m.conv.out_channels % 8:
            return 'transition widths must be multiples of 8'
    return None


def _check_supported(net):
    modes = set(bool(b.training) for b in _bn_list(net))
    if len(modes) != 1:
        raise NotImplementedError('DenseNet (B200): BatchNorm layers must be all in train mode or all in eval mode')
    if any(not b.track_running_stats for b in _bn_list(net)):
        raise NotImplementedError('DenseNet (B200): BatchNorm without running statistics is not supported')


_WARNED = set()


def _reference_graph_forward(net, x):
    """The reference's module graph (densenet.py:21-44,57-75,152-159) evaluated with PyTorch's own CUDA operators: used ONLY for
    configurations outside the kernels' envelope (the constructor's defaults -- growth_rate 12, small_inputs=True -- and train-mode
    dropout), none of which the GridNet notebooks use.  Loud by design: a one-time warning names the reason."""
    import torch.nn.functional as F
    f = net.features
    h = f.conv0(x.float())
    if hasattr(f, 'norm0'):
        h = f.pool0(f.relu0(f.norm0(h)))
    for name, m in f.named_children():
        if name.startswith('denseblock'):
            feats = [h]
            for layer in m.children():
                t = torch.cat(feats, 1)
                t = layer.conv1(F.relu(layer.norm1(t)))
                t = layer.conv2(F.relu(layer.norm2(t)))
                if layer.drop_rate > 0:
                    t = F.dropout(t, p=layer.drop_rate, training=layer.training)
                feats.append(t)
            h = torch.cat(feats, 1)
        elif name.startswith('transition'):
            h = m.pool(m.conv(F.relu(m.norm(h))))
    h = F.relu(f.norm_final(h))
    h = F.adaptive_avg_pool2d(h, (1, 1)).view(h.size(0), -1)
    return net.classifier(h) if net.classify else h


def _forward_chunk(net, geo, cst, plan, x, save):
    """x: (n, 3, P, P) fp32|bf16 CUDA.  Returns (out, saved | None)."""
    n, P, dev = x.shape[0], geo.P, x.device
    f = net.features
    bf = torch.bfloat16
    # ---- stem: conv0 (7x7/2) read straight from the NHWC4 patch by the tensor core (no im2col buffer), norm0+ReLU in the
    # epilogue, then 3x3/2 max-pool into block 1's buffer
    xq = tc.stem_pack_input(x)
    b0 = cst.of(f.norm0)
    c0 = f.conv0.out_channels
    act0 = tc.stem_conv_fwd(xq, plan.w('stem').view(7, c0, 32), scale=b0['sc'], shift=b0['sh'], relu=True)
    saved = dict(xq=xq, act0=act0, blocks=[]) if save else None
    blk0 = geo.blocks[0]
    M = n * blk0['H'] * blk0['H']
    C = torch.empty((M, blk0['c_tot']), device=dev, dtype=bf)
    idx0 = torch.empty((M, c0), device=dev, dtype=torch.uint8)
    call('gn_maxpool3s2_fwd', ptr(act0), c0, n, geo.H0, geo.H0, c0, ptr(C), blk0['c_tot'], ptr(idx0), stream())
    if save:
        saved['idx0'] = idx0
    # ---- dense blocks
    for bi, blk in enumerate(geo.blocks):
        H, g, bott = blk['H'], blk['growth'], blk['bott']
        M = n * H * H
        a2s = []
        a2 = None
        cin = blk['c_in']
        for layer in blk['layers']:
            k1, k2 = cst.of(layer.norm1), cst.of(layer.norm2)
            if save or a2 is None:
                a2 = torch.empty((M, bott), device=dev, dtype=bf)
            tc.gemm_bf16(C[:, :cin], plan.w('w1', id(layer))[:, :cin], out=a2, scale=k2['sc'], shift=k2['sh'], relu=True,
                         xf_scale=k1['sc'], xf_shift=k1['sh'])
            tc.conv3x3_bf16(a2, n, H, H, bott, plan.w('wp2', id(layer)), g, C[:, cin:cin + g])
            if save:
                a2s.append(a2)
            cin += g
        rec = dict(C=C, a2=a2s)
        if blk['trans'] is not None:
            tr = blk['trans']
            kt = cst.of(tr.norm)
            ct = blk['c_tot']
            pooled = torch.empty((M // 4, ct), device=dev, dtype=bf)
            call('gn_bnrelu_avgpool2_fwd', ptr(C), ct, n, H, H, ct, ptr(kt['sc']), ptr(kt['sh']), ptr(pooled), ct, stream())
            nxt = geo.blocks[bi + 1]
            Cn = torch.empty((M // 4, nxt['c_tot']), device=dev, dtype=bf)
            tc.gemm_bf16(pooled, plan.w('wt', id(tr))[:, :ct], out=Cn[:, :nxt['c_in']])
            rec['pooled'] = pooled
            if save:
                saved['blocks'].append(rec)
            C = Cn
        elif save:
            saved['blocks'].append(rec)
    # ---- head
    kf = cst.of(f.norm_final)
    last = geo.blocks[-1]
    feat = torch.empty((n, geo.c_final), device=dev, dtype=torch.float32)
    call('gn_bnrelu_gap_fwd', ptr(C), last['c_tot'], n, last['H'] * last['H'], geo.c_final, ptr(kf['sc']), ptr(kf['sh']), ptr(feat), geo.c_final, stream())
    if net.classify:
        J = net.classifier.out_features
        out = torch.empty((n, J), device=dev, dtype=torch.float32)
        call('gn_linear_small_fwd', ptr(feat), geo.c_final, ptr(net.classifier.weight.detach().float().contiguous()),
             ptr(net.classifier.bias.detach().float().contiguous()), n, geo.c_final, J, ptr(out), stream())
    else:
        out = feat
    if save:
        saved['feat'] = feat
        saved['n'] = n
    return out, saved


class _Grads:
    """Gradient accumulators of one backward pass: BatchNorm column sums + the plan's flat fp32 workspace (zeroed here)."""

    def __init__(self, net, cst, plan):
        dev = net.classifier.weight.device
        self.bn_colsum = torch.zeros((2, cst.bn_total), device=dev, dtype=torch.float32)   # row 0: d beta, row 1: d gamma
        self.plan = plan
        plan.gflat.zero_()


def _saved_bytes_per_spot(geo, c0):
    """Bytes of activations _forward_chunk keeps per spot for the backward (bf16 unless noted): packed patch, conv0 output,
    max-pool arg-max bytes, every block's concat buffer, one bottleneck tensor per dense layer, the transition inputs."""
    n = geo.P * geo.P * 8 + geo.H0 * geo.H0 * c0 * 2 + geo.H1 * geo.H1 * c0
    for blk in geo.blocks:
        hw = blk['H'] * blk['H']
        n += hw * blk['c_tot'] * 2 + len(blk['layers']) * hw * blk['bott'] * 2
        if blk['trans'] is not None:
            n += (hw // 4) * blk['c_tot'] * 2
    return n


def _conv1_backward(dz, w1t, dx, dw, bn):
    """Backward of a dense layer's norm1 -> relu1 -> conv1 (densenet.py:12-18,26-27): data gradient with the BatchNorm/ReLU backward
    epilogue and the weight gradient.  One kernel when the views allow it (bottleneck width 128), else the two GEMMs."""
    if tc.conv1x1_bwd_fusable(dz, w1t, dx, bn['ref']):
        tc.conv1x1_bwd_bf16(dz, w1t, dx, bn, dw)
    else:
        tc.gemm_tn_bf16(dz, bn['ref'], dw, bn['sc'], bn['sh'])
        tc.gemm_bf16(dz, w1t, out=dx, bn=bn)


def _backward_chunk(net, geo, cst, plan, saved, dout, grads):
    n, dev, bf = saved['n'], dout.device, torch.bfloat16
    f = net.features
    last = geo.blocks[-1]
    cs = grads.bn_colsum
    tot = cst.bn_total

    def colsum_of(k):
        return cs[:, k['off']:k['off'] + k['n']]

    # ---- head
    Cf = geo.c_final
    if net.classify:
        J = net.classifier.out_features
        dfeat = torch.empty((n, Cf), device=dev, dtype=torch.float32)
        dw = plan.g(('cls_w',), (J, Cf))
        db = plan.g(('cls_b',), (J,))
        call('gn_linear_small_bwd', ptr(dout), ptr(saved['feat']), Cf, ptr(net.classifier.weight.detach().float().contiguous()), n, Cf, J,
             ptr(dfeat), Cf, ptr(dw), ptr(db), stream())
    else:
        dfeat = dout
    kf = cst.of(f.norm_final)
    C = saved['blocks'][-1]['C']
    M = n * last['H'] * last['H']
    dC = torch.empty((M, last['c_tot']), device=dev, dtype=bf)
    call('gn_pool_bnrelu_bwd', ptr(dfeat), Cf, 1, ptr(C), last['c_tot'], n, last['H'], last['H'], Cf, ptr(kf['sc']), ptr(kf['sh']),
         ptr(kf['mean']), ptr(kf['invstd']), ptr(dC), last['c_tot'], ptr(colsum_of(kf)), tot, stream())
    # ---- blocks, last to first
    for bi in range(len(geo.blocks) - 1, -1, -1):
        blk, rec = geo.blocks[bi], saved['blocks'][bi]
        H, g, bott = blk['H'], blk['growth'], blk['bott']
        C = rec['C']
        M = n * H * H
        dz = torch.empty((M, bott), device=dev, dtype=bf)
        cin = blk['c_tot']
        for li in range(len(blk['layers']) - 1, -1, -1):
            layer = blk['layers'][li]
            cin -= g
            k1, k2 = cst.of(layer.norm1), cst.of(layer.norm2)
            a2 = rec['a2'][li]
            dY = dC[:, cin:cin + g]
            tc.conv3x3_wgrad_into(a2, dY, n, H, H, bott, g, plan.g(('dwp2', id(layer)), (9, bott, g)))
            tc.conv3x3_bf16(dY, n, H, H, g, plan.w('wp2t', id(layer)), bott, dz,
                            bn=dict(ref=a2, ref_is_raw=False, sc=k2['sc'], sh=None, p0=k2['beta'], p1=k2['inv_gamma'], colsum=colsum_of(k2)))
            _conv1_backward(dz, plan.w('w1t', id(layer))[:, :bott], dC[:, :cin], plan.g(('dw1', id(layer)), (bott, cin)),
                            dict(ref=C[:, :cin], ref_is_raw=True, sc=k1['sc'], sh=k1['sh'], p0=k1['mean'], p1=k1['invstd'],
                                 colsum=colsum_of(k1), rmw=True))
        if bi > 0:
            prev, prec = geo.blocks[bi - 1], saved['blocks'][bi - 1]
            tr = prev['trans']
            kt = cst.of(tr.norm)
            ctp, c_in = prev['c_tot'], blk['c_in']
            d_out = dC[:, :c_in]
            tc.gemm_tn_bf16(d_out, prec['pooled'], plan.g(('dwt', id(tr)), (c_in, ctp)))
            dP = tc.gemm_bf16(d_out, plan.w('wtt', id(tr))[:, :c_in])
            Hp = prev['H']
            dCp = torch.empty((n * Hp * Hp, ctp), device=dev, dtype=bf)
            call('gn_pool_bnrelu_bwd', ptr(dP), ctp, 0, ptr(prec['C']), ctp, n, Hp, Hp, ctp, ptr(kt['sc']), ptr(kt['sh']), ptr(kt['mean']),
                 ptr(kt['invstd']), ptr(dCp), ctp, ptr(colsum_of(kt)), tot, stream())
            dC = dCp
        else:
            c0 = f.conv0.out_channels
            k0 = cst.of(f.norm0)
            M0 = n * geo.H0 * geo.H0
            dz0 = torch.empty((M0, c0), device=dev, dtype=bf)
            call('gn_maxpool3s2_bnrelu_bwd', ptr(dC), blk['c_tot'], ptr(saved['idx0']), ptr(saved['act0']), c0, n, geo.H0, geo.H0, c0,
                 ptr(k0['sc']), ptr(k0['beta']), ptr(k0['inv_gamma']), ptr(dz0), c0, ptr(colsum_of(k0)), tot, stream())
            tc.stem_conv_wgrad_into(saved['xq'], dz0, c0, plan.g(('stem',), (c0, 224)))


class _TrainConsts:
    """Train-mode counterpart of _Consts: the same per-BN views (sc, sh, mean, invstd, beta, inv_gamma), but sc / sh / mean /
    invstd are filled layer by layer from batch statistics (``fill``) as the forward pass produces the data."""

    def __init__(self, net):
        bns = _bn_list(net)
        dev = net.classifier.weight.device
        sizes = [b.num_features for b in bns]
        tot = sum(sizes)
        g = torch.cat([b.weight.detach() for b in bns]).float()
        be = torch.cat([b.bias.detach() for b in bns]).float().contiguous()
        ig = torch.where(g != 0, 1.0 / g, torch.zeros_like(g)).contiguous()
        buf = torch.empty((4, tot), device=dev, dtype=torch.float32)
        self.bn = {}
        off = 0
        for b, n in zip(bns, sizes):
            self.bn[id(b)] = dict(sc=buf[0, off:off + n], sh=buf[1, off:off + n], mean=buf[2, off:off + n], invstd=buf[3, off:off + n],
                                  beta=be[off:off + n], inv_gamma=ig[off:off + n], off=off, n=n)
            off += n
        self.bn_total = tot
        self.bns = bns

    def of(self, bn):
        return self.bn[id(bn)]

    def fill(self, bn, ssum, ssq, M):
        """Batch statistics (fp64 sums over M rows) -> this BN's constants; running statistics updated as nn.BatchNorm2d does."""
        if M < 2:
            raise ValueError('Expected more than 1 value per channel when training, got %d' % M)
        k = self.of(bn)
        mom = bn.momentum if bn.momentum is not None else 1.0 / (int(bn.num_batches_tracked.item()) + 1)
        call('gn_bn_train_coeffs', ptr(ssum), ptr(ssq), M, ptr(bn.weight.detach()), ptr(bn.bias.detach()), float(bn.eps), float(mom),
             ptr(bn.running_mean), ptr(bn.running_var), ptr(k['sc']), ptr(k['sh']), ptr(k['mean']), ptr(k['invstd']), k['n'], stream())
        return k

    def finish(self):
        torch._foreach_add_([b.num_batches_tracked for b in self.bns], 1)


def _colstats(x2d, C, ssum, ssq):
    call('gn_colstats_bf16', ptr(x2d), x2d.stride(0), x2d.shape[0], C, ptr(ssum), ptr(ssq), stream())


def _affine_relu(x2d, y2d, C, k):
    call('gn_affine_relu_bf16', ptr(x2d), x2d.stride(0), ptr(y2d), y2d.stride(0), x2d.shape[0], C, ptr(k['sc']), ptr(k['sh']), 1, stream())


def _fix_coeffs(colsum2, k, M, accumulate, F, C):
    """F[0] (c0), F[1] (c1) (+)= the BatchNorm mean/variance gradient terms of one BN from its column sums (d_beta, d_gamma)."""
    call('gn_bn_train_fix_coeffs', ptr(colsum2[0]), ptr(colsum2[1]), ptr(k['sc']), ptr(k['invstd']), ptr(k['mean']), M, 1 if accumulate else 0,
         ptr(F[0]), ptr(F[1]), C, stream())


def _fix(dx2d, x2d, C, F0, F1):
    call('gn_bn_train_fix_bf16', ptr(dx2d), dx2d.stride(0), ptr(x2d), x2d.stride(0), dx2d.shape[0], C, ptr(F0), ptr(F1), stream())


def _forward_train(net, geo, cst, plan, x, save):
    """Train-mode forward of one batch (batch statistics span the whole batch: no chunking).  Returns (out, saved | None)."""
    n, dev = x.shape[0], x.device
    f = net.features
    bf = torch.bfloat16
    # fp64 (sum, sum of squares) slots: conv0 | per block: the concat channels, then one bottleneck slot per layer
    slots = f.conv0.out_channels + sum(b['c_tot'] + len(b['layers']) * b['bott'] for b in geo.blocks)
    st = torch.zeros((2, slots), device=dev, dtype=torch.float64)
    so = [0]

    def take(nch):
        v = st[:, so[0]:so[0] + nch]
        so[0] += nch
        return v

    # ---- stem: raw conv0, statistics, norm0 + ReLU as a pass, max-pool
    xq = tc.stem_pack_input(x)
    c0 = f.conv0.out_channels
    z0 = tc.stem_conv_fwd(xq, plan.w('stem').view(7, c0, 32))
    M0 = n * geo.H0 * geo.H0
    s0 = take(c0)
    _colstats(z0, c0, s0[0], s0[1])
    k0 = cst.fill(f.norm0, s0[0], s0[1], M0)
    act0 = torch.empty_like(z0)
    _affine_relu(z0, act0, c0, k0)
    saved = dict(xq=xq, z0=z0, act0=act0, blocks=[]) if save else None
    blk0 = geo.blocks[0]
    M = n * blk0['H'] * blk0['H']
    C = torch.empty((M, blk0['c_tot']), device=dev, dtype=bf)
    idx0 = torch.empty((M, c0), device=dev, dtype=torch.uint8)
    call('gn_maxpool3s2_fwd', ptr(act0), c0, n, geo.H0, geo.H0, c0, ptr(C), blk0['c_tot'], ptr(idx0), stream())
    if save:
        saved['idx0'] = idx0
    stats = None
    for bi, blk in enumerate(geo.blocks):
        H, g, bott = blk['H'], blk['growth'], blk['bott']
        M = n * H * H
        cin = blk['c_in']
        stats = take(blk['c_tot'])
        _colstats(C, cin, stats[0], stats[1])
        a2 = torch.empty((M, bott), device=dev, dtype=bf)
        zs = []
        z = None
        for layer in blk['layers']:
            k1 = cst.fill(layer.norm1, stats[0], stats[1], M)
            if save or z is None:
                z = torch.empty((M, bott), device=dev, dtype=bf)
            tc.gemm_bf16(C[:, :cin], plan.w('w1', id(layer))[:, :cin], out=z, xf_scale=k1['sc'], xf_shift=k1['sh'])
            s2 = take(bott)
            _colstats(z, bott, s2[0], s2[1])
            k2 = cst.fill(layer.norm2, s2[0], s2[1], M)
            _affine_relu(z, a2, bott, k2)
            tc.conv3x3_bf16(a2, n, H, H, bott, plan.w('wp2', id(layer)), g, C[:, cin:cin + g])
            _colstats(C[:, cin:cin + g], g, stats[0, cin:], stats[1, cin:])
            if save:
                zs.append(z)
            cin += g
        rec = dict(C=C, z=zs)
        if blk['trans'] is not None:
            tr = blk['trans']
            ct = blk['c_tot']
            kt = cst.fill(tr.norm, stats[0], stats[1], M)
            pooled = torch.empty((M // 4, ct), device=dev, dtype=bf)
            call('gn_bnrelu_avgpool2_fwd', ptr(C), ct, n, H, H, ct, ptr(kt['sc']), ptr(kt['sh']), ptr(pooled), ct, stream())
            nxt = geo.blocks[bi + 1]
            Cn = torch.empty((M // 4, nxt['c_tot']), device=dev, dtype=bf)
            tc.gemm_bf16(pooled, plan.w('wt', id(tr))[:, :ct], out=Cn[:, :nxt['c_in']])
            rec['pooled'] = pooled
            if save:
                saved['blocks'].append(rec)
            C = Cn
        elif save:
            saved['blocks'].append(rec)
    last = geo.blocks[-1]
    kf = cst.fill(f.norm_final, stats[0], stats[1], n * last['H'] * last['H'])
    feat = torch.empty((n, geo.c_final), device=dev, dtype=torch.float32)
    call('gn_bnrelu_gap_fwd', ptr(C), last['c_tot'], n, last['H'] * last['H'], geo.c_final, ptr(kf['sc']), ptr(kf['sh']), ptr(feat), geo.c_final, stream())
    if net.classify:
        J = net.classifier.out_features
        out = torch.empty((n, J), device=dev, dtype=torch.float32)
        call('gn_linear_small_fwd', ptr(feat), geo.c_final, ptr(net.classifier.weight.detach().float().contiguous()),
             ptr(net.classifier.bias.detach().float().contiguous()), n, geo.c_final, J, ptr(out), stream())
    else:
        out = feat
    if save:
        saved['feat'] = feat
        saved['n'] = n
    cst.finish()
    return out, saved


def _backward_train(net, geo, cst, plan, saved, dout, grads):
    """Train-mode backward: the eval-mode kernel sequence plus the per-channel mean/variance corrections (module docstring)."""
    n, dev, bf = saved['n'], dout.device, torch.bfloat16
    f = net.features
    last = geo.blocks[-1]
    cs = grads.bn_colsum
    tot = cst.bn_total

    def colsum_of(k):
        return cs[:, k['off']:k['off'] + k['n']]

    Cf = geo.c_final
    if net.classify:
        J = net.classifier.out_features
        dfeat = torch.empty((n, Cf), device=dev, dtype=torch.float32)
        call('gn_linear_small_bwd', ptr(dout), ptr(saved['feat']), Cf, ptr(net.classifier.weight.detach().float().contiguous()), n, Cf, J,
             ptr(dfeat), Cf, ptr(plan.g(('cls_w',), (J, Cf))), ptr(plan.g(('cls_b',), (J,))), stream())
    else:
        dfeat = dout
    kf = cst.of(f.norm_final)
    C = saved['blocks'][-1]['C']
    M = n * last['H'] * last['H']
    dC = torch.empty((M, last['c_tot']), device=dev, dtype=bf)
    call('gn_pool_bnrelu_bwd', ptr(dfeat), Cf, 1, ptr(C), last['c_tot'], n, last['H'], last['H'], Cf, ptr(kf['sc']), ptr(kf['sh']),
         ptr(kf['mean']), ptr(kf['invstd']), ptr(dC), last['c_tot'], ptr(colsum_of(kf)), tot, stream())
    F = torch.zeros((2, last['c_tot']), device=dev, dtype=torch.float32)      # (c0, c1) of the block's concat channels
    _fix_coeffs(colsum_of(kf), kf, M, True, F, Cf)
    for bi in range(len(geo.blocks) - 1, -1, -1):
        blk, rec = geo.blocks[bi], saved['blocks'][bi]
        H, g, bott = blk['H'], blk['growth'], blk['bott']
        C = rec['C']
        M = n * H * H
        dz = torch.empty((M, bott), device=dev, dtype=bf)
        a2 = torch.empty((M, bott), device=dev, dtype=bf)
        Fz = torch.empty((2, bott), device=dev, dtype=torch.float32)
        cin = blk['c_tot']
        for li in range(len(blk['layers']) - 1, -1, -1):
            layer = blk['layers'][li]
            cin -= g
            k1, k2 = cst.of(layer.norm1), cst.of(layer.norm2)
            z = rec['z'][li]
            dY = dC[:, cin:cin + g]
            _fix(dY, C[:, cin:cin + g], g, F[0, cin:], F[1, cin:])          # every consumer of these channels has accumulated
            _affine_relu(z, a2, bott, k2)
            tc.conv3x3_wgrad_into(a2, dY, n, H, H, bott, g, plan.g(('dwp2', id(layer)), (9, bott, g)))
            tc.conv3x3_bf16(dY, n, H, H, g, plan.w('wp2t', id(layer)), bott, dz,
                            bn=dict(ref=z, ref_is_raw=True, sc=k2['sc'], sh=k2['sh'], p0=k2['mean'], p1=k2['invstd'], colsum=colsum_of(k2)))
            _fix_coeffs(colsum_of(k2), k2, M, False, Fz, bott)
            _fix(dz, z, bott, Fz[0], Fz[1])
            _conv1_backward(dz, plan.w('w1t', id(layer))[:, :bott], dC[:, :cin], plan.g(('dw1', id(layer)), (bott, cin)),
                            dict(ref=C[:, :cin], ref_is_raw=True, sc=k1['sc'], sh=k1['sh'], p0=k1['mean'], p1=k1['invstd'],
                                 colsum=colsum_of(k1), rmw=True))
            _fix_coeffs(colsum_of(k1), k1, M, True, F, cin)
        c_in = blk['c_in']
        _fix(dC[:, :c_in], C[:, :c_in], c_in, F[0], F[1])
        if bi > 0:
            prev, prec = geo.blocks[bi - 1], saved['blocks'][bi - 1]
            tr = prev['trans']
            kt = cst.of(tr.norm)
            ctp = prev['c_tot']
            d_out = dC[:, :c_in]
            tc.gemm_tn_bf16(d_out, prec['pooled'], plan.g(('dwt', id(tr)), (c_in, ctp)))
            dP = tc.gemm_bf16(d_out, plan.w('wtt', id(tr))[:, :c_in])
            Hp = prev['H']
            Mp = n * Hp * Hp
            dCp = torch.empty((Mp, ctp), device=dev, dtype=bf)
            call('gn_pool_bnrelu_bwd', ptr(dP), ctp, 0, ptr(prec['C']), ctp, n, Hp, Hp, ctp, ptr(kt['sc']), ptr(kt['sh']), ptr(kt['mean']),
                 ptr(kt['invstd']), ptr(dCp), ctp, ptr(colsum_of(kt)), tot, stream())
            F = torch.empty((2, ctp), device=dev, dtype=torch.float32)
            _fix_coeffs(colsum_of(kt), kt, Mp, False, F, ctp)
            dC = dCp
        else:
            c0 = f.conv0.out_channels
            k0 = cst.of(f.norm0)
            M0 = n * geo.H0 * geo.H0
            dz0 = torch.empty((M0, c0), device=dev, dtype=bf)
            call('gn_maxpool3s2_bnrelu_bwd', ptr(dC), blk['c_tot'], ptr(saved['idx0']), ptr(saved['act0']), c0, n, geo.H0, geo.H0, c0,
                 ptr(k0['sc']), ptr(k0['beta']), ptr(k0['inv_gamma']), ptr(dz0), c0, ptr(colsum_of(k0)), tot, stream())
            F0 = torch.empty((2, c0), device=dev, dtype=torch.float32)
            _fix_coeffs(colsum_of(k0), k0, M0, False, F0, c0)
            _fix(dz0, saved['z0'], c0, F0[0], F0[1])
            tc.stem_conv_wgrad_into(saved['xq'], dz0, c0, plan.g(('stem',), (c0, 224)))


class _DenseNetFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, net, *params):
        _lib.require_cuda(x)
        _check_supported(net)
        if x.dim() != 4 or x.shape[1] != 3 or x.shape[2] != x.shape[3]:
            raise ValueError('DenseNet (B200): expected (N, 3, P, P) patches, got %s' % (tuple(x.shape),))
        if x.dtype not in (torch.float32, torch.bfloat16):
            x = x.float()
        x = x.contiguous()
        geo = _Geometry(net, int(x.shape[2]))
        train_bn = bool(net.features.norm_final.training)
        cst = _TrainConsts(net) if train_bn else _Consts(net)
        plan = _Plan.of(net)
        plan.prepare()
        N = x.shape[0]
        need_grad = any(ctx.needs_input_grad[2:])
        ctx.net, ctx.geo, ctx.cst, ctx.plan, ctx.train_bn = net, geo, cst, plan, train_bn
        ctx.plist = list(net.parameters())
        assert len(ctx.plist) == len(params)
        if train_bn:
            if N > MAX_SPOTS_RESIDENT:
                raise ValueError('DenseNet (B200): a train-mode batch is limited to %d spots (batch statistics span the whole batch)' % MAX_SPOTS_RESIDENT)
            out, saved = _forward_train(net, geo, cst, plan, x, need_grad)
            ctx.saved, ctx.x = saved, None
        elif N <= MAX_SPOTS_RESIDENT:
            out, saved = _forward_chunk(net, geo, cst, plan, x, need_grad)
            ctx.saved, ctx.x = saved, None
        else:
            # Chunk-wise forward.  A chunk's activations are KEPT for the backward while device memory allows (180 GB hold four Visium
            # arrays of DenseNet-121 @128 px: train_gridwise at batch 4 then runs without any recomputation); the other chunks are
            # re-run in the backward.  Decided per chunk from the allocator's figures; not under CUDA-graph capture (the graph's pool
            # would pin the memory for good).
            outs, kept = [], {}
            may_keep = need_grad and not torch.cuda.is_current_stream_capturing() and os.environ.get('GRIDNEXT_B200_KEEP_CHUNKS', '1') != '0'
            per_spot = _saved_bytes_per_spot(geo, net.features.conv0.out_channels)
            for i in range(0, N, MAX_SPOTS_RESIDENT):
                xi = x[i:i + MAX_SPOTS_RESIDENT]
                keep = False
                if may_keep:
                    free, _ = torch.cuda.mem_get_info(x.device)
                    avail = free + torch.cuda.memory_reserved(x.device) - torch.cuda.memory_allocated(x.device)
                    # what stays + the transients of one chunk's backward (~0.3x of its saved bytes) + slack for fragmentation
                    keep = avail - per_spot * xi.shape[0] > 0.4 * per_spot * MAX_SPOTS_RESIDENT + (8 << 30)
                if keep:
                    try:
                        o, kept[i] = _forward_chunk(net, geo, cst, plan, xi, True)
                    except torch.cuda.OutOfMemoryError:
                        # the allocator's figures were too optimistic (e.g. cached blocks of a CUDA graph's private pool are counted as
                        # reserved but cannot serve this stream): give everything kept so far back and recompute all chunks in the backward
                        may_keep = False
                        kept.clear()
                        torch.cuda.empty_cache()
                        o = _forward_chunk(net, geo, cst, plan, xi, False)[0]
                else:
                    o = _forward_chunk(net, geo, cst, plan, xi, False)[0]
                outs.append(o)
            out = torch.cat(outs, 0)
            ctx.saved, ctx.kept, ctx.x = None, kept, (x if need_grad else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        net, geo, cst, plan = ctx.net, ctx.geo, ctx.cst, ctx.plan
        if ctx.saved is None and getattr(ctx, 'x', None) is None:
            raise RuntimeError('DenseNet (B200): backward a second time (saved activations were released; use a fresh forward)')
        dout = dout.contiguous().float()
        grads = _Grads(net, cst, plan)
        if ctx.train_bn:
            _backward_train(net, geo, cst, plan, ctx.saved, dout, grads)
            ctx.saved = None
        elif ctx.saved is not None:
            _backward_chunk(net, geo, cst, plan, ctx.saved, dout, grads)
            ctx.saved = None
        else:
            x, kept = ctx.x, ctx.kept
            for i in range(0, x.shape[0], MAX_SPOTS_RESIDENT):
                saved = kept.pop(i, None)
                if saved is None:
                    _, saved = _forward_chunk(net, geo, cst, plan, x[i:i + MAX_SPOTS_RESIDENT], True)
                _backward_chunk(net, geo, cst, plan, saved, dout[i:i + MAX_SPOTS_RESIDENT].contiguous(), grads)
                del saved
            ctx.x = ctx.kept = None
        # map accumulated gradients onto the parameter list: views of two fresh flat buffers
        gret, gunp = plan.finish()
        bn_grads = grads.bn_colsum
        by_param = {}
        for b in cst.bns:
            k = cst.of(b)
            by_param[id(b.weight)] = bn_grads[1, k['off']:k['off'] + k['n']]
            by_param[id(b.bias)] = bn_grads[0, k['off']:k['off'] + k['n']]
        f = net.features
        for nm, m in f.named_children():
            if nm.startswith('denseblock'):
                for layer in m.children():
                    o, nel = plan.gslots[('dw1', id(layer))]
                    by_param[id(layer.conv1.weight)] = gret[o:o + nel]
            elif nm.startswith('transition'):
                o, nel = plan.gslots[('dwt', id(m))]
                by_param[id(m.conv.weight)] = gret[o:o + nel]
        o, nel = plan.gslots[('cls_w',)]
        by_param[id(net.classifier.weight)] = gret[o:o + nel]
        o, nel = plan.gslots[('cls_b',)]
        by_param[id(net.classifier.bias)] = gret[o:o + nel]
        for pid, (o, shape) in plan.uslots.items():
            nel = 1
            for d in shape:
                nel *= d
            by_param[pid] = gunp[o:o + nel]
        out = []
        for p in ctx.plist:
            gp = by_param.get(id(p))
            out.append(None if gp is None else gp.view(p.shape))
        return (None, None) + tuple(out)


class DenseNet(nn.Module):
    r"""Densenet-BC (`"Densely Connected Convolutional Networks" <https://arxiv.org/pdf/1608.06993.pdf>`).

    Args: growth_rate, block_config, compression, num_init_features, bn_size, drop_rate, num_classes,
    small_inputs (True: 3x3 stem for 32x32 images; False: 7x7/2 stem + max-pool), efficient (accepted for API
    compatibility; activations are managed by the kernels), classify (return logits vs. pooled features).
    """

    def __init__(self, growth_rate=12, block_config=(16, 16, 16), compression=0.5,
                 num_init_features=24, bn_size=4, drop_rate=0,
                 num_classes=10, small_inputs=True, efficient=False, classify=True):
        super(DenseNet, self).__init__()
        assert 0 < compression <= 1, 'compression of densenet should be between 0 and 1'
        self.small_inputs = small_inputs
        if small_inputs:
            self.features = nn.Sequential(OrderedDict([
                ('conv0', nn.Conv2d(3, num_init_features, kernel_size=3, stride=1, padding=1, bias=False))]))
        else:
            self.features = nn.Sequential(OrderedDict([
                ('conv0', nn.Conv2d(3, num_init_features, kernel_size=7, stride=2, padding=3, bias=False))]))
            self.features.add_module('norm0', nn.BatchNorm2d(num_init_features))
            self.features.add_module('relu0', nn.ReLU(inplace=True))
            self.features.add_module('pool0', nn.MaxPool2d(kernel_size=3, stride=2, padding=1, ceil_mode=False))
        num_features = num_init_features
        for i, num_layers in enumerate(block_config):
            self.features.add_module('denseblock%d' % (i + 1),
                                     _DenseBlock(num_layers, num_features, bn_size, growth_rate, drop_rate, efficient))
            num_features = num_features + num_layers * growth_rate
            if i != len(block_config) - 1:
                self.features.add_module('transition%d' % (i + 1), _Transition(num_features, int(num_features * compression)))
                num_features = int(num_features * compression)
        self.features.add_module('norm_final', nn.BatchNorm2d(num_features))
        self.classify = classify
        self.classifier = nn.Linear(num_features, num_classes)
        for name, param in self.named_parameters():
            if 'conv' in name and 'weight' in name:
                n = param.size(0) * param.size(2) * param.size(3)
                param.data.normal_().mul_(math.sqrt(2. / n))
            elif 'norm' in name and 'weight' in name:
                param.data.fill_(1)
            elif 'norm' in name and 'bias' in name:
                param.data.fill_(0)
            elif 'classifier' in name and 'bias' in name:
                param.data.fill_(0)

    def forward(self, x):
        why = _outside_kernel_envelope(self)
        if why is not None:
            _lib.require_cuda(x)
            if why not in _WARNED:
                import warnings
                _WARNED.add(why)
                warnings.warn('gridnext_b200 DenseNet: %s is outside the tcgen05 kernels\' envelope; this configuration runs the reference '
                              'module graph with PyTorch\'s CUDA operators (correct, not the optimised path)' % why, stacklevel=2)
            return _reference_graph_forward(self, x)
        return _DenseNetFn.apply(x, self, *self.parameters())
