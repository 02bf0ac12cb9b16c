"""Functional CPU restatement of the GridNet hot path (test infrastructure only).

Everything here works on a flat ``state_dict``-style mapping ``{key: tensor}`` with the
reference's key names, so the same tensors can be loaded into the reference modules (in
``oracle/make_golden.py``) and into the product modules (in ``tests/``).

Follows (paths relative to /root/reference):
  gridnext/gridnet_models.py:81-109   patch_predictions (spot order n = b*H*W + y*W + x)
  gridnext/gridnet_models.py:128-148  hex corrector: hex hex [BN] ReLU hex hex [BN] ReLU hex
  gridnext/gridnet_models.py:165-187  GridNetHexOddr (rot90+flip == transpose; eliminated here)
  gridnext/gridnet_models.py:226-235  GridNetHexMM: cat((count, image), dim=1)
  gridnext/densenet.py:21-159         DenseNet-BC, eval-mode BN (training.py:126)
  gridnext/training.py:141-171        fg-masked mean cross-entropy, labels-1, argmax
  notebooks/Tutorial_visium_count.ipynb cell 12  count MLP
"""
import re
import torch
import torch.nn.functional as F
from .hexconv_ref import hexconv_visium

BN_EPS = 1e-5
BN_MOMENTUM = 0.1

COUNT_MLP_SPEC = ('L', 'L', 'B', 'R', 'L', 'L', 'B', 'R', 'L')


def sub(sd, prefix):
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


# ----------------------------------------------------------------------------- count MLP f
def mlp_forward(sd, x, spec=COUNT_MLP_SPEC, training=False, stats_out=None, emulate_bf16=False):
    """x: (N, G).  BatchNorm1d uses running stats unless ``training`` (GridNetHexMM quirk,
    training.py:126 only puts ``patch_classifier`` in eval).

    ``emulate_bf16``: same fp32 arithmetic with the input, the Linear weights and every stored activation (after a Linear
    that is not followed by BatchNorm, and after each ReLU) rounded to bfloat16 where the B200 path stores bf16, so ReLU
    masks agree with the kernels (see densenet_forward)."""
    e = emulate_bf16
    x = _rb(x, e)
    last = len(spec) - 1
    for i, kind in enumerate(spec):
        p = '%d.' % i
        if kind == 'L':
            x = F.linear(x, _rb(sd[p + 'weight'], e), sd[p + 'bias'])
            if i != last and spec[i + 1] != 'B':
                x = _rg(_rb(x, e), e)
            elif i != last:
                x = _rg(x, e)
                if training:
                    x = _rb(x, e)       # train-mode BN: the raw Linear output is stored (bf16) before its statistics exist
        elif kind == 'B':
            if training:
                mean = x.mean(0)
                var = x.var(0, unbiased=False)
                if stats_out is not None:
                    n = x.shape[0]
                    stats_out[p + 'running_mean'] = (1 - BN_MOMENTUM) * sd[p + 'running_mean'] + BN_MOMENTUM * mean.detach()
                    stats_out[p + 'running_var'] = (1 - BN_MOMENTUM) * sd[p + 'running_var'] + BN_MOMENTUM * var.detach() * n / (n - 1)
            else:
                mean, var = sd[p + 'running_mean'], sd[p + 'running_var']
            x = (x - mean) / torch.sqrt(var + BN_EPS) * sd[p + 'weight'] + sd[p + 'bias']
        elif kind == 'R':
            x = _rb(torch.relu(x), e)
        else:
            raise ValueError(kind)
    return x


# ----------------------------------------------------------------------------- DenseNet f
def _bn_eval(sd, p, x):
    scale = sd[p + 'weight'] / torch.sqrt(sd[p + 'running_var'] + BN_EPS)
    shift = sd[p + 'bias'] - sd[p + 'running_mean'] * scale
    return x * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


def densenet_block_config(sd):
    cfg = {}
    for k in sd:
        m = re.match(r'features\.denseblock(\d+)\.denselayer(\d+)\.', k)
        if m:
            b, l = int(m.group(1)), int(m.group(2))
            cfg[b] = max(cfg.get(b, 0), l)
    return tuple(cfg[b] for b in sorted(cfg))


class _RoundGradBf16(torch.autograd.Function):
    """Identity whose gradient is rounded to bfloat16 (where the B200 path stores a bf16 gradient tensor)."""

    @staticmethod
    def forward(ctx, t):
        return t.view_as(t)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _rb(t, on):
    """Straight-through bfloat16 rounding of the forward value (gradient passes unrounded)."""
    if not on:
        return t
    return t + (t.detach().to(torch.bfloat16).to(t.dtype) - t.detach())


def _rg(t, on):
    return _RoundGradBf16.apply(t) if on else t


def _bn2d(sd, p, x, training, stats_out):
    """nn.BatchNorm2d of the reference's DenseNet (densenet.py:24-29,50,134): running statistics in eval mode, batch statistics
    (biased variance; running stats updated with momentum 0.1 and the unbiased variance) in train mode."""
    if not training:
        return _bn_eval(sd, p, x)
    mean = x.mean((0, 2, 3))
    var = x.var((0, 2, 3), unbiased=False)
    if stats_out is not None:
        n = x.numel() // x.shape[1]
        stats_out[p + 'running_mean'] = (1 - BN_MOMENTUM) * sd[p + 'running_mean'] + BN_MOMENTUM * mean.detach()
        stats_out[p + 'running_var'] = (1 - BN_MOMENTUM) * sd[p + 'running_var'] + BN_MOMENTUM * var.detach() * n / (n - 1)
    xh = (x - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + BN_EPS)
    return xh * sd[p + 'weight'].view(1, -1, 1, 1) + sd[p + 'bias'].view(1, -1, 1, 1)


def densenet_forward(sd, x, classify=True, emulate_bf16=False, training=False, stats_out=None):
    """DenseNet-BC forward (densenet.py:152-159).  x: (N, 3, P, P) float.  ``training``: train-mode BatchNorm (f pre-training,
    training.py:11-98); eval mode is what the grid-wise hot path uses (training.py:126).

    ``emulate_bf16``: same fp32 arithmetic, but every tensor the B200 path stores in bfloat16 (input, conv weights,
    activated operands, conv outputs, pooled transition input, and the gradients dZ / dC; in train mode also the raw conv0 /
    conv1 outputs, which must exist before their statistics do) is rounded at that point, and the transition pools BEFORE
    its 1x1 convolution (they commute; gridnext_b200/densenet.py).  This pins the kernels far tighter than the fp32
    comparison allows, because ReLU masks then agree."""
    e = emulate_bf16
    t = training
    small_inputs = 'features.norm0.weight' not in sd
    w0 = _rb(sd['features.conv0.weight'], e)
    x = _rb(x, e)
    if small_inputs:
        x = _rb(F.conv2d(x, w0, stride=1, padding=1), e)
    else:
        x = _rg(F.conv2d(x, w0, stride=2, padding=3), e)
        x = _rb(x, e and t)
        x = _rb(torch.relu(_bn2d(sd, 'features.norm0.', x, t, stats_out)), e)
        x = F.max_pool2d(x, 3, stride=2, padding=1)
    cfg = densenet_block_config(sd)
    for bi, nl in enumerate(cfg, start=1):
        x = _rg(x, e)
        for li in range(1, nl + 1):
            p = 'features.denseblock%d.denselayer%d.' % (bi, li)
            h = _rb(torch.relu(_bn2d(sd, p + 'norm1.', x, t, stats_out)), e)
            h = _rg(F.conv2d(h, _rb(sd[p + 'conv1.weight'], e)), e)
            h = _rb(h, e and t)
            h = _rb(torch.relu(_bn2d(sd, p + 'norm2.', h, t, stats_out)), e)
            h = _rg(_rb(F.conv2d(h, _rb(sd[p + 'conv2.weight'], e), padding=1), e), e)
            x = torch.cat((x, h), 1)
        if bi != len(cfg):
            p = 'features.transition%d.' % bi
            x = torch.relu(_bn2d(sd, p + 'norm.', x, t, stats_out))
            if e:
                x = _rg(_rb(F.avg_pool2d(x, 2, stride=2), e), e)
                x = _rb(F.conv2d(x, _rb(sd[p + 'conv.weight'], e)), e)
            else:
                x = F.conv2d(x, sd[p + 'conv.weight'])
                x = F.avg_pool2d(x, 2, stride=2)
    x = torch.relu(_bn2d(sd, 'features.norm_final.', x, t, stats_out))
    x = x.mean((2, 3))
    if classify:
        x = F.linear(x, sd['classifier.weight'], sd['classifier.bias'])
    return x


# ----------------------------------------------------------------------------- g corrector
def corrector_layout(use_bn=True):
    """Sequential indices of the hex layers / BN layers (gridnet_models.py:128-148)."""
    return ((0, 1, 4, 5, 8), (2, 6)) if use_bn else ((0, 1, 3, 4, 6), ())


def corrector_forward(sd, x, use_bn=True, training=True, stats_out=None, ksize=1):
    """sd keys '<idx>.kernel0' ...; x: (B, f_dim, H, W) in Visium odd-r layout."""
    hex_idx, bn_idx = corrector_layout(use_bn)

    def hexl(i, t):
        ks = [sd['%d.kernel%d' % (i, j)] for j in range(ksize + 1)]
        return hexconv_visium(t, ks, sd.get('%d.bias_tensor' % i))

    def bn(i, t):
        p = '%d.' % i
        if training:
            mean = t.mean((0, 2, 3))
            var = t.var((0, 2, 3), unbiased=False)
            if stats_out is not None:
                n = t.numel() // t.shape[1]
                stats_out[p + 'running_mean'] = (1 - BN_MOMENTUM) * sd[p + 'running_mean'] + BN_MOMENTUM * mean.detach()
                stats_out[p + 'running_var'] = (1 - BN_MOMENTUM) * sd[p + 'running_var'] + BN_MOMENTUM * var.detach() * n / (n - 1)
        else:
            mean, var = sd[p + 'running_mean'], sd[p + 'running_var']
        t = (t - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + BN_EPS)
        return t * sd[p + 'weight'].view(1, -1, 1, 1) + sd[p + 'bias'].view(1, -1, 1, 1)

    h = hexl(hex_idx[0], x)
    h = hexl(hex_idx[1], h)
    if use_bn:
        h = bn(bn_idx[0], h)
    h = torch.relu(h)
    h = hexl(hex_idx[2], h)
    h = hexl(hex_idx[3], h)
    if use_bn:
        h = bn(bn_idx[1], h)
    h = torch.relu(h)
    return hexl(hex_idx[4], h)


def cartesian_corrector_forward(sd, x, use_bn=True, training=True, stats_out=None):
    """Base GridNet corrector (gridnet_models.py:51-66): Conv2d 3x3, [BN], ReLU, Conv2d 5x5, [BN], ReLU, Conv2d 5x5, [BN], ReLU,
    Conv2d 3x3, all n_classes wide with 'same' zero padding.  sd keys '<idx>.weight' ... of the nn.Sequential."""
    step = 3 if use_bn else 2
    h, idx = x, 0
    for K in (3, 5, 5):
        h = F.conv2d(h, sd['%d.weight' % idx], sd['%d.bias' % idx], padding=K // 2)
        if use_bn:
            p = '%d.' % (idx + 1)
            if training:
                mean = h.mean((0, 2, 3))
                var = h.var((0, 2, 3), unbiased=False)
                if stats_out is not None:
                    n = h.numel() // h.shape[1]
                    stats_out[p + 'running_mean'] = (1 - BN_MOMENTUM) * sd[p + 'running_mean'] + BN_MOMENTUM * mean.detach()
                    stats_out[p + 'running_var'] = (1 - BN_MOMENTUM) * sd[p + 'running_var'] + BN_MOMENTUM * var.detach() * n / (n - 1)
            else:
                mean, var = sd[p + 'running_mean'], sd[p + 'running_var']
            h = (h - mean.view(1, -1, 1, 1)) / torch.sqrt(var.view(1, -1, 1, 1) + BN_EPS)
            h = h * sd[p + 'weight'].view(1, -1, 1, 1) + sd[p + 'bias'].view(1, -1, 1, 1)
        h = torch.relu(h)
        idx += step
    return F.conv2d(h, sd['%d.weight' % idx], sd['%d.bias' % idx], padding=1)


# ----------------------------------------------------------------------------- composite
def spots_from_counts(x):
    """(B, G, H, W) -> (B*H*W, G), spot order n = b*H*W + y*W + x (gridnet_models.py:168,83)."""
    B, G, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(-1, G)


def grid_from_spots(p, B, H, W):
    """(B*H*W, f) -> (B, f, H, W) (gridnet_models.py:106-107)."""
    return p.reshape(B, H, W, -1).permute(0, 3, 1, 2)


def gridnet_count_forward(sd, x, use_bn=True, training=True, stats_out=None, f_training=False, emulate_bf16=False):
    B, G, H, W = x.shape
    f = mlp_forward(sub(sd, 'patch_classifier.'), spots_from_counts(x), training=f_training, emulate_bf16=emulate_bf16)
    return corrector_forward(sub(sd, 'corrector.'), grid_from_spots(f, B, H, W), use_bn, training, stats_out)


def gridnet_image_forward(sd, x, use_bn=True, training=True, stats_out=None):
    B, H, W = x.shape[:3]
    f = densenet_forward(sub(sd, 'patch_classifier.'), x.reshape((-1,) + tuple(x.shape[3:])))
    return corrector_forward(sub(sd, 'corrector.'), grid_from_spots(f, B, H, W), use_bn, training, stats_out)


def gridnet_mm_forward(sd, x_image, x_count, use_bn=True, training=True, stats_out=None, count_f_training=None):
    """GridNetHexMM: count f stays in train mode during the train phase."""
    if count_f_training is None:
        count_f_training = training
    B, H, W = x_image.shape[:3]
    cstats = {} if stats_out is not None else None
    fc = mlp_forward(sub(sd, 'count_classifier.'), spots_from_counts(x_count), training=count_f_training, stats_out=cstats)
    if stats_out is not None:
        stats_out.update({'count_classifier.' + k: v for k, v in cstats.items()})
    fi = densenet_forward(sub(sd, 'image_classifier.'), x_image.reshape((-1,) + tuple(x_image.shape[3:])))
    f = torch.cat((grid_from_spots(fc, B, H, W), grid_from_spots(fi, B, H, W)), 1)
    gstats = {} if stats_out is not None else None
    out = corrector_forward(sub(sd, 'corrector.'), f, use_bn, training, gstats)
    if stats_out is not None:
        stats_out.update({'corrector.' + k: v for k, v in gstats.items()})
    return out


# ----------------------------------------------------------------------------- loss
def masked_ce(outputs, labels, accum_iters=1):
    """training.py:152-160.  outputs (B, C, H, W), labels (B, H, W) int64 with 0 = background.
    Returns (loss, n_correct, n_foreground)."""
    C = outputs.shape[1]
    o = outputs.permute(0, 2, 3, 1).reshape(-1, C)
    l = labels.reshape(-1)
    fg = l > 0
    o = o[fg]
    l = l[fg] - 1
    loss = F.cross_entropy(o, l) / accum_iters
    preds = o.argmax(1)
    return loss, int((preds == l).sum()), int(fg.sum())
