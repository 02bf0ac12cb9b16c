"""GPU parity: hexagonal convolution, fused corrector, masked CE and the count GridNet step against the
CPU oracle and the reference-generated golden vectors.  Everything goes through the C-ABI library."""
import json, os
import numpy as np
import pytest
import torch

from oracle import synth, shapes as S
from oracle import gridnet_ref as R
from oracle.hexconv_ref import hexconv_visium, hexconv_hexagdly, kernel_shapes
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
MAN = json.load(open(os.path.join(GOLDEN, 'manifest.json')))
TOL = 1e-5   # north_star: hex-conv outputs and gradients within 1e-5 fp32


def dev():
    return torch.device('cuda:0')


def rel_err(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


def rand_hex(cin, cout, k, B, H, W, seed):
    g = torch.Generator(); g.manual_seed(seed)
    ks = [torch.randn(s, generator=g) * (1.0 / (cin * 7) ** 0.5) for s in kernel_shapes(cin, cout, k)]
    b = torch.randn(cout, generator=g) * 0.1
    x = torch.randn(B, cin, H, W, generator=g)
    dy = torch.randn(B, cout, H, W, generator=g)
    return ks, b, x, dy


@pytest.mark.parametrize('k', [1, 2, 3])
@pytest.mark.parametrize('cfg', [(7, 32, 2, 78, 64), (32, 32, 1, 78, 64), (32, 7, 3, 78, 64), (3, 5, 2, 7, 9),
                                 (4, 4, 1, 4, 4), (14, 32, 1, 9, 70), (64, 64, 1, 13, 64), (16, 40, 2, 8, 130)])
def test_hexconv_fwd_bwd_matches_oracle(k, cfg):
    from gridnext_b200.hexagdly import hexconv_visium as gpu_hexconv
    cin, cout, B, H, W = cfg
    ks, b, x, dy = rand_hex(cin, cout, k, B, H, W, seed=k * 1000 + cin + H)
    ks_r = [t.clone().double().requires_grad_(True) for t in ks]
    b_r = b.clone().double().requires_grad_(True)
    x_r = x.clone().double().requires_grad_(True)
    y_r = hexconv_visium(x_r, ks_r, b_r)
    y_r.backward(dy.double())
    ks_g = [t.clone().to(dev()).requires_grad_(True) for t in ks]
    b_g = b.clone().to(dev()).requires_grad_(True)
    x_g = x.clone().to(dev()).requires_grad_(True)
    y_g = gpu_hexconv(x_g, ks_g, b_g)
    y_g.backward(dy.to(dev()))
    assert rel_err(y_g, y_r) < TOL
    assert rel_err(x_g.grad, x_r.grad) < TOL
    assert rel_err(b_g.grad, b_r.grad) < TOL
    for a, r in zip(ks_g, ks_r):
        assert rel_err(a.grad, r.grad) < TOL


@pytest.mark.parametrize('gen', ['2', '2u', '1'])
@pytest.mark.parametrize('cfg', [(7, 32, 2, 78, 64), (32, 32, 3, 78, 64), (32, 7, 2, 78, 64), (3, 5, 2, 7, 9), (4, 4, 1, 4, 4), (14, 32, 1, 9, 70),
                                 (32, 32, 17, 78, 64), (20, 12, 3, 11, 33), (32, 32, 5, 27, 64), (16, 24, 2, 53, 32), (32, 32, 40, 78, 64)])
def test_hexconv_tensor_core_path_matches_oracle(cfg, gen, monkeypatch):
    """kernel_size 1, <= 32 channels on tcgen05 (bf16 x 3 split): forward and data gradient still within 1e-5 of the fp64 oracle.
    gen 2 = csrc/hexconv_tc2.cu (fp32 in / out, operands converted in shared memory; grid width <= 64, a multiple of 4 -- other
    shapes fall through to gen 1 = csrc/hexconv_tc.cu with its parity-plane rewrite).  (27, 64): a strip boundary at row 26;
    (53, 32): three strips, half-width rows; B = 40: several strips per CTA (ring / staging phases wrap many times)."""
    from gridnext_b200 import hexagdly as hx
    monkeypatch.setattr(hx, 'TENSOR_CORE_MODE', '1')
    # '2u': the second-generation weight gradient (csrc/hexconv_wgrad_tc2.cu) with one MMA per tap instead of the taps stacked along N
    monkeypatch.setenv('GRIDNEXT_B200_HEXWG2_STACK', '0' if gen == '2u' else '1')
    monkeypatch.setattr(hx, 'TENSOR_CORE_GEN', gen[0])
    cin, cout, B, H, W = cfg
    ks, b, x, dy = rand_hex(cin, cout, 1, B, H, W, seed=77 + cin + H)
    ks_r = [t.clone().double().requires_grad_(True) for t in ks]
    b_r = b.clone().double().requires_grad_(True)
    x_r = x.clone().double().requires_grad_(True)
    y_r = hexconv_visium(x_r, ks_r, b_r)
    y_r.backward(dy.double())
    ks_g = [t.clone().to(dev()).requires_grad_(True) for t in ks]
    b_g = b.clone().to(dev()).requires_grad_(True)
    x_g = x.clone().to(dev()).requires_grad_(True)
    y_g = hx.hexconv_visium(x_g, ks_g, b_g)
    y_g.backward(dy.to(dev()))
    assert rel_err(y_g, y_r) < TOL
    assert rel_err(x_g.grad, x_r.grad) < TOL
    assert rel_err(b_g.grad, b_r.grad) < TOL
    for a, r in zip(ks_g, ks_r):
        assert rel_err(a.grad, r.grad) < TOL


@pytest.mark.parametrize('B', [3, 20])
def test_fused_corrector_on_tensor_cores(B, monkeypatch):
    """The 5-layer corrector with every hex layer on the tensor-core path: each layer is within 1e-5, the composition of five
    (with two train-mode BatchNorms in between) within 2e-5 of the fp64 oracle; B = 20 also takes this path in 'auto' mode."""
    from gridnext_b200 import hexagdly as hx
    from gridnext_b200.gridnet_models import GridNetHexOddr
    import torch.nn as nn
    monkeypatch.setattr(hx, 'TENSOR_CORE_MODE', '1' if B < 16 else 'auto')
    net = GridNetHexOddr(nn.Identity(), (7,), (78, 64), 7).to(dev()).train()
    sd = {k: v.detach().cpu() for k, v in net.corrector.state_dict().items()}
    g = torch.Generator(); g.manual_seed(3)
    x = torch.randn(B, 7, 78, 64, generator=g)
    dy = torch.randn(B, 7, 78, 64, generator=g)
    sd_r = {k: (v.double().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in sd.items()}
    x_r = x.double().requires_grad_(True)
    y_r = R.corrector_forward(sd_r, x_r, use_bn=True, training=True)
    y_r.backward(dy.double())
    x_g = x.to(dev()).requires_grad_(True)
    y_g = net._correct_visium(x_g)
    y_g.backward(dy.to(dev()))
    assert rel_err(y_g, y_r) < 2e-5

    # Gradients: a 5e-6 difference in a pre-ReLU activation flips the ReLU mask of the few cells that sit that close to zero.
    # Each flip changes its own gradient entries by O(1) and, through the train-mode BatchNorm backward (which subtracts batch
    # means), every other entry by ~1/N.  The fp32-FMA path has the same discontinuity (it flips ~10x fewer cells).  The kernels
    # themselves are pinned at 1e-5 per layer above; here the composed gradient must agree in L2 (1 %); a flipped cell may move its own entry by up to 25 % of the max-norm.
    def close(got, ref, name):
        got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
        l2 = float((got - ref).norm() / max(float(ref.norm()), 1e-3 * ref.numel() ** 0.5))
        mx = float((got - ref).abs().max() / max(float(ref.abs().max()), 1e-3))
        assert l2 < 1e-2 and mx < 0.25, (name, l2, mx)
    close(x_g.grad, x_r.grad, 'dx')
    params = dict(net.corrector.named_parameters())
    for name, p in params.items():
        ref = sd_r[name].grad
        if float(ref.abs().max()) < 1e-9:
            # bias in front of a train-mode BatchNorm: its gradient is exactly zero in theory (sum of mean-free terms); what is
            # left is summation noise, which must be negligible against the gradient of the same layer's kernel
            sib = params[name.replace('bias_tensor', 'kernel0')].grad
            assert float(p.grad.abs().max()) < 1e-2 * float(sib.abs().max()), name
            continue
        close(p.grad, ref, name)


def test_hexagdly_module_layout_matches_upstream_composition():
    import gridnext_b200.hexagdly as hx
    torch.manual_seed(3)
    m = hx.Conv2d(5, 6, kernel_size=2).to(dev())
    assert sorted(k for k, _ in m.named_parameters()) == ['bias_tensor', 'kernel0', 'kernel1', 'kernel2']
    assert repr(m) == 'Conv2d(5, 6, kernel_size=2, stride=1)'
    x = torch.randn(2, 5, 64, 78)      # HexagDLy layout (rows, cols): parity on the last index
    y = m(x.to(dev()))
    ref = hexconv_hexagdly(x.double(), [getattr(m, 'kernel%d' % i).detach().cpu().double() for i in range(3)],
                           m.bias_tensor.detach().cpu().double())
    assert tuple(y.shape) == (2, 6, 64, 78)
    assert rel_err(y, ref) < TOL


def _load_into(module, sd):
    missing = module.load_state_dict({k: v.detach().clone() for k, v in sd.items()}, strict=True)
    return module


@pytest.mark.parametrize('path', ['one_launch', 'layered'])
@pytest.mark.parametrize('use_bn,training', [(True, True), (True, False), (False, True)])
@pytest.mark.parametrize('shape', [(2, 7, 78, 64), (3, 14, 9, 7), (1, 7, 5, 70), (20, 7, 78, 64)])
def test_fused_corrector_matches_oracle(use_bn, training, shape, path, monkeypatch):
    """Both executions of the corrector -- the one-launch persistent kernel (csrc/corrector_fused.cu: all stages, BatchNorm
    reductions and weight gradients in one forward and one backward launch) and the layer-by-layer kernels -- against the fp64
    oracle: output, input gradient, every parameter gradient, running statistics.  (1, 7, 5, 70): two column tiles per row."""
    import torch.nn as nn
    from gridnext_b200 import corrector as corr, hexagdly as hx
    from gridnext_b200.gridnet_models import GridNetHexOddr
    monkeypatch.setattr(corr, 'FUSED_MODE', '1' if path == 'one_launch' else '0')
    monkeypatch.setattr(hx, 'TENSOR_CORE_MODE', '0')          # layered: the exact-fp32 kernels (the tensor-core ones have their own test)
    B, f_dim, H, W = shape
    n_cls = 5
    net = GridNetHexOddr(nn.Identity(), (f_dim,), (H, W), n_cls, use_bn=use_bn, f_dim=f_dim)
    sd = synth.synth_state_dict(S.gridnet_shapes({}, f_dim, n_cls, use_bn), 77)
    _load_into(net, sd)
    net.to(dev())
    net.train(training)
    g = torch.Generator(); g.manual_seed(B * 31 + H)
    x = torch.randn(B, f_dim, H, W, generator=g)
    dy = torch.randn(B, n_cls, H, W, generator=g)
    # oracle (fp64)
    sd_r = {k: (v.double().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in R.sub(sd, 'corrector.').items()}
    x_r = x.double().requires_grad_(True)
    stats = {}
    y_r = R.corrector_forward(sd_r, x_r, use_bn=use_bn, training=training, stats_out=stats)
    y_r.backward(dy.double())
    # product
    x_g = x.to(dev()).requires_grad_(True)
    y_g = net._correct_visium(x_g)
    y_g.backward(dy.to(dev()))
    assert rel_err(y_g, y_r) < TOL
    big = B * H * W > 50000
    # Gradients.  Among millions of pre-ReLU activations a few sit within fp32 rounding of zero; a flipped mask changes its own
    # gradient entries by O(1) (see test_fused_corrector_on_tensor_cores), so on the big batch the composed gradient is compared
    # in L2 (1e-3) with a loose max-norm, while the small shapes keep the 2e-5 max-norm check.
    def check(got, ref, name):
        got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
        if big:
            l2 = float((got - ref).norm() / max(float(ref.norm()), 1e-12))
            mx = float((got - ref).abs().max() / max(float(ref.abs().max()), 1e-12))
            assert l2 < 5e-3 and mx < 0.25, (name, l2, mx)       # measured up to 2.2e-3 / 0.075 (B = 20 without BatchNorm: a handful of flipped cells)
        else:
            assert float((got - ref).abs().max()) / float(ref.abs().max()) < 2 * TOL, name
    check(x_g.grad, x_r.grad, 'dx')
    params = dict(net.corrector.named_parameters())
    for name, p in params.items():
        ref = sd_r[name].grad
        if float(ref.abs().max()) < 1e-9:
            # exactly zero in theory (bias in front of a train-mode BatchNorm): fp32 cancellation residue of a sum over B*H*W cells,
            # negligible against the gradient of the same layer's kernel
            sib = params[name.replace('bias_tensor', 'kernel0')].grad
            assert float(p.grad.abs().max()) < 1e-4 * float(sib.abs().max()), name
        else:
            check(p.grad, ref, name)
    if use_bn and training:
        for k, v in stats.items():
            assert rel_err(dict(net.corrector.named_buffers())[k], v) < TOL, k
        assert int(net.corrector[2].num_batches_tracked) == 1


def test_masked_ce_matches_oracle():
    from gridnext_b200.losses import masked_cross_entropy
    g = torch.Generator(); g.manual_seed(5)
    logits = torch.randn(3, 7, 78, 64, generator=g) * 3
    labels = synth.synth_labels(3, 7, seed=2)
    lr = logits.double().requires_grad_(True)
    loss_r, ncorr, nfg = R.masked_ce(lr, labels, accum_iters=2)
    loss_r.backward()
    lg = logits.to(dev()).requires_grad_(True)
    loss_g, acc = masked_cross_entropy(lg, labels.to(dev()), accum_iters=2)
    loss_g.backward()
    acc = acc.tolist()
    assert abs(float(loss_g) - float(loss_r)) < 1e-6
    assert (int(acc[2]), int(acc[1])) == (ncorr, nfg)
    assert rel_err(lg.grad, lr.grad) < TOL
    # all-background batch: CrossEntropyLoss over an empty selection is NaN in the reference too
    loss_e, acc_e = masked_cross_entropy(logits.to(dev()), torch.zeros_like(labels).to(dev()))
    assert torch.isnan(loss_e) and acc_e.tolist()[1] == 0


def test_count_gridnet_step_matches_reference_golden():
    """One train_gridwise iteration of the count model against vectors from the REAL reference."""
    import torch.nn as nn
    from gridnext_b200.gridnet_models import GridNetHexOddr
    from gridnext_b200.training import gridwise_step
    m = MAN['g1_count_gridnet']
    gold = np.load(os.path.join(GOLDEN, 'g1_count_gridnet.npz'))
    G, n_cls, B = m['G'], m['n_cls'], m['B']
    f = nn.Sequential(nn.Linear(G, 500), nn.Linear(500, 100), nn.BatchNorm1d(100), nn.ReLU(),
                      nn.Linear(100, 100), nn.Linear(100, 50), nn.BatchNorm1d(50), nn.ReLU(), nn.Linear(50, n_cls))
    net = GridNetHexOddr(f, (G,), (78, 64), n_cls, use_bn=True)
    sd = synth.synth_state_dict(S.gridnet_shapes(S.mlp_shapes(G, n_cls), n_cls, n_cls), m['seed_w'])
    _load_into(net, sd)
    net.to(dev())
    x = synth.synth_counts(B, G, seed=m['seed_x']).to(dev())
    y = synth.synth_labels(B, n_cls, seed=m['seed_y']).to(dev())
    net.train(); net.patch_classifier.eval()
    out = net(x)
    loss, acc, _ = gridwise_step(net, x, y, nn.CrossEntropyLoss(), 1, True)
    # f runs on bf16 tensor-core GEMMs (north_star: bf16 logits within 2e-2 of the fp32 reference); g is fp32
    assert rel_err(out, torch.from_numpy(gold['out'])) < 2e-2
    assert abs(float(loss) - float(gold['loss'])) < 1e-2 * max(1.0, abs(float(gold['loss'])))
    assert int(acc.tolist()[1]) == int(gold['nfg'])
    assert abs(int(acc.tolist()[2]) - int(gold['ncorr'])) <= 0.01 * int(gold['nfg'])     # near-tied logits may flip under bf16
    n = 0
    for k in gold.files:
        if k.startswith('grad.'):
            p = dict(net.named_parameters())[k[5:]]
            ref = torch.from_numpy(gold[k])
            scale = max(float(ref.abs().max()), 1e-4)
            err = float((p.grad.cpu() - ref).abs().max()) / scale
            if k.startswith('grad.corrector.'):
                assert err < 0.1, (k, err)           # fp32 kernels fed by bf16 features (g alone is pinned at 1e-5 above)
            else:
                # f gradients are random-sign sums: ReLU-mask flips between bf16 and fp32 forward values move them by
                # ~sqrt(flips / N); tests/test_gpu_count_mlp.py pins them at 3e-2 against the bf16-emulating oracle
                cos = float(torch.nn.functional.cosine_similarity(p.grad.flatten().double().cpu(), ref.flatten().double(), dim=0))
                assert cos > 0.98 and err < 0.3, (k, cos, err)
            n += 1
    assert n >= 30


# ---------------------------------------------------------------------------------------------------------------------
# Base (Cartesian) GridNet: square-conv corrector on the same tile kernels
@pytest.mark.parametrize('K', [1, 3, 5])
@pytest.mark.parametrize('cfg', [(6, 5, 3, 9, 11), (7, 7, 2, 78, 64), (32, 7, 1, 20, 70), (3, 40, 2, 8, 130), (16, 16, 1, 4, 4)])
def test_square_conv_fwd_bwd_matches_conv2d(K, cfg):
    """gn_sqconv_* == F.conv2d(stride 1, padding K//2): forward, data gradient, weight and bias gradients (fp32, 1e-5)."""
    import torch.nn.functional as F
    from gridnext_b200 import hexagdly as hx
    cin, cout, B, H, W = cfg
    g = torch.Generator(); g.manual_seed(100 * K + cin)
    w = torch.randn(cout, cin, K, K, generator=g) / (cin * K * K) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    x = torch.randn(B, cin, H, W, generator=g)
    dy = torch.randn(B, cout, H, W, generator=g)
    xr, wr, br = x.clone().requires_grad_(True), w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.conv2d(xr, wr, br, padding=K // 2)
    (ref * dy).sum().backward()
    xd, wd, bd, dyd = x.cuda(), w.cuda(), b.cuda(), dy.cuda()
    y = hx.hexconv_fwd(xd, hx.pack_weights([wd], K, cin, cout, 0, 'sq'), bd, cout, K, kind='sq')
    assert rel_err(y, ref) < TOL
    dx = hx.hexconv_fwd(dyd, hx.pack_weights([wd], K, cin, cout, 1, 'sq'), None, cin, K, kind='sq')
    assert rel_err(dx, xr.grad) < TOL
    dwp, db = hx.hexconv_wgrad(xd, dyd, K, kind='sq')
    dw = hx.unpack_grad(dwp, [w.shape], K, cin, cout, 'sq')[0]
    assert rel_err(dw, wr.grad) < TOL and rel_err(db, br.grad) < TOL


@pytest.mark.parametrize('tag', ['c1_cartesian_bn', 'c2_cartesian_nobn'])
def test_cartesian_gridnet_step_matches_reference_golden(tag):
    """gridnext_b200.GridNet (base class, gridnet_models.py:24-109) against the reference's own GridNet: output, loss, every
    gradient and the BatchNorm running statistics after one train_gridwise iteration."""
    import torch.nn as nn
    from gridnext_b200.gridnet_models import GridNet
    from gridnext_b200.training import gridwise_step
    from gridnext_b200.corrector import parse_corrector
    m = MAN[tag]
    gold = np.load(os.path.join(GOLDEN, tag + '.npz'))
    net = GridNet(nn.Linear(4, m['f_dim']), (4,), (m['H'], m['W']), m['n_cls'], use_bn=m['use_bn'], f_dim=m['f_dim'])
    sd = synth.synth_state_dict(S.cartesian_gridnet_shapes({'weight': (m['f_dim'], 4), 'bias': (m['f_dim'],)}, m['f_dim'], m['n_cls'], m['use_bn']), m['seed_w'])
    assert set(net.state_dict().keys()) == set(sd.keys())
    net.load_state_dict(sd)
    net.cuda()
    assert parse_corrector(net.corrector) is not None
    net.train(); net.patch_classifier.eval()
    x, y = torch.from_numpy(gold['x']).cuda(), torch.from_numpy(gold['y']).cuda()
    out = net(x)
    assert rel_err(out, torch.from_numpy(gold['out'])) < 1e-4
    net.load_state_dict(sd)       # undo the running-stat update of the probe above
    loss, acc, _ = gridwise_step(net, x, y, nn.CrossEntropyLoss(), 1, True)
    assert abs(float(loss.detach()) - float(gold['loss'])) < 1e-5
    assert (int(acc.tolist()[2]), int(acc.tolist()[1])) == (int(gold['ncorr']), int(gold['nfg']))
    params = dict(net.named_parameters())
    got_sd = net.state_dict()
    zero_theory = ('corrector.0.bias', 'corrector.3.bias', 'corrector.6.bias') if m['use_bn'] else ()   # a bias in front of train-mode BN
    for k in gold.files:
        if k.startswith('grad.') and k[5:] in zero_theory:
            scale = float(np.abs(gold['grad.' + k[5:-4] + 'weight']).sum((1, 2, 3)).max())
            assert float(params[k[5:]].grad.abs().max()) < 1e-5 * scale, k
        elif k.startswith('grad.'):
            assert rel_err(params[k[5:]].grad, torch.from_numpy(gold[k])) < 1e-3, k
        elif k.startswith('after.'):
            assert rel_err(got_sd[k[6:]], torch.from_numpy(gold[k])) < 1e-5, k
