"""GPU parity of the DenseNet f network (forward logits and parameter gradients) against the CPU oracle and the
reference-generated golden vectors.  bf16 tensor-core path: north_star tolerance 2e-2 on logits."""
import json, os
import numpy as np
import pytest
import torch

from oracle import synth, shapes as S
from oracle import gridnet_ref as R
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
MAN = json.load(open(os.path.join(GOLDEN, 'manifest.json')))


def build(tag_or_cfg, seed):
    from gridnext_b200.densenet import DenseNet
    kw = tag_or_cfg
    net = DenseNet(num_classes=7, small_inputs=False, efficient=False, drop_rate=0, **kw)
    sd = synth.synth_state_dict(S.densenet_shapes(kw['growth_rate'], tuple(kw['block_config']), kw['num_init_features'], kw['bn_size']), seed)
    net.load_state_dict(sd)            # strict: key set identical to the reference's
    return net.cuda().eval(), sd


def relmax(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


@pytest.mark.parametrize('tag', ['d2_densenet_tiny_p32', 'd1_densenet121_p64'])
def test_densenet_matches_reference_golden(tag):
    m = MAN[tag]
    gold = np.load(os.path.join(GOLDEN, tag + '.npz'))
    kw = dict(growth_rate=m['growth_rate'], block_config=tuple(m['block_config']), num_init_features=m['num_init_features'], bn_size=m['bn_size'])
    net, sd = build(kw, m['seed_w'])
    g = torch.Generator(); g.manual_seed(m['seed_x'])
    x = torch.randn(m['N'], 3, m['P'], m['P'], generator=g)
    logits = net(x.cuda())
    assert relmax(logits.detach(), gold['logits']) < 2e-2
    g = torch.Generator(); g.manual_seed(m['seed_dy'])
    dy = torch.randn(logits.shape, generator=g)
    (logits * dy.cuda()).sum().backward()
    params = dict(net.named_parameters())
    norms = dict(zip(gold['grad_norm_keys'].tolist(), gold['grad_norm_vals'].tolist()))
    # bf16 activations / gradients (north_star: bf16 path, 2e-2 on logits): parameter gradients are compared by norm
    # (<= 12 %), direction (cosine >= 0.98 per tensor) and max-norm relative error (<= 0.2) -- ReLU-mask flips under bf16 rounding
    # are discrete events, so a handful of spots gives percent-level noise; the fp32 oracle pins the math.
    bad = []
    for k, v in norms.items():
        got = float(params[k].grad.norm())
        if abs(got - v) > 0.12 * max(v, 1e-3):
            bad.append((k, got, v))
    assert not bad, bad[:10]
    for k in gold.files:
        if k.startswith('grad.'):
            ref = torch.from_numpy(gold[k])
            got = params[k[5:]].grad.cpu()
            cos = float(torch.nn.functional.cosine_similarity(got.flatten().double(), ref.flatten().double(), dim=0))
            assert cos > 0.98, (k, cos)              # 120 layers of bf16 backward: noise grows towards the stem
            assert relmax(got, ref) < 0.3, k                 # fp32 reference vs bf16 path; the tight check is the bf16-emulating oracle below


@pytest.mark.parametrize('kw,P,N,deep', [(dict(growth_rate=8, block_config=(2, 2), num_init_features=16, bn_size=2), 32, 5, False),
                                          (dict(growth_rate=16, block_config=(3, 4), num_init_features=32, bn_size=4), 48, 4, False),
                                          (dict(growth_rate=32, block_config=(6, 12, 24, 16), num_init_features=64, bn_size=4), 64, 3, True)])
def test_densenet_matches_bf16_emulating_oracle(kw, P, N, deep):
    """Kernel parity proper: the oracle rounds to bfloat16 exactly where the B200 path stores bf16 (operands, activations,
    dZ / dC), so ReLU masks agree and what is left is fp32 summation order: logits 5e-3, and on the shallow networks
    every parameter gradient within 3e-2 of its max-norm (an order of magnitude tighter than the fp32 comparison above).
    A random-init DenseNet-121 is chaotic in its gradients: the emulating oracle itself moves by a median 10 % per tensor
    against the fp32 oracle (tools/emul_dist.py), so there the gradient check is statistical (direction + median)."""
    net, sd = build(kw, 91)
    g = torch.Generator(); g.manual_seed(17)
    x = torch.randn(N, 3, P, P, generator=g)
    dy = torch.randn(N, 7, generator=g)
    sd_r = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in sd.items()}
    ref = R.densenet_forward(sd_r, x, emulate_bf16=True)
    (ref * dy).sum().backward()
    logits = net(x.cuda())
    (logits * dy.cuda()).sum().backward()
    assert relmax(logits.detach(), ref.detach()) < 5e-3
    errs, coss = [], []
    for k, p in net.named_parameters():
        r = sd_r[k].grad
        assert torch.isfinite(p.grad).all(), k
        errs.append((relmax(p.grad, r), k))
        if deep:
            coss.append((float(torch.nn.functional.cosine_similarity(p.grad.flatten().double().cpu(), r.flatten().double(), dim=0)), k))
    errs.sort(reverse=True)
    if deep:
        # statistical: a change of fp32 summation order inside one kernel moves individual tensors of the late blocks (48 pixel
        # rows at this size) by a few percent; every tensor keeps its direction (0.95), all but a handful stay above 0.98
        coss.sort()
        assert coss[0][0] > 0.95, coss[:3]
        assert sum(1 for c, _ in coss if c < 0.98) <= len(coss) // 50, coss[:8]
        assert errs[len(errs) // 2][0] < 0.15 and errs[0][0] < 0.6, errs[:3]     # measured: median 0.10, max 0.42 (one BN weight of block 2)
    else:
        assert errs[0][0] < 3e-2, errs[:3]


def test_densenet121_p128_matches_oracle_and_argmax():
    """Full benchmark shape (3x128x128) on a handful of spots: logits vs the fp32 oracle, bf16 tolerance + argmax agreement."""
    kw = dict(growth_rate=32, block_config=(6, 12, 24, 16), num_init_features=64, bn_size=4)
    net, sd = build(kw, 77)
    g = torch.Generator(); g.manual_seed(5)
    x = torch.rand(6, 3, 128, 128, generator=g)
    x = (x - 0.45) / 0.225
    with torch.no_grad():
        ref = R.densenet_forward(sd, x)
        out = net(x.cuda()).cpu()
    assert relmax(out, ref) < 2e-2
    top2 = ref.topk(2, 1).values
    clear = (top2[:, 0] - top2[:, 1]) > 2e-2 * ref.abs().max()
    assert torch.equal(out.argmax(1)[clear], ref.argmax(1)[clear])


@pytest.mark.parametrize('kw,P,N,tight', [(dict(growth_rate=8, block_config=(2, 2), num_init_features=16, bn_size=2), 32, 6, True),
                                           (dict(growth_rate=16, block_config=(3, 3), num_init_features=32, bn_size=4), 32, 8, False),
                                           (dict(growth_rate=8, block_config=(4, 2), num_init_features=16, bn_size=2), 32, 8, False),
                                           (dict(growth_rate=32, block_config=(2, 2, 2), num_init_features=64, bn_size=4), 64, 12, False)])
def test_densenet_train_mode_bn_matches_oracle(kw, P, N, tight):
    """Train-mode BatchNorm (f pre-training, training.py:11-98): batch statistics, running-stat update, and the full
    BatchNorm gradient (mean / variance terms included), against the bf16-emulating oracle
    (oracle.densenet_forward(training=True) is pinned to the reference's DenseNet in tests/test_oracle_golden.py).
    Batch-statistic BatchNorm over a few hundred samples makes the gradients chaotic in the rounding: multiplying every
    tensor of the emulating oracle by (1 + 1e-6 * randn) before it is rounded to bf16 -- the size of an fp32 summation-order
    difference -- moves the (3, 3) network's gradients by a median 15 % (max 49 %) and the logits by 4e-3.  Those cases are
    therefore checked statistically (median, direction); the (2, 2) network is benign enough for a 5e-2 max-norm check."""
    net, sd = build(kw, 23)
    net.train()
    g = torch.Generator(); g.manual_seed(29)
    x = torch.randn(N, 3, P, P, generator=g)
    dy = torch.randn(N, 7, generator=g)
    sd_r = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in sd.items()}
    stats = {}
    ref = R.densenet_forward(sd_r, x, emulate_bf16=True, training=True, stats_out=stats)
    (ref * dy).sum().backward()
    logits = net(x.cuda())
    (logits * dy.cuda()).sum().backward()
    assert relmax(logits.detach(), ref.detach()) < 1e-2
    got_sd = net.state_dict()
    for k, v in stats.items():
        assert relmax(got_sd[k], v) < 5e-3, k
    for k, v in got_sd.items():
        if k.endswith('num_batches_tracked'):
            assert int(v) == int(sd[k]) + 1, k
    errs = []
    # every consumer of norm0's output is itself a train-mode BatchNorm, so the loss is invariant to scaling (gamma0, beta0)
    # together: gamma*dgamma + beta*dbeta = 0 and both gradients are cancellation residue, two orders of magnitude below
    # their neighbours (the fp32 and the bf16-emulating oracle disagree on them by 100 %).  They are compared on the scale
    # of the first norm1's gradient instead of their own.
    sib = float(sd_r['features.denseblock1.denselayer1.norm1.bias'].grad.abs().max())
    for k, p in net.named_parameters():
        assert torch.isfinite(p.grad).all(), k
        if k.startswith('features.norm0.'):
            errs.append((float((p.grad.cpu() - sd_r[k].grad).abs().max()) / sib, k))
        else:
            errs.append((relmax(p.grad, sd_r[k].grad), k))
    errs.sort(reverse=True)
    if tight:
        assert errs[0][0] < 5e-2, errs[:5]
    else:
        assert errs[len(errs) // 2][0] < 0.2 and errs[0][0] < 0.6, errs[:5]
        for k, p in net.named_parameters():
            if not k.startswith('features.norm0.'):
                cos = float(torch.nn.functional.cosine_similarity(p.grad.flatten().double().cpu(), sd_r[k].grad.flatten().double(), dim=0))
                assert cos > 0.9, (k, cos)
    # and the fp32 reference semantics (no rounding emulation): logits within the bf16 tolerance
    with torch.no_grad():
        ref32 = R.densenet_forward(sd, x, training=True)
    assert relmax(logits.detach(), ref32) < 3e-2


def test_densenet_train_step_then_eval_uses_updated_running_stats():
    """train() forward updates the running statistics in place; the following eval() forward must see them."""
    kw = dict(growth_rate=8, block_config=(2, 2), num_init_features=16, bn_size=2)
    net, sd = build(kw, 31)
    g = torch.Generator(); g.manual_seed(3)
    x = torch.randn(8, 3, 32, 32, generator=g)
    net.train()
    with torch.no_grad():
        net(x.cuda())
    stats = {}
    with torch.no_grad():
        R.densenet_forward(sd, x, training=True, stats_out=stats)
    sd2 = dict(sd)
    sd2.update(stats)
    net.eval()
    with torch.no_grad():
        out = net(x.cuda()).cpu()
        ref = R.densenet_forward(sd2, x)
    assert relmax(out, ref) < 2e-2


def test_densenet_rejects_unsupported_modes():
    from gridnext_b200.densenet import DenseNet
    net = DenseNet(num_classes=7, small_inputs=False, growth_rate=8, block_config=(2, 2), num_init_features=16, bn_size=2).cuda()
    net.eval()
    net.features.norm0.train()
    with pytest.raises(NotImplementedError):
        net(torch.zeros(2, 3, 32, 32, device='cuda'))
    with pytest.raises(RuntimeError):
        net.eval()(torch.zeros(1, 3, 32, 32))


def test_default_constructor_runs_forward_and_backward():
    """The reference's DEFAULT DenseNet() (densenet.py:93-95: growth_rate 12, block_config (16, 16, 16), 24 stem channels,
    small_inputs=True) lies outside the tcgen05 kernels' channel envelope (12 is not a multiple of 8): it runs the reference module
    graph with PyTorch's CUDA operators, announced by a warning, and matches the fp32 oracle; train-mode dropout takes the same path."""
    import warnings
    from gridnext_b200.densenet import DenseNet
    torch.manual_seed(0)
    net = DenseNet(num_classes=10).cuda().eval()
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    g = torch.Generator(); g.manual_seed(4)
    x = torch.randn(3, 3, 32, 32, generator=g)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter('always')
        out = net(x.cuda())
    assert list(out.shape) == [3, 10]
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    with torch.no_grad():
        ref = R.densenet_forward(sd, x)
    assert relmax(out.detach(), ref) < 2e-2                    # cuDNN's default TF32 convolutions vs the fp32 oracle
    out.sum().backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())
    drop = DenseNet(growth_rate=8, block_config=(2, 2), num_init_features=16, bn_size=2, num_classes=7, small_inputs=False, drop_rate=0.2).cuda().train()
    y = drop(torch.randn(4, 3, 32, 32, device='cuda'))
    assert list(y.shape) == [4, 7] and torch.isfinite(y).all()
    with pytest.raises(RuntimeError):
        DenseNet(num_classes=10)(torch.zeros(1, 3, 32, 32))     # CPU tensors are still refused: there is no CPU path


@pytest.mark.parametrize('signed', [False, True])
@pytest.mark.parametrize('N,H,C,ldo', [(3, 16, 64, 256), (2, 8, 16, 16), (1, 64, 64, 96)])
def test_stem_maxpool_forward_backward_match_torch(N, H, C, ldo, signed):
    """pool0 (densenet.py:111, MaxPool2d(3, 2, 1)): values bit-equal, arg-max taps equal to torch's first-maximum rule (ties included:
    the input is quantised to a few levels), and the fused pool / ReLU / BatchNorm-scale backward against autograd."""
    import torch.nn.functional as F
    from gridnext_b200._lib import call, ptr, stream
    g = torch.Generator(); g.manual_seed(11)
    x = (torch.randn((N, C, H, H), generator=g) * 2).round() / 2            # many exact ties
    if not signed:
        x = torch.relu(x)
    x = (x + 0.0).to(torch.bfloat16)                                        # no -0.0: torch's max treats it as a tie with +0.0
    ref, ridx = F.max_pool2d(x.float(), 3, 2, 1, return_indices=True)
    Ho = H // 2
    oy = torch.arange(Ho).view(1, 1, Ho, 1); ox = torch.arange(Ho).view(1, 1, 1, Ho)
    ky = ridx // H - (2 * oy - 1); kx = ridx % H - (2 * ox - 1)
    rtap = (ky * 3 + kx).permute(0, 2, 3, 1).reshape(-1, C)
    xin = x.permute(0, 2, 3, 1).reshape(-1, C).contiguous().cuda()
    out = torch.zeros((N * Ho * Ho, ldo), dtype=torch.bfloat16, device='cuda')
    idx = torch.empty((N * Ho * Ho, C), dtype=torch.uint8, device='cuda')
    call('gn_maxpool3s2_fwd', ptr(xin), C, N, H, H, C, ptr(out), ldo, ptr(idx), stream())
    assert torch.equal(out[:, :C].float().cpu(), ref.permute(0, 2, 3, 1).reshape(-1, C))
    assert torch.equal(idx.cpu().long(), rtap)
    if signed:
        return
    # backward: dz = scatter(dpool) * [act > 0] * sc, column sums sum g and p1 * (sum g*act - p0 * sum g)
    dp = (torch.randn((N * Ho * Ho, C), generator=g) * 0.1).to(torch.bfloat16)
    sc, p0, p1 = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g) * 0.1, torch.rand(C, generator=g) + 0.5
    xr = x.float().requires_grad_(True)
    F.max_pool2d(xr, 3, 2, 1).backward(dp.float().reshape(N, Ho, Ho, C).permute(0, 3, 1, 2))
    gm = (xr.grad * (x.float() > 0)).permute(0, 2, 3, 1).reshape(-1, C)
    dpb = torch.zeros((N * Ho * Ho, ldo), dtype=torch.bfloat16, device='cuda'); dpb[:, :C] = dp.cuda()
    dz = torch.empty((N * H * H, C), dtype=torch.bfloat16, device='cuda')
    colsum = torch.zeros((2, C), dtype=torch.float32, device='cuda')
    scd, p0d, p1d = sc.cuda(), p0.cuda(), p1.cuda()
    call('gn_maxpool3s2_bnrelu_bwd', ptr(dpb), ldo, ptr(idx), ptr(xin), C, N, H, H, C, ptr(scd), ptr(p0d), ptr(p1d), ptr(dz), C, ptr(colsum), C,
         stream())
    act = xin.float().cpu()
    assert float((dz.float().cpu() - gm * sc).abs().max()) <= 1e-2 * float((gm * sc).abs().max())
    assert float((colsum[0].cpu() - gm.sum(0)).abs().max()) <= 1e-4 * max(1.0, float(gm.sum(0).abs().max()))
    ref_x = p1 * ((gm * act).sum(0) - p0 * gm.sum(0))
    assert float((colsum[1].cpu() - ref_x).abs().max()) <= 1e-4 * max(1.0, float(ref_x.abs().max()))
