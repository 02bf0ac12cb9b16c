"""GPU parity of the count-MLP f on the tcgen05 kernels (bf16 operands, fp32 accumulation; north_star: bf16 within 2e-2)
against the fp32 CPU oracle, forward and every parameter gradient, through GridNetHexOddr.patch_predictions."""
import pytest
import torch
import torch.nn as nn

from oracle import synth, shapes as S
from oracle import gridnet_ref as R

pytestmark = pytest.mark.gpu


def relmax(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


def tutorial_mlp(G, n_cls):
    return nn.Sequential(nn.Linear(G, 500), nn.Linear(500, 100), nn.BatchNorm1d(100), nn.ReLU(),
                         nn.Linear(100, 100), nn.Linear(100, 50), nn.BatchNorm1d(50), nn.ReLU(), nn.Linear(50, n_cls))


@pytest.mark.parametrize('B,G,H,W', [(2, 1000, 78, 64), (1, 5000, 78, 64), (3, 64, 4, 4), (1, 333, 10, 8)])
def test_count_mlp_forward_backward_matches_oracle(B, G, H, W):
    from gridnext_b200.gridnet_models import GridNetHexOddr
    from gridnext_b200.count_mlp import compile_count_mlp
    n_cls = 7
    f = tutorial_mlp(G, n_cls)
    net = GridNetHexOddr(f, (G,), (H, W), n_cls, use_bn=True)
    sd = synth.synth_state_dict(S.gridnet_shapes(S.mlp_shapes(G, n_cls), n_cls, n_cls), 11)
    net.load_state_dict(sd)
    net.cuda()
    net.train(); net.patch_classifier.eval()
    assert compile_count_mlp(net.patch_classifier) is not None
    g = torch.Generator(); g.manual_seed(4)
    x = torch.log1p(torch.poisson(torch.ones(B, G, H, W), generator=g))
    dy = torch.randn(B, n_cls, H, W, generator=g)
    out = net.patch_predictions(x.cuda())
    assert tuple(out.shape) == (B, n_cls, H, W)
    (out * dy.cuda()).sum().backward()
    def oracle(emulate):
        sd_r = {k[len('patch_classifier.'):]: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v)
                for k, v in sd.items() if k.startswith('patch_classifier.')}
        ref = R.grid_from_spots(R.mlp_forward(sd_r, R.spots_from_counts(x), emulate_bf16=emulate), B, H, W)
        (ref * dy).sum().backward()
        return ref.detach(), sd_r

    ref32, _ = oracle(False)
    assert relmax(out.detach(), ref32) < 2e-2                   # north_star: bf16 logits within 2e-2 of the fp32 reference
    # gradients are sums of random-sign terms: a handful of ReLU-mask flips between fp32 and bf16 forward values moves
    # them by ~sqrt(flips / N) (10 % here), so the kernels are pinned against the oracle that rounds where they round
    ref16, sd_r = oracle(True)
    assert relmax(out.detach(), ref16) < 5e-3
    for k, p in net.patch_classifier.named_parameters():
        assert p.grad is not None, k
        assert relmax(p.grad, sd_r[k].grad) < 3e-2, k


def _sd_req(sd, prefix=''):
    return {k[len(prefix):]: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v)
            for k, v in sd.items() if k.startswith(prefix)}


ZERO_BIASES = ('0.bias', '1.bias', '4.bias', '5.bias')


def _check_grads(f, sd_r):
    """Train-mode BatchNorm removes any per-channel shift: the biases of the Linear layers in front of a BatchNorm (and of
    the Linear in front of those) have exactly zero gradient.  Both sides compute rounding residue there; it is compared on
    the scale of the matching weight gradient's row sums instead of its own."""
    for k, p in f.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all(), k
        r = sd_r[k].grad
        if k in ZERO_BIASES:
            scale = float(sd_r[k[:-4] + 'weight'].grad.abs().sum(1).max())
            assert float(p.grad.abs().max()) < 3e-2 * scale, (k, float(p.grad.abs().max()), scale)
        else:
            assert relmax(p.grad, r) < 3e-2, k


@pytest.mark.parametrize('B,G,H,W', [(2, 1000, 78, 64), (3, 64, 4, 4), (1, 333, 10, 8)])
def test_count_mlp_train_mode_bn_matches_oracle(B, G, H, W):
    """GridNetHexMM leaves the count f in TRAIN mode (training.py:126 only touches patch_classifier): batch-statistic
    BatchNorm1d, running-stat update and the full BatchNorm gradient on the tensor-core path."""
    from gridnext_b200.gridnet_models import GridNetHexOddr
    from gridnext_b200.count_mlp import compile_count_mlp
    n_cls = 7
    f = tutorial_mlp(G, n_cls)
    net = GridNetHexOddr(f, (G,), (H, W), n_cls, use_bn=True)
    sd = synth.synth_state_dict(S.gridnet_shapes(S.mlp_shapes(G, n_cls), n_cls, n_cls), 13)
    net.load_state_dict(sd)
    net.cuda()
    net.train()                                         # f stays in train mode
    assert compile_count_mlp(net.patch_classifier) is not None
    g = torch.Generator(); g.manual_seed(6)
    x = torch.log1p(torch.poisson(torch.ones(B, G, H, W), generator=g))
    dy = torch.randn(B, n_cls, H, W, generator=g)
    out = net.patch_predictions(x.cuda())
    (out * dy.cuda()).sum().backward()

    def oracle(emulate):
        sd_r = _sd_req(sd, 'patch_classifier.')
        stats = {}
        ref = R.grid_from_spots(R.mlp_forward(sd_r, R.spots_from_counts(x), training=True, stats_out=stats, emulate_bf16=emulate), B, H, W)
        (ref * dy).sum().backward()
        return ref.detach(), sd_r, stats

    ref32, _, stats32 = oracle(False)
    assert relmax(out.detach(), ref32) < 2e-2
    ref16, sd_r, stats = oracle(True)
    assert relmax(out.detach(), ref16) < 5e-3
    got = net.patch_classifier.state_dict()
    for k, v in stats32.items():
        assert relmax(got[k], v) < 5e-3, k
    assert int(got['2.num_batches_tracked']) == int(sd['patch_classifier.2.num_batches_tracked']) + 1
    _check_grads(net.patch_classifier, sd_r)


@pytest.mark.parametrize('train', [False, True])
def test_count_mlp_spot_major_batches(train):
    """f pre-training input (training.py:11-98): an (N, G) spot batch through count_mlp.forward_spots."""
    from gridnext_b200.count_mlp import compile_count_mlp
    G, n_cls, N = 200, 7, 96
    f = tutorial_mlp(G, n_cls)
    sd = synth.synth_state_dict(S.mlp_shapes(G, n_cls), 19)
    f.load_state_dict(sd)
    f.cuda().train(train)
    g = torch.Generator(); g.manual_seed(8)
    x = torch.log1p(torch.poisson(torch.ones(N, G), generator=g))
    dy = torch.randn(N, n_cls, generator=g)
    out = compile_count_mlp(f).forward_spots(x.cuda())
    (out * dy.cuda()).sum().backward()
    sd_r = _sd_req(sd)
    ref = R.mlp_forward(sd_r, x, training=train, emulate_bf16=True)
    (ref * dy).sum().backward()
    assert relmax(out.detach(), ref.detach()) < 5e-3
    if train:
        _check_grads(f, sd_r)
    else:
        for k, p in f.named_parameters():
            assert relmax(p.grad, sd_r[k].grad) < 3e-2, k


def test_train_spotwise_runs_count_mlp_and_densenet(capsys):
    """training.train_spotwise (training.py:11-98): two epochs on a separable toy problem, for both f families."""
    from gridnext_b200.training import train_spotwise
    from gridnext_b200.densenet import DenseNet
    from torch.utils.data import TensorDataset, DataLoader
    g = torch.Generator(); g.manual_seed(2)
    # count MLP
    y = torch.randint(0, 3, (128,), generator=g)
    x = torch.rand(128, 40, generator=g) + 2.0 * torch.nn.functional.one_hot(y, 40).float()
    dl = {'train': DataLoader(TensorDataset(x, y), batch_size=32), 'val': DataLoader(TensorDataset(x, y), batch_size=64)}
    torch.manual_seed(0)
    f = tutorial_mlp(40, 3)
    opt = torch.optim.Adam(f.parameters(), lr=3e-3)
    f, vh, th = train_spotwise(f, dl, nn.CrossEntropyLoss(), opt, num_epochs=6)
    # histories are per-epoch mean LOSSES (training.py:83-86); fp32 CPU run of the same problem: val 0.80 -> 0.014
    assert len(vh) == 6 and len(th) == 6 and vh[-1] < 0.2 and th[-1] < th[0], (vh, th)
    # DenseNet
    y = torch.randint(0, 2, (48,), generator=g)
    x = torch.randn(48, 3, 32, 32, generator=g) * 0.3 + (y.float() * 2 - 1).view(-1, 1, 1, 1)
    dl = {'train': DataLoader(TensorDataset(x, y), batch_size=16), 'val': DataLoader(TensorDataset(x, y), batch_size=48)}
    net = DenseNet(growth_rate=8, block_config=(2, 2), num_init_features=16, bn_size=2, num_classes=2, small_inputs=False)
    opt = torch.optim.Adam(net.parameters(), lr=1e-3)
    net, vh, th = train_spotwise(net, dl, nn.CrossEntropyLoss(), opt, num_epochs=4)
    # reference DenseNet on CPU, three seeds: val loss 0.44..0.73 -> 0.13..0.35 in four epochs
    assert vh[-1] < vh[0] and th[-1] < th[0] and vh[-1] < 0.65, (vh, th)
    capsys.readouterr()


def test_count_mlp_foreign_modules_use_the_generic_path():
    from gridnext_b200.count_mlp import compile_count_mlp
    f = tutorial_mlp(32, 7).cuda()
    f.train()
    assert compile_count_mlp(f) is not None
    f.eval()
    assert compile_count_mlp(f) is not None
    assert compile_count_mlp(nn.Sequential(nn.Linear(8, 8), nn.Tanh(), nn.Linear(8, 3))) is None
    assert compile_count_mlp(nn.Linear(8, 3)) is None
