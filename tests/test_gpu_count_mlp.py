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


def test_count_mlp_train_mode_bn_and_foreign_modules_use_the_generic_path():
    """GridNetHexMM leaves the count f in train mode (training.py:126 only touches patch_classifier): not compiled."""
    from gridnext_b200.count_mlp import compile_count_mlp
    f = tutorial_mlp(32, 7).cuda()
    f.train()
    assert compile_count_mlp(f) is None
    f.eval()
    assert compile_count_mlp(f) is not None
    assert compile_count_mlp(nn.Sequential(nn.Linear(8, 8), nn.Tanh(), nn.Linear(8, 3))) is None
    assert compile_count_mlp(nn.Linear(8, 3)) is None
