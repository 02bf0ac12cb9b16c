"""End-to-end GPU parity of the composite models and the training loop (rows a2, a5, a9, b of SURVEY.md section 8):
GridNetHexMM against vectors from the REAL reference, GridNetHexOddr with the DenseNet f against the oracle, train_gridwise
and all_fgd_predictions through the reference's own call signatures."""
import json, os
import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import synth, shapes as S
from oracle import gridnet_ref as R
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
MAN = json.load(open(os.path.join(GOLDEN, 'manifest.json')))


def relmax(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


def tutorial_mlp(G, n_cls):
    return nn.Sequential(nn.Linear(G, 500), nn.Linear(500, 100), nn.BatchNorm1d(100), nn.ReLU(),
                         nn.Linear(100, 100), nn.Linear(100, 50), nn.BatchNorm1d(50), nn.ReLU(), nn.Linear(50, n_cls))


def test_multimodal_4x4_matches_reference_golden():
    """Tutorial_multimodal.ipynb's dummy run: [count | image] concat (2, 14, 4, 4) -> (2, 7, 4, 4); the count f stays in
    train mode (training.py:126 only puts patch_classifier in eval), so it runs through the generic module path."""
    from gridnext_b200.gridnet_models import GridNetHexMM
    from gridnext_b200.densenet import DenseNet
    from gridnext_b200.training import gridwise_step
    m = MAN['m1_multimodal_4x4']
    gold = np.load(os.path.join(GOLDEN, 'm1_multimodal_4x4.npz'))
    fi = DenseNet(growth_rate=8, block_config=(2, 2), num_init_features=16, bn_size=2, num_classes=7, small_inputs=False)
    fc = tutorial_mlp(m['Gc'], 7)
    net = GridNetHexMM(fi, fc, (3, m['P'], m['P']), (m['Gc'],), (4, 4), 7)
    sd = synth.synth_state_dict(S.gridnet_mm_shapes(S.densenet_shapes(8, (2, 2), 16, 2), S.mlp_shapes(m['Gc'], 7), 7, 7, 7), m['seed_w'])
    for k in list(sd):
        if k.startswith('patch_classifier.'):
            sd[k] = sd['image_classifier.' + k[len('patch_classifier.'):]]
    assert set(net.state_dict().keys()) == set(sd.keys())
    net.load_state_dict(sd)
    net.cuda()
    net.train(); net.patch_classifier.eval()
    xi, xc, y = (torch.from_numpy(gold[k]).cuda() for k in ('xi', 'xc', 'y'))
    pp = net.patch_predictions([xi, xc])
    assert list(pp.shape) == m['ppred_shape']
    loss, acc, _ = gridwise_step(net, [xi, xc], y, nn.CrossEntropyLoss(), 1, True)
    out = net([xi, xc])
    assert list(out.shape) == m['out_shape']
    assert int(acc.tolist()[1]) == int(gold['nfg'])
    # 32 cells only: BatchNorm over 32 samples amplifies the bf16 error of the image f, hence the loose forward tolerance
    assert abs(float(loss) - float(gold['loss'])) < 5e-2 * max(1.0, abs(float(gold['loss'])))
    for k in gold.files:
        if k.startswith('grad.corrector.') or k.startswith('grad.count_classifier.'):
            p = dict(net.named_parameters())[k[5:]]
            assert p.grad is not None and torch.isfinite(p.grad).all(), k


def test_image_gridnet_matches_oracle():
    """GridNetHexOddr with a DenseNet f on a small grid: spot ordering n = b*H*W + y*W + x, f -> g composition, loss."""
    from gridnext_b200.gridnet_models import GridNetHexOddr
    from gridnext_b200.densenet import DenseNet
    from gridnext_b200.training import gridwise_step
    kw = dict(growth_rate=8, block_config=(2, 2), num_init_features=16, bn_size=2)
    B, H, W, P, n_cls = 2, 6, 8, 32, 5
    f = DenseNet(num_classes=n_cls, small_inputs=False, **kw)
    net = GridNetHexOddr(f, (3, P, P), (H, W), n_cls)
    sd = synth.synth_state_dict(S.gridnet_shapes(S.densenet_shapes(8, (2, 2), 16, 2, num_classes=n_cls), n_cls, n_cls), 23)
    assert set(net.state_dict().keys()) == set(sd.keys())
    net.load_state_dict(sd)
    net.cuda(); net.train(); net.patch_classifier.eval()
    g = torch.Generator(); g.manual_seed(6)
    x = torch.randn(B, H, W, 3, P, P, generator=g)
    y = torch.randint(0, n_cls + 1, (B, H, W), generator=g)
    sd_r = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k and k not in ('bg_const', 'dummy_tensor') else v)
            for k, v in sd.items()}
    fr = R.densenet_forward(R.sub(sd_r, 'patch_classifier.'), x.reshape(-1, 3, P, P), emulate_bf16=True)
    out_r = R.corrector_forward(R.sub(sd_r, 'corrector.'), R.grid_from_spots(fr, B, H, W), True, True)
    loss_r, ncorr, nfg = R.masked_ce(out_r, y)
    loss_r.backward()
    pp = net.patch_predictions(x.cuda())
    assert relmax(pp.detach(), R.grid_from_spots(fr, B, H, W).detach()) < 5e-3          # spot order + f
    loss, acc, _ = gridwise_step(net, x.cuda(), y.cuda(), nn.CrossEntropyLoss(), 1, True)
    assert int(acc.tolist()[1]) == nfg
    assert abs(float(loss) - float(loss_r)) < 2e-2 * max(1.0, abs(float(loss_r)))
    for name in ('corrector.8.kernel0', 'corrector.0.kernel1', 'patch_classifier.classifier.weight'):
        gp = dict(net.named_parameters())[name].grad
        assert relmax(gp, sd_r[name].grad) < 0.1, name


class _Arrays(torch.utils.data.Dataset):
    def __init__(self, n, G, seed):
        self.x = synth.synth_counts(n, G, seed=seed)
        self.y = synth.synth_labels(n, 7, seed=seed + 1)

    def __len__(self):
        return len(self.x)

    def __getitem__(self, i):
        return self.x[i], self.y[i]


def test_train_gridwise_loop_and_fgd_predictions(tmp_path):
    """train_gridwise(model, dataloaders, criterion, optimizer, num_epochs, outfile, f_opt, accum_iters) -> (model, val_hist,
    train_hist) (training.py:101-102,209): loss goes down on a learnable toy problem, checkpoints are written, and
    utils.all_fgd_predictions returns the flattened foreground labels / predictions / softmax."""
    from gridnext_b200.gridnet_models import GridNetHexOddr
    from gridnext_b200.training import train_gridwise
    from gridnext_b200.utils import all_fgd_predictions
    torch.manual_seed(0)
    G = 24
    f = tutorial_mlp(G, 7)
    net = GridNetHexOddr(f, (G,), (78, 64), 7)
    train, val = _Arrays(4, G, 1), _Arrays(2, G, 7)
    # make the problem learnable: inject the label into the first genes of every foreground spot
    for ds in (train, val):
        for i in range(len(ds)):
            for c in range(7):
                ds.x[i][c][ds.y[i] == c + 1] += 3.0
    dls = {'train': torch.utils.data.DataLoader(train, batch_size=2), 'val': torch.utils.data.DataLoader(val, batch_size=2)}
    opt = torch.optim.Adam(net.parameters(), lr=3e-3)
    out = str(tmp_path / 'g.pth')
    model, val_hist, train_hist = train_gridwise(net, dls, nn.CrossEntropyLoss(), opt, num_epochs=4, outfile=out, accum_iters=1)
    assert len(val_hist) == 4 and len(train_hist) == 4 and all(np.isfinite(val_hist)) and all(np.isfinite(train_hist))
    assert train_hist[-1] < train_hist[0]
    assert os.path.exists(out) and os.path.exists(str(tmp_path / 'g.opt'))
    assert set(torch.load(out).keys()) == set(model.state_dict().keys())
    true, pred, smax = all_fgd_predictions(dls['val'], model)
    n_fg = int(sum((val.y[i] > 0).sum() for i in range(len(val))))
    assert true.shape == (n_fg,) and pred.shape == (n_fg,) and smax.shape == (n_fg, 7)
    assert np.allclose(smax.sum(1), 1.0, atol=1e-5) and true.min() >= 0 and true.max() <= 6
