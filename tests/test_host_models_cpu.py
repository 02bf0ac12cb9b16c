"""Host-side model logic on CPU against the REAL reference classes (only where /root/reference is mounted: the build container).
No kernel runs here: the base GridNet with CPU tensors applies its corrector module by module, and patch_predictions of the
hex classes is plain tensor plumbing around user-supplied f modules."""
import os
import sys
import types
import warnings

import pytest
import torch
import torch.nn as nn

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'gridnext')), reason='reference tree not mounted')


@pytest.fixture(scope='module')
def ref_models():
    warnings.filterwarnings('ignore', category=SyntaxWarning)
    from oracle import hexagdly_shim
    sys.modules.setdefault('hexagdly', hexagdly_shim)
    for name in ('matplotlib', 'matplotlib.pyplot', 'mpl_toolkits', 'mpl_toolkits.axes_grid1'):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['mpl_toolkits.axes_grid1'].make_axes_locatable = lambda *a, **k: None
    if REF not in sys.path:
        sys.path.insert(0, REF)
    import gridnext.gridnet_models as gm
    return gm


def _sync(dst, src):
    dst.load_state_dict(src.state_dict(), strict=True)          # identical key sets are part of the contract


@pytest.mark.parametrize('limit', [None, 7])
def test_cartesian_gridnet_cpu_matches_reference(ref_models, limit):
    from gridnext_b200.gridnet_models import GridNet
    torch.manual_seed(3)
    H, W, n_cls, f_dim = 6, 5, 4, 5
    ours = GridNet(nn.Linear(9, f_dim), (9,), (H, W), n_cls, use_bn=True, atonce_patch_limit=limit, f_dim=f_dim)
    ref = ref_models.GridNet(nn.Linear(9, f_dim), (9,), (H, W), n_cls, use_bn=True, atonce_patch_limit=limit, f_dim=f_dim)
    _sync(ours, ref)
    x = torch.randn(3, H, W, 9)
    for m in (ours, ref):
        m.train()
    yo, yr = ours(x), ref(x)
    assert torch.allclose(yo, yr, atol=1e-6)
    yo.square().sum().backward(); yr.square().sum().backward()
    for (k, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
        assert torch.allclose(p.grad, q.grad, atol=1e-5), k
    assert torch.allclose(ours.state_dict()['corrector.1.running_mean'], ref.state_dict()['corrector.1.running_mean'])


def test_multimodal_patch_predictions_cpu_match_reference(ref_models):
    from gridnext_b200.gridnet_models import GridNetHexMM

    def nets():
        fi = nn.Sequential(nn.Flatten(), nn.Linear(3 * 4 * 4, 6))
        fc = nn.Sequential(nn.Linear(10, 8), nn.ReLU(), nn.Linear(8, 3))
        return fi, fc

    for limit in (None, 5):
        torch.manual_seed(5)
        fi, fc = nets()
        ref = ref_models.GridNetHexMM(fi, fc, (3, 4, 4), (10,), (4, 6), 5, use_bn=True, atonce_patch_limit=limit, image_f_dim=6, count_f_dim=3)
        fi2, fc2 = nets()
        ours = GridNetHexMM(fi2, fc2, (3, 4, 4), (10,), (4, 6), 5, use_bn=True, atonce_patch_limit=limit, image_f_dim=6, count_f_dim=3)
        assert set(ours.state_dict()) == set(ref.state_dict())
        _sync(ours, ref)
        xi, xc = torch.randn(2, 4, 6, 3, 4, 4), torch.randn(2, 10, 4, 6)
        po, pr = ours.patch_predictions([xi, xc]), ref.patch_predictions([xi, xc])
        assert tuple(po.shape) == (2, 9, 4, 6) and torch.allclose(po, pr, atol=1e-6)
        assert ours.f_dim == ref.f_dim == 9 and ours.patch_shape == ref.patch_shape
        if limit is None:
            po.sum().backward(); pr.sum().backward()
            for (k, p), (_, q) in zip(ours.named_parameters(), ref.named_parameters()):
                if q.grad is not None:
                    assert torch.allclose(p.grad, q.grad, atol=1e-5), k
        else:
            # reference quirk kept: the re-entrant checkpoint of gridnet_models.py:95-99 re-runs self._ppl in backward, when
            # _set_mode has already pointed patch_classifier at the other modality -> both implementations raise the same error
            for out in (po, pr):
                with pytest.raises(RuntimeError, match='shapes cannot be multiplied'):
                    out.sum().backward()


@pytest.mark.parametrize('accum_iters', [1, 2])
def test_train_gridwise_loop_cpu_matches_reference(ref_models, accum_iters, capsys):
    """training.train_gridwise against the reference's own loop (training.py:101-209) on CPU: same accumulate/step cadence
    (step at batch_ind % accum_iters == 0, :162-169), same histories, same best-weights bookkeeping, same final parameters."""
    import copy
    import gridnext.training as ref_tr
    from torch.utils.data import TensorDataset, DataLoader
    from gridnext_b200.gridnet_models import GridNet
    from gridnext_b200.training import train_gridwise
    torch.manual_seed(11)
    H, W, n_cls, f_dim = 5, 4, 3, 3
    x = torch.randn(6, H, W, 7)
    y = torch.randint(0, n_cls + 1, (6, H, W))
    dl = {'train': DataLoader(TensorDataset(x, y), batch_size=2), 'val': DataLoader(TensorDataset(x[:4], y[:4]), batch_size=2)}
    ref = ref_models.GridNet(nn.Linear(7, f_dim), (7,), (H, W), n_cls, use_bn=True, f_dim=f_dim)
    ours = GridNet(nn.Linear(7, f_dim), (7,), (H, W), n_cls, use_bn=True, f_dim=f_dim)
    _sync(ours, ref)
    ref_dev = ref_tr.device if hasattr(ref_tr, 'device') else None
    results = []
    for model, fn in ((ours, train_gridwise), (ref, ref_tr.train_gridwise)):
        opt = torch.optim.SGD(model.parameters(), lr=0.05)
        m, vh, th = fn(model, dl, nn.CrossEntropyLoss(), opt, num_epochs=3, accum_iters=accum_iters)
        results.append((copy.deepcopy(m.state_dict()), vh, th))
    capsys.readouterr()
    (sd_o, vh_o, th_o), (sd_r, vh_r, th_r) = results
    assert len(vh_o) == len(vh_r) and len(th_o) == len(th_r)
    assert all(abs(float(a) - float(b)) < 1e-6 for a, b in zip(vh_o, vh_r))
    assert all(abs(float(a) - float(b)) < 1e-6 for a, b in zip(th_o, th_r))
    for k in sd_r:
        assert torch.allclose(sd_o[k].float().cpu(), sd_r[k].float().cpu(), atol=1e-6), k


def test_train_spotwise_loop_cpu_matches_reference(ref_models, capsys):
    """training.train_spotwise against the reference's loop (training.py:11-98) on CPU: histories, best-accuracy bookkeeping and
    the returned (best) weights."""
    import copy
    import gridnext.training as ref_tr
    from torch.utils.data import TensorDataset, DataLoader
    from gridnext_b200.training import train_spotwise
    torch.manual_seed(13)
    x = torch.randn(48, 12)
    y = (x[:, 0] > 0).long() + (x[:, 1] > 0.5).long()
    dl = {'train': DataLoader(TensorDataset(x, y), batch_size=16), 'val': DataLoader(TensorDataset(x[:32], y[:32]), batch_size=16)}

    def net():
        torch.manual_seed(17)
        return nn.Sequential(nn.Linear(12, 10), nn.BatchNorm1d(10), nn.ReLU(), nn.Linear(10, 3))

    results = []
    for fn in (train_spotwise, ref_tr.train_spotwise):
        model = net()
        opt = torch.optim.Adam(model.parameters(), lr=1e-2)
        m, vh, th = fn(model, dl, nn.CrossEntropyLoss(), opt, num_epochs=4)
        results.append((copy.deepcopy(m.state_dict()), [float(v) for v in vh], [float(v) for v in th]))
    capsys.readouterr()
    (sd_o, vh_o, th_o), (sd_r, vh_r, th_r) = results
    assert vh_o == pytest.approx(vh_r, abs=1e-6) and th_o == pytest.approx(th_r, abs=1e-6)
    for k in sd_r:
        assert torch.allclose(sd_o[k].float(), sd_r[k].float().cpu(), atol=1e-6), k


@pytest.mark.parametrize('f_only', [False, True])
def test_all_fgd_predictions_cpu_matches_reference(ref_models, f_only):
    """utils.all_fgd_predictions against the reference's evaluation loop (utils.py:20-57) on CPU, and the index helpers."""
    import gridnext.utils as ref_ut
    from torch.utils.data import TensorDataset, DataLoader
    from gridnext_b200 import utils as our_ut
    from gridnext_b200.gridnet_models import GridNet
    torch.manual_seed(21)
    H, W, n_cls = 5, 4, 3
    x = torch.randn(5, H, W, 7)
    y = torch.randint(0, n_cls + 1, (5, H, W))
    dl = DataLoader(TensorDataset(x, y), batch_size=2)
    ref = ref_models.GridNet(nn.Linear(7, n_cls), (7,), (H, W), n_cls, use_bn=True)
    ours = GridNet(nn.Linear(7, n_cls), (7,), (H, W), n_cls, use_bn=True)
    _sync(ours, ref)
    got = our_ut.all_fgd_predictions(dl, ours, f_only=f_only)
    exp = ref_ut.all_fgd_predictions(dl, ref, f_only=f_only)
    assert (got[0] == exp[0]).all() and (got[1] == exp[1]).all()
    assert got[2].shape == exp[2].shape and abs(got[2] - exp[2]).max() < 1e-6
    for col, row in [(0, 0), (1, 1), (127, 77), (64, 10), (3, 5)]:
        assert our_ut.pseudo_hex_to_oddr(col, row) == ref_ut.pseudo_hex_to_oddr(col, row)
        ox, oy = our_ut.pseudo_hex_to_oddr(col, row)
        assert our_ut.oddr_to_pseudo_hex(ox, oy) == ref_ut.oddr_to_pseudo_hex(ox, oy)
        assert our_ut.pseudo_to_true_hex(col, row) == ref_ut.pseudo_to_true_hex(col, row)


def test_position_file_reader_and_window_rule_match_reference(ref_models, tmp_path):
    """imgprocess.read_positions / _window (host side of grid_from_wsi_visium) against utils.visium_get_positions
    (utils.py:246-287) and the window rule of imgprocess.py:188-195, for both Spaceranger position-file generations."""
    import numpy as np
    import gridnext.utils as ref_ut
    from gridnext_b200 import imgprocess as ours
    rows = [('AAAC-1', 1, 0, 0, 1200, 1300), ('AAAG-1', 0, 1, 1, 1397, 1413), ('AAAT-1', 1, 77, 127, 16000, 15800)]
    v2 = tmp_path / 'run_v2' / 'outs' / 'spatial'; v2.mkdir(parents=True)
    with open(v2 / 'tissue_positions.csv', 'w') as fh:
        fh.write('barcode,in_tissue,array_row,array_col,pxl_row_in_fullres,pxl_col_in_fullres\n')
        fh.writelines('%s,%d,%d,%d,%d,%d\n' % r for r in rows)
    v1 = tmp_path / 'run_v1' / 'spatial'; v1.mkdir(parents=True)
    with open(v1 / 'tissue_positions_list.csv', 'w') as fh:
        fh.writelines('%s,%d,%d,%d,%d,%d\n' % r for r in rows)
    for d in (tmp_path / 'run_v2', tmp_path / 'run_v1'):
        got, exp = ours.read_positions(str(d)), ref_ut.visium_get_positions(str(d))
        for k in ('in_tissue', 'array_row', 'array_col', 'pxl_row_in_fullres', 'pxl_col_in_fullres'):
            assert np.array_equal(got[k], exp[k].values), (d, k)
    (tmp_path / 'empty').mkdir()
    for fn in (ours.read_positions, ref_ut.visium_get_positions):
        with pytest.raises(ValueError):
            fn(str(tmp_path / 'empty'))
    import gridnext.imgprocess as ref_ip
    for c in [(0, 0), (3, 1), (127, 77)]:
        assert ours.pseudo_hex_to_cartesian(c) == ref_ip.pseudo_hex_to_cartesian(c)
        assert ours.pseudo_hex_to_oddr(*c) == ref_ip.pseudo_hex_to_oddr(*c) and ours.oddr_to_pseudo_hex(*c) == ref_ip.oddr_to_pseudo_hex(*c)
    # window rule (imgprocess.py:188-195): None -> patch size, float -> fraction of the image WIDTH, int -> pixels, else ValueError
    assert ours._window(224, None, 1000) == 224 and ours._window(224, 0.1, 1000) == 100 and ours._window(224, 96, 1000) == 96
    with pytest.raises(ValueError):
        ours._window(224, '96', 1000)


@pytest.mark.parametrize('kw', [dict(growth_rate=8, block_config=(2, 3), num_init_features=16, bn_size=2, num_classes=5, small_inputs=False),
                                dict(growth_rate=32, block_config=(6, 12, 24, 16), num_init_features=64, bn_size=4, num_classes=7, small_inputs=False),
                                dict(growth_rate=12, block_config=(3, 3, 3), num_init_features=24, bn_size=4, num_classes=10, small_inputs=True,
                                     compression=0.5)])
def test_densenet_constructor_and_init_match_reference(ref_models, kw):
    """DenseNet(...) builds the same module tree (state-dict keys, shapes) and -- from the same RNG state -- the same initial
    parameters as the reference constructor (densenet.py:93-150: He-normal convolutions, unit BatchNorm, zero classifier bias)."""
    import gridnext.densenet as ref_dn
    from gridnext_b200.densenet import DenseNet
    torch.manual_seed(1234)
    ref = ref_dn.DenseNet(**kw)
    torch.manual_seed(1234)
    ours = DenseNet(**kw)
    sr, so = ref.state_dict(), ours.state_dict()
    assert list(sr.keys()) == list(so.keys())
    for k in sr:
        assert sr[k].shape == so[k].shape and torch.equal(sr[k], so[k]), k
    assert [n for n, _ in ours.named_modules()] == [n for n, _ in ref.named_modules()]


@pytest.mark.parametrize('use_bn', [True, False])
def test_cartesian_gridnet_constructor_matches_reference(ref_models, use_bn):
    """Same RNG state -> the same freshly initialised GridNet (key order, buffers bg_const / dummy_tensor, corrector weights),
    and init_weights() applied through .apply() leaves both in the same state (gridnet_models.py:14-20)."""
    from gridnext_b200 import gridnet_models as ours_gm

    def build(gm):
        torch.manual_seed(77)
        return gm.GridNet(nn.Linear(6, 5), (6,), (7, 9), 4, use_bn=use_bn, atonce_patch_limit=3, f_dim=5)

    ref, ours = build(ref_models), build(ours_gm)
    sr, so = ref.state_dict(), ours.state_dict()
    assert list(sr.keys()) == list(so.keys())
    for k in sr:
        assert torch.equal(sr[k], so[k]), k
    for attr in ('patch_shape', 'grid_shape', 'n_classes', 'use_bn', 'atonce_patch_limit', 'f_dim'):
        assert getattr(ours, attr) == getattr(ref, attr), attr
    torch.manual_seed(5); ref.apply(ref_models.init_weights)
    torch.manual_seed(5); ours.apply(ours_gm.init_weights)
    for k, v in ref.state_dict().items():
        assert torch.equal(v, ours.state_dict()[k]), k
