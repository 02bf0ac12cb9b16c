"""Pins the CPU oracle (oracle/) against golden vectors produced by the REAL reference modules
(oracle/make_golden.py, run in the build container) and checks the three hexconv restatements
agree.  CPU only."""
import json, os
import numpy as np
import pytest
import torch

from oracle import synth, shapes as S
from oracle import gridnet_ref as R
from oracle import gather_ref
from oracle.hexconv_ref import (hexconv_visium, hexconv_hexagdly, hexconv_cube_bruteforce, kernel_shapes, n_taps)
from conftest import GOLDEN

MAN = json.load(open(os.path.join(GOLDEN, 'manifest.json')))


def t(x):
    return torch.from_numpy(np.asarray(x))


def make_sd(shapes, seed, grad=True):
    sd = synth.synth_state_dict(shapes, seed)
    if grad:
        for k, v in sd.items():
            if v.is_floating_point() and 'running' not in k and k not in ('bg_const', 'dummy_tensor'):
                v.requires_grad_(True)
    return sd


@pytest.mark.parametrize('k', [1, 2, 3])
@pytest.mark.parametrize('hw', [(9, 8), (8, 9), (5, 6)])
def test_hexconv_three_ways(k, hw):
    H, W = hw
    g = torch.Generator(); g.manual_seed(k * 100 + H)
    ks = [torch.randn(s, dtype=torch.float64, generator=g) for s in kernel_shapes(3, 4, k)]
    b = torch.randn(4, dtype=torch.float64, generator=g)
    x = torch.randn(2, 3, H, W, dtype=torch.float64, generator=g)
    y1 = hexconv_visium(x, ks, b)
    y2 = hexconv_hexagdly(x.transpose(2, 3).contiguous(), ks, b).transpose(2, 3)
    y3 = t(hexconv_cube_bruteforce(x.transpose(2, 3).numpy(), [q.numpy() for q in ks], b.numpy())).transpose(2, 3)
    assert (y1 - y2).abs().max() < 1e-12
    assert (y1 - y3).abs().max() < 1e-12
    assert n_taps(k) == 1 + 3 * k * (k + 1)


def test_hexconv_full_visium_grid_matches_composition():
    g = torch.Generator(); g.manual_seed(5)
    ks = [torch.randn(s, dtype=torch.float64, generator=g) for s in kernel_shapes(5, 6, 1)]
    x = torch.randn(1, 5, 78, 64, dtype=torch.float64, generator=g)
    y1 = hexconv_visium(x, ks, None)
    y2 = hexconv_hexagdly(x.transpose(2, 3).contiguous(), ks, None).transpose(2, 3)
    assert (y1 - y2).abs().max() < 1e-12


def test_state_dict_key_tables_match_reference():
    K = json.load(open(os.path.join(GOLDEN, 'state_dict_keys.json')))

    def same(a, b):
        return {k: tuple(v) for k, v in a.items()} == {k: tuple(v) for k, v in b.items()}
    assert same(K['d1_densenet121_p64'], S.densenet_shapes())
    assert same(K['d2_densenet_tiny_p32'], S.densenet_shapes(8, (2, 3), 16, 2))
    assert same(K['g1'], S.gridnet_shapes(S.mlp_shapes(40, 7), 7, 7))
    assert same(K['g2'], S.gridnet_shapes({'weight': (6, 4), 'bias': (6,)}, 6, 5, False))
    assert same(K['m1'], S.gridnet_mm_shapes(S.densenet_shapes(8, (2, 2), 16, 2), S.mlp_shapes(30, 7), 7, 7, 7))
    assert len(K['d1_densenet121_p64']) == 727


def test_count_gridnet_train_step_matches_reference(golden):
    m = MAN['g1_count_gridnet']
    gold = golden('g1_count_gridnet')
    sd = make_sd(S.gridnet_shapes(S.mlp_shapes(m['G'], m['n_cls']), m['n_cls'], m['n_cls']), m['seed_w'])
    x = synth.synth_counts(m['B'], m['G'], seed=m['seed_x'])
    y = synth.synth_labels(m['B'], m['n_cls'], seed=m['seed_y'])
    stats = {}
    out = R.gridnet_count_forward(sd, x, use_bn=True, training=True, stats_out=stats)
    loss, ncorr, nfg = R.masked_ce(out, y)
    loss.backward()
    assert np.allclose(out.detach().numpy(), gold['out'], rtol=1e-4, atol=1e-5)
    assert abs(float(loss.detach()) - float(gold['loss'])) < 1e-5
    assert (ncorr, nfg) == (int(gold['ncorr']), int(gold['nfg']))
    n = 0
    for k in gold.files:
        if k.startswith('grad.'):
            g = sd[k[5:]].grad.numpy()
            assert np.allclose(g, gold[k], rtol=2e-3, atol=2e-6), k
            n += 1
        if k.startswith('after.corrector.') and 'num_batches' not in k:
            assert np.allclose(stats[k[len('after.corrector.'):]].numpy(), gold[k], rtol=1e-5, atol=1e-6), k
    assert n >= 30
    ge = golden('g1_count_gridnet_eval')
    for k, v in stats.items():            # the reference's eval pass ran after the train step's BN update
        sd['corrector.' + k] = v
    with torch.no_grad():
        oe = R.gridnet_count_forward(sd, x, use_bn=True, training=False)
    assert np.allclose(oe.numpy(), ge['out'], rtol=1e-4, atol=1e-5)


def test_small_grid_nobn_matches_reference(golden):
    m = MAN['g2_small_nobn']
    gold = golden('g2_small_nobn')
    sd = make_sd(S.gridnet_shapes({'weight': (m['f_dim'], 4), 'bias': (m['f_dim'],)}, m['f_dim'], m['n_cls'], False), m['seed_w'])
    x, y = t(gold['x']), t(gold['y'])
    B, _, H, W = x.shape
    f = torch.nn.functional.linear(R.spots_from_counts(x), sd['patch_classifier.weight'], sd['patch_classifier.bias'])
    out = R.corrector_forward(R.sub(sd, 'corrector.'), R.grid_from_spots(f, B, H, W), use_bn=False)
    loss, ncorr, nfg = R.masked_ce(out, y)
    loss.backward()
    assert np.allclose(out.detach().numpy(), gold['out'], rtol=1e-4, atol=1e-5)
    assert abs(float(loss) - float(gold['loss'])) < 1e-5
    assert (ncorr, nfg) == (int(gold['ncorr']), int(gold['nfg']))
    for k in gold.files:
        if k.startswith('grad.'):
            assert np.allclose(sd[k[5:]].grad.numpy(), gold[k], rtol=1e-3, atol=1e-6), k


@pytest.mark.parametrize('tag', ['c1_cartesian_bn', 'c2_cartesian_nobn'])
def test_cartesian_gridnet_matches_reference(golden, tag):
    """Base GridNet (square-conv corrector, gridnet_models.py:51-66) through the reference's own class."""
    m = MAN[tag]
    gold = golden(tag)
    sd = make_sd(S.cartesian_gridnet_shapes({'weight': (m['f_dim'], 4), 'bias': (m['f_dim'],)}, m['f_dim'], m['n_cls'], m['use_bn']), m['seed_w'])
    assert set(sd) == set(json.load(open(os.path.join(GOLDEN, 'state_dict_keys.json')))[tag])
    x, y = t(gold['x']), t(gold['y'])
    B, H, W, _ = x.shape
    f = torch.nn.functional.linear(x.reshape(-1, 4), sd['patch_classifier.weight'], sd['patch_classifier.bias'])
    stats = {}
    out = R.cartesian_corrector_forward(R.sub(sd, 'corrector.'), R.grid_from_spots(f, B, H, W), use_bn=m['use_bn'], stats_out=stats)
    loss, ncorr, nfg = R.masked_ce(out, y)
    loss.backward()
    assert np.allclose(out.detach().numpy(), gold['out'], rtol=1e-4, atol=1e-5)
    assert abs(float(loss) - float(gold['loss'])) < 1e-5
    assert (ncorr, nfg) == (int(gold['ncorr']), int(gold['nfg']))
    for k in gold.files:
        if k.startswith('grad.'):
            assert np.allclose(sd[k[5:]].grad.numpy(), gold[k], rtol=1e-3, atol=1e-6), k
        elif k.startswith('after.corrector.'):
            assert np.allclose(stats[k[len('after.corrector.'):]].numpy(), gold[k], rtol=1e-5, atol=1e-6), k


@pytest.mark.parametrize('tag', ['d1_densenet121_p64', 'd2_densenet_tiny_p32'])
def test_densenet_matches_reference(golden, tag):
    m = MAN[tag]
    gold = golden(tag)
    sd = make_sd(S.densenet_shapes(m['growth_rate'], tuple(m['block_config']), m['num_init_features'], m['bn_size']), m['seed_w'])
    g = torch.Generator(); g.manual_seed(m['seed_x'])
    x = torch.randn(m['N'], 3, m['P'], m['P'], generator=g)
    logits = R.densenet_forward(sd, x)
    assert np.allclose(logits.detach().numpy(), gold['logits'], rtol=1e-4, atol=1e-5)
    g = torch.Generator(); g.manual_seed(m['seed_dy'])
    dy = torch.randn(logits.shape, generator=g)
    (logits * dy).sum().backward()
    norms = dict(zip(gold['grad_norm_keys'].tolist(), gold['grad_norm_vals'].tolist()))
    for k, v in norms.items():
        assert abs(float(sd[k].grad.norm()) - v) <= 2e-3 * max(v, 1e-3), k
    for k in gold.files:
        if k.startswith('grad.'):
            ref = gold[k]
            assert np.allclose(sd[k[5:]].grad.numpy(), ref, rtol=2e-3, atol=2e-3 * np.abs(ref).max()), k  # fp32 noise through 121 layers


def test_densenet_train_mode_matches_reference(golden):
    """Train-mode BatchNorm of f (training.py:11-98 pre-training): logits, every parameter gradient and the updated running
    statistics of the reference's DenseNet in .train()."""
    tag = 'd3_densenet_tiny_train'
    m = MAN[tag]
    gold = golden(tag)
    sd = make_sd(S.densenet_shapes(m['growth_rate'], tuple(m['block_config']), m['num_init_features'], m['bn_size']), m['seed_w'])
    g = torch.Generator(); g.manual_seed(m['seed_x'])
    x = torch.randn(m['N'], 3, m['P'], m['P'], generator=g)
    stats = {}
    logits = R.densenet_forward(sd, x, training=True, stats_out=stats)
    assert np.allclose(logits.detach().numpy(), gold['logits'], rtol=1e-4, atol=1e-5)
    g = torch.Generator(); g.manual_seed(m['seed_dy'])
    dy = torch.randn(logits.shape, generator=g)
    (logits * dy).sum().backward()
    n = 0
    for k in gold.files:
        if k.startswith('grad.'):
            ref = gold[k]
            assert np.allclose(sd[k[5:]].grad.numpy(), ref, rtol=2e-3, atol=2e-3 * np.abs(ref).max()), k
            n += 1
        elif k.startswith('after.') and 'running' in k:
            assert np.allclose(stats[k[6:]].numpy(), gold[k], rtol=1e-5, atol=1e-6), k
    assert n == len([k for k in sd if sd[k].requires_grad])


def test_multimodal_matches_reference(golden):
    m = MAN['m1_multimodal_4x4']
    gold = golden('m1_multimodal_4x4')
    assert m['ppred_shape'] == [2, 14, 4, 4] and m['out_shape'] == [2, 7, 4, 4]   # Tutorial_multimodal.ipynb:615-620
    sd = make_sd(S.gridnet_mm_shapes(S.densenet_shapes(8, (2, 2), 16, 2), S.mlp_shapes(m['Gc'], 7), 7, 7, 7), m['seed_w'])
    # reference aliases image_classifier.* and patch_classifier.* to the same tensors
    for k in list(sd):
        if k.startswith('patch_classifier.'):
            sd[k] = sd['image_classifier.' + k[len('patch_classifier.'):]]
    xi, xc, y = t(gold['xi']), t(gold['xc']), t(gold['y'])
    stats = {}
    out = R.gridnet_mm_forward(sd, xi, xc, training=True, stats_out=stats)
    assert tuple(out.shape) == (2, 7, 4, 4)
    assert np.allclose(out.detach().numpy(), gold['out'], rtol=1e-3, atol=1e-4)
    loss, ncorr, nfg = R.masked_ce(out, y)
    loss.backward()
    assert abs(float(loss) - float(gold['loss'])) < 1e-4
    assert nfg == int(gold['nfg'])
    n = 0
    for k in gold.files:
        if k.startswith('grad.'):   # patch_classifier.* aliases image_classifier.* (same tensors)
            ref = gold[k]
            got = sd[k[5:]].grad.numpy()
            assert np.allclose(got, ref, rtol=5e-3, atol=max(1e-3 * np.abs(ref).max(), 1e-7)), k  # pre-BN bias grads are exactly 0 in theory
            n += 1
        if k.startswith('after.') and 'num_batches' not in k:
            assert np.allclose(stats[k[6:]].numpy(), gold[k], rtol=1e-4, atol=1e-5), k
    assert n > 50


def test_gather_matches_reference(golden):
    m = MAN['p1_gather_p16']
    gold = golden('p1_gather_p16')
    tis, rows, cols, pr, pc = synth.synth_positions(pitch_col=m['pitch_col'], pitch_row=m['pitch_row'],
                                                    org_row=m['org_row'], org_col=m['org_col'])
    img = synth.synth_image(m['Himg'], m['Wimg'], seed=m['img_seed'], smooth=True)
    raw = gather_ref.grid_from_image(img, tis, rows, cols, pr, pc, patch_size=16, window_size=16)
    assert raw.dtype == np.float32
    assert np.array_equal(raw.astype(np.uint8), gold['raw']) and np.array_equal(raw, gold['raw'].astype(np.float32))
    nrm = gather_ref.grid_from_image(img, tis, rows, cols, pr, pc, patch_size=16, window_size=16,
                                     mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    assert np.array_equal(nrm[::11, ::9], gold['nrm_sub'])          # bit-exact float32
    assert np.all(nrm[~synth.tissue_mask()] == 0.0)                  # off-tissue stays exactly zero


def test_template_positions_index_math(golden):
    tp = golden('t1_template_positions')
    assert MAN['t1_template_positions'] == {'n': 4992, 'n_in_tissue': 4525}
    for r, c, xi, yi in zip(tp['array_row'][::37], tp['array_col'][::37], tp['x_ind'][::37], tp['y_ind'][::37]):
        assert gather_ref.pseudo_hex_to_oddr(int(c), int(r)) == (int(xi), int(yi))
    tab = gather_ref.spot_table(tp['in_tissue'], tp['array_row'], tp['array_col'], tp['pxl_row'], tp['pxl_col'])
    assert tab.shape == (4525, 4)
    assert tab[:, 0].max() == 63 and tab[:, 1].max() <= 77
    assert gather_ref.spot_table([1, 1, 1], [0, 0, 0], [0, 0, 0], [0.5, 1.5, 2.5], [2.5, 3.5, -0.5])[:, 2:].tolist() == [[2, 0], [4, 2], [0, 2]]


def test_dataset_assembly_oracle_matches_reference(golden):
    """oracle/datasets_ref.count_grid == the reference's read_annotated_starray run on synthetic files (utils.py:87-166)."""
    from oracle import datasets_ref as D
    gold = golden('a1_starray')
    cstrs = ['%d_%d' % (c, r) for c, r in gold['coords']]
    adict = dict(zip(gold['annot_coords'].tolist(), gold['annot_lbls'].tolist()))
    counts, annots = D.count_grid(gold['cmat'].astype(np.float64), cstrs, adict)
    assert np.array_equal(counts, gold['counts_annot'])
    assert np.array_equal(annots, gold['annots_annot'])
    # the rule of multimodal_datasets.py:237-244 on a hand-checkable case
    c = np.ones((2, 2, 2), np.float32); p = np.ones((2, 2, 3), np.float32); a = np.array([[1, 0], [2, 3]])
    p[1, 0] = 0                       # no image data -> loses label and counts
    c2, p2, a2 = D.mm_fg_consistency(c, p, a)
    assert a2.tolist() == [[1, 0], [0, 3]] and c2[:, 1, 0].tolist() == [0, 0] and c2[:, 0, 1].tolist() == [1, 1]
    assert p2[0, 1].tolist() == [0, 0, 0] and p2[1, 1].tolist() == [1, 1, 1]


def test_spot_cells_host_logic_matches_reference_indexing():
    """gridnext_b200.datasets.spot_cells (host side of the on-device assembly): pseudo-hex -> odd-r cells as utils.py:64-70,
    rint for Cartesian coordinates (utils.py:153-154), IndexError where indexing the reference's grid would raise."""
    from gridnext_b200.datasets import spot_cells
    from oracle import datasets_ref as D
    xs, ys = [0, 1, 3, 126, 127, 64], [0, 1, 77, 0, 77, 10]
    cells = spot_cells(xs, ys).tolist()
    for c, x, y in zip(cells, xs, ys):
        ox, oy = D.pseudo_hex_to_oddr(x, y)
        assert c == oy * 64 + ox
    assert spot_cells([1.4, 2.5, 3.5], [0.6, 1.5, 2.5], visium=False, h_st=9, w_st=11).tolist() == [1 * 11 + 1, 2 * 11 + 2, 2 * 11 + 4]
    with pytest.raises(IndexError):
        spot_cells([400], [3])


def test_patch_grid_oracle_matches_reference_dataset(golden):
    """oracle/datasets_ref.patch_grid == the reference's PatchGridDataset.__getitem__ (image_datasets.py:190-232) run on synthetic
    PNG patch files with Loupe annotations: ToTensor()'d patches placed at odd-r cells, labels + 1 (LabelEncoder order), 0 elsewhere."""
    from oracle import datasets_ref as D
    gold = golden('a2_patchgrid')
    classes = gold['classes'].tolist()
    assert classes == sorted(classes)                                   # LabelEncoder sorts the annotation names
    names = ['tumor', 'stroma', 'immune', 'necrosis']
    coords = [tuple(int(v) for v in c) for c in gold['coords']]
    adict = {'%d_%d' % c: classes.index(names[l]) for c, l, a in zip(coords, gold['label'], gold['annotated']) if a}
    patches = gold['patches'].transpose(0, 3, 1, 2).astype(np.float32) / 255.0          # ToTensor()
    grid, annots = D.patch_grid(patches, coords, adict)
    assert np.array_equal(annots, gold['annots_grid'])
    assert grid.shape == gold['patch_grid'].shape and np.array_equal(grid, gold['patch_grid'])


# ---- round 2: window_size != patch_size (Pillow BICUBIC resize inside the reference's gather loop) ---------------------------
def test_pillow_resize_restatement_equals_pil():
    """oracle.gather_ref.pillow_resize == PIL.Image.resize (default BICUBIC for RGB), bit for bit, up- and down-scaling."""
    from PIL import Image
    rng = np.random.default_rng(0)
    for w, P in [(256, 224), (64, 32), (100, 128), (128, 64), (30, 48), (48, 30), (254, 256), (20, 224), (24, 16), (10, 16)]:
        img = rng.integers(0, 256, (w, w, 3), dtype=np.uint8)
        assert np.array_equal(gather_ref.pillow_resize(img, P), np.array(Image.fromarray(img).resize((P, P)))), (w, P)


def test_product_resize_table_equals_oracle_table():
    """gridnext_b200.imgprocess.pillow_bicubic_table (what the CUDA kernel consumes) == the oracle's restatement."""
    from gridnext_b200.imgprocess import pillow_bicubic_table
    for w, P in [(256, 224), (24, 16), (10, 16), (20, 12), (512, 224), (300, 299)]:
        b, k, ksize, span = pillow_bicubic_table(w, P)
        bo, ko = gather_ref.pillow_coeffs(w, P)
        assert np.array_equal(b, bo) and np.array_equal(k, ko) and ksize == ko.shape[1] and span == int(bo[:, 1].max())


def test_gather_with_resize_matches_reference(golden):
    """grid_from_wsi_visium with window_size 24 / 10 / 0.03 (float) and patch sizes 16 / 16 / 12: the oracle against vectors
    produced by the reference function itself (oracle/make_golden.py --extras2)."""
    m = MAN['p2_gather_resize']
    gold = golden('p2_gather_resize')
    tis, rows, cols, pr, pc = synth.synth_positions(pitch_col=m['pitch_col'], pitch_row=m['pitch_row'], org_row=m['org_row'], org_col=m['org_col'])
    img = synth.synth_image(m['Himg'], m['Wimg'], seed=m['img_seed'], smooth=True)
    sel = np.ix_(gold['cells_y'], gold['cells_x'])
    for key, P, w in (('down_24_to_16', 16, 24), ('up_10_to_16', 16, 10), ('float_0.03_to_12', 12, 0.03)):
        got = gather_ref.grid_from_image(img, tis, rows, cols, pr, pc, patch_size=P, window_size=w)
        assert np.array_equal(got[sel], gold[key].astype(np.float32)), key
    nrm = gather_ref.grid_from_image(img, tis, rows, cols, pr, pc, patch_size=16, window_size=24, mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])
    assert np.array_equal(nrm[::11, ::9], gold['down_24_to_16_nrm_sub'])
