"""Generate tests/golden/*.npz from the REAL reference modules (run in the build container only).

    python -m oracle.make_golden

Imports /root/reference/gridnext/{gridnet_models,densenet,training,imgprocess,utils}.py unmodified,
with ``hexagdly`` replaced by oracle/hexagdly_shim.py (un-vendored dependency, PARITY UNPINNED) and a
matplotlib stub (gridnext/utils.py:8-10 imports it at module load).  Weights come from
oracle/synth.py (key-hashed, independent of construction order) and are pushed into the reference
modules with load_state_dict.  Only inputs that cannot be regenerated and OUTPUTS are stored.
"""
import os, sys, types, warnings, tempfile, json
import numpy as np
import torch
import torch.nn as nn

REF = '/root/reference'
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')


def import_reference():
    warnings.filterwarnings('ignore', category=SyntaxWarning)
    from oracle import hexagdly_shim
    sys.modules['hexagdly'] = hexagdly_shim
    for name in ('matplotlib', 'matplotlib.pyplot', 'mpl_toolkits', 'mpl_toolkits.axes_grid1'):
        m = types.ModuleType(name)
        sys.modules.setdefault(name, m)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']
    sys.modules['mpl_toolkits.axes_grid1'].make_axes_locatable = lambda *a, **k: None
    sys.path.insert(0, REF)
    import gridnext.gridnet_models as gm
    import gridnext.densenet as dn
    import gridnext.training as tr
    import gridnext.imgprocess as ip
    return gm, dn, tr, ip


def count_mlp(G, n_cls):
    # notebooks/Tutorial_visium_count.ipynb cell 12
    return nn.Sequential(nn.Linear(G, 500), nn.Linear(500, 100), nn.BatchNorm1d(100), nn.ReLU(),
                         nn.Linear(100, 100), nn.Linear(100, 50), nn.BatchNorm1d(50), nn.ReLU(),
                         nn.Linear(50, n_cls))


KEYS = {}


def load_synth(model, seed=1234, tag=None):
    from oracle.synth import synth_state_dict, shapes_of
    if tag is not None:
        KEYS[tag] = {k: list(v) for k, v in shapes_of(model).items()}
    sd = synth_state_dict(shapes_of(model), seed)
    model.load_state_dict(sd)
    return sd


def grads_of(model):
    return {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}


def step_like_train_gridwise(model, inputs, labels):
    """One train-phase iteration of training.py:119-171 (no optimizer step)."""
    model.train()
    model.patch_classifier.eval()
    outputs = model(inputs)
    o = outputs.permute((0, 2, 3, 1))
    o = torch.reshape(o, (-1, o.shape[-1]))
    l = torch.reshape(labels, (-1,))
    o = o[l > 0]
    l = l[l > 0] - 1
    loss = nn.CrossEntropyLoss()(o, l)
    _, preds = torch.max(o, 1)
    loss.backward()
    return outputs.detach(), loss.detach(), int(torch.sum(preds == l)), len(l)


def np_(d):
    return {k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in d.items()}


def main():
    os.makedirs(OUT, exist_ok=True)
    gm, dn, tr, ip = import_reference()
    from oracle import synth
    torch.set_num_threads(os.cpu_count())
    manifest = {}

    # ---- G1: hexagdly-layout corrector inside GridNetHexOddr, count f, small G, full 78x64 grid
    G, n_cls, B = 40, 7, 2
    model = gm.GridNetHexOddr(count_mlp(G, n_cls), (G,), (78, 64), n_cls, use_bn=True)
    load_synth(model, 11, 'g1')
    x = synth.synth_counts(B, G, seed=3)
    y = synth.synth_labels(B, n_cls, seed=5)
    out, loss, ncorr, nfg = step_like_train_gridwise(model, x, y)
    g = grads_of(model)
    sd_after = {k: v for k, v in model.state_dict().items() if 'running' in k and k.startswith('corrector')}
    np.savez_compressed(os.path.join(OUT, 'g1_count_gridnet.npz'), out=out.numpy(), loss=loss.numpy(),
                        ncorr=ncorr, nfg=nfg, **{'grad.' + k: v.numpy() for k, v in g.items()},
                        **{'after.' + k: v.numpy() for k, v in sd_after.items()})
    manifest['g1_count_gridnet'] = dict(G=G, n_cls=n_cls, B=B, seed_w=11, seed_x=3, seed_y=5)
    # eval-mode forward (running stats)
    model.eval()
    with torch.no_grad():
        out_eval = model(x)
        pp = model.patch_predictions(x)
    np.savez_compressed(os.path.join(OUT, 'g1_count_gridnet_eval.npz'), out=out_eval.numpy(), ppred=pp.numpy())

    # ---- G2: small odd grid, use_bn False, f_dim != n_classes, identity-ish f
    n_cls, f_dim, B, H, W = 5, 6, 3, 7, 9
    f = nn.Linear(4, f_dim)
    model = gm.GridNetHexOddr(f, (4,), (H, W), n_cls, use_bn=False, f_dim=f_dim)
    load_synth(model, 12, 'g2')
    gx = torch.Generator(); gx.manual_seed(21)
    x = torch.randn(B, 4, H, W, generator=gx)
    y = synth.synth_labels(B, n_cls, H, W, seed=9)
    out, loss, ncorr, nfg = step_like_train_gridwise(model, x, y)
    g = grads_of(model)
    np.savez_compressed(os.path.join(OUT, 'g2_small_nobn.npz'), x=x.numpy(), y=y.numpy(), out=out.numpy(), loss=loss.numpy(),
                        ncorr=ncorr, nfg=nfg, **{'grad.' + k: v.numpy() for k, v in g.items()})
    manifest['g2_small_nobn'] = dict(n_cls=n_cls, f_dim=f_dim, B=B, H=H, W=W, seed_w=12)

    # ---- D1: DenseNet-121 @ 64 px, 2 spots, logits + selected grads; D2: tiny DenseNet @ 32 px
    for tag, kw, P, N, seed in (
            ('d1_densenet121_p64', dict(growth_rate=32, block_config=(6, 12, 24, 16), num_init_features=64, bn_size=4), 64, 2, 31),
            ('d2_densenet_tiny_p32', dict(growth_rate=8, block_config=(2, 3), num_init_features=16, bn_size=2), 32, 3, 32)):
        net = dn.DenseNet(num_classes=7, small_inputs=False, efficient=False, drop_rate=0, **kw)
        load_synth(net, seed, tag)
        net.eval()
        gx = torch.Generator(); gx.manual_seed(seed + 100)
        x = torch.randn(N, 3, P, P, generator=gx)
        logits = net(x)
        gy = torch.Generator(); gy.manual_seed(seed + 200)
        dy = torch.randn(logits.shape, generator=gy)
        (logits * dy).sum().backward()
        g = grads_of(net)
        keep = {k: v for k, v in g.items() if v.numel() <= 40000 and (
            'conv0' in k or 'denselayer1.' in k or 'denselayer2.' in k or 'transition1' in k or 'norm_final' in k
            or 'classifier' in k or k.startswith('features.norm0') or 'denseblock4.denselayer16' in k)}
        norms = {k: float(v.norm()) for k, v in g.items()}
        np.savez_compressed(os.path.join(OUT, tag + '.npz'), logits=logits.detach().numpy(),
                            grad_norm_keys=np.array(list(norms.keys())), grad_norm_vals=np.array(list(norms.values())),
                            **{'grad.' + k: v.numpy() for k, v in keep.items()})
        manifest[tag] = dict(P=P, N=N, seed_w=seed, seed_x=seed + 100, seed_dy=seed + 200, **{k: list(v) if isinstance(v, tuple) else v for k, v in kw.items()})

    # ---- M1: multimodal shape KAT (Tutorial_multimodal.ipynb:615-620) + values, 4x4 grid
    n_cls, Gc, P = 7, 30, 32
    fi = dn.DenseNet(num_classes=n_cls, small_inputs=False, efficient=False, growth_rate=8, block_config=(2, 2),
                     num_init_features=16, bn_size=2, drop_rate=0)
    fc = count_mlp(Gc, n_cls)
    model = gm.GridNetHexMM(fi, fc, (3, P, P), (Gc,), (4, 4), n_cls, use_bn=True, atonce_patch_limit=None)
    load_synth(model, 41, 'm1')
    gx = torch.Generator(); gx.manual_seed(141)
    xi = torch.rand(2, 4, 4, 3, P, P, generator=gx)
    xc = torch.rand(2, Gc, 4, 4, generator=gx)
    y = torch.randint(0, n_cls + 1, (2, 4, 4), generator=gx)
    model.train(); model.patch_classifier.eval()
    with torch.no_grad():
        pp = model.patch_predictions([xi, xc])
    load_synth(model, 41)   # undo count-BN running-stat update from the probe above
    out, loss, ncorr, nfg = step_like_train_gridwise(model, [xi, xc], y)
    g = grads_of(model)
    after = {k: v for k, v in model.state_dict().items() if 'running' in k and not k.startswith('patch_classifier') and not k.startswith('image_classifier')}
    np.savez_compressed(os.path.join(OUT, 'm1_multimodal_4x4.npz'), xi=xi.numpy(), xc=xc.numpy(), y=y.numpy(),
                        ppred=pp.numpy(), out=out.numpy(), loss=loss.numpy(), ncorr=ncorr, nfg=nfg,
                        **{'grad.' + k: v.numpy() for k, v in g.items() if v.numel() <= 20000},
                        **{'after.' + k: v.numpy() for k, v in after.items()})
    manifest['m1_multimodal_4x4'] = dict(n_cls=n_cls, Gc=Gc, P=P, seed_w=41, ppred_shape=list(pp.shape), out_shape=list(out.shape))

    # ---- P1: patch gather through the real grid_from_wsi_visium (PNG on disk + positions csv)
    from PIL import Image
    from torchvision import transforms
    tis, rows, cols, pr, pc = synth.synth_positions(pitch_col=5.5, pitch_row=9.6, org_row=3.0, org_col=2.0)
    Himg, Wimg = 790, 730     # spots close to all four borders -> edge padding exercised
    img = synth.synth_image(Himg, Wimg, seed=7, smooth=True)
    with tempfile.TemporaryDirectory() as td:
        Image.fromarray(img).save(os.path.join(td, 'img.png'))
        sp = os.path.join(td, 'outs', 'spatial'); os.makedirs(sp)
        with open(os.path.join(sp, 'tissue_positions.csv'), 'w') as fh:
            fh.write('barcode,in_tissue,array_row,array_col,pxl_row_in_fullres,pxl_col_in_fullres\n')
            for i in range(len(tis)):
                fh.write('BC%05d-1,%d,%d,%d,%r,%r\n' % (i, tis[i], rows[i], cols[i], float(pr[i]), float(pc[i])))
        try:
            raw = ip.grid_from_wsi_visium(os.path.join(td, 'img.png'), os.path.join(td, 'outs'), patch_size=16, window_size=16)
            nrm = ip.grid_from_wsi_visium(os.path.join(td, 'img.png'), os.path.join(td, 'outs'), patch_size=16, window_size=16,
                                          preprocess_xform=transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]))
        except Exception as e:   # directory layout helper differs between spaceranger versions
            raise
    np.savez_compressed(os.path.join(OUT, 'p1_gather_p16.npz'), raw=raw.numpy().astype(np.uint8), nrm_sub=nrm.numpy()[::11, ::9])  # nrm subsampled: cells [::11, ::9]
    manifest['p1_gather_p16'] = dict(Himg=Himg, Wimg=Wimg, P=16, pitch_col=5.5, pitch_row=9.6, org_row=3.0, org_col=2.0, img_seed=7)

    # ---- T1: template-derived tissue mask & pixel coordinates (visium_templates/tissue_positions.csv)
    import pandas as pd
    df = pd.read_csv(os.path.join(REF, 'gridnext', 'visium_templates', 'tissue_positions.csv'), index_col=0)
    xy = np.array([ip.pseudo_hex_to_oddr(c, r) for c, r in zip(df['array_col'], df['array_row'])])
    np.savez_compressed(os.path.join(OUT, 't1_template_positions.npz'), in_tissue=df['in_tissue'].values.astype(np.uint8),
                        array_row=df['array_row'].values.astype(np.int16), array_col=df['array_col'].values.astype(np.int16),
                        pxl_row=df['pxl_row_in_fullres'].values.astype(np.int32), pxl_col=df['pxl_col_in_fullres'].values.astype(np.int32),
                        x_ind=xy[:, 0].astype(np.int16), y_ind=xy[:, 1].astype(np.int16))
    manifest['t1_template_positions'] = dict(n=len(df), n_in_tissue=int(df['in_tissue'].sum()))

    extras(manifest, (gm, dn, tr, ip))
    with open(os.path.join(OUT, 'state_dict_keys.json'), 'w') as fh:
        json.dump(KEYS, fh, sort_keys=True)
    with open(os.path.join(OUT, 'manifest.json'), 'w') as fh:
        json.dump(manifest, fh, indent=1, sort_keys=True)
    print(json.dumps(manifest, indent=1))


def extras(manifest, mods):
    """Vectors added after the first set (``python -m oracle.make_golden --extras`` regenerates only these)."""
    gm, dn, tr, ip = mods
    # ---- D3: tiny DenseNet in TRAIN mode (f pre-training, training.py:11-98): batch-stat BatchNorm, running-stat update
    tag, kw, P, N, seed = 'd3_densenet_tiny_train', dict(growth_rate=8, block_config=(2, 3), num_init_features=16, bn_size=2), 32, 6, 33
    net = dn.DenseNet(num_classes=7, small_inputs=False, efficient=False, drop_rate=0, **kw)
    load_synth(net, seed)
    net.train()
    gx = torch.Generator(); gx.manual_seed(seed + 100)
    x = torch.randn(N, 3, P, P, generator=gx)
    logits = net(x)
    gy = torch.Generator(); gy.manual_seed(seed + 200)
    dy = torch.randn(logits.shape, generator=gy)
    (logits * dy).sum().backward()
    g = grads_of(net)
    after = {k: v for k, v in net.state_dict().items() if 'running' in k or 'num_batches' in k}
    np.savez_compressed(os.path.join(OUT, tag + '.npz'), logits=logits.detach().numpy(),
                        **{'grad.' + k: v.numpy() for k, v in g.items()}, **{'after.' + k: v.numpy() for k, v in after.items()})
    manifest[tag] = dict(P=P, N=N, seed_w=seed, seed_x=seed + 100, seed_dy=seed + 200, **{k: list(v) if isinstance(v, tuple) else v for k, v in kw.items()})

    # ---- C1 / C2: base (Cartesian) GridNet, corrector = Conv2d 3x3, 5x5, 5x5, 3x3 (gridnet_models.py:51-66), with / without BN
    for tag, use_bn, seed in (('c1_cartesian_bn', True, 51), ('c2_cartesian_nobn', False, 52)):
        n_cls, f_dim, B, H, W = 5, 6, 3, 9, 11
        model = gm.GridNet(nn.Linear(4, f_dim), (4,), (H, W), n_cls, use_bn=use_bn, f_dim=f_dim)
        load_synth(model, seed, tag)
        gx = torch.Generator(); gx.manual_seed(seed + 100)
        x = torch.randn(B, H, W, 4, generator=gx)
        from oracle import synth
        y = synth.synth_labels(B, n_cls, H, W, seed=seed + 200)
        out, loss, ncorr, nfg = step_like_train_gridwise(model, x, y)
        g = grads_of(model)
        after = {k: v for k, v in model.state_dict().items() if 'running' in k}
        np.savez_compressed(os.path.join(OUT, tag + '.npz'), x=x.numpy(), y=y.numpy(), out=out.numpy(), loss=loss.numpy(), ncorr=ncorr, nfg=nfg,
                            **{'grad.' + k: v.numpy() for k, v in g.items()}, **{'after.' + k: v.numpy() for k, v in after.items()})
        manifest[tag] = dict(n_cls=n_cls, f_dim=f_dim, B=B, H=H, W=W, seed_w=seed, use_bn=use_bn)
    # ---- A1: read_annotated_starray (utils.py:87-166) on synthetic Splotch-format files with a one-hot annotation matrix run
    # through the reference's own read_annotfile (incl. its row filter, utils.py:239).  (annot_file=None raises
    # UnboundLocalError in the reference, utils.py:164, so there is no un-annotated vector.)
    import gridnext.utils as ut
    rng = np.random.RandomState(61)
    G, n_spots, n_cls = 12, 300, 6
    all_xy = [(c, r) for r in range(78) for c in range(r % 2, 128, 2)]
    sel = rng.choice(len(all_xy), n_spots, replace=False)
    coords = np.array([all_xy[i] for i in sel], dtype=np.int32)
    cstrs = ['%d_%d' % (c, r) for c, r in coords]
    cmat = rng.poisson(2.0, (G, n_spots)).astype(np.float64) * rng.rand(G, n_spots).round(3)
    lbl = rng.randint(0, n_cls, n_spots)
    lbl[:3] = [n_cls, n_cls + 1, n_cls + 2]                    # three classes that occur exactly once survive the row filter
    onehot = np.zeros((n_cls + 3, n_spots), dtype=int)
    annotated = rng.rand(n_spots) < 0.8
    annotated[:3] = True
    for j in range(n_spots):
        if annotated[j]:
            onehot[lbl[j], j] = 1
    with tempfile.TemporaryDirectory() as td:
        import pandas as pd
        cf, af = os.path.join(td, 'counts.tsv'), os.path.join(td, 'annots.tsv')
        pd.DataFrame(cmat, index=['g%d' % i for i in range(G)], columns=cstrs).to_csv(cf, sep='\t')
        pd.DataFrame(onehot, index=['c%d' % i for i in range(n_cls + 3)], columns=cstrs).to_csv(af, sep='\t')
        cg1, ag1 = ut.read_annotated_starray(cf, af)
        a_coords, a_lbls = ut.read_annotfile(af, Visium=False, afile_delim='\t')
    np.savez_compressed(os.path.join(OUT, 'a1_starray.npz'), cmat=cmat.astype(np.float32), coords=coords,
                        counts_annot=np.transpose(cg1, (2, 0, 1)).astype(np.float32), annots_annot=ag1.astype(np.int64),
                        annot_coords=np.array(list(a_coords)), annot_lbls=np.asarray(a_lbls).astype(np.int64))
    manifest['a1_starray'] = dict(G=G, n_spots=n_spots, n_annot_rows=int(len(a_coords)))
    # ---- A2: PatchGridDataset.__getitem__ (image_datasets.py:190-232) on synthetic patch files (lossless PNG), Loupe-format
    # annotations + a Spaceranger-v2 position file: the patch grid and label grid the reference hands to GridNet
    import gridnext.image_datasets as idt
    from PIL import Image as PILImage
    rng = np.random.RandomState(71)
    n_p, hp = 140, 6
    sel = rng.choice(len(all_xy), n_p, replace=False)
    pcoords = np.array([all_xy[i] for i in sel], dtype=np.int32)                 # (array_col, array_row) pseudo-hex
    ppatches = rng.randint(0, 256, (n_p, hp, hp, 3)).astype(np.uint8)
    names = ['tumor', 'stroma', 'immune', 'necrosis']
    plabel = rng.randint(0, len(names), n_p)
    annotated = rng.rand(n_p) < 0.75
    with tempfile.TemporaryDirectory() as td:
        idir = os.path.join(td, 'arr0'); os.makedirs(idir)
        for i in range(n_p):
            PILImage.fromarray(ppatches[i]).save(os.path.join(idir, 'spot_%d_%d.png' % (pcoords[i, 0], pcoords[i, 1])))
        pf, af = os.path.join(td, 'tissue_positions.csv'), os.path.join(td, 'annots.csv')
        with open(pf, 'w') as fh:
            fh.write('barcode,in_tissue,array_row,array_col,pxl_row_in_fullres,pxl_col_in_fullres\n')
            for i in range(n_p):
                fh.write('BC%04d-1,1,%d,%d,%d,%d\n' % (i, pcoords[i, 1], pcoords[i, 0], 100 + i, 200 + i))
        with open(af, 'w') as fh:
            fh.write('Barcode,annotation\n')
            for i in range(n_p):
                if annotated[i]:
                    fh.write('BC%04d-1,%s\n' % (i, names[plabel[i]]))
        ds = idt.PatchGridDataset([idir], [af], [pf], Visium=True, img_ext='png')
        pg, ag = ds[0]
        classes = list(ds.classes)
    np.savez_compressed(os.path.join(OUT, 'a2_patchgrid.npz'), patches=ppatches, coords=pcoords, label=plabel, annotated=annotated,
                        classes=np.array(classes), patch_grid=pg.numpy(), annots_grid=ag.numpy())
    manifest['a2_patchgrid'] = dict(n=n_p, hp=hp, classes=classes)
    keys_path = os.path.join(OUT, 'state_dict_keys.json')
    if os.path.exists(keys_path):
        old = json.load(open(keys_path))
        old.update(KEYS)
        KEYS.update(old)


def extras2(manifest, mods):
    """Round-2 vectors: grid_from_wsi_visium with window_size != patch_size (Pillow BICUBIC resize inside the reference loop,
    imgprocess.py:188-195,221), a float window, and a non-Normalize transform."""
    from oracle import synth
    gm, dn, tr, ip = mods
    from PIL import Image
    from torchvision import transforms
    tis, rows, cols, pr, pc = synth.synth_positions(pitch_col=5.5, pitch_row=9.6, org_row=3.0, org_col=2.0)
    Himg, Wimg = 790, 730
    img = synth.synth_image(Himg, Wimg, seed=7, smooth=True)
    out = {}
    with tempfile.TemporaryDirectory() as td:
        Image.fromarray(img).save(os.path.join(td, 'img.png'))
        sp = os.path.join(td, 'outs', 'spatial'); os.makedirs(sp)
        with open(os.path.join(sp, 'tissue_positions.csv'), 'w') as fh:
            fh.write('barcode,in_tissue,array_row,array_col,pxl_row_in_fullres,pxl_col_in_fullres\n')
            for i in range(len(tis)):
                fh.write('BC%05d-1,%d,%d,%d,%r,%r\n' % (i, tis[i], rows[i], cols[i], float(pr[i]), float(pc[i])))
        f, d = os.path.join(td, 'img.png'), os.path.join(td, 'outs')
        out['down_24_to_16'] = ip.grid_from_wsi_visium(f, d, patch_size=16, window_size=24).numpy().astype(np.uint8)
        out['up_10_to_16'] = ip.grid_from_wsi_visium(f, d, patch_size=16, window_size=10).numpy().astype(np.uint8)
        out['float_0.03_to_12'] = ip.grid_from_wsi_visium(f, d, patch_size=12, window_size=0.03).numpy().astype(np.uint8)     # int(0.03 * 730) = 21 -> side 20
        nrm = ip.grid_from_wsi_visium(f, d, patch_size=16, window_size=24,
                                      preprocess_xform=transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225]))
        out['down_24_to_16_nrm_sub'] = nrm.numpy()[::11, ::9]
        gray = ip.grid_from_wsi_visium(f, d, patch_size=16, window_size=16,
                                       preprocess_xform=transforms.Compose([transforms.Grayscale(num_output_channels=3), transforms.Normalize([0.5] * 3, [0.25] * 3)]))
        out['crop_16_gray_nrm_sub'] = gray.numpy()[::11, ::9]
    # keep the fixture small: every 4th grid row / column plus the cells next to all four borders
    ys = sorted(set(range(0, 78, 4)) | {1, 76, 77})
    xs = sorted(set(range(0, 64, 4)) | {1, 62, 63})
    for k in list(out):
        if not k.endswith('_sub'):
            out[k] = out[k][np.ix_(ys, xs)]
    out['cells_y'], out['cells_x'] = np.asarray(ys), np.asarray(xs)
    np.savez_compressed(os.path.join(OUT, 'p2_gather_resize.npz'), **out)
    manifest['p2_gather_resize'] = dict(Himg=Himg, Wimg=Wimg, pitch_col=5.5, pitch_row=9.6, org_row=3.0, org_col=2.0, img_seed=7,
                                        cases=sorted(k for k in out if not k.startswith('cells_')), pillow=Image.__version__ if hasattr(Image, '__version__') else None)


if __name__ == '__main__':
    if '--extras2' in sys.argv:
        man = json.load(open(os.path.join(OUT, 'manifest.json')))
        extras2(man, import_reference())
        with open(os.path.join(OUT, 'manifest.json'), 'w') as fh:
            json.dump(man, fh, indent=1, sort_keys=True)
    elif '--extras' in sys.argv:
        man = json.load(open(os.path.join(OUT, 'manifest.json')))
        torch.set_num_threads(os.cpu_count())
        extras(man, import_reference())
        with open(os.path.join(OUT, 'manifest.json'), 'w') as fh:
            json.dump(man, fh, indent=1, sort_keys=True)
        with open(os.path.join(OUT, 'state_dict_keys.json'), 'w') as fh:
            json.dump(KEYS, fh, sort_keys=True)
    else:
        main()
