"""State-dict key -> shape tables of the reference modules, derived from their constructors
(test infrastructure only).  gridnext/densenet.py:93-138, gridnext/gridnet_models.py:24-48,
128-148,194-209; hexagdly.Conv2d parameters kernel0..k, bias_tensor."""
def kernel_shapes(cin, cout, k):
    """hexagdly.Conv2d parameter shapes kernel0..kernel{k} (gridnext_b200/hexagdly.py; the same table as oracle/hexconv_ref.py)."""
    return [(cout, cin, 2 * k + 1 - i, 1 if i == 0 else 2) for i in range(k + 1)]


def bn_shapes(prefix, c):
    return {prefix + 'weight': (c,), prefix + 'bias': (c,), prefix + 'running_mean': (c,),
            prefix + 'running_var': (c,), prefix + 'num_batches_tracked': ()}


def densenet_shapes(growth_rate=32, block_config=(6, 12, 24, 16), num_init_features=64, bn_size=4,
                    num_classes=7, compression=0.5, small_inputs=False):
    s = {}
    if small_inputs:
        s['features.conv0.weight'] = (num_init_features, 3, 3, 3)
    else:
        s['features.conv0.weight'] = (num_init_features, 3, 7, 7)
        s.update(bn_shapes('features.norm0.', num_init_features))
    nf = num_init_features
    for bi, nl in enumerate(block_config, start=1):
        for li in range(1, nl + 1):
            p = 'features.denseblock%d.denselayer%d.' % (bi, li)
            cin = nf + (li - 1) * growth_rate
            s.update(bn_shapes(p + 'norm1.', cin))
            s[p + 'conv1.weight'] = (bn_size * growth_rate, cin, 1, 1)
            s.update(bn_shapes(p + 'norm2.', bn_size * growth_rate))
            s[p + 'conv2.weight'] = (growth_rate, bn_size * growth_rate, 3, 3)
        nf = nf + nl * growth_rate
        if bi != len(block_config):
            p = 'features.transition%d.' % bi
            s.update(bn_shapes(p + 'norm.', nf))
            s[p + 'conv.weight'] = (int(nf * compression), nf, 1, 1)
            nf = int(nf * compression)
    s.update(bn_shapes('features.norm_final.', nf))
    s['classifier.weight'] = (num_classes, nf)
    s['classifier.bias'] = (num_classes,)
    return s


def mlp_shapes(G, n_cls, widths=(500, 100, 100, 50)):
    s = {}
    dims = [G, widths[0], widths[1], None, None, widths[2], widths[3], None, None, n_cls]
    s['0.weight'] = (widths[0], G); s['0.bias'] = (widths[0],)
    s['1.weight'] = (widths[1], widths[0]); s['1.bias'] = (widths[1],)
    s.update(bn_shapes('2.', widths[1]))
    s['4.weight'] = (widths[2], widths[1]); s['4.bias'] = (widths[2],)
    s['5.weight'] = (widths[3], widths[2]); s['5.bias'] = (widths[3],)
    s.update(bn_shapes('6.', widths[3]))
    s['8.weight'] = (n_cls, widths[3]); s['8.bias'] = (n_cls,)
    return s


def corrector_shapes(f_dim, n_cls, use_bn=True, ksize=1, width=32):
    hex_idx = (0, 1, 4, 5, 8) if use_bn else (0, 1, 3, 4, 6)
    chans = [(f_dim, width), (width, width), (width, width), (width, width), (width, n_cls)]
    s = {}
    for i, (ci, co) in zip(hex_idx, chans):
        for j, shp in enumerate(kernel_shapes(ci, co, ksize)):
            s['%d.kernel%d' % (i, j)] = shp
        s['%d.bias_tensor' % i] = (co,)
    if use_bn:
        for i in (2, 6):
            s.update(bn_shapes('%d.' % i, width))
    return s


def with_prefix(prefix, shapes):
    return {prefix + k: v for k, v in shapes.items()}


def cartesian_corrector_shapes(f_dim, n_cls, use_bn=True):
    """nn.Sequential of the base GridNet (gridnet_models.py:51-66)."""
    s, idx = {}, 0
    for cin, K in ((f_dim, 3), (n_cls, 5), (n_cls, 5)):
        s['%d.weight' % idx] = (n_cls, cin, K, K)
        s['%d.bias' % idx] = (n_cls,)
        if use_bn:
            for k, shp in (('weight', (n_cls,)), ('bias', (n_cls,)), ('running_mean', (n_cls,)), ('running_var', (n_cls,)), ('num_batches_tracked', ())):
                s['%d.%s' % (idx + 1, k)] = shp
        idx += 3 if use_bn else 2
    s['%d.weight' % idx] = (n_cls, n_cls, 3, 3)
    s['%d.bias' % idx] = (n_cls,)
    return s


def cartesian_gridnet_shapes(f_shapes, f_dim, n_cls, use_bn=True):
    s = {'bg_const': (1, f_dim), 'dummy_tensor': (1,)}
    s.update(with_prefix('patch_classifier.', f_shapes))
    s.update(with_prefix('corrector.', cartesian_corrector_shapes(f_dim, n_cls, use_bn)))
    return s


def gridnet_shapes(f_shapes, f_dim, n_cls, use_bn=True):
    s = {'bg_const': (1, f_dim), 'dummy_tensor': (1,)}
    s.update(with_prefix('patch_classifier.', f_shapes))
    s.update(with_prefix('corrector.', corrector_shapes(f_dim, n_cls, use_bn)))
    return s


def gridnet_mm_shapes(image_shapes, count_shapes, image_f_dim, count_f_dim, n_cls, use_bn=True):
    f_dim = image_f_dim + count_f_dim
    s = {'bg_const': (1, f_dim), 'dummy_tensor': (1,)}
    s.update(with_prefix('patch_classifier.', image_shapes))
    s.update(with_prefix('corrector.', corrector_shapes(f_dim, n_cls, use_bn)))
    s.update(with_prefix('image_classifier.', image_shapes))
    s.update(with_prefix('count_classifier.', count_shapes))
    return s
