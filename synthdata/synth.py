"""Seeded synthetic weights / inputs shared by the golden generator, tests and bench.py's CPU
arm (test infrastructure only).  Weights never depend on module-construction RNG order:
each tensor is drawn from its own generator seeded by a stable hash of its key."""
import zlib
import numpy as np
import torch


def _gen(key, seed):
    g = torch.Generator()
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)
    return g


def synth_tensor(key, shape, seed=1234):
    """Plausible values per parameter kind, keyed on the reference's state-dict naming."""
    g = _gen(key, seed)
    shape = tuple(shape)
    leaf = key.split('.')[-1]
    if leaf == 'num_batches_tracked':
        return torch.zeros(shape, dtype=torch.long)
    if leaf == 'running_var':
        return torch.rand(shape, generator=g) + 0.5
    if leaf == 'running_mean':
        return 0.1 * torch.randn(shape, generator=g)
    if leaf == 'dummy_tensor':
        return torch.ones(shape)
    if leaf == 'bg_const':
        return torch.zeros(shape)
    is_norm = ('norm' in key) or (len(shape) == 1 and leaf == 'weight')
    if is_norm and leaf == 'weight':
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    if leaf in ('bias', 'bias_tensor'):
        return 0.1 * torch.randn(shape, generator=g)
    fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else int(shape[0])
    if leaf.startswith('kernel'):       # hexagdly kernels: fan-in over the 7-tap hexagon
        fan_in = shape[1] * 7
    return torch.randn(shape, generator=g) * (2.0 / max(fan_in, 1)) ** 0.5


def synth_state_dict(shapes, seed=1234):
    """shapes: {key: shape}.  Returns {key: tensor} (float32 / int64)."""
    return {k: synth_tensor(k, s, seed) for k, s in shapes.items()}


def shapes_of(module):
    return {k: tuple(v.shape) for k, v in module.state_dict().items()}


def tissue_mask(h=78, w=64):
    """Deterministic roundish foreground (~90 % of cells, like the Visium template's 4,525/4,992)."""
    y, x = np.mgrid[0:h, 0:w]
    yy = (y - (h - 1) / 2) / (h / 2)
    xx = (x - (w - 1) / 2) / (w / 2)
    return ((np.abs(yy) ** 6 + np.abs(xx) ** 6) < 0.95)


def synth_labels(B, n_cls, h=78, w=64, seed=0):
    m = torch.from_numpy(tissue_mask(h, w))
    out = torch.zeros(B, h, w, dtype=torch.long)
    for b in range(B):
        g = torch.Generator(); g.manual_seed(seed + b)
        out[b] = torch.randint(1, n_cls + 1, (h, w), generator=g) * m
    return out


def synth_counts(B, G, h=78, w=64, seed=0):
    """log1p(poisson(1)) counts, zero off tissue."""
    m = torch.from_numpy(tissue_mask(h, w)).float()
    out = torch.empty(B, G, h, w)
    for b in range(B):
        g = torch.Generator(); g.manual_seed(1000 + seed + b)
        out[b] = torch.log1p(torch.poisson(torch.ones(G, h, w), generator=g)) * m
    return out


def synth_positions(h=78, w=64, pitch_col=113.25, pitch_row=197.0, org_row=1157.0, org_col=1490.0,
                    all_in_tissue=False, frac=True):
    """Visium-style position table: pseudo-hex array_col = 2x + (row & 1)."""
    rows, cols, pr, pc, tis = [], [], [], [], []
    m = tissue_mask(h, w)
    for r in range(h):
        for x in range(w):
            c = 2 * x + (r & 1)
            rows.append(r); cols.append(c)
            pr.append(org_row + pitch_row * r + (0.5 if (frac and (r + x) % 7 == 0) else 0.0))
            pc.append(org_col + pitch_col * c + (0.5 if (frac and (r * 3 + x) % 5 == 0) else 0.0))
            tis.append(1 if (all_in_tissue or m[r, x]) else 0)
    return (np.asarray(tis), np.asarray(rows), np.asarray(cols), np.asarray(pr), np.asarray(pc))


def synth_image(H, W, seed=7, smooth=False):
    rng = np.random.default_rng(seed)
    if not smooth:
        return rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    y, x = np.mgrid[0:H, 0:W]
    img = np.stack([(127 + 120 * np.sin(x / 37.0 + c) * np.cos(y / 53.0 - c)) for c in range(3)], -1)
    return np.clip(img + rng.normal(0, 3, img.shape), 0, 255).astype(np.uint8)
