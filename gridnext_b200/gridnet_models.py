"""GridNet model classes with the reference's constructor signatures, attributes and state-dict keys.

Mirrors /root/reference/gridnext/gridnet_models.py: ``init_weights`` (:14), ``GridNet`` (:23),
``GridNetHex`` (:122), ``GridNetHexOddr`` (:159), ``GridNetHexMM`` (:193).  Users subclass and override
``_init_corrector()`` / ``patch_predictions()`` exactly as with the reference.

What differs underneath: the hexagonal corrector runs as one fused autograd node on hand-written
sm_100a kernels in the Visium layout (no rot90/flip copies); recognised f networks (the tutorial count
MLP as an ``nn.Sequential`` and ``gridnext_b200.densenet.DenseNet``) run on the tensor-core kernels
with the layout shuffle of patch_predictions folded into their operand loads.
"""
import torch
import torch.nn as nn
import torch.utils.checkpoint as cp

from . import hexagdly
from .corrector import parse_corrector, run_corrector


def init_weights(m):
    """Xavier-uniform weights / zero biases for Conv2d and Linear, unit-gamma BatchNorm2d (reference gridnet_models.py:14-20)."""
    if type(m) in (nn.Conv2d, nn.Linear):
        nn.init.xavier_uniform_(m.weight)
        nn.init.zeros_(m.bias)
    elif type(m) is nn.BatchNorm2d:
        nn.init.ones_(m.weight)
        nn.init.zeros_(m.bias)


def _fast_f(module):
    """Return a callable running ``module`` on the fused kernels for spot-major input, or None."""
    try:
        from .count_mlp import compile_count_mlp
    except ImportError:
        return None
    return compile_count_mlp(module)


class GridNet(nn.Module):
    """Base class: Cartesian grids; the square-conv corrector (gridnet_models.py:51-66) runs on the same tile kernels as the
    hexagonal one (a K x K window is a parity-independent tap table), BatchNorm/ReLU fused the same way."""

    def __init__(self, patch_classifier, patch_shape, grid_shape, n_classes,
                 use_bn=True, atonce_patch_limit=None, f_dim=None):
        super().__init__()
        self.patch_classifier = patch_classifier
        self.patch_shape, self.grid_shape = patch_shape, grid_shape
        self.n_classes, self.use_bn, self.atonce_patch_limit = n_classes, use_bn, atonce_patch_limit
        self.f_dim = n_classes if f_dim is None else f_dim
        self.corrector = self._init_corrector()
        # the two buffers below exist only for state-dict compatibility with the reference (gridnet_models.py:44-48)
        self.bg = torch.zeros((1, self.f_dim), requires_grad=True)
        self.register_buffer("bg_const", self.bg)
        self.dummy = torch.ones(1, dtype=torch.float32, requires_grad=True)
        self.register_buffer("dummy_tensor", self.dummy)

    def _init_corrector(self):
        n, layers = self.n_classes, []
        for cin, k in ((self.f_dim, 3), (n, 5), (n, 5)):
            layers.append(nn.Conv2d(cin, n, k, padding=k // 2))
            if self.use_bn:
                layers.append(nn.BatchNorm2d(n))
            layers.append(nn.ReLU())
        layers.append(nn.Conv2d(n, n, 3, padding=1))
        return nn.Sequential(*layers)

    def foreground_classifier(self, x):
        if torch.max(x) == 0:
            return self.bg_const
        return self.patch_classifier(x.unsqueeze(0))

    def _ppl(self, patch_list, dummy_arg=None):
        assert dummy_arg is not None
        return self.patch_classifier(patch_list)

    def _f_on_spots(self, patch_list):
        """f over a flat spot list (N, ...).  ``atonce_patch_limit`` keeps the reference's contract (gridnet_models.py:88-104):
        chunks of that many spots, each under activation checkpointing when f is being trained."""
        limit = self.atonce_patch_limit
        if limit is None:
            return self._ppl(patch_list, self.dummy_tensor)
        recompute = torch.is_grad_enabled() and any(p.requires_grad for p in self.patch_classifier.parameters())
        outs = []
        for chunk in patch_list.split(limit, dim=0):
            if recompute:
                outs.append(cp.checkpoint(self._ppl, chunk, self.dummy_tensor, use_reentrant=True))
            else:
                outs.append(self._ppl(chunk, self.dummy_tensor))
        return torch.cat(outs, 0)

    def patch_predictions(self, x):
        patch_list = torch.reshape(x, (-1,) + tuple(self.patch_shape))
        patch_pred_list = self._f_on_spots(patch_list)
        patch_pred_grid = torch.reshape(patch_pred_list, (-1,) + tuple(self.grid_shape) + (self.f_dim,))
        return patch_pred_grid.permute((0, 3, 1, 2))

    def forward(self, x):
        ppg = self.patch_predictions(x)
        stages = parse_corrector(self.corrector) if ppg.is_cuda else None
        if stages is not None:
            return run_corrector(stages, ppg, self.training)
        return self.corrector(ppg)       # user-defined corrector: module by module, like the reference


class GridNetHex(GridNet):
    """Hexagonally packed grids; input/outputs in HexagDLy's own addressing (B, C, rows, cols)."""

    def __init__(self, patch_classifier, patch_shape, grid_shape, n_classes,
                 use_bn=True, atonce_patch_limit=None, f_dim=None):
        super(GridNetHex, self).__init__(patch_classifier, patch_shape, grid_shape, n_classes,
                                         use_bn, atonce_patch_limit, f_dim)

    def _init_corrector(self):
        # hex hex [BN] ReLU | hex hex [BN] ReLU | hex, 32 channels wide (module order = the reference's state-dict indices)
        width = 32

        def hexl(cin, cout):
            return hexagdly.Conv2d(in_channels=cin, out_channels=cout, kernel_size=1, stride=1, bias=True)

        seq = []
        for cin in (self.f_dim, width):
            seq += [hexl(cin, width), hexl(width, width)]
            if self.use_bn:
                seq.append(nn.BatchNorm2d(width))
            seq.append(nn.ReLU())
        seq.append(hexl(width, self.n_classes))
        return nn.Sequential(*seq)

    def _correct_visium(self, grid):
        """Apply the corrector to a (B, C, H, W) tensor whose hex parity is on the row index."""
        stages = parse_corrector(self.corrector)
        if stages is not None:
            # nn.Conv2d stages of a user corrector (e.g. notebooks/register_concat.ipynb) see the HexagDLy layout in the
            # reference, i.e. the transposed grid: their kernels are applied with the spatial axes swapped
            return run_corrector(stages, grid, self.training, sq_transposed=True)
        # user-defined corrector: module by module in HexagDLy layout, like the reference
        return self.corrector(grid.transpose(2, 3).contiguous()).transpose(2, 3)

    def forward(self, x):
        # HexagDLy addressing: parity on the last index == Visium layout transposed
        ppg = self.patch_predictions(x)
        return self._correct_visium(ppg.transpose(2, 3)).transpose(2, 3)


class GridNetHexOddr(GridNetHex):
    """Visium odd-right indexing:  1D spot features (B, feats, H, W) | >1D (B, H, W, feats...)
    -> (B, n_class, H, W)."""

    def patch_predictions(self, x):
        if len(x.shape) == 4:
            fast = _fast_f(self.patch_classifier)
            if fast is not None and x.is_cuda and self.atonce_patch_limit is None:
                # count slab (B, G, H, W) consumed directly: no permute/reshape copy (K2 in SURVEY 2.2)
                out = fast.forward_grid(x, self.f_dim)
                if out is not None:
                    return out
            return super(GridNetHexOddr, self).patch_predictions(x.permute((0, 2, 3, 1)))
        return super(GridNetHexOddr, self).patch_predictions(x)

    def forward(self, x):
        # rot90+flip of the reference (gridnet_models.py:177-185) is a transpose into HexagDLy layout and
        # back; the kernels take the row parity directly, so nothing is moved.
        return self._correct_visium(self.patch_predictions(x))


class GridNetHexMM(GridNetHexOddr):
    """Two f networks (image: >1D inputs, count: 1D inputs); features concatenated [count | image]."""

    def __init__(self, image_classifier, count_classifier, image_shape, count_shape, grid_shape, n_classes,
                 use_bn=True, atonce_patch_limit=None, image_f_dim=None, count_f_dim=None):
        image_f_dim = n_classes if image_f_dim is None else image_f_dim
        count_f_dim = n_classes if count_f_dim is None else count_f_dim
        super().__init__(image_classifier, image_shape, grid_shape, n_classes, use_bn, atonce_patch_limit, image_f_dim + count_f_dim)
        self.image_classifier, self.count_classifier = image_classifier, count_classifier
        self.image_shape, self.count_shape = image_shape, count_shape
        self.image_f_dim, self.count_f_dim = image_f_dim, count_f_dim

    def _set_mode(self, mode):
        """Point patch_classifier / patch_shape / f_dim at one modality ('image' | 'count'), or restore the concatenated width."""
        if mode in ('image', 'count'):
            self.patch_classifier = getattr(self, mode + '_classifier')
            self.patch_shape = getattr(self, mode + '_shape')
            self.f_dim = getattr(self, mode + '_f_dim')
        else:
            self.f_dim = self.count_f_dim + self.image_f_dim

    def patch_predictions(self, x):
        x_image, x_count = x
        grids = []
        for mode, x_mode in (('count', x_count), ('image', x_image)):          # concatenation order of the reference: [count | image]
            self._set_mode(mode)
            grids.append(super().patch_predictions(x_mode))
        self._set_mode('concat')
        return torch.cat(grids, dim=1)
