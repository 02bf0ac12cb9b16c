"""Dataset tensor assembly on the device.

The reference builds every array's input tensors in python loops over spots
  * ``utils.read_annotated_starray`` (/root/reference/gridnext/utils.py:144-166): ``counts_grid[y, x] = cmat[spot]``,
    ``annots_grid[y, x] = label + 1`` (0 = background), then ``count_datasets.py:292-293`` permutes to channels-first;
  * ``PatchGridDataset.__getitem__`` (image_datasets.py:205-232): ``patch_grid[y, x] = patch`` for every patch file, labels
    only for annotated spots;
  * ``MultiModalGridDataset.__getitem__`` (multimodal_datasets.py:237-244): a 78x64 double loop of ``.max()`` calls that
    zeroes spots lacking image data or annotations.
Parsing files stays host code; these functions take the parsed arrays (count matrix, spot coordinates, integer labels) and
produce the same tensors with one pass over the output on the GPU (csrc/grid_assemble.cu).  There is no CPU fallback.
"""
import numpy as np
import torch

from . import _lib
from ._lib import ptr, stream, call
from .utils import pseudo_hex_to_oddr

VISIUM_H_ST, VISIUM_W_ST = 78, 64


def spot_cells(xs, ys, visium=True, h_st=VISIUM_H_ST, w_st=VISIUM_W_ST, device=None):
    """Array coordinates of each spot -> int32 row-major cell index y*w_st + x (Visium: pseudo-hex -> odd-r, utils.py:64-70;
    otherwise rint of the Cartesian coordinates, utils.py:153-154).  Spots falling outside the array raise, as indexing the
    reference's grid would."""
    cells = np.empty(len(xs), dtype=np.int32)
    for i, (cx, cy) in enumerate(zip(xs, ys)):
        if visium:
            x, y = pseudo_hex_to_oddr(int(cx), int(cy))
        else:
            x, y = int(np.rint(float(cx))), int(np.rint(float(cy)))
        if not (-w_st <= x < w_st and -h_st <= y < h_st):
            raise IndexError('index (%d, %d) is out of bounds for a (%d, %d) array' % (y, x, h_st, w_st))
        cells[i] = (y % h_st) * w_st + (x % w_st)          # numpy's negative-index wrap-around
    t = torch.from_numpy(cells)
    return t.to(device) if device is not None else t


def _inverse(cells, n_cells):
    _lib.require_cuda(cells)
    if cells.dtype != torch.int32:
        raise ValueError('cells must be int32')
    inv = torch.empty(n_cells, device=cells.device, dtype=torch.int32)
    call('gn_cell_inverse', ptr(cells), cells.numel(), ptr(inv), n_cells, stream())
    return inv


def _labels_grid(labels, inv, h_st, w_st):
    annots = torch.empty((h_st, w_st), device=inv.device, dtype=torch.int64)
    if labels is not None:
        _lib.require_cuda(labels)
        labels = labels.to(torch.int64).contiguous()
    call('gn_grid_labels', ptr(labels), ptr(inv), ptr(annots), h_st * w_st, stream())
    return annots


def assemble_count_grid(cmat, cells, labels=None, h_st=VISIUM_H_ST, w_st=VISIUM_W_ST):
    """cmat: (n_genes, n_spots) fp32 CUDA (the Splotch count matrix, genes x spots); cells: int32 cell of every spot (< 0: spot
    not included, e.g. un-annotated when an annotation file is given, utils.py:157-158); labels: int64 per spot or None.
    -> (counts_grid (n_genes, h_st, w_st) fp32, annots_grid (h_st, w_st) int64), the pair CountGridDataset.__getitem__ returns."""
    _lib.require_cuda(cmat, cells)
    if cmat.dim() != 2 or cmat.dtype != torch.float32 or cmat.shape[1] != cells.numel():
        raise ValueError('assemble_count_grid: cmat must be (n_genes, n_spots) float32 with one cell per spot')
    cmat = cmat.contiguous()
    G, n_cells = cmat.shape[0], h_st * w_st
    inv = _inverse(cells.contiguous(), n_cells)
    grid = torch.empty((G, h_st, w_st), device=cmat.device, dtype=torch.float32)
    call('gn_grid_gather_cols', ptr(cmat), cmat.stride(0), ptr(inv), ptr(grid), n_cells, G, n_cells, stream())
    return grid, _labels_grid(labels, inv, h_st, w_st)


def assemble_patch_grid(patches, cells, labels=None, h_st=VISIUM_H_ST, w_st=VISIUM_W_ST):
    """patches: (n_spots, C, h, w) CUDA, any dtype; cells: int32 cell of every patch; labels: int64 per patch, < 0 = not annotated.
    -> (patch_grid (h_st, w_st, C, h, w) in the patches' dtype, annots_grid (h_st, w_st) int64) (image_datasets.py:205-232)."""
    _lib.require_cuda(patches, cells)
    if patches.dim() < 2 or patches.shape[0] != cells.numel():
        raise ValueError('assemble_patch_grid: one cell per patch expected')
    patches = patches.contiguous()
    n_cells = h_st * w_st
    inv = _inverse(cells.contiguous(), n_cells)
    row_bytes = patches[0].numel() * patches.element_size()
    grid = torch.empty((h_st, w_st) + tuple(patches.shape[1:]), device=patches.device, dtype=patches.dtype)
    call('gn_grid_gather_rows', ptr(patches), row_bytes, ptr(inv), ptr(grid), row_bytes, n_cells, row_bytes, stream())
    return grid, _labels_grid(labels, inv, h_st, w_st)


def multimodal_fg_consistency(counts_grid, patch_grid, annots_grid):
    """In place, multimodal_datasets.py:237-244: a spot whose patch has max == 0 loses its label and its counts; a spot without
    a label loses its patch.  counts_grid (G, H, W) fp32, patch_grid (H, W, ...) fp32, annots_grid (H, W) int64; returns them."""
    _lib.require_cuda(counts_grid, patch_grid, annots_grid)
    H, W = annots_grid.shape
    if (patch_grid.dtype != torch.float32 or counts_grid.dtype != torch.float32 or annots_grid.dtype != torch.int64
            or tuple(patch_grid.shape[:2]) != (H, W) or tuple(counts_grid.shape[1:]) != (H, W)):
        raise ValueError('multimodal_fg_consistency: expected counts (G, H, W) fp32, patches (H, W, ...) fp32, annots (H, W) int64')
    if not (counts_grid.is_contiguous() and patch_grid.is_contiguous() and annots_grid.is_contiguous()):
        raise ValueError('multimodal_fg_consistency: tensors must be contiguous (they are modified in place)')
    n_cells = H * W
    flags = torch.empty(n_cells, device=annots_grid.device, dtype=torch.uint8)
    call('gn_mm_fg_consistency', ptr(patch_grid), patch_grid[0, 0].numel(), ptr(counts_grid), n_cells, counts_grid.shape[0], ptr(annots_grid),
         ptr(flags), n_cells, stream())
    return counts_grid, patch_grid, annots_grid
