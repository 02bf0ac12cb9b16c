"""Spot-patch gather from a full-resolution Visium H&E image.

Mirrors /root/reference/gridnext/imgprocess.py: constants (:21-22), ``pseudo_hex_to_oddr`` /
``oddr_to_pseudo_hex`` (:26-38) and ``grid_from_wsi_visium`` (:162-238).  The per-spot python loop
(crop, /255, Normalize, scatter into (78, 64, 3, P, P)) is one CUDA kernel (csrc/patch_gather.cu);
the integer index math (pseudo-hex -> odd-r, rint) runs on the device as well.
Pure-crop case only: ``2 * (w // 2) == patch_size``.
"""
import os
import glob
import numpy as np
import torch

from . import _lib
from ._lib import ptr, stream, call
from .utils import pseudo_hex_to_oddr, oddr_to_pseudo_hex  # noqa: F401  (re-exported like the reference)

VISIUM_H_ST = 78
VISIUM_W_ST = 64


def pseudo_hex_to_cartesian(c):
    """Visium pseudo-hex (col, row) -> Cartesian coordinates with unit neighbour distance (reference imgprocess.py:41-46)."""
    x, y = c
    return (x / 2, y * np.sqrt(3) / 2)


def _window(patch_size, window_size, xdim):
    if window_size is None:
        w = patch_size
    elif isinstance(window_size, float):
        w = int(window_size * xdim)
    elif isinstance(window_size, int):
        w = window_size
    else:
        raise ValueError("Window size must be a float or int")
    return w


def _normalize_params(preprocess_xform):
    """(mean, std) of a torchvision ``Normalize`` (optionally wrapped in a Compose of exactly that)."""
    if preprocess_xform is None:
        return None, None
    x = preprocess_xform
    tr = getattr(x, 'transforms', None)
    if tr is not None and len(tr) == 1:
        x = tr[0]
    if type(x).__name__ == 'Normalize' and hasattr(x, 'mean') and hasattr(x, 'std'):
        return [float(v) for v in x.mean], [float(v) for v in x.std]
    raise NotImplementedError("grid_from_wsi_visium on B200 supports preprocess_xform=None or a torchvision Normalize")


_CONST_CACHE = {}


def _dev_const(values, device):
    """Per-channel constants as a cached device tensor (no host->device copy per call; keeps the step graph-capturable)."""
    if values is None:
        return None
    key = (tuple(float(v) for v in values), str(device))
    t = _CONST_CACHE.get(key)
    if t is None:
        t = torch.tensor(key[0], device=device, dtype=torch.float32)
        _CONST_CACHE[key] = t
    return t


def spot_table(in_tissue, array_row, array_col, pxl_row, pxl_col, device, h_st=VISIUM_H_ST, w_st=VISIUM_W_ST):
    """Device table cells[h_st*w_st][3] = (cx, cy, valid) from a Visium position table.  Returns (cells, n_dropped)."""
    dev = torch.device(device)
    t = torch.as_tensor(np.asarray(in_tissue).astype(np.uint8)).to(dev)
    ar = torch.as_tensor(np.asarray(array_row).astype(np.int32)).to(dev)
    ac = torch.as_tensor(np.asarray(array_col).astype(np.int32)).to(dev)
    pr = torch.as_tensor(np.asarray(pxl_row).astype(np.float64)).to(dev)
    pc = torch.as_tensor(np.asarray(pxl_col).astype(np.float64)).to(dev)
    _lib.require_cuda(t)
    cells = torch.empty((h_st * w_st, 3), device=dev, dtype=torch.int32)
    dropped = torch.empty(1, device=dev, dtype=torch.int32)
    call('gn_spot_table', ptr(t), ptr(ar), ptr(ac), ptr(pr), ptr(pc), int(t.numel()), h_st, w_st, ptr(cells), ptr(dropped), stream())
    return cells, dropped


def gather_patches(img, cells, patch_size, mean=None, std=None, out_dtype=torch.float32, h_st=VISIUM_H_ST, w_st=VISIUM_W_ST, out=None):
    """img: CUDA uint8 (H, W, 3) contiguous; cells from ``spot_table``.  -> (h_st, w_st, 3, P, P)."""
    _lib.require_cuda(img, cells)
    if img.dtype != torch.uint8 or img.dim() != 3 or img.shape[2] != 3 or not img.is_contiguous():
        raise ValueError('gather_patches: image must be a contiguous uint8 (H, W, 3) tensor')
    if out_dtype not in (torch.float32, torch.bfloat16):
        raise ValueError('gather_patches: out_dtype must be float32 or bfloat16')
    H, W = int(img.shape[0]), int(img.shape[1])
    P = int(patch_size)
    if out is None:
        out = torch.empty((h_st, w_st, 3, P, P), device=img.device, dtype=out_dtype)
    m, s = _dev_const(mean, img.device), _dev_const(std, img.device)
    call('gn_patch_gather', ptr(img), 3 * W, H, W, ptr(cells), h_st * w_st, P, ptr(m), ptr(s), ptr(out),
         1 if out_dtype == torch.bfloat16 else 0, stream())
    return out


def normalize_patches(patches_u8, mean=None, std=None, out_dtype=torch.float32, valid=None, out=None):
    """uint8 (..., 3, P, P) CUDA patch grid -> ToTensor + Normalize on the device (absent cells stay 0)."""
    _lib.require_cuda(patches_u8)
    if patches_u8.dtype != torch.uint8 or not patches_u8.is_contiguous() or patches_u8.shape[-3] != 3:
        raise ValueError('normalize_patches: expected a contiguous uint8 (..., 3, P, P) tensor')
    P = int(patches_u8.shape[-1])
    n_cells = patches_u8.numel() // (3 * P * P)
    if out is None:
        out = torch.empty(patches_u8.shape, device=patches_u8.device, dtype=out_dtype)
    elif out.dtype != out_dtype or out.numel() != patches_u8.numel() or not out.is_contiguous():
        raise ValueError('normalize_patches: out must be a contiguous %s tensor of the input size' % out_dtype)
    m, s = _dev_const(mean, out.device), _dev_const(std, out.device)
    call('gn_normalize_u8', ptr(patches_u8), ptr(valid), n_cells, P, ptr(m), ptr(s), ptr(out), 1 if out_dtype == torch.bfloat16 else 0, stream())
    return out


def read_positions(spaceranger_dir):
    """tissue_positions(.csv|_list.csv) -> dict of numpy columns (Spaceranger >= 2 has a header row)."""
    import pandas as pd
    paths = [p for p in glob.glob(spaceranger_dir + '/**/*.csv', recursive=True) if 'tissue_positions' in p]
    if not paths:
        raise ValueError("Cannot location position file for %s" % spaceranger_dir)
    path = paths[0]
    with open(path) as fh:
        has_header = fh.readline().startswith('barcode')
    names = ["in_tissue", "array_row", "array_col", "pxl_row_in_fullres", "pxl_col_in_fullres"]
    df = pd.read_csv(path, index_col=0, header=0) if has_header else pd.read_csv(path, index_col=0, header=None, names=names)
    return {k: df[k].values for k in names}


def grid_from_wsi_visium(fullres_imgfile, spaceranger_dir, patch_size=256, window_size=256,
                         preprocess_xform=None, device=None, out_dtype=torch.float32, return_device='cpu'):
    """Drop-in for the reference function; extra keyword arguments choose where the result lives."""
    from PIL import Image
    Image.MAX_IMAGE_PIXELS = None
    img = np.array(Image.open(fullres_imgfile))
    if img.ndim != 3 or img.shape[2] < 3:
        raise ValueError('expected an RGB image')
    img = np.ascontiguousarray(img[:, :, :3])
    ydim, xdim = img.shape[:2]
    w = _window(patch_size, window_size, xdim)
    if 2 * (w // 2) != patch_size:
        raise NotImplementedError('B200 gather implements the pure-crop case 2*(window//2) == patch_size')
    mean, std = _normalize_params(preprocess_xform)
    dev = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
    pos = read_positions(spaceranger_dir)
    cells, dropped = spot_table(pos['in_tissue'], pos['array_row'], pos['array_col'], pos['pxl_row_in_fullres'],
                                pos['pxl_col_in_fullres'], dev)
    out = gather_patches(torch.from_numpy(img).to(dev), cells, patch_size, mean, std, out_dtype)
    nd = int(dropped.item())
    if nd:
        print("Warning: %d spots outside bounds of Visium array" % nd)
    return out.to(return_device) if return_device is not None else out
