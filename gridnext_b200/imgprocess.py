"""Spot-patch gather from a full-resolution Visium H&E image.

Mirrors /root/reference/gridnext/imgprocess.py: constants (:21-22), ``pseudo_hex_to_oddr`` /
``oddr_to_pseudo_hex`` (:26-38) and ``grid_from_wsi_visium`` (:162-238).  The per-spot python loop
(crop, /255, Normalize, scatter into (78, 64, 3, P, P)) is one CUDA kernel (csrc/patch_gather.cu);
the integer index math (pseudo-hex -> odd-r, rint) runs on the device as well.  ``window_size != patch_size`` (imgprocess.py:188-195,
221: ``Image.fromarray(patch).resize((P, P))``, Pillow's fixed-point BICUBIC) runs through csrc/patch_resize.cu, bit-exact
against Pillow; a ``preprocess_xform`` other than a bare ``Normalize`` is applied to the gathered patches on the device.
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
    """(mean, std) of a torchvision ``Normalize`` (optionally wrapped in a Compose of exactly that), which the gather kernels
    fuse; ``None`` for any other transform (applied to the gathered patches afterwards)."""
    if preprocess_xform is None:
        return None, None
    x = preprocess_xform
    tr = getattr(x, 'transforms', None)
    if tr is not None and len(tr) == 1:
        x = tr[0]
    if type(x).__name__ == 'Normalize' and hasattr(x, 'mean') and hasattr(x, 'std') and len(x.mean) == 3:
        return [float(v) for v in x.mean], [float(v) for v in x.std]
    return None


# ---- Pillow's 8-bit BICUBIC resampling tables (src/libImaging/Resample.c: bicubic_filter, precompute_coeffs, normalize_coeffs_8bpc)
_PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    if x < 0.0:
        x = -x
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


_RESIZE_TABLES = {}


def pillow_bicubic_table(in_size, out_size):
    """-> (bounds int32 [out, 2] = (first tap, tap count), coefficients int32 [out, ksize] with 22 fractional bits, ksize,
    widest tap count): what Pillow's ``Image.resize`` (BICUBIC, box = whole image) uses along one axis, in the same double
    arithmetic and the same order of operations."""
    key = (int(in_size), int(out_size))
    if key in _RESIZE_TABLES:
        return _RESIZE_TABLES[key]
    scale = float(in_size) / out_size
    filterscale = scale if scale >= 1.0 else 1.0
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [_bicubic((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << _PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << _PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    out = (bounds, kk, ksize, int(bounds[:, 1].max()))
    _RESIZE_TABLES[key] = out
    return out


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


_TABLE_DEV = {}


def gather_patches(img, cells, patch_size, mean=None, std=None, out_dtype=torch.float32, h_st=VISIUM_H_ST, w_st=VISIUM_W_ST, out=None,
                   window=None):
    """img: CUDA uint8 (H, W, 3) contiguous; cells from ``spot_table``.  -> (h_st, w_st, 3, P, P).
    ``window``: side of the square cut out around every spot (even); when it differs from ``patch_size`` the window is
    resized with Pillow's BICUBIC arithmetic (bit-exact)."""
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
    if window is not None and int(window) != P:
        ws = int(window)
        if ws < 2 or ws % 2:
            raise ValueError('gather_patches: window must be an even number of pixels >= 2')
        key = (ws, P, str(img.device))
        tab = _TABLE_DEV.get(key)
        if tab is None:
            bounds, kk, ksize, span = pillow_bicubic_table(ws, P)
            tab = (torch.from_numpy(bounds).to(img.device), torch.from_numpy(kk).to(img.device), ksize, span)
            _TABLE_DEV[key] = tab
        call('gn_patch_gather_resize', ptr(img), 3 * W, H, W, ptr(cells), h_st * w_st, ws, P, ptr(tab[0]), ptr(tab[1]), tab[2], tab[3],
             ptr(m), ptr(s), ptr(out), 1 if out_dtype == torch.bfloat16 else 0, stream())
        return out
    call('gn_patch_gather', ptr(img), 3 * W, H, W, ptr(cells), h_st * w_st, P, ptr(m), ptr(s), ptr(out),
         1 if out_dtype == torch.bfloat16 else 0, stream())
    return out


def apply_patch_transform(grid, cells, preprocess_xform):
    """The reference's per-patch ``Compose([ToPILImage(), ToTensor(), preprocess_xform])`` (imgprocess.py:224-230) for an
    arbitrary transform: raw uint8-valued patches of the in-tissue cells are scaled to [0, 1] (ToTensor) and handed to
    ``preprocess_xform`` as ONE device batch (n, 3, P, P) (torchvision's tensor transforms broadcast over leading dimensions);
    transforms that refuse a batch are applied patch by patch.  Like the reference, the result must keep the patch shape."""
    h_st, w_st = grid.shape[:2]
    flat = grid.view(h_st * w_st, *grid.shape[2:])
    valid = cells.view(-1, 3)[:, 2].bool()
    idx = valid.nonzero(as_tuple=True)[0]
    if idx.numel() == 0:
        return grid
    x = flat[idx].float() / 255.0
    try:
        y = preprocess_xform(x)
        if tuple(y.shape) != tuple(x.shape):
            raise ValueError
    except Exception:
        y = torch.stack([preprocess_xform(t) for t in x])
    if tuple(y.shape) != tuple(x.shape):
        raise ValueError('preprocess_xform changed the patch shape %s -> %s' % (tuple(x.shape[1:]), tuple(y.shape[1:])))
    flat[idx] = y.to(grid.dtype)
    return grid


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
    side = 2 * (w // 2)                        # img[y - w//2 : y + w//2, ...] (imgprocess.py:220)
    if side < 2:
        raise ValueError('window of %d pixels is too small' % w)
    fused = _normalize_params(preprocess_xform)
    dev = torch.device(device if device is not None else 'cuda:%d' % torch.cuda.current_device())
    pos = read_positions(spaceranger_dir)
    cells, dropped = spot_table(pos['in_tissue'], pos['array_row'], pos['array_col'], pos['pxl_row_in_fullres'],
                                pos['pxl_col_in_fullres'], dev)
    img_d = torch.from_numpy(img).to(dev)
    if fused is not None:
        out = gather_patches(img_d, cells, patch_size, fused[0], fused[1], out_dtype, window=side)
    else:
        out = gather_patches(img_d, cells, patch_size, None, None, torch.float32, window=side)
        out = apply_patch_transform(out, cells, preprocess_xform).to(out_dtype)
    nd = int(dropped.item())
    if nd:
        print("Warning: %d spots outside bounds of Visium array" % nd)
    return out.to(return_device) if return_device is not None else out
