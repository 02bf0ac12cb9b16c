"""CPU restatement of the reference's dataset tensor assembly (TEST INFRASTRUCTURE ONLY: imported by tests/, never by the product).

Follows, loop for loop:
  read_annotated_starray   /root/reference/gridnext/utils.py:144-166 (+ the channels-first permute of count_datasets.py:292-293)
  PatchGridDataset item    /root/reference/gridnext/image_datasets.py:205-232
  MultiModalGridDataset    /root/reference/gridnext/multimodal_datasets.py:237-244
Pinned against the real ``read_annotated_starray`` run on synthetic files (tests/golden/a1_starray.npz, oracle/make_golden.py).
"""
import numpy as np


def pseudo_hex_to_oddr(col, row):
    """utils.py:64-70."""
    if row % 2 == 0:
        x = col / 2
    else:
        x = (col - 1) / 2
    return int(x), int(row)


def count_grid(cmat, coord_strs, adict=None, h_st=78, w_st=64, visium=True):
    """cmat: (n_genes, n_spots) array; coord_strs: 'x_y' per column; adict: {coord_str: int label} or None (utils.py:144-166)."""
    n_genes = cmat.shape[0]
    counts_grid = np.zeros((h_st, w_st, n_genes), dtype=float)
    annots_grid = np.zeros((h_st, w_st), dtype=int)
    for j, cstr in enumerate(coord_strs):
        if visium:
            x_vis, y_vis = map(int, cstr.split('_'))
            x, y = pseudo_hex_to_oddr(x_vis, y_vis)
        else:
            x_car, y_car = map(float, cstr.split('_'))
            x, y = int(np.rint(x_car)), int(np.rint(y_car))
        if adict is not None:
            if cstr in adict:
                counts_grid[y, x] = cmat[:, j]
                annots_grid[y, x] = adict[cstr] + 1
        else:
            counts_grid[y, x] = cmat[:, j]
            annots_grid[y, x] = 0
    return np.transpose(counts_grid, (2, 0, 1)).astype(np.float32), annots_grid.astype(np.int64)


def patch_grid(patches, coords, adict=None, h_st=78, w_st=64, visium=True):
    """patches: (n, C, h, w); coords: [(a_x, a_y)] (image_datasets.py:205-232)."""
    grid = np.zeros((h_st, w_st) + patches.shape[1:], dtype=patches.dtype)
    annots = np.zeros((h_st, w_st), dtype=np.int64)
    for i, (a_x, a_y) in enumerate(coords):
        x, y = pseudo_hex_to_oddr(a_x, a_y) if visium else (a_x, a_y)
        if adict is not None:
            cstr = '%d_%d' % (a_x, a_y)
            if cstr in adict:
                annots[y, x] = adict[cstr] + 1
        grid[y, x] = patches[i]
    return grid, annots


def mm_fg_consistency(counts_grid, patch_grid_, annots_grid):
    """multimodal_datasets.py:237-244 on copies."""
    c, p, a = counts_grid.copy(), patch_grid_.copy(), annots_grid.copy()
    H, W = a.shape
    for i in range(H):
        for j in range(W):
            if p[i, j].max() == 0:
                a[i, j] = 0
                c[:, i, j] = 0
            if a[i, j] == 0:
                p[i, j] = 0
    return c, p, a
