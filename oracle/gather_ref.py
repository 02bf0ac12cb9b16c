"""numpy restatement of the spot-patch gather (test infrastructure only).

Follows /root/reference/gridnext/imgprocess.py:185-238 (grid_from_wsi_visium) and the
index helper pseudo_hex_to_oddr (imgprocess.py:26-32 == utils.py:64-70):

  * w = patch_size | int(window_size * xdim) | window_size              (:188-195)
  * edge padding by w//2 on both image axes                             (:198)   == clamp-to-edge
  * only in_tissue == 1 spots                                           (:200-202)
  * x_ind = col//2 (even row) | (col-1)//2 (odd row), y_ind = row        (:26-32)
  * pixel centre = int(np.rint(.)) (round half to even)                 (:213-214)
  * patch = img[y-w//2 : y+w//2, x-w//2 : x+w//2]  (side 2*(w//2))      (:220)
  * resize to (P, P): identity when the side already equals P           (:221)
  * optional ToTensor (/255) + Normalize(mean, std) in float32          (:224-230)
  * cells with y_ind >= 78 or x_ind > 64 are skipped (x_ind == 64 raises IndexError upstream)
  * out-of-tissue cells stay exactly 0.0                                (:206)

When the window side 2*(w//2) differs from patch_size the reference resizes with ``Image.fromarray(patch).resize((P, P))``
(:221), i.e. Pillow's default BICUBIC.  Pillow (pinned by the installed wheel, 12.2.0; src/libImaging/Resample.c) resamples
8-bit images in fixed point -- per-axis taps [xmin, xmin+n) around (xx+0.5)*in/out with support 2*max(in/out, 1), bicubic
(a = -0.5) weights normalised in double and rounded to 22 fractional bits, a horizontal pass into a uint8 intermediate
(clip8((2^21 + sum) >> 22)) and then a vertical pass -- restated here in numpy (``pillow_resize``) and pinned against the real
``PIL.Image.resize`` in tests/test_oracle_golden.py and against reference-generated vectors (p2_gather_resize).
"""
import numpy as np

PRECISION_BITS = 32 - 8 - 2


def _bicubic(x):
    a = -0.5
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def pillow_coeffs(in_size, out_size):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the BICUBIC filter over the whole axis.
    -> bounds (out, 2) int64 [first tap, tap count], integer coefficients (out, ksize)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(np.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int64)
    kk = np.zeros((out_size, ksize), np.float64)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        ww = 0.0
        for x in range(xmax):
            w = _bicubic((x + xmin - center + 0.5) * ss)
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :xmax] /= ww
        bounds[xx] = (xmin, xmax)
    ik = np.trunc(np.where(kk < 0, -0.5 + kk * (1 << PRECISION_BITS), 0.5 + kk * (1 << PRECISION_BITS))).astype(np.int64)
    return bounds, ik


def _clip8(v):
    return np.clip(v >> PRECISION_BITS, 0, 255).astype(np.uint8)


def pillow_resize(img, P):
    """(h, w, c) uint8 -> (P, P, c) uint8, equal to np.array(Image.fromarray(img).resize((P, P))) (BICUBIC)."""
    h, w, c = img.shape
    if (h, w) == (P, P):
        return img.copy()
    tmp = img
    if w != P:
        bh, kh = pillow_coeffs(w, P)
        tmp = np.zeros((h, P, c), np.uint8)
        for xx in range(P):
            xmin, n = bh[xx]
            acc = (1 << (PRECISION_BITS - 1)) + (img[:, xmin:xmin + n, :].astype(np.int64) * kh[xx, :n][None, :, None]).sum(1)
            tmp[:, xx, :] = _clip8(acc)
    if h == P:
        return tmp
    bv, kv = pillow_coeffs(h, P)
    out = np.zeros((P, P, c), np.uint8)
    for yy in range(P):
        ymin, n = bv[yy]
        acc = (1 << (PRECISION_BITS - 1)) + (tmp[ymin:ymin + n].astype(np.int64) * kv[yy, :n][:, None, None]).sum(0)
        out[yy] = _clip8(acc)
    return out

VISIUM_H_ST = 78
VISIUM_W_ST = 64


def pseudo_hex_to_oddr(col, row):
    if row % 2 == 0:
        x = col / 2
    else:
        x = (col - 1) / 2
    return int(x), int(row)


def spot_table(in_tissue, array_row, array_col, pxl_row, pxl_col):
    """Integer per-spot table (x_ind, y_ind, x_px, y_px) for in-tissue spots, in file order."""
    rows = []
    for t, r, c, pr, pc in zip(in_tissue, array_row, array_col, pxl_row, pxl_col):
        if int(t) != 1:
            continue
        x_ind, y_ind = pseudo_hex_to_oddr(int(c), int(r))
        rows.append((x_ind, y_ind, int(np.rint(pc)), int(np.rint(pr))))
    return np.asarray(rows, dtype=np.int64).reshape(-1, 4)


def grid_from_image(img, in_tissue, array_row, array_col, pxl_row, pxl_col,
                    patch_size=256, window_size=256, mean=None, std=None,
                    h_st=VISIUM_H_ST, w_st=VISIUM_W_ST):
    """img: (H, W, 3) uint8.  Returns float32 (h_st, w_st, 3, P, P)."""
    ydim, xdim = img.shape[:2]
    if window_size is None:
        w = patch_size
    elif isinstance(window_size, float):
        w = int(window_size * xdim)
    elif isinstance(window_size, int):
        w = window_size
    else:
        raise ValueError("Window size must be a float or int")
    hw = w // 2
    out = np.zeros((h_st, w_st, 3, patch_size, patch_size), dtype=np.float32)
    for x_ind, y_ind, x_px, y_px in spot_table(in_tissue, array_row, array_col, pxl_row, pxl_col):
        ys = np.clip(np.arange(y_px - hw, y_px + hw), 0, ydim - 1)   # edge padding == clamp
        xs = np.clip(np.arange(x_px - hw, x_px + hw), 0, xdim - 1)
        patch = img[ys][:, xs]                                        # (2*hw, 2*hw, 3) u8
        if 2 * hw != patch_size:
            patch = pillow_resize(patch, patch_size)
        patch = np.transpose(patch, (2, 0, 1))
        if y_ind >= h_st or x_ind >= w_st:
            continue
        if mean is not None:
            p = patch.astype(np.float32) / np.float32(255.0)
            m = np.asarray(mean, dtype=np.float32).reshape(3, 1, 1)
            s = np.asarray(std, dtype=np.float32).reshape(3, 1, 1)
            out[y_ind, x_ind] = (p - m) / s
        else:
            out[y_ind, x_ind] = patch.astype(np.float32)
    return out
