"""Restatement of HexagDLy's hexagonal Conv2d (stride 1).  Test infrastructure only.

HexagDLy (PyPI "hexagdly", Steppa & Holch) is a dependency of the reference
(/root/reference/requirements.txt:11; used at gridnext/gridnet_models.py:10,130-147) that is
neither vendored nor pinned.  PARITY UNPINNED: no reference test or golden vector exists
for it.  Its published algorithm is restated here three ways that must agree:

1. ``hexconv_visium``      -- closed form in the Visium odd-r layout (B, C, 78, 64) that
                              GridNetHexOddr computes end to end (gridnet_models.py:173-187:
                              rot90+flip == transpose, so no re-indexing is needed).
2. ``hexconv_hexagdly``    -- the upstream composition of square convolutions in HexagDLy's own
                              layout (B, C, rows, cols) with odd 0-indexed columns shifted down
                              half a cell (in-tree statement of that convention:
                              gridnext/hexagdly_tools.py:68).
3. ``hexconv_cube_bruteforce`` -- a slow loop over cube-coordinate hexagon neighbourhoods.

Parameters follow upstream: for i = 0..k ``kernel_i`` has shape
(Cout, Cin, 2k+1-i, 1 if i == 0 else 2); last index 0 addresses column j-i, 1 column j+i;
row tap ``a`` addresses row r + a - (k - i//2) + (i&1)*(j&1).  ``bias`` is added once.
"""
import numpy as np
import torch
import torch.nn.functional as F


def n_taps(k):
    return 1 + 3 * k * (k + 1)


def kernel_shapes(cin, cout, k):
    return [(cout, cin, 2 * k + 1 - i, 1 if i == 0 else 2) for i in range(k + 1)]


def tap_table(k):
    """[(i, a, side, d_major, d_minor_even, d_minor_odd)] for every tap.

    ``d_major`` is the offset along the axis whose parity matters (HexagDLy column j ==
    Visium row y); ``d_minor_*`` the offset along the other axis (HexagDLy row r == Visium
    column x) for even / odd parity of the *output* cell's major index.
    """
    taps = []
    for i in range(k + 1):
        for side in range(1 if i == 0 else 2):
            for a in range(2 * k + 1 - i):
                dmaj = 0 if i == 0 else (i if side == 1 else -i)
                base = a - (k - i // 2)
                taps.append((i, a, side, dmaj, base, base + (i & 1)))
    assert len(taps) == n_taps(k)
    return taps


def hexconv_visium(x, kernels, bias=None):
    """Closed form in Visium layout. x: (B, Cin, H, W); parity is taken on the row index y."""
    k = len(kernels) - 1
    B, Cin, H, W = x.shape
    Cout = kernels[0].shape[0]
    P = k + 1
    xp = F.pad(x, (P, P, P, P))
    out = x.new_zeros((B, Cout, H, W))
    even_rows = torch.arange(0, H, 2)
    odd_rows = torch.arange(1, H, 2)
    for (i, a, side, dy, dxe, dxo) in tap_table(k):
        w = kernels[i][:, :, a, side]  # (Cout, Cin)
        for rows, dx in ((even_rows, dxe), (odd_rows, dxo)):
            if len(rows) == 0:
                continue
            src = xp[:, :, rows + P + dy][:, :, :, P + dx:P + dx + W]  # (B, Cin, nrows, W)
            out[:, :, rows] = out[:, :, rows] + torch.einsum('oc,bcyx->boyx', w, src)
    if bias is not None:
        out = out + bias.view(1, -1, 1, 1)
    return out


def hexconv_hexagdly(x, kernels, bias=None):
    """Upstream composition in HexagDLy layout. x: (B, Cin, R, Wc); parity on column j."""
    k = len(kernels) - 1
    Wc = x.shape[-1]
    # i = 0: vertical (same column) taps, carries the bias
    out_all = F.conv2d(F.pad(x, (0, 0, k, k)), kernels[0], bias=bias)
    odd = even = None  # upstream naming: "odd" = 1-indexed odd = 0-indexed even columns
    for i in range(1, k + 1):
        if i % 2 == 0:
            t = k - i // 2
            out_all = out_all + F.conv2d(F.pad(x, (i, i, t, t)), kernels[i], dilation=(1, 2 * i))
        else:
            top, bot = k - i // 2, k - (i + 1) // 2
            # 0-indexed even output columns 0,2,4,...: ceil(Wc/2) of them
            n_e = (Wc + 1) // 2
            need_e = 2 * (n_e - 1) + 2 * i + 1  # padded width needed
            xe = F.pad(x, (i, max(0, need_e - Wc - i), top, bot))
            ce = F.conv2d(xe, kernels[i], dilation=(1, 2 * i), stride=(1, 2))[..., :n_e]
            # 0-indexed odd output columns 1,3,5,...: floor(Wc/2) of them
            n_o = Wc // 2
            if n_o > 0:
                need_o = 2 * (n_o - 1) + 2 * i + 1
                # centre column j = 2m+1 -> leftmost tap j-i = 2m + (1-i): pad left by i-1
                xo = F.pad(x, (i - 1, max(0, need_o - Wc - (i - 1)), top - 1, bot + 1))
                co = F.conv2d(xo, kernels[i], dilation=(1, 2 * i), stride=(1, 2))[..., :n_o]
            else:
                co = x.new_zeros(ce.shape[:-1] + (0,))
            odd = ce if odd is None else odd + ce
            even = co if even is None else even + co
    if odd is not None:
        order = torch.empty(Wc, dtype=torch.long)
        order[0::2] = torch.arange(odd.shape[-1])
        order[1::2] = torch.arange(even.shape[-1]) + odd.shape[-1]
        out_all = out_all + torch.cat((odd, even), 3)[..., order]
    return out_all


def hexconv_cube_bruteforce(x, kernels, bias=None):
    """O(cells^2) check in HexagDLy layout (numpy, float64): neighbourhood = hex ball of radius k."""
    x = np.asarray(x, dtype=np.float64)
    ks = [np.asarray(t, dtype=np.float64) for t in kernels]
    k = len(ks) - 1
    B, Cin, R, Wc = x.shape
    Cout = ks[0].shape[0]
    out = np.zeros((B, Cout, R, Wc))

    def cube(r, j):
        q = j
        rr = r - (j - (j & 1)) // 2
        return q, rr

    for r in range(R):
        for j in range(Wc):
            q0, r0 = cube(r, j)
            for r2 in range(R):
                for j2 in range(Wc):
                    q1, r1 = cube(r2, j2)
                    dq, dr = q1 - q0, r1 - r0
                    dist = (abs(dq) + abs(dr) + abs(dq + dr)) // 2
                    if dist > k:
                        continue
                    i = abs(j2 - j)
                    side = 1 if j2 > j else 0
                    a = (r2 - r) + (k - i // 2) - (i & 1) * (j & 1)
                    assert 0 <= a < 2 * k + 1 - i, (r, j, r2, j2, a)
                    out[:, :, r, j] += x[:, :, r2, j2] @ ks[i][:, :, a, side].T
    if bias is not None:
        out += np.asarray(bias, dtype=np.float64).reshape(1, -1, 1, 1)
    return out
