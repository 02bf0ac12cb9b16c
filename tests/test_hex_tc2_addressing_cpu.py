"""CPU check of the ADDRESSING of the second-generation tensor-core hex kernels (no GPU, no numerics of the tensor core).

csrc/hexconv_wgrad_tc2.cu computes the weight gradient of hexagdly.Conv2d (kernel_size 1) as a GEMM over cells: operand rows are
cells, an x row slot holds 72 cells (4 zero cells either side of the 64 columns), the neighbourhood's column shifts are operand START
cells (same row: cell 3 = column -1, three taps stacked one cell apart; rows above / below: cell 3 on even rows, 4 on odd rows, two
taps), row shifts are other slots of a ring of row pairs, and every CTA walks a contiguous range of row pairs with the job sequence
of `W2Gen`.  This file restates exactly that bookkeeping in numpy (float64) and compares the result with autograd through the oracle
(oracle/hexconv_ref.py: hexconv_visium, tap order of tap_table) -- so the tap -> start-cell table, the parity rule, the halo pairs at
range starts / array boundaries and the ring-slot arithmetic are pinned without a GPU.  The same is done for the contiguous tile
ranges (`H2Seg`) of the forward kernel csrc/hexconv_tc2.cu.  The GPU parity of the kernels themselves is tests/test_gpu_corrector.py.
"""
import numpy as np
import pytest
import torch

from oracle import hexconv_ref as R

XR, DR = 5, 3          # W2_XR, W2_DR: ring depths in row pairs


class W2Gen:
    """The job sequence of one CTA (struct W2Gen in csrc/hexconv_wgrad_tc2.cu): type 0 = x pair, 1 = dY pair."""

    def __init__(self, g0, g1, npa):
        self.g, self.g0, self.g1, self.npa = g0, g0, g1, npa
        self.b, self.p = divmod(g0, npa)
        self.sub, self.xi, self.di = 0, 0, 0

    def next(self):
        if self.g >= self.g1:
            return None
        start = self.g == self.g0 or self.p == 0
        if self.sub < 2 and not start:
            self.sub = 2
        b = self.b
        if self.sub == 0:
            job = (0, b, self.p - 1, self.xi); self.xi += 1; self.sub = 1
        elif self.sub == 1:
            job = (0, b, self.p, self.xi); self.xi += 1; self.sub = 2
        elif self.sub == 2:
            job = (0, b, self.p + 1, self.xi); self.xi += 1; self.sub = 3
        else:
            job = (1, b, self.p, self.di); self.di += 1; self.sub = 0
            self.g += 1
            self.p += 1
            if self.p == self.npa:
                self.p, self.b = 0, self.b + 1
        return job


def emulate_wgrad_tc2(x, dy, n_ctas):
    """x (B, Cin, H, W), dy (B, Cout, H, W) float64 -> dWp [7][Cin][Cout], dbias [Cout], following the kernel's bookkeeping."""
    B, Cin, H, W = x.shape
    Cout = dy.shape[1]
    npa = (H + 1) // 2
    total = B * npa
    dW = np.zeros((7, Cin, Cout))
    db = np.zeros(Cout)
    seen = np.zeros(total, dtype=int)
    for cta in range(n_ctas):
        g0, g1 = total * cta // n_ctas, total * (cta + 1) // n_ctas
        gen = W2Gen(g0, g1, npa)
        xring = np.full((XR, 2, 72, Cin), np.nan)         # NaN: a slot read before it was written poisons the result
        dring = np.full((DR, 2, 64, Cout), np.nan)
        while True:
            job = gen.next()
            if job is None:
                break
            typ, b, pair, idx = job
            rows = np.zeros((2, 64, Cout if typ else Cin))
            for r in range(2):
                y = 2 * pair + r
                if 0 <= y < H:
                    rows[r, :W] = (dy if typ else x)[b, :, y, :].T
            if typ == 0:
                slot = xring[idx % XR]
                slot[:] = 0.0                                   # cells 0..3 and 68..71 are the zero cells
                slot[:, 4:68] = rows
                continue
            dring[idx % DR] = rows
            seen[b * npa + pair] += 1
            xi = gen.xi
            xs = [xring[(xi - 3) % XR], xring[(xi - 2) % XR], xring[(xi - 1) % XR]]
            d = dring[idx % DR]
            for r in range(2):
                same = xs[1][r]
                up = xs[1][0] if r else xs[0][1]
                dn = xs[2][0] if r else xs[1][1]
                c_ud = 4 if r else 3
                a = d[r]                                        # [64 cells][Cout]
                for t in range(3):                              # stacked along N: group t starts one cell further on
                    dW[t] += same[3 + t:3 + t + 64].T @ a
                for t in range(2):
                    dW[3 + t] += up[c_ud + t:c_ud + t + 64].T @ a
                    dW[5 + t] += dn[c_ud + t:c_ud + t + 64].T @ a
                db += a.sum(0)
    assert (seen == 1).all()                                    # every dY row pair exactly once over all CTAs
    return dW, db


def oracle_wgrad(x, dy, cin, cout):
    ks = [torch.randn(s, dtype=torch.float64, requires_grad=True) for s in R.kernel_shapes(cin, cout, 1)]
    b = torch.zeros(cout, dtype=torch.float64, requires_grad=True)
    out = R.hexconv_visium(torch.from_numpy(x), ks, b)
    (out * torch.from_numpy(dy)).sum().backward()
    dW = np.stack([ks[i].grad[:, :, a, side].numpy().T for (i, a, side, *_rest) in R.tap_table(1)])
    return dW, b.grad.numpy()


@pytest.mark.parametrize('shape,n_ctas', [((3, 5, 4, 9, 8), 1), ((3, 5, 4, 9, 8), 4), ((2, 3, 3, 78, 64), 7), ((5, 2, 6, 4, 12), 148),
                                          ((1, 4, 4, 7, 64), 3), ((4, 1, 1, 2, 4), 5)])
def test_wgrad_tc2_addressing_matches_oracle(shape, n_ctas):
    B, cin, cout, H, W = shape
    rng = np.random.default_rng(B * 100 + H)
    x = rng.standard_normal((B, cin, H, W))
    dy = rng.standard_normal((B, cout, H, W))
    n_ctas = min(n_ctas, B * ((H + 1) // 2))                    # the launcher never starts more CTAs than row pairs
    dW, db = emulate_wgrad_tc2(x, dy, n_ctas)
    dW_ref, db_ref = oracle_wgrad(x, dy, cin, cout)
    assert np.isfinite(dW).all()
    assert np.abs(dW - dW_ref).max() < 1e-10 * max(1.0, np.abs(dW_ref).max())
    assert np.abs(db - db_ref).max() < 1e-10 * max(1.0, np.abs(db_ref).max())


def h2_segments(n_tiles, tpa, cta, n_ctas):
    """struct H2Seg of csrc/hexconv_tc2.cu: the (array, first row, tiles) segments of one CTA's contiguous tile range."""
    cur, t1 = n_tiles * cta // n_ctas, n_tiles * (cta + 1) // n_ctas
    out = []
    while cur < t1:
        b, p0 = divmod(cur, tpa)
        nt = min(t1 - cur, tpa - p0)
        out.append((b, 2 * p0, nt))
        cur += nt
    return out


@pytest.mark.parametrize('B,H,n_ctas', [(256, 78, 148), (40, 78, 148), (3, 27, 148), (17, 53, 148), (1, 4, 2), (5, 9, 7)])
def test_forward_tile_ranges_cover_every_tile_once(B, H, n_ctas):
    tpa = (H + 1) // 2
    n_tiles = B * tpa
    n_ctas = min(n_ctas, n_tiles)
    seen = np.zeros((B, tpa), dtype=int)
    loads = 0
    for cta in range(n_ctas):
        segs = h2_segments(n_tiles, tpa, cta, n_ctas)
        for b, y0, nt in segs:
            assert nt >= 1 and y0 % 2 == 0 and y0 // 2 + nt <= tpa          # a segment never crosses an array boundary
            seen[b, y0 // 2:y0 // 2 + nt] += 1
            loads += nt + 2                                                  # its row pairs incl. the two halo pairs
        counts = [nt for _, _, nt in segs]
        assert abs(sum(counts) - n_tiles / n_ctas) < 1.0 + 1e-9              # balanced to one tile
    assert (seen == 1).all()
    if (B, H) == (256, 78):
        # the C4 corner: at most 3 segments per CTA -> 8 % more row pairs loaded than tiles computed (the 26-row strips: 15 %)
        assert loads / n_tiles < 1.09


def emulate_fwd_tc2(x, ks):
    """The forward kernel's arithmetic per tile of two grid rows (y even, y + 1), restated: taps stacked along N, the column shifts applied
    as neighbour-LANE sums in the epilogue, and -- the last change -- the rows below accumulated into the SAME two column blocks as the
    rows above.  x (B, Cin, H, W) float64, ks = hexagdly kernels; returns (B, Cout, H, W)."""
    B, Cin, H, W = x.shape
    taps = R.tap_table(1)
    Wt = [ks[i][:, :, a, side].numpy().T for (i, a, side, *_r) in taps]          # [7] x (Cin, Cout)
    Cout = Wt[0].shape[1]
    out = np.zeros((B, Cout, H, W))
    xp = np.zeros((B, H + 4, W, Cin))                                            # rows -2 .. H+1, cell-major [x][channel]
    xp[:, 2:H + 2] = x.transpose(0, 2, 3, 1)
    for b in range(B):
        for y in range(0, H, 2):
            own = xp[b, y + 2:y + 4].reshape(2 * W, Cin)                         # the tile's 2 W accumulator rows: (row of the pair, x)
            up = xp[b, y + 1:y + 3].reshape(2 * W, Cin)                          # rows (y - 1, y): a whole-slot operand shift
            dn = xp[b, y + 3:y + 5].reshape(2 * W, Cin)                          # rows (y + 1, y + 2)
            E_l, E_c, E_r = own @ Wt[0], own @ Wt[1], own @ Wt[2]                # same row: N = 96
            a0 = up @ Wt[3] + dn @ Wt[5]                                         # tap a = 0 of the row above + below: one column block
            a1 = up @ Wt[4] + dn @ Wt[6]                                         # tap a = 1
            for par in range(2):
                if y + par >= H:
                    continue
                s = slice(par * W, (par + 1) * W)
                L, C, Rr = E_l[s].copy(), E_c[s].copy(), E_r[s].copy()
                if par == 0:                                                     # even row: columns x - 1, x
                    L += a0[s]; C += a1[s]
                else:                                                            # odd row: columns x, x + 1
                    C += a0[s]; Rr += a1[s]
                o = C.copy()                                                     # out[x] = L[x - 1] + C[x] + R[x + 1], zero padding at the ends
                o[1:] += L[:-1]
                o[:-1] += Rr[1:]
                out[b, :, y + par, :] = o.T
    return out


@pytest.mark.parametrize('shape', [(2, 3, 4, 9, 8), (1, 5, 2, 78, 64), (3, 2, 2, 4, 4)])
def test_forward_tc2_lane_sums_and_merged_rows_match_oracle(shape):
    B, cin, cout, H, W = shape
    g = torch.Generator(); g.manual_seed(H * 7 + W)
    ks = [torch.randn(s, dtype=torch.float64, generator=g) for s in R.kernel_shapes(cin, cout, 1)]
    x = torch.randn(B, cin, H, W, dtype=torch.float64, generator=g)
    ref = R.hexconv_visium(x, ks).numpy()
    got = emulate_fwd_tc2(x.numpy(), ks)
    assert np.abs(got - ref).max() < 1e-10 * max(1.0, np.abs(ref).max())
