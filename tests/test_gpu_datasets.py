"""GPU parity of the on-device dataset tensor assembly against the reference-generated golden
grid and the CPU restatement of the reference's loops.  Pure data movement: bit-exact."""
import os
import numpy as np
import pytest
import torch

from oracle import datasets_ref as D
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_count_grid_matches_reference_golden():
    from gridnext_b200 import datasets as ds
    gold = np.load(os.path.join(GOLDEN, 'a1_starray.npz'))
    coords = gold['coords']
    adict = dict(zip(gold['annot_coords'].tolist(), gold['annot_lbls'].tolist()))
    cstrs = ['%d_%d' % (c, r) for c, r in coords]
    cells = ds.spot_cells(coords[:, 0], coords[:, 1])
    labels = torch.tensor([adict.get(s, -1) for s in cstrs], dtype=torch.int64)
    cells = torch.where(labels >= 0, cells, torch.full_like(cells, -1))         # only annotated spots are included (utils.py:157-158)
    counts, annots = ds.assemble_count_grid(torch.from_numpy(gold['cmat']).cuda(), cells.cuda(), labels.cuda())
    assert counts.dtype == torch.float32 and tuple(counts.shape) == (12, 78, 64)
    assert np.array_equal(counts.cpu().numpy(), gold['counts_annot'])
    assert np.array_equal(annots.cpu().numpy(), gold['annots_annot'])


@pytest.mark.parametrize('dtype,shape,h_st,w_st,visium', [(torch.float32, (3, 16, 16), 78, 64, True), (torch.uint8, (3, 7, 5), 78, 64, True),
                                                         (torch.float32, (1, 3, 3), 9, 11, False), (torch.bfloat16, (3, 8, 8), 20, 12, False)])
def test_patch_grid_matches_oracle(dtype, shape, h_st, w_st, visium):
    from gridnext_b200 import datasets as ds
    rng = np.random.RandomState(5)
    if visium:
        allc = [(c, r) for r in range(h_st) for c in range(r % 2, 2 * w_st, 2)]
    else:
        allc = [(c, r) for r in range(h_st) for c in range(w_st)]
    n = min(len(allc), 150)
    sel = rng.choice(len(allc), n, replace=False)
    coords = [allc[i] for i in sel] + [allc[sel[0]]]            # a duplicate spot: the later one wins, as in the reference loop
    n += 1
    if dtype == torch.uint8:
        patches = torch.from_numpy(rng.randint(0, 256, (n,) + shape).astype(np.uint8))
    else:
        patches = torch.from_numpy(rng.randn(n, *shape).astype(np.float32)).to(dtype)
    lbl = rng.randint(-1, 5, n)
    lbl[-1] = lbl[0]
    adict = {'%d_%d' % c: int(l) for c, l in zip(coords, lbl) if l >= 0}
    ref_np = patches.float().numpy() if dtype == torch.bfloat16 else patches.numpy()
    ref_grid, ref_annots = D.patch_grid(ref_np, coords, adict, h_st, w_st, visium)
    cells = ds.spot_cells([c[0] for c in coords], [c[1] for c in coords], visium, h_st, w_st)
    grid, annots = ds.assemble_patch_grid(patches.cuda(), cells.cuda(), torch.from_numpy(lbl).cuda(), h_st, w_st)
    assert grid.dtype == dtype and tuple(grid.shape) == (h_st, w_st) + shape
    got = grid.float().cpu().numpy() if dtype == torch.bfloat16 else grid.cpu().numpy()
    assert np.array_equal(got, ref_grid)
    assert np.array_equal(annots.cpu().numpy(), ref_annots)


def test_multimodal_fg_consistency_matches_oracle():
    from gridnext_b200 import datasets as ds
    rng = np.random.RandomState(9)
    H, W, G = 78, 64, 17
    counts = rng.rand(G, H, W).astype(np.float32)
    patches = rng.randn(H, W, 3, 6, 6).astype(np.float32)
    annots = rng.randint(0, 4, (H, W)).astype(np.int64)
    patches[rng.rand(H, W) < 0.3] = 0                           # no image data
    patches[1, 1] = -np.abs(patches[1, 1]); patches[1, 1, 0, 0, 0] = 0   # all <= 0 with max exactly 0: counts as empty, like .max() == 0
    patches[2, 2] = -1.0                                         # max < 0: has data
    patches[3, 3, 1, 2, 3] = np.nan; annots[3, 3] = 2            # NaN: max() is NaN, != 0 -> kept
    c_ref, p_ref, a_ref = D.mm_fg_consistency(counts, patches, annots)
    c, p, a = ds.multimodal_fg_consistency(torch.from_numpy(counts).cuda(), torch.from_numpy(patches).cuda(), torch.from_numpy(annots).cuda())
    assert np.array_equal(a.cpu().numpy(), a_ref)
    assert np.array_equal(c.cpu().numpy(), c_ref)
    assert np.array_equal(p.cpu().numpy(), p_ref, equal_nan=True)


def test_dataset_assembly_rejects_cpu_tensors_and_bad_coordinates():
    from gridnext_b200 import datasets as ds
    with pytest.raises(RuntimeError):
        ds.assemble_count_grid(torch.zeros(2, 3), torch.zeros(3, dtype=torch.int32))
    with pytest.raises(IndexError):
        ds.spot_cells([400], [3])
