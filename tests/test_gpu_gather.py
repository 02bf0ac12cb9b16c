"""GPU parity of the spot-patch gather: bit-exact against the numpy oracle and the reference golden."""
import json, os
import numpy as np
import pytest
import torch

from oracle import synth, gather_ref
from conftest import GOLDEN

pytestmark = pytest.mark.gpu
MAN = json.load(open(os.path.join(GOLDEN, 'manifest.json')))
MEAN, STD = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]


def run_gpu(img, pos, P, mean=None, std=None, dtype=torch.float32):
    from gridnext_b200 import imgprocess as ip
    tis, rows, cols, pr, pc = pos
    cells, dropped = ip.spot_table(tis, rows, cols, pr, pc, 'cuda:0')
    out = ip.gather_patches(torch.from_numpy(img).cuda(), cells, P, mean, std, dtype)
    return out, int(dropped.item()), cells


def test_gather_matches_reference_golden():
    m = MAN['p1_gather_p16']
    gold = np.load(os.path.join(GOLDEN, 'p1_gather_p16.npz'))
    pos = synth.synth_positions(pitch_col=m['pitch_col'], pitch_row=m['pitch_row'], org_row=m['org_row'], org_col=m['org_col'])
    img = synth.synth_image(m['Himg'], m['Wimg'], seed=m['img_seed'], smooth=True)
    raw, dropped, _ = run_gpu(img, pos, 16)
    assert dropped == 0
    assert np.array_equal(raw.cpu().numpy(), gold['raw'].astype(np.float32))
    nrm, _, _ = run_gpu(img, pos, 16, MEAN, STD)
    assert np.array_equal(nrm.cpu().numpy()[::11, ::9], gold['nrm_sub'])        # bit-exact fp32


@pytest.mark.parametrize('P,Himg,Wimg,seed', [(128, 1500, 1700, 1), (32, 400, 377, 2), (64, 333, 1001, 3), (256, 700, 900, 4),
                                              (128, 1500, 1600, 5), (64, 500, 1024, 6), (256, 800, 1008, 7)])   # 16-byte pitch: TMA path
def test_gather_matches_oracle_bit_exact(P, Himg, Wimg, seed):
    # pitch chosen so that spots hang over all four image borders (edge clamp) and odd byte alignments occur
    pos = synth.synth_positions(pitch_col=Wimg / 130.0, pitch_row=Himg / 79.0, org_row=1.0, org_col=1.5)
    img = synth.synth_image(Himg, Wimg, seed=seed)
    ref = gather_ref.grid_from_image(img, *pos, patch_size=P, window_size=P, mean=MEAN, std=STD)
    out, dropped, cells = run_gpu(img, pos, P, MEAN, STD)
    assert dropped == 0
    # index table: bit-exact spot ordering / rint / //2
    tab = gather_ref.spot_table(*pos)
    c = cells.cpu().numpy().reshape(78, 64, 3)
    assert int(c[:, :, 2].sum()) == len(tab)
    for x_ind, y_ind, x_px, y_px in tab[::97]:
        assert tuple(c[y_ind, x_ind]) == (x_px, y_px, 1)
    assert np.array_equal(out.cpu().numpy(), ref)
    ob, _, _ = run_gpu(img, pos, P, MEAN, STD, torch.bfloat16)
    assert torch.equal(ob.cpu(), torch.from_numpy(ref).to(torch.bfloat16))       # RNE of the exact fp32 value
    raw, _, _ = run_gpu(img, pos, P)
    assert np.array_equal(raw.cpu().numpy(), gather_ref.grid_from_image(img, *pos, patch_size=P, window_size=P))


def test_gather_drops_out_of_array_spots_and_rejects_bad_args():
    from gridnext_b200 import imgprocess as ip
    tis = np.array([1, 1, 1, 0]); rows = np.array([0, 78, 1, 2]); cols = np.array([0, 0, 129, 4])
    pr = np.array([5.5, 6.5, 7.0, 8.0]); pc = np.array([2.5, 3.5, 4.0, 5.0])
    cells, dropped = ip.spot_table(tis, rows, cols, pr, pc, 'cuda:0')
    assert int(dropped.item()) == 2           # row 78 and x_ind 64 are outside the 78 x 64 array
    c = cells.cpu().numpy().reshape(78, 64, 3)
    assert tuple(c[0, 0]) == (2, 6, 1)        # rint: 2.5 -> 2, 5.5 -> 6 (half to even)
    assert int(c[:, :, 2].sum()) == 1
    img = torch.zeros(64, 64, 3, dtype=torch.uint8, device='cuda')
    with pytest.raises(ValueError):
        ip.gather_patches(img, cells, 6)      # patch size not a multiple of 4
    with pytest.raises(RuntimeError):
        ip.gather_patches(img.cpu(), cells, 8)


# ---- round 2: window_size != patch_size and general transforms ---------------------------------------------------------------
def test_gather_with_resize_matches_reference_golden():
    """The resize kernel (Pillow's fixed-point BICUBIC restated in CUDA) against vectors from the reference's own
    grid_from_wsi_visium: window 24 -> patch 16, window 10 -> patch 16, float window 0.03 * width -> patch 12.  Bit-exact."""
    from gridnext_b200 import imgprocess as ip
    m = MAN['p2_gather_resize']
    gold = np.load(os.path.join(GOLDEN, 'p2_gather_resize.npz'))
    pos = synth.synth_positions(pitch_col=m['pitch_col'], pitch_row=m['pitch_row'], org_row=m['org_row'], org_col=m['org_col'])
    img = synth.synth_image(m['Himg'], m['Wimg'], seed=m['img_seed'], smooth=True)
    cells, _ = ip.spot_table(*pos, 'cuda:0')
    img_d = torch.from_numpy(img).cuda()
    sel = np.ix_(gold['cells_y'], gold['cells_x'])
    for key, P, w in (('down_24_to_16', 16, 24), ('up_10_to_16', 16, 10), ('float_0.03_to_12', 12, ip._window(12, 0.03, m['Wimg']))):
        out = ip.gather_patches(img_d, cells, P, window=2 * (w // 2))
        assert np.array_equal(out.cpu().numpy()[sel], gold[key].astype(np.float32)), key
    nrm = ip.gather_patches(img_d, cells, 16, MEAN, STD, window=24)
    assert np.array_equal(nrm.cpu().numpy()[::11, ::9], gold['down_24_to_16_nrm_sub'])


@pytest.mark.parametrize('ws,P,Himg,Wimg', [(256, 224, 1500, 1700), (64, 32, 400, 377), (100, 128, 700, 900), (512, 224, 1200, 1300), (30, 48, 300, 310)])
def test_gather_with_resize_matches_oracle_bit_exact(ws, P, Himg, Wimg):
    """Tutorial-sized windows (256 -> 224 is the tutorials' Resize/CenterCrop scale) incl. spots hanging over every border."""
    from gridnext_b200 import imgprocess as ip
    pos = synth.synth_positions(pitch_col=Wimg / 130.0, pitch_row=Himg / 79.0, org_row=1.0, org_col=1.5)
    img = synth.synth_image(Himg, Wimg, seed=ws + P)
    keep = np.zeros(len(pos[0]), bool)
    keep[::23] = True                                   # the numpy oracle resizes patch by patch: a sample of spots, all borders included
    keep[:70] = True; keep[-70:] = True
    tis = np.where(keep, pos[0], 0)
    ref = gather_ref.grid_from_image(img, tis, *pos[1:], patch_size=P, window_size=ws, mean=MEAN, std=STD)
    cells, _ = ip.spot_table(tis, *pos[1:], 'cuda:0')
    out = ip.gather_patches(torch.from_numpy(img).cuda(), cells, P, MEAN, STD, window=ws)
    assert np.array_equal(out.cpu().numpy(), ref)
    ob = ip.gather_patches(torch.from_numpy(img).cuda(), cells, P, MEAN, STD, torch.bfloat16, window=ws)
    assert torch.equal(ob.cpu(), torch.from_numpy(ref).to(torch.bfloat16))


def test_grid_from_wsi_visium_general_transform_and_resize(tmp_path):
    """The drop-in function end to end on files: a non-Normalize transform (Grayscale + Normalize, reference golden) and a
    resizing window; reference: imgprocess.py:162-238."""
    from PIL import Image
    from torchvision import transforms
    from gridnext_b200 import imgprocess as ip
    m = MAN['p2_gather_resize']
    gold = np.load(os.path.join(GOLDEN, 'p2_gather_resize.npz'))
    tis, rows, cols, pr, pc = synth.synth_positions(pitch_col=m['pitch_col'], pitch_row=m['pitch_row'], org_row=m['org_row'], org_col=m['org_col'])
    img = synth.synth_image(m['Himg'], m['Wimg'], seed=m['img_seed'], smooth=True)
    Image.fromarray(img).save(str(tmp_path / 'img.png'))
    sp = tmp_path / 'outs' / 'spatial'
    sp.mkdir(parents=True)
    with open(sp / 'tissue_positions.csv', 'w') as fh:
        fh.write('barcode,in_tissue,array_row,array_col,pxl_row_in_fullres,pxl_col_in_fullres\n')
        for i in range(len(tis)):
            fh.write('BC%05d-1,%d,%d,%d,%r,%r\n' % (i, tis[i], rows[i], cols[i], float(pr[i]), float(pc[i])))
    xf = transforms.Compose([transforms.Grayscale(num_output_channels=3), transforms.Normalize([0.5] * 3, [0.25] * 3)])
    got = ip.grid_from_wsi_visium(str(tmp_path / 'img.png'), str(tmp_path / 'outs'), patch_size=16, window_size=16, preprocess_xform=xf)
    ref = gold['crop_16_gray_nrm_sub']
    assert got.shape == (78, 64, 3, 16, 16) and got.dtype == torch.float32
    assert np.abs(got.numpy()[::11, ::9] - ref).max() < 1e-5          # Grayscale is a float weighted sum: tensor vs PIL rounding differs in the last bits
    down = ip.grid_from_wsi_visium(str(tmp_path / 'img.png'), str(tmp_path / 'outs'), patch_size=16, window_size=24)
    assert np.array_equal(down.numpy()[np.ix_(gold['cells_y'], gold['cells_x'])], gold['down_24_to_16'].astype(np.float32))
