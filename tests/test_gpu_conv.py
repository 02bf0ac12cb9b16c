"""GPU parity of the tcgen05 3x3 convolution (forward and data gradient with fused BN-ReLU backward)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rnd(shape, seed, scale=1.0):
    g = torch.Generator(); g.manual_seed(seed)
    return torch.randn(shape, generator=g) * scale


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.double().abs().max()), 1e-12))


SHAPES = [(3, 32, 32), (5, 16, 16), (9, 8, 8), (33, 4, 4), (7, 7, 7), (4, 2, 2), (2, 56, 56), (1, 14, 14)]


@pytest.mark.parametrize('Nimg,H,W', SHAPES)
@pytest.mark.parametrize('CI,CO', [(128, 32), (16, 16), (64, 48)])
def test_conv3x3_forward(Nimg, H, W, CI, CO):
    from gridnext_b200.tc import conv3x3_pack, conv3x3_bf16
    x = rnd((Nimg, H, W, CI), 1).to(torch.bfloat16)
    w = rnd((CO, CI, 3, 3), 2, (2.0 / (9 * CI)) ** 0.5)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1).reshape(-1, CO)
    xbuf = torch.zeros((Nimg * H * W, CI + 40), dtype=torch.bfloat16, device='cuda')
    xbuf[:, :CI] = x.reshape(-1, CI).cuda()
    obuf = torch.zeros((Nimg * H * W, CO + 64), dtype=torch.bfloat16, device='cuda')
    wp = conv3x3_pack(w.cuda(), 0)
    conv3x3_bf16(xbuf[:, :CI], Nimg, H, W, CI, wp, CO, obuf[:, 32:32 + CO])
    assert rel(obuf[:, 32:32 + CO].float().cpu(), ref) < 1e-2
    assert float(obuf[:, :32].abs().max()) == 0 and float(obuf[:, 32 + CO:].abs().max()) == 0


@pytest.mark.parametrize('Nimg,H,W', SHAPES)
@pytest.mark.parametrize('Cin,Cout', [(128, 32), (16, 16)])
def test_conv3x3_data_gradient_with_bn_relu_backward(Nimg, H, W, Cin, Cout):
    """dZ = conv3x3^T(dY, W) * [A2 > 0] * s2  and the BatchNorm parameter-gradient column sums."""
    from gridnext_b200.tc import conv3x3_pack, conv3x3_bf16
    M = Nimg * H * W
    w = rnd((Cout, Cin, 3, 3), 3, (2.0 / (9 * Cin)) ** 0.5)
    dy = rnd((Nimg, H, W, Cout), 4).to(torch.bfloat16)
    g = torch.Generator(); g.manual_seed(5)
    gamma, beta = torch.rand(Cin, generator=g) + 0.5, torch.randn(Cin, generator=g) * 0.2
    s2 = gamma / torch.sqrt(torch.rand(Cin, generator=g) + 0.5)
    a2 = torch.relu(rnd((M, Cin), 6) + 0.3).to(torch.bfloat16)            # saved activated bottleneck
    # reference
    dA = F.conv_transpose2d(dy.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), padding=1).permute(0, 2, 3, 1).reshape(M, Cin)
    gmask = dA * (a2.float() > 0)
    ref_dz = gmask * s2
    ref_sum_g = gmask.sum(0)
    ref_sum_gx = (gmask * (a2.float() - beta) / gamma).sum(0)
    # product
    dybuf = torch.zeros((M, Cout + 96), dtype=torch.bfloat16, device='cuda')     # dY is a column slice of the concat-gradient buffer
    dybuf[:, 64:64 + Cout] = dy.reshape(M, Cout).cuda()
    wpt = conv3x3_pack(w.cuda(), 1)
    dz = torch.empty((M, Cin), dtype=torch.bfloat16, device='cuda')
    colsum = torch.zeros((2, Cin), dtype=torch.float32, device='cuda')
    conv3x3_bf16(dybuf[:, 64:64 + Cout], Nimg, H, W, Cout, wpt, Cin, dz,
                 bn=dict(ref=a2.cuda(), ref_is_raw=False, sc=s2.cuda(), sh=None, p0=beta.cuda(), p1=(1.0 / gamma).cuda(), colsum=colsum))
    assert rel(dz.float().cpu(), ref_dz) < 1.5e-2
    assert rel(colsum[0].cpu(), ref_sum_g) < 1e-2
    assert rel(colsum[1].cpu(), ref_sum_gx) < 1e-2


@pytest.mark.parametrize('Nimg,H,W', SHAPES)
@pytest.mark.parametrize('CI,CO', [(128, 32), (16, 8), (64, 48)])
def test_conv3x3_weight_gradient(Nimg, H, W, CI, CO):
    from gridnext_b200.tc import conv3x3_wgrad_bf16
    M = Nimg * H * W
    x = torch.relu(rnd((Nimg, H, W, CI), 7)).to(torch.bfloat16)
    dy = rnd((Nimg, H, W, CO), 8).to(torch.bfloat16)
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(False)
    w = torch.zeros((CO, CI, 3, 3), requires_grad=True)
    F.conv2d(xr, w, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    xbuf = torch.zeros((M, CI + 8), dtype=torch.bfloat16, device='cuda'); xbuf[:, :CI] = x.reshape(M, CI).cuda()
    dbuf = torch.zeros((M, CO + 200), dtype=torch.bfloat16, device='cuda'); dbuf[:, 96:96 + CO] = dy.reshape(M, CO).cuda()
    dw = conv3x3_wgrad_bf16(xbuf[:, :CI], dbuf[:, 96:96 + CO], Nimg, H, W, CI, CO)
    assert rel(dw.cpu(), w.grad) < 1e-4


@pytest.mark.parametrize('N,P,CO', [(3, 128, 64), (5, 64, 64), (4, 32, 16), (2, 224, 64), (7, 48, 32), (2, 128, 96), (3, 40, 24)])
def test_stem_conv7x7s2_forward_and_weight_gradient(N, P, CO):
    """conv0 + norm0 + relu0 (densenet.py:107-109) and the conv0 weight gradient, read from NHWC4 patches without im2col."""
    from gridnext_b200 import tc
    x = rnd((N, 3, P, P), 31).to(torch.bfloat16)
    w = rnd((CO, 3, 7, 7), 32, (2.0 / 147) ** 0.5)
    g = torch.Generator(); g.manual_seed(33)
    sc, sh = torch.rand(CO, generator=g) + 0.5, torch.randn(CO, generator=g) * 0.3
    wr = w.to(torch.bfloat16).float()
    conv = F.conv2d(x.float(), wr, stride=2, padding=3)
    ref = torch.relu(conv * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)).permute(0, 2, 3, 1).reshape(-1, CO)
    xq = tc.stem_pack_input(x.cuda())
    assert torch.equal(xq[..., :3].cpu(), x.permute(0, 2, 3, 1)) and float(xq[..., 3].abs().max()) == 0
    out = tc.stem_conv_fwd(xq, tc.stem_pack_weight(w.cuda()), scale=sc.cuda(), shift=sh.cuda(), relu=True)
    assert rel(out.float().cpu(), ref) < 1e-2
    # fp32 input takes the same path
    xq32 = tc.stem_pack_input(x.float().cuda())
    assert torch.equal(xq32, xq)
    # weight gradient: dz is a column slice of a wider buffer
    Ho = P // 2
    dz = rnd((N * Ho * Ho, CO), 34).to(torch.bfloat16)
    wz = torch.zeros((CO, 3, 7, 7), requires_grad=True)
    F.conv2d(x.float(), wz, stride=2, padding=3).backward(dz.float().reshape(N, Ho, Ho, CO).permute(0, 3, 1, 2))
    dbuf = torch.zeros((N * Ho * Ho, CO + 16), dtype=torch.bfloat16, device='cuda'); dbuf[:, 8:8 + CO] = dz.cuda()
    dw = tc.stem_conv_wgrad(xq, dbuf[:, 8:8 + CO], CO)
    assert rel(dw.cpu(), wz.grad) < 2e-5
