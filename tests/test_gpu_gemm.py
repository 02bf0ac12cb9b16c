"""GPU parity of the tcgen05 GEMM family against fp32 torch references on the same bf16-rounded inputs."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def rnd(shape, seed, scale=1.0):
    g = torch.Generator(); g.manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(torch.bfloat16)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / max(float(b.double().abs().max()), 1e-12))


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (300, 128, 64), (1000, 128, 96), (647, 7, 56), (4096, 256, 512),
                                   (2048, 500, 1000), (5000, 32, 1152), (129, 64, 992), (77, 100, 504), (20000, 128, 224)])
def test_gemm_plain(M, N, K):
    from gridnext_b200.tc import gemm_bf16
    a, b = rnd((M, K), 1), rnd((N, K), 2, 0.1)
    ref = a.float() @ b.float().t()
    out32 = gemm_bf16(a.cuda(), b.cuda(), out_dtype=torch.float32)
    assert rel(out32.cpu(), ref) < 1e-5          # fp32 accumulation of exact bf16 products
    out16 = gemm_bf16(a.cuda(), b.cuda())
    assert rel(out16.float().cpu(), ref) < 1e-2
    assert out16.dtype == torch.bfloat16


def test_gemm_epilogue_affine_relu_accumulate_and_strided_views():
    from gridnext_b200.tc import gemm_bf16
    M, N, K, LD = 700, 128, 160, 256
    buf = rnd((M, LD), 3)                         # concat-style buffer: A is its first K columns
    b = rnd((N, K), 4, 0.1)
    g = torch.Generator(); g.manual_seed(5)
    sc, sh = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g)
    ref = torch.relu((buf[:, :K].float() @ b.float().t()) * sc + sh)
    outbuf = torch.zeros((M, 384), dtype=torch.bfloat16, device='cuda')
    gemm_bf16(buf.cuda()[:, :K], b.cuda(), out=outbuf[:, 192:192 + N], scale=sc.cuda(), shift=sh.cuda(), relu=True)
    assert rel(outbuf[:, 192:192 + N].float().cpu(), ref) < 1e-2
    assert float(outbuf[:, :192].abs().max()) == 0 and float(outbuf[:, 192 + N:].abs().max()) == 0
    acc = torch.ones((M, N), dtype=torch.float32, device='cuda')
    gemm_bf16(buf.cuda()[:, :K], b.cuda(), out=acc, accumulate=True)
    assert rel(acc.cpu(), 1.0 + buf[:, :K].float() @ b.float().t()) < 1e-5


@pytest.mark.parametrize('M,N,K', [(500, 128, 64), (3000, 128, 224), (1024, 256, 992), (333, 128, 96)])
def test_gemm_operand_transform_bn_relu(M, N, K):
    from gridnext_b200.tc import gemm_bf16
    a, b = rnd((M, K), 6), rnd((N, K), 7, 0.1)
    g = torch.Generator(); g.manual_seed(8)
    xs, xt = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g) * 0.5
    act = torch.relu(a.float() * xs + xt).to(torch.bfloat16).float()      # what the kernel feeds the tensor core
    ref = act @ b.float().t()
    out = gemm_bf16(a.cuda(), b.cuda(), out_dtype=torch.float32, xf_scale=xs.cuda(), xf_shift=xt.cuda())
    assert rel(out.cpu(), ref) < 2e-3            # bf16 re-rounding of the activated operand may differ by 1 ulp from torch's


@pytest.mark.parametrize('Kp,Mo,No,xf', [(64, 128, 256, False), (1000, 128, 64, False), (20000, 128, 224, True), (5000, 128, 992, True),
                                         (9984, 500, 5000, False), (3000, 256, 512, True), (777, 7, 56, False), (4096, 32, 1152, False)])
def test_gemm_tn_weight_gradient(Kp, Mo, No, xf):
    from gridnext_b200.tc import gemm_tn_bf16
    lda, ldb = (Mo + 7) // 8 * 8 + 8, (No + 7) // 8 * 8 + 24                    # operands are column slices of wider buffers
    a, b = rnd((Kp, lda), 11, 0.5), rnd((Kp, ldb), 12)
    g = torch.Generator(); g.manual_seed(13)
    xs, xt = torch.rand(No, generator=g) + 0.5, torch.randn(No, generator=g) * 0.5
    bb = b[:, :No].float()
    if xf:
        bb = torch.relu(bb * xs + xt).to(torch.bfloat16).float()
    ref = (a[:, :Mo].double().t() @ bb.double()).float() + 1.0
    out = torch.ones((Mo, No), dtype=torch.float32, device='cuda')
    gemm_tn_bf16(a.cuda()[:, :Mo], b.cuda()[:, :No], out, xs.cuda() if xf else None, xt.cuda() if xf else None)
    assert rel(out.cpu(), ref) < (3e-3 if xf else 2e-5)


@pytest.mark.parametrize('M,N,K', [(1000, 224, 128), (4100, 992, 128), (300, 64, 128), (2048, 512, 256)])
def test_gemm_bn_relu_backward_epilogue(M, N, K):
    """Data gradient of a pre-activation 1x1 conv: dC[:, :N] += (dZ @ W) * [bn(C) > 0] * s, plus BN gradient column sums."""
    from gridnext_b200.tc import gemm_bf16
    dz, wt = rnd((M, K), 21, 0.5), rnd((N, K), 22, 0.1)          # wt[n, k] = W[k, n]
    craw = rnd((M, N + 32), 23)                                   # raw concat buffer (wider than N)
    g = torch.Generator(); g.manual_seed(24)
    mean, invstd = torch.randn(N, generator=g) * 0.2, torch.rand(N, generator=g) + 0.5
    gamma, beta = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g) * 0.3
    sc = gamma * invstd
    sh = beta - mean * sc
    dc0 = rnd((M, N + 32), 25)
    acc = dz.float() @ wt.float().t()
    a = craw[:, :N].float() * sc + sh
    gg = acc * (a > 0)
    ref_dc = dc0[:, :N].float() + gg * sc
    ref_g = gg.sum(0)
    ref_gx = (gg * (craw[:, :N].float() - mean) * invstd).sum(0)
    dc = dc0.clone().cuda()
    colsum = torch.zeros((2, N), dtype=torch.float32, device='cuda')
    gemm_bf16(dz.cuda(), wt.cuda(), out=dc[:, :N],
              bn=dict(ref=craw.cuda()[:, :N], ref_is_raw=True, sc=sc.cuda(), sh=sh.cuda(), p0=mean.cuda(), p1=invstd.cuda(), colsum=colsum, rmw=True))
    assert rel(dc[:, :N].float().cpu(), ref_dc) < 1.5e-2
    assert torch.equal(dc[:, N:].cpu(), dc0[:, N:])
    assert rel(colsum[0].cpu(), ref_g) < 2e-3 and rel(colsum[1].cpu(), ref_gx) < 2e-3


@pytest.mark.parametrize('M,N,rmw', [(1000, 224, True), (4100, 992, True), (300, 64, True), (128, 96, False), (40000, 128, True),
                                     (20000, 160, True), (9000, 320, True), (77, 480, True), (66000, 256, True)])
def test_conv1x1_backward_fused_dgrad_bn_wgrad(M, N, rmw):
    """gn_conv1x1_bwd_bf16: the data gradient with the BatchNorm/ReLU backward epilogue AND the weight gradient of a dense layer's
    bottleneck 1x1 convolution (densenet.py:12-18,26-27) in one kernel, against fp32 torch on the same bf16 inputs and against
    the two separate kernels it replaces."""
    from gridnext_b200.tc import conv1x1_bwd_bf16, conv1x1_bwd_fusable, gemm_bf16, gemm_tn_bf16
    K = 128
    dz, wt = rnd((M, K), 31, 0.5), rnd((N, K), 32, 0.1)
    craw = rnd((M, N + 32), 33)
    g = torch.Generator(); g.manual_seed(34)
    mean, invstd = torch.randn(N, generator=g) * 0.2, torch.rand(N, generator=g) + 0.5
    gamma, beta = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g) * 0.3
    sc = gamma * invstd
    sh = beta - mean * sc
    dc0 = rnd((M, N + 32), 35)
    acc = dz.float() @ wt.float().t()
    a = craw[:, :N].float() * sc + sh
    gg = acc * (a > 0)
    ref_dc = (dc0[:, :N].float() if rmw else 0) + gg * sc
    ref_g = gg.sum(0)
    ref_gx = (gg * (craw[:, :N].float() - mean) * invstd).sum(0)
    ref_dw = (dz.double().t() @ torch.relu(a).to(torch.bfloat16).double()).float() + 1.0
    dzc, wtc, crawc = dz.cuda(), wt.cuda(), craw.cuda()
    consts = dict(sc=sc.cuda(), sh=sh.cuda(), p0=mean.cuda(), p1=invstd.cuda())

    dc = dc0.clone().cuda()
    colsum = torch.zeros((2, N), dtype=torch.float32, device='cuda')
    dw = torch.ones((K, N), dtype=torch.float32, device='cuda')
    assert conv1x1_bwd_fusable(dzc, wtc, dc[:, :N], crawc[:, :N])
    conv1x1_bwd_bf16(dzc, wtc, dc[:, :N], dict(ref=crawc[:, :N], ref_is_raw=True, colsum=colsum, rmw=rmw, **consts), dw)
    torch.cuda.synchronize()
    assert rel(dc[:, :N].float().cpu(), ref_dc) < 1.5e-2
    assert torch.equal(dc[:, N:].cpu(), dc0[:, N:])
    assert rel(colsum[0].cpu(), ref_g) < 2e-3 and rel(colsum[1].cpu(), ref_gx) < 2e-3
    assert rel(dw.cpu(), ref_dw) < 3e-3

    # the two kernels it replaces, on the same inputs
    dc2 = dc0.clone().cuda()
    colsum2 = torch.zeros((2, N), dtype=torch.float32, device='cuda')
    dw2 = torch.ones((K, N), dtype=torch.float32, device='cuda')
    gemm_tn_bf16(dzc, crawc[:, :N], dw2, consts['sc'], consts['sh'])
    gemm_bf16(dzc, wtc, out=dc2[:, :N], bn=dict(ref=crawc[:, :N], ref_is_raw=True, colsum=colsum2, rmw=rmw, **consts))
    assert rel(dc[:, :N].float().cpu(), dc2[:, :N].float().cpu()) < 1e-2          # bf16 re-rounding of the += in L2
    assert rel(dw.cpu(), dw2.cpu()) < 1e-4
    assert rel(colsum.cpu(), colsum2.cpu()) < 1e-4


@pytest.mark.parametrize('M,N,K', [(40037, 128, 64), (50000, 128, 224), (45000, 96, 480), (38912, 128, 96), (64000, 128, 1000), (300, 128, 64),
                                   (5000, 100, 160), (129, 128, 992), (200000, 128, 352)])
def test_gemm_forward_conv1_operand_transform_in_tensor_memory(M, N, K):
    """The forward conv1 shape of a dense layer (bf16 output through the TMA epilogue, 64 < N <= 128; weights resident in shared memory
    when they fit and there are enough row tiles, streamed otherwise): BatchNorm+ReLU on the A operand is applied on the way into TENSOR MEMORY (tcgen05.st) and the MMA reads A from there
    (gemm_tc.cu, XT).  Checked with the BN2+ReLU epilogue of densenet.py:12-18 against fp32 torch on the same bf16 inputs."""
    from gridnext_b200.tc import gemm_bf16
    ld = (K + 7) // 8 * 8 + 32                                    # the operand is a column slice of a wider concat buffer
    a, b = rnd((M, ld), 41), rnd((N, K), 42, 0.1)
    g = torch.Generator(); g.manual_seed(43)
    xs, xt = torch.rand(K, generator=g) + 0.5, torch.randn(K, generator=g) * 0.5
    s2, t2 = torch.rand(N, generator=g) + 0.5, torch.randn(N, generator=g) * 0.2
    act = torch.relu(a[:, :K].float() * xs + xt).to(torch.bfloat16).float()
    ref = torch.relu((act @ b.float().t()) * s2 + t2)
    out = torch.full((M, N + 8), 7.0, dtype=torch.bfloat16, device='cuda')
    gemm_bf16(a.cuda()[:, :K], b.cuda(), out=out[:, :N], scale=s2.cuda(), shift=t2.cuda(), relu=True, xf_scale=xs.cuda(), xf_shift=xt.cuda())
    torch.cuda.synchronize()
    assert rel(out[:, :N].float().cpu(), ref) < 1e-2              # bf16 output rounding
    assert bool((out[:, N:] == 7.0).all())                       # nothing written beside the view
