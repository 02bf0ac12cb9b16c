#!/usr/bin/env python
"""One launch of each g-side / gather kernel at a roofline-relevant shape, for `ncu --set full -k regex:...` captures.

    python tools/ncu_targets.py [gather hexconv corrector]
(development tool; the shapes are the C4 sweep's large-batch corner and the C2 gather)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from gridnext_b200 import hexagdly as hx, imgprocess as ip

H, W = 78, 64
dev = 'cuda'


def gather():
    from synthdata import synth
    tis, rows, cols, pr, pc = synth.synth_positions(all_in_tissue=True)
    img = torch.randint(0, 256, (16512, 16000, 3), device=dev, dtype=torch.uint8)
    cells, _ = ip.spot_table(tis, rows, cols, pr, pc, torch.device(dev))
    for P in (128, 64):
        out = torch.empty((H, W, 3, P, P), device=dev, dtype=torch.bfloat16)
        for _ in range(2):
            ip.gather_patches(img, cells, P, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225], torch.bfloat16, out=out)
    torch.cuda.synchronize()


def hexconv():
    for mode, C, B in (('0', 32, 64), ('0', 4, 256), ('1', 32, 64)):
        hx.TENSOR_CORE_MODE = mode
        conv = hx.Conv2d(C, C, 1).to(dev)
        ks = hx._kernels(conv)
        x = torch.randn(B, C, H, W, device=dev)
        dy = torch.randn(B, C, H, W, device=dev)
        wp = hx.pack_weights(ks, 1, C, C, 0)
        for _ in range(2):
            hx.hexconv_fwd(x, wp, conv.bias_tensor, C, 1)
            hx.hexconv_wgrad(x, dy, 1)
    torch.cuda.synchronize()


def hexwg():
    """The second-generation weight gradient and the BatchNorm backward at the C4 corner (256 arrays, 32 channels)."""
    from gridnext_b200._lib import call, ptr, stream
    hx.TENSOR_CORE_MODE = '1'
    B, C = 256, 32
    x = torch.randn(B, C, H, W, device=dev)
    dy = torch.randn(B, C, H, W, device=dev)
    sc = torch.rand(C, device=dev) + 0.5
    sh = torch.randn(C, device=dev) * 0.1
    mi = torch.cat([torch.zeros(C, device=dev), torch.ones(C, device=dev)])
    sums = torch.zeros(2 * C, device=dev, dtype=torch.float64)
    dH = torch.empty_like(x)
    dg, db = torch.empty(C, device=dev), torch.empty(C, device=dev)
    for _ in range(2):
        hx.hexconv_wgrad(x, dy, 1, sc, sh)
        call('gn_bn_act_bwd', ptr(dy), ptr(x), ptr(sc), ptr(sh), ptr(mi), ptr(sums), float(B * H * W), 1, ptr(dH), ptr(dg), ptr(db), B, C, H * W, 1, stream())
    torch.cuda.synchronize()


def corrector():
    from gridnext_b200.gridnet_models import GridNetHexOddr
    from gridnext_b200.losses import masked_cross_entropy
    hx.TENSOR_CORE_MODE = 'auto'
    for B in (1, 64):
        net = GridNetHexOddr(nn.Identity(), (7,), (H, W), 7).to(dev).train()
        x = torch.randn(B, 7, H, W, device=dev, requires_grad=True)
        labels = torch.randint(0, 8, (B, H, W), device=dev)
        for _ in range(2):
            out = net._correct_visium(x)
            loss, _ = masked_cross_entropy(out, labels)
            loss.backward()
    torch.cuda.synchronize()


if __name__ == '__main__':
    for w in (sys.argv[1:] or ['gather', 'hexconv', 'corrector']):
        globals()[w]()
    print('ok')
