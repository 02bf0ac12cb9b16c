#!/usr/bin/env python
"""Roofline micro-benchmarks of the g-side / gather / count-MLP kernels (BASELINE configs[0], [3]).

    python tools/kbench_g.py [hexconv corrector ce gather mlp]
One JSON line per case: CUDA-event time (median), algorithmic GB/s (bytes per SURVEY 8d) and fraction of the measured HBM
peak, or TFLOP/s where the case is compute-bound."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn
from gridnext_b200 import hexagdly as hx, imgprocess as ip
from gridnext_b200.gridnet_models import GridNetHexOddr
from gridnext_b200.losses import masked_cross_entropy

REPS = int(os.environ.get('KB_REPS', '5'))
HBM = 6544.3
try:
    HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
H, W = 78, 64
dev = 'cuda'


def timeit(fn, name, nbytes=None, flops=None, **kw):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(REPS):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    rec = dict(case=name, ms=round(ms, 4), **kw)
    if nbytes is not None:
        rec['gbs'] = round(nbytes / ms / 1e6, 0); rec['hbm_frac'] = round(nbytes / ms / 1e6 / HBM, 3)
    if flops is not None:
        rec['tflops'] = round(flops / ms / 1e9, 2)
    print(json.dumps(rec), flush=True)


def bench_hexconv():
    for k in (1, 2, 3):
        T = 1 + 3 * k * (k + 1)
        for C in (4, 8, 16, 32, 64):
            for B in (1, 16, 256):
                if B * C * H * W * 4 * 3 > 20e9:
                    continue
                conv = hx.Conv2d(C, C, k).to(dev)
                ks = hx._kernels(conv)
                x = torch.randn(B, C, H, W, device=dev)
                dy = torch.randn(B, C, H, W, device=dev)
                wp = hx.pack_weights(ks, k, C, C, 0)
                wpt = hx.pack_weights(ks, k, C, C, 1)
                cells = B * H * W
                timeit(lambda: hx.hexconv_fwd(x, wp, conv.bias_tensor, C, k), 'hexconv_fwd', nbytes=4.0 * 2 * C * cells, flops=2.0 * T * C * C * cells, k=k, C=C, B=B)
                timeit(lambda: (hx.hexconv_fwd(dy, wpt, None, C, k), hx.hexconv_wgrad(x, dy, k)), 'hexconv_bwd', nbytes=4.0 * 3 * C * cells + 4.0 * 2 * C * cells,
                       flops=4.0 * T * C * C * cells, k=k, C=C, B=B)


def bench_corrector():
    for n_cls in (7, 32):
        for B in (1, 12, 64, 256):
            net = GridNetHexOddr(nn.Identity(), (n_cls,), (H, W), n_cls).to(dev).train()
            x = torch.randn(B, n_cls, H, W, device=dev, requires_grad=True)
            labels = torch.randint(0, n_cls + 1, (B, H, W), device=dev)

            def step():
                out = net._correct_visium(x)
                loss, _ = masked_cross_entropy(out, labels)
                loss.backward()
            cells = B * H * W
            fwd = 4.0 * ((n_cls + 32) + 3 * 64 + (32 + n_cls)); bwd = 4.0 * ((32 + 2 * n_cls) + 3 * 96 + (n_cls + 64))
            timeit(step, 'corrector_fwd_bwd_ce', nbytes=(fwd + bwd + 16.0 * n_cls + 8) * cells, n_cls=n_cls, B=B, spots_per_s=None)


def bench_ce():
    for B in (12, 256):
        logits = torch.randn(B, 7, H, W, device=dev, requires_grad=True)
        labels = torch.randint(0, 8, (B, H, W), device=dev)

        def step():
            loss, _ = masked_cross_entropy(logits, labels)
            loss.backward()
        timeit(step, 'masked_ce_fwd_bwd', nbytes=(8.0 * 7 + 8) * B * H * W, B=B)


def bench_gather():
    from synthdata import synth
    tis, rows, cols, pr, pc = synth.synth_positions(all_in_tissue=True)
    img = torch.randint(0, 256, (16512, 16000, 3), device=dev, dtype=torch.uint8)
    cells, _ = ip.spot_table(tis, rows, cols, pr, pc, torch.device(dev))
    for P in (128, 64, 256):
        for dt, eb in ((torch.bfloat16, 2), (torch.float32, 4)):
            out = torch.empty((H, W, 3, P, P), device=dev, dtype=dt)
            timeit(lambda: ip.gather_patches(img, cells, P, [0.485, 0.456, 0.406], [0.229, 0.224, 0.225], dt, out=out), 'patch_gather',
                   nbytes=float(H * W) * 3 * P * P * (1 + eb), P=P, out=str(dt).split('.')[-1])


def bench_mlp():
    from gridnext_b200.training import gridwise_step
    for B, G in ((1, 5000), (4, 5000), (12, 2000)):
        f = nn.Sequential(nn.Linear(G, 500), nn.Linear(500, 100), nn.BatchNorm1d(100), nn.ReLU(), nn.Linear(100, 100), nn.Linear(100, 50),
                          nn.BatchNorm1d(50), nn.ReLU(), nn.Linear(50, 7))
        net = GridNetHexOddr(f, (G,), (H, W), 7).to(dev)
        net.train(); net.patch_classifier.eval()
        x = torch.log1p(torch.poisson(torch.ones(B, G, H, W, device=dev)))
        y = torch.randint(0, 8, (B, H, W), device=dev)
        crit = nn.CrossEntropyLoss()

        def step():
            gridwise_step(net, x, y, crit, 1, True)
            for p in net.parameters():
                p.grad = None
        macs = G * 500 + 500 * 100 + 100 * 100 + 100 * 50 + 50 * 7
        timeit(step, 'count_gridnet_fwd_bwd (C1)', flops=2.0 * (3 * macs - G * 500) * B * H * W, nbytes=4.0 * G * 2 * B * H * W, B=B, G=G)


if __name__ == '__main__':
    which = sys.argv[1:] or ['hexconv', 'corrector', 'ce', 'gather', 'mlp']
    for w in which:
        globals()['bench_' + w]()
