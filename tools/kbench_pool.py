#!/usr/bin/env python
"""Pooling / element-wise kernels of the DenseNet path at the DenseNet-121 @128 px / 4,992-spot shapes (development tool).

    python tools/kbench_pool.py [names...]      names: maxpool_fwd maxpool_bwd avgpool_fwd pool_bwd stem_fwd stem_wgrad
One JSON line per case: CUDA-event time (median), algorithmic GB/s and its fraction of the measured copy bandwidth.
"""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gridnext_b200._lib import call, ptr, stream

NSP = int(os.environ.get('KB_SPOTS', '4992'))
REPS = int(os.environ.get('KB_REPS', '5'))
PEAK = 6544.3
dev, bf = 'cuda', torch.bfloat16


def timeit(fn, nbytes, name, **kw):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(REPS):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    gbs = nbytes / ms / 1e6
    print(json.dumps(dict(case=name, ms=round(ms, 4), gbs=round(gbs), hbm_frac=round(gbs / PEAK, 3), **kw)), flush=True)


def main(names):
    sel = lambda n: not names or n in names
    n = NSP
    if sel('maxpool_fwd') or sel('maxpool_bwd'):
        H0, c0, ct = 64, 64, 256
        M0, M1 = n * H0 * H0, n * (H0 // 2) ** 2
        act0 = torch.relu(torch.randn(M0, c0, device=dev)).to(bf)
        C = torch.empty(M1, ct, device=dev, dtype=bf)
        idx0 = torch.empty(M1, c0, device=dev, dtype=torch.uint8)
        if sel('maxpool_fwd'):
            timeit(lambda: call('gn_maxpool3s2_fwd', ptr(act0), c0, n, H0, H0, c0, ptr(C), ct, ptr(idx0), stream()),
                   2.0 * M0 * c0 + 3.0 * M1 * c0, 'maxpool_fwd')
        if sel('maxpool_bwd'):
            call('gn_maxpool3s2_fwd', ptr(act0), c0, n, H0, H0, c0, ptr(C), ct, ptr(idx0), stream())
            dC = (torch.randn(M1, ct, device=dev) * 0.1).to(bf)
            dz0 = torch.empty(M0, c0, device=dev, dtype=bf)
            sc, p0, p1 = torch.rand(c0, device=dev) + 0.5, torch.randn(c0, device=dev) * 0.1, torch.rand(c0, device=dev) + 0.5
            colsum = torch.zeros(2, 4096, device=dev)
            timeit(lambda: call('gn_maxpool3s2_bnrelu_bwd', ptr(dC), ct, ptr(idx0), ptr(act0), c0, n, H0, H0, c0, ptr(sc), ptr(p0), ptr(p1),
                                ptr(dz0), c0, ptr(colsum), 4096, stream()),
                   3.0 * M1 * c0 + 4.0 * M0 * c0, 'maxpool_bwd')
        del act0, C, idx0
        torch.cuda.empty_cache()
    if sel('stem_fwd') or sel('stem_wgrad'):
        from gridnext_b200 import tc
        P, CO = 128, 64
        x = torch.randn(n, 3, P, P, device=dev).to(bf)
        xq = tc.stem_pack_input(x)
        w = torch.randn(CO, 3, 7, 7, device=dev) * 0.1
        wq = tc.stem_pack_weight(w)
        sc, sh = torch.rand(CO, device=dev) + 0.5, torch.randn(CO, device=dev) * 0.1
        M0 = n * (P // 2) ** 2
        out = torch.empty(M0, CO, device=dev, dtype=bf)
        if sel('stem_fwd'):
            timeit(lambda: tc.stem_conv_fwd(xq, wq, scale=sc, shift=sh, relu=True, out=out), 8.0 * n * P * P + 2.0 * M0 * CO, 'stem_fwd')
        if sel('stem_wgrad'):
            dz = (torch.randn(M0, CO, device=dev) * 0.1).to(bf)
            dwq = torch.zeros(CO, 224, device=dev)
            timeit(lambda: tc.stem_conv_wgrad_into(xq, dz, CO, dwq), 8.0 * n * P * P + 2.0 * M0 * CO, 'stem_wgrad')
        del x, xq, out
        torch.cuda.empty_cache()
    for H, ct in ((32, 256), (16, 512), (8, 1024)):
        M = n * H * H
        C = (torch.randn(M, ct, device=dev) * 0.5).to(bf)
        sc, sh = torch.rand(ct, device=dev) + 0.5, torch.randn(ct, device=dev) * 0.1
        if sel('avgpool_fwd'):
            pooled = torch.empty(M // 4, ct, device=dev, dtype=bf)
            timeit(lambda: call('gn_bnrelu_avgpool2_fwd', ptr(C), ct, n, H, H, ct, ptr(sc), ptr(sh), ptr(pooled), ct, stream()),
                   2.0 * M * ct * 1.25, 'avgpool_fwd', H=H, C=ct)
        if sel('pool_bwd'):
            dP = (torch.randn(M // 4, ct, device=dev) * 0.1).to(bf)
            dC = torch.empty(M, ct, device=dev, dtype=bf)
            colsum = torch.zeros(2, 4096, device=dev)
            timeit(lambda: call('gn_pool_bnrelu_bwd', ptr(dP), ct, 0, ptr(C), ct, n, H, H, ct, ptr(sc), ptr(sh), ptr(sh), ptr(sc), ptr(dC), ct,
                                ptr(colsum), 4096, stream()),
                   2.0 * M * ct * 2.25, 'pool_bwd', H=H, C=ct)
        del C
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main(sys.argv[1:])
