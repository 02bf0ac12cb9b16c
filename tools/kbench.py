#!/usr/bin/env python
"""Kernel micro-benchmarks at the DenseNet-121 @128 px / 4,992-spot shapes (development tool, not the bench contract).

    python tools/kbench.py [names...]       names: gemm_xf gemm_noxf bwd1x1 gemm_bn gemm_tn conv_fwd conv_dgrad conv_wgrad stem
Prints one JSON line per case: CUDA-event time (median of reps), algorithmic TFLOP/s and GB/s.
"""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gridnext_b200 import tc

NSP = int(os.environ.get('KB_SPOTS', '4992'))
REPS = int(os.environ.get('KB_REPS', '5'))
dev = 'cuda'
bf = torch.bfloat16


def timeit(fn, flops, nbytes, name, **kw):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(REPS):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    print(json.dumps(dict(case=name, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1), gbs=round(nbytes / ms / 1e6, 0), **kw)), flush=True)


def blocks():
    # (H, c_in, layers)
    return [(32, 64, 6), (16, 128, 12), (8, 256, 24), (4, 512, 16)]


def main(names):
    sel = lambda n: not names or n in names
    for bi, (H, c0, L) in enumerate(blocks()):
        if os.environ.get('KB_BLOCKS') and str(bi + 1) not in os.environ['KB_BLOCKS']:
            continue
        M = NSP * H * H
        ct = c0 + 32 * L
        C = (torch.randn(M, ct, device=dev) * 0.5).to(bf)
        a2 = torch.relu(torch.randn(M, 128, device=dev)).to(bf)
        dC = (torch.randn(M, ct, device=dev) * 0.1).to(bf)
        dz = (torch.randn(M, 128, device=dev) * 0.1).to(bf)
        for cin in sorted({c0, c0 + 32 * (L // 2), ct - 32}):
            w1 = (torch.randn(128, cin, device=dev) * 0.05).to(bf)
            w1t = w1.t().contiguous()
            sc, sh = torch.rand(cin, device=dev) + 0.5, torch.randn(cin, device=dev) * 0.1
            s2, t2 = torch.rand(128, device=dev) + 0.5, torch.randn(128, device=dev) * 0.1
            colsum = torch.zeros(2, cin, device=dev)
            if sel('gemm_xf'):
                timeit(lambda: tc.gemm_bf16(C[:, :cin], w1, out=a2, scale=s2, shift=t2, relu=True, xf_scale=sc, xf_shift=sh),
                       2.0 * M * 128 * cin, 2.0 * M * (cin + 128), 'gemm_xf', block=bi + 1, cin=cin)
            if sel('gemm_noxf'):     # speed probe: the same GEMM without the BN+ReLU operand transform in shared memory
                timeit(lambda: tc.gemm_bf16(C[:, :cin], w1, out=a2, scale=s2, shift=t2, relu=True),
                       2.0 * M * 128 * cin, 2.0 * M * (cin + 128), 'gemm_noxf', block=bi + 1, cin=cin)
            if sel('bwd1x1'):
                dw = torch.zeros(128, cin, device=dev)
                timeit(lambda: tc.conv1x1_bwd_bf16(dz, w1t, dC[:, :cin], dict(ref=C[:, :cin], ref_is_raw=True, sc=sc, sh=sh, p0=sh, p1=sc, colsum=colsum, rmw=True), dw),
                       4.0 * M * 128 * cin, 2.0 * M * (128 + 3 * cin), 'bwd1x1', block=bi + 1, cin=cin)
            if sel('gemm_bn'):
                timeit(lambda: tc.gemm_bf16(dz, w1t, out=dC[:, :cin], bn=dict(ref=C[:, :cin], ref_is_raw=True, sc=sc, sh=sh, p0=sh, p1=sc, colsum=colsum, rmw=True)),
                       2.0 * M * 128 * cin, 2.0 * M * (128 + 3 * cin), 'gemm_bn', block=bi + 1, cin=cin)
            if sel('gemm_tn'):
                dw = torch.zeros(128, cin, device=dev)
                timeit(lambda: tc.gemm_tn_bf16(dz, C[:, :cin], dw, sc, sh), 2.0 * M * 128 * cin, 2.0 * M * (128 + cin), 'gemm_tn', block=bi + 1, cin=cin)
        w2 = torch.randn(32, 128, 3, 3, device=dev) * 0.03
        wp, wpt = tc.conv3x3_pack(w2, 0), tc.conv3x3_pack(w2, 1)
        s2 = torch.rand(128, device=dev) + 0.5
        colsum2 = torch.zeros(2, 128, device=dev)
        cin = c0
        if sel('conv_fwd'):
            timeit(lambda: tc.conv3x3_bf16(a2, NSP, H, H, 128, wp, 32, C[:, cin:cin + 32]), 2.0 * 9 * M * 128 * 32, 2.0 * M * (128 + 32), 'conv_fwd', block=bi + 1)
        if sel('conv_dgrad'):
            timeit(lambda: tc.conv3x3_bf16(dC[:, cin:cin + 32], NSP, H, H, 32, wpt, 128, dz,
                                           bn=dict(ref=a2, ref_is_raw=False, sc=s2, sh=None, p0=s2, p1=s2, colsum=colsum2)),
                   2.0 * 9 * M * 128 * 32, 2.0 * M * (32 + 128 + 128), 'conv_dgrad', block=bi + 1)
        if sel('conv_wgrad'):
            timeit(lambda: tc.conv3x3_wgrad_bf16(a2, dC[:, cin:cin + 32], NSP, H, H, 128, 32), 2.0 * 9 * M * 128 * 32, 2.0 * M * (128 + 32), 'conv_wgrad', block=bi + 1)
        del C, a2, dC, dz
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main(sys.argv[1:])
