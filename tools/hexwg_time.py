"""Timing of the tensor-core weight gradient of the hexagonal convolution: generation 1 (parity-plane rewrite + planes kernel + a separate
channel sum for the bias) vs generation 2 (csrc/hexconv_wgrad_tc2.cu, taps stacked along N or one MMA per tap), C = 32, k = 1
(development tool).  CUDA-graph replay, median of 7; accuracy of each variant against the fp32-FMA kernel printed beside the time."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gridnext_b200 import hexagdly as hx

HBM = 6544.3
try:
    HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
C = 32
sc = torch.rand(C, device='cuda') + 0.5
sh = torch.randn(C, device='cuda') * 0.1
for B in (16, 64, 256):
    x = torch.randn(B, C, 78, 64, device='cuda')
    dy = torch.randn(B, C, 78, 64, device='cuda')
    hx.TENSOR_CORE_MODE = '0'
    ref_w, ref_b = hx.hexconv_wgrad(x, dy, 1, sc, sh)
    hx.TENSOR_CORE_MODE = '1'
    for gen, stack in (('1', '1'), ('2', '1'), ('2', '0')):
        hx.TENSOR_CORE_GEN = gen
        os.environ['GRIDNEXT_B200_HEXWG2_STACK'] = stack
        for name, fn in (('wgrad', lambda: hx.hexconv_wgrad(x, dy, 1)), ('wgrad+bn_prologue', lambda: hx.hexconv_wgrad(x, dy, 1, sc, sh))):
            w, b = fn(); torch.cuda.synchronize()
            err = None
            if name != 'wgrad':
                err = [float((w - ref_w).abs().max() / ref_w.abs().max()), float((b - ref_b).abs().max() / ref_b.abs().max())]
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            g.replay(); torch.cuda.synchronize()
            ts = []
            for _ in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[3]
            by = 8.0 * C * B * 78 * 64
            print(json.dumps(dict(case=name, gen=gen, stack=stack, B=B, ms=round(ms, 4), gbs=round(by / ms / 1e6, 1), hbm_frac=round(by / ms / 1e6 / HBM, 3),
                                  err_vs_fp32=err)), flush=True)
            g.reset()
