"""Timing of the tensor-core hexagonal convolution, generation 1 (parity planes) vs generation 2 (in-kernel conversion), at the C4
corner C = 32, k = 1 for a few batch sizes (development tool).  CUDA-graph replay, median of 7."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gridnext_b200 import hexagdly as hx

hx.TENSOR_CORE_MODE = '1'
HBM = 6544.3
try:
    HBM = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'MEASURED_PEAKS.json')))['hbm_gbs']
except Exception:
    pass
C = 32
conv = hx.Conv2d(C, C, 1).cuda()
ks = hx._kernels(conv)
wp = hx.pack_weights(ks, 1, C, C, 0)
sc = torch.rand(C, device='cuda') + 0.5
sh = torch.randn(C, device='cuda') * 0.1
for B in (16, 64, 256):
    x = torch.randn(B, C, 78, 64, device='cuda')
    st = torch.zeros(2 * C, device='cuda', dtype=torch.float64)
    for gen in ('1', '2', '2cp'):
        hx.TENSOR_CORE_GEN = gen[0]
        os.environ['GRIDNEXT_B200_H2_CPASYNC'] = '1' if gen == '2cp' else '0'      # gen 2 with 16-byte asynchronous copies instead of TMA boxes on the input
        for name, fn in (('fwd', lambda: hx.hexconv_fwd(x, wp, conv.bias_tensor, C, 1)),
                         ('fwd+bn_prologue+stats', lambda: hx.hexconv_fwd(x, wp, conv.bias_tensor, C, 1, sc, sh, st))):
            fn(); torch.cuda.synchronize()
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
            print(json.dumps(dict(case=name, gen=gen, B=B, ms=round(ms, 4), gbs=round(by / ms / 1e6, 1), hbm_frac=round(by / ms / 1e6 / HBM, 3))), flush=True)
            g.reset()
