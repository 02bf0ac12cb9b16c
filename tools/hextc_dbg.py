"""Where the second-generation hex forward spends its time: the kernel with parts switched off (development tool; GRIDNEXT_B200_H2_DBG:
1 no output stores, 2 no MMAs, 4 no conversion, 8 no TMEM reads).  Results are wrong by construction.  CUDA-graph replay, median of 7."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gridnext_b200 import hexagdly as hx
hx.TENSOR_CORE_MODE = '1'
C, B = 32, 256
conv = hx.Conv2d(C, C, 1).cuda()
wp = hx.pack_weights(hx._kernels(conv), 1, C, C, 0)
x = torch.randn(B, C, 78, 64, device='cuda')
sc = torch.rand(C, device='cuda') + 0.5
sh = torch.randn(C, device='cuda') * 0.1
st = torch.zeros(2 * C, device='cuda', dtype=torch.float64)
for dbg in (0, 1, 2, 4, 8, 9, 6, 13, 15):
    os.environ['GRIDNEXT_B200_H2_DBG'] = str(dbg)
    for name, fn in (('fwd', lambda: hx.hexconv_fwd(x, wp, conv.bias_tensor, C, 1)), ('fwd+pro+stats', lambda: hx.hexconv_fwd(x, wp, conv.bias_tensor, C, 1, sc, sh, st))):
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
        print(json.dumps(dict(case=name, dbg=dbg, ms=round(sorted(ts)[3], 4))), flush=True)
        g.reset()
