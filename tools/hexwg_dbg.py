"""Where the second-generation hex weight gradient spends its time: the kernel with parts switched off (development tool;
GRIDNEXT_B200_HEXWG2_DBG: 1 no MMAs, 2 no conversion / operand stores, 4 no global copies).  Results are wrong by construction."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gridnext_b200 import hexagdly as hx
hx.TENSOR_CORE_MODE = '1'
C, B = 32, 256
x = torch.randn(B, C, 78, 64, device='cuda'); dy = torch.randn(B, C, 78, 64, device='cuda')
for stack in ('1', '0'):
    for dbg in (0, 1, 2, 4, 3, 5, 6, 7):
        os.environ['GRIDNEXT_B200_HEXWG2_STACK'] = stack
        os.environ['GRIDNEXT_B200_HEXWG2_DBG'] = str(dbg)
        for want_bias in (True, False):
            fn = lambda: hx.hexconv_wgrad(x, dy, 1, want_bias=want_bias)
            fn(); torch.cuda.synchronize()
            ts = []
            for _ in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            print(json.dumps(dict(stack=stack, dbg=dbg, bias=want_bias, ms=round(sorted(ts)[3], 4))), flush=True)
