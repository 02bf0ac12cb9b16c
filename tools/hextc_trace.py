"""Pipeline timeline of the second-generation tensor-core hex convolution (development tool): CTA 0 records clock64() at
producer issue / converter stage-ready, ring-ready, done / MMA ready, issued / epilogue accumulator-ready, read, done."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gridnext_b200 import hexagdly as hx, _lib

hx.TENSOR_CORE_MODE = '1'
hx.TENSOR_CORE_GEN = '2'
C, B = 32, int(os.environ.get('B', '256'))
conv = hx.Conv2d(C, C, 1).cuda()
wp = hx.pack_weights(hx._kernels(conv), 1, C, C, 0)
x = torch.randn(B, C, 78, 64, device='cuda')
for _ in range(3):
    hx.hexconv_fwd(x, wp, conv.bias_tensor, C, 1)
tr = torch.zeros(13 * 64, device='cuda', dtype=torch.int64)
_lib.call('gn_hexconv_tc2_set_trace', _lib.ptr(tr))
hx.hexconv_fwd(x, wp, conv.bias_tensor, C, 1)
torch.cuda.synchronize()
_lib.call('gn_hexconv_tc2_set_trace', None)
t = tr.cpu().view(13, 64)
t0 = int(t[t > 0].min())
names = ['prod_issue', 'cv_stage', 'cv_ring', 'cv_done', 'mma_ready', 'mma_issued', 'ep_full', 'ep_read', 'ep_done', 'ep_prebar', 'ep_postbar', 'w13_full', 'w13_read']
print('cycles relative to the first event; columns = pair / tile index')
for i, n in enumerate(names):
    print('%-10s' % n, ' '.join('%7d' % (int(v) - t0 if v > 0 else -1) for v in t[i, 4:24]))
