#!/usr/bin/env python
"""Which Python frames launch the small element-wise kernels of one eager training step (development tool)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torch.profiler import profile, ProfilerActivity

cfg = sys.argv[1] if len(sys.argv) > 1 else 'c2'
dev = torch.device('cuda', 0)
wl = bench.Workload(cfg, dev, 0, 1)
for i in range(3):
    wl.step_resident(0)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    wl.step_resident(0)
    torch.cuda.synchronize()
cnt = collections.Counter()
for ev in prof.events():
    if ev.name in ('aten::fill_', 'aten::zero_', 'aten::zeros', 'aten::copy_', 'aten::clone', 'aten::zeros_like', 'aten::add_', 'aten::mul'):
        st = [s for s in (ev.stack or []) if 'site-packages/torch' not in s and 'fill_probe' not in s][:3]
        cnt[(ev.name, ' <- '.join(st))] += 1
for (name, st), n in cnt.most_common(25):
    print(n, name, st)
k = collections.Counter(ev.name for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA)
for name, n in k.most_common(12):
    print('KERNEL', n, name[:100])
