# checker script: compares the CUDA path with oracle/, like the tests; not part of the product path
import sys, json, os
sys.path.insert(0, '.')
import numpy as np, torch
from oracle import synth, shapes as S, gridnet_ref as R
from gridnext_b200.densenet import DenseNet
cfgs = {'tiny': (dict(growth_rate=8, block_config=(2, 3), num_init_features=16, bn_size=2), 32, 3),
        'tiny16': (dict(growth_rate=16, block_config=(2, 2), num_init_features=16, bn_size=2), 32, 16),
        'd121': (dict(growth_rate=32, block_config=(6, 12, 24, 16), num_init_features=64, bn_size=4), 64, 4)}
which = sys.argv[1] if len(sys.argv) > 1 else 'tiny'
kw, P, N = cfgs[which]
net = DenseNet(num_classes=7, small_inputs=False, **kw)
sd = synth.synth_state_dict(S.densenet_shapes(kw['growth_rate'], kw['block_config'], kw['num_init_features'], kw['bn_size']), 32)
net.load_state_dict(sd); net = net.cuda().eval()
g = torch.Generator(); g.manual_seed(1)
x = torch.randn(N, 3, P, P, generator=g); dy = torch.randn(N, 7, generator=g)
sdr = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in sd.items()}
ref = R.densenet_forward(sdr, x); (ref * dy).sum().backward()
out = net(x.cuda()); (out * dy.cuda()).sum().backward()
print('logits rel err', float((out.cpu() - ref).abs().max() / ref.abs().max()))
for k, p in net.named_parameters():
    r = sdr[k].grad
    e = float((p.grad.cpu() - r).abs().max() / max(float(r.abs().max()), 1e-12))
    cos = float(torch.nn.functional.cosine_similarity(p.grad.cpu().flatten(), r.flatten(), dim=0))
    flag = ' <<<' if e > 0.05 else ''
    print('%-55s rel %.4f cos %.5f |ref| %.3e%s' % (k, e, cos, float(r.abs().max()), flag))
