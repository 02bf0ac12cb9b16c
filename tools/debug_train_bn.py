"""[checker script: compares the CUDA path with oracle/, like the tests; not part of the product path] Per-parameter gradient error of the train-mode DenseNet path against the bf16-emulating oracle, in execution order."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import synth, shapes as S
from oracle import gridnet_ref as R
from gridnext_b200.densenet import DenseNet

cfgs = {'a': (dict(growth_rate=16, block_config=(3, 3), num_init_features=32, bn_size=4), 32, 8),
        'b': (dict(growth_rate=8, block_config=(4, 2), num_init_features=16, bn_size=2), 32, 8),
        'c': (dict(growth_rate=8, block_config=(2, 2), num_init_features=16, bn_size=2), 32, 6)}
kw, P, N = cfgs[sys.argv[1] if len(sys.argv) > 1 else 'a']
N = int(sys.argv[2]) if len(sys.argv) > 2 else N
net = DenseNet(num_classes=7, small_inputs=False, efficient=False, drop_rate=0, **kw)
sd = synth.synth_state_dict(S.densenet_shapes(kw['growth_rate'], tuple(kw['block_config']), kw['num_init_features'], kw['bn_size']), 23)
net.load_state_dict(sd)
net.cuda().train()
g = torch.Generator(); g.manual_seed(29)
x = torch.randn(N, 3, P, P, generator=g)
dy = torch.randn(N, 7, generator=g)
sd_r = {k: (v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k, v in sd.items()}
stats = {}
ref = R.densenet_forward(sd_r, x, emulate_bf16=True, training=True, stats_out=stats)
(ref * dy).sum().backward()
out = net(x.cuda())
(out * dy.cuda()).sum().backward()
got = net.state_dict()
for k, v in stats.items():
    old = sd[k]
    bg, br = (got[k].cpu() - 0.9 * old) / 0.1, (v - 0.9 * old) / 0.1
    print('%-60s batch-stat err %.2e (max %.3g)' % (k, float((bg - br).abs().max() / br.abs().max()), float(br.abs().max())))
print('logits', float((out.detach().cpu() - ref.detach()).abs().max() / ref.abs().max()))
for k, p in net.named_parameters():
    r = sd_r[k].grad
    e = float((p.grad.cpu() - r).abs().max() / r.abs().max())
    print('%-55s %.4f  |ref| %.3g' % (k, e, float(r.abs().max())))
