# checker script: compares the CUDA path with oracle/, like the tests; not part of the product path
import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import torch
from oracle import synth, shapes as S, gridnet_ref as R
from gridnext_b200.densenet import DenseNet
def relmax(a,b):
    a,b=a.double().cpu(),b.double().cpu(); return float((a-b).abs().max()/max(float(b.abs().max()),1e-12))
kw=dict(growth_rate=32, block_config=(6,12,24,16), num_init_features=64, bn_size=4)
for N,P in ((3,64),(12,64)):
    net=DenseNet(num_classes=7, small_inputs=False, **kw)
    sd=synth.synth_state_dict(S.densenet_shapes(32,(6,12,24,16),64,4),91); net.load_state_dict(sd); net=net.cuda().eval()
    g=torch.Generator(); g.manual_seed(17)
    x=torch.randn(N,3,P,P,generator=g); dy=torch.randn(N,7,generator=g)
    sd_r={k:(v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k,v in sd.items()}
    ref=R.densenet_forward(sd_r,x,emulate_bf16=True); (ref*dy).sum().backward()
    out=net(x.cuda()); (out*dy.cuda()).sum().backward()
    errs=sorted(((relmax(p.grad, sd_r[k].grad),k) for k,p in net.named_parameters()), reverse=True)
    print(N,P,'logits',relmax(out,ref.detach()))
    print(errs[:8]); import statistics; print('median',statistics.median(e for e,_ in errs), 'frac<3e-2', sum(e<3e-2 for e,_ in errs)/len(errs))
