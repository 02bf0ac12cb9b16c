# checker script: compares the CUDA path with oracle/, like the tests; not part of the product path
import sys; sys.path.insert(0,'/root/repo')
import torch, torch.nn as nn
from oracle import synth, shapes as S, gridnet_ref as R
from gridnext_b200.gridnet_models import GridNetHexOddr
def relmax(a,b):
    a,b=a.double().cpu(),b.double().cpu(); return float((a-b).abs().max()/max(float(b.abs().max()),1e-12))
B,G,H,W,n_cls=2,1000,78,64,7
f=nn.Sequential(nn.Linear(G,500),nn.Linear(500,100),nn.BatchNorm1d(100),nn.ReLU(),nn.Linear(100,100),nn.Linear(100,50),nn.BatchNorm1d(50),nn.ReLU(),nn.Linear(50,n_cls))
net=GridNetHexOddr(f,(G,),(H,W),n_cls)
sd=synth.synth_state_dict(S.gridnet_shapes(S.mlp_shapes(G,n_cls),n_cls,n_cls),11); net.load_state_dict(sd); net.cuda(); net.train(); net.patch_classifier.eval()
g=torch.Generator(); g.manual_seed(4)
x=torch.log1p(torch.poisson(torch.ones(B,G,H,W),generator=g)); dy=torch.randn(B,n_cls,H,W,generator=g)
out=net.patch_predictions(x.cuda()); (out*dy.cuda()).sum().backward()
for mode in ('fp32','bf16in'):
    sd_r={k[len('patch_classifier.'):]:(v.clone().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k,v in sd.items() if k.startswith('patch_classifier.')}
    xs=x if mode=='fp32' else x.bfloat16().float()
    sdf=sd_r if mode=='fp32' else {k:(R._rb(v,True) if k.endswith('weight') and v.dim()==2 else v) for k,v in sd_r.items()}
    ref=R.grid_from_spots(R.mlp_forward(sdf,R.spots_from_counts(xs)),B,H,W); (ref*dy).sum().backward()
    print(mode,'out',relmax(out.detach(),ref.detach()))
    for k,p in net.patch_classifier.named_parameters(): print('   ',k,relmax(p.grad,sd_r[k].grad), float(sd_r[k].grad.abs().max()), float(sd_r[k].grad.abs().mean()))
