# checker script: compares the CUDA path with oracle/, like the tests; not part of the product path
import sys; sys.path.insert(0,'/root/repo')
import torch, torch.nn as nn
from oracle import gridnet_ref as R
from gridnext_b200 import hexagdly as hx
from gridnext_b200.gridnet_models import GridNetHexOddr
def rel_err(a,b):
    a,b=a.detach().double().cpu(),b.detach().double().cpu(); return float((a-b).abs().max()/max(float(b.abs().max()),1e-12))
for mode in ('0','1'):
    hx.TENSOR_CORE_MODE=mode
    torch.manual_seed(0)
    net=GridNetHexOddr(nn.Identity(),(7,),(78,64),7).cuda().train()
    sd={k:v.detach().cpu() for k,v in net.corrector.state_dict().items()}
    g=torch.Generator(); g.manual_seed(3)
    B=3
    x=torch.randn(B,7,78,64,generator=g); dy=torch.randn(B,7,78,64,generator=g)
    sd_r={k:(v.double().requires_grad_(True) if v.is_floating_point() and 'running' not in k else v) for k,v in sd.items()}
    x_r=x.double().requires_grad_(True); y_r=R.corrector_forward(sd_r,x_r,use_bn=True,training=True); y_r.backward(dy.double())
    x_g=x.cuda().requires_grad_(True); y_g=net._correct_visium(x_g); y_g.backward(dy.cuda())
    print('mode',mode,'y',rel_err(y_g,y_r),'dx',rel_err(x_g.grad,x_r.grad))
    for name,p in net.corrector.named_parameters():
        ref=sd_r[name].grad; scale=max(float(ref.abs().max()),1e-3)
        print('   ',name, float((p.grad.double().cpu()-ref).abs().max())/scale)
