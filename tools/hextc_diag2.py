# checker script: compares the CUDA path with oracle/, like the tests; not part of the product path
import sys; sys.path.insert(0,'/root/repo')
import torch
from gridnext_b200 import hexagdly as hx
from oracle.hexconv_ref import kernel_shapes
torch.manual_seed(0)
for (cin,cout,B,pro) in ((32,32,3,False),(32,32,3,True),(7,32,3,False),(32,7,3,False)):
    g=torch.Generator(); g.manual_seed(1)
    ks=[(torch.randn(s,generator=g)*(1.0/(cin*7)**0.5)).cuda() for s in kernel_shapes(cin,cout,1)]
    b=(torch.randn(cout,generator=g)*0.1).cuda()
    x=torch.randn(B,cin,78,64,generator=g).cuda()
    sc=(torch.rand(cin,generator=g)+0.5).cuda() if pro else None
    sh=(torch.randn(cin,generator=g)*0.3).cuda() if pro else None
    wp=hx.pack_weights(ks,1,cin,cout,0)
    outs={}
    for mode in ('0','1'):
        hx.TENSOR_CORE_MODE=mode
        st=torch.zeros(2*cout,device='cuda',dtype=torch.float64)
        outs[mode]=(hx.hexconv_fwd(x,wp,b,cout,1,sc,sh,st).double(), st.clone())
    # fp64 reference via FFMA inputs in double on GPU is not available: use mode 0 as the reference (6e-7 accurate)
    d=(outs['1'][0]-outs['0'][0])
    print(cin,cout,'pro' if pro else '', 'max rel', float(d.abs().max()/outs['0'][0].abs().max()), 'mean err', float(d.mean()), 'rms err', float(d.pow(2).mean().sqrt()),
          'stats rel', float(((outs['1'][1]-outs['0'][1]).abs()/outs['0'][1].abs().clamp_min(1e-9)).max()))
