import sys; sys.path.insert(0,'/root/repo')
import torch
from gridnext_b200 import hexagdly as hx
hx.TENSOR_CORE_MODE='1'
C,B=32,256
conv=hx.Conv2d(C,C,1).cuda(); ks=hx._kernels(conv)
x=torch.randn(B,C,78,64,device='cuda'); wp=hx.pack_weights(ks,1,C,C,0)
st=torch.zeros(2*C,device='cuda',dtype=torch.float64)
for _ in range(3): hx.hexconv_fwd(x,wp,conv.bias_tensor,C,1,None,None,st)
torch.cuda.synchronize()
