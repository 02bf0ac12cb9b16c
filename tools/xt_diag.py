#!/usr/bin/env python
"""Development: which (M, K) shapes of the forward conv1 GEMM run (each in its own process)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CHILD = r'''
import sys, torch
sys.path.insert(0, %r)
from gridnext_b200 import tc
M, K, ld, reps = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
C = (torch.randn(M, ld, device='cuda') * 0.5).to(torch.bfloat16)
w = (torch.randn(128, K, device='cuda') * 0.05).to(torch.bfloat16)
a2 = torch.empty(M, 128, device='cuda', dtype=torch.bfloat16)
sc, sh = torch.rand(K, device='cuda') + 0.5, torch.randn(K, device='cuda') * 0.1
s2, t2 = torch.rand(128, device='cuda') + 0.5, torch.randn(128, device='cuda') * 0.1
for i in range(reps):
    tc.gemm_bf16(C[:, :K], w, out=a2, scale=s2, shift=t2, relu=True, xf_scale=sc, xf_shift=sh)
    torch.cuda.synchronize()
ref = torch.relu((torch.relu(C[:4096, :K].float() * sc + sh).to(torch.bfloat16).float() @ w.float().t()) * s2 + t2)
print('ok', float((a2[:4096].float() - ref).abs().max()))
''' % ROOT
for M, K, ld, reps in [(50000, 224, 256, 3), (500000, 224, 256, 3), (5111808, 224, 256, 2), (5111808, 256, 256, 2), (5111808, 192, 256, 2),
                       (5111808, 160, 256, 2), (5111808, 224, 512, 2), (1277952, 224, 256, 4)]:
    r = subprocess.run([sys.executable, '-c', CHILD, str(M), str(K), str(ld), str(reps)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=120)
    print(M, K, ld, reps, '->', (r.stdout.strip() or r.stderr.strip().splitlines()[-1][:160]), flush=True)
