import os, sys, torch
sys.path.insert(0, os.getcwd())
from open_clip_b200 import ops, _lib as L
M, d = 51200, 768
x4 = torch.randn(M, 4*d, device="cuda").to(torch.bfloat16)
x = torch.randn(M, d, device="cuda").to(torch.bfloat16)
w = (torch.randn(d, 4*d, device="cuda")*0.02).to(torch.bfloat16)
o = torch.empty_like(x)
for _ in range(3):
    ops.gemm(x4, w, out=o)
torch.cuda.synchronize()
