import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mmduet_b200 import ops
from mmduet_b200._lib import EPI_BF16, EPI_RESID_F32
torch.manual_seed(0)
M = 23328
x = (torch.randn(M, 1152, device="cuda") * 0.5).bfloat16()
w = (torch.randn(4304, 1152, device="cuda") * 0.03).bfloat16()
b = torch.randn(4304, device="cuda")
out = torch.empty(M, 4304, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.gemm(x, w, bias=b, out=out)           # fc1 shape, BN=256
h = (torch.randn(M, 4304, device="cuda") * 0.5).bfloat16()
w2 = (torch.randn(1152, 4304, device="cuda") * 0.02).bfloat16()
res = torch.randn(M, 1152, device="cuda")
for _ in range(3):
    ops.gemm(h, w2, bias=b[:1152].contiguous(), out=res, epi=EPI_RESID_F32)   # fc2 shape, BN=192, fp32 residual RMW
torch.cuda.synchronize()
