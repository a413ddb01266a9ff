import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops
dev = torch.device("cuda:0"); bf = torch.bfloat16
M, D = 50432, 768
x = torch.randn(M, D, device=dev).to(bf); g = torch.ones(D, device=dev); b = torch.zeros(D, device=dev)
y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12)
dy = torch.randn_like(x); dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev)
for _ in range(3):
    ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db, addend=dy)
torch.cuda.synchronize(); print("ok")
