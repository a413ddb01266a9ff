"""One launch each of the epilogue-heavy GEMMs (GELU, dGELU, residual) for `ncu --set full --import-source on`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops
dev = torch.device("cuda:0"); bf = torch.bfloat16
M, N, K = 32768, 3072, 768
A = torch.randn(M, K, device=dev).to(bf); W = torch.randn(N, K, device=dev).to(bf); bias = torch.zeros(N, device=dev)
out = torch.empty(M, N, device=dev, dtype=bf); out2 = torch.empty_like(out); aux = torch.randn(M, N, device=dev).to(bf)
for _ in range(2):
    ops.gemm_raw(A, False, W, False, M, N, K, out, epi=0, bias=bias)                       # plain
    ops.gemm_raw(A, False, W, False, M, N, K, out, epi=1, bias=bias, out2=out2)            # GELU
    ops.gemm_raw(A, False, W, False, M, N, K, out, epi=2, aux=aux)                         # dGELU
W2 = torch.randn(768, 768, device=dev).to(bf); o2 = torch.empty(M, 768, device=dev, dtype=bf); res = torch.randn(M, 768, device=dev).to(bf)
b2 = torch.zeros(768, device=dev)
for _ in range(2):
    ops.gemm_raw(A, False, W2, False, M, 768, 768, o2, epi=0, bias=b2, residual=res)
torch.cuda.synchronize(); print("ok")
