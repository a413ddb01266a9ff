"""CTA-pair GEMM bring-up: every operand layout / epilogue the pair kernel serves, vs torch fp32, incl. an odd tile count."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from b200mm import ops
dev = torch.device("cuda:0"); bf = torch.bfloat16
torch.manual_seed(0)
def rel(a, b): return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()
bad = 0
for (M, N, K) in [(2048, 768, 768), (1157, 768, 1024), (4096, 3072, 768), (3000, 2304, 768)]:
    x = torch.randn(M, K, device=dev).to(bf); w = (torch.randn(N, K, device=dev) / K ** 0.5).to(bf)
    b = torch.randn(N, device=dev); r = torch.randn(M, N, device=dev).to(bf)
    ref = x.float() @ w.float().t() + b
    e = [rel(ops.linear_fwd(x, w, b), ref), rel(ops.linear_fwd(x, w, b, residual=r), ref + r.float())]
    z, a = ops.linear_gelu_fwd(x, w, b)
    e += [rel(z, ref), rel(a, F.gelu(ref))]
    dy = torch.randn(M, N, device=dev).to(bf)
    e.append(rel(ops.linear_dgrad(dy, w), dy.float() @ w.float()))
    zz = torch.randn(M, K, device=dev).to(bf); zf = zz.float().requires_grad_(True); F.gelu(zf).sum().backward()
    e.append(rel(ops.linear_dgrad(dy, w, gelu_z=zz), (dy.float() @ w.float()) * zf.grad))
    dw = torch.zeros(N, K, device=dev); ops.linear_wgrad(dy, x, dw)
    e.append(rel(dw, dy.float().t() @ x.float()))
    torch.cuda.synchronize()
    ok = all(v < 1e-2 for v in e)
    bad += not ok
    print(("OK  " if ok else "FAIL"), (M, N, K), [f"{v:.2e}" for v in e], flush=True)
sys.exit(1 if bad else 0)
