import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F, b200mm
from b200mm import ops
dev = torch.device("cuda:0"); bf16 = torch.bfloat16
def rel(a, b): return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()
for (N, H, W) in [(1, 8, 8), (2, 3, 6), (1, 3, 8), (1, 4, 6), (2, 8, 6), (2, 3, 8), (1, 3, 6), (1, 3, 7), (1, 5, 6), (4, 3, 6)]:
    torch.manual_seed(1)
    C = 64
    x = torch.randn(N, C, H, W, device=dev).to(bf16).float()
    w = (torch.randn(C, C, 3, 3, device=dev) * 0.05).to(bf16).float()
    ref = F.conv2d(x, w, None, 1, 1)
    w_ohwi = w.permute(0, 2, 3, 1).reshape(C, 9 * C).to(bf16).contiguous()
    xn = x.permute(0, 2, 3, 1).reshape(-1, C).to(bf16).contiguous()
    y, _, _ = ops.conv_fwd(xn, N, H, W, C, w_ohwi, 3, 1, 1)
    got = y.float().view(N, H, W, C).permute(0, 3, 1, 2)
    print((N, H, W), "rel err", round(rel(got, ref), 4), flush=True)
print("with stats, test order")
for (N, H, W) in [(2, 9, 61), (1, 3, 6), (1, 3, 6)]:
    torch.manual_seed(31)
    C = 64
    x = torch.randn(N, C, H, W, device=dev).to(bf16).float()
    w = (torch.randn(C, C, 3, 3, device=dev) * 0.05).to(bf16).float()
    ref = F.conv2d(x, w, None, 1, 1)
    w_ohwi = w.permute(0, 2, 3, 1).reshape(C, 9 * C).to(bf16).contiguous()
    xn = x.permute(0, 2, 3, 1).reshape(-1, C).to(bf16).contiguous()
    stats = torch.zeros(2 * C, device=dev)
    y, _, _ = ops.conv_fwd(xn, N, H, W, C, w_ohwi, 3, 1, 1, col_stats=stats)
    got = y.float().view(N, H, W, C).permute(0, 3, 1, 2)
    print((N, H, W), "rel err", round(rel(got, ref), 4), "stats", round(rel(stats[:C], y.float().sum(0)), 5), flush=True)
