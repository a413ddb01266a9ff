import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn.functional as F, b200mm
from b200mm import ops
dev = torch.device("cuda:0"); bf16 = torch.bfloat16
def rel(a, b): return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()
shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(1, 3, 6)]
for (N, H, W) in shapes:
    torch.manual_seed(31)
    C = 64
    x = torch.randn(N, C, H, W, device=dev).to(bf16).float()
    w = (torch.randn(C, C, 3, 3, device=dev) * 0.05).to(bf16).float()
    ref = F.conv2d(x, w, None, 1, 1)
    w_ohwi = w.permute(0, 2, 3, 1).reshape(C, 9 * C).to(bf16).contiguous()
    xn = x.permute(0, 2, 3, 1).reshape(-1, C).to(bf16)
    stats = torch.zeros(2 * C, device=dev)
    y = torch.full((N * H * W, C), 7.0, device=dev, dtype=bf16)
    ops.conv_fwd(xn, N, H, W, C, w_ohwi, 3, 1, 1, col_stats=stats, out=y)
    got = y.float().view(N, H, W, C).permute(0, 3, 1, 2)
    print((N, H, W), "rel err", round(rel(got, ref), 4), "untouched", int((y == 7.0).sum()), "of", y.numel(), flush=True)
