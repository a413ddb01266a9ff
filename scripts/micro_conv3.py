"""layer1 3x3 conv timing (B = 256, 56 x 56, 64 -> 64): halo-resident kernels vs TMA-im2col implicit GEMM
(run twice: B200MM_HALO_CONV=1 / 0)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, b200mm
from b200mm import ops
dev = torch.device("cuda:0"); bf16 = torch.bfloat16
N, H, W, C = int(os.environ.get("PB", 256)), 56, 56, 64
x = torch.randn(N * H * W, C, device=dev).to(bf16); dy = torch.randn(N * H * W, C, device=dev).to(bf16)
w = (torch.randn(C, 9 * C, device=dev) * 0.05).to(bf16); dw = torch.zeros(C, 9 * C, device=dev)
stats = torch.zeros(128, device=dev)
flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)
def timeit(fn, n=10):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        flush.zero_(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2] * 1e3
print("HALO_CONV =", os.environ.get("B200MM_HALO_CONV", "1"))
print("conv fwd (+stats) %.1f us" % timeit(lambda: ops.conv_fwd(x, N, H, W, C, w, 3, 1, 1, col_stats=stats)))
print("conv fwd          %.1f us" % timeit(lambda: ops.conv_fwd(x, N, H, W, C, w, 3, 1, 1)))
print("conv wgrad        %.1f us" % timeit(lambda: ops.conv_wgrad(dy, x, N, H, W, C, 3, 1, 1, dw)))
