"""Stem conv timing: direct kernels vs im2col + GEMM lowering (B = 256, 224 x 224)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, b200mm
from b200mm import ops
dev = torch.device("cuda:0"); bf16 = torch.bfloat16
N = int(os.environ.get("PB", 256))
img = torch.randn(N, 3, 224, 224, device=dev)
wp = torch.zeros(64, 152, device=dev, dtype=bf16); wp[:, :147] = (torch.randn(64, 147, device=dev) * 0.05).to(bf16)
stats = torch.zeros(128, device=dev); dw = torch.zeros(64, 152, device=dev)
dy = torch.randn(N * 112 * 112, 64, device=dev).to(bf16)
flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)
def timeit(fn, n=10):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        flush.zero_(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2] * 1e3
print("direct fwd   %.1f us" % timeit(lambda: ops.stem_conv_fwd(img, wp, col_stats=stats)))
print("direct wgrad %.1f us" % timeit(lambda: ops.stem_conv_wgrad(img, dy, dw)))
cols = ops.im2col_nchw_f32(img, 7, 2, 3, 152)[0]
print("im2col       %.1f us" % timeit(lambda: ops.im2col_nchw_f32(img, 7, 2, 3, 152)))
print("gemm fwd     %.1f us" % timeit(lambda: ops.linear_fwd(cols, wp, col_stats=stats)))
print("gemm wgrad   %.1f us" % timeit(lambda: ops.linear_wgrad(dy, cols, dw)))
