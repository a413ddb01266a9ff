"""BatchNorm forward (given column statistics) / backward timings per ResNet-50 shape, cold (L2 flushed) and warm."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops
dev = torch.device("cuda:0"); bf = torch.bfloat16
flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)
def timeit(fn, cold, iters=10):
    for _ in range(3): fn()
    ts = []
    for _ in range(iters):
        if cold: flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3
shapes = [(50176, 256), (12544, 512), (200704, 128), (50176, 1024), (12544, 2048), (802816, 64), (200704, 512)]
if len(sys.argv) > 1: shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]]
for M, C in shapes:
    x = torch.randn(M, C, device=dev).to(bf); dout = torch.randn(M, C, device=dev).to(bf)
    g = torch.rand(C, device=dev) + 0.5; b = torch.randn(C, device=dev) * 0.1
    rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev)
    xf = x.float(); stats = torch.cat([xf.sum(0), (xf * xf).sum(0)]).contiguous()
    out, mean, rstd = ops.batchnorm_fwd(x, g, b, rm, rv, relu=True, col_stats=stats)
    dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
    f = lambda: ops.batchnorm_fwd(x, g, b, rm, rv, relu=True, col_stats=stats)
    w = lambda: ops.batchnorm_bwd(dout, None, x, mean, rstd, g, dg, db, relu=True, beta=b)
    el = M * C
    for name, fn, by in (("fwd", f, 4.0 * el), ("bwd", w, 10.0 * el)):
        tc, tw = timeit(fn, True), timeit(fn, False)
        print(f"[{M:7d} x {C:4d}] {name}: cold {tc:6.1f} us ({by/tc/1e6:5.0f} GB/s)  warm {tw:6.1f} us ({by/tw/1e6:5.0f} GB/s)", flush=True)
