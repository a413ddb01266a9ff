"""Column-sum (bias gradient) timing, cold L2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops
dev = torch.device("cuda:0"); bf = torch.bfloat16
flush = torch.empty(512 << 20, device=dev, dtype=torch.uint8)
def timeit(fn, n=10):
    for _ in range(2): fn()
    ts = []
    for _ in range(n):
        flush.zero_(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort(); return ts[len(ts) // 2] * 1e3
for M, N in [(32768, 768), (32768, 2304), (32768, 3072), (50432, 768), (50432, 3072)]:
    x = torch.randn(M, N, device=dev).to(bf); out = torch.zeros(N, device=dev)
    t = timeit(lambda: ops.colsum(x, out))
    print(f"[{M} x {N}] {t:6.1f} us  {M * N * 2 / t / 1e6:5.2f} TB/s", flush=True)
