"""LayerNorm backward timing (cold L2) for the text / ViT shapes."""
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
for M, D in [(32768, 768), (50432, 768), (16448, 1024), (16384, 1024)]:
    x = torch.randn(M, D, device=dev).to(bf); g = torch.ones(D, device=dev); b = torch.zeros(D, device=dev)
    y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12)
    dy = torch.randn_like(x); dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev)
    t0 = timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db))
    t1 = timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db, addend=dy))
    by = 3.0 * M * D * 2
    print(f"[{M} x {D}] bwd {t0:6.1f} us ({by / t0 / 1e6:5.2f} TB/s)   with addend {t1:6.1f} us ({(by + M * D * 2) / t1 / 1e6:5.2f} TB/s)", flush=True)
