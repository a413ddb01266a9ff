"""Fixed cost of a GEMM launch: back-to-back launches (no events, no flush between them) of one-tile-per-CTA problems
and of the short convolution shapes, next to a trivial elementwise kernel."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops, _lib
dev = torch.device("cuda:0"); bf = torch.bfloat16
lib = _lib.load()


def b2b(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def graph_b2b(fn, n=50):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n): fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


small = torch.zeros(1024, device=dev)
print(f"elementwise add on 1024 floats: eager {b2b(lambda: small.add_(1.0)):.2f} us / launch, graph {graph_b2b(lambda: small.add_(1.0)):.2f}")
for (M, N, K, mode) in [(128 * 148, 256, 64, "plain"), (128 * 148, 256, 256, "plain"), (128 * 148, 256, 1024, "plain"),
                        (128 * 148 * 2, 256, 256, "plain"), (128 * 148 * 4, 256, 256, "plain"),
                        (50176, 1024, 256, "stats"), (50176, 1024, 256, "plain"), (50176, 256, 1024, "plain"),
                        (12544, 2048, 512, "plain"), (12544, 512, 2048, "plain"), (200704, 512, 128, "plain"),
                        (802816, 256, 64, "plain"), (32768, 768, 768, "plain"), (32768, 3072, 768, "plain")]:
    A = torch.randn(M, K, device=dev).to(bf)
    Bm = (torch.randn(N, K, device=dev) / K ** 0.5).to(bf)
    out = torch.zeros(M, N, device=dev, dtype=bf)
    st = torch.zeros(2 * N, device=dev) if mode == "stats" else None
    fn = lambda: ops.gemm_raw(A, False, Bm, False, M, N, K, out, epi=0, col_stats=st)
    te, tg = b2b(fn), graph_b2b(fn)
    extra = ""
    if "ab" in sys.argv:
        lib.b200mm_tune(1, 0)
        extra = f" | B-resident off {graph_b2b(fn):6.1f}"
        lib.b200mm_tune(1, 1)
    fl = 2.0 * M * N * K
    by = 2.0 * (M * K + N * K + M * N)
    print(f"[{M:7d} x{N:5d} x{K:5d}] {mode:5s} eager {te:7.1f} us  graph {tg:7.1f} us   ({fl / tg / 1e6:6.0f} TF/s, {by / tg / 1e3:6.0f} GB/s; "
          f"floors: tensor {fl / 1359.7e6:5.1f} us, hbm {by / 6553.6e3:5.1f} us){extra}", flush=True)
