"""Tile-width sweep for the text-tower GEMM shapes (block_n 128 vs 256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops
dev = torch.device("cuda:0"); bf = torch.bfloat16
flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
def timeit(fn, iters=8):
    for _ in range(2): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]
for (M, N, K, a_mn, b_mn, epi, res, splits) in [(32768, 768, 768, 0, 0, 0, True, 1), (32768, 768, 768, 0, 1, 0, False, 1),
                                        (32768, 768, 3072, 0, 0, 0, True, 1), (32768, 2304, 768, 0, 0, 0, False, 1),
                                        (50432, 768, 768, 0, 0, 0, True, 1), (768, 768, 32768, 1, 1, 4, False, 17),
                                        (768, 3072, 32768, 1, 1, 4, False, 5), (2304, 768, 32768, 1, 1, 4, False, 6)]:
    A = torch.randn((K, M) if a_mn else (M, K), device=dev).to(bf)
    Bm = torch.randn((K, N) if b_mn else (N, K), device=dev).to(bf)
    out = torch.zeros(M, N, device=dev, dtype=torch.float32 if epi == 4 else bf)
    r = torch.randn(M, N, device=dev).to(bf) if res else None
    bias = torch.zeros(N, device=dev) if epi == 0 and not b_mn else None
    for bn in (128, 256):
        ms = timeit(lambda: ops.gemm_raw(A, bool(a_mn), Bm, bool(b_mn), M, N, K, out, epi=epi, bias=bias, residual=r,
                                         splits=splits, block_n=bn))
        print(f"M{M} N{N} K{K} a{a_mn} b{b_mn} epi{epi} res{int(res)} splits{splits} bn{bn}: {ms*1e3:7.1f} us {2.0*M*N*K/ms/1e9:7.0f} TF/s", flush=True)
