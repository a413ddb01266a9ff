"""Micro-benchmark of individual ops at config-2 shapes (for ncu captures and quick A/B timing)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops
dev = torch.device("cuda:0")
bf = torch.bfloat16
which = sys.argv[1:] or ["bn", "col2im", "gemm"]
def timeit(name, fn, bytes_=None, flops=None, iters=10):
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    for _ in range(2): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()   # evict L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    s = f"{name:44s} {ms*1e3:8.1f} us"
    if bytes_: s += f"  {bytes_/ms/1e6:7.0f} GB/s"
    if flops: s += f"  {flops/ms/1e9:7.0f} TF/s"
    print(s, flush=True)
if "bn" in which:
    for (M, C) in [(802816, 64), (802816, 256), (200704, 512), (50176, 1024), (12544, 2048)]:
        x = torch.randn(M, C, device=dev).to(bf); g = torch.ones(C, device=dev); b = torch.zeros(C, device=dev)
        rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev)
        out, mean, rstd = ops.batchnorm_fwd(x, g, b, rm, rv)
        timeit(f"bn_fwd  [{M},{C}]", lambda: ops.batchnorm_fwd(x, g, b, rm, rv), bytes_=3 * M * C * 2)
        dout = torch.randn(M, C, device=dev).to(bf); dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
        timeit(f"bn_bwd  [{M},{C}]", lambda: ops.batchnorm_bwd(dout, out, x, mean, rstd, g, dg, db), bytes_=7 * M * C * 2)
        timeit(f"bn_bwd  [{M},{C}] mask from x", lambda: ops.batchnorm_bwd(dout, None, x, mean, rstd, g, dg, db, beta=b), bytes_=5 * M * C * 2)
        st = torch.zeros(2 * C, device=dev); st[:C] = x.float().sum(0); st[C:] = (x.float() ** 2).sum(0)
        timeit(f"bn_fwd  [{M},{C}] given stats", lambda: ops.batchnorm_fwd(x, g, b, rm, rv, col_stats=st), bytes_=2 * M * C * 2)
if "col2im" in which:
    for (N, H, C, s) in [(256, 56, 64, 1), (256, 56, 128, 2), (256, 14, 256, 1)]:
        Ho = H // s
        dcols = torch.randn(N * Ho * Ho, 9 * C, device=dev).to(bf)
        x = torch.randn(N * H * H, C, device=dev).to(bf)
        timeit(f"col2im N{N} H{H} C{C} s{s}", lambda: ops.col2im(dcols, N, H, H, C, 3, s, 1), bytes_=dcols.numel() * 2 + N * H * H * C * 2)
        timeit(f"im2col N{N} H{H} C{C} s{s}", lambda: ops.im2col(x, N, H, H, C, 3, s, 1), bytes_=dcols.numel() * 2 + N * H * H * C * 2)
    img = torch.randn(256, 3, 224, 224, device=dev)
    timeit("im2col_nchw stem", lambda: ops.im2col_nchw_f32(img, 7, 2, 3, 152), bytes_=img.numel() * 4 + 256 * 112 * 112 * 152 * 2)
    a0 = torch.randn(256 * 112 * 112, 64, device=dev).to(bf)
    timeit("maxpool_fwd", lambda: ops.maxpool_fwd(a0, 256, 112, 112, 64), bytes_=a0.numel() * 2 * 1.375)
if "gemm_dgelu" in which:
    shapes = [(32768, 3072, 768, 1, 2)]
    which.append("gemm")
elif "gemm_res" in which:
    shapes = [(802816, 256, 64, 1, 7)]
    which.append("gemm")
elif "gemmx" in which:
    shapes = [(32768, 3072, 768, 1, 0), (32768, 3072, 768, 0, 2), (32768, 3072, 768, 1, 2), (32768, 3072, 768, 0, 0),
              (32768, 3072, 768, 0, 1), (32768, 768, 768, 0, 0), (32768, 768, 768, 0, 7), (32768, 768, 3072, 0, 7),
              (802816, 256, 64, 1, 0), (802816, 256, 64, 1, 7), (802816, 256, 64, 1, 8), (200704, 512, 128, 1, 7),
              (200704, 512, 128, 1, 8), (50176, 1024, 256, 1, 7), (50176, 1024, 256, 1, 8)]
    which.append("gemm")
elif "gemm1" in which:
    shapes = [(802816, 256, 64, 0, 0)]
elif "gemm" in which:
    shapes = [(32768, 3072, 768, 0, 1), (32768, 3072, 768, 1, 2), (32768, 2304, 768, 0, 0), (32768, 768, 768, 0, 0),
                                 (802816, 256, 64, 0, 0), (802816, 576, 64, 1, 0), (802816, 64, 576, 0, 0), (200704, 1152, 128, 1, 0),
                                 (3211264, 64, 152, 0, 0), (50176, 2304, 256, 1, 0), (8192, 8192, 8192, 0, 0)]
if "gemm" in which or "gemm1" in which:
    for (M, N, K, b_mn, epi) in shapes:
        res = None
        inplace = False
        if epi == 8:     # plain store, residual aliasing the output (TMA reduce-add)
            epi, inplace = 0, True
        if epi == 7:     # plain store + residual
            epi, res = 0, torch.randn(M, N, device=dev).to(bf)
        A = torch.randn(M, K, device=dev).to(bf)
        Bm = torch.randn(N, K, device=dev).to(bf) if not b_mn else torch.randn(K, N, device=dev).to(bf)
        out = torch.empty(M, N, device=dev, dtype=bf); out2 = torch.empty(M, N, device=dev, dtype=bf) if epi == 1 else None
        if inplace:
            out.zero_(); res = out
        aux = torch.randn(M, N, device=dev).to(bf) if epi == 2 else None
        bias = torch.zeros(N, device=dev) if epi != 2 else None   # dgrad epilogues carry no bias
        timeit(f"gemm M{M} N{N} K{K} b_mn{b_mn} epi{epi}{'+res' if res is not None else ''}{' (in place)' if inplace else ''}", lambda: ops.gemm_raw(A, False, Bm, bool(b_mn), M, N, K, out, epi=epi, bias=bias, aux=aux, out2=out2, residual=res),
               flops=2.0 * M * N * K, bytes_=2.0 * (M * K + N * K + M * N * (2 if epi == 1 else 1) + (M * N if epi == 2 else 0)))
if "attn" in which:
    for (B, H, S) in [(256, 12, 128), (256, 12, 197)]:
        qkv = torch.randn(B * S, 3 * H * 64, device=dev).to(bf)
        mask = torch.ones(B, S, dtype=torch.int64, device=dev)
        kb = ops.mask_to_bias(mask)
        for pdrop in (0.0, 0.1):
            out, lse = ops.attention_fwd(qkv, kb, B, H, S, p_drop=pdrop, seed=1)
            dout = torch.randn_like(out)
            fl = 4.0 * S * S * 64 * B * H
            timeit(f"attn_fwd B{B} H{H} S{S} p{pdrop}", lambda: ops.attention_fwd(qkv, kb, B, H, S, p_drop=pdrop, seed=1),
                   flops=fl, bytes_=qkv.numel() * 2 + out.numel() * 2)
            timeit(f"attn_bwd B{B} H{H} S{S} p{pdrop}", lambda: ops.attention_bwd(qkv, kb, out, dout, lse, B, H, S, p_drop=pdrop, seed=1),
                   flops=2.5 * fl, bytes_=2 * qkv.numel() * 2 + 2 * out.numel() * 2)
if "ln" in which:
    for (M, D) in [(32768, 768), (50432, 768), (16384, 1024)]:
        x = torch.randn(M, D, device=dev).to(bf); g = torch.ones(D, device=dev); b = torch.zeros(D, device=dev)
        y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12)
        dy = torch.randn_like(x); dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev)
        timeit(f"ln_fwd [{M},{D}]", lambda: ops.layernorm_fwd(x, g, b, 1e-12), bytes_=2 * M * D * 2)
        timeit(f"ln_bwd [{M},{D}]", lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db), bytes_=3 * M * D * 2)
        timeit(f"ln_bwd [{M},{D}] +addend", lambda: ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db, addend=dy), bytes_=4 * M * D * 2)
        timeit(f"colsum [{M},{D}]", lambda: ops.colsum(dy, db), bytes_=M * D * 2)
