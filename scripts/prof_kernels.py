"""One or two launches of each non-GEMM hot kernel at its BASELINE-config shape, for `ncu --set full` (attention,
LayerNorm, BatchNorm, Adam, preprocessing): the kernels north_star asks ncu evidence for besides the GEMM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops
dev = torch.device("cuda:0"); bf = torch.bfloat16
which = set(sys.argv[1:]) or {"attn", "ln", "bn", "adam", "pre"}
if "attn" in which:
    for B, H, S, p in ((256, 12, 128, 0.1), (256, 12, 197, 0.0)):
        qkv = torch.randn(B * S, 3 * H * 64, device=dev).to(bf)
        kb = ops.mask_to_bias(torch.ones(B, S, dtype=torch.int64, device=dev))
        out, lse = ops.attention_fwd(qkv, kb, B, H, S, p_drop=p, seed=7)
        dout = torch.randn_like(out)
        ops.attention_bwd(qkv, kb, out, dout, lse, B, H, S, p_drop=p, seed=7)
if "ln" in which:
    M, D = 32768, 768
    x = torch.randn(M, D, device=dev).to(bf); g = torch.ones(D, device=dev); b = torch.zeros(D, device=dev)
    y, mean, rstd = ops.layernorm_fwd(x, g, b, 1e-12)
    dy = torch.randn_like(x); dg = torch.zeros(D, device=dev); db = torch.zeros(D, device=dev)
    ops.layernorm_bwd(dy, x, mean, rstd, g, dg, db)
if "bn" in which:
    for M, C in ((802816, 256), (50176, 256), (12544, 512)):
        x = torch.randn(M, C, device=dev).to(bf); dout = torch.randn(M, C, device=dev).to(bf)
        g = torch.rand(C, device=dev) + 0.5; b = torch.randn(C, device=dev) * 0.1
        rm = torch.zeros(C, device=dev); rv = torch.ones(C, device=dev)
        xf = x.float(); stats = torch.cat([xf.sum(0), (xf * xf).sum(0)]).contiguous(); del xf
        out, mean, rstd = ops.batchnorm_fwd(x, g, b, rm, rv, relu=True, col_stats=stats)
        dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
        ops.batchnorm_bwd(dout, None, x, mean, rstd, g, dg, db, relu=True, beta=b)
        del x, dout, out
if "adam" in which:
    n = 161_700_000 // 64 * 64
    p_ = torch.randn(n, device=dev); g_ = torch.randn(n, device=dev); m_ = torch.zeros(n, device=dev); v_ = torch.zeros(n, device=dev)
    sh = torch.empty(n, device=dev, dtype=bf)
    ops.adam_step(p_, g_, m_, v_, sh, lr=2e-5, step=1)
if "pre" in which:
    imgs = torch.randint(0, 256, (256, 224, 224, 3), dtype=torch.uint8, device=dev)
    ops.u8_normalize(imgs)
    big = [torch.randint(0, 256, (480, 640, 3), dtype=torch.uint8) for _ in range(64)]
    buf, table = ops.pack_images(big)
    ops.preprocess_u8_packed(buf.to(dev), table.to(dev))
torch.cuda.synchronize()
print("ok")
