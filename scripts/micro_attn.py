"""Attention micro-benchmark: forward / backward at the text-tower shape (B=256, H=12, S=128) and others, dropout off / on,
warp-specialised kernels (attention_ws.cu) vs the single-role ones (B200MM_ATTN_WS=0).  CUDA events on the launching
stream; the working set (198 MB fwd) exceeds L2, so back-to-back launches see cold operands.
Also prints the HBM floor: fwd 4 tiles (Q, K, V, O) + LSE, bwd 8 tiles (Q, K, V, O, dO, dQ, dK, dV) + LSE per head."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops
dev = torch.device("cuda:0"); bf = torch.bfloat16
peak = 6553.6e9
try:
    peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] * 1e9
except Exception:
    pass
shapes = [(256, 12, 128), (64, 16, 256), (256, 12, 197)] if len(sys.argv) < 2 else [tuple(int(x) for x in sys.argv[1:4])]
res = []
for B, H, S in shapes:
    qkv = torch.randn(B * S, 3 * H * 64, device=dev).to(bf)
    kb = ops.mask_to_bias(torch.ones(B, S, dtype=torch.int64, device=dev))
    for ws in ("1", "0"):
        os.environ["B200MM_ATTN_WS"] = ws
        for p in (0.0, 0.1):
            out, lse, dmask = ops.attention_fwd(qkv, kb, B, H, S, p_drop=p, seed=7, save_mask=True)   # as the towers call it
            dout = torch.randn_like(out)
            def timed(fn, n=10):
                for _ in range(3):
                    fn()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    fn()
                e1.record(); torch.cuda.synchronize()
                return e0.elapsed_time(e1) / n * 1e3
            tf = timed(lambda: ops.attention_fwd(qkv, kb, B, H, S, p_drop=p, seed=7, save_mask=True))
            tb = timed(lambda: ops.attention_bwd(qkv, kb, out, dout, lse, B, H, S, p_drop=p, seed=7, drop_mask=dmask))
            heads = B * H
            fb = heads * (4 * S * 128 + 4 * S)
            bb = heads * (8 * S * 128 + 4 * S)
            r = dict(B=B, H=H, S=S, ws=int(ws), p_drop=p, fwd_us=round(tf, 1), bwd_us=round(tb, 1),
                     fwd_floor_us=round(fb / peak * 1e6, 1), bwd_floor_us=round(bb / peak * 1e6, 1),
                     fwd_frac_of_hbm=round(fb / peak * 1e6 / tf, 3), bwd_frac_of_hbm=round(bb / peak * 1e6 / tb, 3),
                     fwd_tflops=round(heads * 4 * S * S * 64 / tf / 1e6, 1), bwd_tflops=round(heads * 10 * S * S * 64 / tb / 1e6, 1))
            res.append(r)
            print(json.dumps(r), flush=True)
os.environ["B200MM_ATTN_WS"] = "1"
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/micro_attn.json", "w"), indent=1)
