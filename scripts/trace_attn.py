"""B200MM_ATTN_TRACE=1 python scripts/trace_attn.py : per-head hand-over timeline of CTA 0 of the WS backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops
dev = torch.device("cuda:0")
B, H, S, p = 256, 12, 128, float(sys.argv[1]) if len(sys.argv) > 1 else 0.0
qkv = torch.randn(B * S, 3 * H * 64, device=dev).to(torch.bfloat16)
kb = ops.mask_to_bias(torch.ones(B, S, dtype=torch.int64, device=dev))
out, lse, mask = ops.attention_fwd(qkv, kb, B, H, S, p_drop=p, seed=7, save_mask=True)
dout = torch.randn_like(out)
os.environ["B200MM_ATTN_TRACE"] = "0"
for _ in range(2):
    ops.attention_bwd(qkv, kb, out, dout, lse, B, H, S, p_drop=p, seed=7, drop_mask=mask)
torch.cuda.synchronize()
os.environ["B200MM_ATTN_TRACE"] = "1"
ops.attention_bwd(qkv, kb, out, dout, lse, B, H, S, p_drop=p, seed=7, drop_mask=mask)
torch.cuda.synchronize()
