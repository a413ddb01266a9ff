"""Two launches each of the attention forward and backward at the text-tower shape, for `ncu -k regex:attn_`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops
dev = torch.device("cuda:0")
B, H, S = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (256, 12, 128)
p = float(sys.argv[4]) if len(sys.argv) >= 5 else 0.1
qkv = torch.randn(B * S, 3 * H * 64, device=dev).to(torch.bfloat16)
kb = ops.mask_to_bias(torch.ones(B, S, dtype=torch.int64, device=dev))
for _ in range(2):
    out, lse = ops.attention_fwd(qkv, kb, B, H, S, p_drop=p, seed=7)
    dout = torch.randn_like(out)
    dq = ops.attention_bwd(qkv, kb, out, dout, lse, B, H, S, p_drop=p, seed=7)
torch.cuda.synchronize()
print("ok", float(dq.float().abs().mean()))
