import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops
dev = torch.device("cuda:0"); bf = torch.bfloat16
B, H, S = 256, 12, int(sys.argv[1]) if len(sys.argv) > 1 else 128
qkv = torch.randn(B * S, 3 * H * 64, device=dev).to(bf)
kb = ops.mask_to_bias(torch.ones(B, S, dtype=torch.int64, device=dev))
out, lse = ops.attention_fwd(qkv, kb, B, H, S)
dout = torch.randn_like(out)
for _ in range(3):
    ops.attention_bwd(qkv, kb, out, dout, lse, B, H, S)
torch.cuda.synchronize()
print("ok")
