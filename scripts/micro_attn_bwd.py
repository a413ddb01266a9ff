import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from b200mm import ops
dev = torch.device("cuda:0"); bf = torch.bfloat16
S = int(sys.argv[1]) if len(sys.argv) > 1 else 128
B, H = (64, 16) if S > 256 else (256, 12)
qkv = torch.randn(B * S, 3 * H * 64, device=dev).to(bf)
kb = ops.mask_to_bias(torch.ones(B, S, dtype=torch.int64, device=dev))
out, lse = ops.attention_fwd(qkv, kb, B, H, S)
dout = torch.randn_like(out)
for _ in range(3):
    ops.attention_bwd(qkv, kb, out, dout, lse, B, H, S)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.attention_bwd(qkv, kb, out, dout, lse, B, H, S)
e1.record(); torch.cuda.synchronize()
print("ok", f"B{B} H{H} S{S} bwd {e0.elapsed_time(e1) / 5 * 1e3:.1f} us")
