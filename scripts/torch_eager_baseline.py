"""Stock PyTorch on the same B200 ("the bar on the GPU box is the stock PyTorch eager/SDPA path running
the same module"): the oracle module (transformers DistilBERT + torchvision ResNet-50, the reference's own libraries) in
bf16 autocast with SDPA attention, channels_last convolutions (cuDNN), torch.optim.Adam(fused=True), batch 256.
A reported baseline only; PB=<batch> selects the batch; writes gpurun_out/torch_eager_baseline_b<batch>.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from oracle import reference_model as R

dev = torch.device("cuda:0")
B, S, STEPS = int(os.environ.get("PB", 256)), 128, 8
torch.backends.cudnn.benchmark = True
torch.manual_seed(42)
cfg = R.TowerConfig()
model = R.MultimodalClassifier(2, cfg)
model.bert = R.build_distilbert(cfg, eager=False)      # SDPA (flash) attention, the library default
model = model.to(dev).to(memory_format=torch.channels_last)
model.train()
crit = nn.CrossEntropyLoss()
opt = torch.optim.Adam(model.parameters(), lr=2e-5, fused=True)
d = {k: v.to(dev) for k, v in R.synthetic_batch(B, S).items()}
d["image"] = d["image"].contiguous(memory_format=torch.channels_last)

def step():
    opt.zero_grad(set_to_none=True)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        out = model(d["text"], d["image"], d["text_mask"])
        loss = crit(out.float(), d["label"])
    loss.backward()
    opt.step()

for _ in range(4):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(STEPS):
    step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / STEPS
rep = {"what": "stock PyTorch %s eager, bf16 autocast, SDPA, channels_last, fused Adam; ResNet-50 + DistilBERT-multilingual train step" % torch.__version__,
       "batch": B, "seq_len": S, "ms_per_step": ms, "samples_per_s": B / ms * 1e3}
print(json.dumps(rep))
os.makedirs("profiles", exist_ok=True); os.makedirs("gpurun_out", exist_ok=True)
json.dump(rep, open(f"gpurun_out/torch_eager_baseline_b{B}.json", "w"), indent=1)
