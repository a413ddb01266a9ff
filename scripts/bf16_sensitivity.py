"""How well conditioned is the random-init reference network?  (CPU, a few seconds.)
Rounds (a) only the conv/linear weights, (b) weights + every stored activation of the fp32 oracle ResNet-50 to bf16
and reports the relative change of every Bottleneck output and of the 1000-way logits.  Evidence for DESIGN.md
'parity': the deviation of a bf16 engine from the fp32 oracle behind the deep image blocks is a property of the
network at random init, not of the implementation."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
from torchvision.models.resnet import Bottleneck
from oracle import reference_model as R
from oracle import bf16_emulation as E

torch.manual_seed(42)
net = R.build_resnet(R.TowerConfig()).train()
x = torch.randn(int(os.environ.get("SB", 8)), 3, 224, 224)
rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-12)).item()

def blocks_and_logits(model, emulate):
    outs, hooks = [], []
    for m in model.modules():
        if isinstance(m, Bottleneck):
            hooks.append(m.register_forward_hook(lambda mod, i, o: outs.append(o.detach().clone())))
    class W(nn.Module):
        def __init__(s, r): super().__init__(); s.resnet = r
        def forward(s, x): return s.resnet(x)
    w = W(model)
    with torch.no_grad():
        if emulate:
            with E.bf16_storage(w):
                y = w(x)
        else:
            y = w(x)
    for h in hooks: h.remove()
    return outs, y

ref, yref = blocks_and_logits(net, False)
import copy
net_w = E.round_gemm_weights_(copy.deepcopy(net))
o1, y1 = blocks_and_logits(net_w, False)
o2, y2 = blocks_and_logits(net_w, True)
rep = {"weights_only": {"blocks": [rel(a, b) for a, b in zip(o1, ref)], "logits": rel(y1, yref)},
       "weights_and_activations": {"blocks": [rel(a, b) for a, b in zip(o2, ref)], "logits": rel(y2, yref)},
       "activations_only_vs_rounded_weight_fp32": {"blocks": [rel(a, b) for a, b in zip(o2, o1)], "logits": rel(y2, y1)}}
print(json.dumps(rep, indent=1))
os.makedirs("profiles", exist_ok=True)
json.dump(rep, open("profiles/bf16_sensitivity_r01.json", "w"), indent=1)
