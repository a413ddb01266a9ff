"""Full-size parity report (BASELINE north_star criteria) on one B200: engine (bf16) vs the oracle module run in
fp32 on the same GPU (TF32 off), identical random-init weights and synthetic inputs, dropout 0.
Writes profiles/parity_r01.json: per-layer activation errors, logits error, 200-step loss trajectory, argmax agreement.
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
import b200mm
from oracle import reference_model as R

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
B, S, STEPS = int(os.environ.get("PB", 32)), 128, int(os.environ.get("PSTEPS", 200))

def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()

torch.manual_seed(42)
oracle = R.zero_dropout(R.MultimodalClassifier(2)).to(dev)
eng = b200mm.MultimodalClassifier(2, text_config=b200mm.TextConfig(dropout=0.0, attention_dropout=0.0),
                                  head_dropout=0.0, device=dev)
eng.load_reference_state_dict(oracle.state_dict())
oracle.train(); eng.train()
data = {k: v.to(dev) for k, v in R.synthetic_batch(B, S).items()}

ref_text, ref_img = [], []
hooks = [oracle.bert.embeddings.register_forward_hook(lambda m, i, o: ref_text.append(o.detach()))]
for layer in oracle.bert.transformer.layer:
    hooks.append(layer.register_forward_hook(lambda m, i, o: ref_text.append((o[0] if isinstance(o, tuple) else o).detach())))
for stage in (oracle.resnet.layer1, oracle.resnet.layer2, oracle.resnet.layer3, oracle.resnet.layer4):
    for blk in stage:
        hooks.append(blk.register_forward_hook(lambda m, i, o: ref_img.append(o.detach())))
with torch.no_grad():
    ref_logits = oracle(data["text"], data["image"], data["text_mask"])
for h in hooks: h.remove()
eng.text.capture, eng.img.capture = [], []
with torch.no_grad():
    got_logits = eng._engine_forward(data["text"], data["image"], data["text_mask"], training=True)
rep = {"config": f"ResNet-50 + DistilBERT-multilingual, batch {B}, seq {S}, 224px, dropout 0, train-mode BN",
       "text_layer_rel_err": [rel(g.view(B, S, -1), r) for g, r in zip(eng.text.capture, ref_text)],
       "resnet_block_rel_err": [rel(g.float().view(N, H, W, -1).permute(0, 3, 1, 2), r) for (g, N, H, W), r in zip(eng.img.capture, ref_img)],
       "logits_rel_err": rel(got_logits, ref_logits)}
eng.text.capture = eng.img.capture = None
print(json.dumps(rep)); sys.stdout.flush()

# ---- loss trajectory: fresh batches every step (a stream of synthetic data, like an epoch of the loader)
crit = nn.CrossEntropyLoss()
opt_ref = torch.optim.Adam(oracle.parameters(), lr=2e-5)
opt = b200mm.FusedAdam(eng.parameters(), lr=2e-5)
ref_losses, losses, agree, total = [], [], 0, 0
t0 = time.time()
for step in range(STEPS):
    d = {k: v.to(dev) for k, v in R.synthetic_batch(B, S, seed=5000 + step).items()}
    l, out_ref = R.train_step(oracle, d, crit, opt_ref)
    opt.zero_grad()
    logits, lf, _ = eng.train_step_fused(d["text"], d["image"], d["text_mask"], d["label"])
    opt.step()
    ref_losses.append(l.item()); losses.append(lf.item())
    agree += (logits.argmax(1) == out_ref.argmax(1)).sum().item(); total += B
rel_gap = [abs(a - b) / abs(b) for a, b in zip(losses, ref_losses)]
rep.update({"steps": STEPS, "lr": 2e-5, "loss_engine": losses, "loss_oracle": ref_losses,
            "max_rel_loss_gap": max(rel_gap), "mean_rel_loss_gap": sum(rel_gap) / len(rel_gap),
            "argmax_agreement": agree / total, "seconds": time.time() - t0})
os.makedirs("profiles", exist_ok=True)
json.dump(rep, open("profiles/parity_r01.json", "w"), indent=1)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rep, open("gpurun_out/parity_r01.json", "w"), indent=1)
print({k: v for k, v in rep.items() if k not in ("loss_engine", "loss_oracle")})
