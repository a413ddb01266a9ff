"""Full-size parity report for the ViT + BERT-family configurations (BASELINE configs 3 / 4): engine (bf16) vs the
oracle module in fp32 on the same GPU (TF32 off), identical random-init weights and synthetic inputs, dropout 0.
PCFG=3: ViT-B/16 + BERT-base (seq 128);  PCFG=4: ViT-L/14 + XLM-R-large (seq 256).
Writes profiles/parity_r01_cfg<N>.json: per-layer activation errors of both towers, logits error, loss trajectory,
argmax agreement."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
import b200mm
from oracle import reference_model as R

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")
CFG = int(os.environ.get("PCFG", 3))
B, STEPS = int(os.environ.get("PB", 16)), int(os.environ.get("PSTEPS", 100))
cfg = R.TowerConfig.vit_b16_bert_base() if CFG == 3 else R.TowerConfig.vit_l14_xlmr_large()
S = 128 if CFG == 3 else 256

def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()

torch.manual_seed(42)
oracle = R.zero_dropout(R.MultimodalClassifier(2, cfg)).to(dev)
tc = b200mm.TextConfig.bert_base(dropout=0.0, attention_dropout=0.0) if CFG == 3 else \
    b200mm.TextConfig.xlmr_large(dropout=0.0, attention_dropout=0.0)
vc = b200mm.ViTConfig.vit_b16() if CFG == 3 else b200mm.ViTConfig.vit_l14()
eng = b200mm.MultimodalClassifier(2, text_config=tc, image_config=vc, head_dropout=0.0, device=dev)
eng.load_reference_state_dict(oracle.state_dict())
oracle.train(); eng.train()
data = {k: v.to(dev) for k, v in R.synthetic_batch(B, S, cfg).items()}

ref_text, ref_img = [], []
hooks = [oracle.bert.embeddings.register_forward_hook(lambda m, i, o: ref_text.append(o.detach()))]
for layer in oracle.bert.encoder.layer:
    hooks.append(layer.register_forward_hook(lambda m, i, o: ref_text.append((o[0] if isinstance(o, tuple) else o).detach())))
hooks.append(oracle.resnet.embeddings.register_forward_hook(lambda m, i, o: ref_img.append(o.detach())))
for layer in oracle.resnet.encoder.layer:
    hooks.append(layer.register_forward_hook(lambda m, i, o: ref_img.append((o[0] if isinstance(o, tuple) else o).detach())))
with torch.no_grad():
    ref_logits = oracle(data["text"], data["image"], data["text_mask"])
for h in hooks: h.remove()
eng.text.capture, eng.img.capture = [], []
with torch.no_grad():
    got_logits = eng._engine_forward(data["text"], data["image"], data["text_mask"], training=True)
rep = {"config": f"BASELINE config {CFG}: {type(oracle.resnet).__name__} + {type(oracle.bert).__name__}, batch {B}, seq {S}, dropout 0",
       "text_layer_rel_err": [rel(g.view(B, S, -1), r) for g, r in zip(eng.text.capture, ref_text)],
       "vit_layer_rel_err": [rel(g.view(B, r.shape[1], -1), r) for g, r in zip(eng.img.capture, ref_img)],
       "logits_rel_err": rel(got_logits, ref_logits)}
eng.text.capture = eng.img.capture = None
print(json.dumps(rep)); sys.stdout.flush()

crit = nn.CrossEntropyLoss()
opt_ref = torch.optim.Adam(oracle.parameters(), lr=2e-5)
opt = b200mm.FusedAdam(eng.parameters(), lr=2e-5)
ref_losses, losses, agree, total = [], [], 0, 0
t0 = time.time()
for step in range(STEPS):
    d = {k: v.to(dev) for k, v in R.synthetic_batch(B, S, cfg, seed=7000 + step).items()}
    l, out_ref = R.train_step(oracle, d, crit, opt_ref)
    opt.zero_grad()
    logits, lf, _ = eng.train_step_fused(d["text"], d["image"], d["text_mask"], d["label"])
    opt.step()
    ref_losses.append(l.item()); losses.append(lf.item())
    agree += (logits.argmax(1) == out_ref.argmax(1)).sum().item(); total += B
gap = [abs(a - b) / abs(b) for a, b in zip(losses, ref_losses)]
rep.update({"steps": STEPS, "lr": 2e-5, "loss_engine": losses, "loss_oracle": ref_losses, "max_rel_loss_gap": max(gap),
            "mean_rel_loss_gap": sum(gap) / len(gap), "argmax_agreement": agree / total, "seconds": time.time() - t0})
for dname in ("profiles", "gpurun_out"):
    os.makedirs(dname, exist_ok=True)
    json.dump(rep, open(f"{dname}/parity_r01_cfg{CFG}.json", "w"), indent=1)
print({k: v for k, v in rep.items() if k not in ("loss_engine", "loss_oracle")})
