import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
import b200mm
from oracle import reference_model as R

def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()

dev = torch.device("cuda:0")
for (img, batch, seq) in [(64, 16, 32), (128, 32, 32)]:
    cfg = R.TowerConfig.tiny(image_size=img)
    torch.manual_seed(42)
    oracle = R.zero_dropout(R.MultimodalClassifier(2, cfg))
    tcfg = b200mm.TextConfig(vocab_size=cfg.vocab_size, max_position_embeddings=cfg.max_position_embeddings,
                             dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, hidden_dim=cfg.hidden_dim,
                             dropout=0.0, attention_dropout=0.0)
    eng = b200mm.MultimodalClassifier(2, text_config=tcfg, image_config=b200mm.ImageConfig(layers=cfg.resnet_layers),
                                      head_dropout=0.0, device=dev)
    eng.load_reference_state_dict(oracle.state_dict())
    data = R.synthetic_batch(batch, seq, cfg)
    d = {k: v.to(dev) for k, v in data.items()}
    oracle.train(); eng.train()
    crit = nn.CrossEntropyLoss()
    # bf16-rounded-weights fp32 oracle on GPU as a second reference
    loss_ref = crit(oracle(data["text"], data["image"], data["text_mask"]), data["label"]); loss_ref.backward()
    eng.zero_grad()
    _, lf, _ = eng.train_step_fused(d["text"], d["image"], d["text_mask"], d["label"])
    print(f"== img {img} batch {batch}: loss {lf.item():.5f} vs {loss_ref.item():.5f}")
    grads = eng.reference_grad_dict()
    for k, p in oracle.named_parameters():
        if p.grad is None: continue
        e = rel(grads[k].view(p.grad.shape), p.grad)
        if e > 0.03 and ("resnet" in k or "fc" in k or "embeddings" in k):
            print(f"  {k:55s} rel {e:.4f}  |ref| {p.grad.norm().item():.3e}")
