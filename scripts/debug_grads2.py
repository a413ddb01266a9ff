import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.nn as nn
import b200mm
from oracle import reference_model as R

def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()

dev = torch.device("cuda:0")
DRY = not torch.cuda.is_available()
img, batch, seq = 64, 16, 32
cfg = R.TowerConfig.tiny(image_size=img)
torch.manual_seed(42)
oracle = R.zero_dropout(R.MultimodalClassifier(2, cfg))
tcfg = b200mm.TextConfig(vocab_size=cfg.vocab_size, max_position_embeddings=cfg.max_position_embeddings,
                         dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, hidden_dim=cfg.hidden_dim,
                         dropout=0.0, attention_dropout=0.0)
if not DRY:
    eng = b200mm.MultimodalClassifier(2, text_config=tcfg, image_config=b200mm.ImageConfig(layers=cfg.resnet_layers),
                                      head_dropout=0.0, device=dev)
    eng.load_reference_state_dict(oracle.state_dict())
data = R.synthetic_batch(batch, seq, cfg)
oracle.train()
if not DRY:
    d = {k: v.to(dev) for k, v in data.items()}; eng.train()
cap = {}
def hook(name):
    def f(mod, gin, gout):
        cap[name] = gout[0].detach().clone()
    return f
rn = oracle.resnet
for m in rn.modules():
    if isinstance(m, nn.ReLU): m.inplace = False
for name, mod in [("avgpool", rn.avgpool), ("layer4", rn.layer4), ("layer3", rn.layer3), ("layer2", rn.layer2), ("layer1", rn.layer1),
                  ("fc", rn.fc), ("l4.ds", rn.layer4[0].downsample[1]), ("l4.conv3", rn.layer4[0].conv3), ("l4.bn2", rn.layer4[0].bn2),
                  ("l4.conv2", rn.layer4[0].conv2), ("l4.bn1", rn.layer4[0].bn1), ("l4.conv1", rn.layer4[0].conv1),
]:
    mod.register_full_backward_hook(hook(name))
crit = nn.CrossEntropyLoss()
loss_ref = crit(oracle(data["text"], data["image"], data["text_mask"]), data["label"]); loss_ref.backward()
print({k: tuple(v.shape) for k, v in cap.items()})
if DRY: sys.exit(0)
eng.zero_grad()
eng.img.debug = {}
eng.train_step_fused(d["text"], d["image"], d["text_mask"], d["label"])
dbg = eng.img.debug
def nchw(x, like):
    N, C, H, W = like.shape
    return x.float().view(N, H, W, C).permute(0, 3, 1, 2)
print("dpooled vs avgpool gout", rel(dbg["dpooled"].view_as(cap["avgpool"]), cap["avgpool"]))
names = ["layer4", "layer3", "layer2", "layer1"]
for i, n in enumerate(names):
    print("d_out into", n, rel(nchw(dbg["d_block_out"][i], cap[n]), cap[n]))
d_y3, dz, d_a2, d_y2, d_a1, d_y1 = dbg["inner"][0]
print("l4 d_y3 (bn3 gin = conv3 gout)", rel(nchw(d_y3, cap["l4.conv3"]), cap["l4.conv3"]))
print("l4 dz (identity gout)", rel(nchw(dz, cap["l4.ds"]), cap["l4.ds"]))
print("l4 d_a2 (bn2... relu gout)", rel(nchw(d_a2, cap["l4.bn2"]), cap["l4.bn2"]), "(vs bn2 gout, which is post-relu-bwd)")
print("l4 d_y2 (conv2 gout)", rel(nchw(d_y2, cap["l4.conv2"]), cap["l4.conv2"]))
print("l4 d_y1 (conv1 gout)", rel(nchw(d_y1, cap["l4.conv1"]), cap["l4.conv1"]))
# relu mask agreement at layer4 output
