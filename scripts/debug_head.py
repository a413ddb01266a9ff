import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import test_model_gpu as T
dev = torch.device("cuda:0")
oracle, eng, data = T._pair_head(dev)
d = {k: v.to(dev) for k, v in data.items()}
args = ("text", "image", "text_mask", "caption_text", "caption_text_mask")
oracle.train(); eng.train()
caps = {}
def hook(name):
    return lambda m, i, o: caps.__setitem__(name, o.detach())
oracle.text_fc.register_forward_hook(hook("text"))
oracle.caption_text_fc.register_forward_hook(hook("caption"))
oracle.image_model.register_forward_hook(hook("image"))
oracle.image_model.image_model.register_forward_hook(hook("resnet"))
oracle.text_model.register_forward_hook(hook("text_cls"))
oracle.fusion_layer.register_forward_hook(hook("fused"))
oracle.fusion_layer.attention_layer.register_forward_hook(hook("attw"))
ref = oracle(*[data[k] for k in args])
with torch.no_grad():
    eng._prep()
    fused = eng._features(*[d[k] for k in args], True)
(B, cat, sv_t, sv_c, feat, f1, f1d, pi, s_img, z_att, a, m_a, r_a, y, w, r_lin, m_r, r_r, fused) = eng._saved
rel = T.rel
print("text_cls", rel(sv_t[4], caps["text_cls"]))
print("text   ", rel(cat[:, :512], caps["text"]))
print("resnet ", rel(feat, caps["resnet"]))
print("image  ", rel(cat[:, 512:1024], caps["image"]))
print("caption", rel(cat[:, 1024:], caps["caption"]))
print("attw   ", rel(w, caps["attw"]))
print("fused  ", rel(fused, caps["fused"]))
lg, _, _, _ = eng._output(fused, None, False, bn_train=True)
print("logits ", rel(lg, ref.detach()))
z_ref = oracle.output_fc[0](caps["fused"]).squeeze(1)
print("pre-BN logit std over batch", z_ref.std().item(), "mean", z_ref.mean().item())
