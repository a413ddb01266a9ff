"""CPU fp32 oracle for the organiser 2C baseline (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates, module for module, the reference's late-fusion classifier and training step:

  * ``MultimodalClassifier``         example_scripts/Multimodal_example_task2C.txt:152-197
  * ``train_step`` (zero_grad/fwd/CE/bwd/Adam)          ...txt:205-217, loss/optim ...txt:248-249
  * ``train`` / ``test`` / ``evaluate`` loops           ...txt:200-242, 259-280

The only change w.r.t. the reference is the one the task mandates: there is no network, so
``AutoModel.from_pretrained('distilbert-base-multilingual-cased')`` (.txt:158) becomes
``DistilBertModel(DistilBertConfig(vocab_size=119547))`` and ``models.resnet50(pretrained=True)``
(.txt:164) becomes ``resnet50(weights=None)`` -- same architectures, random init.  The arithmetic itself
lives in un-vendored third-party code the reference pins in poetry.lock (transformers 4.39.2,
torchvision 0.17.2, torch 2.2.2); here transformers 5.5 / torchvision 0.26 / torch 2.11 execute the same
post-LN DistilBERT and ResNet-50 v1.5 maths.

PARITY PINNING: the reference holds no golden activations for this path.  The oracle is therefore
pinned against OUTPUTS OF THE REFERENCE ITSELF RUN IN THE BUILD CONTAINER: tests/golden/make_reference_golden.py executes
the source text of the reference's classes and loop functions verbatim (organiser .txt:152-242; participant .py:307-392,
476-499, 562-685, 689-879) -- only the network-bound constructors return from-config modules -- and commits logits, loss,
every parameter's gradient norm, a whole train()/test()/evaluate() pass and the TSVs it wrote
(tests/golden/reference_run_golden.pt); tests/test_cpu.py demands that this oracle and the repo's host-side loops
reproduce them.  What that cannot pin is the third-party arithmetic of the reference's LOCKED versions (4.39.2 / 0.17.2 /
2.2.2 are not installable here).  The other reference artefacts -- TSV schema, combine_preds known answers, scorer,
k-fold split -- are covered in oracle/ensemble.py and tests/golden/.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.nn as nn

PAD_ID = 0
TRAIN_PRIOR = 603 / 2143  # fraction of 'propaganda' in the reference's train split


@dataclass
class TowerConfig:
    """Shapes of the two towers. Defaults = BASELINE.json configs 1-2 (DistilBERT-multilingual + ResNet-50)."""
    vocab_size: int = 119547
    max_position_embeddings: int = 512
    dim: int = 768
    n_layers: int = 6
    n_heads: int = 12
    hidden_dim: int = 3072
    dropout: float = 0.1
    attention_dropout: float = 0.1
    head_dropout: float = 0.3   # bert_drop, .txt:160
    resnet_layers: tuple = (3, 4, 6, 3)  # ResNet-50
    resnet_width: int = 64
    image_size: int = 224
    num_classes: int = 2
    # --- BASELINE configs 3-5 / HEAD-script towers (same head, towers swapped)
    text_arch: str = "distilbert"        # 'distilbert' | 'bert' | 'roberta' (RoBERTa / XLM-R)
    layer_norm_eps: float = 1e-12
    pad_token_id: int = 0
    type_vocab_size: int = 2
    pooling: str = "last"                # 'last' = h[:, -1] (.txt:178) ; 'cls' = h[:, 0] (HEAD .py:359-360)
    image_arch: str = "resnet"           # 'resnet' | 'vit'
    vit_patch: int = 16
    vit_dim: int = 768
    vit_layers: int = 12
    vit_heads: int = 12
    vit_hidden: int = 3072

    @staticmethod
    def tiny(**kw) -> "TowerConfig":
        """A small instance of the same graph for fast CPU/GPU parity tests."""
        base = dict(vocab_size=1024, max_position_embeddings=128, dim=128, n_layers=2, n_heads=2,
                    hidden_dim=256, resnet_layers=(1, 1, 1, 1), image_size=64)
        base.update(kw)
        return TowerConfig(**base)

    @staticmethod
    def vit_b16_bert_base(**kw) -> "TowerConfig":
        """BASELINE configs 3 / 5: ViT-B/16 + BERT-base (AraBERT-shaped, vocab 64000), CLS pooling."""
        base = dict(vocab_size=64000, n_layers=12, text_arch="bert", pooling="cls", image_arch="vit")
        base.update(kw)
        return TowerConfig(**base)

    @staticmethod
    def vit_l14_xlmr_large(**kw) -> "TowerConfig":
        """BASELINE config 4: ViT-L/14 + XLM-R-large."""
        base = dict(vocab_size=250002, max_position_embeddings=514, dim=1024, n_layers=24, n_heads=16,
                    hidden_dim=4096, text_arch="roberta", layer_norm_eps=1e-5, pad_token_id=1, type_vocab_size=1,
                    pooling="cls", image_arch="vit", vit_patch=14, vit_dim=1024, vit_layers=24, vit_heads=16,
                    vit_hidden=4096)
        base.update(kw)
        return TowerConfig(**base)

    @staticmethod
    def tiny_vit_bert(text_arch: str = "bert", **kw) -> "TowerConfig":
        """Small instance of the config-3/4/5 graph (ViT + BERT / RoBERTa-style text tower)."""
        rob = text_arch == "roberta"
        base = dict(vocab_size=1024, max_position_embeddings=130, dim=128, n_layers=2, n_heads=2, hidden_dim=256,
                    text_arch=text_arch, layer_norm_eps=1e-5 if rob else 1e-12, pad_token_id=1 if rob else 0,
                    type_vocab_size=1 if rob else 2, pooling="cls", image_arch="vit", image_size=64, vit_patch=16,
                    vit_dim=128, vit_layers=2, vit_heads=2, vit_hidden=256)
        base.update(kw)
        return TowerConfig(**base)


def build_distilbert(cfg: TowerConfig, eager: bool = True):
    from transformers import DistilBertConfig, DistilBertModel
    hf = DistilBertConfig(vocab_size=cfg.vocab_size, max_position_embeddings=cfg.max_position_embeddings,
                          dim=cfg.dim, n_layers=cfg.n_layers, n_heads=cfg.n_heads, hidden_dim=cfg.hidden_dim,
                          dropout=cfg.dropout, attention_dropout=cfg.attention_dropout)
    if eager:
        hf._attn_implementation = "eager"
    return DistilBertModel(hf)


def build_bert(cfg: TowerConfig, eager: bool = True):
    """BERT-base / AraBERT (HEAD script ``AutoModel.from_pretrained(text_model)``, Multimodal_example_task2C.py:317)
    or RoBERTa / XLM-R (``caption_text_model``, :337; BASELINE config 4) from config -- random init, no network."""
    if cfg.text_arch == "bert":
        from transformers import BertConfig as C, BertModel as M
    else:
        from transformers import XLMRobertaConfig as C, XLMRobertaModel as M
    hf = C(vocab_size=cfg.vocab_size, hidden_size=cfg.dim, num_hidden_layers=cfg.n_layers,
           num_attention_heads=cfg.n_heads, intermediate_size=cfg.hidden_dim,
           max_position_embeddings=cfg.max_position_embeddings, hidden_dropout_prob=cfg.dropout,
           attention_probs_dropout_prob=cfg.attention_dropout, layer_norm_eps=cfg.layer_norm_eps,
           pad_token_id=cfg.pad_token_id, type_vocab_size=cfg.type_vocab_size)
    if eager:
        hf._attn_implementation = "eager"
    return M(hf)


def build_text(cfg: TowerConfig, eager: bool = True):
    return build_distilbert(cfg, eager) if cfg.text_arch == "distilbert" else build_bert(cfg, eager)


def build_vit(cfg: TowerConfig, eager: bool = True):
    """transformers ViTModel without pooler: CLS token, pre-LN blocks, final LayerNorm (what the reference's
    commented ``vit_base_patch16_224`` / ``ViTModel`` lines point at; Multimodal_example_task2C.py:82,
    mm_model_mm_example_task2C.py:66-67)."""
    from transformers import ViTConfig, ViTModel
    hf = ViTConfig(hidden_size=cfg.vit_dim, num_hidden_layers=cfg.vit_layers, num_attention_heads=cfg.vit_heads,
                   intermediate_size=cfg.vit_hidden, image_size=cfg.image_size, patch_size=cfg.vit_patch)
    if eager:
        hf._attn_implementation = "eager"
    return ViTModel(hf, add_pooling_layer=False)


def build_image(cfg: TowerConfig):
    return build_resnet(cfg) if cfg.image_arch == "resnet" else build_vit(cfg)


def build_resnet(cfg: TowerConfig):
    from torchvision.models.resnet import ResNet, Bottleneck
    # torchvision.models.resnet50 == ResNet(Bottleneck, [3, 4, 6, 3]); num_classes stays 1000 (.txt:164-165)
    return ResNet(Bottleneck, list(cfg.resnet_layers), num_classes=1000)


class MultimodalClassifier(nn.Module):
    """Reference: example_scripts/Multimodal_example_task2C.txt:152-197 (same attribute names)."""

    def __init__(self, num_classes: int = 2, cfg: TowerConfig | None = None):
        super().__init__()
        cfg = cfg or TowerConfig()
        self.cfg = cfg
        self.bert = build_text(cfg)                            # .txt:158
        self.bert_drop = nn.Dropout(cfg.head_dropout)          # .txt:160
        self.bert_fc = nn.Linear(cfg.dim, 512)                 # .txt:161
        self.resnet = build_image(cfg)                         # .txt:164
        self.resnet_fc = nn.Linear(1000 if cfg.image_arch == "resnet" else cfg.vit_dim, 512)   # .txt:165
        self.fusion_fc = nn.Linear(1024, 512)                  # .txt:168
        self.output_fc = nn.Linear(512, num_classes)           # .txt:170

    def forward(self, text, image, mask):
        bert_output = self.bert(text, attention_mask=mask, return_dict=False)   # .txt:175
        tok = -1 if self.cfg.pooling == "last" else 0                           # HEAD script pools CLS (.py:359-360)
        bert_output = self.bert_drop(bert_output[0][:, tok, :])                 # .txt:178 (LAST position)
        bert_output = self.bert_fc(bert_output)                                 # .txt:179
        if self.cfg.image_arch == "resnet":
            resnet_output = self.resnet(image)                                  # .txt:183
        else:
            resnet_output = self.resnet(pixel_values=image).last_hidden_state[:, 0]   # ViT CLS feature
        resnet_output = self.resnet_fc(resnet_output)                           # .txt:184
        features = torch.cat((bert_output, resnet_output), dim=1)               # .txt:190
        features = self.fusion_fc(features)                                     # .txt:193
        return self.output_fc(features)                                         # .txt:195


def zero_dropout(model: nn.Module) -> nn.Module:
    """Parity runs use p = 0 (bit-parity with torch's RNG stream is not a goal)."""
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    if hasattr(model, "bert") and hasattr(model.bert, "transformer"):
        for layer in model.bert.transformer.layer:
            att = layer.attention
            if hasattr(att, "dropout") and isinstance(att.dropout, nn.Dropout):
                att.dropout.p = 0.0
            if hasattr(att, "dropout_prob"):
                att.dropout_prob = 0.0
            if hasattr(att, "config"):
                att.config.attention_dropout = 0.0
    return model


def synthetic_batch(batch: int, seq_len: int, cfg: TowerConfig | None = None, seed: int = 1234,
                    device: str = "cpu"):
    """Synthetic inputs of SURVEY.md §8d: N(0,1) pixels, uniform ids with random real lengths, pad beyond."""
    cfg = cfg or TowerConfig()
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(batch, 3, cfg.image_size, cfg.image_size, generator=g)
    lo = min(1000, cfg.vocab_size // 2)
    ids = torch.randint(lo, cfg.vocab_size, (batch, seq_len), generator=g)
    lengths = torch.randint(min(8, seq_len), seq_len + 1, (batch,), generator=g)
    lengths[0] = seq_len
    pos = torch.arange(seq_len).unsqueeze(0)
    mask = (pos < lengths.unsqueeze(1)).long()
    ids = ids * mask + cfg.pad_token_id * (1 - mask)
    labels = (torch.rand(batch, generator=g) < TRAIN_PRIOR).long()
    return {"text": ids.to(device), "text_mask": mask.to(device), "image": image.to(device),
            "label": labels.to(device)}


def train_step(model, batch, criterion, optimizer):
    """One iteration of the reference loop body, .txt:205-217. Returns (loss, logits)."""
    optimizer.zero_grad()
    output = model(batch["text"], batch["image"], batch["text_mask"])
    loss = criterion(output, batch["label"])
    loss.backward()
    optimizer.step()
    return loss.detach(), output.detach()


def train(model, train_loader, criterion, optimizer, device):
    """Reference: .txt:200-223 (tqdm removed)."""
    model.train()
    train_loss = 0.0
    correct = 0
    n = 0
    for data in train_loader:
        optimizer.zero_grad()
        text = data["text"].to(device)
        image = data["image"].to(device)
        mask = data["text_mask"].to(device)
        labels = data["label"].to(device)
        output = model(text, image, mask)
        loss = criterion(output, labels)
        loss.backward()
        optimizer.step()
        train_loss += loss.item() * labels.size(0)
        _, predicted = torch.max(output, 1)
        correct += (predicted == labels).sum().item()
        n += labels.size(0)
    return train_loss / n, correct / n


def test(model, test_loader, criterion, device):
    """Reference: .txt:225-242."""
    model.eval()
    test_loss = 0.0
    correct = 0
    n = 0
    with torch.no_grad():
        for data in test_loader:
            output = model(data["text"].to(device), data["image"].to(device), data["text_mask"].to(device))
            labels = data["label"].to(device)
            loss = criterion(output, labels)
            test_loss += loss.item() * labels.size(0)
            _, predicted = torch.max(output, 1)
            correct += (predicted == labels).sum().item()
            n += labels.size(0)
    return test_loss / n, correct / n


def evaluate(model, test_loader, device, out_path="task2C_TeamName.tsv", run_id="DistilBERT+ResNet"):
    """Reference: .txt:259-280 -- argmax predictions to a 3-column TSV."""
    model.eval()
    rows = []
    id2l = {0: "not_propaganda", 1: "propaganda"}
    with torch.no_grad():
        for data in test_loader:
            output = model(data["text"].to(device), data["image"].to(device), data["text_mask"].to(device))
            _, predicted = torch.max(output, 1)
            for i, l in zip(data["id"], predicted.tolist()):
                rows.append((i, id2l[l]))
    with open(out_path, "w") as f:
        f.write("id\tlabel\trun_id\n")
        for i, l in rows:
            f.write(f"{i}\t{l}\t{run_id}\n")
    return rows


def cpu_train_throughput(batch: int = 16, seq_len: int = 128, steps: int = 3, warmup: int = 1,
                         threads: int | None = None, cfg: TowerConfig | None = None):
    """CPU baseline of BASELINE.md §4: the oracle's full fp32 train step on the host cores."""
    import time
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(42)
    model = MultimodalClassifier(2, cfg)
    model.train()
    criterion = nn.CrossEntropyLoss()
    optimizer = torch.optim.Adam(model.parameters(), lr=2e-5)
    data = synthetic_batch(batch, seq_len, cfg)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        train_step(model, data, criterion, optimizer)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    best = min(times)
    med = sorted(times)[len(times) // 2]
    return {"samples_per_s": batch / med, "best_s": best, "median_s": med, "cores": threads,
            "batch": batch, "seq_len": seq_len, "steps": steps}


# ------------------------------------------------------------------------------------------------- HEAD-script model
class LLMWithClassificationHead(nn.Module):
    """Reference: example_scripts/Multimodal_example_task2C.py:307-360, ``pooling_type="cls"`` (the only one the
    script selects, :589-591, :603-605); ``AutoModel.from_pretrained`` -> from-config random init."""

    def __init__(self, model: nn.Module):
        super().__init__()
        self.model = model

    def forward(self, input_ids, attention_mask):
        return self.model(input_ids=input_ids, attention_mask=attention_mask).last_hidden_state[:, 0]   # :359-360


class ConcatAttention3(nn.Module):
    """Reference: :476-499."""

    def __init__(self, input_dim, attention_dim):
        super().__init__()
        self.attention_layer = nn.Sequential(nn.Linear(input_dim, input_dim), nn.BatchNorm1d(input_dim), nn.ReLU(),
                                             nn.Softmax(dim=1))
        self.reduce = nn.Sequential(nn.Linear(input_dim, attention_dim), nn.BatchNorm1d(attention_dim), nn.ReLU())

    def forward(self, text_features, image_features, caption_features):
        x = torch.cat((text_features, image_features, caption_features), dim=1)
        return self.reduce(self.attention_layer(x) * x)


class CustomDenseNet161(nn.Module):
    """Reference: :562-585 -- timm ``resnet18`` with ``reset_classifier(0)`` (timm is not installed here; timm's
    ResNet-18 is torchvision's BasicBlock ResNet with the same state-dict keys) + ``fine_tune`` MLP."""

    def __init__(self, layers=(2, 2, 2, 2)):
        super().__init__()
        from torchvision.models.resnet import BasicBlock, ResNet
        self.image_model = ResNet(BasicBlock, list(layers), num_classes=1000)
        self.image_model.fc = nn.Identity()                       # reset_classifier(0)
        self.fine_tune = nn.Sequential(nn.Linear(512, 512), nn.ReLU(inplace=True), nn.Dropout(p=0.35),
                                       nn.Linear(512, 512))

    def forward(self, x):
        return self.fine_tune(self.image_model(x))


class MultimodalClassifierHEAD(nn.Module):
    """Reference: :587-685 with ``fusion_method="concatenation"`` (same attribute names / state-dict keys)."""

    def __init__(self, text_cfg: TowerConfig, caption_cfg: TowerConfig, resnet_layers=(2, 2, 2, 2)):
        super().__init__()
        self.text_model = LLMWithClassificationHead(build_text(text_cfg))
        self.text_dropout = nn.Dropout(0.3)
        self.text_fc = nn.Sequential(nn.Linear(text_cfg.dim, 512), nn.BatchNorm1d(512), nn.ReLU())
        self.caption_text_model = LLMWithClassificationHead(build_text(caption_cfg))
        self.caption_text_dropout = nn.Dropout(0.3)
        self.caption_text_fc = nn.Sequential(nn.Linear(caption_cfg.dim, 512), nn.BatchNorm1d(512), nn.ReLU())
        self.image_model = CustomDenseNet161(resnet_layers)
        self.fusion_layer = ConcatAttention3(3 * 512, 512)
        self.output_fc = nn.Sequential(nn.Linear(512, 1), nn.BatchNorm1d(1))

    def forward(self, text, image, mask, caption_text, caption_text_mask):
        t = self.text_fc(self.text_dropout(self.text_model(text, attention_mask=mask)))
        c = self.caption_text_fc(self.caption_text_dropout(
            self.caption_text_model(caption_text, attention_mask=caption_text_mask)))
        i = self.image_model(image)
        out = self.output_fc(self.fusion_layer(t, i, c))
        return out.squeeze(1)
