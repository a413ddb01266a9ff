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

PARITY PINNING: the reference holds no golden activations for this path, so forward /
backward numerics are pinned by this oracle only ("parity unpinned" by reference artefacts); what the
reference artefacts do pin -- TSV schema, combine_preds known answers -- is covered in
oracle/ensemble.py and tests/golden/.
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

    @staticmethod
    def tiny(**kw) -> "TowerConfig":
        """A small instance of the same graph for fast CPU/GPU parity tests."""
        base = dict(vocab_size=1024, max_position_embeddings=128, dim=128, n_layers=2, n_heads=2,
                    hidden_dim=256, resnet_layers=(1, 1, 1, 1), image_size=64)
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
        self.bert = build_distilbert(cfg)                      # .txt:158
        self.bert_drop = nn.Dropout(cfg.head_dropout)          # .txt:160
        self.bert_fc = nn.Linear(cfg.dim, 512)                 # .txt:161
        self.resnet = build_resnet(cfg)                        # .txt:164
        self.resnet_fc = nn.Linear(1000, 512)                  # .txt:165
        self.fusion_fc = nn.Linear(1024, 512)                  # .txt:168
        self.output_fc = nn.Linear(512, num_classes)           # .txt:170

    def forward(self, text, image, mask):
        bert_output = self.bert(text, attention_mask=mask, return_dict=False)   # .txt:175
        bert_output = self.bert_drop(bert_output[0][:, -1, :])                  # .txt:178 (LAST position)
        bert_output = self.bert_fc(bert_output)                                 # .txt:179
        resnet_output = self.resnet(image)                                      # .txt:183
        resnet_output = self.resnet_fc(resnet_output)                           # .txt:184
        features = torch.cat((bert_output, resnet_output), dim=1)               # .txt:190
        features = self.fusion_fc(features)                                     # .txt:193
        return self.output_fc(features)                                         # .txt:195


def zero_dropout(model: nn.Module) -> nn.Module:
    """Parity runs use p = 0 (bit-parity with torch's RNG stream is not a goal)."""
    for m in model.modules():
        if isinstance(m, nn.Dropout):
            m.p = 0.0
    if hasattr(model, "bert"):
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
    ids = ids * mask + PAD_ID * (1 - mask)
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
